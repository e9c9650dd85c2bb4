def stop_gradient(x):
  return x


class Precision:
  HIGHEST = 'highest'
