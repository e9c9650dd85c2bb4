class TrainState:
  pass
