def external_configurable(fn, module=None, name=None):
  return fn
