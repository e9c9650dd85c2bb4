"""gin stand-in: decorators are the identity; `bind` records values that the flax stand-in's
Module.__init__ applies by class name (what gin's injection would do)."""
from . import config  # noqa: F401


def configurable(*args, **kwargs):
  if len(args) == 1 and callable(args[0]) and not kwargs:
    return args[0]
  return lambda x: x


def add_config_file_search_path(*a, **k):
  pass


def bind(cls_name, **values):
  from flax import linen
  linen._GIN.setdefault(cls_name, {}).update(values)


def clear():
  from flax import linen
  linen._GIN.clear()
