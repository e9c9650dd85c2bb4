from . import scope  # noqa: F401


class FrozenDict(dict):
  def __hash__(self):
    return hash(tuple(sorted(self.items())))
