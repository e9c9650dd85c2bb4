def tree_map(fn, tree, *rest):
  import dataclasses
  if isinstance(tree, dict):
    return {k: tree_map(fn, v, *[r[k] for r in rest]) for k, v in tree.items()}
  if isinstance(tree, (list, tuple)):
    return type(tree)(tree_map(fn, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
  if dataclasses.is_dataclass(tree):
    kw = {}
    for f in dataclasses.fields(tree):
      v = getattr(tree, f.name)
      kw[f.name] = None if v is None else tree_map(fn, v, *[getattr(r, f.name) for r in rest])
    return type(tree)(**kw)
  if tree is None:
    return None
  return fn(tree, *rest)


def tree_reduce(fn, tree, initializer=None):
  acc = initializer
  leaves = []

  def collect(x):
    leaves.append(x)
    return x
  tree_map(collect, tree)
  for l in leaves:
    acc = fn(acc, l)
  return acc
