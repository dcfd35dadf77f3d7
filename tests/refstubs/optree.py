"""Stand-in for the two optree calls the reference's Python layer makes
(tree_flatten_with_path / tree_unflatten) with optree's default semantics for the node types
that layer builds: dict children in sorted-key order, namedtuple / tuple / list children in
positional order, everything else a leaf."""


class PyTreeSpec:
    def __init__(self, kind, meta, children):
        self.kind, self.meta, self.children = kind, meta, children

    @property
    def num_leaves(self):
        if self.kind == "leaf":
            return 1
        return sum(c.num_leaves for c in self.children)


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(type(x), "_fields")


def _flatten(tree, path, paths, leaves):
    if isinstance(tree, dict):
        keys = sorted(tree)
        return PyTreeSpec("dict", (type(tree), keys),
                          [_flatten(tree[k], path + (k,), paths, leaves) for k in keys])
    if _is_namedtuple(tree):
        return PyTreeSpec("namedtuple", type(tree),
                          [_flatten(v, path + (i,), paths, leaves) for i, v in enumerate(tree)])
    if isinstance(tree, (tuple, list)):
        return PyTreeSpec("seq", type(tree),
                          [_flatten(v, path + (i,), paths, leaves) for i, v in enumerate(tree)])
    paths.append(path)
    leaves.append(tree)
    return PyTreeSpec("leaf", None, [])


def tree_flatten_with_path(tree):
    paths, leaves = [], []
    spec = _flatten(tree, (), paths, leaves)
    return paths, leaves, spec


def _unflatten(spec, it):
    if spec.kind == "leaf":
        return next(it)
    kids = [_unflatten(c, it) for c in spec.children]
    if spec.kind == "dict":
        cls, keys = spec.meta
        return cls(zip(keys, kids))
    if spec.kind == "namedtuple":
        return spec.meta(*kids)
    return spec.meta(kids)


def tree_unflatten(treespec, leaves):
    return _unflatten(treespec, iter(leaves))
