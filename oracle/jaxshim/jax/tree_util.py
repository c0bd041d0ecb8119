"""``jax.tree_util`` stand-in for dict / list / tuple trees (see ../README.md)."""


def tree_leaves(t):
    if isinstance(t, dict):
        return [l for k in sorted(t) for l in tree_leaves(t[k])]
    if isinstance(t, (list, tuple)):
        return [l for v in t for l in tree_leaves(v)]
    return [] if t is None else [t]


def tree_map(f, t, *rest):
    if isinstance(t, dict):
        return {k: tree_map(f, t[k], *[r[k] for r in rest]) for k in t}
    if isinstance(t, (list, tuple)):
        return type(t)(tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(t))
    return None if t is None else f(t, *rest)


def tree_flatten(t):
    return tree_leaves(t), None
