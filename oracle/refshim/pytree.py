"""ORACLE (test infrastructure) - the small part of jax.tree_util / equinox's pytree behaviour the reference's loader
and `eqx.tree_at` calls rely on (utils.py:189-218, deeplabv3.py:209, googlenet.py:327, experimental.py:71-80).

Nodes: Module instances (children = dataclass-style fields in annotation order, base classes first), lists, tuples,
dicts; `None` is an empty node; everything else (arrays, ints, bools, strings, callables, StateIndex) is a leaf.
"""
from __future__ import annotations

from typing import Any, Callable, List


def field_names(obj) -> List[str]:
    names: List[str] = []
    for base in reversed(type(obj).__mro__):
        for name in base.__dict__.get("__annotations__", {}):
            if name not in names:
                names.append(name)
    return [n for n in names if n in obj.__dict__]


def _is_module(x) -> bool:
    from .fake_equinox import Module

    return isinstance(x, Module)


def _children(x):
    """(kind, keys, values) of a node, or None for a leaf"""
    if _is_module(x):
        names = field_names(x)
        return "module", names, [x.__dict__[n] for n in names]
    if isinstance(x, (list, tuple)):
        return type(x).__name__, list(range(len(x))), list(x)
    if isinstance(x, dict):
        keys = sorted(x)
        return "dict", keys, [x[k] for k in keys]
    return None


def _rebuild(x, keys, values):
    if _is_module(x):
        new = object.__new__(type(x))
        new.__dict__.update(x.__dict__)
        for k, v in zip(keys, values):
            new.__dict__[k] = v
        return new
    if isinstance(x, list):
        return list(values)
    if isinstance(x, tuple):
        return tuple(values)
    return dict(zip(keys, values))


class TreeDef:
    def __init__(self, skeleton, is_leaf):
        self.skeleton, self.is_leaf = skeleton, is_leaf


def tree_flatten(tree, is_leaf: Callable[[Any], bool] = None):
    leaves: List[Any] = []

    def rec(x):
        if x is None:
            return
        if is_leaf is not None and is_leaf(x):
            leaves.append(x)
            return
        node = _children(x)
        if node is None:
            leaves.append(x)
            return
        for v in node[2]:
            rec(v)

    rec(tree)
    return leaves, TreeDef(tree, is_leaf)


def tree_leaves(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[0]


def tree_unflatten(treedef: TreeDef, leaves):
    it = iter(leaves)

    def rec(x):
        if x is None:
            return None
        if treedef.is_leaf is not None and treedef.is_leaf(x):
            return next(it)
        node = _children(x)
        if node is None:
            return next(it)
        return _rebuild(x, node[1], [rec(v) for v in node[2]])

    return rec(treedef.skeleton)


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, treedef = tree_flatten(tree, is_leaf)
    others = [tree_flatten(r, is_leaf)[0] if _children(tree) is not None else [r] for r in rest]
    if _children(tree) is None and tree is not None and not (is_leaf and is_leaf(tree)):
        return f(tree, *rest)            # a leaf at the root (utils.py:217 maps over an iterator object)
    return tree_unflatten(treedef, [f(*xs) for xs in zip(leaves, *others)])


class _Slot:
    """unique stand-in for a leaf while `where` is evaluated (equinox.tree_at does the same)"""


def tree_at(where, pytree, replace=None, replace_fn=None):
    def copy(x):
        if x is None:
            return None
        node = _children(x)
        if node is None:
            return _Slot()
        return _rebuild(x, node[1], [copy(v) for v in node[2]])

    shadow = copy(pytree)
    picked = where(shadow)
    single = not isinstance(picked, (list, tuple))
    picked = [picked] if single else list(picked)
    if replace_fn is None:
        repl = [replace] if single else list(replace)
    else:
        repl = None

    def rec(orig, sh):
        for i, p in enumerate(picked):
            if sh is p:
                return replace_fn(orig) if repl is None else repl[i]
        if orig is None:
            return None
        node = _children(orig)
        if node is None:
            return orig
        sh_vals = _children(sh)[2]
        return _rebuild(orig, node[1], [rec(o, s) for o, s in zip(node[2], sh_vals)])

    return rec(pytree, shadow)
