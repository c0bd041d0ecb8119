"""``flax.linen`` stand-in: Module with dataclass fields, ``setup`` / ``@compact`` submodules,
``self.param``, ``init`` and ``apply`` over a nested ``{'params': {...}}`` dict — the subset the
reference's layers, blocks and models use (test infrastructure, see ../../README.md)."""
import dataclasses
import functools
from typing import Any, Optional

import numpy as _np

from jax import _core as _jc
from jax import random as _random

from . import initializers  # noqa: F401

_UNSET = object()
_stack = []          # modules whose method is executing (innermost last)


class ScopeParamNotFoundError(KeyError):
    pass


class ScopeParamShapeError(ValueError):
    pass


def compact(fn):
    fn._compact = True
    return fn


def _wrap_method(fn):
    @functools.wraps(fn)
    def run(self, *a, **k):
        self._ensure_setup()
        _stack.append(self)
        try:
            return fn(self, *a, **k)
        finally:
            _stack.pop()
    run._wrapped = True
    return run


class Module:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        ann = dict(cls.__dict__.get("__annotations__", {}))
        ann.pop("parent", None), ann.pop("name", None)
        ann["parent"] = Any
        ann["name"] = Optional[str]
        cls.__annotations__ = ann
        cls.parent = dataclasses.field(default=_UNSET, kw_only=True, repr=False, compare=False)
        cls.name = dataclasses.field(default=None, kw_only=True)
        dataclasses.dataclass(cls, eq=False, repr=False)
        for k, v in list(cls.__dict__.items()):
            if k in ("setup", "param", "init", "apply", "clone", "variables") or not callable(v) or isinstance(v, type):
                continue
            if (k == "__call__" or not k.startswith("_")) and not getattr(v, "_wrapped", False):
                setattr(cls, k, _wrap_method(v))

    # ------------------------------------------------------------------ construction
    def __post_init__(self):
        object.__setattr__(self, "_state", dict(root=None, mode=None, rng=None, setup_done=False, in_setup=False,
                                                 auto={}, taken=set()))
        if self.parent is _UNSET:
            object.__setattr__(self, "parent", _stack[-1] if _stack else None)
        p = self.parent
        if p is not None:
            if self.name is None and not p._state["in_setup"]:
                i = p._state["auto"].get(type(self).__name__, 0)
                p._state["auto"][type(self).__name__] = i + 1
                object.__setattr__(self, "name", f"{type(self).__name__}_{i}")
            if self.name is not None:
                p._claim(self.name)

    def _claim(self, name):
        if name in self._state["taken"]:
            raise ValueError(f"Could not create submodule '{name}' in {type(self).__name__}: name already in use")
        self._state["taken"].add(name)

    def __setattr__(self, k, v):
        st = self.__dict__.get("_state")
        if st is not None and st["in_setup"] and isinstance(v, Module) and v.parent is self and v.name is None:
            object.__setattr__(v, "name", k)
            self._claim(k)
        object.__setattr__(self, k, v)

    def setup(self):
        pass

    def _ensure_setup(self):
        st = self._state
        if st["setup_done"]:
            return
        if self._root()._state["mode"] is None:
            raise RuntimeError(f"Can't call methods of the unbound module {type(self).__name__}: use .init / .apply")
        st["setup_done"], st["in_setup"] = True, True
        _stack.append(self)
        try:
            self.setup()
        finally:
            _stack.pop()
            st["in_setup"] = False

    # ------------------------------------------------------------------ scopes
    def _root(self):
        m = self
        while m.parent is not None:
            m = m.parent
        return m

    def _scope(self, create):
        if self.parent is None:
            return self._state["root"]
        ps = self.parent._scope(create)
        if self.name is None:
            raise RuntimeError(f"submodule {type(self).__name__} has no name")
        if create:
            return ps.setdefault(self.name, {})
        return ps.get(self.name, {}) if isinstance(ps, dict) else {}

    def _path(self):
        return "/".join(([] if self.parent is None else [self.parent._path()]) + [self.name or type(self).__name__])

    def param(self, name, init_fn, *init_args, **init_kw):
        root = self._root()._state
        init = root["mode"] == "init"
        scope = self._scope(init)
        if name in scope:
            v = _jc.asarray(scope[name])
            if init_args and isinstance(init_args[0], (tuple, list)) and tuple(v.shape) != tuple(init_args[0]):
                raise ScopeParamShapeError(f"initializer expected shape {tuple(init_args[0])} for parameter "
                                           f"'{name}' in {self._path()}, got {tuple(v.shape)}")
            return v
        if not init:
            raise ScopeParamNotFoundError(f"could not find parameter named '{name}' in scope '{self._path()}'")
        root["rng"], k = _random.split(root["rng"])
        v = _jc.asarray(init_fn(k, *init_args, **init_kw))
        scope[name] = v
        return v

    # ------------------------------------------------------------------ entry points
    def clone(self, **updates):
        kw = {f.name: getattr(self, f.name) for f in dataclasses.fields(self) if f.init}
        kw.update(updates)
        return type(self)(**kw)

    def _bound(self, mode, tree, rng=None):
        m = self.clone(parent=None)
        m._state.update(root=tree, mode=mode, rng=rng)
        return m

    def apply(self, variables, *args, rngs=None, method=None, mutable=False, **kwargs):
        tree = variables["params"] if isinstance(variables, dict) and "params" in variables else {}
        m = self._bound("apply", tree)
        fn = getattr(m, method.__name__ if callable(method) else (method or "__call__"))
        return fn(*args, **kwargs)

    def init(self, rngs, *args, method=None, **kwargs):
        rng = rngs["params"] if isinstance(rngs, dict) else rngs
        tree = {}
        m = self._bound("init", tree, rng)
        getattr(m, method or "__call__")(*args, **kwargs)
        return {"params": tree} if tree else {}

    def init_with_output(self, rngs, *args, **kwargs):
        rng = rngs["params"] if isinstance(rngs, dict) else rngs
        tree = {}
        out = self._bound("init", tree, rng)(*args, **kwargs)
        return out, ({"params": tree} if tree else {})

    def __repr__(self):
        fs = ", ".join(f"{f.name}={getattr(self, f.name)!r}" for f in dataclasses.fields(self)
                       if f.name not in ("parent",))
        return f"{type(self).__name__}({fs})"
