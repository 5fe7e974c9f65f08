"""Minimal `omegaconf` (2.0 surface used by the reference: DictConfig / ListConfig with attribute and item access, `${a.b.c}`
interpolation resolved on access against the root, `${now:FORMAT}`, OmegaConf.create / to_container / to_yaml / load / merge)."""
import copy
import re
import time
from collections.abc import MutableMapping, MutableSequence

import yaml

_INTERP = re.compile(r"\$\{([^${}]+)\}")


class _Node:
    __slots__ = ()

    def _root(self):
        n = self
        while n._parent is not None:
            n = n._parent
        return n

    def _resolve(self, value):
        if not isinstance(value, str) or "${" not in value:
            return value
        root = self._root()
        m = _INTERP.fullmatch(value)
        if m:
            return _lookup(root, m.group(1))
        out = value
        for _ in range(16):
            new = _INTERP.sub(lambda mm: str(_lookup(root, mm.group(1))), out)
            if new == out:
                break
            out = new
        return out


def _lookup(root, expr):
    expr = expr.strip()
    if expr.startswith("now:"):
        return time.strftime(expr[4:])
    if expr.startswith("env:"):
        import os

        key, _, default = expr[4:].partition(",")
        return os.environ.get(key, default or None)
    node = root
    for part in expr.split("."):
        node = node[int(part)] if isinstance(node, ListConfig) else node[part]
    return node


def _wrap(value, parent):
    if isinstance(value, (DictConfig, ListConfig)):
        value = _unwrap(value, resolve=False)
    if isinstance(value, dict):
        return DictConfig(value, parent)
    if isinstance(value, (list, tuple)):
        return ListConfig(list(value), parent)
    return value


def _unwrap(value, resolve):
    if isinstance(value, DictConfig):
        return {k: _unwrap(value._get(k, resolve), resolve) for k in value._content}
    if isinstance(value, ListConfig):
        return [_unwrap(value._get(i, resolve), resolve) for i in range(len(value._content))]
    return value


class DictConfig(_Node, MutableMapping):
    __slots__ = ("_content", "_parent")

    def __init__(self, content=None, parent=None):
        object.__setattr__(self, "_parent", parent)
        object.__setattr__(self, "_content", {})
        for k, v in (content or {}).items():
            self._content[k] = _wrap(v, self)

    def _get(self, key, resolve=True):
        v = self._content[key]
        return self._resolve(v) if resolve else v

    def __getitem__(self, key):
        return self._get(key)

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        try:
            return self._get(key)
        except KeyError:
            raise AttributeError(f"Missing key {key}") from None

    def __setitem__(self, key, value):
        self._content[key] = _wrap(value, self)

    __setattr__ = __setitem__

    def __delitem__(self, key):
        del self._content[key]

    def __iter__(self):
        return iter(self._content)

    def __len__(self):
        return len(self._content)

    def __contains__(self, key):
        return key in self._content

    def get(self, key, default=None):
        return self._get(key) if key in self._content else default

    def keys(self):
        return self._content.keys()

    def __repr__(self):
        return repr(_unwrap(self, resolve=False))

    def __deepcopy__(self, memo):
        return DictConfig(copy.deepcopy(_unwrap(self, resolve=False), memo), None) if self._parent is None else _unwrap(self, True)


class ListConfig(_Node, MutableSequence):
    __slots__ = ("_content", "_parent")

    def __init__(self, content=None, parent=None):
        object.__setattr__(self, "_parent", parent)
        object.__setattr__(self, "_content", [])
        for v in content or []:
            self._content.append(_wrap(v, self))

    def _get(self, i, resolve=True):
        v = self._content[i]
        return self._resolve(v) if resolve else v

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._get(j) for j in range(*i.indices(len(self._content)))]
        return self._get(i)

    def __setitem__(self, i, value):
        self._content[i] = _wrap(value, self)

    def __delitem__(self, i):
        del self._content[i]

    def __len__(self):
        return len(self._content)

    def insert(self, i, value):
        self._content.insert(i, _wrap(value, self))

    def __iter__(self):
        return (self._get(i) for i in range(len(self._content)))

    def __repr__(self):
        return repr(_unwrap(self, resolve=False))

    def __eq__(self, other):
        return list(self) == list(other)


def _merge(dst: dict, src: dict):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


class OmegaConf:
    @staticmethod
    def create(obj=None):
        if isinstance(obj, str):
            obj = yaml.safe_load(obj)
        if isinstance(obj, (DictConfig, ListConfig)):
            obj = _unwrap(obj, resolve=False)
        return _wrap(obj if obj is not None else {}, None)

    @staticmethod
    def load(path):
        with open(path) as fh:
            return OmegaConf.create(yaml.safe_load(fh) or {})

    @staticmethod
    def to_container(cfg, resolve=False, **_):
        return _unwrap(cfg, resolve)

    @staticmethod
    def to_yaml(cfg, resolve=False, **_):
        return yaml.safe_dump(_unwrap(cfg, resolve), sort_keys=False)

    @staticmethod
    def merge(*cfgs):
        out = {}
        for c in cfgs:
            _merge(out, _unwrap(c, resolve=False) if isinstance(c, DictConfig) else dict(c))
        return OmegaConf.create(out)

    @staticmethod
    def set_struct(cfg, value):
        return None

    @staticmethod
    def is_config(obj):
        return isinstance(obj, (DictConfig, ListConfig))

    @staticmethod
    def select(cfg, key, default=None):
        try:
            return _lookup(cfg, key)
        except (KeyError, IndexError):
            return default


def open_dict(cfg):
    import contextlib

    return contextlib.nullcontext(cfg)
