import importlib
import os

_ORIGINAL_CWD = [None]


def get_original_cwd():
    return _ORIGINAL_CWD[0] or os.getcwd()


def to_absolute_path(path):
    return path if os.path.isabs(path) else os.path.join(get_original_cwd(), path)


def get_class(path):
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


get_method = get_class


class _AttrDict(dict):
    """Container handed to the callee for a nested config node: item and attribute access, may hold instantiated objects."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    __setattr__ = dict.__setitem__


def _materialise(value):
    """Nested nodes: a node with `_target_` is instantiated (the reference relies on it: models/mask3d.py:56 takes
    `config.backbone` as the ready backbone module, conf/model/mask3d.yaml:36-47), any other mapping becomes an attribute-access
    container, lists are walked, leaves (already interpolated by the config) pass through."""
    from omegaconf import DictConfig, ListConfig

    if isinstance(value, (DictConfig, dict)):
        if "_target_" in value:
            return instantiate(value)
        return _AttrDict({k: _materialise(value[k]) for k in value.keys()})
    if isinstance(value, (ListConfig, list, tuple)):
        return [_materialise(v) for v in value]
    return value


def instantiate(config, *args, **kwargs):
    """`_target_(*args, **{config items}, **kwargs)` with interpolated values; keyword arguments override config items."""
    if config is None:
        return None
    target = config["_target_"]
    fn = get_class(target) if isinstance(target, str) else target
    params = {k: _materialise(config[k]) for k in config.keys() if k not in ("_target_", "_recursive_", "_convert_")}
    params.update(kwargs)
    return fn(*args, **params)


call = instantiate
