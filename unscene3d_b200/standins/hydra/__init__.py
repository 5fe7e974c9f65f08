"""Minimal `hydra` with 1.0 semantics (the reference pins hydra-core 1.0.5): `@hydra.main(config_path, config_name)` composes the
primary config with its defaults list (`group: option` entries, `# @package _group_ | _global_ | a.b` headers) and the
command-line overrides (`a.b=c`, `group/sub=option`, `+a.b=c`); `hydra.utils.instantiate` builds nested `_target_` nodes before
the call (models/mask3d.py:56 receives `config.backbone` as a module)."""
import functools
import inspect
import os
import re
import sys

import yaml
from omegaconf import DictConfig, OmegaConf

from . import utils  # noqa: F401

__version__ = "1.0.5+us3d-standin"
_PACKAGE = re.compile(r"#\s*@package\s+(\S+)")


def _load_group_file(conf_dir, group, option):
    path = os.path.join(conf_dir, group, option if option.endswith((".yaml", ".yml")) else option + ".yaml")
    with open(path) as fh:
        text = fh.read()
    m = _PACKAGE.search(text.split("\n", 3)[0] + "\n" + "\n".join(text.split("\n")[1:3]))
    package = m.group(1) if m else "_group_"  # hydra 1.0 default when the header is absent: the group
    if package == "_group_":
        package = group.replace("/", ".")
    elif package == "_global_":
        package = ""
    return package, yaml.safe_load(text)


def _place(root: dict, package: str, content):
    if content is None:
        return
    node = root
    parts = [p for p in package.split(".") if p]
    for p in parts[:-1]:
        node = node.setdefault(p, {})
    if not parts:
        _deep_merge(root, content)
    elif isinstance(content, dict) and isinstance(node.get(parts[-1]), dict):
        _deep_merge(node[parts[-1]], content)
    else:
        node[parts[-1]] = content


def _deep_merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_merge(dst[k], v)
        else:
            dst[k] = v


def _set_path(root: dict, dotted: str, value):
    node = root
    parts = dotted.split(".")
    for p in parts[:-1]:
        nxt = node.get(p) if isinstance(node, dict) else node[int(p)]
        if nxt is None:
            nxt = node[p] = {}
        node = nxt
    if isinstance(node, list):
        node[int(parts[-1])] = value
    else:
        node[parts[-1]] = value


def compose(conf_dir: str, config_name: str, overrides=()):
    with open(os.path.join(conf_dir, config_name if config_name.endswith((".yaml", ".yml")) else config_name + ".yaml")) as fh:
        primary = yaml.safe_load(fh) or {}
    defaults = primary.pop("defaults", [])
    choices = []
    for d in defaults:
        if isinstance(d, dict):
            choices += list(d.items())
    values = []
    for ov in overrides:
        key, _, val = ov.partition("=")
        key = key.lstrip("+")
        if os.path.isdir(os.path.join(conf_dir, key)):  # a config group: data/datasets=freemask
            for i, (g, _) in enumerate(choices):
                if g == key:
                    choices[i] = (g, val)
                    break
            else:
                choices.append((key, val))
        else:
            values.append((key, yaml.safe_load(val) if val != "" else None))
    cfg = {}
    _deep_merge(cfg, primary)  # hydra 1.0: the primary config first, the defaults list on top of it
    for group, option in choices:
        if option in (None, "null"):
            continue
        package, content = _load_group_file(conf_dir, group, str(option))
        _place(cfg, package, content)
    for key, val in values:
        _set_path(cfg, key, val)
    cfg.pop("hydra", None)
    return OmegaConf.create(cfg)


def main(config_path=None, config_name=None, **_):
    def decorator(fn):
        @functools.wraps(fn)
        def wrapper(cfg_passthrough=None):
            if cfg_passthrough is not None:  # a decorated function called with a config runs as it is (as in hydra)
                return fn(cfg_passthrough)
            conf_dir = os.path.join(os.path.dirname(os.path.abspath(inspect.getsourcefile(fn))), config_path or "")
            utils._ORIGINAL_CWD[0] = os.getcwd()
            cfg = compose(conf_dir, config_name, [a for a in sys.argv[1:] if "=" in a])
            return fn(cfg)

        return wrapper

    return decorator
