"""Stand-ins for the third-party packages the reference's ENTRY POINT needs and this image lacks: `hydra`
(1.0 semantics: defaults-list composition with `# @package` headers, command-line overrides, `instantiate`),
`omegaconf` (attribute / item access, `${a.b}` and `${now:...}` interpolation), `pytorch_lightning` (LightningModule base and a
plain single-device fit loop), plus import-only stubs (pyviz3d, matplotlib.cm, open3d, plyfile, trimesh, imageio).

They exist so that the UNMODIFIED `main_instance_segmentation.py` + `conf/` + `trainer/trainer.py` can run their training step on
the CUDA shim inside this repo's tests (tests/test_entry_point.py); they are not part of the operator path and a real
installation of any of these packages takes precedence: `install()` only adds what `import` cannot find."""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REAL = ("hydra", "omegaconf", "pytorch_lightning")
IMPORT_ONLY = ("pyviz3d", "matplotlib", "open3d", "plyfile", "trimesh", "imageio", "wandb", "torchmetrics", "natsort", "fire", "albumentations",
               "volumentations")


class _Anything(types.ModuleType):
    """Import-only module: any attribute is a class that can be constructed and called and does nothing."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None,
                              "__getattr__": lambda self, n: (lambda *a, **k: None)})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _Anything(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        pass


def _importable(name):
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def install():
    """Returns the names that were stood in for."""
    import collections
    import collections.abc

    if not hasattr(collections, "Set"):  # the reference targets Python 3.8 (utils/utils.py:340 subclasses collections.Set)
        collections.Set = collections.abc.Set
    used = []
    need_real = [n for n in REAL if not _importable(n)]
    if need_real and HERE not in sys.path:
        sys.path.append(HERE)  # at the END: anything installed wins
    used += need_real
    stubs = [n for n in IMPORT_ONLY if not _importable(n)]
    if stubs and not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder(stubs))
    return used + stubs
