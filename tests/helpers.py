"""Shared test plumbing: load a model package against a chosen MinkowskiEngine implementation.

The model files (ours under unscene3d_b200/models, the reference's under /root/reference/models)
bind `MinkowskiEngine` at import time, so the same source can be imported twice under different
package names — once over the CPU oracle, once over the CUDA shim — inside one process.
"""
import collections
import collections.abc
import contextlib
import importlib
import importlib.util
import os
import sys
import types
import zlib

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
if REPO not in sys.path:
    sys.path.insert(0, REPO)

_ME_KEYS = ["MinkowskiEngine", "MinkowskiEngine.MinkowskiOps", "MinkowskiEngine.MinkowskiPooling", "MinkowskiEngine.utils"]


@contextlib.contextmanager
def minkowski_as(modules: dict):
    """Temporarily make `import MinkowskiEngine` resolve to `modules` (name -> module object)."""
    keys = list(dict.fromkeys(_ME_KEYS + list(modules)))
    saved = {k: sys.modules.get(k) for k in keys}
    sys.modules.update(modules)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def oracle_me_modules():
    """Every operator module the model files import, backed by the CPU oracle."""
    from oracle import me_cpu, ops_cpu

    mods = me_cpu.as_module_tree()
    mods.update(ops_cpu.as_module_tree())
    return mods


def _stub_modules():
    """Import-only dependencies of the reference's model files that are absent here."""
    import types

    hydra = types.ModuleType("hydra")
    hydra.main = lambda *a, **k: (lambda fn: fn)
    return {"hydra": hydra}


def load_package(pkg_dir: str, alias: str, me_modules: dict):
    """Import the package at `pkg_dir` under the name `alias` with MinkowskiEngine := me_modules."""
    if alias in sys.modules:
        return sys.modules[alias]
    with minkowski_as(me_modules):
        spec = importlib.util.spec_from_file_location(alias, os.path.join(pkg_dir, "__init__.py"),
                                                      submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[alias] = mod
        try:
            spec.loader.exec_module(mod)
        except Exception:
            for k in [k for k in sys.modules if k == alias or k.startswith(alias + ".")]:
                del sys.modules[k]
            raise
    return mod


def our_models_on_oracle():
    from oracle import ops_cpu
    import unscene3d_b200  # noqa: F401  (shims on sys.path: SetCriterion binds `custom_cuda_utils` when it is instantiated)

    mod = load_package(os.path.join(REPO, "unscene3d_b200", "models"), "oracle_backed_models", oracle_me_modules())
    mod.mask3d.CrossAttentionLayer.attention_core = staticmethod(ops_cpu.multihead_cross_attention)
    mod.criterion.SetCriterion.mask_loss_core = staticmethod(ops_cpu.mask_losses)
    mod.mask3d.Mask3D.segment_attention_core = staticmethod(ops_cpu.segment_attention_masks)
    mod.position_embedding.PositionEmbeddingCoordsSine.fourier_core = staticmethod(ops_cpu.fourier_posenc)
    mod.modules.resnet_block._ResidualBase.block_core = staticmethod(lambda block, x: None)  # the reference's own sequence
    mod.res16unet.Res16UNetBase.transition_core = staticmethod(lambda conv_layer, norm, x: None)
    mod.res16unet.Res16UNetBase.stage_core = staticmethod(lambda blocks, x: None)
    return mod


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "models"))


def staged_reference_root():
    """Where the UNMODIFIED reference Python files can be imported from: /root/reference in the build container, the copy
    staged by oracle/build_ref.py (git-ignored oracle/_ref/reference, shipped with the snapshot) on the GPU box."""
    for root in (REFERENCE, os.path.join(REPO, "oracle", "_ref", "reference")):
        if os.path.isdir(os.path.join(root, "models")):
            return root
    return None


def reference_models_on_shim():
    """The UNMODIFIED reference `models` package over the CUDA drop-in modules (unscene3d_b200/shims): `import MinkowskiEngine`,
    `torch_scatter`, `pointnet2._ext`, `detectron2`, `custom_cuda_utils` inside the reference files resolve to libus3d."""
    root = staged_reference_root()
    assert root is not None
    collections.Set = collections.abc.Set
    cur = sys.modules.get("models")
    if cur is not None and getattr(cur, "_us3d_backend", None) == "shim":
        return cur
    import unscene3d_b200  # noqa: F401  (shims first on sys.path)

    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k.startswith("third_party")]:
        del sys.modules[k]
    for k in _ME_KEYS:  # a CPU test of the same process may have left the oracle's modules under these names
        m = sys.modules.get(k)
        if m is not None and "shims" not in (getattr(m, "__file__", "") or ""):
            del sys.modules[k]
    saved = {k: sys.modules.get(k) for k in _stub_modules()}
    sys.modules.update({k: v for k, v in _stub_modules().items() if k not in sys.modules})
    sys.path.insert(0, root)
    try:
        mod = importlib.import_module("models")
        for sub in ("res16unet", "mask3d", "matcher", "criterion"):
            importlib.import_module("models." + sub)
    finally:
        sys.path.remove(root)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
    mod._us3d_backend = "shim"
    import MinkowskiEngine

    assert "shims" in MinkowskiEngine.__file__, MinkowskiEngine.__file__
    return mod


def reference_models_on_oracle():
    """The UNMODIFIED reference `models` package over the oracle (build container only).
    The reference uses absolute imports (`from models.resnet import ...`), so it has to be imported
    under its own name with /root/reference on sys.path."""
    assert have_reference()
    collections.Set = collections.abc.Set  # py3.12 skew, utils/utils.py:340
    if "models" in sys.modules and getattr(sys.modules["models"], "__file__", "").startswith(REFERENCE):
        return sys.modules["models"]
    import unscene3d_b200  # noqa: F401  puts the pure-Python detectron2 / custom_cuda_utils shims on sys.path

    mods = oracle_me_modules()
    mods.update(_stub_modules())
    with minkowski_as(mods):
        sys.path.insert(0, REFERENCE)
        try:
            for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k.startswith("third_party")]:
                del sys.modules[k]
            mod = importlib.import_module("models")
            importlib.import_module("models.res16unet")
            importlib.import_module("models.mask3d")
            importlib.import_module("models.matcher")
            importlib.import_module("models.criterion")
        finally:
            sys.path.remove(REFERENCE)
    return mod


class Cfg:
    """Stand-in for the hydra node the backbones read (conf/model/mask3d.yaml:36-47)."""

    def __init__(self, **kw):
        self.bn_momentum = 0.02
        self.conv1_kernel_size = 3
        self.dilations = [1, 1, 1, 1]
        self.__dict__.update(kw)


def deterministic_state(module: torch.nn.Module, seed: int = 0):
    """Name-keyed, construction-order-independent weights (same recipe as
    unscene3d_b200.utils.seeded_init, restated here so tests do not depend on it)."""
    out = {}
    for name, t in module.state_dict().items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros_like(t)
        elif name.endswith("running_mean"):
            out[name] = (torch.rand(t.shape, generator=g) - 0.5) * 0.2
        elif name.endswith("running_var"):
            out[name] = 0.5 + torch.rand(t.shape, generator=g)
        elif ".bn.weight" in name or name.endswith("norm.weight"):
            out[name] = 0.5 + torch.rand(t.shape, generator=g)
        elif ".bn.bias" in name or name.endswith("norm.bias"):
            out[name] = (torch.rand(t.shape, generator=g) - 0.5) * 0.4
        else:
            fan = t.shape[-2] * (t.shape[0] if t.ndim == 3 else 1) if t.ndim >= 2 else max(t.numel(), 1)
            bound = (3.0 / fan) ** 0.5 * 1.4
            out[name] = (torch.rand(t.shape, generator=g) * 2 - 1) * bound
    return out


def random_scene(n_target: int, seed: int, batch: int = 1, extent: int = 24, negative: bool = True):
    """Small random voxel cloud on a few noisy sheets (surface-like, with negative coordinates)."""
    rng = np.random.default_rng(seed)
    coords = []
    for b in range(batch):
        pts = []
        while sum(len(p) for p in pts) < n_target * 2:
            u = rng.integers(0, extent, size=(n_target, 2))
            axis = rng.integers(0, 3)
            h = rng.integers(0, extent) + (rng.random(n_target) < 0.3).astype(np.int64)
            p = np.insert(u, axis, h, axis=1)
            pts.append(p)
        p = np.concatenate(pts)
        if negative:
            p = p - extent // 2
        _, first = np.unique(p, axis=0, return_index=True)
        p = p[np.sort(first)][:n_target]
        coords.append(np.concatenate([np.full((p.shape[0], 1), b), p], 1))
    return np.concatenate(coords).astype(np.int32)


# ---- ReLU-mask replay -----------------------------------------------------------------------------------------------
# A parameter gradient is a sum over ~10^6 elements gated by ReLU masks.  Two implementations whose forward passes differ by
# 1e-6 .. 4e-5 (summation order, the fp32-faithful bf16 split) disagree on the sign of a handful of near-zero
# pre-activations; each flipped mask entry is a 100 % error on one element, so free-running gradients agree only to
# ~sqrt(#flips / #elements) (1e-2) however exact the backward kernels are.  To pin the backward path at the north-star
# tolerance the discrete decisions are taken out: the CUDA run records the mask of every ReLU it executes, in call order, and
# the oracle replays them (relu(x) := x * mask).  Everything continuous is then compared at 1e-3.  (tests/test_mask3d.py
# does the same with the decoder's boolean attention masks.)
@contextlib.contextmanager
def record_relu_masks(engine):
    masks = []
    orig = engine.MinkowskiReLU.forward

    def forward(self, x):
        out = orig(self, x)
        masks.append((out.F.detach() > 0).cpu())
        return out

    engine.MinkowskiReLU.forward = forward
    try:
        yield masks
    finally:
        engine.MinkowskiReLU.forward = orig


@contextlib.contextmanager
def replay_relu_masks(me_cpu, masks, stats=None):
    """`stats`, if given, receives (#entries where the oracle's own mask differs, #entries) per ReLU."""
    it = iter(masks)
    orig = me_cpu.MinkowskiReLU.forward

    def forward(self, x):
        m = next(it)
        assert tuple(m.shape) == tuple(x.F.shape), f"ReLU call order differs: mask {tuple(m.shape)} vs features {tuple(x.F.shape)}"
        if stats is not None:
            stats.append((int(((x.F.detach() > 0) != m).sum()), m.numel()))
        return x._like(x.F * m.to(x.F.dtype))

    me_cpu.MinkowskiReLU.forward = forward
    try:
        yield
    finally:
        me_cpu.MinkowskiReLU.forward = orig
