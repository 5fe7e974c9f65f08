"""unscene3d_b200 — B200-native (sm_100a) backend for UnScene3D's self-training hot path.

Importing the package loads libus3d.so (fails loudly if it is not built) and makes the drop-in
operator modules (`MinkowskiEngine`, `torch_scatter`, `pointnet2`, `custom_cuda_utils`, ...) importable
under the names the reference uses.
"""
import os
import sys

from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)

SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def install_shims():
    """Put the drop-in modules first on sys.path."""
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)


install_shims()
