"""Furthest point sampling pinned to the REFERENCE KERNEL ITSELF: oracle/_ref/libfps_ref.so is the reference's
third_party/pointnet2/_ext_src/src/sampling_gpu.cu (:72-176 kernel, :178-215 launcher) compiled for sm_100a from the
reference tree by oracle/build_ref.py.  us3d_fps (cluster / DSMEM kernel, csrc/decoder_ops.cu) must return the same indices,
and so must the numpy restatement (oracle/ops_cpu.furthest_point_sampling) the other tests use.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "oracle", "_ref", "libfps_ref.so")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libfps_ref.so not built (oracle/build_ref.py)")]


def reference_fps(points: torch.Tensor, m: int) -> torch.Tensor:
    """What pointnet2._ext.furthest_point_sampling does (sampling.cpp:71-90): temp = 1e10, output int32 zeros."""
    lib = ctypes.CDLL(LIB)
    lib.fps_ref.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.fps_ref.restype = ctypes.c_int
    b, n, _ = points.shape
    temp = torch.full((b, n), 1e10, dtype=torch.float32, device=points.device)
    out = torch.zeros(b, m, dtype=torch.int32, device=points.device)
    rc = lib.fps_ref(b, n, m, points.contiguous().data_ptr(), temp.data_ptr(), out.data_ptr())
    torch.cuda.synchronize()
    assert rc == 0, rc
    return out


@pytest.mark.parametrize("n,m,span", [(1, 1, 3), (31, 8, 2), (300, 20, 5), (513, 100, 6), (4097, 100, 12), (20000, 100, 40),
                                       (200000, 100, 150), (340000, 100, 200)])
def test_us3d_fps_equals_the_reference_kernel(n, m, span):
    import unscene3d_b200  # noqa: F401
    from oracle import ops_cpu
    from unscene3d_b200.engine import functional as Fn

    rng = np.random.default_rng(7 * n + m)
    pts = rng.integers(-span, span + 1, size=(2, n, 3)).astype(np.float32)  # integer voxel coordinates: exact distance ties
    m = min(m, n)
    x = torch.from_numpy(pts).cuda()
    want = reference_fps(x, m).cpu().numpy()
    got = Fn.furthest_point_sampling(x, m).cpu().numpy()
    assert np.array_equal(got, want)
    if n <= 20000:
        for b in range(2):
            assert np.array_equal(ops_cpu.furthest_point_sampling(pts[b], m), want[b])


def test_us3d_fps_equals_the_reference_kernel_on_real_valued_points():
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200.engine import functional as Fn

    g = torch.Generator().manual_seed(3)
    x = (torch.rand(3, 50000, 3, generator=g) * 8 - 4).cuda()
    assert torch.equal(Fn.furthest_point_sampling(x, 64).cpu(), reference_fps(x, 64).cpu())
