"""No-GPU checks of the drop-in boundary: libus3d.so loads, exports every function include/us3d.h
declares (and nothing the header does not), and argument validation fails loudly without touching a
device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "us3d.h")
DEBUG_HEADER = os.path.join(ROOT, "include", "us3d_debug.h")


def header_functions(path=HEADER):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(us3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from unscene3d_b200 import _lib

    declared = header_functions()
    assert len(declared) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/us3d.h but not exported by libus3d.so"
    assert sorted(_lib.PROTOTYPES) == declared, "ctypes prototypes and header disagree"


def test_library_exports_nothing_that_is_not_declared():
    """The converse: every us3d_* symbol the shared library exports is declared — in include/us3d.h (the drop-in surface) or in
    include/us3d_debug.h (profiling / tuning hooks, which the product's operator path does not call)."""
    import subprocess

    from unscene3d_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted({line.split()[-1] for line in out.splitlines() if line.split() and line.split()[-1].startswith("us3d_")})
    declared = set(header_functions()) | set(header_functions(DEBUG_HEADER))
    assert not [n for n in exported if n not in declared], [n for n in exported if n not in declared]
    assert all(n.startswith("us3d_debug_") for n in header_functions(DEBUG_HEADER))
    assert not [n for n in header_functions(DEBUG_HEADER) if n not in exported]


def test_abi_version_matches_header():
    from unscene3d_b200 import _lib

    ver = int(re.search(r"#define US3D_ABI_VERSION (\d+)", open(HEADER).read()).group(1))
    assert _lib.ABI_VERSION == ver


def test_argument_validation_reports_errors_without_a_device():
    from unscene3d_b200 import _lib

    lib = _lib.lib
    assert lib.us3d_hash_capacity(1000) == 2048
    assert lib.us3d_hash_capacity(0) == 1024
    rc = lib.us3d_spconv_gather(0, 4, 0, 10, 99, 0, 4, 4, 0, 0, 0, 0, 0, 4, 0, 0, 0)
    assert rc != 0 and b"kvol" in lib.us3d_last_error()
    with pytest.raises(_lib.Us3dError):
        _lib.check(lib.us3d_kernel_map(0, 10, (ctypes.c_int32 * 3)(), 1, 0, 0, 1000, 0, 0, 128, 0))


def test_shim_modules_import_under_reference_names():
    import unscene3d_b200  # noqa: F401
    import MinkowskiEngine as ME
    import MinkowskiEngine.MinkowskiOps as me
    from MinkowskiEngine import MinkowskiNetwork, MinkowskiReLU, SparseTensor  # noqa: F401
    from MinkowskiEngine.MinkowskiPooling import MinkowskiAvgPooling  # noqa: F401

    for name in ["SparseTensor", "MinkowskiConvolution", "MinkowskiConvolutionTranspose", "KernelGenerator", "RegionType",
                 "MinkowskiBatchNorm", "MinkowskiInstanceNorm", "MinkowskiReLU", "MinkowskiAvgPooling", "MinkowskiSumPooling",
                 "MinkowskiAvgUnpooling", "MinkowskiNetwork", "utils", "MinkowskiAlgorithm", "SparseTensorQuantizationMode", "TensorField"]:
        assert hasattr(ME, name), name
    assert [ME.RegionType(i) for i in range(3)]  # evaluated at import by models/modules/common.py:70
    assert hasattr(me, "cat") and hasattr(me, "SparseTensor")
    assert all(hasattr(ME.utils, n) for n in ("sparse_quantize", "sparse_collate", "batched_coordinates"))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "unscene3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"


def test_launch_list_rejects_unknown_ops_without_a_device():
    """us3d_run_ops validates the list before anything is launched: an unknown op kind is an argument error (no GPU needed), an
    empty list is a no-op."""
    import ctypes

    from unscene3d_b200 import _lib

    lib = _lib.lib
    assert lib.us3d_run_ops(None, 0, None) == 0
    op = _lib.Op()
    op.kind = 99
    rc = lib.us3d_run_ops(ctypes.byref(op), 1, None)
    assert rc != 0 and b"unknown op kind 99" in lib.us3d_last_error()
    flat = (ctypes.c_longlong * 27)(*([77] + [0] * 26))
    f = (ctypes.c_double * 2)(0.0, 0.0)
    assert lib.us3d_run_ops_flat(flat, f, 1, None) != 0 and b"unknown op kind 77" in lib.us3d_last_error()


def test_launch_list_records_and_gradient_arena_host_logic():
    """Host logic behind the launch lists (no device): every op is one row of 27 int64 (kind, 16 pointers, 10 integers) + 2 floats —
    the layout us3d_run_ops_flat reads —, held tensors stay referenced until the list is dropped; the zero-filled arena of the weight
    gradients hands out disjoint, 256-byte aligned, all-zero slices and starts a new buffer when one is exhausted."""
    import torch

    from unscene3d_b200 import _lib
    from unscene3d_b200.engine import blocks as B
    from unscene3d_b200.engine import functional as Fn

    ops = B._Ops()
    ops.emit(_lib.OP_BN_APPLY, (11, 12, 13), (1, 2))
    ops.emit(_lib.OP_CONV, tuple(range(100, 116)), tuple(range(10)), 1e-5, 0.02)
    assert ops.n == 2 and len(ops.a) == 2 * 27 and len(ops.f) == 4
    assert ops.a[:27] == [_lib.OP_BN_APPLY, 11, 12, 13] + [0] * 13 + [1, 2] + [0] * 8
    assert ops.a[27] == _lib.OP_CONV and ops.a[28:44] == list(range(100, 116)) and ops.a[44:54] == list(range(10))
    assert ops.f[2:] == [1e-5, 0.02]
    t = torch.zeros(3)
    assert ops.hold(t, None) is t and ops.keep[0] is t
    assert ctypes_sizeof_op() == 4 + 4 + 16 * 8 + 10 * 8 + 2 * 4  # int kind (+ padding), p[16], v[10], f[2]

    arena = Fn._ZeroArena()
    dev = torch.device("cpu")
    a = arena.take(1000, dev)
    b = arena.take(70, dev)
    assert a.numel() == 1000 and b.numel() == 70 and float(a.abs().sum()) == 0.0 and float(b.abs().sum()) == 0.0
    assert b.data_ptr() - a.data_ptr() == 1024 * 4  # slices start on 64-float (256-byte) granules of the buffer
    a.fill_(1.0)
    assert float(b.abs().sum()) == 0.0
    big = arena.take((16 << 20) + 5, dev)  # larger than what is left: a fresh zero buffer, the earlier slices stay valid
    assert big.numel() == (16 << 20) + 5 and float(big[:1000].abs().sum()) == 0.0 and float(a.sum()) == 1000.0


def ctypes_sizeof_op():
    import ctypes

    from unscene3d_b200 import _lib

    return ctypes.sizeof(_lib.Op)
