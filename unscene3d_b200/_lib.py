"""ctypes binding of libus3d.so (include/us3d.h).  The library is the product: if it is missing or
does not export a declared symbol the import FAILS — there is no Python/CPU fallback."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("US3D_LIB") or os.path.join(_HERE, "csrc", "libus3d.so")  # US3D_LIB: A/B builds of the same ABI (scripts/)

_i, _ll, _f = ctypes.c_int, ctypes.c_longlong, ctypes.c_float
_p = ctypes.c_void_p

class BnFuse(ctypes.Structure):
    """us3d_bn_fuse_t (include/us3d.h): BatchNorm statistics folded into us3d_spconv_gather_mt_bn."""

    _fields_ = [("ws", _p), ("mean", _p), ("invstd", _p), ("running_mean", _p), ("running_var", _p), ("num_batches_tracked", _p),
                ("eps", _f), ("momentum", _f)]


class Op(ctypes.Structure):
    """us3d_op_t (include/us3d.h): one launch of a launch list (us3d_run_ops)."""

    _fields_ = [("kind", _i), ("p", _p * 16), ("v", _ll * 10), ("f", _f * 2)]


OP_CONV, OP_BN_APPLY, OP_BN_BACKWARD, OP_WGRAD, OP_ADD = 1, 2, 3, 4, 5

# name -> argtypes (all functions return int unless listed in _RESTYPE)
PROTOTYPES = {
    "us3d_abi_version": [],
    "us3d_last_error": [],
    "us3d_launch_count": [],
    "us3d_reset_launch_count": [],
    "us3d_hash_capacity": [_i],
    "us3d_coords_unique": [_p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p, _p, ctypes.POINTER(_i), _p],
    "us3d_coords_unique_h": [_p, _i, _i, _p, _p],
    "us3d_kernel_map": [_p, _i, ctypes.POINTER(ctypes.c_int32), _i, _p, _p, _i, _p, _p, _i, _p],
    "us3d_spconv_gather": [_p, _i, _p, _i, _i, _p, _i, _i, _i, _i, _p, _p, _p, _i, _i, _p, _p],
    "us3d_spconv_tc_supported": [_i, _i],
    "us3d_spconv_packed_bytes": [_i, _i, _i, _i],
    "us3d_spconv_pack_weights": [_p, _i, _i, _i, _i, _i, _i, _p, _p],
    "us3d_split_bf16": [_p, _i, _i, _i, _p, _p, _p],
    "us3d_spconv_gather_mt": [_p, _p, _i, _p, _i, _i, _p, _i, _i, _i, _p, _p, _p, _i, _i, _p, _p, _p, _ll, _p],
    "us3d_spconv_gather_mt_bn": [_p, _p, _i, _p, _i, _i, _p, _i, _i, _i, _p, _p, _p, _i, _i, _p, _p, _p, _ll, _p, _p],
    "us3d_run_ops": [_p, _i, _p],
    "us3d_run_ops_flat": [_p, _p, _i, _p],
    "us3d_spconv_gather_mt_workspace_bytes": [_i, _i, _i],
    "us3d_spconv_partition_size": [],
    "us3d_spconv_partition": [_p, _i, _i, _p, _p],
    "us3d_spconv_wgrad_tc_supported": [_i, _i],
    "us3d_spconv_wgrad_planes": [_p, _p, _p, _p, _p, _i, _i, _p, _i, _i, _i, _p, _p, _p],
    "us3d_permute_planes": [_p, _p, _p, _i, _i, _p, _p, _p],
    "us3d_neighbour_pattern_keys": [_p, _i, _i, _p, _p, _p],
    "us3d_kernel_map_reorder": [_p, _i, _i, _p, _p, _p, _i, _p],
    "us3d_spconv_wgrad": [_p, _i, _p, _i, _i, _p, _i, _p, _p, _i, _i, _p],
    "us3d_bn_stats": [_p, _i, _i, _i, _p, _p, _p],
    "us3d_bn_finalize": [_p, _p, _i, _i, _f, _f, _p, _p, _p, _p, _p],
    "us3d_bn_apply": [_p, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _p, _i, _p],
    "us3d_bn_bwd_reduce": [_p, _i, _p, _i, _p, _i, _i, _i, _p, _p, _i, _p, _p],
    "us3d_bn_bwd_apply": [_p, _i, _p, _i, _p, _i, _i, _i, _p, _p, _p, _i, _p, _p, _i, _p, _i, _p, _p, _p],
    "us3d_bn_batch_stats": [_p, _i, _i, _i, _f, _f, _p, _p, _p, _p, _p, _p],
    "us3d_bn_backward": [_p, _i, _p, _i, _p, _i, _i, _i, _p, _p, _p, _i, _i, _p, _p, _i, _p, _i, _p, _p, _p],
    "us3d_bn_workspace_bytes": [_i],
    "us3d_bn_stats_fused": [_p, _i, _i, _i, _f, _f, _p, _p, _p, _p, _p, _p, _p],
    "us3d_bn_apply_planes": [_p, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _p, _i, _p, _p, _p],
    "us3d_bn_backward_planes": [_p, _i, _p, _i, _p, _i, _i, _i, _p, _p, _p, _i, _i, _p, _p, _i, _p, _i, _p, _p, _p, _p, _p],
    "us3d_spconv_pack_pair": [_p, _i, _i, _i, _i, _i, _p, _p, _p],
    "us3d_spconv_pack_many": [_p, _i, _i, _p],
    "us3d_stem_conv_supported": [_i, _i, _i],
    "us3d_stem_conv_fwd": [_p, _i, _p, _i, _i, _p, _i, _i, _p, _p, _i, _p],
    "us3d_stem_conv_wgrad": [_p, _i, _p, _i, _i, _p, _i, _p, _i, _i, _p],
    "us3d_relu": [_p, _p, _ll, _p],
    "us3d_relu_bwd": [_p, _p, _p, _ll, _p],
    "us3d_add": [_p, _p, _p, _ll, _p],
    "us3d_copy2d": [_p, _i, _p, _i, _i, _i, _p],
    "us3d_pool_fwd": [_p, _i, _p, _i, _i, _i, _p, _p],
    "us3d_pool_bwd": [_p, _p, _p, _i, _p, _i, _i, _i, _p, _p],
    "us3d_furthest_point_sampling": [_p, _i, _i, _i, _p, _p, _p],
    "us3d_segment_mean_fwd": [_p, _p, _i, _i, _i, _p, _p, _p],
    "us3d_segment_mean_bwd": [_p, _p, _p, _i, _i, _p, _p],
    "us3d_segment_mean_f64": [_p, _p, _i, _i, _i, _p, _p, _p, _p],
    "us3d_ncut_gram": [_p, _i, _i, _p, _p, _p, _p],
    "us3d_ncut_threshold": [_p, _p, _i, _p, _p, _f, ctypes.c_double, _p, _p, _p, _p],
    "us3d_ncut_matvec": [_p, _i, ctypes.c_double, _p, _p, _p, _p],
    "us3d_project_features_2d3d": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _f, _i, _p, _p, _p, _p],
    "us3d_felzenszwalb_segment_h": [_p, _p, _p, _i, _i, _f, _i, _p, _p, _i],
    "us3d_project_voxels_to_planes": [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "us3d_project_voxels_to_planes_bwd": [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p],
    "us3d_ncut_lanczos_workspace_bytes": [_i],
    "us3d_ncut_lanczos": [_p, _i, ctypes.c_double, _p, _p, _p, _p, _i, _i, _i, ctypes.c_double, _p, _ll, _p, _p],
    "us3d_xattn_workspace_bytes": [_i, _i, _i, _i, _i],
    "us3d_xattn_fwd": [_p, _p, _p, _p, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _i, _f, _p, _p, _p, _p],
    "us3d_xattn_bwd": [_p, _p, _p, _p, _ll, _ll, _ll, _ll, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p],
    "us3d_freemask_soft_masks": [_p, _i, _i, _p, _p, _p],
    "us3d_freemask_row_stats": [_p, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _p, _p],
    "us3d_freemask_weighted_inter": [_p, _i, _i, _i, _f, _p, _p, _p],
    "us3d_freemask_separate_h": [_p, _i, _i, _p, _p, _p, _p, _p, _i, _ll],
    "us3d_pooled_mask_bits": [_p, _p, _p, _i, _p, _i, _p, _p],
    "us3d_mask_loss_fwd": [_p, _i, _i, _p, _i, _p, _p, _i, _p, _f, _p, _p, _p],
    "us3d_mask_loss_bwd": [_p, _i, _i, _p, _i, _p, _p, _i, _p, _f, _p, _p, _p, _p],
    "us3d_fourier_posenc": [_p, _i, _i, _p, _p, _p, _i, _i, _p, _p],
    "us3d_matcher_cost": [_p, _i, _i, _p, _i, _p, _i, _p, _f, _f, _f, _p, _p],
}
_RESTYPE = {"us3d_last_error": ctypes.c_char_p, "us3d_launch_count": _ll, "us3d_reset_launch_count": None,
            "us3d_spconv_packed_bytes": _ll, "us3d_xattn_workspace_bytes": _ll,
            "us3d_spconv_gather_mt_workspace_bytes": _ll, "us3d_ncut_lanczos_workspace_bytes": _ll}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m unscene3d_b200.csrc.build` "
            "(or __graft_entry__.build()).  unscene3d_b200 has no fallback path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError = stale library
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, _i)
    return lib


lib = _load()
ABI_VERSION = lib.us3d_abi_version()


class Us3dError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise Us3dError(lib.us3d_last_error().decode() or f"libus3d error {rc}")


def launch_count() -> int:
    return int(lib.us3d_launch_count())


def reset_launch_count():
    lib.us3d_reset_launch_count()
