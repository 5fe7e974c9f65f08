"""`custom_cuda_utils` (utils/cuda_utils/cuda_utils.cpp:49-54).  models/noise_robust_loss.py:11 imports it at module
import; its kernels only execute when cost_noise_robust != 0 (models/criterion.py:170), which the self-training
configuration never sets (conf/matcher/hungarian_matcher.yaml:6), so the entry points exist and fail loudly."""


def _not_built(name):
    def fn(*a, **k):
        raise NotImplementedError(f"custom_cuda_utils.{name}: the tri-plane noise-robust loss is off by default "
                                  "(cost_noise_robust = 0) and is not built in this round")

    fn.__name__ = name
    return fn


project_sparse_voxels_to_planes = _not_built("project_sparse_voxels_to_planes")
project_sparse_voxels_to_planes_backward = _not_built("project_sparse_voxels_to_planes_backward")
trilinear_interpolate = _not_built("trilinear_interpolate")
trilinear_interpolate_backward = _not_built("trilinear_interpolate_backward")
