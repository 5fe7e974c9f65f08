"""`raycast_cuda` (utils/cuda_utils/raycast_cuda.cpp): the 3D -> 2D rendering of voxel features.  utils/cuda_utils/raycast_image.py
imports it at module import next to `project_features_cuda`; neither the training nor the pseudo-mask path calls it, so the entry
points exist and say so when called."""


def _not_built(name):
    def fn(*a, **k):
        raise NotImplementedError(f"raycast_cuda.{name}: the 3D -> 2D ray casting is not on the training or pseudo-mask path and is not built")

    fn.__name__ = name
    return fn


raycast_features = _not_built("raycast_features")
raycast_features_backward = _not_built("raycast_features_backward")
raycast_interpolate_features = _not_built("raycast_interpolate_features")
raycast_interpolate_backward = _not_built("raycast_interpolate_backward")
