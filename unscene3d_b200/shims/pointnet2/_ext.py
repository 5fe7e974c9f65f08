"""`pointnet2._ext` as bound by third_party/pointnet2/_ext_src/src/bindings.cpp:9-22."""
from unscene3d_b200.engine.functional import furthest_point_sampling  # noqa: F401


def _off_path(name):
    def fn(*a, **k):
        raise NotImplementedError(f"pointnet2._ext.{name} is not used by the UnScene3D hot path (SURVEY.md §2.2 N1)")

    fn.__name__ = name
    return fn


gather_points = _off_path("gather_points")
gather_points_grad = _off_path("gather_points_grad")
ball_query = _off_path("ball_query")
group_points = _off_path("group_points")
group_points_grad = _off_path("group_points_grad")
three_nn = _off_path("three_nn")
three_interpolate = _off_path("three_interpolate")
three_interpolate_grad = _off_path("three_interpolate_grad")
