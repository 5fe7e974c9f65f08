"""Drop-in `pointnet2` package: only `_ext.furthest_point_sampling` is on the hot path
(third_party/pointnet2/pointnet2_utils.py:22-30, models/mask3d.py:228)."""
