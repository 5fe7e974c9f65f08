"""`ME.utils.*` (datasets/utils.py:266-287, 403-432)."""
from unscene3d_b200.engine.utils import batched_coordinates, sparse_collate, sparse_quantize  # noqa: F401
