"""`import MinkowskiEngine.MinkowskiOps as me` (models/res16unet.py:1, models/mask3d.py:4)."""
from unscene3d_b200.engine import SparseTensor, cat  # noqa: F401
