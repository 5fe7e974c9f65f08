"""`from MinkowskiEngine.MinkowskiPooling import MinkowskiAvgPooling` (models/mask3d.py:5)."""
from unscene3d_b200.engine import (MinkowskiAvgPooling, MinkowskiAvgUnpooling, MinkowskiMaxPooling,  # noqa: F401
                                   MinkowskiSumPooling)
