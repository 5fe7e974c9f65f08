"""Model definitions of the hot path (host-side mirror of the reference's `models/` package).

`load_model(name)` mirrors the reference registry (models/__init__.py:18-31) used by the pseudo-mask
path (pseudo_masks/unscene3d_pseudo_main.py:59).
"""
from . import criterion, mask3d, matcher, res16unet
from .criterion import SetCriterion
from .mask3d import Mask3D
from .matcher import HungarianMatcher
from .res16unet import (Res16UNet14, Res16UNet14A, Res16UNet18B, Res16UNet18D, Res16UNet34, Res16UNet34A,
                        Res16UNet34C, Res16UNet34CMultiRes, Res16UNet34D, Custom30M)

MODELS = [getattr(res16unet, a) for a in dir(res16unet) if "Net" in a and isinstance(getattr(res16unet, a), type)]


def get_models():
    return MODELS


def load_model(name):
    table = {m.__name__: m for m in MODELS}
    if name not in table:
        print("Invalid model index. Options are:")
        for m in MODELS:
            print(f"\t* {m.__name__}")
        return None
    return table[name]
