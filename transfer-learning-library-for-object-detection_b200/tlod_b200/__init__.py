"""tlod_b200 -- B200-native (sm_100a) RoI / proposal hot path of
live-group/Transfer-Learning-Library-for-Object-Detection.

Host side of the C ABI in include/tlod_b200.h.  The reference-facing operator
surface (``model.roi_align.modules.roi_align.RoIAlignAvg`` ...) lives in the
sibling ``model`` package; put this directory on ``sys.path`` the way the
reference's ``_init_paths.py`` adds ``lib/``.
"""
from . import _lib  # noqa: F401  (raises if libtlod_b200.so is missing)
from . import functional  # noqa: F401
from . import sharding  # noqa: F401
from .autograd import (DALossFunction, ImageDALossFunction, image_da_losses, GradReverse, RoIAlignAvgFunction, RoIAlignFunction,  # noqa: F401
                       RoICropFunction, RoICropPoolFunction, RoIPoolFunction, RPNLossFunction, SpaceToDepthFunction, da_losses, rpn_losses, space_to_depth,
                       grad_reverse)
from ._lib import TlodError, launch_count  # noqa: F401

__all__ = ["functional", "RoIAlignFunction", "RoIAlignAvgFunction", "RoICropFunction", "RoICropPoolFunction", "RoIPoolFunction", "GradReverse", "grad_reverse", "DALossFunction",
           "da_losses", "ImageDALossFunction", "image_da_losses", "RPNLossFunction", "rpn_losses", "SpaceToDepthFunction", "space_to_depth", "TlodError", "launch_count"]
