"""maplab_b200 — B200-native loop-closure query path of ethz-asl/maplab.

Host-side mirror of the reference API over the C-ABI library libmaplab_lc_b200.so
(include/maplab_lc_b200.h). Importing the package does not load CUDA; the first detector does.
"""
from . import capi  # noqa: F401
from .capi import Detector, MlcError, default_settings, default_ransac_settings  # noqa: F401
from .loop_detector import LoopDetector, ProjectedImage  # noqa: F401
