"""rfsi_b200 -- B200-native RFSI prediction engine (near.obs + random-forest forward pass).

The compute lives in librfsi_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/rfsi_b200.h); this package is the host-side mirror of the reference's R call
shapes.  There is no CPU fallback.
"""
from ._lib import RfsiError, device_count, launch_count  # noqa: F401
from .api import Prediction, Raster, feature_map, near_obs, near_obs_days, predict, predict_fused, raster_extract  # noqa: F401
from .forest import FlatForest, RangerForest, flatten_ranger, from_sklearn, ranger_dict_from_flat  # noqa: F401

__version__ = "0.1.0"
