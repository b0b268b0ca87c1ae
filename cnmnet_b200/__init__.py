"""cnmnet_b200 -- B200-native (sm_100a) implementation of the CNMNet inference hot path.

Python keeps the reference's ``depthnet`` module API; every forward lands in hand-written CUDA
behind the C ABI of ``include/cnmnet_b200.h`` (``libcnmnet_b200.so``).  There is no CPU or
PyTorch fallback: calls raise if the library is missing or no CUDA device is present.
"""
__version__ = "0.1"

from .depthnet.depthNet_model import depthNet, DepthRefineNet  # noqa: F401
from .depthnet.depth_util import (Depth2normal, get_normal_by_planes, get_pixel_coordinates,  # noqa: F401
                                  process_camera_parameters)
from .depthnet.inverse_warp import inverse_warp, pixel2cam  # noqa: F401
from .pipeline import CNMPipeline  # noqa: F401
from .video import VideoDepthStream, reference_tuples  # noqa: F401
from .metrics import depth_metrics  # noqa: F401
from . import sevenscenes  # noqa: F401
