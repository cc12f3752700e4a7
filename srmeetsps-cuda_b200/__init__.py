"""srmeetsps-cuda_b200: B200-native (sm_100a) SRmeetsPS outer loop behind the reference's
SRPS / DataHandler surface.  The compute path is libsrps_b200.so (hand-written CUDA, C ABI in
include/srps_c_api.h); this package is the host-side mirror of the reference interface."""
from .context import Context, SRPSError            # noqa: F401
from .srps import (DataHandler, ImageDataHandler, MatFileDataHandler, Preferences, SRPS,  # noqa: F401
                   preprocess_depth)
from .snapshot import read_snapshot, write_snapshot  # noqa: F401
