"""manhattanslam_b200 -- B200-native RGB-D front-end for ManhattanSLAM.

Python mirror of the reference's operator surfaces (ORBextractor, ORBmatcher window searches,
PlaneDetection pre-stage, SurfelFusion) over the C-ABI library libmsl_frontend.so whose kernels are
hand-written sm_100a CUDA.  Python is harness/plumbing only; the product is the C ABI in
include/msl_frontend.h plus the C++ adapters in adapters/.
"""
from ._lib import MslError, lib  # noqa: F401
from .orb import ORBextractor, KP_DTYPE  # noqa: F401
from .surfel import SurfelFusion, SURFEL_DTYPE, SEED_DTYPE  # noqa: F401
from .plane import PlaneDetection, BLOCK_DTYPE  # noqa: F401
from .matcher import ORBmatcher, frame_geom, GEOM_DTYPE  # noqa: F401

__version__ = "0.1.0"
from .glue import FrameGlue  # noqa: F401
