"""edgegan_b200 -- B200-native (sm_100a) EdgeGAN training / inference hot path.

Host side mirrors the reference's `edgegan.nn` / `edgegan.models` surface; all
arithmetic runs in hand-written CUDA kernels behind the C ABI in include/edgegan_b200.h.
"""
from .config import Flags, update_flags  # noqa: F401

__version__ = "0.1.0"
