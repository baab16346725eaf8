"""viennaray_b200 -- a B200-native (sm_100a CUDA) implementation of ViennaRay's
Monte Carlo flux hot path behind a C ABI (include/viennaray_b200.h)."""
__version__ = "0.1.0"
