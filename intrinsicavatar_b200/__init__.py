"""intrinsicavatar_b200: B200-native (sm_100a) implementation of IntrinsicAvatar's per-frame render path.

The numeric work lives in ``libia_b200.so`` (csrc/, C ABI in include/ia_b200.h); this package is the
host-side mirror of the reference's model interface.  Nothing here falls back to CPU/PyTorch compute.
"""
__all__ = ["capi", "engine", "model", "snarf", "body", "weights", "synthetic", "parallel"]
