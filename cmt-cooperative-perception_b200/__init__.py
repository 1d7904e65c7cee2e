"""cmtcoop_b200 -- B200-native (sm_100a) implementation of the CMT / CMTCoop cross-modal
token-fusion hot path behind the reference's mmdet3d_plugin class names.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/cmtcoop_b200.h)
  _lib.py, ops.py  ctypes binding of the C ABI on torch tensors (no CPU fallback)
  plugin/          host-side mirror of the reference classes (CmtHead, CmtTransformer, FlashMHA, ...)
  synth.py         deterministic synthetic configs / weights / inputs for tests and bench
"""
__version__ = "0.1.0"
