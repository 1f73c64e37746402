"""weathermodel_b200 -- B200 (sm_100a) implementation of the WeatherModel training hot path.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/wm_b200.h) -> libwm_b200.so
  _lib.py, ops.py  ctypes binding and torch-tensor front end of the C ABI
  (host-side mirror of the reference's src/ tree is added next to these)
"""
__version__ = "0.1.0"
