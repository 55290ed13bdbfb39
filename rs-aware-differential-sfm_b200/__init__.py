"""rs-aware-differential-sfm_b200: B200-native dense optimisation core of RS-aware differential SfM.

Product path = csrc/ (sm_100a CUDA kernels behind the extern "C" ABI of include/rsdsfm.h) plus the
host-side mirror of the reference's C++ surface (host/).  `capi` is the ctypes binding used by the
tests and bench.py; `synth` generates the analytic benchmark inputs.
"""
