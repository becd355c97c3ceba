"""pyp_b200 — B200-native (sm_100a) engine behind nextPYP's CSP refine3d / reconstruct3d path.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C-ABI of
include/cspb200.h), the ctypes binding, and the host-side mirror of the reference's
binary front-ends.  See DESIGN.md.
"""
__version__ = "0.1.0"
