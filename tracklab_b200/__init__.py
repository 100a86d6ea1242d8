"""B200-native (sm_100a) implementation of TrackLab's per-frame tracking hot path.

Only the hot path named by BASELINE.json lives here: the C-ABI CUDA library (``csrc/`` ->
``libtrackkern.so``), its ctypes binding, and the host-side mirror of the reference's
``tracklab.pipeline`` module API. There is no CPU fallback: importing the compute wrappers without
the built library, or calling them without a CUDA device, raises.
"""
__version__ = "0.1.0"
