"""B200-native hot path of danielmccusker/active-particle-jamming (2D overdamped SPP step).

Layout: csrc/ (sm_100a CUDA kernels + the C ABI of include/apj_b200.h), host/ (C++ mirror of the
reference's Engine / code/classes interface that calls the ABI), device.py (ctypes binding used
by tests and bench.py), slab.py (one box over several GPUs: host-side decomposition arithmetic), sweep.py (a phase-diagram input.txt dealt out over the GPUs of a node, one `jam --sweep` process per GPU). There is no CPU fallback anywhere in this package.
"""
from .device import DeviceEngine, ApjError, load_library, PI, PI2  # noqa: F401
from .slab import SlabRank, SlabBox, DistSlab  # noqa: F401
from ._build import build_library  # noqa: F401
