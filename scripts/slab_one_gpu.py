"""Profiling helper: the slab path with several ranks sharing ONE GPU (one host thread per rank), so that ncu can see
the SLAB variants of the kernels on a single-GPU box. Usage: slab_one_gpu.py [particles] [ranks] [steps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
from active_particle_jamming_b200.slab import SlabBox
from bench import synthetic_state

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2097152
ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 120
R, L, x, y, phi = synthetic_state(n, 0.9, 12345)
box = SlabBox(n, L, ranks, seed=12345, max_neighbors=64)
for r in box.local:
    r.set_timeout(3.0)        # under ncu the ranks' kernels are serialised: a wait that cannot be met gives up quickly
box.upload(x=x, y=y, R=R, phi=phi)
box.skip_self_term_once()
box.set_activity(0.05, 0.5)
box.step(steps)
print("slab_one_gpu: n=%d ranks=%d steps=%d counters=%s checksum=%016x" % (n, ranks, steps, box.counters(), box.checksum()))
box.close()
