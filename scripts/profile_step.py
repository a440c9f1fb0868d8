"""Short run for ncu: N particles, crude relax, then a few steps (no graph so every launch is a kernel)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from active_particle_jamming_b200 import DeviceEngine, PI
from active_particle_jamming_b200.device import FLAG_NO_GRAPH

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rho = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
G = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rng = np.random.default_rng(3)
R = 1 + 0.1 * rng.standard_normal(N)
L = float(np.sqrt(3.14159265 * np.sum(R * R) / rho))
x = rng.uniform(-L / 2, L / 2, N); y = rng.uniform(-L / 2, L / 2, N); phi = rng.uniform(-PI, PI, N)
e = DeviceEngine(N, L, seed=1, flags=FLAG_NO_GRAPH, lanes_per_particle=G)
e.upload(x=x, y=y, R=R, phi=phi); e.mark_origin()
e.set_activity(0.0, 0.5); e.step(400)
e.set_activity(0.05, 0.5); e.step(nsteps)
print(e.counters())
