import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from _util import device_from_state, rel_err
from test_gpu_parity import relaxed_oracle
N, rho = 65536, 0.9
o, _ = relaxed_oracle(N, rho, seed=40, l_s=0.05, l_n=0.5, presteps=6)
L = o.scalars()["L"]
for cutoff in (20.0, 60.0, 140.0):
    vel, ori, pair = o.spatial_correlations(cutoff)
    with device_from_state(o.state(), seed=1) as e:
        c = e.spatial_correlations(cutoff)
    gv, go = c["vel_sum"][0] / c["counts"][0], c["ori_sum"][0] / c["counts"][0]
    norm = 2 * L * L / (2 * 3.14159265 * 0.1 * float(N) * float(N))
    print("cutoff", cutoff, "counts", c["counts"][0][:4], c["counts"][0][-2:], "nan", np.isnan(gv).sum(), np.isnan(vel).sum())
    print("  vel err", np.nanmax(np.abs(gv - vel)), "ori err", np.nanmax(np.abs(go - ori)), "pair err", np.nanmax(np.abs(c["pair_sum"][0] * norm - pair)), "pair scale", np.nanmax(pair))
    k = int(np.nanargmax(np.abs(gv - vel))); print("  worst vel bin", k, gv[k], vel[k], " worst pair bin", int(np.nanargmax(np.abs(c["pair_sum"][0] * norm - pair))))
