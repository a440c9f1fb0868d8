"""Third validation gate of north_star: LONG runs agree statistically with the reference algorithm on
the polar order parameter, the MSD (value and exponent) and the density-fluctuation scaling.

Trajectories cannot be compared beyond a few hundred steps (the orientation dynamics amplifies last-bit
differences by a decade per ~25 steps, DESIGN.md "Parity"), so this compares ENSEMBLES: S independent
systems on the GPU (batched as replicas of one handle, Philox noise) against S independent runs of the
oracle (oracle/apj_oracle.c, pinned bit-for-bit to the reference build) started from different initial
conditions and driven by a different noise stream. Every statistic is compared as ensemble mean against
ensemble mean with the standard error measured on the ensembles themselves (z-score <= 4 plus a small
absolute floor), at two points of the phase diagram (ordered and disordered): 12 systems of N = 4096 over
2e4 steps each side. The oracle runs are independent serial programs: one subprocess per system, so the CPU
side takes about a minute on the GPU box's host cores."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if __name__ == "__main__":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

from _util import random_system
from oracle.pyoracle import OracleSim

pytestmark = pytest.mark.gpu

S, N, RHO = 12, 4096, 0.9
RELAX, RAMP, STEPS, EVERY = 200, 200, 20000, 100
RADII = (5.0, 8.0, 12.0, 17.0, 24.0)


def _gpu_ensemble(l_s, l_n, seed0):
    from active_particle_jamming_b200 import DeviceEngine
    parts = [random_system(N, RHO, seed0 + s) for s in range(S)]
    Ls = [p[1] for p in parts]
    R, x, y, phi = (np.concatenate([p[k] for p in parts]) for k in (0, 2, 3, 4))
    with DeviceEngine(N, Ls, n_systems=S, seed=seed0) as e:
        e.upload(x=x, y=y, R=R, phi=phi)
        e.skip_self_term_once()
        e.set_activity(0.0, l_n); e.step(RELAX)                      # relax(): jamming.cpp:482-525
        e.set_activity(l_s, l_n); e.set_ramp(RAMP); e.step(RAMP); e.set_ramp(0)
        e.mark_origin()
        order, msd, area = [], [], []
        for _ in range(STEPS // EVERY):
            e.step(EVERY)
            order.append(e.order_orientation()[0].copy())
            msd.append(e.msd().copy())
            area.append([e.fluct_area(np.full(S, r)).copy() for r in RADII])
    return np.array(Ls), np.array(order), np.array(msd), np.array(area)        # (T,S), (T,S), (T,R,S)


def _oracle_system(l_s, l_n, seed0, s):
    """One oracle run (system s of the ensemble): L, order(t), msd(t), area(t, radius)."""
    R, L, x, y, phi = random_system(N, RHO, seed0 + s)
    o = OracleSim.from_arrays(R, x, y, phi, RHO)
    o.topology(); o.assign(); o.build(); o.mark_origin()
    o.set_params(0.0, l_n); o.run_philox(seed0 + 7, 0, RELAX, replica=s)
    for k in range(RAMP):                                        # CFself ramps linearly (jamming.cpp:518)
        o.set_params(l_s * k / RAMP, l_n); o.run_philox(seed0 + 7, RELAX + k, 1, replica=s)
    o.set_params(l_s, l_n)
    o.mark_origin()
    so, sm, sa = [], [], []
    for k in range(STEPS // EVERY):
        o.run_philox(seed0 + 7, RELAX + RAMP + k * EVERY, EVERY, replica=s)
        so.append(o.order()); sm.append(o.msd()); sa.append([o.fluct_area(r) for r in RADII])
    o.close()
    return dict(L=L, order=so, msd=sm, area=sa)


def _oracle_ensemble(l_s, l_n, seed0):
    """S independent serial runs, one subprocess each (this file run as a script)."""
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), json.dumps([l_s, l_n, seed0, s])], stdout=subprocess.PIPE, text=True,
                              cwd=ROOT) for s in range(S)]
    outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
    assert all(p.returncode == 0 for p in procs)
    Ls = np.array([o["L"] for o in outs])
    order = np.array([o["order"] for o in outs]).T
    msd = np.array([o["msd"] for o in outs]).T
    area = np.transpose(np.array([o["area"] for o in outs]), (1, 2, 0))
    return Ls, order, msd, area


def _stats(Ls, order, msd, area):
    """Per-system statistics (arrays over the ensemble)."""
    T = order.shape[0]
    t = EVERY * np.arange(1, T + 1)
    half = T // 2
    out = {"order": order[half:].mean(axis=0), "log_msd_end": np.log(msd[-1])}
    sel = t >= 500
    lt = np.log(t[sel])
    out["msd_exponent"] = np.array([np.polyfit(lt, np.log(msd[sel, s]), 1)[0] for s in range(order.shape[1])])
    # density fluctuations (Fluctuations.h:51-87): rms fluctuation of the disk area inside a circle of radius r
    # around the COM vs its mean, over time; the scaling exponent is the log-log slope over the radii
    mean_a = area[half:].mean(axis=0)                                # (R,S)
    rms_a = area[half:].std(axis=0)
    out["fluct_exponent"] = np.array([np.polyfit(np.log(mean_a[:, s]), np.log(rms_a[:, s]), 1)[0] for s in range(order.shape[1])])
    out["mean_area_r8"] = mean_a[1]
    return out


@pytest.mark.parametrize("l_s,l_n", [(0.3, 0.15), (0.3, 0.9)])          # ordered (flocking) / disordered
def test_ensemble_statistics_match_the_reference_algorithm(l_s, l_n):
    g = _stats(*_gpu_ensemble(l_s, l_n, seed0=4000))
    o = _stats(*_oracle_ensemble(l_s, l_n, seed0=9000))
    floors = {"order": 0.01, "log_msd_end": 0.05, "msd_exponent": 0.03, "fluct_exponent": 0.06, "mean_area_r8": 0.5}
    report = {}
    for k in g:
        gm, om = g[k].mean(), o[k].mean()
        sem = np.sqrt(g[k].var(ddof=1) / S + o[k].var(ddof=1) / S)
        report[k] = (gm, om, sem)
        assert abs(gm - om) <= 4.0 * sem + floors[k], "%s: gpu %.4f vs oracle %.4f (sem %.4f)" % (k, gm, om, sem)
    # the two points are physically different: the gate is not vacuous
    if l_n < 0.5:
        assert report["order"][0] > 0.5 and report["order"][1] > 0.5
    else:
        assert report["order"][0] < 0.3 and report["order"][1] < 0.3
    print({k: tuple(round(float(x), 4) for x in v) for k, v in report.items()})


if __name__ == "__main__":       # worker of _oracle_ensemble
    a = json.loads(sys.argv[1])
    print(json.dumps(_oracle_system(float(a[0]), float(a[1]), int(a[2]), int(a[3]))))
