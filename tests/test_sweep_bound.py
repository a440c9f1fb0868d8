"""The bound behind the skin-aware sweep length of the CUDA step kernel (apj_device.cuh, APJ_CLASSES), checked on the
REFERENCE dynamics with the oracle alone (no GPU): for every pair of the Verlet list, the current distance d, the
distance d_b at the last list build and the skin-test value D = sqrt(l1) + sqrt(l2) of newSkinList
(reference jamming.cpp:596-611, COM-drift corrected, evaluated on the same state) satisfy d >= d_b - D. Hence an entry
built at d_b >= rn + tau cannot interact (d < rn) while D <= tau -- which is why a step may stop after the list classes
with (k+1) * skin / 4 >= D and still give the bits of the full sweep. Also pins the class-selection rule itself."""
import numpy as np
import pytest

from _util import random_system
from oracle.pyoracle import PI, OracleSim

RN, RS = 2.8, 4.2
SKIN = RS - RN
EPS = 1e-9


def wrap(d, L):
    return d - L * np.round(d / L)


def skin_value(o, L):
    sc = o.scalars()
    dx = wrap(o.x - o.xo - sc["COMx"] + sc["COMoldx"], L)
    dy = wrap(o.y - o.yo - sc["COMy"] + sc["COMoldy"], L)
    c = np.sort(np.hypot(dx, dy))
    return c[-1] + c[-2]


def pair_dist(o, i, j, L, x=None, y=None):
    x = o.x if x is None else x
    y = o.y if y is None else y
    return np.hypot(wrap(x[j] - x[i], L), wrap(y[j] - y[i], L))


def class_for(D):      # apj_class_for (csrc/apj_device.cuh): lowest class that is safe to stop after
    k = 0
    for c in range(3):
        if D > (c + 1) * SKIN / 4 - EPS:
            k = c + 1
    return k


@pytest.mark.parametrize("l_s,l_n", [(0.5, 0.3), (0.1, 1.0)])
def test_pair_distance_never_drops_below_build_distance_minus_skin_value(l_s, l_n):
    N, rho, seed = 1024, 0.9, 17
    R, L, x, y, phi = random_system(N, rho, seed)
    o = OracleSim.from_arrays(R, x, y, phi, rho)
    o.topology(); o.assign(); o.build(); o.mark_origin()
    o.set_params(0.0, l_n); o.run_philox(seed, 0, 60)
    o.set_params(l_s, l_n)
    built = None
    classes_seen, rebuilds, worst = set(), 0, np.inf
    for t in range(60, 460):
        px, py = o.x.copy(), o.y.copy()
        r0 = o.scalars()["resetCounter"]
        o.run_philox(seed, t, 1)
        if o.scalars()["resetCounter"] != r0:                  # this step began with newSkinList firing: lists were
            off, idx = o.verlet()                               # rebuilt from (px, py), and x_old = (px, py)
            i = np.repeat(np.arange(N), np.diff(off)); j = idx.astype(np.int64)
            built = (i, j, pair_dist(o, i, j, L, px, py))
            assert np.all(built[2] < RS)
            rebuilds += 1
        if built is None:
            continue
        i, j, d_b = built
        D = skin_value(o, L)                                    # what the NEXT step's newSkinList evaluates
        if D > SKIN:
            continue                                            # that step rebuilds before it sweeps
        d = pair_dist(o, i, j, L)
        worst = min(worst, float(np.min(d - (d_b - D))))
        k = class_for(D)
        classes_seen.add(k)
        if k < 3:                                               # entries beyond class k are out of reach this step
            beyond = d_b >= RN + (k + 1) * SKIN / 4
            assert not np.any(d[beyond] < RN), "an omitted list entry interacts"
    assert rebuilds >= 3 and worst >= -1e-9, (rebuilds, worst)
    assert classes_seen >= {0, 1, 2, 3}                         # the run visits every sweep length
    o.close()


def test_class_rule():
    q = SKIN / 4
    assert [class_for(v) for v in (0.0, q - 2 * EPS, q, 2 * q - 2 * EPS, 2 * q, 3 * q - 2 * EPS, 3 * q, SKIN)] == [0, 0, 1, 1, 2, 2, 3, 3]
