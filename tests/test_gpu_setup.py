"""Device-side setup (Engine::initCells / Engine::topology as kernels, jamming.cpp:285-410) and the Ovito overlap
hue (Cell::over, :653-656), through the C ABI. The lattice draws from the Philox stream (the reference's stream is
time-seeded: its initial condition can only be matched in distribution), so the checks are the recipe's invariants
plus exact agreement of everything derived from the drawn state with the oracle."""
import numpy as np
import pytest

from _util import device_from_state
from oracle.pyoracle import PI, OracleSim
from test_gpu_parity import relaxed_oracle

pytestmark = pytest.mark.gpu


def test_init_lattice_follows_the_reference_recipe():
    from active_particle_jamming_b200 import DeviceEngine
    N, rho, seed = 10000, 0.9, 77
    L = DeviceEngine.lattice_box_length(N, rho, seed)[0]
    with DeviceEngine(N, L, seed=5) as e:
        e.init_lattice(seed)
        d = e.download()
        # radii 1 + N(0,1)/10 (:296) and the box they imply (:305), summed in another order than the kernel
        assert abs(d["R"].mean() - 1.0) < 4 * 0.1 / np.sqrt(N) and abs(d["R"].std() - 0.1) < 0.004
        assert abs(L - np.sqrt(3.14159265 * np.sum(d["R"] ** 2) / rho)) <= 1e-11 * L
        # jittered offset-row lattice (:311-325): spacing L / floor(sqrt N), even rows shifted by 1, jitter N(0,1)/10
        rootN = int(np.sqrt(N))
        i = np.arange(N)
        x0 = -L / 2 + (L / rootN) * (i % rootN) + np.where((i // rootN) % 2 == 0, 1.0, 0.0)
        y0 = -L / 2 + (L / rootN) * (i // rootN)
        jx = (d["x"] - x0 + L / 2) % L - L / 2
        jy = (d["y"] - y0 + L / 2) % L - L / 2
        assert abs(jx.std() - 0.1) < 0.004 and abs(jy.std() - 0.1) < 0.004 and abs(jx.mean()) < 0.005 and abs(jy.mean()) < 0.005
        assert abs(np.corrcoef(jx, jy)[0, 1]) < 0.04 and abs(np.corrcoef(jx, d["R"])[0, 1]) < 0.04
        assert np.all(d["x"] >= -L / 2) and np.all(d["x"] < L / 2) and np.all(d["y"] >= -L / 2) and np.all(d["y"] < L / 2)
        # polarity U[-PI, PI) on the 2^-32 lattice, cos / sin of it (:327-329)
        assert d["phi"].min() >= -PI and d["phi"].max() < PI and abs(d["phi"].mean()) < 0.08 and abs(d["phi"].std() - 2 * PI / np.sqrt(12)) < 0.04
        assert np.max(np.abs(d["cosp"] - np.cos(d["phi"]))) <= 4e-16 and np.max(np.abs(d["sinp"] - np.sin(d["phi"]))) <= 4e-16
        for f, g in (("x_real", "x"), ("x0", "x"), ("x_old", "x"), ("y_real", "y"), ("y0", "y"), ("y_old", "y")):
            assert np.array_equal(d[f], d[g])
        assert not d["vx"].any() and not d["vy"].any()
        c = e.get_com(0)
        assert abs(c["COM"][0] - d["x"].mean()) <= 1e-11 * L and np.array_equal(c["COM"], c["COM0"]) and np.array_equal(c["COM"], c["COM_old"])
        # binning and lists were built from that state: exactly the oracle's
        o = OracleSim.from_arrays(d["R"], d["x"], d["y"], d["phi"], rho, L=L)
        o.topology(); o.assign(); o.build()
        assert np.array_equal(e.download(["box"])["box"], o.box)
        assert np.array_equal(e.pair_set(), o.pair_set())
        o.close()
        e.set_activity(0.1, 0.3); e.skip_self_term_once(); e.step(50)       # and it runs
        assert e.counters()["step"] == 50
        first = d
    # the same seed gives the same bits; another seed another state; replicas draw independently
    with DeviceEngine(N, L, seed=9) as e:
        e.init_lattice(seed)
        assert np.array_equal(e.download(["x"])["x"], first["x"]) and np.array_equal(e.download(["R"])["R"], first["R"])
    Ls = DeviceEngine.lattice_box_length(2048, [0.84, 1.0], 3, n_systems=2)
    assert Ls[0] > Ls[1]
    with DeviceEngine(2048, Ls, n_systems=2, seed=1) as e:
        e.init_lattice(3)
        d = e.download(["R", "x"])
        assert not np.array_equal(d["R"][:2048], d["R"][2048:])
        assert abs(Ls[1] - np.sqrt(3.14159265 * np.sum(d["R"][2048:] ** 2) / 1.0)) <= 1e-11 * Ls[1]


def test_box_table_equals_reference_topology():
    """Engine::topology (:356-410) as a kernel: centres and the 3x3 periodic neighbour table, reference numbering."""
    o, _ = relaxed_oracle(3000, 0.9, seed=2, presteps=4)
    s = o.state()
    with device_from_state(s) as e:
        centres, nbrs = e.box_table()
        assert np.array_equal(nbrs, o.box_neighbors().reshape(-1, 9))
        b = e.geometry()["b"]
        L, Lh = s["L"], s["L"] / 2
        i = np.arange(b)
        c1 = ((-Lh + i * L / b) + (-Lh + (i + 1) * L / b)) / 2.            # jamming.cpp:379-385, same rounding
        assert np.array_equal(centres[:, 0], np.tile(c1, b)) and np.array_equal(centres[:, 1], np.repeat(c1, b))
    o.close()


@pytest.mark.parametrize("N,rho,lanes", [(4096, 1.0, 4), (1500, 0.9, 1), (100, 1.1, 8)])
def test_overlap_hue_is_the_references_integer(N, rho, lanes):
    """Cell::over of print_video (:653-656, :855-870): 240 minus 240 |overlap| per overlapping pair, accumulated
    into an int with truncation at every update, in the reference's visiting order -- bit for bit (the oracle's
    restatement is pinned to the compiled reference's neighborInteractions in test_oracle_vs_reference.py)."""
    o, _ = relaxed_oracle(N, rho, seed=50 + lanes, l_s=0.3, l_n=0.3, presteps=25)
    o.assign(); o.build()
    want = o.overlap_hue()
    with device_from_state(o.state(), lanes_per_particle=lanes) as e:
        got = e.overlap_hue()
    o.close()
    assert np.array_equal(got, want)
    assert want.min() < 236 and want.max() <= 240        # above jamming every particle overlaps; the int may go negative
