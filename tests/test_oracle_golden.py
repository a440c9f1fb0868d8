"""Pins the oracle (oracle/apj_oracle.c) to outputs of the REFERENCE ITSELF: tests/golden/*.npz were
written by tests/golden/make_golden.py from oracle/_ref (the reference's own jamming.cpp + classes/*.h,
NDIM=2). Integer/index work and -- because the oracle follows the reference's operation order and is
built with the same flags -- all fp64 state must be BIT-IDENTICAL. Runs anywhere (no GPU, no /root/reference)."""
import numpy as np
import pytest

from _util import GOLDEN_CASES, STATE, golden_state, half_pairs, load_golden, oracle_from_state


@pytest.fixture(scope="module", params=GOLDEN_CASES)
def gold(request):
    return load_golden(request.param)


def test_topology_matches_reference(gold):
    s0 = golden_state(gold, "s0_")
    o = oracle_from_state(s0, gold["rho"])
    sc = o.scalars()
    assert sc["b"] == s0["b"] and sc["nbox"] == s0["nbox"]
    assert sc["lp"] == s0["lp"] and sc["Lover2"] == s0["Lover2"]         # bit-exact (jamming.cpp:361-365)
    assert np.array_equal(o.box_neighbors(), gold["box_neighbors"])        # jamming.cpp:391-408
    o.close()


@pytest.mark.parametrize("bruteforce", [False, True])
def test_assign_cells_matches_reference(gold, bruteforce):
    s0 = golden_state(gold, "s0_")
    o = oracle_from_state(s0, gold["rho"])
    o.box[:] = -1
    o.assign(bruteforce=bruteforce)                                         # jamming.cpp:527-548 (Q8)
    assert np.array_equal(o.box, s0["box"])
    off, idx = o.cell_lists()
    assert np.array_equal(off, gold["s0_cl_off"]) and np.array_equal(idx, gold["s0_cl_idx"])
    o.close()


def test_verlet_lists_match_reference(gold):
    s0 = golden_state(gold, "s0_")
    o = oracle_from_state(s0, gold["rho"])
    o.assign()
    o.build()                                                               # jamming.cpp:550-585 (Q1: half list)
    off, idx = o.verlet()
    assert np.array_equal(off, gold["s0_vl_off"])
    assert np.array_equal(idx, gold["s0_vl_idx"])                          # same order, not just the same set
    i = np.repeat(np.arange(gold["N"]), np.diff(off))
    assert np.all(idx > i)
    o.close()


def test_injected_steps_bit_identical(gold):
    """K x Engine::calculate_next_positions with randuni() injected (jamming.cpp:837-853, :667)."""
    s0 = golden_state(gold, "s0_")
    o = oracle_from_state(s0, gold["rho"])
    o.assign()
    o.build()
    cks = set(int(c) for c in gold["checkpoints"])
    for k in range(gold["K"]):
        o.step(gold["noise"][k], fast_assign=(k % 2 == 0))                 # both binning forms along the way
        assert o.scalars()["resetCounter"] == gold["resets"][k], "rebuild decision differs at step %d" % k
        if k + 1 in cks:
            ref = golden_state(gold, "s%d_" % (k + 1))
            for f in STATE:
                assert np.array_equal(getattr(o, f), ref[f]), "field %s differs after step %d" % (f, k + 1)
            assert np.array_equal(o.box, ref["box"])
            sc = o.scalars()
            for name in ("COMx", "COMy", "COMoldx", "COMoldy", "COM0x", "COM0y"):
                assert sc[name] == ref[name], name
    off, idx = o.verlet()
    assert np.array_equal(off, gold["end_vl_off"]) and np.array_equal(idx, gold["end_vl_idx"])
    o.close()


def _end_state(gold):
    s = golden_state(gold, "s%d_" % gold["K"])
    return oracle_from_state(s, gold["rho"])


def test_engine_observables_match_reference(gold):
    o = _end_state(gold)
    assert o.order() == float(gold["order"])                               # jamming.cpp:776-789
    assert np.array_equal(o.orientation(), gold["orientation"])            # :791-806
    assert o.msd() == float(gold["msd"])                                   # :808-823
    o.close()


def test_fluctuations_match_reference(gold):
    o = _end_state(gold)
    from oracle.pyoracle import OracleSim
    got = [OracleSim.lib().orc_fluct_overlap(*a) for a in gold["fluct_overlap_args"]]
    assert np.array_equal(got, gold["fluct_overlap"])                      # Fluctuations.h:89-120
    f = o.fluct_init(gold["steps"], 10, gold["rho"])                       # Engine::fluct_int = 10 (jamming.cpp:60)
    assert f.time_interval == float(gold["fluct_time_interval"]) and f.rad_interval == float(gold["fluct_rad_interval"])
    for row in gold["fluct_seq"]:                                          # state machine incl. the flush call (Q13)
        o.fluct_measure(f)
        assert (f.current_radius, f.current_value, f.counter) == (row[0], row[1], int(row[2]))
    o.close()


def test_correlations_match_reference(gold):
    o = _end_state(gold)
    o.assign()
    off, idx = o.cell_lists()
    assert np.array_equal(off, gold["end_cl_off"]) and np.array_equal(idx, gold["end_cl_idx"])
    nc, npb = int(gold["corr_dims"][0]), int(gold["corr_dims"][1])
    cutoff = 20.0                                                          # local build: jamming.cpp:153
    assert (nc, npb) == (int(np.ceil(cutoff / 2.0)), int(np.ceil(cutoff / 0.1)))
    vel, ori, pair = o.spatial_correlations(cutoff)                        # Correlations.h:71-167
    assert np.array_equal(vel, gold["corr_vel"], equal_nan=True)
    assert np.array_equal(ori, gold["corr_ori"], equal_nan=True)
    assert np.array_equal(pair, gold["corr_pair"])
    assert np.array_equal(o.vel_dist(gold["l_s"]), gold["vel_dist"])        # Correlations.h:179-187
    assert np.array_equal(o.density_distribution(), gold["dens_dist"])     # Fluctuations.h:122-139
    o.close()


def test_pair_set_is_complete_bruteforce(gold):
    """The half lists hold exactly {(i<j): d2 < rs2} -- checked against an O(N^2) numpy sweep."""
    s0 = golden_state(gold, "s0_")
    L = s0["L"]
    dx = s0["x"][None, :] - s0["x"][:, None]
    dy = s0["y"][None, :] - s0["y"][:, None]
    dx = np.where(dx < -L / 2, dx + L, np.where(dx >= L / 2, dx - L, dx))
    dy = np.where(dy < -L / 2, dy + L, np.where(dy >= L / 2, dy - L, dy))
    d2 = dx * dx + dy * dy
    i, j = np.nonzero(np.triu(d2 < 4.2 * 4.2, k=1))
    brute = np.stack([i, j], axis=1)
    assert np.array_equal(half_pairs(gold["s0_vl_off"], gold["s0_vl_idx"]), brute)
