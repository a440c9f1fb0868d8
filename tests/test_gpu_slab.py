"""Slab decomposition of ONE periodic box (BASELINE config 4) against (i) the periodic
single-handle engine -- positions must be BIT-IDENTICAL, because Philox counters are keyed by particle
id and list order by global cell, so the decomposition can only change the rounding of COM -- and
(ii) the oracle (pair set bit-exact, injected single step <= 1e-12). Several ranks share the one GPU of
the test box (one host thread per rank, peers are plain device pointers); the exchange code is the
same peer-store / peer-atomic / flag path that runs over NVLink between GPUs."""
import numpy as np
import pytest

from _util import DEV2ORC, device_from_state, random_system, rel_err, wrapped_abs_diff
from oracle.pyoracle import PI, OracleSim
from test_gpu_parity import relaxed_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-12


def slab_from_state(s, nranks, seed=12345, **kw):
    from active_particle_jamming_b200.slab import SlabBox
    box = SlabBox(len(s["x"]), s["L"], nranks, seed=seed, **kw)
    box.set_activity(s["CFself"], s["CTnoise"])
    box.upload(**{d: s[o] for d, o in DEV2ORC.items()})
    box.set_com(com=[s["COMx"], s["COMy"]], com0=[s["COM0x"], s["COM0y"]], com_old=[s["COMoldx"], s["COMoldy"]])
    box.set_reset_counter(int(s["resetCounter"]))
    return box


@pytest.mark.parametrize("nranks,lanes", [(1, 4), (2, 4), (3, 1), (4, 2)])
def test_slab_pairs_and_single_step_match_oracle(nranks, lanes):
    """Gates 1 and 2 of north_star on the decomposed box."""
    N, rho = 4096, 0.9
    o, rng = relaxed_oracle(N, rho, seed=300 + nranks, l_s=0.5, l_n=0.3)
    L = o.scalars()["L"]
    box = slab_from_state(o.state(), nranks, lanes_per_particle=lanes)
    try:
        assert sum(box.n_own()) == N
        o.assign(); o.build()
        assert np.array_equal(box.pair_set(), o.pair_set())
        for k in range(3):
            # the first step starts from identical state: 1e-12 (gate 2). The next two run on without
            # re-synchronising; last-bit differences grow ~1 decade per 25 steps (DESIGN.md "Parity")
            tol = TOL if k == 0 else 1e-10
            nz = rng.uniform(-PI, PI, N)
            o.step(nz)
            box.step_injected(nz)
            d = box.download()
            assert np.max(wrapped_abs_diff(d["x"], o.x, L)) <= tol * L and np.max(wrapped_abs_diff(d["y"], o.y, L)) <= tol * L
            assert rel_err(d["cosp"], o.cosp) <= tol and rel_err(d["sinp"], o.sinp) <= tol
            assert rel_err(d["vx"], o.vx, floor=1e-2) <= tol and rel_err(d["x_real"], o.xr, floor=L) <= tol
            sc, c = o.scalars(), box.get_com()
            assert abs(c["COM"][0] - sc["COMx"]) <= tol * L and abs(c["COM"][1] - sc["COMy"]) <= tol * L
            assert box.counters()["resetCounter"] == sc["resetCounter"]
    finally:
        box.close()
        o.close()


@pytest.mark.parametrize("nranks,lanes,N,flags", [(1, 4, 4096, 0), (2, 4, 4096, 0), (3, 2, 6000, 0), (4, 1, 20000, 0), (8, 1, 40000, 0),
                                                  (3, 1, 20000, 2), (2, 1, 20000, 2 | 8 | 16)])   # flags 2: split tail (fold + commit kernel) on the slabs; 8: persistent kernel on 3 blocks per rank
def test_slab_free_running_bit_identical_to_periodic(nranks, lanes, N, flags):
    """300 Philox steps with frequent rebuilds and migration across slab edges (lambda_s = 0.5): every
    rank count gives the bits of the single-GPU periodic engine."""
    rho, seed = 0.9, 77
    o, _ = relaxed_oracle(N, rho, seed=seed + nranks, l_s=0.5, l_n=0.3, presteps=40)
    s = o.state()
    L = s["L"]
    o.close()
    # 40 relaxation steps leave a few dense clusters of the random start: lists of up to ~50 entries
    e = device_from_state(s, seed=seed, lanes_per_particle=lanes, max_neighbors=64)
    box = slab_from_state(s, nranks, seed=seed, lanes_per_particle=lanes, max_neighbors=64, flags=flags)
    try:
        own0 = box.n_own()
        moved = 0
        for chunk in (1, 50, 120, 129):
            e.step(chunk); box.step(chunk)
            ce, cb = e.counters(), box.counters()
            assert cb["step"] == ce["step"] and cb["resetCounter"] == ce["resetCounter"]
            own = box.n_own()
            assert sum(own) == N
            moved += int(own != own0)
            own0 = own
            a, b = e.download(), box.download()
            for f in ("x", "y", "cosp", "sinp", "x_real", "y_real", "x_old", "y_old", "x0", "y0", "R"):
                assert np.array_equal(a[f], b[f]), "%s differs after %d steps" % (f, ce["step"])
            assert np.array_equal(a["box"], b["box"])
            assert np.array_equal(e.pair_set(), box.pair_set())
        assert ce["resetCounter"] >= 3, "the run was meant to cross several rebuilds"
        if nranks > 1:
            assert moved > 0, "no particle ever changed rank: migration untested"
        assert box.checksum() == e.checksum()                                # the fingerprint bench.py prints per --gpus N
        # observables: additive shares of the ranks against the periodic engine
        assert abs(box.order_orientation()[0][0] - e.order_orientation()[0][0]) <= TOL
        assert abs(box.msd()[0] - e.msd()[0]) <= TOL * max(1.0, e.msd()[0])
        assert np.array_equal(box.occupancy_hist(), e.occupancy_hist())
        assert abs(box.list_stats()[0] - e.list_stats()[0]) <= 1e-12
    finally:
        e.close()
        box.close()


def test_slab_mark_origin_and_relax_ramp():
    """start() :191-203 and the relax() ramp on a decomposed box."""
    N, rho, seed = 8192, 0.9, 5
    R, L, x, y, phi = random_system(N, rho, seed)
    from active_particle_jamming_b200 import DeviceEngine
    from active_particle_jamming_b200.slab import SlabBox
    e = DeviceEngine(N, L, seed=seed, lanes_per_particle=2)
    box = SlabBox(N, L, 3, seed=seed, lanes_per_particle=2)
    try:
        for q in (e, box):
            q.set_activity(0.0, 0.5)
            q.upload(x=x, y=y, R=R, phi=phi)
            q.step(60)
            q.set_activity(0.3, 0.5)
            q.set_ramp(50)
            q.step(80)
            q.mark_origin()
            q.step(40)
        a, b = e.download(), box.download()
        for f in ("x", "y", "cosp", "x0", "y0", "x_real"):
            assert np.array_equal(a[f], b[f]), f
        assert abs(box.msd()[0] - e.msd()[0]) <= TOL * max(1.0, e.msd()[0])
        ca, cb = e.get_com(), box.get_com()
        assert np.max(np.abs(ca["COM0"] - cb["COM0"])) <= TOL * L
    finally:
        e.close()
        box.close()


@pytest.mark.parametrize("nranks,cutoff", [(2, 20.0), (3, 20.0), (4, 60.0)])
def test_slab_spatial_correlations_match_periodic(nranks, cutoff):
    """Correlations::spatialCorrelations on the decomposed box: own pairs + the pairs that cross a rank's
    right edge (edge particles of the neighbour shipped by the host) add up to the periodic handle's raw sums --
    the pair COUNTS exactly, the fp64 sums to summation-order rounding."""
    N, rho = 65536, 0.9
    o, _ = relaxed_oracle(N, rho, seed=90 + nranks, l_s=0.1, l_n=0.4, presteps=8)
    s = o.state()
    o.close()
    with device_from_state(s, seed=3, lanes_per_particle=1, max_neighbors=64) as e:
        want = e.spatial_correlations(cutoff)
    box = slab_from_state(s, nranks, seed=3, lanes_per_particle=1, max_neighbors=64)
    try:
        got = box.spatial_correlations(cutoff)
        assert np.array_equal(got["counts"], want["counts"])
        for f in ("ori_sum", "vel_sum", "pair_sum"):
            assert rel_err(got[f], want[f], floor=1.0) <= 1e-10, f
        # ... and in the middle of a run: particles have drifted (some over a slab edge, still owned by the old rank), the
        # cells are those of the last rebuild
        with device_from_state(s, seed=3, lanes_per_particle=1, max_neighbors=64) as e:
            e.step(57); box.step(57)
            want, got = e.spatial_correlations(cutoff), box.spatial_correlations(cutoff)
        assert np.array_equal(got["counts"], want["counts"])
        for f in ("ori_sum", "vel_sum", "pair_sum"):
            assert rel_err(got[f], want[f], floor=1.0) <= 1e-10, f
        from active_particle_jamming_b200 import ApjError
        with pytest.raises(ApjError):
            box.spatial_correlations(400.0)                     # slabs narrower than the cutoff: refused, not miscounted
    finally:
        box.close()


def test_slab_checkpoint_per_rank_files(tmp_path):
    """apj_save_checkpoint / apj_load_checkpoint on slab handles: every rank writes and reads its own share (no gather
    of the box). The restored box holds the stored bits, continues the Philox stream at the stored step and takes
    its next step within 1e-12 of the uninterrupted run."""
    from active_particle_jamming_b200.slab import SlabBox
    from active_particle_jamming_b200 import ApjError
    N, rho, nranks = 12000, 0.9, 3
    o, _ = relaxed_oracle(N, rho, seed=71, l_s=0.3, l_n=0.4, presteps=30)
    s = o.state()
    L = s["L"]
    o.close()
    a = slab_from_state(s, nranks, seed=9, lanes_per_particle=1)
    b = SlabBox(N, L, nranks, seed=1, lanes_per_particle=1)            # another seed: the key travels with the file
    try:
        a.step(70)
        prefix = str(tmp_path / "box")
        a.save_checkpoint(prefix)
        import os
        assert sorted(os.listdir(tmp_path)) == ["box.rank%dof%d" % (r, nranks) for r in range(nranks)]
        b.load_checkpoint(prefix)
        da, db = a.download(), b.download()
        for f in ("x", "y", "cosp", "sinp", "x_real", "y_real", "x0", "y0", "x_old", "y_old", "R", "vx", "vy", "phi"):
            assert np.array_equal(da[f], db[f]), f
        ca, cb = a.counters(), b.counters()
        assert ca["step"] == cb["step"] == 70 and ca["resetCounter"] == cb["resetCounter"]
        assert np.array_equal(a.get_com()["COM"], b.get_com()["COM"]) and np.array_equal(a.get_com()["COM_old"], b.get_com()["COM_old"])
        assert a.checksum() == b.checksum()
        a.step(1); b.step(1)
        da, db = a.download(), b.download()
        assert np.max(wrapped_abs_diff(da["x"], db["x"], L)) <= TOL * L and rel_err(da["cosp"], db["cosp"]) <= TOL
        a.step(60); b.step(60)
        assert a.counters()["resetCounter"] == b.counters()["resetCounter"]
        # a file of another rank is refused
        with pytest.raises(ApjError):
            b.local[0].load_checkpoint(prefix + ".rank1of%d" % nranks)
    finally:
        a.close()
        b.close()


def test_slab_capacity_overflow_is_loud():
    from active_particle_jamming_b200 import ApjError
    from active_particle_jamming_b200.slab import SlabBox
    N, rho = 4096, 0.9
    R, L, x, y, phi = random_system(N, rho, 3)
    box = SlabBox(N, L, 2, capacity=N // 2 - 200)
    try:
        with pytest.raises(ApjError):
            box.upload(x=x, y=y, R=R, phi=phi)
    finally:
        box.close()


def test_slab_across_gpus_matches_periodic():
    """Two or more real GPUs (one process per GPU, cudaIpc peer mappings over NVLink): skipped on a
    single-GPU box, where the same code paths run with several ranks on one device (tests above)."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 8)),
                        "--master-addr", "127.0.0.1", "--master-port", "29641", os.path.join(root, "scripts", "slab_dist_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
