"""CUDA path vs the oracle (oracle/apj_oracle.c, itself pinned bit-for-bit to the reference) on
seeded synthetic systems, every call through the C ABI. Small sizes compare full state; the sizes
BASELINE.json names (65 536 at phi = 1.0, 1 048 576 at phi = 0.9) are compared directly too -- the
O(N) oracle finishes them in seconds -- plus size-independent properties (Newton's third law,
COM consistency, determinism, in-box positions)."""
import numpy as np
import pytest

from _util import DEV2ORC, device_from_state, random_system, rel_err, wrapped_abs_diff
from oracle.pyoracle import PI, OracleSim

pytestmark = pytest.mark.gpu
TOL = 1e-12


def relaxed_oracle(N, rho, seed, l_s=0.05, l_n=0.5, presteps=60):
    """Random synthetic configuration pushed through `presteps` oracle steps (the first
    ones at CFself = 0, like relax()) so overlaps are physical and xnew = cosp holds."""
    R, L, x, y, phi = random_system(N, rho, seed)
    o = OracleSim.from_arrays(R, x, y, phi, rho)
    o.topology()
    o.assign()
    o.build()
    o.mark_origin()
    rng = np.random.default_rng(seed + 1)
    o.set_params(0.0, l_n)
    for _ in range(presteps // 2):
        o.step(rng.uniform(-PI, PI, N))
    o.set_params(l_s, l_n)
    for _ in range(presteps - presteps // 2):
        o.step(rng.uniform(-PI, PI, N))
    o.mark_origin()
    return o, rng


def assert_state_close(d, o, L, tol, what=""):
    assert np.max(wrapped_abs_diff(d["x"], o.x, L)) <= tol * L, what + " x"
    assert np.max(wrapped_abs_diff(d["y"], o.y, L)) <= tol * L, what + " y"
    for f in ("x_real", "y_real", "x_old", "y_old", "x0", "y0"):
        assert rel_err(d[f], getattr(o, DEV2ORC[f]), floor=L) <= tol, what + " " + f
    assert rel_err(d["cosp"], o.cosp) <= tol and rel_err(d["sinp"], o.sinp) <= tol, what + " cos/sin"
    dphi = np.abs(d["phi"] - o.phi)
    assert np.max(np.minimum(dphi, np.abs(6.28318531 - dphi))) <= tol * PI, what + " phi"
    assert rel_err(d["vx"], o.vx, floor=1e-2) <= tol and rel_err(d["vy"], o.vy, floor=1e-2) <= tol, what + " v"


@pytest.mark.parametrize("lanes", [1, 2, 4, 8])
@pytest.mark.parametrize("N,rho", [(4096, 0.9), (1500, 1.0)])
def test_step_by_step_resynchronised(N, rho, lanes):
    """Every step starts from the oracle's state: pair sets bit-exact after each rebuild, each single
    step within 1e-12 (gate 2 of north_star), for every lanes-per-particle variant of the kernel."""
    o, rng = relaxed_oracle(N, rho, seed=N + lanes, l_s=0.5, l_n=0.3)
    L = o.scalars()["L"]
    e = device_from_state(o.state(), lanes_per_particle=lanes)
    try:
        o.assign(); o.build()
        assert np.array_equal(e.pair_set(), o.pair_set())
        for k in range(12):
            e.close()
            e = device_from_state(o.state(), lanes_per_particle=lanes)
            nz = rng.uniform(-PI, PI, N)
            rebuilt = o.step(nz)
            e.step_injected(nz)
            assert_state_close(e.download(), o, L, TOL, "step %d" % k)
            assert e.counters()["resetCounter"] == o.scalars()["resetCounter"]
            if rebuilt:
                assert np.array_equal(e.pair_set(), o.pair_set())
                assert np.array_equal(e.download(["box"])["box"], o.box)
            sc, c = o.scalars(), e.get_com()
            assert abs(c["COM"][0] - sc["COMx"]) <= TOL * L and abs(c["COM"][1] - sc["COMy"]) <= TOL * L
    finally:
        e.close()
        o.close()


@pytest.mark.parametrize("lanes", [1, 4, 8])
@pytest.mark.parametrize("N", [100, 144, 256])               # b = 3, 3-4, 5 boxes per side (rn is the largest radius drawn): the smallest grids the reference is defined on
def test_smallest_boxes(N, lanes):
    """b = floor(L / 2 rn) = 3 is the smallest grid whose 3x3 neighbourhood does not alias (jamming.cpp:359; apj_create
    refuses b < 3): one tile column wraps onto itself through the periodic seam on both sides. Step-by-step and
    200 free-running Philox steps against the oracle."""
    rho = 0.9
    o, rng = relaxed_oracle(N, rho, seed=N + lanes, l_s=0.5, l_n=0.3)
    L = o.scalars()["L"]
    assert 3 <= o.scalars()["b"] <= (3 if N == 100 else 5)
    e = device_from_state(o.state(), lanes_per_particle=lanes, seed=5)
    try:
        o.assign(); o.build()
        assert np.array_equal(e.pair_set(), o.pair_set())
        for k in range(8):
            nz = rng.uniform(-PI, PI, N)
            rebuilt = o.step(nz)
            e.step_injected(nz)
            if k == 0:
                assert_state_close(e.download(), o, L, TOL, "step 0")
            if rebuilt:
                assert np.array_equal(e.pair_set(), o.pair_set())
        e.close()
        e = device_from_state(o.state(), lanes_per_particle=lanes, seed=5)
        o.run_philox(5, 0, 1); e.step(1)
        assert_state_close(e.download(), o, L, TOL, "philox step 1")
        nreb = o.run_philox(5, 1, 199); e.step(199)
        c = e.counters()
        assert c["step"] == 200 and nreb > 0
        d = e.download()
        assert np.all((d["x"] >= -L / 2) & (d["x"] < L / 2) & (d["y"] >= -L / 2) & (d["y"] < L / 2))
        o2 = OracleSim.from_arrays(d["R"], d["x"], d["y"], d["phi"], rho, L=L)   # a fresh build on the device's own state gives the oracle's pairs
        o2.topology(); o2.assign(); o2.build()
        e.force_rebuild()
        assert np.array_equal(e.pair_set(), o2.pair_set())
        o2.close()
    finally:
        e.close()
        o.close()


@pytest.mark.parametrize("flags", [0, 1])   # CUDA graph / direct launches
def test_free_running_philox_matches_oracle(flags):
    """apj_step (Philox4x32-10 noise generated in the kernel, speculative multi-step launches with the
    on-device rebuild) against the oracle drawing the same counter-based stream on the CPU."""
    N, rho, seed = 4096, 0.9, 99
    o, _ = relaxed_oracle(N, rho, seed=3, l_s=0.3, l_n=0.5)
    L = o.scalars()["L"]
    with device_from_state(o.state(), seed=seed, flags=flags) as e:
        first = e.counters()["step"]
        assert first == 0
        o.run_philox(seed, 0, 1)
        e.step(1)
        assert_state_close(e.download(), o, L, TOL, "philox step 1")        # same noise bits -> 1e-12
        # 99 more steps (crossing several speculative launch groups and on-device rebuilds). Last-bit
        # differences grow ~1 decade per 25 steps (chaotic orientation dynamics), hence 1e-6 here.
        nreb = o.run_philox(seed, 1, 99)
        e.step(99)
        c = e.counters()
        assert c["step"] == 100
        assert c["resetCounter"] == o.scalars()["resetCounter"] and nreb > 0  # rebuilds on the same steps
        assert_state_close(e.download(), o, L, 1e-6, "philox step 100")
        assert np.array_equal(e.pair_set(), o.pair_set())
    o.close()


def test_replicas_equal_single_systems():
    """n_systems > 1 (phase-diagram sweep batching): every replica evolves exactly like the same
    system alone -- bit-identical, the block decomposition and reduction order are per system."""
    N = 2048
    specs = [(0.84, 0.1, 0.2, 5), (0.9, 0.5, 0.5, 6), (1.0, 1.0, 0.05, 7)]
    os_ = [relaxed_oracle(N, rho, seed=sd, l_s=ls, l_n=ln)[0] for rho, ls, ln, sd in specs]
    from active_particle_jamming_b200 import DeviceEngine
    Ls = [o.scalars()["L"] for o in os_]
    rng = np.random.default_rng(0)
    noise = rng.uniform(-PI, PI, (20, len(specs), N))
    batch = DeviceEngine(N, Ls, n_systems=len(specs))
    batch.set_activity([s[1] for s in specs], [s[2] for s in specs])
    st = [o.state() for o in os_]
    batch.upload(box=np.concatenate([s["box"] for s in st]), **{d: np.concatenate([s[k] for s in st]) for d, k in DEV2ORC.items()})
    for i, s in enumerate(st):
        batch.set_com(i, com=[s["COMx"], s["COMy"]], com0=[s["COM0x"], s["COM0y"]], com_old=[s["COMoldx"], s["COMoldy"]])
        batch.set_reset_counter(int(s["resetCounter"]), i)
    for k in range(20):
        batch.step_injected(noise[k].ravel())
    db = batch.download()
    for i, s in enumerate(st):
        with device_from_state(s) as single:
            for k in range(20):
                single.step_injected(noise[k, i])
            ds = single.download()
            for f in ds:
                assert np.array_equal(ds[f], db[f][i * N:(i + 1) * N]), (i, f)
            assert single.counters()["resetCounter"] == batch.counters(i)["resetCounter"]
            assert np.array_equal(single.pair_set(), batch.pair_set(i))
    o_ord, o_vec = batch.order_orientation()
    for i, o in enumerate(os_):
        for k in range(20):
            o.step(noise[k, i])
        assert abs(o_ord[i] - o.order()) <= 1e-7 and np.max(np.abs(o_vec[i] - o.orientation())) <= 1e-7
        o.close()
    batch.close()


def test_relax_schedule_ramp_and_first_step():
    """relax() (jamming.cpp:482-525): CFself = 0 phase, then the linear ramp (:518); and the very first
    step of a fresh Engine, whose alignment sum has no self term (Cell.h:75)."""
    N, rho = 1024, 0.9
    R, L, x, y, phi = random_system(N, rho, 21)
    o = OracleSim.from_arrays(R, x, y, phi, rho)
    o.topology(); o.assign(); o.build()
    from active_particle_jamming_b200 import DeviceEngine
    e = DeviceEngine(N, L, seed=5)
    # relax() starts with x_old = x_real = -100 (Cell.h:64-67): the first skin test always fires (Q7)
    e.upload(x=o.x, y=o.y, R=o.R, phi=o.phi, cosp=o.cosp, sinp=o.sinp, box=o.box, x_old=o.xo, y_old=o.yo,
             x_real=o.xr, y_real=o.yr, x0=o.x0, y0=o.y0)
    o.set_com([0, 0], [0, 0], [0, 0]); e.set_com(0, com=[0, 0], com0=[0, 0], com_old=[0, 0])
    e.skip_self_term_once()
    CF, trelax, ttherm = 0.4, 30, 40
    o.set_params(0.0, 0.5); e.set_activity(0.0, 0.5)
    o.run_philox(5, 0, trelax); e.step(trelax)
    assert_state_close(e.download(), o, L, 1e-7, "CFself=0 phase")
    e.set_activity(CF, 0.5); e.set_ramp(ttherm)
    for t_ in range(ttherm):
        o.set_params(CF - (ttherm - t_) * CF / ttherm, 0.5)                 # jamming.cpp:518
        o.run_philox(5, trelax + t_, 1)
    e.step(ttherm)
    assert_state_close(e.download(), o, L, 1e-6, "ramp phase")
    o.set_params(CF, 0.5)
    o.run_philox(5, trelax + ttherm, 5); e.step(5)
    assert_state_close(e.download(), o, L, 1e-6, "after ramp")
    assert e.counters()["resetCounter"] == o.scalars()["resetCounter"]
    e.close(); o.close()


def test_mark_origin_and_observables_vs_oracle():
    N, rho = 4096, 0.9
    o, rng = relaxed_oracle(N, rho, seed=17, l_s=0.4, l_n=0.2)
    with device_from_state(o.state()) as e:
        e.mark_origin(); o.mark_origin()                                    # start() :191-203
        c, sc = e.get_com(), o.scalars()
        L = sc["L"]
        assert abs(c["COM0"][0] - sc["COM0x"]) <= TOL * L and abs(c["COM_old"][1] - sc["COMoldy"]) <= TOL * L
        for k in range(40):
            nz = rng.uniform(-PI, PI, N)
            o.step(nz); e.step_injected(nz)
        order, orient = e.order_orientation()
        assert abs(order[0] - o.order()) <= 1e-7 and np.max(np.abs(orient[0] - o.orientation())) <= 1e-7
        assert rel_err(e.msd()[0], o.msd(), floor=1e-3) <= 1e-7
        # observables from IDENTICAL state: re-upload the oracle's state, then 1e-12
        with device_from_state(o.state()) as e2:
            order, orient = e2.order_orientation()
            assert abs(order[0] - o.order()) <= TOL and np.max(np.abs(orient[0] - o.orientation())) <= TOL
            assert rel_err(e2.msd()[0], o.msd(), floor=1e-3) <= TOL
            for radius in (3.0, 7.5, L / 4, L / 2):
                assert rel_err(e2.fluct_area(radius)[0], o.fluct_area(radius), floor=1.0) <= 1e-11
            o.assign()
            assert np.array_equal(e2.occupancy_hist()[0], o.density_distribution().astype(np.int64))
            assert np.array_equal(e2.vel_hist(0.4 / 50)[0], np.rint(o.vel_dist(0.4) * N).astype(np.int64))
            off_o, idx_o = o.cell_lists(); off_d, idx_d = e2.cell_lists()
            assert np.array_equal(off_o, off_d) and np.array_equal(idx_o, idx_d)


def test_error_paths():
    from active_particle_jamming_b200 import ApjError, DeviceEngine
    with pytest.raises(ApjError) as ei:
        DeviceEngine(64, 10.0)                                              # b = floor(10/5.6) = 1 < 3 (jamming.cpp:359)
    assert ei.value.code == -1
    with pytest.raises(ApjError) as ei:
        DeviceEngine(1024, 60.0, lanes_per_particle=3)
    assert ei.value.code == -1
    e = DeviceEngine(1024, 60.0)
    with pytest.raises(ApjError) as ei:
        e.step(1)                                                           # no state uploaded
    assert ei.value.code == -4
    with pytest.raises(ApjError):
        e.upload(x=np.zeros(1024), y=np.zeros(1024))                        # R, phi missing
    e.close()
    # list capacity beyond what the layout holds (96): everything in one spot
    R, L, x, y, phi = random_system(1024, 0.9, 3)
    e = DeviceEngine(1024, L, max_neighbors=8)
    with pytest.raises(ApjError) as ei:
        e.upload(x=x / 50, y=y / 50, R=R, phi=phi)
    assert ei.value.code == -3 and "max_neighbors" in str(ei.value)
    e.close()
    # non-positive radius: found by the pack kernel
    e = DeviceEngine(1024, L)
    Rb = R.copy(); Rb[17] = 0.0
    with pytest.raises(ApjError) as ei:
        e.upload(x=x, y=y, R=Rb, phi=phi)
    assert ei.value.code == -1 and "radii" in str(ei.value)
    e.close()
    with device_from_state(relaxed_oracle(1024, 0.9, 1)[0].state()) as e:
        with pytest.raises(ApjError) as ei:
            e.spatial_correlations(19.0)
        assert ei.value.code == -1


def test_list_capacity_grows_instead_of_truncating():
    """A list longer than max_neighbors must never be truncated: the system stays stale, the
    library re-allocates the list storage and re-runs the chain -- at upload and in the middle of apj_step
    (a clustering run) -- and the pair set / trajectory equal those of a handle created large enough."""
    N, rho = 4096, 0.9
    o, _ = relaxed_oracle(N, rho, seed=21, l_s=0.5, l_n=0.3, presteps=20)     # 20 steps from a random start: lists up to ~40
    s = o.state()
    o.assign(); o.build()
    ref_pairs = o.pair_set()
    o.close()
    with device_from_state(s, seed=5, max_neighbors=8) as small, device_from_state(s, seed=5, max_neighbors=96) as big:
        assert small.counters()["overflow"] == 0
        assert np.array_equal(small.pair_set(), ref_pairs) and np.array_equal(big.pair_set(), ref_pairs)
        small.step(200); big.step(200)
        a, b = small.download(), big.download()
        for f in ("x", "y", "cosp", "sinp", "x_real", "x_old"):
            assert np.array_equal(a[f], b[f]), f
        assert small.counters()["resetCounter"] == big.counters()["resetCounter"] >= 2
    # growth in the middle of a run: a passively relaxed (homogeneous) state has lists of <= ~21 entries; switching the
    # propulsion on makes the density fluctuate and the longest list grows to ~35 within 200 steps (measured on the oracle)
    o, _ = relaxed_oracle(N, rho, seed=9, l_s=0.0, l_n=0.3, presteps=100)
    s = o.state()
    o.close()
    with device_from_state(s, seed=3, max_neighbors=24) as e, device_from_state(s, seed=3, max_neighbors=96) as g:
        assert g.counters()["list_max"] <= 24
        for q in (e, g):
            q.set_activity(0.5, 0.3)
            q.step(400)
        assert g.counters()["list_max"] > 24, "the run was meant to outgrow the initial list capacity"
        a, b = e.download(), g.download()
        for f in ("x", "y", "cosp", "sinp", "x_old"):
            assert np.array_equal(a[f], b[f]), f
        assert e.counters()["resetCounter"] == g.counters()["resetCounter"] and e.counters()["overflow"] == 0


def DeviceEngine_(*a, **k):
    from active_particle_jamming_b200 import DeviceEngine
    return DeviceEngine(*a, **k)


def test_download_into_caller_buffers_and_checksum():
    """apj_download_state scatters by id on the device into caller-owned (here page-locked) arrays; the
    state fingerprint is a function of the state only (not of the memory order, which a rebuild changes)."""
    import torch
    N = 20000
    o, _ = relaxed_oracle(N, 0.9, seed=8, presteps=10)
    s = o.state()
    o.close()
    with device_from_state(s, seed=1) as e:
        bufs = {k: torch.empty(N, dtype=torch.float64, pin_memory=True).numpy() for k in ("x", "y", "cosp", "sinp", "R", "phi")}
        bufs["box"] = torch.empty(N, dtype=torch.int32, pin_memory=True).numpy()
        d = e.download(list(bufs), out=bufs)
        assert all(d[k] is bufs[k] for k in bufs)
        for dev, orc in (("x", "x"), ("y", "y"), ("cosp", "cosp"), ("sinp", "sinp"), ("R", "R"), ("phi", "phi")):
            assert np.array_equal(d[dev], s[orc]), dev                        # upload -> download is the identity
        c0 = e.checksum()
        e.force_rebuild()                                                     # permutes memory, not the state
        assert e.checksum() == c0
        e.step(3)
        assert e.checksum() != c0
        with pytest.raises(ValueError):
            e.download(["x"], out={"x": np.zeros(N - 1)})
    # phi only: cos/sin derived on the device (jamming.cpp:332-333), within an ulp of the host's
    with DeviceEngine_(N, s["L"]) as e:
        e.upload(x=s["x"], y=s["y"], R=s["R"], phi=s["phi"])
        d = e.download(["cosp", "sinp", "x_real", "x0", "vx"])
        assert np.max(np.abs(d["cosp"] - np.cos(s["phi"]))) <= 4e-16 and np.max(np.abs(d["sinp"] - np.sin(s["phi"]))) <= 4e-16
        assert np.array_equal(d["x_real"], s["x"]) and np.array_equal(d["x0"], s["x"]) and not d["vx"].any()
    # cos/sin only: phi derived
    with DeviceEngine_(N, s["L"]) as e:
        e.upload(x=s["x"], y=s["y"], R=s["R"], cosp=s["cosp"], sinp=s["sinp"])
        assert np.max(np.abs(e.download(["phi"])["phi"] - np.arctan2(s["sinp"], s["cosp"]))) <= 1e-15


def test_queued_observables_equal_the_synchronous_ones():
    """apj_obs_enqueue / apj_obs_fetch (reductions queued on the stream, read back in one copy) return the raw sums
    of the very kernels behind apj_order_orientation / apj_msd / apj_fluct_area / COM: identical bits, any number
    of tickets per fetch, across steps queued in between."""
    N = 8192
    o, _ = relaxed_oracle(N, 0.9, seed=14, l_s=0.3, l_n=0.4, presteps=30)
    s = o.state()
    o.close()
    with device_from_state(s, seed=2) as e:
        want, tickets = [], []
        for k in range(5):
            e.step(7)
            order, orient = e.order_orientation()
            want.append((order[0], orient[0].copy(), e.msd()[0], e.fluct_area(3.0 + k)[0], e.get_com(0)["COM"].copy()))
        e2 = device_from_state(s, seed=2)
        try:
            for k in range(5):
                e2.step(7)
                tickets += [e2.obs_enqueue(e2.OBS_ORDER), e2.obs_enqueue(e2.OBS_MSD), e2.obs_enqueue(e2.OBS_FLUCT, 3.0 + k),
                            e2.obs_enqueue(e2.OBS_COM)]
            assert tickets == list(range(tickets[0], tickets[0] + 20))
            raw = e2.obs_fetch(tickets[0], 20)[:, 0, :]
            for k in range(5):
                ox, oy = raw[4 * k]
                assert np.sqrt(ox * ox + oy * oy + 0.0 * 0.0) / N == want[k][0] and ox / N == want[k][1][0] and oy / N == want[k][1][1]
                assert raw[4 * k + 1][0] / N == want[k][2]
                assert raw[4 * k + 2][0] == want[k][3]
                assert np.max(np.abs(raw[4 * k + 3] / N - want[k][4])) <= 1e-12 * s["L"]   # engine COM: folded by the step kernel
            assert np.array_equal(e2.obs_fetch(tickets[7], 3)[:, 0, :], raw[7:10])       # any run of tickets
            with pytest.raises(Exception):
                e2.obs_fetch(tickets[-1] + 1, 1)                                          # not issued yet
        finally:
            e2.close()


def test_remote_settings_correlations_cutoff_140():
    """The reference's cluster build measures Correlations with cutoff = 140 (jamming.cpp:155-162): 70 correlation
    bins, 1400 g(r) bins, ~8700 partners per particle. Raw sums against the oracle's boxPairs form at N = 65 536."""
    N, rho = 65536, 0.9
    o, _ = relaxed_oracle(N, rho, seed=40, l_s=0.05, l_n=0.5, presteps=6)
    L = o.scalars()["L"]
    assert L / 2 > 140.0
    o.assign(); o.build()        # as start() does before every spatialCorrelations (jamming.cpp:245-246): the boxPairs walk needs fresh cell lists
    vel, ori, pair = o.spatial_correlations(140.0)
    with device_from_state(o.state(), seed=1) as e:
        c = e.spatial_correlations(140.0)
    o.close()
    assert c["counts"].shape == (1, 70) and c["pair_sum"].shape == (1, 1400)
    with np.errstate(invalid="ignore", divide="ignore"):
        gv, go = c["vel_sum"][0] / c["counts"][0], c["ori_sum"][0] / c["counts"][0]
    norm = 2 * L * L / (2 * 3.14159265 * 0.1 * float(N) * float(N))
    assert not np.isnan(gv).any() and c["counts"][0].sum() > 0.4 * N * 17000 / 2
    # the per-bin sums run over up to 1e7 pairs each, in different orders: 1e-10 of the O(1) correlation values
    assert rel_err(gv, vel, floor=1e-2) <= 1e-9 and rel_err(go, ori, floor=1e-2) <= 1e-9
    assert rel_err(c["pair_sum"][0] * norm, pair, floor=1e-3) <= 1e-10


def _pair_hash(off, idx, n, chunk=1 << 24):
    """Order-independent 64-bit hash of the half-list pair set {(i, j)}: sum of a mixed (i << 32 | j)."""
    rows = np.repeat(np.arange(n, dtype=np.uint64), np.diff(off))
    acc = np.uint64(0)
    with np.errstate(over="ignore"):
        for a in range(0, rows.size, chunk):
            z = (rows[a:a + chunk] << np.uint64(32)) | idx[a:a + chunk].astype(np.uint64)
            z ^= z >> np.uint64(30); z *= np.uint64(0xbf58476d1ce4e5b9)
            z ^= z >> np.uint64(27); z *= np.uint64(0x94d049bb133111eb)
            z ^= z >> np.uint64(31)
            acc += z.sum(dtype=np.uint64)
    return int(acc)


def test_headline_16m_box_parity():
    """The configuration the metric is quoted on (BASELINE.json configs[3], bench workload box16m): N = 16 777 216,
    phi = 0.9. Neighbour-pair set against the O(N) oracle -- per-particle half-list lengths array_equal, pair hash
    equal, total equal -- and one injected-noise step <= 1e-12 (gates 1 and 2 of north_star at full size)."""
    N, rho = 16777216, 0.9
    R, L, x, y, phi = random_system(N, rho, 2026)
    o = OracleSim.from_arrays(R, x, y, phi, rho)
    o.set_params(0.05, 0.5)
    o.topology(); o.assign(); o.build(); o.mark_origin()
    off_o, idx_o = o.verlet()
    from active_particle_jamming_b200 import DeviceEngine
    with DeviceEngine(N, L, seed=4, max_neighbors=64) as e:
        e.set_activity(0.05, 0.5)
        e.upload(x=o.x, y=o.y, R=o.R, phi=o.phi, cosp=o.cosp, sinp=o.sinp)
        e.skip_self_term_once()                                               # the oracle's first step has x_new = 0 (Cell.h:75,102)
        assert np.array_equal(e.download(["box"])["box"], o.box)
        off_d, idx_d = e.pair_list()
        assert off_d[-1] == off_o[-1] == idx_o.size
        assert np.array_equal(np.diff(off_d), np.diff(off_o))
        assert _pair_hash(off_d, idx_d, N) == _pair_hash(off_o, idx_o, N)
        del off_d, idx_d, off_o, idx_o
        nz = np.random.default_rng(7).uniform(-PI, PI, N)
        o.step(nz); e.step_injected(nz)
        assert_state_close(e.download(), o, L, TOL, "16M step")
        sc, c = o.scalars(), e.get_com()
        assert abs(c["COM"][0] - sc["COMx"]) <= TOL * L and abs(c["COM"][1] - sc["COMy"]) <= TOL * L
    o.close()


def test_inhomogeneous_density_grows_the_tile():
    """A clustered configuration (empty half box) forces tiles far from the mean occupancy: the
    shared-memory tile capacity must adapt, and the pair set must still be exact."""
    N, rho = 4096, 0.45
    R, L, x, y, phi = random_system(N, rho, 31)
    x = x / 2 - L / 4                                                       # everything in the left half
    o = OracleSim.from_arrays(R, x, y, phi, rho)
    o.topology(); o.assign(); o.build()
    from active_particle_jamming_b200 import DeviceEngine
    with DeviceEngine(N, L, max_neighbors=96) as e:
        e.set_activity(0.1, 0.1)
        e.upload(x=o.x, y=o.y, R=o.R, phi=o.phi, cosp=o.cosp, sinp=o.sinp)
        assert np.array_equal(e.pair_set(), o.pair_set())
        t = e.tuning()
        assert t["tile_max"] <= t["tile_cap"]
    o.close()


@pytest.mark.parametrize("N,rho", [(65536, 1.0), (1048576, 0.9)])
def test_full_size_parity_and_properties(N, rho):
    """BASELINE.json configs[1] and [2] sizes: direct comparison with the oracle (pair set bit-exact,
    single step 1e-12) and size-independent properties of a free run."""
    l_s, l_n = 0.05, 0.5
    o, rng = relaxed_oracle(N, rho, seed=N % 1000, l_s=l_s, l_n=l_n, presteps=12)
    sc = o.scalars(); L = sc["L"]
    o.assign(); o.build()
    with device_from_state(o.state(), seed=77) as e:
        assert np.array_equal(e.pair_set(), o.pair_set())                   # gate 1 at full size
        nz = rng.uniform(-PI, PI, N)
        o.step(nz); e.step_injected(nz)
        assert_state_close(e.download(), o, L, TOL, "full-size step")       # gate 2 at full size
    # free run with on-device Philox noise and rebuilds; twice, for determinism
    runs = []
    for rep in range(2):
        with device_from_state(o.state(), seed=77) as e:
            e.step(150)
            runs.append((e.download(), e.counters(), e.get_com()))
    (d, cnt, c), (d2, cnt2, _) = runs
    for f in d:
        assert np.array_equal(d[f], d2[f]), "run-to-run difference in " + f
    assert cnt2["resetCounter"] == cnt["resetCounter"]
    assert np.all(d["x"] >= -L / 2) and np.all(d["x"] < L / 2) and np.all(d["y"] >= -L / 2) and np.all(d["y"] < L / 2)
    # x - x_real is a whole number of box lengths (PBC is a single wrap per step, Cell.h:168-175)
    kx = (d["x_real"] - d["x"]) / L
    assert np.max(np.abs(kx - np.rint(kx))) < 1e-9
    # COM held by the engine == mean of x_real (calculate_COM, jamming.cpp:761-774)
    assert abs(c["COM"][0] - d["x_real"].mean()) <= 1e-11 * L and abs(c["COM"][1] - d["y_real"].mean()) <= 1e-11 * L
    # Newton's third law: sum_i (R_i v_i - CFself R_i n_i) = sum of pair forces = 0
    fx = np.sum(d["R"] * d["vx"] - l_s * d["R"] * d["cosp"]); fy = np.sum(d["R"] * d["vy"] - l_s * d["R"] * d["sinp"])
    scale = np.sum(np.abs(d["R"] * d["vx"])) + 1.0
    assert abs(fx) <= 1e-11 * scale and abs(fy) <= 1e-11 * scale
    assert np.allclose(d["cosp"] ** 2 + d["sinp"] ** 2, 1.0, atol=1e-14)
    assert np.allclose(np.cos(d["phi"]), d["cosp"], atol=1e-14)
    assert cnt["rebuilds"] >= 1 and cnt["overflow"] == 0 and cnt["step"] == 150
    o.close()


@pytest.mark.parametrize("lanes", [1, 4])
@pytest.mark.parametrize("l_s", [0.05, 0.5])
def test_skin_aware_sweep_length_is_bit_identical_to_full_sweep(l_s, lanes):
    """A step sweeps only the list classes that can have come within rn given the skin-test value
    (apj_device.cuh APJ_CLASSES); the omitted entries all fail d2 < rn2, so trajectories must be
    BIT-identical to sweeping the full lists -- over many rebuilds, with the same resetCounter."""
    N, rho, seed, steps = 4096, 0.9, 7, 600
    o, _ = relaxed_oracle(N, rho, seed=11, l_s=l_s, l_n=0.5)
    s = o.state()
    o.close()
    out = []
    for on in (1, 0):
        with device_from_state(s, seed=seed, lanes_per_particle=lanes) as e:
            e.set_sweep_truncation(on)
            e.step(steps)
            # x_old reset WITHOUT a list build (start() :203): the shortening stays in force through it, the bound now
            # being skinBase (pair displacement between the build and the reset) + the skin-test value since the reset
            e.step(7); e.mark_origin()
            mid = e.sweep_stats()
            e.step(60)
            st = e.sweep_stats()
            assert not on or mid["active"]        # (how many classes the next steps skip depends on how old the lists were at the reset)
            e.step(steps - 67)
            st = e.sweep_stats()
            out.append((e.download(), e.counters(), st))
    steps *= 2
    (a, ca, sa), (b, cb, sb) = out
    assert ca["step"] == cb["step"] == steps
    assert ca["resetCounter"] == cb["resetCounter"] and ca["resetCounter"] >= 2
    for f in ("x", "y", "x_real", "y_real", "cosp", "sinp", "vx", "vy", "phi", "x_old", "y_old"):
        assert np.array_equal(a[f], b[f]), f
    assert sa["active"] and not sb["active"]           # the shortening was in force in the first run
    assert sb["retried"] == 0 and sa["retried"] <= steps // 10


def test_split_tail_equals_fused_tail():
    """Large systems end a step with apj_reduce_commit_kernel instead of the step kernel's last block
    (APJ_FLAG_SPLIT_TAIL / _FUSED_TAIL force either form). Only the rounding of COM may differ (another
    fixed summation order), which enters nothing but the skin test: positions are bit-identical, rebuilds
    fall on the same steps."""
    N, rho, seed, steps = 6000, 0.9, 5, 400
    o, _ = relaxed_oracle(N, rho, seed=21, l_s=0.3, l_n=0.5)
    s = o.state()
    L = s["L"]
    o.close()
    out = []
    for flags in (2, 4, 2 | 1, 2 | 8 | 16):     # 16 | 8: pipelined persistent step kernel on 3 blocks, ~8 tiles per block
        with device_from_state(s, seed=seed, lanes_per_particle=1, flags=flags) as e:
            e.step(1); e.step(steps - 1)
            out.append((e.download(), e.counters(), e.get_com(0)))
    (a, ca, ma), (b, cb, mb), (c, cc, mc), (d, cd, md) = out
    assert cd["step"] == steps and cd["resetCounter"] == ca["resetCounter"]
    for f in ("x", "y", "x_real", "y_real", "cosp", "sinp", "vx", "vy", "phi", "x_old", "y_old"):
        assert np.array_equal(a[f], d[f]), "persistent grid of 3 blocks: " + f
    assert np.max(np.abs(np.asarray(ma["COM"]) - np.asarray(md["COM"]))) <= 1e-12 * L   # one partial per block: another fixed order
    assert ca["step"] == cb["step"] == cc["step"] == steps
    assert ca["resetCounter"] == cb["resetCounter"] == cc["resetCounter"] and ca["resetCounter"] >= 2
    assert ca["launches"] > cb["launches"]                      # one more kernel per step
    for f in ("x", "y", "x_real", "y_real", "cosp", "sinp", "vx", "vy", "phi", "x_old", "y_old"):
        assert np.array_equal(a[f], b[f]) and np.array_equal(a[f], c[f]), f
    assert np.max(np.abs(np.asarray(ma["COM"]) - np.asarray(mb["COM"]))) <= 1e-12 * L
    assert np.array_equal(np.asarray(ma["COM"]), np.asarray(mc["COM"]))   # graph / direct launches: same bits


@pytest.mark.parametrize("tb", [256, 128])
def test_ring_step_kernel_is_bit_identical_to_the_classic_one(monkeypatch, tb):
    """APJ_STEP_RING=1 (experimental): one block per SM, consumer groups around a ring of tile buffers
    refilled by the last warp to finish a sweep, one partial per warp folded in the classic order. Same bits as the
    classic split-tail path -- positions, COM, rebuild steps -- over ~8 tiles per block with refills and rebuilds."""
    from active_particle_jamming_b200 import DeviceEngine
    N, rho, seed = 300000, 0.9, 11
    L = DeviceEngine.lattice_box_length(N, rho, seed)[0]
    out = []
    for ring in (0, 1):
        monkeypatch.setenv("APJ_STEP_RING", str(ring))
        monkeypatch.setenv("APJ_TB", str(tb))
        with DeviceEngine(N, L, seed=3, lanes_per_particle=1, flags=2) as e:    # 2: split tail
            e.init_lattice(seed)
            e.set_activity(0.0, 0.5); e.step(120)                                # passive relaxation off the lattice
            e.set_activity(0.1, 0.3); e.step(1)
            tun = e.tuning()
            e.step(170)
            tun["smem_bytes"] = min(tun["smem_bytes"], e.tuning()["smem_bytes"])  # (a denser tile would send the handle back to the classic kernel)
            out.append((e.download(), e.counters(), e.get_com(0), tun))
    (a, ca, ma, ta), (b, cb, mb, tb_) = out
    assert ta["tb"] == tb_["tb"] == tb
    assert tb_["smem_bytes"] > 200000 > ta["smem_bytes"] > 0                     # the ring kernel was in use: 5 / 8 tile buffers fill the SM
    assert ca["step"] == cb["step"] == 291 and ca["resetCounter"] == cb["resetCounter"] and ca["resetCounter"] >= 2
    for f in a:
        assert np.array_equal(a[f], b[f]), f
    for k in ("COM", "COM0", "COM_old"):
        assert np.array_equal(np.asarray(ma[k]), np.asarray(mb[k])), k


def test_checkpoint_restart_continues_the_run(tmp_path):
    """apj_save_checkpoint / apj_load_checkpoint (an addition; the reference cannot resume): the restored
    handle holds the stored bits, continues the Philox stream at the stored step, takes its first step within
    1e-12 of the uninterrupted run and rebuilds on the same steps."""
    from active_particle_jamming_b200 import DeviceEngine
    from active_particle_jamming_b200.device import ApjError
    N, rho = 4096, 0.9
    o, _ = relaxed_oracle(N, rho, seed=31, l_s=0.3, l_n=0.5)
    s = o.state()
    L = s["L"]
    o.close()
    path = str(tmp_path / "run.apjckpt")
    with device_from_state(s, seed=9) as a:
        a.step(150)
        a.save_checkpoint(path)
        ca, da, ma = a.counters(), a.download(), a.get_com(0)
        a.step(1)
        d1 = a.download()
        a.step(200)
        c2 = a.counters()
    with DeviceEngine(N, L, seed=4242) as b:                   # another key: the checkpoint's must win
        b.load_checkpoint(path)
        cb, db, mb = b.counters(), b.download(), b.get_com(0)
        assert cb["step"] == ca["step"] == 150 and cb["resetCounter"] == ca["resetCounter"]
        for f in da:
            if f != "box":                                     # box ids are re-derived from the positions at load
                assert np.array_equal(da[f], db[f]), f
        for k in ("COM", "COM0", "COM_old"):
            assert np.array_equal(np.asarray(ma[k]), np.asarray(mb[k])), k
        b.step(1)
        e1 = b.download()
        assert np.max(wrapped_abs_diff(e1["x"], d1["x"], L)) <= TOL * L and np.max(wrapped_abs_diff(e1["y"], d1["y"], L)) <= TOL * L
        assert rel_err(e1["cosp"], d1["cosp"]) <= TOL and rel_err(e1["sinp"], d1["sinp"]) <= TOL      # same noise stream
        b.step(200)
        c3 = b.counters()
        assert c3["step"] == c2["step"] == 351 and abs(c3["resetCounter"] - c2["resetCounter"]) <= 1
    with DeviceEngine(N // 2, L, seed=1) as c:
        with pytest.raises(ApjError):
            c.load_checkpoint(path)                            # another shape
        with pytest.raises(ApjError):
            c.load_checkpoint(str(tmp_path / "missing"))
