"""Live cross-check of the oracle restatement against the compiled reference (oracle/_ref), beyond
the committed golden vectors: other sizes, densities and activities, longer runs, the reference's
own initCells + relax. Skipped where oracle/_ref is absent (it cannot be built without
/root/reference; the prebuilt .so travels to the GPU box with the snapshot)."""
import numpy as np
import pytest

from _util import STATE, oracle_from_state
from oracle.pyoracle import PI, OracleSim, RefEngine, have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libapj_ref.so not built (no /root/reference here)")


def test_constants_are_the_truncated_literals():
    L = RefEngine.lib()
    assert L.apjref_PI() == 3.14159265 and L.apjref_PI2() == 6.28318531 and L.apjref_ndim() == 2   # jamming.cpp:2-4


@pytest.mark.parametrize("N,rho,l_s,l_n,seed,steps", [(512, 0.84, 0.3, 0.1, 7, 300), (2048, 1.0, 0.05, 0.5, 8, 120),
                                                      (333, 0.9, 2.0, 1.0, 9, 200)])
def test_long_injected_run_bit_identical(N, rho, l_s, l_n, seed, steps):
    RefEngine.seed(seed)
    r = RefEngine(N, 1000, l_s, l_n, rho)
    r.init_cells()
    r.topology()
    r.assign()
    r.build()
    r.mark_origin()
    s0 = r.get_state()
    # first step of a fresh Engine: Cell::x_new is still 0 (Cell.h:75) -> state carries xnew=ynew=0
    o = oracle_from_state(s0, rho)
    o.assign(bruteforce=True)
    o.build()
    rng = np.random.default_rng(seed)
    rebuilds = 0
    for k in range(steps):
        nz = rng.uniform(-PI, PI, N)
        r.step(nz)
        rebuilds += o.step(nz)
    s1 = r.get_state()
    for f in STATE:
        assert np.array_equal(getattr(o, f), s1[f]), f
    assert np.array_equal(o.box, s1["box"])
    assert o.scalars()["resetCounter"] == s1["resetCounter"] and rebuilds == s1["resetCounter"]
    ro, ri = r.verlet()
    oo, oi = o.verlet()
    assert np.array_equal(ro, oo) and np.array_equal(ri, oi)
    assert o.order() == r.order() and o.msd() == r.msd()
    r.close()
    o.close()


def test_delta_norm_matches_reference():
    RefEngine.seed(1)
    r = RefEngine(128, 10, 0.1, 0.1, 0.9)
    r.init_cells()
    r.topology()
    L = r.scalars()["L"]
    o = OracleSim(128, L, 0.9)
    for d in np.concatenate([np.linspace(-2.6 * L, 2.6 * L, 301), [L / 2, -L / 2, np.nextafter(L / 2, 0), 0.0]]):
        assert o.l.orc_delta_norm(o.h, d) == r.l.apjref_delta_norm(r.h, d)     # jamming.cpp:872-880
    r.close()
    o.close()


def test_uniform_mapping_matches_shimmed_randuni():
    """oracle's u32 -> U[-PI,PI) map equals the (shimmed) boost::uniform_real path the reference uses."""
    import random
    RefEngine.seed(4242)
    mt = np.random.MT19937()
    # std::mt19937 seeded with 4242 == numpy's legacy seeding of MT19937
    mt._legacy_seeding(4242)
    raw = mt.random_raw(64)
    got = [RefEngine.randuni() for _ in range(64)]
    want = [OracleSim.lib().orc_u32_to_randuni(int(u)) for u in raw]
    assert got == want
    assert all(-PI <= g < PI for g in got)


@pytest.mark.parametrize("N,rho,seed", [(400, 1.0, 3), (1500, 0.9, 4)])
def test_overlap_hue_matches_reference_neighbor_interactions(N, rho, seed):
    """Cell::over (the Ovito hue; jamming.cpp:653-656): the oracle's restatement equals what the reference's own
    neighborInteractions leaves behind on a filmed step -- an int accumulated with truncation at every update,
    so the ORDER of the pair visits matters."""
    RefEngine.seed(seed)
    r = RefEngine(N, 1000, 0.3, 0.3, rho)
    r.init_cells(); r.topology(); r.assign(); r.build(); r.mark_origin()
    rng = np.random.default_rng(seed)
    for _ in range(15):
        r.step(rng.uniform(-PI, PI, N))
    r.assign(); r.build()
    s = r.get_state()
    o = oracle_from_state(s, rho)
    o.assign(bruteforce=True); o.build()
    want = r.overlap_hue()
    got = o.overlap_hue()
    assert np.array_equal(got, want)
    assert want.min() < 236 and want.max() == 240          # some particles overlap visibly, some not at all
    r.close()
    o.close()
