"""Scratch probe run on the GPU box: parity vs oracle + first timings."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from active_particle_jamming_b200 import DeviceEngine, PI
from oracle.pyoracle import OracleSim

def make(N, rho, seed):
    rng = np.random.default_rng(seed)
    R = 1 + 0.1 * rng.standard_normal(N)
    L = OracleSim.box_length(R, rho)
    x = rng.uniform(-L / 2, L / 2, N); y = rng.uniform(-L / 2, L / 2, N); phi = rng.uniform(-PI, PI, N)
    return R, L, x, y, phi

N = 4096; rho = 0.9
R, L, x, y, phi = make(N, rho, 1)
o = OracleSim.from_arrays(R, x, y, phi, rho); o.set_params(0.05, 0.5); o.topology(); o.assign(); o.build()
# relax a bit on the oracle so overlaps are physical
rng = np.random.default_rng(2)
o.xr[:] = o.x; o.yr[:] = o.y
o.mark_origin()
for t in range(200):
    o.step(rng.uniform(-PI, PI, N))
st0 = o.state()
e = DeviceEngine(N, L, seed=7)
e.set_activity(0.05, 0.5)
e.upload(x=st0["x"], y=st0["y"], x_real=st0["xr"], y_real=st0["yr"], x0=st0["x0"], y0=st0["y0"], x_old=st0["xo"], y_old=st0["yo"],
         R=st0["R"], phi=st0["phi"], cosp=st0["cosp"], sinp=st0["sinp"], vx=st0["vx"], vy=st0["vy"], box=st0["box"])
e.set_com(0, com=[st0["COMx"], st0["COMy"]], com0=[st0["COM0x"], st0["COM0y"]], com_old=[st0["COMoldx"], st0["COMoldy"]])
print("geometry", e.geometry(), o.scalars()["lp"], o.scalars()["b"])
o.assign(); o.build()
ps_o = o.pair_set(); ps_g = e.pair_set()
print("pair sets equal:", np.array_equal(ps_o, ps_g), len(ps_o), len(ps_g))
d = e.download()
print("box equal:", np.array_equal(d["box"], o.box))
maxerr = {}
for t in range(50):
    nz = rng.uniform(-PI, PI, N)
    rb = o.step(nz); e.step_injected(nz)
    d = e.download()
    for k, ko in [("x", "x"), ("y", "y"), ("x_real", "xr"), ("y_real", "yr"), ("phi", "phi"), ("cosp", "cosp"), ("sinp", "sinp"), ("vx", "vx"), ("vy", "vy"), ("x_old", "xo")]:
        err = np.max(np.abs(d[k] - getattr(o, ko)))
        maxerr[k] = max(maxerr.get(k, 0), err)
print("50 injected steps max abs err:", {k: float("%.3g" % v) for k, v in maxerr.items()})
print("counters", e.counters(), "oracle resets", o.scalars()["resetCounter"])
print("COM", e.get_com()["COM"], o.scalars()["COMx"], o.scalars()["COMy"])
# philox multi-step
o2 = OracleSim.from_arrays(R, x, y, phi, rho); o2.set_params(0.05, 0.5); o2.topology(); o2.assign(); o2.build(); o2.xr[:] = o2.x; o2.yr[:] = o2.y; o2.mark_origin()
e2 = DeviceEngine(N, L, seed=99); e2.set_activity(0.05, 0.5)
e2.upload(x=o2.x, y=o2.y, R=o2.R, phi=o2.phi, cosp=o2.cosp, sinp=o2.sinp); e2.mark_origin(); e2.skip_self_term_once()
nreb = o2.run_philox(99, 0, 100)
e2.step(100)
d2 = e2.download()
dxw = np.abs(d2["x"] - o2.x); dxw = np.minimum(dxw, L - dxw)
print("100 philox steps: max|dx|", dxw.max(), "oracle rebuilds", nreb, e2.counters())
print("order", e2.order_orientation(), o2.order(), o2.orientation(), "msd", e2.msd(), o2.msd())
print("fluct", e2.fluct_area(5.0), o2.fluct_area(5.0))

# other lane counts: pair set + 20 injected steps
for G in (1, 2, 8):
    eg = DeviceEngine(N, L, seed=7, lanes_per_particle=G); eg.set_activity(0.05, 0.5)
    og = OracleSim.from_arrays(R, x, y, phi, rho); og.set_params(0.05, 0.5); og.topology(); og.assign(); og.build(); og.xr[:] = og.x; og.yr[:] = og.y; og.mark_origin()
    eg.upload(x=og.x, y=og.y, R=og.R, phi=og.phi, cosp=og.cosp, sinp=og.sinp); eg.mark_origin(); eg.skip_self_term_once()
    ok = np.array_equal(og.pair_set(), eg.pair_set())
    rg = np.random.default_rng(5)
    for t in range(20):
        nz = rg.uniform(-PI, PI, N); og.step(nz); eg.step_injected(nz)
    dg = eg.download()
    dxw = np.abs(dg["x"] - og.x); dxw = np.minimum(dxw, L - dxw)
    print("G", G, "pairs", ok, "max|dx|", dxw.max(), "max|dphi|", np.max(np.abs(dg["phi"] - og.phi)), eg.counters()["resetCounter"], og.scalars()["resetCounter"])
    eg.close()

# timing
for Nb, rho_b, Gs in [(65536, 1.0, (1, 2, 4)), (262144, 0.9, (1, 2)), (1048576, 0.9, (1,)), (4194304, 0.9, (1,))]:
  R, L, x, y, phi = make(Nb, rho_b, 3)
  for G in Gs:
    eb = DeviceEngine(Nb, L, seed=1, lanes_per_particle=G); eb.set_activity(0.05, 0.5)
    t0 = time.time()
    eb.upload(x=x, y=y, R=R, phi=phi); eb.mark_origin()
    tb = time.time() - t0
    eb.set_activity(0.0, 0.5); eb.step(300)     # crude relax
    eb.set_activity(0.05, 0.5); eb.step(100); eb.step(1)
    c0 = eb.counters()
    nsteps = 2000 if Nb <= 65536 else 300
    eb.timer_begin(); eb.step(nsteps); ms = eb.timer_end()
    c1 = eb.counters()
    kms, kc = eb.time_step_kernel(200)
    print(eb.tuning())
    print("N %d G %d: upload+build %.3fs | %d steps %.4f ms/step %.3e p-steps/s rebuilds %d list_max %d | step kernel %.4f ms -> %.0f GB/s at 197 B" % (
        Nb, G, tb, nsteps, ms / nsteps, Nb * nsteps / (ms * 1e-3), c1["rebuilds"] - c0["rebuilds"], c1["list_max"], kms, Nb * 197 / (kms * 1e-3) / 1e9))
    eb.close()
