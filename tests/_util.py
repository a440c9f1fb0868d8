"""Shared helpers of the test-suite: golden loading and state hand-over between the three
implementations (reference build / oracle restatement / CUDA path through the C ABI)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCAL = ["L", "Lover2", "lp", "b", "nbox", "COMx", "COMy", "COM0x", "COM0y", "COMoldx", "COMoldy", "resetCounter", "t",
        "CFself", "CTnoise"]
STATE = ["x", "y", "xr", "yr", "x0", "y0", "xo", "yo", "R", "phi", "cosp", "sinp", "vx", "vy", "xnew", "ynew"]
GOLDEN_CASES = ["n256_dense_fast", "n1024_base", "n100_tiny"]
# device field name <- oracle/reference field name
DEV2ORC = dict(x="x", y="y", x_real="xr", y_real="yr", x0="x0", y0="y0", x_old="xo", y_old="yo", R="R", phi="phi",
               cosp="cosp", sinp="sinp", vx="vx", vy="vy")


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    g["N"], g["rho"], g["l_s"], g["l_n"] = int(g["meta"][0]), float(g["meta"][1]), float(g["meta"][2]), float(g["meta"][3])
    g["K"], g["steps"] = int(g["meta"][5]), int(g["meta"][6])
    return g


def golden_state(g, prefix):
    s = {k: g[prefix + k] for k in STATE if prefix + k in g}
    s["box"] = g[prefix + "box"]
    s.update(dict(zip(SCAL, g[prefix + "scal"])))
    return s


def oracle_from_state(s, dens):
    """OracleSim holding exactly the state dict `s` (as returned by RefEngine.get_state /
    OracleSim.state / golden_state)."""
    from oracle.pyoracle import OracleSim
    o = OracleSim(len(s["x"]), s["L"], dens)
    for k in STATE:
        if k in s:
            getattr(o, k)[:] = s[k]
    o.Rinv[:] = 1.0 / o.R
    o.box[:] = s["box"]
    o.set_params(s["CFself"], s["CTnoise"])
    o.topology()
    o.set_com([s["COMx"], s["COMy"]], [s["COM0x"], s["COM0y"]], [s["COMoldx"], s["COMoldy"]])
    o.set_reset_counter(int(s["resetCounter"]))
    return o


def device_from_state(s, seed=12345, **kw):
    """DeviceEngine (C ABI) holding the state dict `s`; lists are built by the upload."""
    from active_particle_jamming_b200 import DeviceEngine
    e = DeviceEngine(len(s["x"]), s["L"], seed=seed, **kw)
    e.set_activity(s["CFself"], s["CTnoise"])
    e.upload(box=s["box"], **{d: s[o] for d, o in DEV2ORC.items()})
    e.set_com(0, com=[s["COMx"], s["COMy"]], com0=[s["COM0x"], s["COM0y"]], com_old=[s["COMoldx"], s["COMoldy"]])
    e.set_reset_counter(int(s["resetCounter"]))
    return e


def half_pairs(off, idx):
    from oracle.pyoracle import pairs_canonical
    i = np.repeat(np.arange(len(off) - 1, dtype=np.int64), np.diff(off))
    return pairs_canonical(i, np.asarray(idx, dtype=np.int64))


def wrapped_abs_diff(a, b, L):
    d = np.abs(np.asarray(a) - np.asarray(b))
    return np.minimum(d, np.abs(L - d))


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor): relative where values are O(1) or larger, absolute below."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def random_system(N, rho, seed):
    """Synthetic initial condition of SURVEY §8(d): radii 1 + N(0,1)/10 (jamming.cpp:296), L from the
    reference formula (:305), uniform random positions and angles."""
    from oracle.pyoracle import OracleSim, PI
    rng = np.random.default_rng(seed)
    R = 1 + 0.1 * rng.standard_normal(N)
    L = OracleSim.box_length(R, rho)
    x = rng.uniform(-L / 2, L / 2, N)
    y = rng.uniform(-L / 2, L / 2, N)
    phi = rng.uniform(-PI, PI, N)
    return R, L, x, y, phi
