"""Generates tests/golden/*.npz by running the REFERENCE ITSELF (oracle/_ref/libapj_ref.so, built
by oracle/Makefile from /root/reference/code/jam/jamming.cpp + code/classes/*.h with NDIM=2 and the
Boost.Random shim). Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The reference owns no tests or golden vectors, and /root/reference does not exist on
the GPU box, so these files are what pins the oracle (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_golden.py) to outputs of the reference's own code. Every array is produced by
reference functions only: Engine::initCells / topology / assignCellsToGrid / buildVerletLists /
relax / calculate_next_positions (noise injected in place of randuni() at jamming.cpp:667) and the
observables of jamming.cpp:776-823, classes/Fluctuations.h, classes/Correlations.h.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.pyoracle import RefEngine, PI, build  # noqa: E402

STATE = ["x", "y", "xr", "yr", "x0", "y0", "xo", "yo", "R", "phi", "cosp", "sinp", "vx", "vy", "xnew", "ynew"]
SCAL = ["L", "Lover2", "lp", "b", "nbox", "COMx", "COMy", "COM0x", "COM0y", "COMoldx", "COMoldy", "resetCounter", "t",
        "CFself", "CTnoise"]

CASES = [
    # name, N, rho, l_s, l_n, seed, K injected steps, checkpoints
    ("n256_dense_fast", 256, 1.0, 1.0, 0.3, 11, 100, (1, 2, 50, 100)),
    ("n1024_base", 1024, 0.9, 0.05, 0.5, 1234, 24, (1, 24)),
    ("n100_tiny", 100, 0.84, 0.2, 1.0, 5, 40, (1, 40)),
]


def snapshot(r, prefix, out):
    s = r.get_state()
    for k in STATE:
        out[prefix + k] = s[k]
    out[prefix + "box"] = s["box"]
    out[prefix + "scal"] = np.array([s[k] for k in SCAL])


def make(name, N, rho, l_s, l_n, seed, K, cks):
    out = {}
    RefEngine.seed(seed)
    steps = 1000  # Engine::totalSteps; sets Fluctuations::time_interval = steps/(10*10)
    r = RefEngine(N, steps, l_s, l_n, rho)
    r.init_cells()
    r.topology()
    r.assign()
    r.build()
    r.relax()          # jamming.cpp:482-525 (local: 2000 + 2000 steps)
    r.mark_origin()    # :191-203
    # a fresh assign+build from the start state: the pair-set parity object
    r.assign()
    r.build()
    snapshot(r, "s0_", out)
    off, idx = r.verlet()
    out["s0_vl_off"], out["s0_vl_idx"] = off, idx
    coff, cidx = r.cell_lists()
    out["s0_cl_off"], out["s0_cl_idx"] = coff, cidx
    out["box_neighbors"] = r.box_neighbors()

    rng = np.random.default_rng(seed + 1000)
    # U[-PI,PI) on the 2^-32 lattice of boost::uniform_real over mt19937
    u = rng.integers(0, 2 ** 32, size=(K, N), dtype=np.uint64).astype(np.float64)
    noise = u / 4294967296.0 * (PI - (-PI)) + (-PI)
    out["noise"] = noise
    resets = []
    for k in range(K):
        r.step(noise[k])
        resets.append(r.scalars()["resetCounter"])
        if (k + 1) in cks:
            snapshot(r, "s%d_" % (k + 1), out)
    out["resets"] = np.array(resets)
    out["checkpoints"] = np.array(cks)
    # lists the reference is holding after the last step (built at its last rebuild)
    off, idx = r.verlet()
    out["end_vl_off"], out["end_vl_idx"] = off, idx

    # observables on the end state -- reference code paths
    out["order"] = np.array(r.order())
    out["orientation"] = r.orientation()
    out["msd"] = np.array(r.msd())
    tmp = tempfile.mkdtemp()
    r.attach_observers(tmp)
    out["fluct_overlap_args"] = np.array([[1.0, 3.0, 2.5], [0.9, 5.0, 5.2], [1.1, 3.0, 1.0], [1.0, 10.0, 10.9], [1.2, 3.0, 3.0]])
    out["fluct_overlap"] = np.array([r.fluct_overlap(*a) for a in out["fluct_overlap_args"]])
    # Fluctuations state machine, repeated calls on the same end state (Q13): record every call
    seq = []
    for _ in range(25):
        f = r.fluct_measure()
        seq.append([f["current_radius"], f["current_value"], f["counter"]])
    out["fluct_seq"] = np.array(seq)
    out["fluct_time_interval"] = np.array(f["time_interval"])
    out["fluct_rad_interval"] = np.array(f["rad_interval"])
    # spatialCorrelations needs fresh CellLists (start() :245-246)
    r.assign()
    r.build()
    coff, cidx = r.cell_lists()
    out["end_cl_off"], out["end_cl_idx"] = coff, cidx
    d = r.corr_dims()
    out["corr_dims"] = np.array([d["nc"], d["np"], d["noBins"], d["correlation_time"]])
    v, o, p = r.spatial_correlations()
    out["corr_vel"], out["corr_ori"], out["corr_pair"] = v, o, p
    out["vel_dist"] = r.vel_dist()
    out["dens_dist"] = r.density_distribution()
    out["meta"] = np.array([N, rho, l_s, l_n, seed, K, steps], dtype=np.float64)
    r.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written:", sum(v.nbytes for v in out.values()) // 1024, "KiB raw;", int(out["resets"][-1]), "skin rebuilds")


REF = "/root/reference"


def read_tree(root):
    out = {}
    for dp, _, files in os.walk(root):
        for f in sorted(files):
            out[os.path.relpath(os.path.join(dp, f), root)] = open(os.path.join(dp, f), "rb").read().decode("latin-1")
    return out


def make_print_bytes():
    """Every Print::print_* of the REFERENCE's Print.h driven by tests/print_probe.cpp: exact bytes."""
    import json
    import subprocess
    tmp = tempfile.mkdtemp()
    exe = os.path.join(tmp, "probe")
    subprocess.run(["g++", "-O1", "-std=c++11", "-w", "-I", os.path.join(REF, "code", "classes"),
                    os.path.join(HERE, "..", "print_probe.cpp"), "-o", exe], check=True)
    root = os.path.join(tmp, "out") + "/"
    os.makedirs(root + "local_output")
    subprocess.run([exe, root], check=True, stderr=subprocess.DEVNULL)
    tree = read_tree(os.path.join(root, "local_output", "probe"))
    json.dump(tree, open(os.path.join(HERE, "print_bytes.json"), "w"), indent=1, sort_keys=True)
    print("print_bytes.json:", len(tree), "files")


def make_run_shape(N=1024, steps=1000, l_s=0.05, l_n=0.5, rho=0.9, seed=77):
    """One full run of the reference driver (Engine::start, jamming.cpp:173-283): the shape of every
    output file (lines, columns, first column) + the physics scalars of its summary."""
    import json
    tmp = tempfile.mkdtemp()
    RefEngine.seed(seed)
    r = RefEngine(N, steps, l_s, l_n, rho)
    secs = r.run_start(tmp)
    tree = read_tree(os.path.join(tmp, "local_output", "apjref", "run0"))
    shape = {}
    for name, text in tree.items():
        lines = text.split("\n")
        assert lines[-1] == ""
        lines = lines[:-1]
        shape[name] = {"lines": len(lines), "columns": sorted(set(len(l.split("\t")) for l in lines)),
                       "first_column": [l.split("\t")[0] for l in lines][:400]}
    def col(name, k):
        return [float(l.split("\t")[k]) for l in tree[name].split("\n")[:-1]]
    meta = {"N": N, "steps": steps, "l_s": l_s, "l_n": l_n, "rho": rho, "seconds": secs,
            "order_mean": float(np.mean(col("dat/order.dat", 1))), "order_std": float(np.std(col("dat/order.dat", 1))),
            "msd_last": col("dat/MSD.dat", 1)[-1], "fluct": [col("dat/fluct.dat", 0), col("dat/fluct.dat", 1)],
            "densDist": col("dat/densDist.dat", 1), "corr": col("dat/corr.dat", 1), "orientationCorr": col("dat/orientationCorr.dat", 1),
            "pairCorr_sum": float(np.nansum(col("dat/pairCorr.dat", 1))), "velDist": col("dat/velDist.dat", 1),
            "summary": tree["dat/summary.dat"]}
    json.dump({"shape": shape, "meta": meta}, open(os.path.join(HERE, "run_shape.json"), "w"), indent=1, sort_keys=True)
    r.close()
    print("run_shape.json: reference run took %.2f s" % secs)


if __name__ == "__main__":
    build()
    for c in CASES:
        make(*c)
    make_print_bytes()
    make_run_shape()
