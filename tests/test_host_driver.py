"""The C++ host mirror of the reference's Engine / code/classes interface
(active_particle_jamming_b200/host): same CLI, same output tree and line formats, hot path through
the C ABI. CPU part: Print is byte-identical to the reference's Print (golden bytes written by
tests/golden/make_golden.py from /root/reference/code/classes/Print.h), the driver keeps the
reference's usage/exit-code behaviour and refuses to run without a GPU. GPU part: a full run
produces the file tree of a reference run, and its physics agrees statistically."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "active_particle_jamming_b200", "host")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def jam():
    from active_particle_jamming_b200 import _build
    return _build.build_host()


def read_tree(root):
    out = {}
    for dp, _, files in os.walk(root):
        for f in sorted(files):
            out[os.path.relpath(os.path.join(dp, f), root)] = open(os.path.join(dp, f), "rb").read().decode("latin-1")
    return out


def test_print_is_byte_identical_to_reference(tmp_path):
    exe = str(tmp_path / "probe")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(HOST, "classes"), os.path.join(ROOT, "tests", "print_probe.cpp"),
                    "-o", exe], check=True)
    root = str(tmp_path / "out") + "/"
    os.makedirs(root + "local_output")
    subprocess.run([exe, root], check=True)
    got = read_tree(os.path.join(root, "local_output", "probe"))
    want = json.load(open(os.path.join(GOLD, "print_bytes.json")))
    assert sorted(got) == sorted(want)                       # the 14 files of Print.h:79-92
    for name in want:
        assert got[name] == want[name], name


def test_usage_and_exit_code(jam):
    r = subprocess.run([jam, "only", "three", "args"], capture_output=True, text=True)
    assert r.returncode == 1                                 # reference jamming.cpp:901-912
    assert r.stdout.startswith("Incorrect number of arguments. Need: \n- full run ID\n- single run ID\n- number of cells")
    assert r.stdout.rstrip().endswith("Program exit status (1)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_cuda_available(), reason="only meaningful without a GPU")
def test_driver_has_no_cpu_path(jam, tmp_path):
    env = dict(os.environ, APJ_OUTPUT_ROOT=str(tmp_path), APJ_SEED="1")
    r = subprocess.run([jam, "full", "run0", "256", "100", "0.1", "0.5", "0.9"], capture_output=True, text=True, env=env)
    assert r.returncode != 0 and "no CPU fallback" in r.stdout


def test_host_classes_never_compute_the_hot_path():
    src = open(os.path.join(HOST, "classes", "Cell.h")).read()
    assert "exit(118)" in src                                # Cell::update aborts: the update is the kernel's epilogue
    drv = open(os.path.join(HOST, "jam", "jamming.cpp")).read() + open(os.path.join(HOST, "classes", "Batch.h")).read()
    for call in ("apj_step", "apj_force_rebuild", "apj_order_orientation", "apj_msd", "apj_mark_origin", "apj_fluct_area",
                 "apj_spatial_correlations", "apj_vel_hist", "apj_occupancy_hist"):
        assert call in drv


@pytest.mark.gpu
def test_full_run_matches_reference_output_tree(jam, tmp_path):
    g = json.load(open(os.path.join(GOLD, "run_shape.json")))
    m = g["meta"]
    env = dict(os.environ, APJ_OUTPUT_ROOT=str(tmp_path), APJ_SEED="20240607")
    r = subprocess.run([jam, "apjref", "run0", str(m["N"]), str(m["steps"]), str(m["l_s"]), str(m["l_n"]), str(m["rho"])],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    tree = read_tree(os.path.join(str(tmp_path), "local_output", "apjref", "run0"))
    assert sorted(tree) == sorted(g["shape"])
    for name, sh in g["shape"].items():
        lines = tree[name].split("\n")[:-1]
        assert len(lines) == sh["lines"], name
        assert sorted(set(len(l.split("\t")) for l in lines)) == sh["columns"], name
        if name not in ("dat/summary2.dat", "dat/fluct.dat"):     # first column = time / abscissa / label: deterministic
            assert [l.split("\t")[0] for l in lines][:400] == sh["first_column"], name

    def col(name, k):
        return np.array([float(l.split("\t")[k]) for l in tree[name].split("\n")[:-1]])
    # fluct.dat: the expected areas are a deterministic function of the radius schedule (up to L, which
    # depends on the radii drawn): the first radius is exactly 3
    assert abs(col("dat/fluct.dat", 0)[0] - m["fluct"][0][0]) < 1e-3
    # statistical agreement with the reference run (different random initial condition and noise stream)
    assert abs(col("dat/order.dat", 1).mean() - m["order_mean"]) < 0.2
    assert 0.2 * m["msd_last"] < col("dat/MSD.dat", 1)[-1] < 5 * m["msd_last"]
    dd, dref = col("dat/densDist.dat", 1), np.array(m["densDist"])
    assert abs(dd.sum() - dref.sum()) < 1e-9 * max(1.0, dref.sum()) + 1.5            # same number of boxes per sample (b may differ by one)
    assert abs(np.argmax(dd) - np.argmax(dref)) <= 2
    vd = col("dat/velDist.dat", 1)
    assert abs(vd.sum() - np.sum(m["velDist"])) < 0.05
    summ = tree["dat/summary.dat"].split("\n")
    ref_summ = m["summary"].split("\n")
    assert [l.split("\t")[0] for l in summ] == [l.split("\t")[0] for l in ref_summ]       # labels, byte for byte
    assert summ[3] == ref_summ[3] and summ[1] == ref_summ[1]                             # steps+1 (Q16), N
    gr = col("dat/pairCorr.dat", 1)
    assert abs(np.nansum(gr) - m["pairCorr_sum"]) < 0.15 * m["pairCorr_sum"]


@pytest.mark.gpu
def test_video_frames_carry_the_overlap_hue(jam, tmp_path):
    """print_video (jamming.cpp:855-870; makevid is a compile-time 0 in the reference, APJ_MAKEVID=1 films here): one XYZ
    frame per nSkip steps in the reference's layout (classes/Print.h:129-148), third column = Cell::over, the integer hue
    apj_overlap_hue evaluates for the step of the frame (240 = no overlap, lower = compressed)."""
    N, steps = 400, 300
    env = dict(os.environ, APJ_OUTPUT_ROOT=str(tmp_path), APJ_SEED="7", APJ_MAKEVID="1")
    r = subprocess.run([jam, "vid", "run0", str(N), str(steps), "0.1", "0.3", "1.0"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = open(os.path.join(str(tmp_path), "local_output", "vid", "run0", "vid", "ovito.txt")).read().split("\n")[:-1]
    frames = steps // 100 + 1                                    # t = 0, 100, 200, 300
    assert len(lines) == frames * (N + 2)
    hues = []
    for f in range(frames):
        blk = lines[f * (N + 2):(f + 1) * (N + 2)]
        assert blk[0] == str(N) and blk[1] == "time step comment"
        rows = [l.split("\t") for l in blk[2:]]
        assert all(len(c) == 7 for c in rows) and [int(c[0]) for c in rows] == list(range(N))
        hues.append(np.array([int(c[2]) for c in rows]))
    h = np.concatenate(hues)
    assert h.max() == 240 and h.min() < 236 and np.all(h <= 240)  # rho = 1.0: above jamming, many overlapping pairs
    assert np.mean(h < 240) > 0.3


@pytest.mark.gpu
def test_device_side_init_runs_the_reference_cadence(jam, tmp_path):
    """APJ_DEVICE_INIT=1 (default for N >= 4M): Engine::initCells runs on the device (apj_init_lattice), vector<Cell> is
    never built for the run; the output tree, the cadence and the summary are those of a host-initialised run."""
    N, steps = 2048, 400
    outs = {}
    for mode in ("0", "1"):
        root = tmp_path / ("m" + mode)
        os.makedirs(root)
        env = dict(os.environ, APJ_OUTPUT_ROOT=str(root), APJ_SEED="11", APJ_DEVICE_INIT=mode, APJ_MAKEVID="1")
        r = subprocess.run([jam, "di", "run0", str(N), str(steps), "0.1", "0.3", "0.9"], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[mode] = read_tree(os.path.join(str(root), "local_output", "di", "run0"))
    a, b = outs["0"], outs["1"]
    assert sorted(a) == sorted(b)
    for name in a:
        la, lb = a[name].split("\n"), b[name].split("\n")
        assert len(la) == len(lb), name
    def col(tree, name, k):
        return np.array([float(l.split("\t")[k]) for l in tree[name].split("\n")[:-1]])
    assert np.array_equal(col(a, "dat/order.dat", 0), col(b, "dat/order.dat", 0))                    # same time stamps
    assert abs(col(a, "dat/order.dat", 1).mean() - col(b, "dat/order.dat", 1).mean()) < 0.25         # other random lattice, same physics
    assert 0.3 < col(b, "dat/MSD.dat", 1)[-1] / col(a, "dat/MSD.dat", 1)[-1] < 3.0
    sa, sb = a["dat/summary.dat"].split("\n"), b["dat/summary.dat"].split("\n")
    assert sa[1] == sb[1] and sa[3] == sb[3]                                                        # N, steps + 1
    frames = b["vid/ovito.txt"].split("\n")
    assert frames[0] == str(N) and len(frames) - 1 == (steps // 100 + 1) * (N + 2)
    radii = np.array([float(l.split("\t")[1]) for l in frames[2:N + 2]])
    assert abs(radii.mean() - 1.0) < 0.02 and abs(radii.std() - 0.1) < 0.02                           # the records were materialised from the device


@pytest.mark.gpu
def test_engine_mirrors_expose_reference_state(tmp_path):
    """Cell / Box mirrors: pull_cells, pull_cell_lists, pull_verlet_lists give the reference's views."""
    prog = r'''
#define main jam_main
#include "jam/jamming.cpp"
#undef main
int main() {
    gen.seed(5); g_seed = 5;
    Engine e("mirror", "run0", 1024, 100, 0.1, 0.5, 0.9);
    e.initCells(); e.topology(); e.attach_device();
    e.assignCellsToGrid(); e.buildVerletLists();
    for (int k = 0; k < 25; k++) e.calculate_next_positions();
    e.assignCellsToGrid(); e.buildVerletLists(); e.pull_cells(); e.pull_cell_lists(); e.pull_verlet_lists();
    long inlist = 0, pairs = 0, bad = 0;
    for (int p = 0; p < e.nbox; p++) for (int i : e.grid[p].CellList) { inlist++; if (e.cell[i].box != p) bad++; }
    for (int i = 0; i < e.N; i++) for (int j : e.cell[i].VerletList) {
        pairs++;
        double dx = e.delta_norm(e.cell[j].x[0] - e.cell[i].x[0]), dy = e.delta_norm(e.cell[j].x[1] - e.cell[i].x[1]);
        if (!(j > i) || !(dx*dx + dy*dy < e.rs2)) bad++;
    }
    printf("%ld %ld %ld %d %g %g\n", inlist, pairs, bad, e.b, e.COM[0], e.calculateOrderParameter());
    return bad != 0;
}'''
    src = tmp_path / "mirror.cpp"
    src.write_text(prog)
    from active_particle_jamming_b200 import _build
    exe = str(tmp_path / "mirror")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", HOST, str(src), "-o", exe, "-L" + os.path.dirname(_build.LIB), "-lapj_b200",
                    "-Wl,-rpath," + os.path.dirname(_build.LIB)], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    inlist, pairs, bad = map(int, r.stdout.split()[:3])
    assert inlist == 1024 and bad == 0 and 7.0 < pairs / 1024 < 9.0      # half-list length ~7.9 at phi = 0.9


def test_sweep_rejects_malformed_input(jam, tmp_path):
    bad = tmp_path / "input.txt"
    bad.write_text("sw run1 256 100 0.1 0.5\n")                       # six fields instead of seven
    r = subprocess.run([jam, "--sweep", str(bad)], capture_output=True, text=True, env=dict(os.environ, APJ_OUTPUT_ROOT=str(tmp_path)))
    assert r.returncode == 2 and "expected <fullRunID> <runID> <N> <steps>" in r.stdout
    r = subprocess.run([jam, "--sweep", str(tmp_path / "missing.txt")], capture_output=True, text=True)
    assert r.returncode == 2 and "cannot open" in r.stdout


@pytest.mark.gpu
def test_sweep_runs_input_lines_as_batched_replicas(jam, tmp_path):
    """jam --sweep input.txt: the reference's job array (create-arrays.sh / jamming.sh, one process per line) as
    replicas of one device handle. Every line gets the full reference output tree; the physics of each replica
    follows ITS parameters (order parameter high at low noise, low at high noise), and a replica of a batch agrees
    statistically with the same line run alone."""
    g = json.load(open(os.path.join(GOLD, "run_shape.json")))
    lines = [("runA", 0.3, 0.10, 0.9), ("runB", 0.3, 0.95, 0.9), ("runC", 0.1, 0.10, 0.84), ("runD", 0.3, 0.10, 0.9)]
    inp = tmp_path / "input.txt"
    inp.write_text("".join("sweep %s 1024 2000 %g %g %g\n" % l for l in lines) + "sweep runE 576 1000 0.3 0.5 0.9\n")
    env = dict(os.environ, APJ_OUTPUT_ROOT=str(tmp_path), APJ_SEED="77")
    r = subprocess.run([jam, "--sweep", str(inp), "3"], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("batch of") == 3                              # 3 + 1 (batch size) + 1 (other shape)
    order = {}
    for name, l_s, l_n, rho in lines + [("runE", 0.3, 0.5, 0.9)]:
        tree = read_tree(os.path.join(str(tmp_path), "local_output", "sweep", name))
        assert sorted(tree) == sorted(g["shape"]), name                 # the 14 files of a reference run
        col = np.array([float(l.split("\t")[1]) for l in tree["dat/order.dat"].split("\n")[:-1]])
        assert len(col) == (1000 if name == "runE" else 2000) // 100 + 1
        order[name] = col[len(col) // 2:].mean()
    assert order["runA"] > 0.8 and order["runD"] > 0.8 and order["runB"] < 0.35      # each replica follows its own noise
    assert abs(order["runA"] - order["runD"]) < 0.15                                 # same point, different batch: same physics


@pytest.mark.gpu
def test_sweep_skips_a_bad_line_instead_of_aborting_the_batch(jam, tmp_path):
    """one impossible line (a box of fewer than 3 x 3 cells) used to exit(720) the whole per-GPU sweep
    process. It is reported and skipped on its own; its batch mates run; the exit status says that lines were skipped."""
    inp = tmp_path / "input.txt"
    inp.write_text("sw good1 256 300 0.3 0.5 0.9\nsw tiny 16 300 0.3 0.5 0.9\nsw good2 256 300 0.3 0.5 0.9\nsw neg 256 300 0.3 0.5 -1\n")
    env = dict(os.environ, APJ_OUTPUT_ROOT=str(tmp_path), APJ_SEED="5")
    r = subprocess.run([jam, "--sweep", str(inp), "8"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "input.txt:2: skipped (box too small" in r.stdout and "input.txt:4: skipped" in r.stdout and "2 line(s)" in r.stdout
    for name in ("good1", "good2"):
        assert os.path.exists(os.path.join(str(tmp_path), "local_output", "sw", name, "dat", "summary.dat"))
    assert not os.path.exists(os.path.join(str(tmp_path), "local_output", "sw", "tiny"))


def test_host_topology_and_lattice_match_the_reference_algorithm(tmp_path):
    """Engine::initCells / Engine::topology are host-side setup (no device): the grid (b, lp, the 3x3 periodic
    neighbour table in the reference's numbering, jamming.cpp:356-410) must equal the oracle's for the same L, and
    the lattice must follow the reference's recipe (:285-354): radii 1 + N(0,1)/10, L from the packing fraction with
    the truncated PI, offset rows, all particles inside the box."""
    from oracle.pyoracle import OracleSim
    prog = r'''
#define main jam_main
#include "jam/jamming.cpp"
#undef main
int main() {
    gen.seed(11); g_seed = 11;
    Engine e("probe", "run0", 1024, 100, 0.1, 0.5, 0.9);
    e.initCells(); e.topology();
    printf("%.17g %d %d %.17g\n", e.L, e.b, e.nbox, e.lp);
    for (int p = 0; p < e.nbox; p++) { for (int k = 0; k < 9; k++) printf("%d ", e.grid[p].neighbors[k]); printf("\n"); }
    double sr = 0, sr2 = 0, area = 0; int outside = 0;
    for (int i = 0; i < e.N; i++) {
        const Cell& c = e.cell[i];
        sr += c.R; sr2 += c.R*c.R; area += c.R*c.R;
        if (c.x[0] < -e.Lover2 || c.x[0] >= e.Lover2 || c.x[1] < -e.Lover2 || c.x[1] >= e.Lover2) outside++;
        if (fabs(c.cosp - cos(c.phi)) > 1e-15 || fabs(c.Rinv*c.R - 1.0) > 1e-15) outside += 1000;
    }
    printf("%.17g %.17g %.17g %d\n", sr/e.N, sr2/e.N, area, outside);
    return 0;
}'''
    src = tmp_path / "topo.cpp"
    src.write_text(prog)
    from active_particle_jamming_b200 import _build
    exe = str(tmp_path / "topo")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", HOST, str(src), "-o", exe, "-L" + os.path.dirname(_build.LIB), "-lapj_b200",
                    "-Wl,-rpath," + os.path.dirname(_build.LIB)], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().split("\n")
    L, b, nbox, lp = float(lines[0].split()[0]), int(lines[0].split()[1]), int(lines[0].split()[2]), float(lines[0].split()[3])
    nb = np.array([[int(v) for v in l.split()] for l in lines[1:1 + nbox]])
    mean_r, mean_r2, area, outside = lines[1 + nbox].split()
    o = OracleSim(1024, L, 0.9)
    o.topology()
    sc = o.scalars()
    assert (b, nbox) == (int(sc["b"]), int(sc["nbox"])) and lp == sc["lp"]
    assert np.array_equal(nb, o.box_neighbors())                       # same table, same numbering
    o.close()
    assert L == np.sqrt(3.14159265 * float(area) / 0.9)                # jamming.cpp:305, truncated PI
    assert abs(float(mean_r) - 1.0) < 0.02 and abs(float(mean_r2) - 1.01) < 0.03 and int(outside) == 0
