"""active_particle_jamming_b200.sweep: the reference's job array (input.txt, one line per phase-diagram point)
dealt out over the GPUs of a node, one `jam --sweep` process per GPU. Checked here with a stand-in for the jam
binary that records what it was given (the real driver is exercised on the GPU in tests/test_host_driver.py)."""
import os
import stat
import subprocess
import sys

import pytest

from active_particle_jamming_b200 import sweep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_input(path, spec):
    with open(path, "w") as f:
        k = 0
        for n, steps, count in spec:
            for _ in range(count):
                k += 1
                f.write("grid run%d %d %d 0.05 0.5 0.9\n" % (k, n, steps))
        f.write("\n")
    return k


def test_partition_balances_and_keeps_shapes_contiguous(tmp_path):
    p = str(tmp_path / "input.txt")
    total = write_input(p, [(4096, 10000, 100), (1024, 5000, 30), (100000, 2000, 3)])
    runs = sweep.parse_input(p)
    assert len(runs) == total
    shares = sweep.partition(runs, 8, 64)
    assert sorted(l for s in shares for l in s) == sorted(t for t, _, _ in runs)          # every line exactly once
    load = [sum(int(l.split()[2]) * int(l.split()[3]) for l in s) for s in shares]
    assert max(load) <= 1.6 * (sum(load) / 8)                                             # balanced by N x steps
    for s in shares:                                                                      # same-shape lines are contiguous
        shapes = [tuple(l.split()[2:4]) for l in s]
        changes = sum(1 for a, b in zip(shapes, shapes[1:]) if a != b)
        assert changes <= len(set(shapes)) - 1 + 2
    assert all(len(s) > 0 for s in sweep.partition(runs, 2, 64))
    one = sweep.partition(runs[:3], 8, 64)
    assert sum(1 for s in one if s) == 3                                                  # fewer runs than GPUs


def test_malformed_input_is_rejected(tmp_path):
    p = tmp_path / "input.txt"
    p.write_text("grid run1 4096 1000 0.05 0.5\n")
    with pytest.raises(ValueError):
        sweep.parse_input(str(p))
    assert sweep.main([str(p), "--gpus", "2"]) == 2
    assert sweep.main([str(tmp_path / "nope.txt")]) == 2


def test_one_process_per_gpu_with_its_share(tmp_path):
    fake = tmp_path / "fakejam"
    log = tmp_path / "log"
    os.makedirs(log)
    fake.write_text("#!/bin/bash\n"
                    "# stand-in for host/bin/jam: record device, seed, output root, arguments and the lines received\n"
                    "out=%s/dev$APJ_DEVICE\n"
                    "echo \"$APJ_DEVICE $APJ_SEED $APJ_OUTPUT_ROOT $1 $3\" > $out.env\n"
                    "cp \"$2\" $out.lines\n"
                    "[ \"$APJ_DEVICE\" = \"2\" ] && exit 7\n"
                    "exit 0\n" % log)
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    p = str(tmp_path / "input.txt")
    total = write_input(p, [(4096, 1000, 40)])
    r = subprocess.run([sys.executable, "-m", "active_particle_jamming_b200.sweep", p, "--gpus", "4", "--batch", "16",
                        "--output-root", str(tmp_path), "--seed", "100", "--jam", str(fake)],
                       capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert r.returncode == 7, r.stdout + r.stderr                                         # the failing GPU's status is reported
    seen = []
    for g in range(4):
        dev, seed, root, flag, batch = open(str(log / ("dev%d.env" % g))).read().split()
        assert (dev, seed, root, flag, batch) == (str(g), str(100 + g), str(tmp_path), "--sweep", "16")
        seen += [l.strip() for l in open(str(log / ("dev%d.lines" % g))) if l.strip()]
        assert ("gpu %d: 10 runs, exit status %d" % (g, 7 if g == 2 else 0)) in r.stdout
    assert len(seen) == total and len(set(seen)) == total
