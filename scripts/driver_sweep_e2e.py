#!/usr/bin/env python
"""BASELINE config[0] as a job array: 64 (lambda_n, lambda_s) points of N = 1024, 1e5 steps each, through `jam --sweep` (replicas of
ONE device handle, every line with its own reference-format output tree). Prints one JSON line."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 64
with tempfile.TemporaryDirectory() as d:
    os.makedirs(os.path.join(d, "local_output"))
    inp = os.path.join(d, "input.txt")
    with open(inp, "w") as f:
        for k in range(reps):
            f.write("sweep run%d %d %d %g %g 0.9\n" % (k, N, steps, 0.02 + 0.01 * (k % 8), 0.1 + 0.1 * (k // 8)))
    jam = os.path.join(ROOT, "active_particle_jamming_b200", "host", "bin", "jam")
    env = dict(os.environ, APJ_OUTPUT_ROOT=d, APJ_SEED="7")
    t0 = time.perf_counter()
    p = subprocess.run([jam, "--sweep", inp, str(reps)], capture_output=True, text=True, env=env)
    secs = time.perf_counter() - t0
    done = sum(os.path.exists(os.path.join(d, "local_output", "sweep", "run%d" % k, "dat", "summary.dat")) for k in range(reps))
print(json.dumps({"config": "%d runs of N=%d, %d steps (+4000 relax) as replicas of one handle, full reference output tree each" % (reps, N, steps),
                  "seconds": secs, "rc": p.returncode, "runs_completed": done, "seconds_per_run": secs / reps,
                  "particle_steps_per_s": reps * N * (steps + 4001) / secs}))
