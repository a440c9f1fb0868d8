#!/usr/bin/env python
"""BASELINE config[0] end to end: the reference's own Engine::start() (oracle/_ref: init, relax, 1e5 steps with its observable
cadence and its output files, one host core) against the shipped jam driver (host/bin/jam: same CLI, same files, device
behind the C ABI). Prints one JSON line. Usage: driver_e2e.py [N] [steps]"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import RefEngine, have_ref
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
l_s, l_n, rho = 0.05, 0.5, 0.9
out = {"config": "N=%d, %d steps, lambda_s=%g lambda_n=%g rho=%g; full driver run incl. init, relax (2000 + 2000), observables, 14 output files" % (N, steps, l_s, l_n, rho)}
with tempfile.TemporaryDirectory() as d:
    if have_ref():
        RefEngine.seed(7)
        r = RefEngine(N, steps, l_s, l_n, rho)
        out["reference_seconds_one_core"] = r.run_start(os.path.join(d, "ref"))
        r.close()
    os.makedirs(os.path.join(d, "ours", "local_output"))
    jam = os.path.join(ROOT, "active_particle_jamming_b200", "host", "bin", "jam")
    env = dict(os.environ, APJ_OUTPUT_ROOT=os.path.join(d, "ours"), APJ_SEED="7")
    t0 = time.perf_counter()
    p = subprocess.run([jam, "e2e", "run0", str(N), str(steps), str(l_s), str(l_n), str(rho)], capture_output=True, text=True, env=env)
    out["jam_seconds"] = time.perf_counter() - t0
    out["jam_rc"] = p.returncode
if "reference_seconds_one_core" in out:
    out["speedup_vs_one_core"] = out["reference_seconds_one_core"] / out["jam_seconds"]
print(json.dumps(out))
