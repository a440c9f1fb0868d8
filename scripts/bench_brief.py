"""One-line digest of a bench.py JSON line (GPU-visit scripts)."""
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])      # NCCL may print its version line first
    r = d["roofline"]; e = d.get("e2e") or {}
    print("value %.4e  ms/step %.4f  kernel_ms %.4f  frac %.3f  rebuild %.2f ms x %d  e2e %.3e (%.3f s)  checksum %s  sweep %s" % (
        d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["rebuild"]["ms_per_rebuild"], d["rebuild"]["rebuilds_in_timed_region"],
        e.get("value", 0), e.get("job_seconds", 0), d["state_checksum"]["value"], d["config"]["sweep"]["steps_per_class"]))
except Exception as ex:
    print("no bench line:", ex)
