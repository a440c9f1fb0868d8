#!/bin/bash
# A/B: default library vs an experimental variant, tests on the default only
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for v in libapj_b200.so $VARIANTS; do
  APJ_B200_LIB=$v timeout 600 python bench.py --no-cpu --steps 1000 --warmup 200 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "$v rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$v.json")); print("$v", "value %.4e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel_ms %.4f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"], d["config"]["tuning"], "e2e %.3e"%d["e2e"]["value"])
except Exception as e: print("$v failed", e, open("gpurun_out/bench_$v.err").read()[-500:])
PY
done
