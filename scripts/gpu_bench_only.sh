#!/bin/bash
# bench only (no tests), default library + $VARIANTS
mkdir -p gpurun_out
for v in libapj_b200.so $VARIANTS; do
  APJ_B200_LIB=$v timeout 600 python bench.py --no-cpu --steps 1000 --warmup 200 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "$v rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$v.json")); print("$v", "value %.4e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel_ms %.4f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"])
except Exception as e: print("$v failed", e, open("gpurun_out/bench_$v.err").read()[-500:])
PY
done
