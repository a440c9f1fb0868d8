#!/bin/bash
# larger launch groups for the launch-bound sizes (same box)
mkdir -p gpurun_out
O=gpurun_out/grp2
for m in 16 32 64; do
  for w in "jam1k --steps 50000" "jam65k --steps 10000" "sweep512 --steps 2000"; do set -- $w
    APJ_GROUP=$m timeout 600 python bench.py --workload $1 $2 $3 --warmup 200 --no-cpu --no-e2e > ${O}_$1_$m.json 2> ${O}_$1_$m.err; echo -n "$1 m=$m: "; python scripts/bench_brief.py ${O}_$1_$m.json | cut -c1-100
  done; done
