#!/bin/bash
# same-box A/B of library variants on the launch-bound workloads: VARIANTS="libapj_b200.so libapj_x.so"
mkdir -p gpurun_out
O=gpurun_out/ab2
for rep in 1 2; do for v in $VARIANTS; do
  for w in "jam1k --steps 50000" "sweep512 --steps 2000"; do set -- $w
    APJ_B200_LIB=$v timeout 600 python bench.py --workload $1 $2 $3 --warmup 200 --no-cpu --no-e2e > ${O}_$1_$v.json 2> ${O}_$1_$v.err; echo -n "$v $1 rep$rep: "; python scripts/bench_brief.py ${O}_$1_$v.json | cut -c1-60
  done; done; done
