#!/bin/bash
# speculative steps per launch group (APJ_GROUP): idle launches after a rebuild fires against idle chains at every group end
mkdir -p gpurun_out
O=gpurun_out/grp
for m in 16 8 4 2; do
  APJ_GROUP=$m timeout 600 python bench.py --no-cpu --no-e2e --steps 1000 --warmup 100 > ${O}_box_$m.json 2> ${O}_box_$m.err; echo -n "box16m m=$m: "; python scripts/bench_brief.py ${O}_box_$m.json | cut -c1-110
  APJ_GROUP=$m timeout 600 python bench.py --workload box16m_hot --no-cpu --no-e2e --steps 200 --warmup 50 > ${O}_hot_$m.json 2> ${O}_hot_$m.err; echo -n "hot m=$m: "; python scripts/bench_brief.py ${O}_hot_$m.json | cut -c1-110
done
for m in 16 8; do
  APJ_GROUP=$m timeout 600 python bench.py --workload obs1m --steps 10000 --warmup 200 --no-cpu --no-e2e > ${O}_obs_$m.json 2> ${O}_obs_$m.err; echo -n "obs1m m=$m: "; python scripts/bench_brief.py ${O}_obs_$m.json | cut -c1-110
done
