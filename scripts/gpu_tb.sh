#!/bin/bash
# 128- against 256-particle work blocks (APJ_TB) for the one-lane-per-particle kernel, every workload that uses it
mkdir -p gpurun_out
O=gpurun_out/tb
b() { name=$1; shift; for tb in 256 128; do APJ_TB=$tb timeout 600 python bench.py "$@" > ${O}_${name}_$tb.json 2> ${O}_${name}_$tb.err; echo "$name tb=$tb rc=$?"; python scripts/bench_brief.py ${O}_${name}_$tb.json; done; }
b box16m --no-cpu --steps 1000 --warmup 100
b hot --workload box16m_hot --no-cpu --no-e2e --steps 200 --warmup 50
b obs1m --workload obs1m --steps 10000 --warmup 200 --no-cpu --no-e2e
b sweep512 --workload sweep512 --steps 2000 --warmup 200 --no-cpu --no-e2e
