#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r2i}
timeout 1800 python -m pytest tests -m gpu -q --durations=4 > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 ${O}_pytest.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$?"; python scripts/bench_brief.py ${O}_bench_$name.json; tail -2 ${O}_bench_$name.err; }
b box2m --particles 2097152 --no-cpu --no-e2e
for l in 1 2; do APJ_LANES=$l b jam65k_l$l --workload jam65k --steps 10000 --warmup 200 --no-cpu --no-e2e; done
for l in 1 2 4 8; do APJ_LANES=$l b jam1k_l$l --workload jam1k --steps 20000 --warmup 200 --no-cpu --no-e2e; done
for n in 16384 131072 524288; do for l in 1 2 4; do APJ_LANES=$l b n${n}_l$l --particles $n --no-relax --steps 2000 --warmup 200 --no-cpu --no-e2e; done; done
T=/tmp/ncu_$$; mkdir -p $T
cap() { out=$1; k=$2; skip=$3; cnt=$4; shift 4; timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt --kill 1 -o $T/$out -f "$@" > ${O}_ncu_$out.log 2>&1; rc=$?; python scripts/ncu_summary.py $T/$out.ncu-rep 14 > ${O}_${out}_ncu_full.txt 2>&1; echo "ncu $out rc=$rc $(head -1 ${O}_${out}_ncu_full.txt)"; rm -f $T/$out.ncu-rep; }
cap slab_apj_step_kernel apj_step_kernel 500 3 python scripts/slab_one_gpu.py 2097152 1 600
rm -rf $T
