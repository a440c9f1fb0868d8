#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r2n}
timeout 1800 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 ${O}_pytest.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$?"; python scripts/bench_brief.py ${O}_bench_$name.json; tail -2 ${O}_bench_$name.err; }
b base --no-cpu --no-e2e
b hot --workload box16m_hot --steps 200 --warmup 50 --no-cpu --no-e2e
T=/tmp/ncu_$$; mkdir -p $T
timeout 300 ncu --set full --clock-control none --import-source on -k regex:apj_verlet_build_kernel -s 3 -c 1 --kill 1 -o $T/vb -f python bench.py --no-relax --no-cpu --no-e2e --steps 64 --warmup 16 > ${O}_ncu_vb.log 2>&1; python scripts/ncu_summary.py $T/vb.ncu-rep 10 > ${O}_apj_verlet_build_kernel_ncu_full.txt 2>&1; head -24 ${O}_apj_verlet_build_kernel_ncu_full.txt
rm -rf $T
