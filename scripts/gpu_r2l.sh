#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r2l}
b() { name=$1; shift; timeout 900 python bench.py "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$?"; python scripts/bench_brief.py ${O}_bench_$name.json; tail -2 ${O}_bench_$name.err; }
b base --no-cpu --no-e2e
APJ_B200_LIB=libapj_bb4.so b bb4 --no-cpu --no-e2e
APJ_B200_LIB=libapj_nocls.so b nocls --no-cpu --no-e2e
T=/tmp/ncu_$$; mkdir -p $T
APJ_B200_LIB=libapj_nocls.so timeout 300 ncu --set full --clock-control none -k regex:apj_step_kernel -s 150 -c 1 --kill 1 -o $T/nc -f python bench.py --no-relax --no-cpu --no-e2e --steps 64 --warmup 16 > ${O}_ncu_nocls.log 2>&1; python scripts/ncu_summary.py $T/nc.ncu-rep 6 > ${O}_nocls_step_ncu_full.txt 2>&1; head -24 ${O}_nocls_step_ncu_full.txt
timeout 300 ncu --set full --clock-control none -k regex:apj_verlet_build_kernel -s 3 -c 1 --kill 1 -o $T/vb -f python bench.py --no-relax --no-cpu --no-e2e --steps 64 --warmup 16 > ${O}_ncu_vb.log 2>&1; python scripts/ncu_summary.py $T/vb.ncu-rep 6 > ${O}_apj_verlet_build_kernel_ncu_full.txt 2>&1; head -12 ${O}_apj_verlet_build_kernel_ncu_full.txt
rm -rf $T
