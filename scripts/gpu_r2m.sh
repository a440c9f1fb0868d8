#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r2m}
b() { name=$1; shift; timeout 900 python bench.py "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$?"; python scripts/bench_brief.py ${O}_bench_$name.json; python -c "
import json,sys
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print('   tuning', d['config']['tuning'])"; tail -2 ${O}_bench_$name.err; }
APJ_B200_LIB=libapj_tb512.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size or split_tail or skin_aware" > ${O}_pytest.log 2>&1; echo "pytest tb512 rc=$?"; tail -3 ${O}_pytest.log
b base --no-cpu --no-e2e
APJ_B200_LIB=libapj_tb512.so b tb512 --no-cpu --no-e2e
T=/tmp/ncu_$$; mkdir -p $T
APJ_B200_LIB=libapj_tb512.so timeout 300 ncu --set full --clock-control none -k regex:apj_step_kernel -s 150 -c 1 --kill 1 -o $T/nc -f python bench.py --no-relax --no-cpu --no-e2e --steps 64 --warmup 16 > ${O}_ncu_tb512.log 2>&1; python scripts/ncu_summary.py $T/nc.ncu-rep 6 > ${O}_tb512_step_ncu_full.txt 2>&1; head -32 ${O}_tb512_step_ncu_full.txt
rm -rf $T
