#!/bin/bash
# per-instruction shared-memory wavefronts of the step kernel's gathers (16-byte {x,y} / {cos,sin} against 8-byte R)
mkdir -p gpurun_out; T=/tmp/ncu_$$; mkdir -p $T
timeout 300 ncu --set full --clock-control none --import-source on -k regex:apj_step_kernel -s 150 -c 1 --kill 1 -o $T/nc -f python bench.py --no-cpu --no-e2e --steps 64 --warmup 16 > gpurun_out/lds_ncu.log 2>&1
python scripts/ncu_lds.py $T/nc.ncu-rep > gpurun_out/lds_step_kernel.txt 2>&1; cat gpurun_out/lds_step_kernel.txt
rm -rf $T
