#!/bin/bash
# one full ncu capture of the step kernel in steady state (library selected by APJ_B200_LIB)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apj_step_kernel -s 400 -c 2 -o gpurun_out/step_full_${TAG:-cur} python bench.py --no-relax --no-cpu --steps 200 --warmup 100 > gpurun_out/ncu_full_${TAG:-cur}.log 2>&1; echo "ncu rc=$?"
