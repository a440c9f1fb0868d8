#!/bin/bash
# tests, then full ncu captures of the step kernel (steady state) and of one active Verlet build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apj_step_kernel -s 400 -c 2 -o gpurun_out/step_full python bench.py --no-relax --no-cpu --steps 200 --warmup 100 > gpurun_out/ncu_full.log 2>&1; echo "ncu step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apj_verlet_build -s 3 -c 1 -o gpurun_out/build_full python bench.py --no-relax --no-cpu --steps 20 --warmup 10 > gpurun_out/ncu_build.log 2>&1; echo "ncu build rc=$?"
