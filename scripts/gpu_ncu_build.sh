#!/bin/bash
# full ncu capture of the Verlet build kernel (first 24 launches: the active ones are a few of them)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apj_verlet_build -c 10 -o gpurun_out/build_full python bench.py --no-relax --no-cpu --steps 100 --warmup 20 > gpurun_out/ncu_build.log 2>&1; echo "ncu build rc=$?"
