#!/bin/bash
# One GPU-box visit: smoke, GPU test-suite, bench, launch list, full ncu capture of the step kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "$1" = "prof" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --no-relax --no-cpu --steps 64 --warmup 16 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apj_step_kernel -s 150 -c 2 -o gpurun_out/step_full python bench.py --no-relax --no-cpu --steps 64 --warmup 16 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
