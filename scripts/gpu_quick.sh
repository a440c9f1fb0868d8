#!/bin/bash
# tests + a short bench (no CPU leg) [+ launch list with "list"]
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu $BENCH_ARGS > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
if [ "$1" = "list" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --no-relax --no-cpu --steps 200 --warmup 100 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
fi
