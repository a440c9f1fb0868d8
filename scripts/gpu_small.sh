#!/bin/bash
# small-system path: parity + the launch-bound bench lines
mkdir -p gpurun_out
O=gpurun_out/small
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_statistics.py -m gpu -q -x -k "not full_size and not headline" > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
b() { name=$1; shift; timeout 600 python bench.py "$@" > ${O}_${name}.json 2> ${O}_${name}.err; echo "$name rc=$?"; python scripts/bench_brief.py ${O}_${name}.json; }
b jam1k --workload jam1k --steps 100000 --warmup 200 --no-cpu --no-e2e
b jam65k --workload jam65k --steps 10000 --warmup 200 --no-cpu --no-e2e
b sweep512 --workload sweep512 --steps 2000 --warmup 200 --no-cpu --no-e2e
APJ_TB=192 b sweep512_tb192 --workload sweep512 --steps 2000 --warmup 200 --no-cpu --no-e2e
APJ_TB=128 b sweep512_tb128 --workload sweep512 --steps 2000 --warmup 200 --no-cpu --no-e2e

