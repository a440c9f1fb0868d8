#!/bin/bash
# A/B of the pipelined persistent step kernel (APJ_STEP_PIPE=1) against the classic one
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r2c}
APJ_DEBUG_LAUNCH=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -m gpu -x -q -k "split_tail or free_running or skin_aware" > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 ${O}_pytest.log; grep "apj_b200: launch" ${O}_pytest.log | head -5
for pipe in 0 1; do
  APJ_STEP_PIPE=$pipe timeout 600 python bench.py --no-cpu --no-e2e > ${O}_bench_pipe$pipe.json 2> ${O}_bench_pipe$pipe.err; echo "bench pipe=$pipe rc=$?"; python scripts/bench_brief.py ${O}_bench_pipe$pipe.json; tail -3 ${O}_bench_pipe$pipe.err
done
if [ "$1" = "prof" ]; then
for pipe in ${PROF_PIPES:-1}; do
  APJ_STEP_PIPE=$pipe timeout 400 ncu --set full --clock-control none --import-source on -k regex:apj_step -s 150 -c 1 --kill 1 -o ${O}_step_pipe$pipe -f python bench.py --no-relax --no-cpu --no-e2e --steps 64 --warmup 16 > ${O}_ncu_pipe$pipe.log 2>&1; echo "ncu pipe=$pipe rc=$?"
done
fi
