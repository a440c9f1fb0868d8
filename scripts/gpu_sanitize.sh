#!/bin/bash
# compute-sanitizer over the parity tests: memcheck on everything but the full-size cases and the
# multi-rank slab tests (several ranks on one device need concurrent kernels; the sanitizer serialises them),
# racecheck on the golden cases, the setup kernels and the pipelined step kernel
mkdir -p gpurun_out
S="compute-sanitizer --print-limit 10"
timeout 700 $S --tool memcheck python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_setup.py -m gpu -q -k "not full_size and not headline and not inhomogeneous" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck.log | tail -3
timeout 300 $S --tool memcheck python -m pytest tests/test_gpu_slab.py -m gpu -q -k "bit_identical and 1-4-4096 or one_rank" > gpurun_out/sanitize_memcheck_slab.log 2>&1; echo "memcheck slab rc=$?"; grep -E "ERROR SUMMARY|passed|failed|deselected" gpurun_out/sanitize_memcheck_slab.log | tail -3
timeout 300 $S --tool racecheck python -m pytest tests/test_gpu_golden.py tests/test_gpu_setup.py -m gpu -q -x -k "pair_set or single_step or hue or lattice" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | tail -3
APJ_STEP_PIPE=1 timeout 300 $S --tool racecheck python -m pytest tests/test_gpu_golden.py -m gpu -q -x -k "single_step or free_running" > gpurun_out/sanitize_racecheck_pipe.log 2>&1; echo "racecheck pipe rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck_pipe.log | tail -3
timeout 300 $S --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "queued or mark_origin_and_observables or cutoff_140" > gpurun_out/sanitize_racecheck_obs.log 2>&1; echo "racecheck obs rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck_obs.log | tail -3
