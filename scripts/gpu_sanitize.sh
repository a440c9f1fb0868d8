#!/bin/bash
# compute-sanitizer over the parity tests: memcheck on everything but the full-size cases and the
# multi-rank slab tests (several ranks on one device need concurrent kernels; the sanitizer serialises them),
# racecheck on the golden cases
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py -m gpu -q -k "not full_size" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck.log | tail -3
timeout 120 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_golden.py -m gpu -q -x -k "pair_set or single_step" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | tail -3
