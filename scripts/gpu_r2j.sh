#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r2j}
b() { name=$1; shift; timeout 900 python bench.py "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$?"; python scripts/bench_brief.py ${O}_bench_$name.json; python -c "
import json,sys
d=json.loads([l for l in open('${O}_bench_$name.json') if l.startswith('{')][-1]); print('   tuning', d['config']['tuning'])"; tail -2 ${O}_bench_$name.err; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -m gpu -x -q -k "split_tail or free_running or skin_aware or step_by_step" > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
b base --no-cpu --no-e2e
APJ_B200_LIB=libapj_early.so b early --no-cpu --no-e2e
APJ_B200_LIB=libapj_b5.so APJ_TILE_SLACK=0 b b5 --no-cpu --no-e2e
APJ_B200_LIB=libapj_b5.so b b5_slack --no-cpu --no-e2e
