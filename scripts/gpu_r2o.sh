#!/bin/bash
# GPU suite on the build with cached control reads + short-group graphs, then the driver end-to-end numbers
mkdir -p gpurun_out /tmp/o/local_output
O=gpurun_out/r2o
timeout 1800 python -m pytest tests -m gpu -q -x --durations=4 > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 ${O}_pytest.log
for k in 1 2; do APJ_TIMING=1 APJ_OUTPUT_ROOT=/tmp/o APJ_SEED=7 active_particle_jamming_b200/host/bin/jam e2e run$k 1024 100000 0.05 0.5 0.9 2>&1 >/dev/null | grep timing; done
python scripts/driver_e2e.py 1024 100000 > ${O}_drv_1k.json 2>/dev/null; cat ${O}_drv_1k.json
python scripts/driver_e2e.py 65536 2000 > ${O}_drv_65k.json 2>/dev/null; cat ${O}_drv_65k.json
python scripts/driver_sweep_e2e.py 1024 100000 64 > ${O}_drv_sweep.json 2>&1; cat ${O}_drv_sweep.json
timeout 600 python bench.py --workload jam1k --steps 100000 --warmup 200 --no-cpu > ${O}_bench_jam1k.json 2>/dev/null; python scripts/bench_brief.py ${O}_bench_jam1k.json
