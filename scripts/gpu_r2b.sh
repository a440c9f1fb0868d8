#!/bin/bash
# Round 2, visit B: GPU suite, bench (default), launch list, ncu of verlet_build + step kernel
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r2b}
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 ${O}_pytest.log
timeout 600 python bench.py --no-cpu > ${O}_bench_k1000.json 2> ${O}_bench_k1000.err; echo "bench k1000 rc=$?"; python scripts/bench_brief.py ${O}_bench_k1000.json; tail -3 ${O}_bench_k1000.err
timeout 600 python bench.py --no-cpu --steps 20 --warmup 5 > ${O}_bench_k20.json 2> ${O}_bench_k20.err; echo "bench k20 rc=$?"; python scripts/bench_brief.py ${O}_bench_k20.json
if [ "$1" = "prof" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 900 --csv --log-file ${O}_launches.csv python bench.py --no-relax --no-cpu --no-e2e --steps 200 --warmup 100 > ${O}_under_ncu.log 2>&1; echo "ncu list rc=$?"
for k in ${KERNELS:-apj_verlet_build_kernel apj_step_kernel}; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-0} -c 1 --kill 1 -o ${O}_$k -f python bench.py --no-relax --no-cpu --no-e2e --steps 64 --warmup 16 > ${O}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
fi
