#!/bin/bash
# Round 2, visit A: smoke, GPU suite (incl. the 16M parity test), bench at the driver's flags and at the default, ncu of the chain kernels
mkdir -p gpurun_out
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > ${O}_gpu.txt 2>&1
free -g >> ${O}_gpu.txt; nproc >> ${O}_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 ${O}_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench_k20.json 2> ${O}_bench_k20.err; echo "bench k20 rc=$?"; cat ${O}_bench_k20.json; tail -3 ${O}_bench_k20.err
timeout 600 python bench.py --no-cpu > ${O}_bench_k1000.json 2> ${O}_bench_k1000.err; echo "bench k1000 rc=$?"; cat ${O}_bench_k1000.json; tail -3 ${O}_bench_k1000.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > ${O}_bench_ref.json 2> ${O}_bench_ref.err; echo "bench ref rc=$?"; cat ${O}_bench_ref.json
if [ "$1" = "prof" ]; then
for k in apj_verlet_build_kernel apj_reorder_kernel apj_bin_count_kernel apj_make_tiles_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 --kill 1 -o ${O}_$k -f python bench.py --no-relax --no-cpu --no-e2e --steps 16 --warmup 3 > ${O}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
fi
