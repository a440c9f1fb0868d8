#!/bin/bash
# ring step kernel (APJ_STEP_RING=1; APJ_TB=128: 128-particle tiles) against the classic one: parity tests that run large
# single systems, bench A/B, one ncu capture
mkdir -p gpurun_out
O=gpurun_out/ring
if [ -z "$SKIP_TESTS" ]; then
for tb in 256 128; do
APJ_TB=$tb APJ_STEP_RING=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "${TESTS:-full_size or split_tail}" > ${O}_pytest_$tb.log 2>&1; echo "pytest(ring, tb $tb) rc=$?"; tail -3 ${O}_pytest_$tb.log
done
fi
for v in ${VARIANTS:-256.0 256.1 128.0 128.1}; do
  APJ_TB=${v%.*} APJ_STEP_RING=${v#*.} timeout 300 python bench.py --no-cpu --no-e2e --steps ${STEPS:-300} --warmup 50 > ${O}_bench_$v.json 2> ${O}_bench_$v.err; echo "tb.ring=$v rc=$?"
  python scripts/bench_brief.py ${O}_bench_$v.json; tail -2 ${O}_bench_$v.err
done
if [ -n "$NCU" ]; then
T=/tmp/ncu_$$; mkdir -p $T
APJ_TB=${NCU} APJ_STEP_RING=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:apj_step_ring_kernel -s 150 -c 1 --kill 1 -o $T/nc -f python bench.py --no-cpu --no-e2e --steps 64 --warmup 16 > ${O}_ncu.log 2>&1; python scripts/ncu_summary.py $T/nc.ncu-rep 30 > ${O}_step_ring_ncu_full.txt 2>&1; head -40 ${O}_step_ring_ncu_full.txt
rm -rf $T
fi
