#!/bin/bash
# ncu of the step kernel on the launch-bound workload (N = 1024)
mkdir -p gpurun_out; T=/tmp/ncu_$$; mkdir -p $T
O=gpurun_out/small
timeout 300 ncu --set full --clock-control none --import-source on -k regex:apj_step_kernel -s 3000 -c 2 --kill 1 -o $T/nc -f python bench.py --workload jam1k --no-cpu --no-e2e --steps 2000 --warmup 200 > ${O}_ncu.log 2>&1; python scripts/ncu_summary.py $T/nc.ncu-rep 40 > ${O}_jam1k_step_ncu_full.txt 2>&1; head -64 ${O}_jam1k_step_ncu_full.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 120 --csv --log-file ${O}_jam1k_launches.csv python bench.py --workload jam1k --no-cpu --no-e2e --steps 2000 --warmup 200 > /dev/null 2>&1; python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("${O}_jam1k_launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4]].append(float(r[-1].replace(',','')))
for k,v in agg.items(): print(k[:60], len(v), sum(v)/len(v))
PY
rm -rf $T
