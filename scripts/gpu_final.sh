#!/bin/bash
# Final single-GPU visit: what the driver runs (GPU suite, smoke, bench at its flags) + the default bench + launch list
mkdir -p gpurun_out
O=gpurun_out/${TAG:-fin}
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 ${O}_smoke.log
timeout 1800 python -m pytest tests -m gpu -q --durations=4 > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 ${O}_pytest.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$?"; python scripts/bench_brief.py ${O}_bench_$name.json; tail -2 ${O}_bench_$name.err; }
b ref --impl reference --gpus 1 --steps 20 --warmup 5
b k20 --gpus 1 --steps 20 --warmup 5
b k1000
b obs1m --workload obs1m --steps 10000 --warmup 200 --no-cpu
b jam65k --workload jam65k --steps 10000 --warmup 200 --no-cpu
b sweep512 --workload sweep512 --steps 2000 --warmup 200 --no-cpu
b jam1k --workload jam1k --steps 100000 --warmup 200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 900 --csv --log-file ${O}_launches.csv python bench.py --no-relax --no-cpu --no-e2e --steps 200 --warmup 100 > ${O}_under_ncu.log 2>&1; echo "ncu list rc=$?"
