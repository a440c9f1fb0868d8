#!/bin/bash
# multi-GPU visit: bit-identity check of the slab path against the single-GPU engine, then bench.py at N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/slab_dist_check.py > gpurun_out/slab_check_$N.log 2>&1; echo "check rc=$?"; tail -3 gpurun_out/slab_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1000 --warmup 200 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n$N.json | cut -c1-1500; tail -3 gpurun_out/bench_n$N.err
