#!/bin/bash
# Scaling run on one 8-GPU box: the same 16M box at 1, 2, 4, 8 GPUs (slabs), one JSON line each; the
# state_checksum of the four lines must be identical. Run as: gpurun --gpus 8 -- bash scripts/gpu_scale.sh
mkdir -p gpurun_out
O=gpurun_out/${TAG:-scale}
nvidia-smi topo -m > ${O}_topo.txt 2>&1
run() { n=$1; tag=$2; shift 2
  if [ "$n" = "1" ]; then timeout 900 python bench.py --gpus 1 "$@" > ${O}_${tag}.json 2> ${O}_${tag}.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > ${O}_${tag}.json 2> ${O}_${tag}.err; fi
  echo "n=$n $tag rc=$?"; python scripts/bench_brief.py ${O}_${tag}.json; tail -2 ${O}_${tag}.err | cut -c1-300; }
for n in ${NS:-1 2 4 8}; do run $n n$n --no-cpu ${BENCH_ARGS:---steps 1000 --warmup 200}; done
if [ "$1" = "pipe" ]; then
  for n in 2 8; do APJ_STEP_PIPE=1 run $n pipe_n$n --no-cpu --no-e2e ${BENCH_ARGS:---steps 1000 --warmup 200}; done
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("${O}_n*.json") + glob.glob("${O}_pipe_n*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f.split("/")[-1], "gpus", d["n_gpus"], "value %.4e" % d["value"], "checksum", d["state_checksum"]["value"], "reset", d["state_checksum"]["resetCounter"],
              "kernel_ms %.4f" % r["kernel_ms"], "parts", {k: round(v, 4) for k, v in r["parts_ms"].items() if k != "how"}, "e2e %.3e" % (d.get("e2e") or {}).get("value", 0))
    except Exception as e:
        print(f, "failed", e)
PY
