#!/usr/bin/env python
"""Multi-GPU check of the slab path, run under torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/slab_dist_check.py [--particles 200000] [--steps 300]

Every rank steps its slab of ONE periodic box (peer memory over NVLink through cudaIpc mappings);
rank 0 also steps the same box on the periodic single-GPU engine. The two must agree BIT FOR BIT
(positions, orientations, rebuild count, neighbour-pair set). Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=200000)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--lambda_s", type=float, default=0.5)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from active_particle_jamming_b200 import DeviceEngine
    from active_particle_jamming_b200.slab import DistSlab
    from _util import random_system
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    N, rho, seed = a.particles, 0.9, 31
    R, L, x, y, phi = random_system(N, rho, seed)
    box = DistSlab(N, L, device=local, seed=seed, max_neighbors=64, lanes_per_particle=1)
    ref = DeviceEngine(N, L, device=local, seed=seed, max_neighbors=64, lanes_per_particle=1) if rank == 0 else None
    ok, report = True, {}
    for q in (box, ref):
        if q is None:
            continue
        q.set_activity(0.0, 0.3)
        q.upload(x=x, y=y, R=R, phi=phi)
        q.step(40)
        q.set_activity(a.lambda_s, 0.3)
        q.step(a.steps)
    own = [None] * world
    dist.all_gather_object(own, box.local[0].info()["n_own"])
    got = box.download(["x", "y", "cosp", "sinp", "x_real", "y_real", "x_old", "y_old"])
    pairs = box.pair_set()
    order = box.order_orientation()[0][0]
    msd = box.msd()[0]
    chk = box.checksum()
    corr = box.spatial_correlations(20.0)                 # edge sets travel between the ranks (NCCL p2p), pairs across edges counted once
    # per-rank checkpoint files, restored into a fresh set of handles
    import tempfile
    prefix = os.path.join(tempfile.gettempdir(), "apj_slab_ckpt_%d" % os.getppid())
    box.save_checkpoint(prefix)
    box2 = DistSlab(N, L, device=local, seed=1, max_neighbors=64, lanes_per_particle=1)
    box2.load_checkpoint(prefix)
    chk2 = box2.checksum()
    step2 = box2.counters()["step"]
    box2.close()
    try:
        os.remove("%s.rank%dof%d" % (prefix, rank, world))
    except OSError:
        pass
    if rank == 0:
        want = ref.download(["x", "y", "cosp", "sinp", "x_real", "y_real", "x_old", "y_old"])
        same = {k: bool(np.array_equal(got[k], want[k])) for k in want}
        cb, cr = box.counters(), ref.counters()
        report = {"check": "slab vs periodic engine", "n_gpus": world, "particles": N, "steps": a.steps + 40, "owned": own,
                  "bit_identical": same, "resetCounter": [cb["resetCounter"], cr["resetCounter"]], "step": [cb["step"], cr["step"]],
                  "pairs_equal": bool(np.array_equal(pairs, ref.pair_set())), "n_pairs": int(len(pairs)),
                  "order": [order, float(ref.order_orientation()[0][0])], "msd": [float(msd), float(ref.msd()[0])]}
        rc = ref.spatial_correlations(20.0)
        report["checksum"] = ["%016x" % chk, "%016x" % ref.checksum(), "%016x" % chk2]
        report["corr_counts_equal"] = bool(np.array_equal(corr["counts"], rc["counts"]))
        report["corr_sums_rel_err"] = float(max(np.max(np.abs(corr[k] - rc[k]) / np.maximum(np.abs(rc[k]), 1.0)) for k in ("ori_sum", "vel_sum", "pair_sum")))
        report["checkpoint_step"] = [int(step2), int(cb["step"])]
        ok = all(same.values()) and cb["resetCounter"] == cr["resetCounter"] and report["pairs_equal"] and sum(own) == N \
            and abs(report["order"][0] - report["order"][1]) <= 1e-12 and chk == ref.checksum() == chk2 and report["corr_counts_equal"] \
            and report["corr_sums_rel_err"] <= 1e-10 and step2 == cb["step"]
        report["ok"] = bool(ok)
        print(json.dumps(report))
        ref.close()
    box.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
