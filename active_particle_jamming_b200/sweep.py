"""Phase-diagram sweep over the GPUs of a node (BASELINE config 5): the reference's PBS job array
(code/jam/create-arrays.sh writes one "<fullRunID> <runID> <N> <steps> <lambda_s> <lambda_n> <rho>" line per
point into input.txt; code/jam/jamming.sh starts one single-core process per line) becomes one
`jam --sweep` process per GPU, each running its share of the lines as batched replicas of one device
handle (host/jam/jamming.cpp, host/classes/Batch.h). Replicas are independent: no communication.

    python -m active_particle_jamming_b200.sweep input.txt --gpus 8 [--batch 64] [--output-root DIR] [--seed S]

Lines are dealt out so that every GPU receives runs of equal shape (N, steps) in contiguous blocks --
`jam --sweep` batches consecutive lines of one shape -- and the shares are balanced by N x steps.
Exit status: 0 if every process succeeded, else the first non-zero status."""
import argparse
import os
import subprocess
import sys
import tempfile

from ._build import HOST_BIN


def parse_input(path):
    """[(line text, N, steps)] of the non-blank lines; raises ValueError on a malformed line (7 fields)."""
    runs = []
    with open(path) as f:
        for k, ln in enumerate(f, 1):
            if not ln.strip():
                continue
            p = ln.split()
            try:
                if len(p) != 7:
                    raise ValueError
                n, steps = int(p[2]), int(p[3])
                float(p[4]); float(p[5]); float(p[6])
            except ValueError:
                raise ValueError("%s:%d: expected <fullRunID> <runID> <N> <steps> <lambda_s> <lambda_n> <rho>" % (path, k))
            runs.append((" ".join(p), n, steps))
    return runs


def partition(runs, n_gpus, batch):
    """Deal the runs out to `n_gpus` shares in blocks of up to `batch` same-shape runs, always giving the next
    block to the least loaded share (load = sum of N x steps). Returns a list of lists of line texts."""
    shapes = {}
    for text, n, steps in runs:
        shapes.setdefault((n, steps), []).append(text)
    blocks = []
    for (n, steps), lines in shapes.items():
        # blocks no larger than needed to occupy all GPUs, and never larger than one device batch
        per = max(1, min(batch, -(-len(lines) // n_gpus)))
        for k in range(0, len(lines), per):
            blk = lines[k:k + per]
            blocks.append((n * steps * len(blk), blk))
    blocks.sort(key=lambda b: -b[0])
    shares, load = [[] for _ in range(n_gpus)], [0] * n_gpus
    for cost, blk in blocks:
        g = load.index(min(load))
        shares[g].extend(blk)
        load[g] += cost
    return shares


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("input")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64, help="replicas per device handle")
    ap.add_argument("--output-root", default=None, help="APJ_OUTPUT_ROOT of the runs (must contain local_output/)")
    ap.add_argument("--seed", type=int, default=None, help="APJ_SEED (default: wall clock, as the reference)")
    ap.add_argument("--jam", default=os.environ.get("APJ_JAM", HOST_BIN), help="the jam driver binary")
    a = ap.parse_args(argv)
    try:
        runs = parse_input(a.input)
    except (OSError, ValueError) as e:
        print(e)
        return 2
    if a.gpus < 1 or a.batch < 1:
        print("--gpus and --batch must be >= 1")
        return 2
    shares = partition(runs, a.gpus, a.batch)
    procs = []
    with tempfile.TemporaryDirectory(prefix="apj_sweep_") as tmp:
        for g, lines in enumerate(shares):
            if not lines:
                continue
            part = os.path.join(tmp, "input_gpu%d.txt" % g)
            with open(part, "w") as f:
                f.write("\n".join(lines) + "\n")
            env = dict(os.environ, APJ_DEVICE=str(g))
            if a.output_root:
                env["APJ_OUTPUT_ROOT"] = a.output_root
            if a.seed is not None:
                env["APJ_SEED"] = str(a.seed + g)
            procs.append((g, len(lines), subprocess.Popen([a.jam, "--sweep", part, str(a.batch)], env=env)))
        rc = 0
        for g, n, p in procs:
            r = p.wait()
            print("gpu %d: %d runs, exit status %d" % (g, n, r))
            rc = rc or r
    return rc


if __name__ == "__main__":
    sys.exit(main())
