#!/usr/bin/env python
"""Benchmark of the B200 hot path: fp64 particle-steps/s of Engine::calculate_next_positions
(reference code/jam/jamming.cpp:837-853) through the C ABI of include/apj_b200.h.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code on the host cores

One bench step = ONE Euler step of every particle of the workload (one call of
calculate_next_positions); `value` = particles x K / time of the K timed steps, state resident in
HBM (CUDA events on the engine's stream, max over ranks); `e2e` = the same metric for a whole job
through the host-facing API with HOST buffers: upload of the full state, the K steps with the
reference driver's observable read-backs, download of the full state -- copies inside the timed
region. Workloads (BASELINE.json configs): see WORKLOADS. One JSON line on stdout (rank 0).

`state_checksum` (apj_state_checksum: a 64-bit fingerprint of {id, x, y, cos, sin} summed over particles,
taken right after the K timed steps, plus resetCounter) does not depend on the decomposition: the lines of
`--gpus 1/2/4/8` on `box16m` must carry the SAME value -- the bit-identity of the slab path, proved on the
hardware the number was measured on.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PI = 3.14159265  # reference jamming.cpp:3

WORKLOADS = {
    # name: (particles, packing fraction, lambda_s, lambda_n, replicas, BASELINE.json config it is)
    "box16m": (16777216, 0.9, 0.05, 0.5, 1, "configs[3]: N=16M periodic box (the configuration the 1e10 target is quoted on); slabs along x when n_gpus > 1"),
    "jam65k": (65536, 1.0, 0.05, 0.5, 1, "configs[1]: N=65,536 above jamming, single B200"),
    "obs1m": (1048576, 0.9, 0.05, 0.5, 1, "configs[2]: N=1M single B200"),
    "sweep512": (4096, 0.9, 0.05, 0.5, 64, "configs[4]: 64 replicas x N=4096 per GPU"),
    "jam1k": (1024, 0.9, 0.05, 0.5, 1, "configs[0]: N=1024, the reference's own CPU-runnable case"),
    "box16m_hot": (16777216, 0.9, 0.5, 0.5, 1, "configs[3] geometry at lambda_s = 0.5: the rebuild chain fires every ~10 steps"),
}
# observable cadence of the reference driver (jamming.cpp:56-62, :155-162 local settings, :211-255)
FLUCT_INT, NSKIP, TIME_AVG, T_CORR, CUTOFF = 10, 100, 10, 10, 20.0
CPU_SAMPLE_N = 16384      # particles of one reference Engine in the CPU legs
RELAX = (2000, 2000)      # trelax, tthermalize of the reference's local relax() (jamming.cpp:489-499)


def synthetic_state(n, rho, seed):
    """radii 1 + N(0,1)/10 (jamming.cpp:296), L = sqrt(PI sum R^2 / rho) (:305) summed
    in index order, uniform random positions and angles."""
    rng = np.random.default_rng(seed)
    R = 1.0 + 0.1 * rng.standard_normal(n)
    L = float(np.sqrt(PI * np.cumsum(R * R)[-1] / rho))    # cumsum = index-order accumulation, like the reference loop
    x = rng.uniform(-L / 2, L / 2, n)
    y = rng.uniform(-L / 2, L / 2, n)
    phi = rng.uniform(-PI, PI, n)
    return R, L, x, y, phi


# ----------------------------------------------------------------------------------------------
# clocks
class ClockSampler:
    """In-process NVML sampler (a thread polling every ~2 ms): SM clock, power and the clock-event reasons
    over the warm-up + timed region. nvidia-smi -lms cannot deliver inside a 20 ms timed region."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index):
        self.rows, self.h, self.stop_flag, self.mx = [], None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    index = int(vis.split(",")[index])
                except Exception:
                    pass
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.rows.append((time.time(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                  int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)), nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:
                pass
            time.sleep(0.002)

    def window(self, t0, t1):
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        sm = [r[1] for r in rows]
        reasons = sorted({name for r in rows for name, bit in self.REASONS if r[2] & bit})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx, "reasons": reasons, "samples": len(sm),
                "power_w_max": max((r[3] for r in rows), default=None), "how": "NVML polled in-process every ~2 ms over warm-up + timed region"}

    def stop(self):
        self.stop_flag = True


# ----------------------------------------------------------------------------------------------
# CPU legs: the reference's own code (oracle/_ref) on the host cores, one serial Engine per core --
# the reference's parallel model is a job array of single-core runs (code/jam/jamming.sh:2-4).
def cpu_worker(n, rho, l_s, l_n, seed, warm, steps):
    from oracle import pyoracle
    if pyoracle.have_ref():
        E = pyoracle.RefEngine
        E.seed(seed)
        r = E(n, 1000, l_s, l_n, rho)
        r.init_cells(); r.topology(); r.assign(); r.build()      # reference lattice init (jamming.cpp:285-354)
        r.set_params(0.0, l_n); r.steps(warm // 2)
        r.set_params(l_s, l_n); r.steps(warm - warm // 2)
        r.mark_origin()
        secs = r.time_steps(steps)
        kind = "reference"
    else:                                                            # reference build absent: the oracle port
        R, L, x, y, phi = synthetic_state(n, rho, seed)
        o = pyoracle.OracleSim.from_arrays(R, x, y, phi, rho)
        o.topology(); o.assign(); o.build(); o.mark_origin()
        o.set_params(0.0, l_n); o.run_philox(seed, 0, warm // 2)
        o.set_params(l_s, l_n); o.run_philox(seed, warm // 2, warm - warm // 2)
        t0 = time.perf_counter(); o.run_philox(seed, warm, steps); secs = time.perf_counter() - t0
        kind = "port"
    print(json.dumps({"secs": secs, "kind": kind}))


def run_cpu(rho, l_s, l_n, warm, steps, n=CPU_SAMPLE_N):
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--_cpu_worker", json.dumps([n, rho, l_s, l_n, 1000 + k, warm, steps])],
                              stdout=subprocess.PIPE, text=True, cwd=ROOT) for k in range(cores)]
    outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
    rate = sum(n * steps / o["secs"] for o in outs)
    slowest = max(o["secs"] for o in outs)
    return {"value": rate, "unit": "particle-steps/s", "cores": cores, "kind": outs[0]["kind"],
            "sample": "%d independent serial Engines (one per host core, the reference's job-array model), N=%d each at phi=%g "
                      "lambda_s=%g lambda_n=%g, reference lattice init + %d warm-up steps, %d timed calculate_next_positions() each"
                      % (cores, n, rho, l_s, l_n, warm, steps),
            "per_core": rate / cores, "secs": slowest, "engines": cores, "particles_timed": cores * n, "steps_timed": steps}


CPU_MIN_STEPS = 1500      # the CPU legs time at least this many steps per Engine (rebuilds included), whatever --steps says


def run_cadence(e, K, reps, queued=True):
    """The reference driver's measurement loop (jamming.cpp:207-255) for steps t = 0..K-1 through the C ABI:
    measureFluctuations every FLUCT_INT steps, again + order / orientation / COM / MSD every NSKIP, and TIME_AVG
    times per run assignCellsToGrid + buildVerletLists + spatialCorrelations + velDist + density_distribution,
    followed by T_CORR steps of the orientation autocorrelation. The reductions are QUEUED on the stream
    (apj_obs_enqueue) and read back in one copy per NSKIP block (apj_obs_fetch); the histograms and the pair
    correlations (10 x per run) are synchronous calls. Returns the number of observable calls."""
    calls = 0
    every = max(1, K // TIME_AVG)
    events = sorted(set(list(range(0, K, FLUCT_INT)) + list(range(0, K, NSKIP)) + list(range(every, K, every))
                        + [t + k for t in range(every, K, every) for k in range(T_CORR) if t + k < K]))
    radius = np.full(reps, 3.0)
    queued = queued and hasattr(e, "obs_enqueue")
    first, pending = None, 0
    OBS_COM, OBS_ORDER, OBS_MSD, OBS_FLUCT = 0, 1, 2, 3        # include/apj_b200.h APJ_OBS_*

    def q(kind, param=None):
        nonlocal first, pending, calls
        calls += 1
        if not queued:
            {OBS_FLUCT: lambda: e.fluct_area(radius), OBS_ORDER: e.order_orientation, OBS_MSD: e.msd,
             OBS_COM: lambda: (e.get_com(0) if reps > 1 or not hasattr(e, "local") else e.get_com())}[kind]()
            return
        tk = e.obs_enqueue(kind, param)
        first = tk if first is None else first
        pending += 1

    def flush():
        nonlocal first, pending
        if pending:
            e.obs_fetch(first, pending)
        first, pending = None, 0

    done = 0                                   # steps executed; event at t is evaluated after step t+1 (calculate_next_positions first)
    for t in events:
        if t + 1 > done:
            e.step(t + 1 - done); done = t + 1
        if t % FLUCT_INT == 0:
            q(OBS_FLUCT, radius)
        if t % NSKIP == 0:
            q(OBS_FLUCT, radius); q(OBS_ORDER); q(OBS_MSD); q(OBS_COM)
            flush()
        if t % every == 0 and t != 0:
            e.force_rebuild(); q(OBS_ORDER); e.spatial_correlations(CUTOFF); e.vel_hist(np.full(reps, 0.001)); e.occupancy_hist()
            calls += 3
        if t >= every and (t % every) < T_CORR:
            q(OBS_ORDER)
    if done < K:
        e.step(K - done)
    flush()
    return calls


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="box16m", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="override the workload's particle count")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-relax", action="store_true", help="shorten the relax() schedule (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end job (profiling runs)")
    ap.add_argument("--cadence", action="store_true", help="timed region runs the reference driver's observable cadence (default for obs1m)")
    ap.add_argument("--_cpu_worker", default=None)
    a = ap.parse_args()
    if a._cpu_worker:
        return cpu_worker(*json.loads(a._cpu_worker))

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_tot, rho, l_s, l_n, reps, what = WORKLOADS[a.workload]
    if a.particles:
        n_tot = a.particles
    K, W = max(1, a.steps), max(3, a.warmup)
    cadence = a.cadence or a.workload == "obs1m"

    if a.impl == "reference":
        if rank != 0:
            return
        steps_cpu = max(K, CPU_MIN_STEPS)
        c = run_cpu(rho, l_s, l_n, max(W, 200), steps_cpu, n=min(CPU_SAMPLE_N, n_tot))
        line = {"impl": "reference", "metric": "particle-steps/sec (fp64)", "value": c["value"], "unit": "particle-steps/s",
                "n_gpus": a.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * c["secs"] / steps_cpu, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": a.workload, "what": what, "phi": rho, "lambda_s": l_s, "lambda_n": l_n,
                           "particles": c["particles_timed"], "particles_timed": c["particles_timed"], "engines": c["engines"],
                           "particles_per_engine": min(CPU_SAMPLE_N, n_tot), "steps_timed": c["steps_timed"], "workload_particles": n_tot,
                           "note": "BOUNDED SAMPLE, not the workload's configuration: the reference is a serial program whose parallel model is a job "
                                   "array, so the arm runs one Engine of N=%d per host core for max(--steps, %d) steps (rebuilds inside) and reports the "
                                   "summed particle-steps/s. Its cost per particle-step grows with N (cache misses, O(N nbox) binning), so this rate is "
                                   "an upper bound for the reference on the %d-particle workload." % (CPU_SAMPLE_N, CPU_MIN_STEPS, n_tot)},
                "cpu_baseline": {"value": c["value"], "unit": c["unit"], "cores": c["cores"], "kind": c["kind"], "sample": c["sample"]},
                "e2e": {"value": c["value"], "unit": c["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from active_particle_jamming_b200 import DeviceEngine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == a.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def pinned_empty(n, dtype=np.float64):
        return torch.empty(int(n), dtype=torch.float64 if dtype == np.float64 else torch.int32, pin_memory=True).numpy()

    def pinned(arr):
        t = pinned_empty(arr.size, arr.dtype)
        t[...] = arr
        return t

    # ---- this rank's share.
    #  * one box (reps == 1), n_gpus > 1: the SAME periodic box of n_tot particles, cut along x into slabs of
    #    whole cell columns, one slab per GPU; halo / partial / migration exchange on the device over NVLink
    #    peer memory (include/apj_b200.h "slab mode"). Total work is fixed: strong scaling.
    #  * replica workloads (reps > 1): independent ensembles per GPU, no communication: weak scaling.
    slab = world > 1 and reps == 1
    if slab:
        from active_particle_jamming_b200.slab import DistSlab
        R, L, x, y, phi = synthetic_state(n_tot, rho, 12345)               # every rank draws the same box, keeps its slab
        e = DistSlab(n_tot, L, device=local, seed=12345)
        me = e.local[0]
        e.upload(x=x, y=y, R=R, phi=phi)
        del R, x, y, phi
        n_loc = me.info()["n_own"]
        parallelism = "slab%d (periodic slabs of whole cell columns along x; device-side peer-memory halo push fused into the " \
                      "step kernel, all-rank partial exchange + commit, peer-atomic migration at rebuild)" % world
        scaling = "strong"
    else:
        n_loc = n_tot
        parallelism = "single" if world == 1 else "replicas: %d independent ensembles per GPU, no communication" % reps
        scaling = "strong" if world == 1 else "weak"
        seed0 = 12345 if reps == 1 else 12345 + rank * reps                # one box: the same box whatever the GPU count
        R, L, x, y, phi = synthetic_state(n_loc, rho, seed0)
        Ls = [L] * reps
        if reps > 1:
            parts = [synthetic_state(n_loc, rho, seed0 + s) for s in range(reps)]
            Ls = [p[1] for p in parts]
            R, x, y, phi = (np.concatenate([p[k] for p in parts]) for k in (0, 2, 3, 4))
        host = {k: pinned(v) for k, v in dict(x=x, y=y, R=R, phi=phi).items()}
        e = DeviceEngine(n_loc, Ls, n_systems=reps, device=local, seed=12345 if reps == 1 else 12345 + rank)
        me = e
        e.upload(**host)
    e.skip_self_term_once()
    trelax, ttherm = (50, 50) if a.no_relax else RELAX                     # relax(): jamming.cpp:482-525
    e.set_activity(0.0, l_n); e.step(trelax)
    e.set_activity(l_s, l_n); e.set_ramp(ttherm); e.step(ttherm); e.set_ramp(0)
    e.mark_origin(); e.set_reset_counter(0)

    # ---- device-resident throughput
    smi = ClockSampler(local)
    t0w = time.time()
    e.step(W)
    c0 = me.counters()
    sw0 = me.sweep_stats()
    barrier()
    me.timer_begin()
    obs_calls = run_cadence(e, K, reps) if cadence else (e.step(K) or 0)
    ms = me.timer_end()
    barrier()
    t1w = time.time()
    checksum = e.checksum()                                                # collective in slab mode
    c1 = me.counters()
    sw1 = me.sweep_stats()
    ms = max_over_ranks(ms)
    clocks = smi.window(t0w, t1w)
    smi.stop()
    n_loc = me.info()["n_own"] if slab else n_loc
    particles = float(n_tot) if slab else sum_over_ranks(float(n_loc * reps))
    value = particles * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel (fused step kernel): event pair around every launch
    n_full, list_max = e.list_stats() if slab else e.list_stats(0)
    barrier()
    kms, committed = me.time_step_kernel(min(max(K, 16), 512))
    barrier()
    parts = me.time_step_parts(min(max(K, 16), 128))                       # {step kernel, fold + commit, slab commit (rendezvous)}
    per_rank = None
    if world > 1:                                                          # per-rank kernel time and share (load balance of the slabs)
        t = torch.tensor([kms, float(n_loc)] + parts, dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = {"kernel_ms": [float(v[0]) for v in allt], "particles": [int(v[1]) for v in allt],
                    "step_kernel_ms": [float(v[2]) for v in allt], "fold_commit_ms": [float(v[3]) for v in allt],
                    "slab_commit_ms": [float(v[4]) for v in allt]}
        kms = max(per_rank["kernel_ms"])
    b_alg = 128.0 + 4.0 * n_full                                           # bytes per particle-step
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = n_loc * reps * b_alg / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "frac_of_nominal_8000": achieved / 8000.0,
                "kernel": "apj_step_kernel (+ apj_reduce_commit_kernel)", "kernel_ms": kms,
                "parts_ms": {"step_kernel": parts[0], "fold_commit": parts[1], "slab_commit": parts[2],
                             "how": "events between the kernels of single steps (apj_time_step_parts)"}, "bytes_per_particle_step": b_alg, "n_full": n_full,
                "algorithmic": "achieved = particles x (128 + 4 n_full) B / kernel time: SURVEY 8(d)'s per-particle-step figure with the FULL list "
                               "length, whatever the kernel really moves (16-bit list entries, skin-truncated sweeps); `traffic` is the measured DRAM figure",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}
    if per_rank:
        roofline["per_rank"] = per_rank
        roofline["note"] = "per GPU: this rank's particles x bytes / slowest rank's kernel time"
    try:
        # ncu capture of the single-GPU workload (profiles/traffic.json); only that configuration has a measured figure
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(a.workload)
        roofline["traffic"] = tr if (world == 1 and not a.particles) else None
        if roofline["traffic"]:
            roofline["frac_from_traffic"] = tr / (kms * 1e-3) / 1e9 / peak
    except Exception:
        pass

    # ---- the rebuild chain (assignCellsToGrid + buildVerletLists as a counting sort + tile-staged list build)
    nreb = c1["rebuilds"] - c0["rebuilds"]
    barrier()
    me.timer_begin(); e.force_rebuild(); reb_ms = max_over_ranks(me.timer_end())
    rebuild = {"ms_per_rebuild": reb_ms, "rebuilds_in_timed_region": nreb, "rebuilds_per_1000_steps": 1000.0 * nreb / K,
               "amortised_share_of_step": (nreb * reb_ms / K) / (ms / K) if ms > 0 else None,
               "how": "CUDA events around one apj_force_rebuild (chain of %d launches + control read-back)" % (15 if slab else 11)}

    # ---- end to end: a whole job through the host-facing API, host buffers in, host buffers out
    e2e = None
    if not a.no_e2e:
        F14 = ("x", "y", "x_real", "y_real", "x0", "y0", "x_old", "y_old", "R", "phi", "cosp", "sinp", "vx", "vy")
        cap = me.info()["capacity"] if slab else n_loc * reps
        bufs = {k: pinned_empty(cap) for k in F14}                         # page-locked host buffers, reused in and out
        if slab:
            ids_pin = pinned_empty(cap, np.int32)
            ids_host, st_host = me.download_local(F14, out=bufs, ids_out=ids_pin)
            com = [me.get_com(0)]
            nloc0 = ids_host.size
            up = {k: st_host[k] for k in F14}
        else:
            bufs["box"] = pinned_empty(cap, np.int32)
            st_host = e.download(out=bufs)
            com = [e.get_com(s) for s in range(reps)]
            up = {k: st_host[k] for k in F14}
        barrier()
        tj0 = time.perf_counter()
        if slab:
            e._each(lambda r: r.upload_local(ids_host, **up))              # H2D: 14 fp64 + 1 int32 (id) per owned particle
            e.set_com(com=com[0]["COM"], com0=com[0]["COM0"], com_old=com[0]["COM_old"])
        else:
            e.upload(box=st_host["box"], **up)                             # H2D: 14 fp64 + 1 int32 per particle
            for s in range(reps):
                e.set_com(s, com=com[s]["COM"], com0=com[s]["COM0"], com_old=com[s]["COM_old"])
        done = 0
        while done < K:                                                    # driver cadence: order/orientation/COM/MSD every 100 steps (:218-239)
            n = min(100, K - done)
            e.step(n); done += n
            e.order_orientation(); e.msd(); me.get_com(0)
        out = me.download_local(F14, out=bufs, ids_out=ids_pin)[1] if slab else e.download(out=bufs)   # D2H: full state
        barrier()
        tj1 = time.perf_counter()
        tj = max_over_ranks(tj1 - tj0)
        nb = particles if slab else sum_over_ranks(float(n_loc * reps))
        e2e = {"value": particles * K / tj, "unit": "particle-steps/s", "h2d_bytes_per_step": (14 * 8 + 4) * nb / K,
               "d2h_bytes_per_step": ((14 * 8 + 4) * nb + 40 * ((K + 99) // 100) * reps * world) / K, "job_seconds": tj,
               "what": ("apj_slab_upload" if slab else "apj_upload_state") + " (page-locked host SoA) + K x apj_step with order/orientation/MSD/COM "
                       "read back every 100 steps + " + ("apj_slab_download" if slab else "apj_download_state") + " into page-locked host SoA; "
                       "bytes are totals over all GPUs"}
        assert np.all(np.isfinite(out["x"])) and np.all(np.isfinite(out["cosp"]))

    line = {"metric": "particle-steps/sec (fp64)", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": a.workload, "what": what, "particles": int(particles), "particles_per_gpu": int(particles / world), "phi": rho,
                       "lambda_s": l_s, "lambda_n": l_n, "dt": 0.1, "parallelism": parallelism, "relax": [trelax, ttherm],
                       "l2": "state + lists per GPU = %.0f MB, larger than the 126 MB L2: no flush between steps" % (n_loc * reps * (108 + 4 * n_full) / 1e6)
                       if n_loc * reps * 150 > 200e6 else "state fits L2 (%.0f MB): L2-resident by nature of the workload, no flush" % (n_loc * reps * 150 / 1e6),
                       "rebuilds_in_timed_region": nreb, "list_max": list_max, "tuning": me.tuning(),
                       "observables": ("reference driver cadence inside the timed region (jamming.cpp:211-255): fluct every %d steps, order/orientation/COM/MSD "
                                       "every %d, rebuild + spatialCorrelations(cutoff %g) + velDist + density_distribution %d x per run, %d autocorrelation steps "
                                       "after each; %d observable calls" % (FLUCT_INT, NSKIP, CUTOFF, TIME_AVG, T_CORR, obs_calls)) if cadence else "none in the timed region",
                       "sweep": dict(sw1, steps_per_class=[b - a for a, b in zip(sw0["steps_per_class"], sw1["steps_per_class"])],
                                     retried=sw1["retried"] - sw0["retried"])},
            "state_checksum": {"value": "%016x" % checksum, "resetCounter": c1["resetCounter"], "step": c1["step"],
                               "what": "apj_state_checksum after the timed steps: sum over particles of hash(id, bits of x, y, cos, sin) mod 2^64; "
                                       "identical for every --gpus N of the same workload (decomposition-independent, bit-exact state)"},
            "roofline": roofline, "rebuild": rebuild, "gpu_launches": c1["launches"] - c0["launches"], "clocks": clocks}
    if e2e:
        line["e2e"] = e2e
    if rank == 0 and world == 1 and not a.no_cpu:
        c = run_cpu(rho, l_s, l_n, 200, CPU_MIN_STEPS, n=min(CPU_SAMPLE_N, n_tot))
        line["cpu_baseline"] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample", "per_core")}
    e.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
