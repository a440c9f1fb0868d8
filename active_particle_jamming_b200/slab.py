"""Slab mode: ONE periodic box cut along x into slabs of whole cell columns, one rank per GPU
(BASELINE config 4). Host side only: who owns which column, which particle goes to which
rank, how per-rank answers are merged. All exchange between ranks while stepping happens on the
device, inside the CUDA library (peer stores over NVLink; include/apj_b200.h "slab mode").

Two front ends over the same `SlabRank` handle:
  * `SlabBox`   -- all ranks in THIS process, one host thread per rank (tests: several ranks on one GPU);
  * `DistSlab`  -- one rank per process under torchrun; torch.distributed carries the plumbing only
                   (the 64-byte IPC handles, COM / observable sums, barriers).
The decomposition arithmetic (`slab_columns`, `owner_of`, `merge_by_id`) is plain numpy and is what the
CPU (gloo) tests cover.
"""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .device import (ApjError, DeviceEngine, STATE_FIELDS, _Config, _State, _dp, _f64, _ip, _lp, _p,
                     load_library)

_u64p = C.POINTER(C.c_uint64)


def _bind(L):
    if getattr(L, "_slab_bound", False):
        return L
    L.apj_slab_create.argtypes = [C.POINTER(_Config), C.c_double, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_void_p)]
    L.apj_slab_info.argtypes = [C.c_void_p, _lp]
    L.apj_slab_export.argtypes = [C.c_void_p, C.c_void_p, _u64p]
    L.apj_slab_connect.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_int32]
    L.apj_slab_ready.argtypes = [C.c_void_p]
    L.apj_slab_set_timeout.argtypes = [C.c_void_p, C.c_double]
    L.apj_slab_upload.argtypes = [C.c_void_p, C.POINTER(_State), _ip, C.c_int64]
    L.apj_slab_download.argtypes = [C.c_void_p, C.POINTER(_State), _ip, C.c_int64, _lp]
    L.apj_slab_get_pairs.argtypes = [C.c_void_p, _ip, C.c_int64, _lp]
    L.apj_slab_export_edge.argtypes = [C.c_void_p, C.c_double, C.c_int64, _dp, _lp]
    L.apj_slab_spatial_correlations.argtypes = [C.c_void_p, C.c_double, C.c_int64, _dp, _ip, _dp, _dp, _dp, _dp]
    L._slab_bound = True
    return L


# ---- decomposition arithmetic (host, numpy) -------------------------------------------------------
def grid_b(L, rn=2.8):
    """Engine::topology (reference jamming.cpp:361-365): cells per side."""
    return int(np.floor(L / (2 * rn)))


def slab_columns(b, nranks):
    """[(first column, columns)] per rank: rank r owns [r*b//nranks, (r+1)*b//nranks)."""
    edges = [r * b // nranks for r in range(nranks + 1)]
    return [(edges[r], edges[r + 1] - edges[r]) for r in range(nranks)]


def owner_of(x, L, b, nranks):
    """Rank that owns each particle, from the floor bin of x (the device re-bins with the reference's
    nearest-centre rule and migrates the rare particle that sits exactly on a column edge)."""
    col = np.clip(np.floor((np.asarray(x) + L / 2) / (L / b)).astype(np.int64), 0, b - 1)
    edges = np.array([r * b // nranks for r in range(nranks + 1)], dtype=np.int64)
    return (np.searchsorted(edges, col, side="right") - 1).astype(np.int32)


def merge_by_id(parts, n, fields):
    """parts: per rank (ids, {field: values}); returns {field: array[n]} in original particle order."""
    out = {}
    seen = np.zeros(n, dtype=np.int32)
    for ids, _ in parts:
        np.add.at(seen, ids, 1)
    if not np.all(seen == 1):
        raise ValueError("slab merge: %d particles missing, %d duplicated" % (int((seen == 0).sum()), int((seen > 1).sum())))
    for f in fields:
        a = np.zeros(n, dtype=np.int32 if f == "box" else np.float64)
        for ids, d in parts:
            a[ids] = d[f]
        out[f] = a
    return out


def merge_pairs(parts):
    """Union of the per-rank (id_i, id_j) lists, sorted like DeviceEngine.pair_set."""
    p = np.concatenate([q.reshape(-1, 2) for q in parts], axis=0).astype(np.int64)
    order = np.lexsort((p[:, 1], p[:, 0]))
    return p[order]


# ---- one rank -------------------------------------------------------------------------------------
class SlabRank(DeviceEngine):
    """One slab of the global box on one GPU (handle created by apj_slab_create)."""

    def __init__(self, n, L, rank, nranks, device=0, capacity=0, seed=12345, max_neighbors=0, steps_per_launch=0,
                 dt=0.0, rn=0.0, rs_factor=0.0, flags=0, tile_slots=0, lanes_per_particle=0):
        self.lib = _bind(load_library())
        self.n = int(n)
        self.n_systems = 1
        self.ntot = self.n
        self.rank, self.nranks, self.device = int(rank), int(nranks), int(device)
        cfg = _Config(self.n, 1, int(device), dt, rn, rs_factor, int(seed), int(max_neighbors),
                      int(steps_per_launch), int(flags), int(tile_slots), int(lanes_per_particle), 0)
        h = C.c_void_p()
        rc = self.lib.apj_slab_create(C.byref(cfg), float(L), self.rank, self.nranks, int(capacity), C.byref(h))
        if rc != 0:
            raise ApjError(rc, self.lib.apj_last_error(None).decode())
        self.h = h
        self.L = float(L)

    def info(self):
        o = np.zeros(8, dtype=np.int64)
        self._chk(self.lib.apj_slab_info(self.h, _p(o, _lp)))
        return dict(zip(["rank", "nranks", "col0", "ncols", "capacity", "ghost_capacity", "n_own", "arena_bytes"], map(int, o)))

    def export(self):
        """(64-byte IPC handle, local device pointer) of this rank's peer-visible arena."""
        hb = (C.c_ubyte * 64)()
        ptr = C.c_uint64(0)
        self._chk(self.lib.apj_slab_export(self.h, hb, C.byref(ptr)))
        return bytes(hb), int(ptr.value)

    def connect(self, peer, handle=None, ptr=0, peer_device=-1):
        hb = (C.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        self._chk(self.lib.apj_slab_connect(self.h, int(peer), hb, int(ptr), int(peer_device)))

    def ready(self):
        self._chk(self.lib.apj_slab_ready(self.h))

    def set_timeout(self, seconds):
        self._chk(self.lib.apj_slab_set_timeout(self.h, float(seconds)))

    def upload_local(self, ids, **fields):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        st = _State()
        keep = [ids]
        for k, v in fields.items():
            if v is None:
                continue
            if k not in STATE_FIELDS:
                raise KeyError(k)
            a = _f64(v, ids.size)
            setattr(st, k, _p(a))
            keep.append(a)
        self._chk(self.lib.apj_slab_upload(self.h, C.byref(st), _p(ids, _ip), ids.size))

    def download_local(self, fields=None, out=None, ids_out=None):
        """(ids, {field: values}) of the owned particles in device order. `out` / `ids_out` may hold
        preallocated (e.g. page-locked) arrays of at least n_own entries."""
        fields = list(fields) if fields is not None else STATE_FIELDS + ["box"]
        n = C.c_int64(0)
        st = _State()
        self._chk(self.lib.apj_slab_download(self.h, C.byref(st), None, 0, C.byref(n)))
        m = max(n.value, 1)
        ids = ids_out if ids_out is not None and ids_out.size >= m and ids_out.dtype == np.int32 else np.empty(m, dtype=np.int32)
        out = dict(out) if out is not None else {}
        for k in fields:
            want = np.int32 if k == "box" else np.float64
            a = out.get(k)
            if a is None or a.size < m or a.dtype != want or not a.flags["C_CONTIGUOUS"]:
                a = out[k] = np.empty(m, dtype=want)
            if k == "box":
                st.box = _p(a, _ip)
            else:
                setattr(st, k, _p(a))
        self._chk(self.lib.apj_slab_download(self.h, C.byref(st), _p(ids, _ip), ids.size, C.byref(n)))
        return ids[:n.value], {k: out[k][:n.value] for k in fields}

    def pairs(self):
        tot = C.c_int64(0)
        self._chk(self.lib.apj_slab_get_pairs(self.h, None, 0, C.byref(tot)))
        p = np.zeros((max(tot.value, 1), 2), dtype=np.int32)
        self._chk(self.lib.apj_slab_get_pairs(self.h, _p(p, _ip), tot.value, C.byref(tot)))
        return p[:tot.value]

    def export_edge(self, width):
        """(6, n) array x | y | cos | sin | vx | vy of the owned particles within `width` of this slab's left edge."""
        n = C.c_int64(0)
        self._chk(self.lib.apj_slab_export_edge(self.h, float(width), 0, None, C.byref(n)))
        out = np.zeros((6, max(n.value, 1)))
        if n.value:
            self._chk(self.lib.apj_slab_export_edge(self.h, float(width), n.value, _p(out), C.byref(n)))
        return out[:, :n.value]

    def spatial_correlations_share(self, cutoff, ext):
        """This rank's share of the raw correlation sums; `ext` = export_edge(cutoff) of the rank to the right."""
        nc, npb = int(np.ceil(cutoff / 2.0)), int(np.ceil(cutoff / 0.1))
        g = self.geometry()
        b, lp, Lh = g["b"], g["lp"], g["Lover2"]
        ext = np.asarray(ext, dtype=np.float64).reshape(6, -1)
        row = np.clip(np.floor((ext[1] + Lh) / lp).astype(np.int64), 0, b - 1)      # cell row of every edge particle
        order = np.argsort(row, kind="stable")
        ext = np.ascontiguousarray(ext[:, order])
        row_start = np.zeros(b + 1, dtype=np.int32)
        row_start[1:] = np.cumsum(np.bincount(row, minlength=b))
        cnt, ori, vel, pair = np.zeros(nc), np.zeros(nc), np.zeros(nc), np.zeros(npb)
        self._chk(self.lib.apj_slab_spatial_correlations(self.h, float(cutoff), ext.shape[1], _p(ext), _p(row_start, _ip),
                                                         _p(cnt), _p(ori), _p(vel), _p(pair)))
        return np.concatenate([cnt, ori, vel, pair])

    # periodic-box entry points that have a slab counterpart
    def upload(self, **fields):
        raise ApjError(-4, "slab rank: use upload_local (or SlabBox / DistSlab .upload)")

    def step_injected(self, noise):
        noise = _f64(noise, self.n)      # indexed by GLOBAL particle id
        self._chk(self.lib.apj_step_injected(self.h, _p(noise)))

    def list_stats(self, system=0):
        o = np.zeros(2, dtype=np.int64)
        self._chk(self.lib.apj_list_stats(self.h, 0, _p(o, _lp)))
        return int(o[0]), int(o[1])      # (sum of list lengths over owned particles, longest)


# ---- front ends -----------------------------------------------------------------------------------
class _SlabFront:
    """What both front ends share: partition on upload, merge on download, sums of observables.
    Subclasses provide `local` (the SlabRank objects of this process), `_each(fn)` (run fn on every
    local rank concurrently) and `_allsum(array)` (sum over ALL ranks of the box)."""

    def _partition_upload(self, fields):
        x = np.asarray(fields["x"], dtype=np.float64)
        own = owner_of(x, self.L, self.b, self.nranks)
        xr = np.asarray(fields.get("x_real", fields["x"]), dtype=np.float64)
        yr = np.asarray(fields.get("y_real", fields["y"]), dtype=np.float64)
        com = np.array([np.sum(xr) / self.n, np.sum(yr) / self.n])   # calculate_COM on the whole box

        def up(r):
            ids = np.nonzero(own == r.rank)[0].astype(np.int32)
            r.upload_local(ids, **{k: np.asarray(v)[ids] for k, v in fields.items() if v is not None})
            r.set_com(0, com=com, com0=com, com_old=com)
        self._each(up)

    def upload(self, **fields):
        self._partition_upload(fields)

    def set_activity(self, CFself=None, CTnoise=None):
        self._each(lambda r: r.set_activity(CFself, CTnoise))

    def set_ramp(self, t):
        self._each(lambda r: r.set_ramp(t))

    def set_com(self, com=None, com0=None, com_old=None):
        self._each(lambda r: r.set_com(0, com=com, com0=com0, com_old=com_old))

    def set_reset_counter(self, v):
        self._each(lambda r: r.set_reset_counter(v))

    def skip_self_term_once(self, on=True):
        self._each(lambda r: r.skip_self_term_once(on))

    def step(self, n=1):
        self._each(lambda r: r.step(n))

    def step_injected(self, noise):
        noise = _f64(noise, self.n)
        self._each(lambda r: r.step_injected(noise))

    def force_rebuild(self):
        self._each(lambda r: r.force_rebuild())

    def mark_origin(self):
        self._each(lambda r: r.mark_origin())
        part = sum(r.get_com()["COM"] for r in self.local)          # each rank holds its share sum/N
        com = self._allsum(part)
        self.set_com(com=com, com0=com, com_old=com)

    def counters(self):
        return self.local[0].counters()

    def save_checkpoint(self, prefix):
        """One file per rank, `prefix`.rank<r>of<n> (the 16M box is never gathered)."""
        self._each(lambda r: r.save_checkpoint("%s.rank%dof%d" % (prefix, r.rank, r.nranks)))

    def load_checkpoint(self, prefix):
        self._each(lambda r: r.load_checkpoint("%s.rank%dof%d" % (prefix, r.rank, r.nranks)))

    def get_com(self):
        return self.local[0].get_com()

    def n_own(self):
        return [r.info()["n_own"] for r in self.local]

    def checksum(self):
        """Fingerprint of the whole box: the ranks' shares add modulo 2^64 (equals DeviceEngine.checksum of
        the same state on one GPU)."""
        mine = sum(r.checksum() for r in self.local) & 0xffffffffffffffff
        parts = self._allsum(np.array([float(mine >> 43), float((mine >> 22) & 0x1fffff), float(mine & 0x3fffff)]))
        return (int(parts[0]) * (1 << 43) + int(parts[1]) * (1 << 22) + int(parts[2])) & 0xffffffffffffffff

    # observables: additive shares
    def order_orientation(self):
        part = sum(r.order_orientation()[1][0] for r in self.local)
        v = self._allsum(part)
        return np.array([np.sqrt(v[0] * v[0] + v[1] * v[1] + 0.0 * 0.0)]), v.reshape(1, 2)

    def msd(self):
        return self._allsum(np.array([sum(r.msd()[0] for r in self.local)]))

    def fluct_area(self, radius):
        return self._allsum(np.array([sum(r.fluct_area(radius)[0] for r in self.local)]))

    def vel_hist(self, dv):
        return self._allsum(sum(r.vel_hist(dv) for r in self.local).astype(np.float64)).astype(np.int64)

    def occupancy_hist(self):
        return self._allsum(sum(r.occupancy_hist() for r in self.local).astype(np.float64)).astype(np.int64)

    def spatial_correlations(self, cutoff):
        """Correlations::spatialCorrelations raw sums of the whole box (same dict as DeviceEngine.spatial_correlations):
        every rank hands the particles near its left edge to its left neighbour, counts its own pairs and the pairs
        that cross its right edge, and the shares are added."""
        nc = int(np.ceil(cutoff / 2.0))
        # per local rank: edge set of the rank to its right. Ownership only changes at a rebuild, so an owned particle may
        # sit up to the skin (1.4) beyond its slab's edge: the edge zone is that much wider than the cutoff
        edges = self._exchange_edges(cutoff + 2.0)
        shares = self._each_indexed(lambda k, r: r.spatial_correlations_share(cutoff, edges[k]))
        tot = self._allsum(np.sum(shares, axis=0))
        return dict(counts=tot[None, :nc], ori_sum=tot[None, nc:2 * nc], vel_sum=tot[None, 2 * nc:3 * nc], pair_sum=tot[None, 3 * nc:])

    def list_stats(self):
        s = self._allsum(np.array([float(sum(r.list_stats()[0] for r in self.local))]))
        return float(s[0]) / self.n, max(r.list_stats()[1] for r in self.local)

    def close(self):
        for r in self.local:
            r.close()


class SlabBox(_SlabFront):
    """All `nranks` slabs in this process, one host thread per rank (collective calls block until every
    rank's stream is done, so they must be issued concurrently). `devices[r]` is rank r's CUDA device;
    several ranks may share one device (tests)."""

    def __init__(self, n, L, nranks, devices=None, **kw):
        self.n, self.L, self.nranks = int(n), float(L), int(nranks)
        self.b = grid_b(self.L, kw.get("rn", 0.0) or 2.8)
        devices = list(devices) if devices is not None else [0] * self.nranks
        self.local = [SlabRank(n, L, r, nranks, device=devices[r], **kw) for r in range(self.nranks)]
        self.pool = ThreadPoolExecutor(max_workers=self.nranks)
        for r in self.local:
            for q in self.local:
                if q is not r:
                    r.connect(q.rank, ptr=q.export()[1], peer_device=q.device)
        for r in self.local:
            r.ready()

    def _each(self, fn):
        return list(self.pool.map(fn, self.local))

    def _allsum(self, a):
        return np.asarray(a, dtype=np.float64)

    def _each_indexed(self, fn):
        return list(self.pool.map(lambda kr: fn(*kr), enumerate(self.local)))

    def _exchange_edges(self, width):
        mine = self._each(lambda r: r.export_edge(width))
        return [mine[(k + 1) % self.nranks] for k in range(self.nranks)]

    def download(self, fields=None):
        fields = list(fields) if fields is not None else STATE_FIELDS + ["box"]
        return merge_by_id([r.download_local(fields) for r in self.local], self.n, fields)

    def pair_set(self):
        return merge_pairs([r.pairs() for r in self.local])

    def close(self):
        super().close()
        self.pool.shutdown()


class DistSlab(_SlabFront):
    """One slab per process (torchrun: one process per GPU). torch.distributed is plumbing only."""

    def __init__(self, n, L, device=None, **kw):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.n, self.L = int(n), float(L)
        self.rank, self.nranks = dist.get_rank(), dist.get_world_size()
        self.b = grid_b(self.L, kw.get("rn", 0.0) or 2.8)
        self.device = torch.cuda.current_device() if device is None else int(device)
        me = SlabRank(n, L, self.rank, self.nranks, device=self.device, **kw)
        self.local = [me]
        self.cuda = dist.get_backend() == "nccl"
        handle, _ = me.export()
        mine = torch.tensor(list(handle), dtype=torch.uint8)
        if self.cuda:
            mine = mine.cuda(self.device)
        allh = [torch.zeros_like(mine) for _ in range(self.nranks)]
        dist.all_gather(allh, mine)
        for r in range(self.nranks):
            if r != self.rank:
                me.connect(r, handle=bytes(allh[r].cpu().numpy().tobytes()))
        me.ready()
        dist.barrier()

    def _each(self, fn):
        self.dist.barrier()          # ranks enter a collective call together (bounded device-side waits)
        return [fn(self.local[0])]

    def _each_indexed(self, fn):
        self.dist.barrier()
        return [fn(0, self.local[0])]

    def _exchange_edges(self, width):
        """My left-edge particles go to the rank on my left; I receive those of the rank on my right."""
        torch, dist = self.torch, self.dist
        mine = self.local[0].export_edge(width)
        if self.nranks == 1:
            return [mine]
        left, right = (self.rank - 1) % self.nranks, (self.rank + 1) % self.nranks
        dev = torch.device("cuda", self.device) if self.cuda else torch.device("cpu")
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.nranks)]
        dist.all_gather(counts, torch.tensor([mine.shape[1]], dtype=torch.int64, device=dev))
        send = torch.as_tensor(np.ascontiguousarray(mine)).to(dev)
        recv = torch.zeros((6, int(counts[right].item())), dtype=torch.float64, device=dev)
        ops = [dist.P2POp(dist.isend, send, left), dist.P2POp(dist.irecv, recv, right)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return [recv.cpu().numpy()]

    def _allsum(self, a):
        t = self.torch.as_tensor(np.asarray(a, dtype=np.float64))
        if self.cuda:
            t = t.cuda(self.device)
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def download(self, fields=None):
        """Every rank receives the whole box (tests / small systems)."""
        fields = list(fields) if fields is not None else STATE_FIELDS + ["box"]
        mine = self.local[0].download_local(fields)
        parts = [None] * self.nranks
        self.dist.all_gather_object(parts, mine)
        return merge_by_id(parts, self.n, fields)

    def pair_set(self):
        parts = [None] * self.nranks
        self.dist.all_gather_object(parts, self.local[0].pairs())
        return merge_pairs(parts)
