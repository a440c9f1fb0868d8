"""ctypes binding of the C ABI in include/apj_b200.h (the same stub a reference-side maintainer
would write). No torch types cross this boundary; there is no CPU fallback:
if the CUDA library is missing or no device is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

from ._build import LIB

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)

# reference code/jam/jamming.cpp:3-4 (truncated literals are part of the algorithm)
PI = 3.14159265
PI2 = 6.28318531

FLAG_NO_GRAPH = 1


class ApjError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("apj_b200 error %d: %s" % (code, msg))
        self.code = code


class _Config(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_systems", C.c_int32), ("device", C.c_int32), ("dt", C.c_double),
                ("rn", C.c_double), ("rs_factor", C.c_double), ("seed", C.c_uint64), ("max_neighbors", C.c_int32),
                ("steps_per_launch", C.c_int32), ("flags", C.c_int32), ("tile_slots", C.c_int32),
                ("lanes_per_particle", C.c_int32), ("reserved", C.c_int32)]


STATE_FIELDS = ["x", "y", "x_real", "y_real", "x0", "y0", "x_old", "y_old", "R", "phi", "cosp", "sinp", "vx", "vy"]


class _State(C.Structure):
    _fields_ = [(n, _dp) for n in STATE_FIELDS] + [("box", _ip)]


EXPORTS = ["apj_version", "apj_last_error", "apj_create", "apj_destroy", "apj_set_activity", "apj_set_ramp",
           "apj_upload_state", "apj_download_state", "apj_state_checksum", "apj_lattice_box_length", "apj_init_lattice", "apj_get_box_table", "apj_overlap_hue", "apj_set_com", "apj_get_com", "apj_mark_origin",
           "apj_skip_self_term_once", "apj_step", "apj_step_injected", "apj_force_rebuild", "apj_sync", "apj_save_checkpoint", "apj_load_checkpoint",
           "apj_get_counters", "apj_get_sweep_stats", "apj_set_sweep_truncation", "apj_get_tuning", "apj_set_reset_counter", "apj_get_geometry", "apj_get_pair_list", "apj_get_cell_lists", "apj_list_stats",
           "apj_order_orientation", "apj_msd", "apj_fluct_area", "apj_obs_enqueue", "apj_obs_fetch", "apj_spatial_correlations", "apj_vel_hist",
           "apj_occupancy_hist", "apj_timer_begin", "apj_timer_end", "apj_time_step_kernel", "apj_time_step_parts",
           # slab mode (bound in slab.py)
           "apj_slab_create", "apj_slab_info", "apj_slab_export", "apj_slab_connect", "apj_slab_set_timeout", "apj_slab_ready",
           "apj_slab_upload", "apj_slab_download", "apj_slab_get_pairs", "apj_slab_export_edge", "apj_slab_spatial_correlations"]

_lib = None


def load_library():
    """dlopen the in-tree CUDA library; fail loudly (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise ApjError(-2, "CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB)
    L = C.CDLL(LIB)
    L.apj_version.restype = C.c_char_p
    L.apj_last_error.restype = C.c_char_p
    L.apj_last_error.argtypes = [C.c_void_p]
    L.apj_create.argtypes = [C.POINTER(_Config), _dp, C.POINTER(C.c_void_p)]
    L.apj_destroy.argtypes = [C.c_void_p]
    L.apj_set_activity.argtypes = [C.c_void_p, _dp, _dp]
    L.apj_set_ramp.argtypes = [C.c_void_p, C.c_int64]
    L.apj_upload_state.argtypes = [C.c_void_p, C.POINTER(_State)]
    L.apj_download_state.argtypes = [C.c_void_p, C.POINTER(_State)]
    L.apj_state_checksum.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.apj_lattice_box_length.argtypes = [C.c_int64, C.c_int32, C.c_uint64, _dp, C.c_int32, _dp]
    L.apj_init_lattice.argtypes = [C.c_void_p, C.c_uint64]
    L.apj_get_box_table.argtypes = [C.c_void_p, C.c_int32, _dp, _ip]
    L.apj_overlap_hue.argtypes = [C.c_void_p, _ip]
    L.apj_set_com.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, _dp]
    L.apj_get_com.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, _dp]
    L.apj_mark_origin.argtypes = [C.c_void_p]
    L.apj_skip_self_term_once.argtypes = [C.c_void_p, C.c_int32]
    L.apj_step.argtypes = [C.c_void_p, C.c_int64]
    L.apj_step_injected.argtypes = [C.c_void_p, _dp]
    L.apj_force_rebuild.argtypes = [C.c_void_p]
    L.apj_sync.argtypes = [C.c_void_p]
    L.apj_save_checkpoint.argtypes = [C.c_void_p, C.c_char_p]
    L.apj_load_checkpoint.argtypes = [C.c_void_p, C.c_char_p]
    L.apj_get_counters.argtypes = [C.c_void_p, C.c_int32, _lp]
    L.apj_get_tuning.argtypes = [C.c_void_p, _ip]
    L.apj_get_sweep_stats.argtypes = [C.c_void_p, C.c_int32, _dp]
    L.apj_set_sweep_truncation.argtypes = [C.c_void_p, C.c_int32]
    L.apj_set_reset_counter.argtypes = [C.c_void_p, C.c_int32, C.c_int64]
    L.apj_get_geometry.argtypes = [C.c_void_p, C.c_int32, _dp]
    L.apj_get_pair_list.argtypes = [C.c_void_p, C.c_int32, _lp, _ip, C.c_int64, _lp]
    L.apj_get_cell_lists.argtypes = [C.c_void_p, C.c_int32, _lp, _ip]
    L.apj_list_stats.argtypes = [C.c_void_p, C.c_int32, _lp]
    L.apj_order_orientation.argtypes = [C.c_void_p, _dp, _dp]
    L.apj_msd.argtypes = [C.c_void_p, _dp]
    L.apj_fluct_area.argtypes = [C.c_void_p, _dp, _dp]
    L.apj_obs_enqueue.argtypes = [C.c_void_p, C.c_int32, _dp, _lp]
    L.apj_obs_fetch.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _dp]
    L.apj_spatial_correlations.argtypes = [C.c_void_p, C.c_double, _dp, _dp, _dp, _dp]
    L.apj_vel_hist.argtypes = [C.c_void_p, _dp, _lp]
    L.apj_occupancy_hist.argtypes = [C.c_void_p, _lp]
    L.apj_timer_begin.argtypes = [C.c_void_p]
    L.apj_timer_end.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.apj_time_step_kernel.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_float), _lp]
    L.apj_time_step_parts.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_float)]
    _lib = L
    return L


def _p(a, t=_dp):
    return None if a is None else a.ctypes.data_as(t)


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError("expected %d values, got %d" % (n, a.size))
    return a


class DeviceEngine:
    """One apj_engine handle: `n_systems` independent systems of `n` particles on one GPU."""

    def __init__(self, n, L, n_systems=1, device=0, seed=12345, max_neighbors=0, steps_per_launch=0,
                 dt=0.0, rn=0.0, rs_factor=0.0, flags=0, tile_slots=0, lanes_per_particle=0):
        self.lib = load_library()
        self.n = int(n)
        self.n_systems = int(n_systems)
        Ls = _f64(np.broadcast_to(np.asarray(L, dtype=np.float64), (self.n_systems,)).copy(), self.n_systems)
        cfg = _Config(self.n, self.n_systems, int(device), dt, rn, rs_factor, int(seed), int(max_neighbors),
                      int(steps_per_launch), int(flags), int(tile_slots), int(lanes_per_particle), 0)
        h = C.c_void_p()
        rc = self.lib.apj_create(C.byref(cfg), _p(Ls), C.byref(h))
        if rc != 0:
            raise ApjError(rc, self.lib.apj_last_error(None).decode())
        self.h = h
        self.ntot = self.n * self.n_systems

    # -- plumbing ------------------------------------------------------------------------
    def _chk(self, rc):
        if rc != 0:
            raise ApjError(rc, self.lib.apj_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.apj_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- parameters ----------------------------------------------------------------------
    def set_activity(self, CFself=None, CTnoise=None):
        a = None if CFself is None else _f64(np.broadcast_to(np.asarray(CFself, dtype=np.float64), (self.n_systems,)).copy())
        b = None if CTnoise is None else _f64(np.broadcast_to(np.asarray(CTnoise, dtype=np.float64), (self.n_systems,)).copy())
        self._chk(self.lib.apj_set_activity(self.h, _p(a), _p(b)))

    def set_ramp(self, tthermalize):
        self._chk(self.lib.apj_set_ramp(self.h, int(tthermalize)))

    def skip_self_term_once(self, on=True):
        self._chk(self.lib.apj_skip_self_term_once(self.h, int(bool(on))))

    # -- state ---------------------------------------------------------------------------
    def upload(self, **fields):
        st = _State()
        keep = []
        for k, v in fields.items():
            if v is None:
                continue
            if k == "box":
                a = np.ascontiguousarray(v, dtype=np.int32)
                st.box = _p(a, _ip)
            elif k in STATE_FIELDS:
                a = _f64(v, self.ntot)
                setattr(st, k, _p(a))
            else:
                raise KeyError(k)
            keep.append(a)
        self._chk(self.lib.apj_upload_state(self.h, C.byref(st)))

    def download(self, fields=None, out=None):
        """{field: array[n_systems * n]} in original particle order. `out` may hold preallocated (e.g.
        page-locked) arrays to receive the fields."""
        fields = list(fields) if fields is not None else STATE_FIELDS + ["box"]
        st = _State()
        out = dict(out) if out is not None else {}
        for k in fields:
            want = np.int32 if k == "box" else np.float64
            a = out.get(k)
            if a is None:
                a = out[k] = np.empty(self.ntot, dtype=want)
            if a.dtype != want or a.size != self.ntot or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("download: out[%r] must be a contiguous %s array of %d values" % (k, want.__name__, self.ntot))
            if k == "box":
                st.box = _p(a, _ip)
            else:
                setattr(st, k, _p(a))
        self._chk(self.lib.apj_download_state(self.h, C.byref(st)))
        return {k: out[k] for k in fields}

    @staticmethod
    def lattice_box_length(n, dens, seed, n_systems=1, device=0):
        """Box length of every system for the radii apj_init_lattice(seed) will draw (jamming.cpp:305)."""
        lib = load_library()
        d = _f64(np.broadcast_to(np.asarray(dens, dtype=np.float64), (n_systems,)).copy())
        out = np.zeros(n_systems)
        rc = lib.apj_lattice_box_length(int(n), int(n_systems), int(seed), _p(d), int(device), _p(out))
        if rc != 0:
            raise ApjError(rc, lib.apj_last_error(None).decode())
        return out

    def init_lattice(self, seed):
        """Engine::initCells on the device (radii, jittered lattice, polarity), then binning + lists."""
        self._chk(self.lib.apj_init_lattice(self.h, int(seed)))

    def box_table(self, system=0):
        nbox = self.geometry(system)["nbox"]
        c, nb = np.zeros((nbox, 2)), np.zeros((nbox, 9), dtype=np.int32)
        self._chk(self.lib.apj_get_box_table(self.h, int(system), _p(c), _p(nb, _ip)))
        return c, nb

    def overlap_hue(self):
        o = np.zeros(self.ntot, dtype=np.int32)
        self._chk(self.lib.apj_overlap_hue(self.h, _p(o, _ip)))
        return o

    def checksum(self):
        """64-bit fingerprint of {id, x, y, cos, sin} over the particles this handle owns (apj_state_checksum)."""
        v = C.c_uint64(0)
        self._chk(self.lib.apj_state_checksum(self.h, C.byref(v)))
        return int(v.value)

    def set_com(self, system=0, com=None, com0=None, com_old=None):
        a = [None if v is None else _f64(v, 2) for v in (com, com0, com_old)]
        self._chk(self.lib.apj_set_com(self.h, int(system), _p(a[0]), _p(a[1]), _p(a[2])))

    def get_com(self, system=0):
        c, c0, co = np.zeros(2), np.zeros(2), np.zeros(2)
        self._chk(self.lib.apj_get_com(self.h, int(system), _p(c), _p(c0), _p(co)))
        return dict(COM=c, COM0=c0, COM_old=co)

    def mark_origin(self):
        self._chk(self.lib.apj_mark_origin(self.h))

    # -- stepping ------------------------------------------------------------------------
    def step(self, n=1):
        self._chk(self.lib.apj_step(self.h, int(n)))

    def step_injected(self, noise):
        noise = _f64(noise, self.ntot)
        self._chk(self.lib.apj_step_injected(self.h, _p(noise)))

    def force_rebuild(self):
        self._chk(self.lib.apj_force_rebuild(self.h))

    def sync(self):
        self._chk(self.lib.apj_sync(self.h))

    def save_checkpoint(self, path):
        self._chk(self.lib.apj_save_checkpoint(self.h, os.fsencode(path)))

    def load_checkpoint(self, path):
        self._chk(self.lib.apj_load_checkpoint(self.h, os.fsencode(path)))

    def counters(self, system=0):
        o = np.zeros(8, dtype=np.int64)
        self._chk(self.lib.apj_get_counters(self.h, int(system), _p(o, _lp)))
        return dict(step=int(o[0]), resetCounter=int(o[1]), rebuilds=int(o[2]), list_max=int(o[3]), overflow=int(o[4]),
                    launches=int(o[5]), discarded=int(o[6]), nbox=int(o[7]))

    def sweep_stats(self, system=0):
        o = np.zeros(8, dtype=np.float64)
        self._chk(self.lib.apj_get_sweep_stats(self.h, int(system), _p(o, _dp)))
        return dict(retried=int(o[0]), active=bool(o[1]), skinD=float(o[2]), kmin=int(o[3]), steps_per_class=[int(v) for v in o[4:8]])

    def set_sweep_truncation(self, on):
        self._chk(self.lib.apj_set_sweep_truncation(self.h, 1 if on else 0))

    def tuning(self):
        o = np.zeros(8, dtype=np.int32)
        self._chk(self.lib.apj_get_tuning(self.h, _p(o, _ip)))
        return dict(zip(["lanes", "tb", "ppb", "tile_cap", "tile_max", "smem_bytes", "nblk", "steps_per_launch"], map(int, o)))

    def set_reset_counter(self, value, system=0):
        self._chk(self.lib.apj_set_reset_counter(self.h, int(system), int(value)))

    def geometry(self, system=0):
        o = np.zeros(5)
        self._chk(self.lib.apj_get_geometry(self.h, int(system), _p(o)))
        return dict(L=o[0], Lover2=o[1], lp=o[2], b=int(o[3]), nbox=int(o[4]))

    def pair_list(self, system=0):
        """(offsets[N+1], idx) half list by particle id, partners ascending."""
        off = np.zeros(self.n + 1, dtype=np.int64)
        tot = C.c_int64(0)
        self._chk(self.lib.apj_get_pair_list(self.h, int(system), _p(off, _lp), None, 0, C.byref(tot)))
        idx = np.zeros(max(tot.value, 1), dtype=np.int32)
        self._chk(self.lib.apj_get_pair_list(self.h, int(system), _p(off, _lp), _p(idx, _ip), tot.value, C.byref(tot)))
        return off, idx[:tot.value]

    def pair_set(self, system=0):
        off, idx = self.pair_list(system)
        i = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(off))
        return np.stack([i, idx.astype(np.int64)], axis=1)

    def list_stats(self, system=0):
        """(mean full-list length n_full, longest list) of the lists currently held."""
        o = np.zeros(2, dtype=np.int64)
        self._chk(self.lib.apj_list_stats(self.h, int(system), _p(o, _lp)))
        return o[0] / self.n, int(o[1])

    def cell_lists(self, system=0):
        nbox = self.geometry(system)["nbox"]
        off = np.zeros(nbox + 1, dtype=np.int64)
        idx = np.zeros(self.n, dtype=np.int32)
        self._chk(self.lib.apj_get_cell_lists(self.h, int(system), _p(off, _lp), _p(idx, _ip)))
        return off, idx

    # -- observables ---------------------------------------------------------------------
    def order_orientation(self):
        o, v = np.zeros(self.n_systems), np.zeros((self.n_systems, 2))
        self._chk(self.lib.apj_order_orientation(self.h, _p(o), _p(v)))
        return o, v

    def msd(self):
        o = np.zeros(self.n_systems)
        self._chk(self.lib.apj_msd(self.h, _p(o)))
        return o

    def fluct_area(self, radius):
        r = _f64(np.broadcast_to(np.asarray(radius, dtype=np.float64), (self.n_systems,)).copy())
        o = np.zeros(self.n_systems)
        self._chk(self.lib.apj_fluct_area(self.h, _p(r), _p(o)))
        return o

    OBS_COM, OBS_ORDER, OBS_MSD, OBS_FLUCT = 0, 1, 2, 3

    def obs_enqueue(self, kind, param=None):
        """Queue a reduction on the stream (no host synchronisation); returns its ticket."""
        p = None if param is None else _f64(np.broadcast_to(np.asarray(param, dtype=np.float64), (self.n_systems,)).copy())
        tk = C.c_int64(0)
        self._chk(self.lib.apj_obs_enqueue(self.h, int(kind), _p(p), C.byref(tk)))
        return int(tk.value)

    def obs_fetch(self, first, count):
        """Raw sums of tickets first .. first+count-1: array (count, n_systems, 2), one copy + one sync."""
        out = np.zeros((int(count), self.n_systems, 2))
        self._chk(self.lib.apj_obs_fetch(self.h, int(first), int(count), _p(out)))
        return out

    def spatial_correlations(self, cutoff):
        nc, npb = int(np.ceil(cutoff / 2.0)), int(np.ceil(cutoff / 0.1))
        cnt, ori, vel = (np.zeros((self.n_systems, nc)) for _ in range(3))
        pair = np.zeros((self.n_systems, npb))
        self._chk(self.lib.apj_spatial_correlations(self.h, float(cutoff), _p(cnt), _p(ori), _p(vel), _p(pair)))
        return dict(counts=cnt, ori_sum=ori, vel_sum=vel, pair_sum=pair)

    def vel_hist(self, dv):
        d = _f64(np.broadcast_to(np.asarray(dv, dtype=np.float64), (self.n_systems,)).copy())
        o = np.zeros((self.n_systems, 100), dtype=np.int64)
        self._chk(self.lib.apj_vel_hist(self.h, _p(d), _p(o, _lp)))
        return o

    def occupancy_hist(self):
        o = np.zeros((self.n_systems, 50), dtype=np.int64)
        self._chk(self.lib.apj_occupancy_hist(self.h, _p(o, _lp)))
        return o

    # -- timing --------------------------------------------------------------------------
    def timer_begin(self):
        self._chk(self.lib.apj_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float(0)
        self._chk(self.lib.apj_timer_end(self.h, C.byref(ms)))
        return ms.value

    def time_step_parts(self, n):
        """Mean ms of {step kernel, fold + commit kernel, slab commit} over n single steps."""
        o = (C.c_float * 3)()
        self._chk(self.lib.apj_time_step_parts(self.h, int(n), o))
        return [float(v) for v in o]

    def time_step_kernel(self, n):
        ms = C.c_float(0)
        k = C.c_int64(0)
        self._chk(self.lib.apj_time_step_kernel(self.h, int(n), C.byref(ms), C.byref(k)))
        return ms.value, k.value
