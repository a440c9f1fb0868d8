"""TEST INFRASTRUCTURE ONLY. ctypes front-ends for the two CPU checkers:

* ``OracleSim``  -> oracle/libapj_oracle.so  (plain-C restatement, apj_oracle.c)
* ``RefEngine``  -> oracle/_ref/libapj_ref.so (the reference's own jamming.cpp + classes/*.h
  compiled from /root/reference by oracle/Makefile; absent on machines without the reference
  tree unless the prebuilt .so travelled with the snapshot).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product package (active_particle_jamming_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libapj_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libapj_ref.so")

PI = 3.14159265  # reference code/jam/jamming.cpp:3
PI2 = 6.28318531  # :4

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)


def build(verbose=False):
    """Compile the oracle (always) and oracle/_ref (when /root/reference is present)."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def have_ref():
    return os.path.exists(REF_SO)


def _ptr(a, t=_dp):
    return a.ctypes.data_as(t) if a is not None else None


FIELDS = ["x", "y", "xr", "yr", "x0", "y0", "xo", "yo", "R", "Rinv", "phi", "cosp", "sinp",
          "vx", "vy", "Fx", "Fy", "xnew", "ynew"]
SCALARS = ["L", "Lover2", "lp", "b", "nbox", "COMx", "COMy", "COM0x", "COM0y", "COMoldx", "COMoldy",
           "resetCounter", "t", "CFself", "CTnoise"]


class _FluctC(C.Structure):
    _fields_ = [("current_radius", C.c_double), ("current_value", C.c_double), ("rad_interval", C.c_double),
                ("time_interval", C.c_double), ("dens", C.c_double), ("counter", C.c_int)]


class OracleSim:
    """SoA simulation state driven by the C restatement."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(ORACLE_SO):
                build()
            L = C.CDLL(ORACLE_SO)
            L.orc_new.restype = C.c_void_p
            L.orc_new.argtypes = [C.c_long, C.c_double, C.c_double]
            L.orc_field.restype = _dp
            L.orc_field.argtypes = [C.c_void_p, C.c_int]
            L.orc_box.restype = _ip
            L.orc_box.argtypes = [C.c_void_p]
            L.orc_box_length.restype = C.c_double
            L.orc_box_length.argtypes = [_dp, C.c_long, C.c_double]
            L.orc_delta_norm.restype = C.c_double
            L.orc_delta_norm.argtypes = [C.c_void_p, C.c_double]
            for n in ["orc_free", "orc_finish_init", "orc_topology", "orc_assign_bruteforce", "orc_assign",
                      "orc_build_verlet", "orc_save_old", "orc_update", "orc_calculate_com", "orc_mark_origin"]:
                getattr(L, n).argtypes = [C.c_void_p]
                getattr(L, n).restype = None
            L.orc_get_scalars.argtypes = [C.c_void_p, _dp]
            L.orc_set_com.argtypes = [C.c_void_p, _dp, _dp, _dp]
            L.orc_set_params.argtypes = [C.c_void_p, C.c_double, C.c_double]
            L.orc_set_reset_counter.argtypes = [C.c_void_p, C.c_long]
            L.orc_get_verlet.restype = C.c_long
            L.orc_get_verlet.argtypes = [C.c_void_p, _lp, _ip, C.c_long]
            L.orc_get_cell_lists.restype = C.c_long
            L.orc_get_cell_lists.argtypes = [C.c_void_p, _lp, _ip]
            L.orc_get_box_neighbors.argtypes = [C.c_void_p, _ip]
            L.orc_overlap_hue.argtypes = [C.c_void_p, _ip]
            L.orc_new_skin_list.restype = C.c_int
            L.orc_new_skin_list.argtypes = [C.c_void_p]
            L.orc_neighbor_interactions.argtypes = [C.c_void_p, _dp]
            L.orc_step.restype = C.c_int
            L.orc_step.argtypes = [C.c_void_p, _dp, C.c_int]
            L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
            L.orc_u32_to_randuni.restype = C.c_double
            L.orc_u32_to_randuni.argtypes = [C.c_uint32]
            L.orc_philox_noise.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_long, _dp]
            L.orc_run_philox.restype = C.c_long
            L.orc_run_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_long, C.c_int]
            L.orc_order.restype = C.c_double
            L.orc_order.argtypes = [C.c_void_p]
            L.orc_orientation.argtypes = [C.c_void_p, _dp]
            L.orc_msd.restype = C.c_double
            L.orc_msd.argtypes = [C.c_void_p]
            L.orc_fluct_overlap.restype = C.c_double
            L.orc_fluct_overlap.argtypes = [C.c_double] * 3
            L.orc_fluct_area.restype = C.c_double
            L.orc_fluct_area.argtypes = [C.c_void_p, C.c_double]
            L.orc_fluct_init.argtypes = [C.POINTER(_FluctC), C.c_double, C.c_int, C.c_int, C.c_double]
            L.orc_fluct_measure.restype = C.c_int
            L.orc_fluct_measure.argtypes = [C.POINTER(_FluctC), C.c_void_p, _dp]
            L.orc_density_distribution.argtypes = [C.c_void_p, _dp]
            L.orc_spatial_correlations.argtypes = [C.c_void_p, C.c_double, _dp, _dp, _dp]
            L.orc_vel_dist.argtypes = [C.c_void_p, C.c_double, _dp]
            cls._lib = L
        return cls._lib

    def __init__(self, N, L, dens):
        self.l = self.lib()
        self.N = int(N)
        self.h = C.c_void_p(self.l.orc_new(self.N, float(L), float(dens)))
        for k, name in enumerate(FIELDS):
            setattr(self, name, np.ctypeslib.as_array(self.l.orc_field(self.h, k), shape=(self.N,)))
        self.box = np.ctypeslib.as_array(self.l.orc_box(self.h), shape=(self.N,))

    def close(self):
        if self.h:
            for name in FIELDS + ["box"]:
                setattr(self, name, None)
            self.l.orc_free(self.h)
            self.h = None

    @staticmethod
    def box_length(R, dens):
        R = np.ascontiguousarray(R, dtype=np.float64)
        return OracleSim.lib().orc_box_length(_ptr(R), len(R), float(dens))

    @classmethod
    def from_arrays(cls, R, x, y, phi, dens, L=None):
        """Mirror of initCells' tail (jamming.cpp:291-353) on caller-supplied arrays."""
        R = np.ascontiguousarray(R, dtype=np.float64)
        s = cls(len(R), cls.box_length(R, dens) if L is None else L, dens)
        s.R[:] = R
        s.x[:] = x
        s.y[:] = y
        s.phi[:] = phi
        s.l.orc_finish_init(s.h)
        return s

    def scalars(self):
        o = np.zeros(15)
        self.l.orc_get_scalars(self.h, _ptr(o))
        return dict(zip(SCALARS, o))

    def set_com(self, com=None, com0=None, com_old=None):
        a = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in (com, com0, com_old)]
        self.l.orc_set_com(self.h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]))

    def set_params(self, CFself, CTnoise):
        self.l.orc_set_params(self.h, float(CFself), float(CTnoise))

    def set_reset_counter(self, v):
        self.l.orc_set_reset_counter(self.h, int(v))

    def topology(self):
        self.l.orc_topology(self.h)

    def assign(self, bruteforce=False):
        (self.l.orc_assign_bruteforce if bruteforce else self.l.orc_assign)(self.h)

    def build(self):
        self.l.orc_build_verlet(self.h)

    def verlet(self):
        off = np.zeros(self.N + 1, dtype=np.int64)
        tot = self.l.orc_get_verlet(self.h, _ptr(off, _lp), None, 0)
        idx = np.zeros(max(tot, 1), dtype=np.int32)
        self.l.orc_get_verlet(self.h, _ptr(off, _lp), _ptr(idx, _ip), tot)
        return off, idx[:tot]

    def pair_set(self):
        """Sorted (i<j) pair array of the half lists."""
        off, idx = self.verlet()
        i = np.repeat(np.arange(self.N, dtype=np.int64), np.diff(off))
        return pairs_canonical(i, idx.astype(np.int64))

    def cell_lists(self):
        nbox = int(self.scalars()["nbox"])
        off = np.zeros(nbox + 1, dtype=np.int64)
        idx = np.zeros(self.N, dtype=np.int32)
        self.l.orc_get_cell_lists(self.h, _ptr(off, _lp), _ptr(idx, _ip))
        return off, idx

    def box_neighbors(self):
        nbox = int(self.scalars()["nbox"])
        o = np.zeros((nbox, 9), dtype=np.int32)
        self.l.orc_get_box_neighbors(self.h, _ptr(o, _ip))
        return o

    def overlap_hue(self):
        """Cell::over after neighborInteractions of a filmed step (jamming.cpp:653-656), by particle index."""
        o = np.zeros(self.N, dtype=np.int32)
        self.l.orc_overlap_hue(self.h, _ptr(o, _ip))
        return o

    def save_old(self):
        self.l.orc_save_old(self.h)

    def mark_origin(self):
        self.l.orc_mark_origin(self.h)

    def calculate_com(self):
        self.l.orc_calculate_com(self.h)

    def new_skin_list(self):
        return bool(self.l.orc_new_skin_list(self.h))

    def neighbor_interactions(self, noise):
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        self.l.orc_neighbor_interactions(self.h, _ptr(noise))

    def update(self):
        self.l.orc_update(self.h)

    def step(self, noise, fast_assign=True):
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        return bool(self.l.orc_step(self.h, _ptr(noise), int(fast_assign)))

    def run_philox(self, seed, first_step, n, replica=0, fast_assign=True):
        return self.l.orc_run_philox(self.h, int(seed), int(first_step), int(replica), int(n), int(fast_assign))

    @staticmethod
    def philox_noise(seed, step, N, replica=0):
        o = np.zeros(N)
        OracleSim.lib().orc_philox_noise(int(seed), int(step), int(replica), int(N), _ptr(o))
        return o

    @staticmethod
    def philox4x32_10(ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        OracleSim.lib().orc_philox4x32_10(c, k, o)
        return list(o)

    def order(self):
        return self.l.orc_order(self.h)

    def orientation(self):
        o = np.zeros(2)
        self.l.orc_orientation(self.h, _ptr(o))
        return o

    def msd(self):
        return self.l.orc_msd(self.h)

    def fluct_area(self, radius):
        return self.l.orc_fluct_area(self.h, float(radius))

    def fluct_init(self, totalSteps, skip, dens):
        f = _FluctC()
        self.l.orc_fluct_init(C.byref(f), self.scalars()["L"], int(totalSteps), int(skip), float(dens))
        return f

    def fluct_measure(self, f):
        o = np.zeros(2)
        flushed = self.l.orc_fluct_measure(C.byref(f), self.h, _ptr(o))
        return (bool(flushed), o)

    def density_distribution(self, dist=None):
        dist = np.zeros(50) if dist is None else dist
        self.l.orc_density_distribution(self.h, _ptr(dist))
        return dist

    def spatial_correlations(self, cutoff, acc=None):
        nc, npb = int(np.ceil(cutoff / 2.0)), int(np.ceil(cutoff / 0.1))
        if acc is None:
            acc = (np.zeros(nc), np.zeros(nc), np.zeros(npb))
        self.l.orc_spatial_correlations(self.h, float(cutoff), _ptr(acc[0]), _ptr(acc[1]), _ptr(acc[2]))
        return acc

    def vel_dist(self, CFself, acc=None):
        acc = np.zeros(100) if acc is None else acc
        self.l.orc_vel_dist(self.h, float(CFself), _ptr(acc))
        return acc

    def state(self):
        d = {n: getattr(self, n).copy() for n in FIELDS}
        d["box"] = self.box.copy()
        d.update(self.scalars())
        return d


def pairs_canonical(i, j):
    """Unique, lexicographically sorted (min,max) pairs as an (M,2) int64 array."""
    a = np.minimum(i, j)
    b = np.maximum(i, j)
    p = np.unique(np.stack([a, b], axis=1), axis=0)
    return p


class _RefState(C.Structure):
    _names = ["x", "y", "xr", "yr", "x0", "y0", "xo", "yo", "R", "phi", "cosp", "sinp", "vx", "vy", "xnew", "ynew", "Fx", "Fy"]
    _fields_ = [(n, _dp) for n in _names] + [("box", _ip)]


class RefEngine:
    """The reference's own Engine (jamming.cpp:43-117), NDIM=2, behind oracle/ref_harness.cpp."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO + " (reference build unavailable here)")
            L = C.CDLL(REF_SO)
            L.apjref_new.restype = C.c_void_p
            L.apjref_new.argtypes = [C.c_long, C.c_long, C.c_double, C.c_double, C.c_double]
            for n in ["apjref_delete", "apjref_init_cells", "apjref_topology", "apjref_assign", "apjref_build",
                      "apjref_neighbor_interactions", "apjref_calculate_com", "apjref_save_old", "apjref_step",
                      "apjref_relax", "apjref_mark_origin"]:
                getattr(L, n).argtypes = [C.c_void_p]
                getattr(L, n).restype = None
            L.apjref_seed.argtypes = [C.c_uint]
            L.apjref_inject_uniform.argtypes = [_dp, C.c_long]
            L.apjref_injected_consumed.restype = C.c_long
            L.apjref_randuni.restype = C.c_double
            L.apjref_randnorm.restype = C.c_double
            L.apjref_PI.restype = C.c_double
            L.apjref_PI2.restype = C.c_double
            L.apjref_set_particles.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double]
            L.apjref_get_state.argtypes = [C.c_void_p, C.POINTER(_RefState)]
            L.apjref_set_state.argtypes = [C.c_void_p, C.POINTER(_RefState)]
            L.apjref_get_scalars.argtypes = [C.c_void_p, _dp]
            L.apjref_set_com.argtypes = [C.c_void_p, _dp, _dp, _dp]
            L.apjref_set_params.argtypes = [C.c_void_p, C.c_double, C.c_double]
            L.apjref_set_reset_counter.argtypes = [C.c_void_p, C.c_long]
            L.apjref_new_skin_list.restype = C.c_int
            L.apjref_new_skin_list.argtypes = [C.c_void_p]
            L.apjref_steps.argtypes = [C.c_void_p, C.c_long]
            L.apjref_delta_norm.restype = C.c_double
            L.apjref_delta_norm.argtypes = [C.c_void_p, C.c_double]
            L.apjref_get_verlet.restype = C.c_long
            L.apjref_get_verlet.argtypes = [C.c_void_p, _lp, _ip, C.c_long]
            L.apjref_get_cell_lists.restype = C.c_long
            L.apjref_get_cell_lists.argtypes = [C.c_void_p, _lp, _ip, C.c_long]
            L.apjref_get_box_neighbors.argtypes = [C.c_void_p, _ip]
            L.apjref_num_box_pairs.restype = C.c_long
            L.apjref_num_box_pairs.argtypes = [C.c_void_p]
            L.apjref_order.restype = C.c_double
            L.apjref_order.argtypes = [C.c_void_p]
            L.apjref_orientation.argtypes = [C.c_void_p, _dp]
            L.apjref_msd.restype = C.c_double
            L.apjref_msd.argtypes = [C.c_void_p]
            L.apjref_attach_observers.argtypes = [C.c_void_p, C.c_char_p]
            L.apjref_fluct_overlap.restype = C.c_double
            L.apjref_fluct_overlap.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
            L.apjref_fluct_measure.argtypes = [C.c_void_p, _dp]
            L.apjref_density_distribution.argtypes = [C.c_void_p, _dp]
            L.apjref_corr_dims.argtypes = [C.c_void_p, _ip]
            L.apjref_spatial_correlations.argtypes = [C.c_void_p, _dp, _dp, _dp]
            L.apjref_vel_dist.argtypes = [C.c_void_p, _dp]
            L.apjref_overlap_hue.argtypes = [C.c_void_p, _ip]
            L.apjref_autocorrelation.argtypes = [C.c_void_p, C.c_int, _dp]
            L.apjref_run_start.restype = C.c_double
            L.apjref_run_start.argtypes = [C.c_void_p, C.c_char_p]
            L.apjref_time_steps.restype = C.c_double
            L.apjref_time_steps.argtypes = [C.c_void_p, C.c_long]
            cls._lib = L
        return cls._lib

    def __init__(self, N, steps, l_s, l_n, rho):
        self.l = self.lib()
        self.N = int(N)
        self.h = C.c_void_p(self.l.apjref_new(self.N, int(steps), float(l_s), float(l_n), float(rho)))

    def close(self):
        if self.h:
            self.l.apjref_delete(self.h)
            self.h = None

    @classmethod
    def seed(cls, s):
        cls.lib().apjref_seed(int(s) & 0xFFFFFFFF)

    @classmethod
    def randuni(cls):
        return cls.lib().apjref_randuni()

    def init_cells(self):
        self.l.apjref_init_cells(self.h)

    def set_particles(self, R, x, y, phi, L=0.0):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (R, x, y, phi)]
        self.l.apjref_set_particles(self.h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), float(L))

    def topology(self):
        self.l.apjref_topology(self.h)

    def get_state(self):
        st = _RefState()
        out = {}
        for n in _RefState._names:
            out[n] = np.zeros(self.N)
            setattr(st, n, _ptr(out[n]))
        out["box"] = np.zeros(self.N, dtype=np.int32)
        st.box = _ptr(out["box"], _ip)
        self.l.apjref_get_state(self.h, C.byref(st))
        out.update(self.scalars())
        return out

    def set_state(self, **fields):
        st = _RefState()
        keep = []
        for n, v in fields.items():
            if n == "box":
                a = np.ascontiguousarray(v, dtype=np.int32)
                st.box = _ptr(a, _ip)
            else:
                a = np.ascontiguousarray(v, dtype=np.float64)
                setattr(st, n, _ptr(a))
            keep.append(a)
        self.l.apjref_set_state(self.h, C.byref(st))

    def scalars(self):
        o = np.zeros(15)
        self.l.apjref_get_scalars(self.h, _ptr(o))
        return dict(zip(SCALARS, o))

    def set_com(self, com=None, com0=None, com_old=None):
        a = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in (com, com0, com_old)]
        self.l.apjref_set_com(self.h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]))

    def set_params(self, CFself, CTnoise):
        self.l.apjref_set_params(self.h, float(CFself), float(CTnoise))

    def set_reset_counter(self, v):
        self.l.apjref_set_reset_counter(self.h, int(v))

    def assign(self):
        self.l.apjref_assign(self.h)

    def build(self):
        self.l.apjref_build(self.h)

    def new_skin_list(self):
        return bool(self.l.apjref_new_skin_list(self.h))

    def overlap_hue(self):
        """Cell::over left by the reference's own neighborInteractions on a filmed step (scratch engines only:
        forces and alignment sums accumulate as in any call)."""
        o = np.zeros(self.N, dtype=np.int32)
        self.l.apjref_overlap_hue(self.h, _ptr(o, _ip))
        return o

    def save_old(self):
        self.l.apjref_save_old(self.h)

    def mark_origin(self):
        self.l.apjref_mark_origin(self.h)

    def calculate_com(self):
        self.l.apjref_calculate_com(self.h)

    def relax(self):
        self.l.apjref_relax(self.h)

    def step(self, noise=None):
        """calculate_next_positions(); with `noise`, randuni() returns noise[i] for particle i."""
        if noise is not None:
            noise = np.ascontiguousarray(noise, dtype=np.float64)
            self.l.apjref_inject_uniform(_ptr(noise), len(noise))
        self.l.apjref_step(self.h)
        if noise is not None:
            used = self.l.apjref_injected_consumed()
            self.l.apjref_clear_injection()
            assert used == self.N, used

    def steps(self, n):
        self.l.apjref_steps(self.h, int(n))

    def time_steps(self, n):
        return self.l.apjref_time_steps(self.h, int(n))

    def verlet(self):
        off = np.zeros(self.N + 1, dtype=np.int64)
        tot = self.l.apjref_get_verlet(self.h, _ptr(off, _lp), None, 0)
        idx = np.zeros(max(tot, 1), dtype=np.int32)
        self.l.apjref_get_verlet(self.h, _ptr(off, _lp), _ptr(idx, _ip), tot)
        return off, idx[:tot]

    def pair_set(self):
        off, idx = self.verlet()
        i = np.repeat(np.arange(self.N, dtype=np.int64), np.diff(off))
        return pairs_canonical(i, idx.astype(np.int64))

    def cell_lists(self):
        nbox = int(self.scalars()["nbox"])
        off = np.zeros(nbox + 1, dtype=np.int64)
        idx = np.zeros(self.N, dtype=np.int32)
        self.l.apjref_get_cell_lists(self.h, _ptr(off, _lp), _ptr(idx, _ip), self.N)
        return off, idx

    def box_neighbors(self):
        nbox = int(self.scalars()["nbox"])
        o = np.zeros((nbox, 9), dtype=np.int32)
        self.l.apjref_get_box_neighbors(self.h, _ptr(o, _ip))
        return o

    def order(self):
        return self.l.apjref_order(self.h)

    def orientation(self):
        o = np.zeros(2)
        self.l.apjref_orientation(self.h, _ptr(o))
        return o

    def msd(self):
        return self.l.apjref_msd(self.h)

    def attach_observers(self, location):
        os.makedirs(os.path.join(location, "local_output"), exist_ok=True)
        if not location.endswith("/"):
            location += "/"
        self.l.apjref_attach_observers(self.h, location.encode())

    def fluct_overlap(self, r, R, d):
        return self.l.apjref_fluct_overlap(self.h, r, R, d)

    def fluct_measure(self):
        o = np.zeros(5)
        self.l.apjref_fluct_measure(self.h, _ptr(o))
        return dict(current_radius=o[0], current_value=o[1], counter=int(o[2]), time_interval=o[3], rad_interval=o[4])

    def density_distribution(self):
        o = np.zeros(50)
        self.l.apjref_density_distribution(self.h, _ptr(o))
        return o

    def corr_dims(self):
        o = np.zeros(4, dtype=np.int32)
        self.l.apjref_corr_dims(self.h, _ptr(o, _ip))
        return dict(nc=int(o[0]), np=int(o[1]), noBins=int(o[2]), correlation_time=int(o[3]))

    def spatial_correlations(self):
        d = self.corr_dims()
        v, o, p = np.zeros(d["nc"]), np.zeros(d["nc"]), np.zeros(d["np"])
        self.l.apjref_spatial_correlations(self.h, _ptr(v), _ptr(o), _ptr(p))
        return v, o, p

    def vel_dist(self):
        o = np.zeros(100)
        self.l.apjref_vel_dist(self.h, _ptr(o))
        return o

    def run_start(self, location):
        os.makedirs(os.path.join(location, "local_output"), exist_ok=True)
        if not location.endswith("/"):
            location += "/"
        return self.l.apjref_run_start(self.h, location.encode())
