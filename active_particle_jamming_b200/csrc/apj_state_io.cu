// Host <-> device hand-over of the particle state (apj_upload_state / apj_download_state and their
// slab and checkpoint callers). The caller's arrays are the per-particle fields of `struct Cell`
// (reference code/classes/Cell.h:15-43) as host SoA in ORIGINAL particle order; the device keeps
// 16-byte records in cell order. Everything between the two layouts runs on the GPU:
//
//   upload:   caller SoA --cudaMemcpyAsync, one per field--> staging planes in HBM
//             apj_pack_kernel: interleave {x,y} ..., Rinv = 1/R (jamming.cpp:298), cos/sin of an uploaded
//             phi (:332-333) or phi of uploaded cos/sin, box renumbering, id, radius validation
//   download: apj_unpack_kernel: scatter by particle id into the staging planes (periodic box) or
//             keep device order (slab), box renumbering; one cudaMemcpyAsync per requested field
//
// No per-call host allocation, no host loop over particles (the only host arithmetic left in the
// upload is the index-order COM sum of calculate_COM, jamming.cpp:761-774, which runs while the
// copies are in flight).
#include "apj_device.cuh"

namespace {

constexpr int IO_TB = 256;

__global__ void __launch_bounds__(IO_TB) apj_pack_kernel(const DevState st, const ApjStage sg, const unsigned present,
                                                         const long long n, const int slab) {
    const long long g = (long long)blockIdx.x * IO_TB + threadIdx.x;
    if (g >= n) return;
    const bool has_xr = present & (1u << 2), has_x0 = present & (1u << 4), has_xo = present & (1u << 6);
    const bool has_phi = present & (1u << 9), has_cs = (present & (3u << 10)) == (3u << 10), has_v = present & (1u << 12);
    const bool has_box = present & (1u << 14);
    const double2 xy = make_double2(sg.f[0][g], sg.f[1][g]);
    const double R = sg.f[8][g];
    if (!(R > 0.0)) atomicOr(sg.flag, 1);
    double phi, c, s;
    if (has_phi) phi = sg.f[9][g];
    if (has_cs) { c = sg.f[10][g]; s = sg.f[11][g]; }
    if (!has_phi) phi = atan2(s, c);
    if (!has_cs) sincos(phi, &s, &c);                       // jamming.cpp:332-333
    const double2 xr = has_xr ? make_double2(sg.f[2][g], sg.f[3][g]) : xy;
    const double2 x0 = has_x0 ? make_double2(sg.f[4][g], sg.f[5][g]) : xr;
    const double2 xo = has_xo ? make_double2(sg.f[6][g], sg.f[7][g]) : xy;
    const double2 v = has_v ? make_double2(sg.f[12][g], sg.f[13][g]) : make_double2(0.0, 0.0);
    int id, box = -1;
    if (slab) {
        id = sg.ids[g];
        if (id < 0 || id >= st.N) { atomicOr(sg.flag, 2); id = 0; }
    } else {
        const int sys = (int)(g / st.N);
        id = (int)(g - (long long)sys * st.N);
        if (has_box) {                                      // reference numbering i + j*b -> internal cy + cx*b
            const SysCtl* __restrict__ ctl = st.ctl + sys;
            const int bx = sg.box[g], b = ctl->b;
            if (bx >= 0 && bx < ctl->nbox) { const int qx = bx % b, qy = bx / b; box = qy + qx * b; }
        }
    }
    st.XY[0][g] = xy;
    st.CS[0][g] = make_double2(c, s);
    st.XR[0][g] = xr;
    st.X0[0][g] = x0;
    st.XO[0][g] = xo;
    st.V[0][g] = v;
    st.RR[0][g] = make_double2(R, 1.0 / R);                 // Cell::R, Cell::Rinv (jamming.cpp:297-298)
    st.PHI[0][g] = phi;
    st.ID[0][g] = id;
    st.BOX[0][g] = box;
}

// want: bit k set -> plane k is read back (bit 14: box). by_id: scatter to system * N + id (periodic box);
// otherwise device order (slab: the ids travel separately).
__global__ void __launch_bounds__(IO_TB) apj_unpack_kernel(const DevState st, const ApjStage sg, const unsigned want,
                                                           const int by_id) {
    const long long g = (long long)blockIdx.x * IO_TB + threadIdx.x;
    if (g >= st.ntot) return;
    const int sys = (int)(g / st.cap);
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    if (g - ctl->p0 >= ctl->n_own) return;
    const int cur = ctl->cur, gen = ctl->gen;
    const long long d = by_id ? (long long)sys * st.N + st.ID[gen][g] : g;
    if (want & (3u << 0)) { const double2 a = st.XY[cur][g]; sg.f[0][d] = a.x; sg.f[1][d] = a.y; }
    if (want & (3u << 2)) { const double2 a = st.XR[cur][g]; sg.f[2][d] = a.x; sg.f[3][d] = a.y; }
    if (want & (3u << 4)) { const double2 a = st.X0[gen][g]; sg.f[4][d] = a.x; sg.f[5][d] = a.y; }
    if (want & (3u << 6)) { const double2 a = st.XO[gen][g]; sg.f[6][d] = a.x; sg.f[7][d] = a.y; }
    if (want & (1u << 8)) sg.f[8][d] = st.RR[gen][g].x;
    if (want & (1u << 9)) sg.f[9][d] = st.PHI[gen][g];
    if (want & (3u << 10)) { const double2 a = st.CS[cur][g]; sg.f[10][d] = a.x; sg.f[11][d] = a.y; }
    if (want & (3u << 12)) { const double2 a = st.V[gen][g]; sg.f[12][d] = a.x; sg.f[13][d] = a.y; }
    if (want & (1u << 14)) {
        const int bi = st.BOX[gen][g], b = ctl->b;
        sg.box[d] = bi < 0 ? -1 : (bi / b + ctl->col0) + (bi % b) * b;   // reference numbering i + j*b of the GLOBAL grid
    }
}

// 64-bit fingerprint of the state by particle id: sum over particles of mix(id, bits of x, y, cos, sin).
// Addition is commutative, so the value does not depend on the particle order or on how the box is cut
// into slabs; ranks add their shares.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ull;
    z ^= z >> 27; z *= 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(IO_TB) apj_checksum_kernel(const DevState st, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    for (long long g = (long long)blockIdx.x * IO_TB + threadIdx.x; g < st.ntot; g += (long long)gridDim.x * IO_TB) {
        const int sys = (int)(g / st.cap);
        const SysCtl* __restrict__ ctl = st.ctl + sys;
        if (g - ctl->p0 >= ctl->n_own) continue;
        const double2 xy = st.XY[ctl->cur][g], cs = st.CS[ctl->cur][g];
        unsigned long long h = mix64(((unsigned long long)sys << 32 | (unsigned)st.ID[ctl->gen][g]) + 0x9e3779b97f4a7c15ull);
        h = mix64(h ^ (unsigned long long)__double_as_longlong(xy.x));
        h = mix64(h ^ (unsigned long long)__double_as_longlong(xy.y));
        h = mix64(h ^ (unsigned long long)__double_as_longlong(cs.x));
        h = mix64(h ^ (unsigned long long)__double_as_longlong(cs.y));
        acc += h;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

}  // namespace

void apj_launch_pack(const DevState& st, cudaStream_t s, const ApjStage& sg, unsigned present, long long n, int slab) {
    if (n > 0) apj_pack_kernel<<<(unsigned)((n + IO_TB - 1) / IO_TB), IO_TB, 0, s>>>(st, sg, present, n, slab);
}
void apj_launch_unpack(const DevState& st, cudaStream_t s, const ApjStage& sg, unsigned want, int by_id) {
    apj_unpack_kernel<<<(unsigned)((st.ntot + IO_TB - 1) / IO_TB), IO_TB, 0, s>>>(st, sg, want, by_id);
}
void apj_launch_checksum(const DevState& st, cudaStream_t s, unsigned long long* out) {
    const long long blocks = (st.ntot + IO_TB - 1) / IO_TB;
    apj_checksum_kernel<<<(unsigned)(blocks < 148 * 8 ? (blocks > 0 ? blocks : 1) : 148 * 8), IO_TB, 0, s>>>(st, out);
}
