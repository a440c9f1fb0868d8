// Device-side setup and the one per-frame quantity of the Ovito dump:
//
//   Engine::initCells  (reference code/jam/jamming.cpp:285-354, 2D branch): polydisperse radii 1 + N(0,1)/10, box
//                      length from the packing fraction, jittered offset-row lattice, random polarity -- drawn from
//                      the counter-based Philox stream (the reference's time-seeded mt19937 stream is irreproducible
//                      by construction, SURVEY 8c: RNG parity is distributional), so a 16M-particle run never
//                      builds vector<Cell> on the host;
//   Engine::topology   (:356-410): box centres and the 3x3 periodic neighbour table in the reference's numbering
//                      p = i + j*b, for callers that want the table (the kernels derive both on the fly);
//   Cell::over         (:653-656, :855-870; classes/Print.h:129-148): the overlap hue of print_video.
#include <algorithm>
#include "apj_device.cuh"

namespace {

constexpr int SU_TB = 256;

// variates of particle (sys, i): {radius normal, x jitter normal, y jitter normal, polarity angle}
struct InitDraw { double zR, zx, zy, phi; };
__device__ __forceinline__ double u01(unsigned w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }   // (0, 1)
__device__ __forceinline__ InitDraw init_draw(unsigned i, unsigned sys, unsigned long long seed, bool all) {
    const unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
    InitDraw d;
    const uint4 a = apj_philox4(i, 0x696e6974u, 0u, sys, k0, k1);       // counter word 1 = 'init': disjoint from the step stream
    const double r1 = sqrt(-2.0 * log(u01(a.x)));
    double s1, c1;
    sincospi(2.0 * u01(a.y), &s1, &c1);
    d.zR = r1 * c1;                                                      // Box-Muller: two independent normals per pair
    d.zx = r1 * s1;
    d.zy = 0.0; d.phi = 0.0;
    if (all) {
        const double r2 = sqrt(-2.0 * log(u01(a.z)));
        d.zy = r2 * cospi(2.0 * u01(a.w));
        const uint4 b = apj_philox4(i, 0x696e6974u, 1u, sys, k0, k1);
        d.phi = apj_u32_to_randuni(b.x);                                 // randuni() (:331)
    }
    return d;
}

// sum of R^2 per system: block partials, then one block per system (fixed order)
__global__ void __launch_bounds__(SU_TB) apj_radii_partial_kernel(const long long n, const int bps, const unsigned long long seed, double* __restrict__ part) {
    const int sys = blockIdx.x / bps, blk = blockIdx.x - sys * bps;
    double a = 0.0;
    for (long long i = (long long)blk * SU_TB + threadIdx.x; i < n; i += (long long)bps * SU_TB) {
        const double R = 1. + init_draw((unsigned)i, (unsigned)sys, seed, false).zR / 10;   // :296
        a += R * R;
    }
    __shared__ double sa[SU_TB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sa[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < SU_TB / 32; w++) a += sa[w];
        part[blockIdx.x] = a;
    }
}
__global__ void __launch_bounds__(SU_TB) apj_radii_final_kernel(const int bps, const double* __restrict__ part, double* __restrict__ out) {
    const int sys = blockIdx.x;
    double a = 0.0;
    for (int k = threadIdx.x; k < bps; k += SU_TB) a += part[(size_t)sys * bps + k];
    __shared__ double sa[SU_TB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sa[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < SU_TB / 32; w++) a += sa[w];
        out[sys] = a;
    }
}

__global__ void __launch_bounds__(SU_TB) apj_init_lattice_kernel(const DevState st, const unsigned long long seed) {
    const long long g = (long long)blockIdx.x * SU_TB + threadIdx.x;
    if (g >= st.ntot) return;
    const int sys = (int)(g / st.N);
    const int i = (int)(g - (long long)sys * st.N);
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    const double L = ctl->L, Lover2 = ctl->Lover2;
    const InitDraw d = init_draw((unsigned)i, (unsigned)sys, seed, true);
    const double R = 1. + d.zR / 10;                                     // :296
    const int rootN = (int)sqrt((double)st.N);                           // :311
    const double spacing = L / rootN;
    const int j = i / rootN;
    double x = -Lover2 + spacing * (i % rootN) + d.zx / 10.;             // :322-324
    double y = -Lover2 + spacing * j + d.zy / 10.;
    if (j % 2 == 0) x += 1.0;
    double phi = d.phi;                                                  // :327-329; periodicAngles (Cell.h:160-166)
    if (phi >= APJ_PI) phi -= APJ_PI2; else if (phi < -APJ_PI) phi += APJ_PI2;
    if (x >= Lover2) x -= L; else if (x < -Lover2) x += L;               // PBC, single wrap (Cell.h:168-175)
    if (y >= Lover2) y -= L; else if (y < -Lover2) y += L;
    double sn, cs;
    sincos(phi, &sn, &cs);
    const double2 xy = make_double2(x, y);
    st.XY[0][g] = xy; st.XR[0][g] = xy; st.X0[0][g] = xy; st.XO[0][g] = xy;
    st.CS[0][g] = make_double2(cs, sn);
    st.V[0][g] = make_double2(0.0, 0.0);
    st.RR[0][g] = make_double2(R, 1.0 / R);                              // :297-298
    st.PHI[0][g] = phi;
    st.ID[0][g] = i;
    st.BOX[0][g] = -1;
}

__global__ void __launch_bounds__(SU_TB) apj_box_table_kernel(const DevState st, const int sys, double* __restrict__ centres, int* __restrict__ nbrs) {
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    const int b = ctl->b;
    const double L = ctl->L, Lh = ctl->Lover2;
    for (int p = blockIdx.x * SU_TB + threadIdx.x; p < b * b; p += gridDim.x * SU_TB) {
        const int i = p % b, j = p / b;                                  // serial index p = i + j*b (:370-374)
        const double mn0 = -Lh + i * L / b, mx0 = -Lh + (i + 1) * L / b;   // :379-382
        const double mn1 = -Lh + j * L / b, mx1 = -Lh + (j + 1) * L / b;
        centres[2 * (size_t)p] = (mn0 + mx0) / 2.;                       // :385
        centres[2 * (size_t)p + 1] = (mn1 + mx1) / 2.;
        for (int dj = -1; dj <= 1; dj++)
            for (int di = -1; di <= 1; di++) {
                const int ii = (i + di + b) % b, jj = (j + dj + b) % b;
                nbrs[9 * (size_t)p + (di + 1) + (dj + 1) * 3] = ii + jj * b;   // :392-409
            }
    }
}

// Cell::over as neighborInteractions accumulates it on a filmed step (:653-656): starting from 240, every
// overlapping pair subtracts 240 |overlap| from BOTH partners -- into an int, so each update truncates and the
// result depends on the order of the updates. The reference's order for particle p: partners with a lower index
// first, ascending (they reach p from their own lists as the outer loop climbs), then p's own half list, which
// buildVerletLists filled neighbour box by neighbour box (offset (di,dj) at position (di+1) + 3 (dj+1)), ascending
// index inside a box. One thread per particle collects its overlapping partners from the full list, orders them
// that way and replays the integer updates.
__global__ void __launch_bounds__(128) apj_overlap_hue_kernel(const DevState st, int* __restrict__ over_by_id) {
    extern __shared__ __align__(16) unsigned char hue_smem[];
    const int sys = blockIdx.x / st.maxblk, blk = blockIdx.x - sys * st.maxblk;
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    if (blk >= ctl->nblk) return;
    const TileDesc d = st.tiles[(long long)sys * st.maxblk + blk];
    const int cur = ctl->cur, gen = ctl->gen, b = ctl->b;
    const double L = ctl->L, Lh = ctl->Lover2;
    const int lgG = st.G == 1 ? 0 : (st.G == 2 ? 1 : (st.G == 4 ? 2 : 3));
    const unsigned short* __restrict__ lst = reinterpret_cast<const unsigned short*>(st.list32 + ((long long)sys * st.maxblk + blk) * st.max_quads * st.tb * 4);
    for (int t = threadIdx.x; t < d.n; t += blockDim.x) {
        const long long g = (long long)d.g0 + t;
        const double2 me = st.XY[cur][g];
        const double Ri = st.RR[gen][g].x;
        const int my_id = st.ID[gen][g], my_box = st.BOX[gen][g];
        const int mcx = my_box / b, mcy = my_box - mcx * b;
        constexpr int MAXP = 24;
        long long key[MAXP];
        double val[MAXP];
        int np = 0;
        const int n = st.cnt[g];
        for (int e = 0; e < n; e++) {
            const int wd = e >> 1, sub = wd & (st.G - 1), kk = wd >> lgG;
            int slot = (int)(lst[(((size_t)(kk >> 2) * st.tb + (size_t)t * st.G + sub) * 4 + (kk & 3)) * 2 + (e & 1)] >> 4) - 1;
            long long jg = -1;
            for (int p = 0; p < (d.info & 0xff); p++) {
                if (slot < d.plen[p]) { jg = (long long)d.pstart[p] + slot; break; }
                slot -= d.plen[p];
            }
            if (jg < 0) continue;
            const double2 q = st.XY[cur][jg];
            const double dx = apj_delta_norm(q.x - me.x, L, Lh), dy = apj_delta_norm(q.y - me.y, L, Lh);
            const double d2 = apj_d2(dx, dy);
            if (!(d2 < st.rn2)) continue;
            const double sumR = Ri + st.RR[gen][jg].x;
            if (!(d2 < sumR * sumR)) continue;
            const double overlap = sumR / sqrt(d2) - 1;                  // :643, literally
            const int jid = st.ID[gen][jg];
            long long k;
            if (jid < my_id) k = jid;                                    // reached from j's list while the outer loop is at j < i
            else {                                                       // my own half list: neighbour-box order, then index
                const int jb = st.BOX[gen][jg];
                const int jcx = jb / b, jcy = jb - (jb / b) * b;
                int di = jcx - mcx, dj = jcy - mcy;
                if (di > 1) di -= b; else if (di < -1) di += b;
                if (dj > 1) dj -= b; else if (dj < -1) dj += b;
                k = ((long long)1 << 40) + ((long long)((di + 1) + (dj + 1) * 3) << 32) + jid;
            }
            if (np < MAXP) { key[np] = k; val[np] = 240 * fabs(overlap); np++; }
        }
        for (int a = 1; a < np; a++) {                                   // insertion sort by key
            const long long ka = key[a]; const double va = val[a];
            int q = a - 1;
            while (q >= 0 && key[q] > ka) { key[q + 1] = key[q]; val[q + 1] = val[q]; q--; }
            key[q + 1] = ka; val[q + 1] = va;
        }
        int over = 240;
        for (int a = 0; a < np; a++) over = (int)((double)over - val[a]);   // int -= double: truncates at every update
        over_by_id[(long long)sys * st.N + my_id] = over;
    }
}

}  // namespace

void apj_launch_radii_sums(cudaStream_t s, long long n, int n_sys, unsigned long long seed, double* d_part, double* d_sum) {
    const int bps = (int)std::max<long long>(1, std::min<long long>((n + SU_TB - 1) / SU_TB, 1024));
    apj_radii_partial_kernel<<<n_sys * bps, SU_TB, 0, s>>>(n, bps, seed, d_part);
    apj_radii_final_kernel<<<n_sys, SU_TB, 0, s>>>(bps, d_part, d_sum);
}
void apj_launch_init_lattice(const DevState& st, cudaStream_t s, unsigned long long seed) {
    apj_init_lattice_kernel<<<(unsigned)((st.ntot + SU_TB - 1) / SU_TB), SU_TB, 0, s>>>(st, seed);
}
void apj_launch_box_table(const DevState& st, cudaStream_t s, int sys, double* d_centres, int* d_neighbors) {
    apj_box_table_kernel<<<148 * 4, SU_TB, 0, s>>>(st, sys, d_centres, d_neighbors);
}
void apj_launch_overlap_hue(const DevState& st, cudaStream_t s, int* d_over_by_id) {
    apj_overlap_hue_kernel<<<st.n_sys * st.maxblk, 128, 0, s>>>(st, d_over_by_id);
}
