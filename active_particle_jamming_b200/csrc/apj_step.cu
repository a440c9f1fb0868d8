// Fused step kernel: one launch = one Engine::calculate_next_positions() (reference
// code/jam/jamming.cpp:837-853) for every batched system, evaluated speculatively:
//
//   newSkinList (:587-617)          -> per-particle COM-corrected displacement, block top-2
//   neighborInteractions (:623-669) -> gather sweep over the full Verlet list (no atomics,
//                                      SURVEY Q18): soft-disk force + Vicsek sum, then
//                                      phi = atan2 + CTnoise * U[-PI,PI)
//   Cell::update (classes/Cell.h:92-118,157-175) -> wrap, sincos, propulsion, Euler, PBC
//   calculate_COM (:761-774)        -> block partial sums of x_real
//
// Work decomposition. A block (TB = 128 or 256 threads) owns PPB = TB/G consecutive particles of one cell
// column (TileDesc, built at every rebuild). Everything those particles can interact with
// lives in <= 6 contiguous runs of the cell-ordered arrays (three columns x rows
// cy0-1..cy1+1, split where y wraps). One thread issues TMA bulk copies (cp.async.bulk +
// mbarrier) of those runs of {x,y}, {cos,sin}, {R,1/R} and of the block's list words into
// shared memory: one memory round trip per block. The neighbour sweep runs out of shared
// memory with G lanes per particle (G = 1 for large systems; G = 2/4/8 spreads a small system
// over enough warps to hide latency), partial sums are combined with __shfl_xor and handed to
// one thread per particle for the transcendental epilogue, which runs lane-dense. Lists are
// distance-sorted, so the interacting partners come first and the force branch is coherent.
//
// Results go to the *other* half of the ping-pong buffers. The last block to finish reduces
// the per-block partials in a fixed order (deterministic) and either COMMITS the step (flips
// ctl.cur, advances ctl.step, stores COM) or, if the skin test of the state the step STARTED
// from says the lists were too old, marks the system stale and leaves the state untouched: the
// rebuild chain then re-bins from the intact input buffers and the step is re-run. The rebuild
// decision therefore falls on exactly the step the reference takes it on, with no host round
// trip and no extra pass over the particles.
#include "apj_device.cuh"
#include "apj_trig.h"

namespace {

struct PairAcc { double Fx, Fy, ax, ay; };

// Message passing between blocks (partial -> ticket atomic -> reader): an acquire-release fence at GPU scope is what the
// pattern needs; __threadfence() is fence.sc.gpu, whose sequentially consistent form (MEMBAR.SC + ERRBAR) sits ~1 us on
// the critical path of a 10 us step of a small system.
__device__ __forceinline__ void apj_fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// shared-state-space loads by 32-bit address (keeps the sweep free of generic->shared address
// arithmetic; the addresses derive from a register the mbarrier wait "produces", so the compiler
// cannot hoist them above the wait)
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// 1/sqrt(a) for the overlap term: hardware approximation (rsqrt.approx.f64: ~20 good bits) refined by ONE
// cubically convergent step, y1 = y0 + y0 e (1/2 + 3/8 e) with the residual e = 1 - a y0^2 formed by an fma:
// truncation 5/16 e^3 < 2^-61, rounding ~1 ulp -- the accuracy of CUDA's rsqrt() (1 ulp) at a third of its
// instructions and none of its special-case branches. a = 0 (coincident particles) gives NaN, as the
// reference's inf * 0 does; denormal a does not occur (a is a squared distance of O(1)).
__device__ __forceinline__ double apj_rsqrt(const double a) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
    const double h = a * y0;
    const double e = fma(-h, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double q = y0 * e;
    return fma(q, p, y0);
}

// top-2 of one non-negative value per lane over the warp (the multiset newSkinList's scan keeps,
// jamming.cpp:607-608). Non-negative doubles order like their bit patterns, so two 32-bit REDUX.MAX per
// value replace the five shuffle + merge rounds.
__device__ __forceinline__ void apj_warp_top2(const double v, double& t1, double& t2) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const unsigned who = __ballot_sync(0xffffffffu, hi == mh && lo == ml);
    const bool first = (int)(threadIdx.x & 31) == __ffs((int)who) - 1;       // one holder of the maximum steps aside
    const unsigned hi2 = first ? 0u : hi, lo2 = first ? 0u : lo;
    const unsigned nh = __reduce_max_sync(0xffffffffu, hi2);
    const unsigned nl = __reduce_max_sync(0xffffffffu, hi2 == nh ? lo2 : 0u);
    t1 = __hiloint2double((int)mh, (int)ml);
    t2 = __hiloint2double((int)nh, (int)nl);
}

// one neighbour: reference jamming.cpp:633-662 seen from particle i (gather form). `off` is the
// byte offset of the neighbour's 16-byte tile slot; offset 0 is the sentinel slot, parked at
// (1e300, 1e300) so that it fails the d2 < rn2 test like any far particle.
// interacting part of one neighbour (d2 < rn2 already established)
__device__ __forceinline__ void pair_force(PairAcc& a, const double Ri, const double dx, const double dy, const double d2,
                                           const unsigned at, const unsigned dCS, const unsigned dRR) {
    const double sumR = Ri + lds_f64(at + dRR);
    if (d2 < sumR * sumR) {                           // they also overlap (:641)
        // overlap = sumR / sqrt(d2) - 1 (:643), evaluated as sumR * rsqrt(d2) - 1: apj_rsqrt is
        // good to ~1 ulp, so the quotient differs from the reference's by <= ~2 ulp
        // (4e-16), far inside the 1e-12 gate, at a fraction of the instruction count.
        const double overlap = sumR * apj_rsqrt(d2) - 1;
        a.Fx -= overlap * dx;
        a.Fy -= overlap * dy;
    }
    const double2 cs = lds_f64x2(at + dCS);
    a.ax += cs.x;                                     // add up orientations of neighbours (:659-660)
    a.ay += cs.y;
}

// one list word = two neighbours: reference jamming.cpp:633-662 seen from particle i (gather
// form). Each 16-bit entry is the byte offset of the neighbour's 16-byte tile slot; offset 0 is
// the sentinel slot, parked at (1e300, 1e300) so that it fails the d2 < rn2 test like any far
// particle. Both position loads and distance tests are issued before either branch, so the two
// shared-memory round trips overlap.
template <bool WRAP>
__device__ __forceinline__ void word_terms(PairAcc& a, const double2 me, const double Ri, const unsigned word,
                                           const unsigned sXY, const unsigned dCS, const unsigned dRR,
                                           const double L, const double Lh, const double rn2) {
    const unsigned at0 = sXY + (word & 0xffffu), at1 = sXY + (word >> 16);
    const double2 q0 = lds_f64x2(at0), q1 = lds_f64x2(at1);
    double dx0 = q0.x - me.x, dy0 = q0.y - me.y, dx1 = q1.x - me.x, dy1 = q1.y - me.y;
    if (WRAP) {                                       // delta_norm (:872-880)
        dx0 = apj_wrap1(dx0, L, Lh); dy0 = apj_wrap1(dy0, L, Lh);
        dx1 = apj_wrap1(dx1, L, Lh); dy1 = apj_wrap1(dy1, L, Lh);
    }
    const double d20 = apj_d2(dx0, dy0), d21 = apj_d2(dx1, dy1);
    if (d20 < rn2) pair_force(a, Ri, dx0, dy0, d20, at0, dCS, dRR);   // they're neighbours (:637)
    if (d21 < rn2) pair_force(a, Ri, dx1, dy1, d21, at1, dCS, dRR);
}

constexpr int QREG = 2;   // list quads (4 words = 8 entries each) a thread holds in registers at any time

// one quad of list words (bounds: the lane's list has nw words, this quad starts at word w0)
template <bool WRAP>
__device__ __forceinline__ void quad_terms(PairAcc& acc, const uint4 qd, const int w0, const int nw, const double2 me, const double Ri,
                                           const unsigned sXY, const unsigned dCS, const unsigned dRR, const double L, const double Lh,
                                           const double rn2) {
    word_terms<WRAP>(acc, me, Ri, qd.x, sXY, dCS, dRR, L, Lh, rn2);
    if (w0 + 1 < nw) word_terms<WRAP>(acc, me, Ri, qd.y, sXY, dCS, dRR, L, Lh, rn2);
    if (w0 + 2 < nw) word_terms<WRAP>(acc, me, Ri, qd.z, sXY, dCS, dRR, L, Lh, rn2);
    if (w0 + 3 < nw) word_terms<WRAP>(acc, me, Ri, qd.w, sXY, dCS, dRR, L, Lh, rn2);
}

// nw = list words of this lane; q[] = its first two quads (requested before the tile landed); gq = the lane's quad
// column in global memory (stride TB). Two quads rotate through the registers: while one is swept, the one after
// the next is already on its way into the registers of the one just finished (a quad is ~200 instructions of
// sweep, far longer than the load), so lists of any length run at the same pace on 8 list registers.
template <int TB, bool WRAP>
__device__ __forceinline__ void sweep(PairAcc& acc, const int nw, const uint4 (&q)[QREG], const uint4* __restrict__ gq,
                                      const double2 me, const double Ri, const unsigned sXY, const unsigned dCS,
                                      const unsigned dRR, const double L, const double Lh, const double rn2) {
    uint4 qa = q[0], qb = q[1];
    for (int w0 = 0; w0 < nw; w0 += 8) {
        quad_terms<WRAP>(acc, qa, w0, nw, me, Ri, sXY, dCS, dRR, L, Lh, rn2);
        if (w0 + 8 < nw) qa = __ldg(gq + (size_t)((w0 >> 2) + 2) * TB);
        if (w0 + 4 < nw) {
            quad_terms<WRAP>(acc, qb, w0 + 4, nw, me, Ri, sXY, dCS, dRR, L, Lh, rn2);
            if (w0 + 12 < nw) qb = __ldg(gq + (size_t)((w0 >> 2) + 3) * TB);
        }
    }
}

// phi = atan2(y_new, x_new) + CTnoise * randuni()   (jamming.cpp:667), then Cell::update: periodicAngles (single
// wrap, Cell.h:160-166), cosp = cos(phi), sinp = sin(phi). (ax, ay) = alignment sum, nz = the noise term. phi itself
// is only formed when the caller wants it (observables / downloads) or the fast form does not apply.
__device__ __forceinline__ void apj_new_orientation(const double ax, const double ay, const double nz, const bool want_phi,
                                                    double& phi, double& cs, double& sn) {
    phi = 0.0;
    const double r2 = ax * ax + ay * ay;
#ifndef APJ_EXACT_TRIG
    const double anz = fabs(nz);
    const bool fast = anz >= 1e-7 && anz <= 3.0 && r2 > 1e-200 && r2 < 1e200;
#else
    const bool fast = false;
#endif
    if (fast) {
        // cos and sin of theta + eta by the angle-addition identity, theta = atan2(ay, ax) never formed:
        // (cos theta, sin theta) = (ax, ay) / r, one rsqrt and one sincos(eta) instead of atan2 + sincos.
        // Both forms are within ~1e-15 of the exact value (measured 1.2e-15 apart over 1.2e7 samples),
        // far inside the 1e-12 gate. The reference's wrap uses truncated constants: when theta + eta
        // leaves [-PI, PI) it shifts phi by PI2 = 6.28318531, i.e. rotates (cos, sin) by PI2 - 2 pi.
        // Whether it wraps follows from the signs (theta = atan2(ay, ax) carries the sign BIT of ay, so
        // ay = -0.0 with ax < 0 is theta = -pi): for eta > 0 (theta >= 0 required) phi >= PI iff
        // sin(phi) < 0, or cos(phi) < 0 and sin(phi) <= sin(PI); mirrored for eta < 0.
        double sn_n, cn_n;
        apj_sincos_pm3(nz, &sn_n, &cn_n);                 // |nz| <= 3 here (apj_trig.h): half the instructions of sincos()
        const double inv = rsqrt(r2);
        const double c0 = ax * inv, s0 = ay * inv;
        cs = c0 * cn_n - s0 * sn_n;
        sn = s0 * cn_n + c0 * sn_n;
        constexpr double SIN_PI = 3.5897930298416118e-09;    // sin(3.14159265) evaluated on the double constant
        constexpr double DPI2 = 2.8204138795420667e-09;      // 6.28318531 (as a double) - 2 pi
        if (nz > 0.0 && !signbit(ay) && (sn < 0.0 || (cs < 0.0 && sn <= SIN_PI))) {
            const double c = cs; cs = c + sn * DPI2; sn = sn - c * DPI2;      // phi -= PI2
        } else if (nz < 0.0 && signbit(ay) && (sn > 0.0 || (cs < 0.0 && sn > -SIN_PI))) {
            const double c = cs; cs = c - sn * DPI2; sn = sn + c * DPI2;      // phi += PI2
        }
    }
    if (!fast || want_phi) {
        phi = atan2(ay, ax) + nz;
        if (phi >= APJ_PI) phi -= APJ_PI2;
        else if (phi < -APJ_PI) phi += APJ_PI2;
        if (!fast) sincos(phi, &sn, &cs);
    }
}

// End of a step, one thread per system: a = {sum x_real, sum y_real, top-1, top-2 displacement^2} of the
// state the step STARTED from; kcls = the distance classes the launch swept.
__device__ __forceinline__ void apj_commit(SysCtl* __restrict__ ctl, const DevState& st, const double4 a, const int kcls) {
    const double D = sqrt(a.z) + sqrt(a.w);
    if (D > st.skin) {                       // jamming.cpp:611
        ctl->stale = 1;                      // lists too old for the state this step read:
        ctl->save_old = 1;                   // drop the speculative result, rebuild, re-run
        ctl->n_discarded += 1;
    } else if (kcls < APJ_CLASSES - 1 && apj_class_for(ctl->skinBase + D, st) > kcls) {
        ctl->kmin = apj_class_for(ctl->skinBase + D, st);   // swept too few classes for this state: drop the result, re-run longer
        ctl->n_retried += 1;
    } else {
        ctl->COM[0] = a.x / st.N;            // calculate_COM (jamming.cpp:761-774)
        ctl->COM[1] = a.y / st.N;
        ctl->cur ^= 1;
        ctl->step += 1;
        ctl->no_self_once = 0;
        ctl->skinDD = fmax(D - ctl->skinD, 0.0);
        ctl->skinD = D;
        ctl->kmin = 0;
        ctl->n_class[kcls] += 1;
    }
}

// one warp: fold `cnt` partials {sum x, sum y, top-1, top-2} in a fixed order -- lane l takes l, l + 32, ..., eight
// independent loads in flight, then a shuffle tree. Every lane returns the total.
__device__ __forceinline__ double4 apj_fold_partials(const double4* __restrict__ src, const int cnt) {
    const int lane = threadIdx.x & 31;
    double4 a = make_double4(0.0, 0.0, 0.0, 0.0);
    for (int k0 = 0; k0 < cnt; k0 += 32 * 8) {
        double2 v01[8], v23[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int k = k0 + u * 32 + lane;
            v01[u] = make_double2(0.0, 0.0); v23[u] = make_double2(0.0, 0.0);
            if (k < cnt) {
                v01[u] = __ldcg(reinterpret_cast<const double2*>(src + k));
                v23[u] = __ldcg(reinterpret_cast<const double2*>(src + k) + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a.x += v01[u].x; a.y += v01[u].y;
            apj_top2_merge(a.z, a.w, v23[u].x, v23[u].y);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
        const double b1 = __shfl_xor_sync(0xffffffffu, a.z, o), b2 = __shfl_xor_sync(0xffffffffu, a.w, o);
        apj_top2_merge(a.z, a.w, b1, b2);
    }
    return a;
}

#ifndef APJ_PF_DIST
#define APJ_PF_DIST 0   // blocks ahead; 0 = off. Measured on B200 (N = 16M): 444 (3 per SM) cuts the kernel 6 % from a cold L2 (ncu)
                       // but costs ~1.5 % in steady state, where consecutive launches overlap -- kept as a build option
#endif
#ifndef APJ_BLOCKS_256
#define APJ_BLOCKS_256 4
#endif
#ifndef APJ_BLOCKS_128
#define APJ_BLOCKS_128 8
#endif
#ifndef APJ_BLOCKS_192
#define APJ_BLOCKS_192 6   // 56 registers per thread: 36 warps per SM
#endif
#ifndef APJ_BLOCKS_224
#define APJ_BLOCKS_224 5   // 56 registers per thread: 35 warps per SM
#endif
constexpr int apj_blocks_for(int tb) { return tb == 512 ? 2 : tb == 256 ? APJ_BLOCKS_256 : (tb == 224 ? APJ_BLOCKS_224 : (tb == 192 ? APJ_BLOCKS_192 : APJ_BLOCKS_128)); }
template <int TB, int G, bool INJECT, bool SLAB, bool SPLIT>
__global__ void __launch_bounds__(TB, apj_blocks_for(TB))
apj_step_kernel(const DevState st, const double* __restrict__ noise_by_id, const int always_full) {
    constexpr int PPB = TB / G;                        // particles per block
    constexpr int EWARPS = (PPB + 31) / 32;            // warps that run the epilogue
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ TileDesc sd;
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ double4 s_red[EWARPS];

    const int sys = st.n_sys == 1 ? 0 : blockIdx.x / st.maxblk;
    int blk = (int)(blockIdx.x - sys * st.maxblk);
    if (SLAB) {   // the grid starts with the last owned column, then wraps to the first: both halo-pushing columns run early
        const int nb = st.ctl[sys].nblk, sh = st.ctl[sys].blk_shift;
        if (blk < nb) { blk += sh; if (blk >= nb) blk -= nb; }
    }
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const long long bg = (long long)sys * st.maxblk + blk;
    int desc_word = 0;                                 // fetched alongside ctl: one round trip, not two
    if (t < 16) desc_word = apj_ldg_l2keep(reinterpret_cast<const int*>(st.tiles + bg) + t);
    // Cross-block software prefetch. A step streams ~3 GB through the 126 MB L2, so nothing a block needs
    // is still cached from the previous step and its head would be two dependent HBM round trips
    // (descriptor, then tile + lists). Blocks are dispatched in index order, so each block asks L2 for
    // what block blk + PF will need about one block lifetime later: the last warp reads that block's
    // descriptor (itself prefetched by block blk - PF) and issues bulk L2 prefetches of its tile pieces,
    // list quads and per-particle arrays, and touches the descriptor of block blk + 2 PF.
    constexpr int PF = APJ_PF_DIST;
    int fdesc = 0;
    if (PF > 0 && wid == TB / 32 - 1) {
        if (lane < 16 && blk + PF < st.maxblk) fdesc = __ldg(reinterpret_cast<const int*>(st.tiles + bg + PF) + lane);
        if (lane == 16 && blk + 2 * PF < st.maxblk) apj_prefetch_l2(st.tiles + bg + 2 * PF);
    }
    SysCtl* __restrict__ ctl = st.ctl + sys;
    const int nblk = ctl->nblk;
    const long long step = ctl->step;
    if (blk >= nblk || ctl->stale || step >= ctl->target) return;  // uniform over the system's blocks

    const int cur = ctl->cur, gen = ctl->gen;
    __shared__ int s_kcls;                             // distance classes this launch sweeps (0..kcls): uniform, derived by one thread
    // tile: slot 0 is the sentinel, slots 1.. are the concatenated pieces; three arrays of 16-byte
    // records {x,y}, {cos,sin}, {R,1/R}, each tile_cap+1 records long
    double2* __restrict__ sXYp = reinterpret_cast<double2*>(smem_raw);
    double2* __restrict__ sCSp = sXYp + (st.tile_cap + 1);
    double2* __restrict__ sRRp = sCSp + (st.tile_cap + 1);
    double4* __restrict__ sAcc = reinterpret_cast<double4*>(sRRp + (st.tile_cap + 1));   // G > 1 only

    // Stage the tile: <= 6 pieces x 3 arrays, one TMA bulk copy each. Warp 0 holds the descriptor in
    // registers (lanes 0..15); lane 3*p + a issues the copy of piece p of array a, so the <= 18 copies
    // leave in one instruction instead of a serial loop on one thread, and they leave BEFORE the block
    // barrier that publishes the descriptor to the other warps.
    if (wid == 0) {
        if (lane < 16) reinterpret_cast<int*>(&sd)[lane] = desc_word;
        if (lane == 0) { apj_mbar_init(&s_bar, 1); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
        __syncwarp();
        const int np = __shfl_sync(0xffffffffu, desc_word, 3) & 0xff;
        const int mp = lane / 3, arr = lane - mp * 3;
        int off = 1, my_off = 1, my_len = 0, my_start = 0;
#pragma unroll
        for (int q = 0; q < APJ_MAX_PIECES; q++) {
            const int start = __shfl_sync(0xffffffffu, desc_word, 4 + q);
            const int len = __shfl_sync(0xffffffffu, desc_word, 4 + APJ_MAX_PIECES + q);
            if (q < np) {
                if (q == mp) { my_off = off; my_len = len; my_start = start; }
                off += len;
            }
        }
        if (lane == 0) apj_mbar_expect_tx(&s_bar, (unsigned)(off - 1) * 48u);
        __syncwarp();
        if (mp < np && my_len > 0) {
            const double2* src = arr == 0 ? st.XY[cur] : (arr == 1 ? st.CS[cur] : st.RR[gen]);
            double2* dst = arr == 0 ? sXYp : (arr == 1 ? sCSp : sRRp);
            apj_bulk_g2s(dst + my_off, src + my_start, (unsigned)my_len * 16u, &s_bar);
        }
    }
    if (t == 32 % TB) { sXYp[0] = make_double2(1e300, 1e300); s_kcls = apj_sweep_class(ctl, st); }
    __syncthreads();
    const int kcls = s_kcls;

    const int npieces = sd.info & 0xff;
    const bool wraps = (sd.info & APJ_INFO_WRAPS) != 0;
    (void)npieces;

    // per-thread loads that do not depend on the tile: in flight while the TMA copies land
    const int p = t / G, sub = t - p * G;              // sweep mapping: G adjacent lanes per particle
    const bool sweeping = p < sd.n;
    const int n = sweeping ? (int)((st.cntk[sd.g0 + p] >> (8 * kcls)) & 0xffu) : 0;
    const int nwords = (n + 1) >> 1;
    const int nw = (nwords - sub + G - 1) / G;         // list words of this lane (<= 0: none)
    const uint4* __restrict__ gq = reinterpret_cast<const uint4*>(st.list32) + bg * (long long)st.max_quads * TB + t;
    uint4 q[QREG];                                     // the first two quads do not wait for cnt (stale words are never swept)
#pragma unroll
    for (int j = 0; j < QREG; j++) q[j] = (j < st.max_quads) ? __ldg(gq + (size_t)j * TB) : make_uint4(0u, 0u, 0u, 0u);
    const bool active = t < sd.n;                      // epilogue mapping: thread t <-> particle t
    const long long g = (long long)sd.g0 + (active ? t : 0);
    // x_old / x_real are only needed at the very end of the epilogue. APJ_LATE_XO asks L2 for them here and loads them
    // right after the sweep, which keeps 8 registers out of the sweep: measured +0.6 % at 64 registers / 4 blocks per SM,
    // and the 48-register / 5-block build it was meant for gains only 3.6 % -- default: early.
#ifndef APJ_LATE_XO
#define APJ_EARLY_XO 1
#endif
#ifdef APJ_EARLY_XO
    double2 xo = make_double2(0.0, 0.0), xr = xo;
#endif
    int id = 0;
    if (t < PPB) {
#ifdef APJ_EARLY_XO
        xo = st.XO[gen][g]; xr = st.XR[cur][g];
#else
        apj_prefetch_l2(st.XO[gen] + g); apj_prefetch_l2(st.XR[cur] + g);
#endif
        id = st.ID[gen][g];
    }
    const double L = ctl->L, Lh = ctl->Lover2;
    const double rn2 = st.rn2;
    // randuni() of this step (jamming.cpp:667): depends on (id, step) only, so it is drawn while the tile is in flight
    double u = 0.0;
    if (t < PPB) {
        if (INJECT) {
            u = noise_by_id[(long long)sys * st.N + id];
        } else {
            const unsigned w = apj_philox_word0((unsigned)id, (unsigned)step, (unsigned)(step >> 32), (unsigned)sys,
                                                (unsigned)st.seed, (unsigned)(st.seed >> 32));
            u = apj_u32_to_randuni(w);
        }
    }

    if (PF > 0 && wid == TB / 32 - 1 && blk + PF < nblk) {   // this warp's own loads are in flight: ask L2 for block blk + PF
        const int fnp = __shfl_sync(0xffffffffu, fdesc, 3) & 0xff;
        const int fg0 = __shfl_sync(0xffffffffu, fdesc, 0), fn = __shfl_sync(0xffffffffu, fdesc, 1);
        const int mp = lane / 3, arr = lane - mp * 3;
        const int fstart = __shfl_sync(0xffffffffu, fdesc, 4 + (mp < APJ_MAX_PIECES ? mp : 0));
        const int flen = __shfl_sync(0xffffffffu, fdesc, 4 + APJ_MAX_PIECES + (mp < APJ_MAX_PIECES ? mp : 0));
        if (lane < 3 * APJ_MAX_PIECES) {
            if (mp < fnp && flen > 0) {
                const double2* src = arr == 0 ? st.XY[cur] : (arr == 1 ? st.CS[cur] : st.RR[gen]);
                apj_bulk_prefetch_l2(src + fstart, (unsigned)flen * 16u);
            }
        } else if (fn > 0 && fn <= PPB) {
            const int a = lane - 3 * APJ_MAX_PIECES;
            if (a == 0) apj_bulk_prefetch_l2(st.XO[gen] + fg0, (unsigned)fn * 16u);
            else if (a == 1) apj_bulk_prefetch_l2(st.XR[cur] + fg0, (unsigned)fn * 16u);
            else if (a == 2) apj_bulk_prefetch_l2(st.ID[gen] + fg0, (unsigned)fn * 4u);
            else if (a == 3) apj_bulk_prefetch_l2(st.cntk + fg0, (unsigned)fn * 4u);
            else if (a == 4) apj_bulk_prefetch_l2(reinterpret_cast<const uint4*>(st.list32) + (bg + PF) * (long long)st.max_quads * TB,
                                                  (unsigned)(st.max_quads < QREG ? st.max_quads : QREG) * TB * 16u);
        }
    }

    unsigned sXY = apj_smem_addr(sXYp);
    apj_mbar_wait(&s_bar, 0u, sXY);
    const unsigned dCS = (unsigned)(st.tile_cap + 1) * 16u, dRR = 2u * dCS;

    // ---- neighborInteractions ----
    PairAcc acc = {0.0, 0.0, 0.0, 0.0};
    {
        const unsigned own = (unsigned)(sd.own_slot + (sweeping ? p : 0)) * 16u;
        const double2 me = lds_f64x2(sXY + own);
        const double Ri = lds_f64(sXY + dRR + own);
        if (wraps) sweep<TB, true>(acc, nw, q, gq, me, Ri, sXY, dCS, dRR, L, Lh, rn2);
        else sweep<TB, false>(acc, nw, q, gq, me, Ri, sXY, dCS, dRR, L, Lh, rn2);
    }
    if (G > 1) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            acc.Fx += __shfl_xor_sync(0xffffffffu, acc.Fx, o);
            acc.Fy += __shfl_xor_sync(0xffffffffu, acc.Fy, o);
            acc.ax += __shfl_xor_sync(0xffffffffu, acc.ax, o);
            acc.ay += __shfl_xor_sync(0xffffffffu, acc.ay, o);
        }
        if (sub == 0) sAcc[p] = make_double4(acc.Fx, acc.Fy, acc.ax, acc.ay);
        __syncthreads();
        if (wid >= EWARPS) return;                     // only the epilogue warps continue
        const double4 a = sAcc[t < PPB ? t : 0];
        acc.Fx = a.x; acc.Fy = a.y; acc.ax = a.z; acc.ay = a.w;
    }

    double sum_x = 0.0, sum_y = 0.0, top1 = 0.0, top2 = 0.0;
    if (active) {
#ifndef APJ_EARLY_XO
        double2 xo, xr;                                // issued now, consumed at the end of the epilogue (L2 hits: prefetched in the head)
        asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(xo.x), "=d"(xo.y) : "l"(st.XO[gen] + g));
        asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(xr.x), "=d"(xr.y) : "l"(st.XR[cur] + g));
#endif
        const unsigned own = (unsigned)(sd.own_slot + t) * 16u;
        const double2 me = lds_f64x2(sXY + own), mcs = lds_f64x2(sXY + dCS + own), mrr = lds_f64x2(sXY + dRR + own);
        const double Ri = mrr.x;
        double Fx = acc.Fx, Fy = acc.Fy, ax = acc.ax, ay = acc.ay;
        if (!ctl->no_self_once) { ax += mcs.x; ay += mcs.y; }  // self term: Cell::update left x_new = cosp (Cell.h:102-103)

        // ---- phi = atan2(y_new, x_new) + CTnoise * randuni()   (jamming.cpp:667), then Cell::update:
        //      periodicAngles (single wrap, Cell.h:160-166), cosp = cos(phi), sinp = sin(phi) ----
        const double nz = ctl->CTnoise * u;
        const bool want_phi = always_full || step + 1 == ctl->target;   // phi itself is only read by observables / downloads
        double phi, sn, cs;
        apj_new_orientation(ax, ay, nz, want_phi, phi, cs, sn);
        double CF = ctl->CFself;
        if (ctl->ramp_len > 0) {                      // relax() ramp (jamming.cpp:518)
            const long long t_ = step - ctl->ramp_t0;
            if (t_ < ctl->ramp_len) CF = CF - (double)(ctl->ramp_len - t_) * CF / (double)ctl->ramp_len;
        }
        Fx += cs * CF * Ri;
        Fy += sn * CF * Ri;
        const double vx = Fx * mrr.y, vy = Fy * mrr.y;   // Rinv = 1/R stored at upload (jamming.cpp:298)
        const double dx = vx * st.dt, dy = vy * st.dt;
        double x = me.x + dx, y = me.y + dy;
        if (x >= Lh) x -= L; else if (x < -Lh) x += L;   // Cell::PBC, single wrap (Cell.h:168-175)
        if (y >= Lh) y -= L; else if (y < -Lh) y += L;

        st.XY[cur ^ 1][g] = make_double2(x, y);
        st.CS[cur ^ 1][g] = make_double2(cs, sn);
        // ---- newSkinList: displacement since the last rebuild, COM drift removed (of the state the step STARTED from) ----
        {
            const double ddx = apj_delta_norm(((me.x - xo.x) - ctl->COM[0]) + ctl->COM_old[0], L, Lh);
            const double ddy = apj_delta_norm(((me.y - xo.y) - ctl->COM[1]) + ctl->COM_old[1], L, Lh);
            top1 = apj_d2(ddx, ddy);
        }
        const double xrn = xr.x + dx, yrn = xr.y + dy;
        st.XR[cur ^ 1][g] = make_double2(xrn, yrn);
        if (SLAB) {   // halo exchange fused into the epilogue: boundary columns are stored straight into the
                      // neighbours' ghost slots (peer memory over NVLink), in the half they read next step
            if (sd.info & APJ_INFO_PUSH_LEFT) {    // column 0 starts at slot 0 -> left neighbour's RIGHT ghost column
                const long long pg = (long long)st.cap + st.gcap + g;
                apj_peer(st, st.left, st.XY[cur ^ 1])[pg] = make_double2(x, y);
                apj_peer(st, st.left, st.CS[cur ^ 1])[pg] = make_double2(cs, sn);
            }
            if (sd.info & APJ_INFO_PUSH_RIGHT) {   // last owned column -> right neighbour's LEFT ghost column
                const long long pg = (long long)st.cap + (g - ctl->last_col_start);
                apj_peer(st, st.right, st.XY[cur ^ 1])[pg] = make_double2(x, y);
                apj_peer(st, st.right, st.CS[cur ^ 1])[pg] = make_double2(cs, sn);
            }
        }
        if (want_phi) {   // fields only observables read
            st.V[gen][g] = make_double2(vx, vy);
            st.PHI[gen][g] = phi;
        }
        sum_x = xrn; sum_y = yrn;
    }

    // ---- block reduction: sum of x_real, top-2 displacement^2 ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum_x += __shfl_xor_sync(0xffffffffu, sum_x, o);
        sum_y += __shfl_xor_sync(0xffffffffu, sum_y, o);
    }
    apj_warp_top2(top1, top1, top2);
    if (EWARPS > 1) {
        if (lane == 0) s_red[wid] = make_double4(sum_x, sum_y, top1, top2);
        __syncthreads();       // all EWARPS == all warps of the block when EWARPS > 1 (G == 1 or 2)
        if (wid != 0) return;
        if (wid == 0) {
#pragma unroll
            for (int w = 1; w < EWARPS; w++) {
                const double4 b = s_red[w];
                sum_x += b.x; sum_y += b.y;
                apj_top2_merge(top1, top2, b.z, b.w);
            }
        }
    }
    // warp 0 only from here. Two-level, fixed-order reduction of the per-block partials: the last
    // block of each group of 32 folds its group, the last group to finish folds the groups and
    // commits. Deterministic (membership and order are fixed) and O(sqrt)-deep in the tail.
    if (SPLIT) {
        // Large systems: the block only leaves its partial. apj_reduce_commit_kernel (next in the stream)
        // folds them in a fixed order and commits, so no block ends with a fence + atomic round trip
        // during which one warp keeps the block's registers and shared memory alive (~10 % of the
        // resident warp slots at N = 16M).
        if (SLAB && (sd.info & (APJ_INFO_PUSH_LEFT | APJ_INFO_PUSH_RIGHT))) { __syncwarp(); __threadfence_system(); }
        if (lane == 0) st.partials[bg] = make_double4(sum_x, sum_y, top1, top2);
        return;
    }
    const int grp = blk >> 5, ngrp = (nblk + 31) >> 5;
    const int gsize = min(32, nblk - (grp << 5));
    unsigned* __restrict__ gticket = st.gticket + (long long)sys * st.maxgrp;
    double4* __restrict__ gpart = st.gpartials + (long long)sys * st.maxgrp;
    unsigned last = 0;
    if (SLAB) __syncwarp();   // peer stores of the other lanes are ordered before lane 0's system-scope fence
    if (lane == 0) {
        st.partials[bg] = make_double4(sum_x, sum_y, top1, top2);
        // only blocks that stored into a neighbour's ghost slots need the (expensive) system-scope fence
        if (SLAB && (sd.info & (APJ_INFO_PUSH_LEFT | APJ_INFO_PUSH_RIGHT))) __threadfence_system(); else apj_fence_gpu();
        last = (atomicAdd(gticket + grp, 1u) == (unsigned)gsize - 1u) ? 1u : 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    apj_fence_gpu();
    double4 a = make_double4(0.0, 0.0, 0.0, 0.0);
    if (lane < gsize) {
        const double4* q = st.partials + (long long)sys * st.maxblk + (grp << 5) + lane;
        const double2 b01 = __ldcg(reinterpret_cast<const double2*>(q));
        const double2 b23 = __ldcg(reinterpret_cast<const double2*>(q) + 1);
        a = make_double4(b01.x, b01.y, b23.x, b23.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
        const double b1 = __shfl_xor_sync(0xffffffffu, a.z, o), b2 = __shfl_xor_sync(0xffffffffu, a.w, o);
        apj_top2_merge(a.z, a.w, b1, b2);
    }
#ifndef APJ_NO_GROUP_SHORTCUT
    if (!SLAB && ngrp == 1) {
        // a system of <= 32 work blocks (N = 1024 .. 8192: the reference's own sizes, the replicas of a sweep) has ONE group:
        // its fold IS the system's -- folding the single group partial again adds zeros and merges (0, 0), the same bits --
        // so the second ticket, fence and round trip (~1.5 us of a ~10 us step) are skipped
        if (lane == 0) {
            gticket[grp] = 0u;
            apj_commit(ctl, st, a, kcls);
        }
        return;
    }
#endif
    last = 0;
    if (lane == 0) {
        gticket[grp] = 0u;
        gpart[grp] = a;
        if (SLAB) __threadfence_system(); else apj_fence_gpu();
        last = (atomicAdd(&ctl->ticket, 1u) == (unsigned)ngrp - 1u) ? 1u : 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;

    // ---- last group of this system: fold the group partials, then commit ----
    apj_fence_gpu();
    a = apj_fold_partials(gpart, ngrp);
    if (SLAB) {
        // this rank's partial goes to every rank of the box (own mailbox included); the commit kernel
        // that follows folds them in rank order, so all ranks take the same decision
        if (lane == 0) ctl->ticket = 0u;
        const unsigned long long ep = ctl->seq[0];
        if (lane < st.nranks) {
            SlabMail* m = apj_peer(st, lane, st.mail);
            m->part[ep & 1][st.rank] = a;                              // same lane: the release store below orders it
            apj_st_release_sys(&m->flag[0][st.rank], ep + 1);
        }
        return;
    }
    if (lane == 0) {
        ctl->ticket = 0u;
        apj_commit(ctl, st, a, kcls);
    }
    return;
}

// ---------------------------------------------------------------------------------------------------------
// Pipelined persistent form of the step (one large system, one lane per particle, split tail).
//
// The grid is ONE wave of resident blocks; block b walks the tiles b, b + grid, b + 2 grid, ... A block of the
// classic kernel above spends ~40 % of its life in its head -- descriptor round trip, then tile + lists + per-
// particle arrays, two dependent trips to HBM -- with nothing to compute. Here the trips of tile i+1 run under
// the epilogue of tile i:
//   * the tile buffer is only read by the neighbour sweep. The LAST warp of the block to finish its sweep
//     (shared-memory arrival counter, no block barrier) issues the TMA copies of the next tile straight into
//     the buffer, plus a 64-byte bulk copy of the descriptor of the tile after that, all completing on the one
//     mbarrier whose phase the next sweep waits for;
//   * the own particle's {cos,sin}, {R,1/R} are the last reads of the buffer; x_old / x_real are read from global
//     memory after the sweep -- L2 hits, because the thread asked L2 for them a tile earlier -- while Philox runs;
//   * the first list quad of the next tile travels to the thread's shared-memory slot with cp.async, list
//     lengths and ids are requested before the epilogue starts.
// Warps never wait for each other inside the loop: a warp waits for "tile landed", sweeps, signals, updates its
// particles, moves on. Per-lane running sums of x_real (shared memory) and per-warp running top-2 displacements
// are folded once, in a fixed order, when the block has run out of tiles: one partial per BLOCK.
// Same arithmetic in the same order per particle as the classic kernel: positions are bit-identical to it.
#ifndef APJ_PIPE_BLOCKS
#define APJ_PIPE_BLOCKS 4
#endif
struct PipeSmem {
    TileDesc sd[3];                  // descriptors of tiles i-1 / i / i+1 (slot = visit number % 3)
    unsigned long long full;         // mbarrier: tile (and the next descriptor) landed
    unsigned done;                   // warps that finished sweeping the current tile
    int kcls, pushed;
    double2 top[APJ_TB_MAX / 32];    // per warp: running top-2 displacement^2
    double4 red[APJ_TB_MAX / 32];
};

// whole warp: TMA copies of the tile described by d (shared memory) + the descriptor `next_src` (may be null)
__device__ __forceinline__ void pipe_issue(const DevState& st, const TileDesc* d, const int cur, const int gen, double2* sXYp, double2* sCSp,
                                           double2* sRRp, unsigned long long* bar, const TileDesc* next_src, TileDesc* next_dst) {
    const int lane = threadIdx.x & 31;
    const int np = d->info & 0xff;
    const int mp = lane / 3, arr = lane - mp * 3;
    int off = 1, my_off = 1, my_len = 0, my_start = 0;
#pragma unroll
    for (int q = 0; q < APJ_MAX_PIECES; q++) {
        if (q < np) {
            const int len = d->plen[q];
            if (q == mp) { my_off = off; my_len = len; my_start = d->pstart[q]; }
            off += len;
        }
    }
    if (lane == 0) apj_mbar_expect_tx(bar, (unsigned)(off - 1) * 48u + (next_src ? (unsigned)sizeof(TileDesc) : 0u));
    __syncwarp();
    if (mp < np && my_len > 0) {
        const double2* src = arr == 0 ? st.XY[cur] : (arr == 1 ? st.CS[cur] : st.RR[gen]);
        double2* dst = arr == 0 ? sXYp : (arr == 1 ? sCSp : sRRp);
        apj_bulk_g2s(dst + my_off, src + my_start, (unsigned)my_len * 16u, bar);
    }
    if (lane == 31 && next_src) apj_bulk_g2s(next_dst, next_src, (unsigned)sizeof(TileDesc), bar);
}

template <int TB, bool INJECT, bool SLAB>
__global__ void __launch_bounds__(TB, APJ_PIPE_BLOCKS)
apj_step_pipe_kernel(const DevState st, const double* __restrict__ noise_by_id, const int always_full) {
    constexpr int NW = TB / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(128) PipeSmem sm;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    SysCtl* __restrict__ ctl = st.ctl;
    const int nblk = ctl->nblk;
    const long long step = ctl->step;
    int tile = blockIdx.x;
    if (tile >= nblk || ctl->stale || step >= ctl->target) return;   // uniform over the block
    const int cur = ctl->cur, gen = ctl->gen;
    const int stride = gridDim.x;

    // tile: slot 0 is the sentinel, slots 1.. are the concatenated pieces; three arrays of 16-byte records, then the
    // per-thread running sums of x_real
    double2* __restrict__ sXYp = reinterpret_cast<double2*>(smem_raw);
    double2* __restrict__ sCSp = sXYp + (st.tile_cap + 1);
    double2* __restrict__ sRRp = sCSp + (st.tile_cap + 1);
    double2* __restrict__ sSum = sRRp + (st.tile_cap + 1);
    uint4* __restrict__ sQ0 = reinterpret_cast<uint4*>(sSum + TB);      // per thread: first list quad of the tile about to be swept

    if (t < 16) reinterpret_cast<int*>(&sm.sd[0])[t] = __ldg(reinterpret_cast<const int*>(st.tiles + tile) + t);
    if (t == 32) {
        apj_mbar_init(&sm.full, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        sm.done = 0u; sm.pushed = 0;
        sXYp[0] = make_double2(1e300, 1e300);
        sm.kcls = apj_sweep_class(ctl, st);
    }
    sSum[t] = make_double2(0.0, 0.0);
    if (lane == 0) sm.top[wid] = make_double2(0.0, 0.0);
    __syncthreads();
    if (wid == 0)
        pipe_issue(st, &sm.sd[0], cur, gen, sXYp, sCSp, sRRp, &sm.full, tile + stride < nblk ? st.tiles + tile + stride : nullptr, &sm.sd[1]);
    const int kcls = sm.kcls;
    const double L = ctl->L, Lh = ctl->Lover2;
    const double rn2 = st.rn2;
    const unsigned dCS = (unsigned)(st.tile_cap + 1) * 16u, dRR = 2u * dCS;

    // What a tile needs from global memory besides the tile itself is requested one tile ahead WITHOUT holding
    // registers across the epilogue: the thread's first list quad goes to its own shared-memory slot with
    // cp.async (the second one is loaded when the sweep starts and is consumed a quad later), its x_old / x_real
    // records are pulled into L2 so that the loads after the sweep are L2 hits that Philox covers.
    unsigned cntk_raw;                                 // list lengths of the next tile: used (shifted) only after the tile wait
    int id;
    auto load_early = [&](const TileDesc& d, const int tl) {
        const bool mine = t < d.n;
        const long long g = (long long)d.g0 + (mine ? t : 0);
        const uint4* gq = reinterpret_cast<const uint4*>(st.list32) + (long long)tl * st.max_quads * TB + t;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(apj_smem_addr(sQ0 + t)), "l"(gq) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        apj_prefetch_l2(st.XO[gen] + g);
        apj_prefetch_l2(st.XR[cur] + g);
        cntk_raw = mine ? st.cntk[g] : 0u;
        id = st.ID[gen][g];
    };
    load_early(sm.sd[0], tile);

    int k = 0;                                         // descriptor slot of the current tile
    unsigned phase = 0;
    for (;;) {
        const TileDesc& d = sm.sd[k];
        unsigned sXY = apj_smem_addr(sXYp);
        apj_mbar_wait(&sm.full, phase, sXY);
        phase ^= 1u;
        const bool active = t < d.n;
        const long long g = (long long)d.g0 + (active ? t : 0);
        const bool wraps = (d.info & APJ_INFO_WRAPS) != 0;
        const int my_id = id;
        const int nent = (int)((cntk_raw >> (8 * kcls)) & 0xffu);
        // the tile after this one: ask L2 for its pieces now, so that the refill issued at the end of this sweep is a
        // run of L2 hits (its descriptor arrived with this tile)
        if (wid == NW - 1 && tile + stride < nblk) {
            const TileDesc& dn = sm.sd[k == 2 ? 0 : k + 1];
            const int mp = lane / 3, arr = lane - mp * 3;
            if (mp < (dn.info & 0xff) && mp < APJ_MAX_PIECES) {
                const double2* src = arr == 0 ? st.XY[cur] : (arr == 1 ? st.CS[cur] : st.RR[gen]);
                apj_bulk_prefetch_l2(src + dn.pstart[mp], (unsigned)dn.plen[mp] * 16u);
            }
        }

        // ---- neighborInteractions ----
        PairAcc acc = {0.0, 0.0, 0.0, 0.0};
        const unsigned own = (unsigned)(d.own_slot + (active ? t : 0)) * 16u;
        const double2 me = lds_f64x2(sXY + own);
        double2 mcs, mrr;
        {
            const double Ri = lds_f64(sXY + dRR + own);
            const int nw = (nent + 1) >> 1;
            const uint4* gq = reinterpret_cast<const uint4*>(st.list32) + (long long)tile * st.max_quads * TB + t;
            uint4 q[QREG];
            q[1] = st.max_quads > 1 ? __ldg(gq + TB) : make_uint4(0u, 0u, 0u, 0u);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            q[0] = sQ0[t];
            if (wraps) sweep<TB, true>(acc, nw, q, gq, me, Ri, sXY, dCS, dRR, L, Lh, rn2);
            else sweep<TB, false>(acc, nw, q, gq, me, Ri, sXY, dCS, dRR, L, Lh, rn2);
            // own {cos,sin}, {R,1/R}: the last reads of the tile buffer
            mcs = lds_f64x2(sXY + dCS + own);
            mrr = lds_f64x2(sXY + dRR + own);
        }

        // ---- hand the tile buffer back: the last warp to get here refills it with the block's next tile ----
        const int nxt = tile + stride;
        {
            unsigned last = 0;
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                last = (atomicAdd(&sm.done, 1u) == (unsigned)NW - 1u) ? 1u : 0u;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                if (lane == 0) sm.done = 0u;
                if (nxt < nblk) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    const int k1 = k == 2 ? 0 : k + 1, k2 = k1 == 2 ? 0 : k1 + 1;
                    pipe_issue(st, &sm.sd[k1], cur, gen, sXYp, sCSp, sRRp, &sm.full,
                               nxt + stride < nblk ? st.tiles + nxt + stride : nullptr, &sm.sd[k2]);
                }
            }
        }
        // ---- own-particle inputs of the epilogue: from global memory (the tile buffer is being overwritten) ----
        double2 xo = make_double2(0.0, 0.0), xr = xo;
        if (active) { xo = st.XO[gen][g]; xr = st.XR[cur][g]; }
        // what the next tile needs from global memory is requested now, under the epilogue
        const int info = d.info;
        if (nxt < nblk) load_early(sm.sd[k == 2 ? 0 : k + 1], nxt);

        double sum_x = 0.0, sum_y = 0.0, top1 = 0.0;
        if (active) {
            // randuni() of this step (jamming.cpp:667)
            double u;
            if (INJECT) {
                u = noise_by_id[my_id];
            } else {
                const unsigned w = apj_philox_word0((unsigned)my_id, (unsigned)step, (unsigned)(step >> 32), 0u,
                                                    (unsigned)st.seed, (unsigned)(st.seed >> 32));
                u = apj_u32_to_randuni(w);
            }
            const double Ri = mrr.x;
            double Fx = acc.Fx, Fy = acc.Fy, ax = acc.ax, ay = acc.ay;
            if (!ctl->no_self_once) { ax += mcs.x; ay += mcs.y; }  // self term: Cell::update left x_new = cosp (Cell.h:102-103)
            {   // newSkinList: displacement since the last rebuild, COM drift removed
                const double ddx = apj_delta_norm(((me.x - xo.x) - ctl->COM[0]) + ctl->COM_old[0], L, Lh);
                const double ddy = apj_delta_norm(((me.y - xo.y) - ctl->COM[1]) + ctl->COM_old[1], L, Lh);
                top1 = apj_d2(ddx, ddy);
            }
            const double nz = ctl->CTnoise * u;
            const bool want_phi = always_full || step + 1 == ctl->target;
            double phi, sn, cs;
            apj_new_orientation(ax, ay, nz, want_phi, phi, cs, sn);
            double CF = ctl->CFself;
            if (ctl->ramp_len > 0) {                      // relax() ramp (jamming.cpp:518)
                const long long t_ = step - ctl->ramp_t0;
                if (t_ < ctl->ramp_len) CF = CF - (double)(ctl->ramp_len - t_) * CF / (double)ctl->ramp_len;
            }
            Fx += cs * CF * Ri;
            Fy += sn * CF * Ri;
            const double vx = Fx * mrr.y, vy = Fy * mrr.y;   // Rinv = 1/R stored at upload (jamming.cpp:298)
            const double dx = vx * st.dt, dy = vy * st.dt;
            double x = me.x + dx, y = me.y + dy;
            const double xrn = xr.x + dx, yrn = xr.y + dy;
            if (x >= Lh) x -= L; else if (x < -Lh) x += L;   // Cell::PBC, single wrap (Cell.h:168-175)
            if (y >= Lh) y -= L; else if (y < -Lh) y += L;
            st.XY[cur ^ 1][g] = make_double2(x, y);
            st.CS[cur ^ 1][g] = make_double2(cs, sn);
            st.XR[cur ^ 1][g] = make_double2(xrn, yrn);
            if (SLAB) {   // halo exchange fused into the epilogue (see the classic kernel)
                if (info & APJ_INFO_PUSH_LEFT) {
                    const long long pg = (long long)st.cap + st.gcap + g;
                    apj_peer(st, st.left, st.XY[cur ^ 1])[pg] = make_double2(x, y);
                    apj_peer(st, st.left, st.CS[cur ^ 1])[pg] = make_double2(cs, sn);
                }
                if (info & APJ_INFO_PUSH_RIGHT) {
                    const long long pg = (long long)st.cap + (g - ctl->last_col_start);
                    apj_peer(st, st.right, st.XY[cur ^ 1])[pg] = make_double2(x, y);
                    apj_peer(st, st.right, st.CS[cur ^ 1])[pg] = make_double2(cs, sn);
                }
            }
            if (want_phi) {   // fields only observables read
                st.V[gen][g] = make_double2(vx, vy);
                st.PHI[gen][g] = phi;
            }
            sum_x = xrn; sum_y = yrn;
        }
        // running reductions of the block: x_real per lane, top-2 displacement^2 per warp
        {
            const double2 a = sSum[t];
            sSum[t] = make_double2(a.x + sum_x, a.y + sum_y);
            double t1, t2;
            apj_warp_top2(top1, t1, t2);
            if (lane == 0) {
                double2 m = sm.top[wid];
                apj_top2_merge(m.x, m.y, t1, t2);
                sm.top[wid] = m;
                if (SLAB && (info & (APJ_INFO_PUSH_LEFT | APJ_INFO_PUSH_RIGHT))) sm.pushed = 1;
            }
        }
        if (nxt >= nblk) break;
        tile = nxt;
        k = k == 2 ? 0 : k + 1;
    }

    // ---- the block's partial: lanes in a shuffle tree, warps in order ----
    {
        double2 a = sSum[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
            a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
        }
        if (lane == 0) sm.red[wid] = make_double4(a.x, a.y, sm.top[wid].x, sm.top[wid].y);
        __syncthreads();
        if (wid != 0) return;
        double4 r = sm.red[0];
#pragma unroll
        for (int w = 1; w < NW; w++) {
            const double4 b = sm.red[w];
            r.x += b.x; r.y += b.y;
            apj_top2_merge(r.z, r.w, b.z, b.w);
        }
        if (SLAB && sm.pushed) { __syncwarp(); __threadfence_system(); }
        if (lane == 0) st.partials[blockIdx.x] = r;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Ring form of the step (one large periodic system, one lane per particle, split tail): ONE block per SM made of
// RING_NG = 3 groups of 256 threads around a ring of RING_NB = 5 tile buffers in shared memory (5 x 45 KB fit the
// 227 KB of an SM at the tile sizes of a homogeneous box; denser tiles fall back to the classic kernel).
//   * the block walks the tiles b, b + grid, b + 2 grid, ... (local numbers i = 0, 1, ...); group k sweeps the
//     tiles i = k, k + 3, ...; tile i lives in buffer i % 5;
//   * a buffer is refilled the moment its tile has been swept: the LAST warp of the group to finish the sweep of
//     tile i (shared-memory arrival counter, no barrier) issues the TMA copies of tile i + 5 into the same buffer.
//     The descriptor it needs arrived WITH tile i (a 64-byte bulk copy on the same mbarrier), so a refill is one trip
//     to L2 / HBM and has more than a tile time to land;
//   * what a tile needs from global memory besides the tile itself -- list length, id, the first list quad -- is
//     requested by the same thread right after the sweep of its previous tile and arrives under the epilogue, so a warp
//     goes from the last store of one tile straight into the sweep of the next: the head of the classic kernel
//     (descriptor trip, then tile + lists trip, ~40 % of a block's life) is gone;
//   * warps never wait for each other: one partial per WARP and tile (wpartials), folded in the classic kernel's
//     order by apj_reduce_commit_kernel -- positions AND the committed COM / displacement maxima are bit-identical
//     to the classic split-tail path;
//   * 24 warps share the register file instead of 32: 80 registers per thread.
template <int RING_NB>
struct RingSmem {
    TileDesc sd[RING_NB];            // descriptor of the tile in (or on its way into) buffer b
    TileDesc nd[RING_NB];            // descriptor of the NEXT occupant of buffer b: lands with the tile
    unsigned long long full[RING_NB];
    unsigned done[RING_NB];          // warps that finished sweeping the tile in buffer b
    int round[RING_NB];              // use number (i / 5) of the tile buffer b was last issued for: groups drift, and an mbarrier
                                     // parity wait two phases ahead of the barrier would pass at once (stale tile) -- a warp
                                     // first waits for "issued for MY round", then for the data
    int kcls;
};

// whole warp: `dw` = the 16 descriptor words of a tile (lanes 0..15). Publishes the descriptor next to buffer `b` and
// issues the TMA copies of the tile's pieces (+ the descriptor of the buffer's next occupant) on the buffer's mbarrier.
template <int RING_NB>
__device__ __forceinline__ void ring_issue(const DevState& st, RingSmem<RING_NB>& sm, unsigned char* smem_raw, const int b, const int round,
                                           const int dw, const TileDesc* next_src, const int cur, const int gen,
                                           const unsigned buf_bytes, const unsigned dCS) {
    const int lane = threadIdx.x & 31;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer was last read through the generic proxy
    if (lane < 16) reinterpret_cast<int*>(&sm.sd[b])[lane] = dw;
    const int np = __shfl_sync(0xffffffffu, dw, 3) & 0xff;
    const int mp = lane / 3, arr = lane - mp * 3;
    int off = 1, my_off = 1, my_len = 0, my_start = 0;
#pragma unroll
    for (int q = 0; q < APJ_MAX_PIECES; q++) {
        const int start = __shfl_sync(0xffffffffu, dw, 4 + q);
        const int len = __shfl_sync(0xffffffffu, dw, 4 + APJ_MAX_PIECES + q);
        if (q < np) {
            if (q == mp) { my_off = off; my_len = len; my_start = start; }
            off += len;
        }
    }
    __syncwarp();                                                   // descriptor stores before lane 0's release
    if (lane == 0) {
        apj_mbar_expect_tx(&sm.full[b], (unsigned)(off - 1) * 48u + (next_src ? (unsigned)sizeof(TileDesc) : 0u));
        __threadfence_block();
        asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(apj_smem_addr(&sm.round[b])), "r"(round) : "memory");   // the barrier is in phase `round` now
    }
    __syncwarp();
    if (mp < np && my_len > 0) {
        const double2* src = arr == 0 ? st.XY[cur] : (arr == 1 ? st.CS[cur] : st.RR[gen]);
        double2* dst = reinterpret_cast<double2*>(smem_raw + (size_t)b * buf_bytes + (size_t)arr * dCS);
        apj_bulk_g2s(dst + my_off, src + my_start, (unsigned)my_len * 16u, &sm.full[b]);
    }
    if (lane == 31 && next_src) apj_bulk_g2s(&sm.nd[b], next_src, (unsigned)sizeof(TileDesc), &sm.full[b]);
}

template <int TB, int RING_NG, int RING_NB, bool INJECT>
__global__ void __launch_bounds__(RING_NG * TB, 1)
apj_step_ring_kernel(const DevState st, const double* __restrict__ noise_by_id, const int always_full) {
    static_assert(RING_NG < RING_NB && RING_NG * TB == 768, "24 warps at 80 registers; at least one spare buffer");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(128) RingSmem<RING_NB> sm;
    SysCtl* __restrict__ ctl = st.ctl;
    const int nblk = ctl->nblk;
    const long long step = ctl->step;
    if ((int)blockIdx.x >= nblk || ctl->stale || step >= ctl->target) return;   // uniform over the grid
    const int cur = ctl->cur, gen = ctl->gen;
    const int stride = gridDim.x;
    const int ntile = (nblk - (int)blockIdx.x + stride - 1) / stride;           // tiles of this block
    const unsigned buf_bytes = (unsigned)(st.tile_cap + 1) * 48u;
    const unsigned dCS = (unsigned)(st.tile_cap + 1) * 16u, dRR = 2u * dCS;

    if (threadIdx.x == 0) {
        for (int b = 0; b < RING_NB; b++) {
            apj_mbar_init(&sm.full[b], 1);
            sm.done[b] = 0u;
            sm.round[b] = -1;
            *reinterpret_cast<double2*>(smem_raw + (size_t)b * buf_bytes) = make_double2(1e300, 1e300);   // sentinel slot 0
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        sm.kcls = apj_sweep_class(ctl, st);
    }
    __syncthreads();

    const int gi = threadIdx.x / TB;
    const int t = threadIdx.x - gi * TB, lane = t & 31, wid = t >> 5;
    if (gi >= ntile) return;
    const int kcls = sm.kcls;
    const double L = ctl->L, Lh = ctl->Lover2;
    const double rn2 = st.rn2;
    int tile = blockIdx.x + gi * stride;

    if (gi == 0 && wid == 0) {                                      // first fill of the ring: tiles 0 .. 4 of the block
        int d[RING_NB];
#pragma unroll
        for (int k = 0; k < RING_NB; k++)
            d[k] = (k < ntile && lane < 16) ? apj_ldg_l2keep(reinterpret_cast<const int*>(st.tiles + tile + k * stride) + lane) : 0;
#pragma unroll
        for (int k = 0; k < RING_NB; k++)
            if (k < ntile) ring_issue(st, sm, smem_raw, k, 0, d[k], k + RING_NB < ntile ? st.tiles + tile + (k + RING_NB) * stride : nullptr,
                                      cur, gen, buf_bytes, dCS);
    }

    int b = gi, rnd = 0;                                            // buffer of tile i = i % 5 and its use number i / 5
    int g0 = __ldg(&st.tiles[tile].g0), n = __ldg(&st.tiles[tile].n);
    const uint4* __restrict__ gq = reinterpret_cast<const uint4*>(st.list32) + (long long)tile * st.max_quads * TB + t;
    uint4 q[QREG];
    q[0] = __ldg(gq);
    unsigned cntk_raw = t < n ? st.cntk[g0 + t] : 0u;
    int id = st.ID[gen][g0 + (t < n ? t : 0)];

    for (int i = gi;; i += RING_NG) {
        const bool active = t < n;
        const long long g = (long long)g0 + (active ? t : 0);
        q[1] = st.max_quads > 1 ? __ldg(gq + TB) : make_uint4(0u, 0u, 0u, 0u);   // consumed a quad (~200 instructions) into the sweep

        unsigned sXY = apj_smem_addr(smem_raw + (size_t)b * buf_bytes);
        {
            int r_;
            do { asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(r_) : "r"(apj_smem_addr(&sm.round[b])) : "memory"); } while (r_ != rnd);
        }
        apj_mbar_wait(&sm.full[b], (unsigned)rnd & 1u, sXY);
        const bool wraps = (sm.sd[b].info & APJ_INFO_WRAPS) != 0;
        const unsigned own = (unsigned)(sm.sd[b].own_slot + (active ? t : 0)) * 16u;

        // ---- neighborInteractions ----
        PairAcc acc = {0.0, 0.0, 0.0, 0.0};
        const double2 me = lds_f64x2(sXY + own);
        double2 mcs, mrr;
        {
            const double Ri = lds_f64(sXY + dRR + own);
            const int nent = (int)((cntk_raw >> (8 * kcls)) & 0xffu);
            const int nw = (nent + 1) >> 1;
            if (wraps) sweep<TB, true>(acc, nw, q, gq, me, Ri, sXY, dCS, dRR, L, Lh, rn2);
            else sweep<TB, false>(acc, nw, q, gq, me, Ri, sXY, dCS, dRR, L, Lh, rn2);
            mcs = lds_f64x2(sXY + dCS + own);               // own {cos,sin}, {R,1/R}: the last reads of the buffer
            mrr = lds_f64x2(sXY + dRR + own);
        }

        // ---- hand the buffer back: the last warp of the group to get here refills it with tile i + 5 ----
        {
            unsigned last = 0;
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                last = (atomicAdd(&sm.done[b], 1u) == (unsigned)(TB / 32) - 1u) ? 1u : 0u;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                if (lane == 0) sm.done[b] = 0u;
                if (i + RING_NB < ntile) {
                    __threadfence_block();
                    const int dwn = lane < 16 ? reinterpret_cast<const int*>(&sm.nd[b])[lane] : 0;
                    __syncwarp();
                    ring_issue(st, sm, smem_raw, b, rnd + 1, dwn, i + 2 * RING_NB < ntile ? st.tiles + tile + 2 * RING_NB * stride : nullptr,
                               cur, gen, buf_bytes, dCS);
                }
            }
        }

        // ---- own-particle inputs of the epilogue, then what this thread's next tile needs: all in flight under the epilogue ----
        double2 xo = make_double2(0.0, 0.0), xr = xo;
        if (active) {
            asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(xo.x), "=d"(xo.y) : "l"(st.XO[gen] + g));
            asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(xr.x), "=d"(xr.y) : "l"(st.XR[cur] + g));
        }
        const bool more = i + RING_NG < ntile;
        const int tile_n = tile + RING_NG * stride;
        const int my_id = id;
        if (more) {
            g0 = __ldg(&st.tiles[tile_n].g0); n = __ldg(&st.tiles[tile_n].n);
            gq = reinterpret_cast<const uint4*>(st.list32) + (long long)tile_n * st.max_quads * TB + t;
            q[0] = __ldg(gq);
            cntk_raw = t < n ? st.cntk[g0 + t] : 0u;
            id = st.ID[gen][g0 + (t < n ? t : 0)];
        }

        double sum_x = 0.0, sum_y = 0.0, top1 = 0.0, top2 = 0.0;
        if (active) {
            double u;                                       // randuni() of this step (jamming.cpp:667)
            if (INJECT) {
                u = noise_by_id[my_id];
            } else {
                const unsigned w = apj_philox_word0((unsigned)my_id, (unsigned)step, (unsigned)(step >> 32), 0u,
                                                    (unsigned)st.seed, (unsigned)(st.seed >> 32));
                u = apj_u32_to_randuni(w);
            }
            const double Ri = mrr.x;
            double Fx = acc.Fx, Fy = acc.Fy, ax = acc.ax, ay = acc.ay;
            if (!ctl->no_self_once) { ax += mcs.x; ay += mcs.y; }  // self term: Cell::update left x_new = cosp (Cell.h:102-103)
            const double nz = ctl->CTnoise * u;
            const bool want_phi = always_full || step + 1 == ctl->target;
            double phi, sn, cs;
            apj_new_orientation(ax, ay, nz, want_phi, phi, cs, sn);
            double CF = ctl->CFself;
            if (ctl->ramp_len > 0) {                        // relax() ramp (jamming.cpp:518)
                const long long t_ = step - ctl->ramp_t0;
                if (t_ < ctl->ramp_len) CF = CF - (double)(ctl->ramp_len - t_) * CF / (double)ctl->ramp_len;
            }
            Fx += cs * CF * Ri;
            Fy += sn * CF * Ri;
            const double vx = Fx * mrr.y, vy = Fy * mrr.y;  // Rinv = 1/R stored at upload (jamming.cpp:298)
            const double dx = vx * st.dt, dy = vy * st.dt;
            double x = me.x + dx, y = me.y + dy;
            if (x >= Lh) x -= L; else if (x < -Lh) x += L;  // Cell::PBC, single wrap (Cell.h:168-175)
            if (y >= Lh) y -= L; else if (y < -Lh) y += L;
            st.XY[cur ^ 1][g] = make_double2(x, y);
            st.CS[cur ^ 1][g] = make_double2(cs, sn);
            {   // newSkinList: displacement since the last rebuild, COM drift removed (of the state the step STARTED from)
                const double ddx = apj_delta_norm(((me.x - xo.x) - ctl->COM[0]) + ctl->COM_old[0], L, Lh);
                const double ddy = apj_delta_norm(((me.y - xo.y) - ctl->COM[1]) + ctl->COM_old[1], L, Lh);
                top1 = apj_d2(ddx, ddy);
            }
            const double xrn = xr.x + dx, yrn = xr.y + dy;
            st.XR[cur ^ 1][g] = make_double2(xrn, yrn);
            if (want_phi) {                                 // fields only observables read
                st.V[gen][g] = make_double2(vx, vy);
                st.PHI[gen][g] = phi;
            }
            sum_x = xrn; sum_y = yrn;
        }

        // ---- this warp's partial of the tile (apj_reduce_commit_kernel folds the 8 of a tile in warp order) ----
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum_x += __shfl_xor_sync(0xffffffffu, sum_x, o);
            sum_y += __shfl_xor_sync(0xffffffffu, sum_y, o);
        }
        apj_warp_top2(top1, top1, top2);
        if (lane == 0) st.wpartials[(long long)tile * (TB / 32) + wid] = make_double4(sum_x, sum_y, top1, top2);
        if (!more) break;
        tile = tile_n;
        b += RING_NG;
        if (b >= RING_NB) { b -= RING_NB; rnd++; }
    }
}

// one warp: wait for the partials of all ranks, fold them in rank order, take the decision
__device__ __forceinline__ void apj_slab_commit_warp(const DevState& st, SysCtl* __restrict__ ctl) {
    const int lane = threadIdx.x & 31;
    const unsigned long long ep = ctl->seq[0];
    bool ok = true;
    if (lane < st.nranks && !ctl->slab_err) ok = apj_wait_flag(&st.mail->flag[0][lane], ep + 1, st.timeout_ns);
    const unsigned missing = __ballot_sync(0xffffffffu, !ok);
    if (lane != 0) return;
    ctl->seq[0] = ep + 1;
    if (missing || ctl->slab_err) {   // a peer never showed up: stop stepping, the host reports it
        if (missing && !ctl->slab_diag) ctl->slab_diag = (1 << 16) | (int)missing;
        ctl->slab_err |= 1;
        ctl->target = ctl->step;
        return;
    }
    double4 a = make_double4(0.0, 0.0, 0.0, 0.0);
    for (int r = 0; r < st.nranks; r++) {
        const double2* q = reinterpret_cast<const double2*>(&st.mail->part[ep & 1][r]);
        const double2 b01 = __ldcg(q), b23 = __ldcg(q + 1);
        a.x += b01.x; a.y += b01.y;
        apj_top2_merge(a.z, a.w, b23.x, b23.y);
    }
    apj_commit(ctl, st, a, apj_sweep_class(ctl, st));   // N of the global box; every rank folds the same partials
}
// Split tail of the step kernel (large systems): deterministic two-level fold of the per-block partials
// {sum x_real, sum y_real, top-1, top-2 d2}, then the commit (periodic box) or the push of this rank's
// partial to every rank (slab mode; apj_slab_commit_kernel follows). grid = (chunks, n_sys), RC_TB threads;
// CTA c folds partials [c*RC_CHUNK, (c+1)*RC_CHUNK): RC_PER consecutive ones per thread, a shuffle tree,
// the warps in order; the last CTA to finish folds the chunk results the same way.
constexpr int RC_TB = 256, RC_PER = 4, RC_CHUNK = RC_TB * RC_PER;
__device__ __forceinline__ double4 rc_block_fold(double4 a, double4* s_w) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
        const double b1 = __shfl_xor_sync(0xffffffffu, a.z, o), b2 = __shfl_xor_sync(0xffffffffu, a.w, o);
        apj_top2_merge(a.z, a.w, b1, b2);
    }
    if (lane == 0) s_w[wid] = a;
    __syncthreads();
    if (wid == 0) {
        a = s_w[0];
#pragma unroll
        for (int w = 1; w < RC_TB / 32; w++) {
            const double4 b = s_w[w];
            a.x += b.x; a.y += b.y;
            apj_top2_merge(a.z, a.w, b.z, b.w);
        }
    }
    return a;   // valid in warp 0
}
__device__ __forceinline__ double4 rc_load(const double4* p) {
    const double2 b01 = __ldcg(reinterpret_cast<const double2*>(p)), b23 = __ldcg(reinterpret_cast<const double2*>(p) + 1);
    return make_double4(b01.x, b01.y, b23.x, b23.y);
}
__global__ void __launch_bounds__(RC_TB) apj_reduce_commit_kernel(const DevState st) {
    __shared__ double4 s_w[RC_TB / 32];
    __shared__ unsigned s_last;
    const int sys = blockIdx.y, c = blockIdx.x, t = threadIdx.x;
    SysCtl* __restrict__ ctl = st.ctl + sys;
    if (ctl->stale || ctl->step >= ctl->target) return;   // the step kernel exited on the same test: nothing to fold
    const int nblk = st.persist_grid > 0 ? min(st.persist_grid, ctl->nblk) : ctl->nblk;   // pipelined kernel: one partial per resident block
    const int nchunk = (nblk + RC_CHUNK - 1) / RC_CHUNK;
    if (c >= nchunk) return;
    const double4* __restrict__ part = st.partials + (long long)sys * st.maxblk;
    double4* __restrict__ gpart = st.gpartials + (long long)sys * st.maxgrp;
    double4 a = make_double4(0.0, 0.0, 0.0, 0.0);
#pragma unroll
    for (int u = 0; u < RC_PER; u++) {
        const int k = c * RC_CHUNK + t * RC_PER + u;
        if (k < nblk) {
            double4 b;
            if (st.ring_nb > 0) {            // ring kernel: one partial per warp, folded here in the classic block's order (warp 0, then 1 .. 7)
                const int nwp = st.tb >> 5;
                const double4* wp = st.wpartials + (long long)k * nwp;
                b = rc_load(wp);
                for (int w = 1; w < nwp; w++) {
                    const double4 r = rc_load(wp + w);
                    b.x += r.x; b.y += r.y;
                    apj_top2_merge(b.z, b.w, r.z, r.w);
                }
            } else {
                b = rc_load(part + k);
            }
            a.x += b.x; a.y += b.y;
            apj_top2_merge(a.z, a.w, b.z, b.w);
        }
    }
    a = rc_block_fold(a, s_w);
    if (t == 0) {
        gpart[c] = a;
        apj_fence_gpu();
        s_last = (atomicAdd(&ctl->ticket, 1u) == (unsigned)nchunk - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    apj_fence_gpu();
    a = make_double4(0.0, 0.0, 0.0, 0.0);
    for (int k = t * RC_PER; k < nchunk; k += RC_TB * RC_PER) {   // nchunk <= RC_CHUNK for every supported size: one pass
#pragma unroll
        for (int u = 0; u < RC_PER; u++) {
            if (k + u < nchunk) {
                const double4 b = rc_load(gpart + k + u);
                a.x += b.x; a.y += b.y;
                apj_top2_merge(a.z, a.w, b.z, b.w);
            }
        }
    }
    __syncthreads();   // s_w is reused
    a = rc_block_fold(a, s_w);
    if (t >= 32) return;
    const int lane = t;
    if (st.slab) {
        if (lane == 0) ctl->ticket = 0u;
        const unsigned long long ep = ctl->seq[0];
        if (lane < st.nranks) {
            SlabMail* m = apj_peer(st, lane, st.mail);
            m->part[ep & 1][st.rank] = a;                              // same lane: the release store below orders it
            apj_st_release_sys(&m->flag[0][st.rank], ep + 1);
        }
        __syncwarp();
        apj_slab_commit_warp(st, ctl);   // second half of the step in the same launch: all-rank fold + decision
        return;
    }
    if (lane == 0) {
        ctl->ticket = 0u;
        apj_commit(ctl, st, a, apj_sweep_class(ctl, st));
    }
}

// Slab mode: second half of the step. Waits for the partials of all ranks (pushed by their step
// kernels, see above), folds them in rank order and takes the reference's decision
// (jamming.cpp:611): commit the speculative step or drop it and rebuild. One warp.
__global__ void apj_slab_commit_kernel(const DevState st) {
    SysCtl* __restrict__ ctl = st.ctl;
    if (ctl->stale || ctl->step >= ctl->target) return;   // the step kernel exited on the same test: nothing was sent
    apj_slab_commit_warp(st, ctl);
}

size_t step_smem_bytes(const DevState& st) {
    return (size_t)(st.tile_cap + 1) * 48 + (st.G > 1 ? (size_t)st.ppb * 32 : 0);
}
size_t pipe_smem_bytes(const DevState& st) {   // tile + per-thread running sums + per-thread first list quad
    return (size_t)(st.tile_cap + 1) * 48 + (size_t)st.tb * 32;
}

template <int TB, int G>
int configure(DevState& st) {
    const int bytes = 0;
    auto set = [&](auto k, int) { return apj_allow_max_smem(k); };
    if (!set(apj_step_kernel<TB, G, true, false, false>, bytes) || !set(apj_step_kernel<TB, G, false, false, false>, bytes) ||
        !set(apj_step_kernel<TB, G, true, true, false>, bytes) || !set(apj_step_kernel<TB, G, false, true, false>, bytes)) return -1;
    st.persist_grid = 0;
    if (G == 1) {
        if (!set(apj_step_kernel<TB, 1, true, false, true>, bytes) || !set(apj_step_kernel<TB, 1, false, false, true>, bytes) ||
            !set(apj_step_kernel<TB, 1, true, true, true>, bytes) || !set(apj_step_kernel<TB, 1, false, true, true>, bytes)) return -1;
        if (st.split_tail && st.n_sys == 1 && st.want_persist) {   // pipelined persistent kernel: one wave of resident blocks
            const int pb = (int)pipe_smem_bytes(st);
            if (!set(apj_step_pipe_kernel<TB, true, false>, pb) || !set(apj_step_pipe_kernel<TB, false, false>, pb) ||
                !set(apj_step_pipe_kernel<TB, true, true>, pb) || !set(apj_step_pipe_kernel<TB, false, true>, pb)) return -1;
            int dev = 0, sms = 0, nb = 0;
            if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, apj_step_pipe_kernel<TB, false, false>, TB, pb) != cudaSuccess || nb < 1) return -1;
            st.persist_grid = sms * nb;
            st.persist_sms = sms;
        }
        st.ring_nb = 0;
        if ((TB == 256 || TB == 128) && apj_ring_eligible(st)) {   // ring kernel: one block per SM around nb tile buffers
            const int nb = apj_ring_buffers(st.tb, st.tile_cap);
            if (nb > 0) {
                if (TB == 256 && (!set(apj_step_ring_kernel<256, 3, 5, true>, 0) || !set(apj_step_ring_kernel<256, 3, 5, false>, 0))) return -1;
                if (TB == 128 && (!set(apj_step_ring_kernel<128, 6, 8, true>, 0) || !set(apj_step_ring_kernel<128, 6, 8, false>, 0))) return -1;
                int dev = 0, sms = 0;
                if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
                st.ring_nb = nb;
                st.ring_grid = sms;
                st.persist_grid = 0;                              // (apj_reduce_commit_kernel folds one partial per TILE)
            }
        }
    }
    return 0;
}

template <int TB, int G, bool SPLIT>
void launch_variant(const DevState& st, cudaStream_t s, const double* noise_by_id, int always_full) {
    const int grid = st.n_sys * st.maxblk;
    const size_t smem = step_smem_bytes(st);
    if (st.slab) {
        if (noise_by_id) apj_step_kernel<TB, G, true, true, SPLIT><<<grid, TB, smem, s>>>(st, noise_by_id, always_full);
        else apj_step_kernel<TB, G, false, true, SPLIT><<<grid, TB, smem, s>>>(st, nullptr, always_full);
    } else {
        if (noise_by_id) apj_step_kernel<TB, G, true, false, SPLIT><<<grid, TB, smem, s>>>(st, noise_by_id, always_full);
        else apj_step_kernel<TB, G, false, false, SPLIT><<<grid, TB, smem, s>>>(st, nullptr, always_full);
    }
}
template <int TB>
void launch_pipe(const DevState& st, cudaStream_t s, const double* noise_by_id, int always_full) {
    const int grid = st.persist_grid;
    const size_t smem = pipe_smem_bytes(st);
    if (st.slab) {
        if (noise_by_id) apj_step_pipe_kernel<TB, true, true><<<grid, TB, smem, s>>>(st, noise_by_id, always_full);
        else apj_step_pipe_kernel<TB, false, true><<<grid, TB, smem, s>>>(st, nullptr, always_full);
    } else {
        if (noise_by_id) apj_step_pipe_kernel<TB, true, false><<<grid, TB, smem, s>>>(st, noise_by_id, always_full);
        else apj_step_pipe_kernel<TB, false, false><<<grid, TB, smem, s>>>(st, nullptr, always_full);
    }
}

inline void launch_ring(const DevState& st, cudaStream_t s, const double* noise_by_id, int always_full) {
    const size_t smem = (size_t)st.ring_nb * (st.tile_cap + 1) * 48;
    if (st.tb == 256) {
        if (noise_by_id) apj_step_ring_kernel<256, 3, 5, true><<<st.ring_grid, 768, smem, s>>>(st, noise_by_id, always_full);
        else apj_step_ring_kernel<256, 3, 5, false><<<st.ring_grid, 768, smem, s>>>(st, nullptr, always_full);
    } else {
        if (noise_by_id) apj_step_ring_kernel<128, 6, 8, true><<<st.ring_grid, 768, smem, s>>>(st, noise_by_id, always_full);
        else apj_step_ring_kernel<128, 6, 8, false><<<st.ring_grid, 768, smem, s>>>(st, nullptr, always_full);
    }
}

template <int TB, int G>
void launch(const DevState& st, cudaStream_t s, const double* noise_by_id, int always_full) {
    if (G == 1 && st.split_tail) {
        if (st.ring_nb > 0) { launch_ring(st, s, noise_by_id, always_full); apj_check_launch("apj_step_ring_kernel"); }
        else if (st.persist_grid > 0) { launch_pipe<TB>(st, s, noise_by_id, always_full); apj_check_launch("apj_step_pipe_kernel"); }
        else { launch_variant<TB, 1, true>(st, s, noise_by_id, always_full); apj_check_launch("apj_step_kernel (split tail)"); }
        apj_reduce_commit_kernel<<<dim3((st.maxblk + RC_CHUNK - 1) / RC_CHUNK, st.n_sys), RC_TB, 0, s>>>(st);
        apj_check_launch("apj_reduce_commit_kernel");
    } else {
        launch_variant<TB, G, false>(st, s, noise_by_id, always_full);
        apj_check_launch("apj_step_kernel (fused tail)");
    }
    if (st.slab && !(G == 1 && st.split_tail)) { apj_slab_commit_kernel<<<1, 32, 0, s>>>(st); apj_check_launch("apj_slab_commit_kernel"); }
}

}  // namespace

#if APJ_TB_G1 == 512
#define APJ_DISPATCH_512(CALL) if (st.tb == 512 && st.G == 1) { CALL(512, 1); } else
#else
#define APJ_DISPATCH_512(CALL)
#endif
#define APJ_DISPATCH(CALL)                                                   \
    APJ_DISPATCH_512(CALL)                                                   \
    if (st.tb == 256 && st.G == 1) { CALL(256, 1); }                         \
    else if (st.tb == 192 && st.G == 1) { CALL(192, 1); }                    \
    else if (st.tb == 128 && st.G == 1) { CALL(128, 1); }                    \
    else if (st.tb == 128 && st.G == 2) { CALL(128, 2); }                    \
    else if (st.tb == 128 && st.G == 4) { CALL(128, 4); }                    \
    else if (st.tb == 128 && st.G == 8) { CALL(128, 8); }

int apj_step_blocks_per_sm_limit(int tb) { return apj_blocks_for(tb); }
size_t apj_step_extra_smem(const DevState& st) { return (st.want_persist && st.split_tail && st.n_sys == 1 && st.G == 1) ? (size_t)st.tb * 32 : 0; }

int apj_configure_kernels(DevState& st) {
#define APJ_CFG(TB, G) return configure<TB, G>(st)
    APJ_DISPATCH(APJ_CFG)
#undef APJ_CFG
    return -1;
}

// The kernels of ONE step launched one by one with `between(k)` called after part k (0: step kernel, 1: fold + commit,
// 2: slab commit) -- apj_time_step_parts records events there. Same kernels, same order as apj_launch_step.
template <int TB, int G>
static void launch_parts(const DevState& st, cudaStream_t s, void (*between)(int, void*), void* arg) {
    if (G == 1 && st.split_tail) {
        if (st.ring_nb > 0) launch_ring(st, s, nullptr, 0);
        else if (st.persist_grid > 0) launch_pipe<TB>(st, s, nullptr, 0);
        else launch_variant<TB, 1, true>(st, s, nullptr, 0);
        between(0, arg);
        apj_reduce_commit_kernel<<<dim3((st.maxblk + RC_CHUNK - 1) / RC_CHUNK, st.n_sys), RC_TB, 0, s>>>(st);
        between(1, arg);
    } else {
        launch_variant<TB, G, false>(st, s, nullptr, 0);
        between(0, arg);
        between(1, arg);
    }
    if (st.slab && !(G == 1 && st.split_tail)) apj_slab_commit_kernel<<<1, 32, 0, s>>>(st);
    between(2, arg);
}
void apj_launch_step_parts(const DevState& st, const ApjLaunch& l, void (*between)(int, void*), void* arg) {
#define APJ_PARTS(TB, G) launch_parts<TB, G>(st, l.stream, between, arg)
    APJ_DISPATCH(APJ_PARTS)
#undef APJ_PARTS
    if (l.launch_counter) (*l.launch_counter) += (st.G == 1 && st.split_tail) ? 2 : (st.slab ? 2 : 1);
}

void apj_launch_step(const DevState& st, const ApjLaunch& l, const double* noise_by_id, int always_full) {
#define APJ_LAUNCH(TB, G) launch<TB, G>(st, l.stream, noise_by_id, always_full)
    APJ_DISPATCH(APJ_LAUNCH)
#undef APJ_LAUNCH
    if (l.launch_counter) (*l.launch_counter) += (st.G == 1 && st.split_tail) ? 2 : (st.slab ? 2 : 1);
}
