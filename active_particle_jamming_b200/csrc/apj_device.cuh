// Device-side data model shared by all kernels of the B200 hot path.
//
// HBM layout (all fp64, SoA, particles kept in CELL ORDER -- re-sorted at every Verlet rebuild;
// cells are numbered cy + cx*b so one column of cells is one contiguous run of particles):
//   XY[2]   {x, y}                  16 B, ping-pong by ctl.cur: a step reads half `cur` and
//   CS[2]   {cos(phi), sin(phi)}    16 B  writes half `cur^1`
//   XR[2]   x_real (unwrapped position), ping-pong with XY so a speculative step can be dropped
//   RR, X0, XO, V, PHI, ID, BOX [2]  fields that only move at a rebuild; ping-pong by ctl.gen
//           (RR = {R, 1/R}: Cell::R and Cell::Rinv, jamming.cpp:297-298)
//   tiles   one TileDesc per work block: a run of <= ppb = tb/G consecutive particles of ONE cell
//           column plus the <= 6 contiguous particle runs ("pieces") that hold every cell
//           adjacent to the block's cells. The step kernel copies the pieces into shared
//           memory with TMA bulk copies; all neighbour gathers then hit shared memory.
//   list    full Verlet list as TILE-LOCAL 16-bit entries (byte offset of the neighbour's 16-byte
//           tile slot; slot 0 = sentinel, used to pad odd lists), two per 32-bit word, sorted by
//           distance at build time. Word w of particle p belongs to lane sub = w % G of the G
//           lanes sweeping p, as that lane's word k = w / G; a thread's words are stored as uint4
//           quads in the block's [quad][tb threads] array: one coalesced 16-byte load per thread
//           fetches 8 entries, issued before the tile has landed
// One SysCtl per batched system (replica) holds geometry, parities, COM and the step counter;
// kernels read it at entry, the last block of the step kernel commits to it.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// Truncated literals are part of the reference algorithm (jamming.cpp:3-4).
#define APJ_PI 3.14159265
#define APJ_PI2 6.28318531

#ifndef APJ_TB_G1
#define APJ_TB_G1 256       // threads per work block for one lane per particle (512: experiment)
#endif
#define APJ_TB_MAX (APJ_TB_G1 > 256 ? APJ_TB_G1 : 256)   // largest thread block of the step kernel (particles per block = tb / G)
#define APJ_MAX_PIECES 6
#define APJ_MAX_RANKS 8     // slab mode: GPUs of one box
// Skin-aware sweep length. A list entry whose distance AT BUILD TIME was d_b can only come within
// rn once the two particles have moved d_b - rn relative to each other, and the relative
// displacement of any pair since the build is bounded by the sum of the two largest COM-corrected
// displacements -- the very quantity D = sqrt(l1)+sqrt(l2) the reference's skin test evaluates every
// step (jamming.cpp:596-611). Lists are stored in APJ_CLASSES classes of build distance, uniform in d_b^2
// (class k = floor((d_b^2 - rn^2) * APJ_CLASSES / (rs^2 - rn^2)), class 0 also holds everything inside rn), so a
// step only has to sweep the classes k <= ((rn + D)^2 - rn^2) * APJ_CLASSES / (rs^2 - rn^2).
// The class is chosen from the previous step's D plus a margin and VERIFIED against the step's own D
// at commit time; a launch that swept too few entries is dropped and re-run (same mechanism as the
// speculative rebuild), so the result never differs from sweeping the full list.
#define APJ_CLASSES 4
// TileDesc::info bits
#define APJ_INFO_WRAPS (1 << 8)
#define APJ_INFO_PUSH_LEFT (1 << 9)    // slab mode: the block's particles are the left neighbour's right ghost column
#define APJ_INFO_PUSH_RIGHT (1 << 10)  // ... the right neighbour's left ghost column

struct __align__(64) TileDesc {
    int g0;         // first particle (absolute index) of the block
    int n;          // particles in the block (1..ppb)
    int own_slot;   // tile slot (1-based) of particle g0
    int info;       // npieces | APJ_INFO_* flags | list words of the longest list << 16
    int pstart[APJ_MAX_PIECES];  // absolute particle index where each piece starts
    int plen[APJ_MAX_PIECES];    // particles in each piece (tile slots are the concatenation)
};

struct __align__(128) SysCtl {
    // geometry (Engine::L, Lover2, lp, b, nbox; jamming.cpp:99-103)
    double L, Lover2, lp;
    int b, nbox;
    int cell_base;  // offset of this system's cells in the global cell arrays (nbox+1 slots each)
    int col_base;   // offset of this system's columns in the per-column scratch (b+1 slots each)
    // activity (Engine::CFself / CTnoise) and the relax() ramp (jamming.cpp:516-520)
    double CFself, CTnoise;
    long long ramp_len, ramp_t0;
    // dynamic state
    int cur;         // parity of XY / CS / XR
    int gen;         // parity of the rebuild-permuted arrays
    int stale;       // 1: lists must be rebuilt before the next step can run
    int save_old;    // 1: the pending rebuild was fired by the skin test (saveOldPositions + resetCounter++)
    int no_self_once;  // drop the alignment self term for the next committed step
    int list_max, overflow;
    unsigned ticket;
    int nblk;        // work blocks in use (<= DevState::maxblk)
    int tile_max;    // largest tile (slots) of the current decomposition
    long long step, target;
    long long reset_counter, n_rebuilds, n_discarded;
    double COM[2], COM_old[2], COM0[2];
    // particle range of this system in the arrays: [p0, p0 + n_own). Fixed (p0 = sys*N, n_own = N) for
    // periodic systems; in slab mode n_own changes at every rebuild (migration) and p0 = 0.
    int p0, n_own;
    int n_end;       // scratch of the cell scan: p0 + particles binned into owned cells
    int col0, ncols; // slab mode: owned cell columns [col0, col0 + ncols) of the global b x b grid (periodic: 0, b)
    int last_col_start;  // first particle of the last owned column (index of the right neighbour's ghost copy)
    int blk_shift;       // slab mode: work block the grid starts with (first block of the LAST owned column), so that the
                         // blocks that push halo data to the neighbours run first and their system-scope fences are long
                         // done when the launch drains; 0 otherwise
    int slab_err;    // sticky: 1 peer wait timed out, 2 migrant crossed more than one slab, 4 capacity exceeded
    // Skin-aware sweep length (see APJ_CLASSES below). skinD = sqrt(l1)+sqrt(l2) of newSkinList
    // (jamming.cpp:611) for the state the last committed step started from (0 right after a
    // skin-triggered rebuild); skinDD = its last per-step increase.
    int trunc_ok;    // 1: x_old is the state the lists were built from (last rebuild was fired by the skin test)
    int kmin;        // lowest distance class the next launch must sweep (set when a launch swept too few; 0 after a commit)
    int slab_diag;   // first timed-out wait: (channel + 1) << 16 | mask of the ranks that had not arrived
    double skinD, skinDD;
    // x_old can be reset WITHOUT a list build (start() :203 after relax: apj_mark_origin). The lists then are older than
    // x_old; skinBase bounds the pair displacement between the list build and that reset (twice the largest COM-
    // corrected displacement at the moment of the reset, accumulated over resets), so skinBase + D bounds it since the build.
    double skinBase;
    unsigned long long mark_d2;   // scratch of apj_mark_origin: bit pattern of the largest displacement^2 (atomicMax)
    long long n_retried;   // launches dropped because the sweep length chosen from the previous step's skinD was too short
    long long n_class[4];  // committed steps per swept class (APJ_CLASSES entries)
    unsigned long long seq[3];   // slab mode sequence numbers: step epochs, rebuild phase A, rebuild phase B
};

// ---- slab mode (one global periodic box cut into slabs of whole cell columns, one rank per GPU) ----
// Everything a PEER writes lives in one device allocation per rank (the "arena") with the same layout
// on every rank, so the address of any peer-visible object of rank r is
//   peer_arena[r] + (address on this rank - arena).
// Peers are other GPUs of the box reached over NVLink (cudaIpc mappings, one process per GPU) or, for
// tests, other handles on the same device.
struct __align__(16) MigRec {   // one migrating particle: every per-particle field (128 B)
    double2 xy, cs, xr, rr, x0, xo, v;
    double phi;
    int id, pad;
};
struct __align__(16) SlabMail {
    unsigned long long flag[3][APJ_MAX_RANKS];   // [channel][source rank]: that rank's sequence number on the channel
    double4 part[2][APJ_MAX_RANKS];              // [epoch parity][source rank]: {sum x_real, sum y_real, top1 d2, top2 d2}
    int inbox_count;                             // migrants pushed into `inbox` since the last rebuild
    int ghost_n[2][2];                           // [gen parity][side] particles in the ghost column
    int pad[3];
};

struct DevState {
    int n_sys, N;          // systems, particles per system (slab mode: particles of the GLOBAL box)
    int cap;               // particle slots per system in the arrays (periodic: N; slab mode: capacity of this rank)
    long long ntot;        // n_sys * cap
    int G;                 // lanes per particle in the sweep (1, 2, 4 or 8)
    int tb;                // threads per work block of the step kernel (128 or 256)
    int ppb;               // particles per work block = tb / G
    int maxblk;            // work blocks reserved per system (N/ppb + b + 1)
    int S;                 // list capacity per particle (even)
    int max_rounds;        // ceil(S/2 / G): list words one lane can hold
    int max_quads;         // ceil(max_rounds / 4): uint4 rows of the per-block list array
    int tile_cap;          // shared-memory tile capacity in slots (<= 4094; slot 0 is the sentinel)
    double dt, rn2, rs2, skin;  // skin = rs - rn (jamming.cpp:611)
    double rn, cls_inv;         // cls_inv = APJ_CLASSES / (rs^2 - rn^2): build-distance class of a list entry = floor((d2 - rn2) * cls_inv), >= 0
    int truncate;               // 0 disables the skin-aware sweep length (always the full list)
    int split_tail;             // 1: step kernel leaves per-block partials, apj_reduce_commit_kernel folds and commits (large systems)
    int want_persist;           // APJ_FLAG_PERSIST
    int persist_grid;           // > 0: blocks of the persistent step kernel (one resident wave; one large system only)
    int persist_sms;            // SMs of the device (block b of the persistent grid sits in resident slot b / persist_sms)
    int want_ring;              // APJ_STEP_RING=1: ring form of the step kernel (one block per SM, consumer groups around a ring of tile buffers)
    int ring_nb, ring_grid;     // > 0: tile buffers of a ring block / blocks (one per SM); set by apj_configure_kernels
    unsigned long long seed;
    SysCtl* ctl;
    double2* XY[2];
    double2* CS[2];
    double2* XR[2];
    double2* RR[2];
    double2* X0[2];
    double2* XO[2];
    double2* V[2];
    double* PHI[2];
    int* ID[2];
    int* BOX[2];  // internal cell index cy + cx*b (columns contiguous)
    TileDesc* tiles;      // n_sys * maxblk
    unsigned* list32;     // n_sys * maxblk * max_quads * tb uint4
    int* cnt;             // per particle: list length
    unsigned* cntk;       // per particle: cumulative list length through class k in byte k (byte APJ_CLASSES-1 == cnt)
    int* boxnew;
    int* perm;
    int* cell_count;   // zero between rebuilds
    int* cell_start;   // per system nbox+1 entries, absolute particle indices
    int* cell_cursor;
    int* col_blk;      // per system b+1 entries: first work block of each column
    int* chunk_sums;   // per system scan_chunks entries: scratch of the cell scan
    double4* partials;  // per work block {sum x_real, sum y_real, top1 d2, top2 d2}
    double4* gpartials; // per group of 32 work blocks
    double4* wpartials; // ring kernel: per work block and warp (8 per block); null unless the ring kernel was asked for
    unsigned* gticket;  // per group arrival counter (zero between launches)
    int maxgrp;         // groups reserved per system
    // slab mode (n_sys == 1). Array slots [cap, cap+gcap) hold the left ghost column (global column
    // col0-1), [cap+gcap, cap+2*gcap) the right one (col0+ncols); XY, CS, RR, ID live in the arena.
    int slab, rank, nranks, left, right;
    int gcap, mcap;        // capacity of one ghost column / of the migration inbox
    unsigned long long timeout_ns;   // a peer that does not show up within this time is reported, not waited for
    char* arena;
    char* peer_arena[APJ_MAX_RANKS];   // peer_arena[rank] == arena
    SlabMail* mail;        // in the arena
    MigRec* inbox;         // in the arena, mcap records
    int* gstart;           // in the arena: [gen parity][side][b+1] row starts of the ghost columns (absolute slots)
};

template <class T>
__device__ __forceinline__ T* apj_peer(const DevState& st, int r, T* mine) {
    return reinterpret_cast<T*>(st.peer_arena[r] + (reinterpret_cast<char*>(mine) - st.arena));
}
__device__ __forceinline__ unsigned long long apj_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void apj_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long apj_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *p >= want; false on timeout
__device__ __forceinline__ bool apj_wait_flag(const unsigned long long* p, unsigned long long want, unsigned long long timeout_ns) {
    if (apj_ld_acquire_sys(p) >= want) return true;
    const unsigned long long t0 = apj_globaltimer();
    while (apj_ld_acquire_sys(p) < want) {
        if (apj_globaltimer() - t0 > timeout_ns) return false;
        __nanosleep(200);
    }
    return true;
}

// Engine::delta_norm (jamming.cpp:872-880) for |delta| < 1.5 L: one conditional add of -L or +L
// gives the same value as the reference's while loop (k*L with k = +-1 is exact).
__device__ __forceinline__ double apj_wrap1(double d, double L, double Lh) {
    if (d < -Lh) d += L;
    else if (d >= Lh) d -= L;
    return d;
}
// General form (used where the argument is not bounded by the box).
__device__ __forceinline__ double apj_delta_norm(double d, double L, double Lh) {
    if (d < -Lh) { do { d += L; } while (d < -Lh || d >= Lh); }
    else if (d >= Lh) { do { d -= L; } while (d < -Lh || d >= Lh); }
    return d;
}
// dx*dx + dy*dy with the reference's rounding (two products, one sum; no FMA contraction), so
// that distance predicates are bit-identical to the CPU build (SURVEY §7 "hard parts").
__device__ __forceinline__ double apj_d2(double dx, double dy) {
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// Philox4x32-10 (Salmon et al. SC'11). counter = (id, step_lo, step_hi, system), key = seed.
__device__ __forceinline__ unsigned apj_philox_word0(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                                     unsigned k0, unsigned k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}
// all four output words (device-side initCells draws several variates per particle)
__device__ __forceinline__ uint4 apj_philox4(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// boost::uniform_real<>(-PI, PI) over a 32-bit engine (Boost 1.64 generate_uniform_real):
// u / 2^32 * (PI - (-PI)) + (-PI)   
__device__ __forceinline__ double apj_u32_to_randuni(unsigned u) {
    return __dadd_rn(__dmul_rn(__dmul_rn((double)u, 1.0 / 4294967296.0), APJ_PI - (-APJ_PI)), -APJ_PI);
}

// build-distance class of a list entry (apj_verlet_build_kernel): a pure function of d2, so the list order does
// not depend on the decomposition
__device__ __forceinline__ int apj_entry_class(double d2, double rn2, double cls_inv) {
#ifdef APJ_NO_CLASSES   // experiment: pure slot order (more lanes of a warp on the same slot), no shortened sweeps
    return 0;
#endif
    return min(max((int)((d2 - rn2) * cls_inv), 0), APJ_CLASSES - 1);
}
// last distance class a step must sweep when the skin-test value is D (APJ_CLASSES-1 = full list): an entry of
// class k was built at d_b^2 >= rn^2 + k / cls_inv and can only have come within rn if d_b <= rn + D
__device__ __forceinline__ int apj_class_for(double D, const DevState& st) {
    const double r = st.rn + D;
    const double x = (r * r - st.rn2) * st.cls_inv + 1e-6;
    return x >= (double)(APJ_CLASSES - 1) ? APJ_CLASSES - 1 : (int)x;
}
// class the launch sweeps: from the previous step's D plus a margin for this step's motion (uniform
// over the blocks of a system: reads only fields no block writes before the commit)
__device__ __forceinline__ int apj_sweep_class(const SysCtl* ctl, const DevState& st) {
    if (!st.truncate || !ctl->trunc_ok) return APJ_CLASSES - 1;
    const int k = apj_class_for(ctl->skinBase + ctl->skinD + 2.0 * ctl->skinDD + 0.01, st);
    return k > ctl->kmin ? k : ctl->kmin;
}

// top-2 merge of two (largest, second) pairs -- the multiset the reference's sequential scan
// keeps (jamming.cpp:607-608).
__device__ __forceinline__ void apj_top2_merge(double& a1, double& a2, double b1, double b2) {
    if (b1 > a1) { a2 = fmax(a1, b2); a1 = b1; }
    else { a2 = fmax(a2, b1); }
}

// tile slot (1-based; 0 is the sentinel) of absolute particle index j (j must lie in one of the pieces)
__device__ __forceinline__ int apj_slot_of(const TileDesc& d, int j) {
    int off = 1;
    const int npieces = d.info & 0xff;
#pragma unroll
    for (int p = 0; p < APJ_MAX_PIECES; p++) {
        if (p < npieces) {
            const int r = j - d.pstart[p];
            if (r >= 0 && r < d.plen[p]) return off + r;
            off += d.plen[p];
        }
    }
    return -1;
}

// ---- TMA bulk copy (global -> shared) + mbarrier, sm_90+/sm_100a PTX ----------------------
__device__ __forceinline__ unsigned apj_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void apj_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(apj_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void apj_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(apj_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void apj_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(apj_smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(apj_smem_addr(bar)) : "memory");
}
// shared-state-space loads by 32-bit address (keeps gather loops free of generic->shared address arithmetic)
__device__ __forceinline__ double2 apj_lds_f64x2(unsigned addr) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
// 32-bit load that asks L2 to keep the line (evict_last policy): the tile descriptors are re-read every step while
// ~3 GB of particle data stream through the 126 MB L2 in between; 4 MB of descriptors stay resident this way and
// the head of a step block saves one trip to HBM.
__device__ __forceinline__ int apj_ldg_l2keep(const int* p) {
#ifdef APJ_NO_L2KEEP
    return __ldg(p);
#else
    int v;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
#endif
}
// L2 prefetch hints (no data reaches the SM): one line / a contiguous range (16-byte granules)
__device__ __forceinline__ void apj_prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void apj_bulk_prefetch_l2(const void* p, unsigned bytes) {
    const unsigned long long a = reinterpret_cast<unsigned long long>(p);
    const unsigned long long a0 = a & ~15ull;
    const unsigned n = (unsigned)((a + bytes + 15ull - a0) & ~15ull);
    if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(n) : "memory");
}
// `dep` is threaded through the wait as an in/out operand: shared-memory loads whose address is
// derived from it cannot be scheduled before the barrier completes.
__device__ __forceinline__ void apj_mbar_wait(unsigned long long* bar, unsigned parity, unsigned& dep) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%2], %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done), "+r"(dep) : "r"(apj_smem_addr(bar)), "r"(parity) : "memory");
    }
}

// Staging planes of the host <-> device state hand-over (apj_state_io.cu): the 14 fp64 fields of apj_state
// in its order (x, y, x_real, y_real, x0, y0, x_old, y_old, R, phi, cosp, sinp, vx, vy), box ids, particle
// ids (slab upload) and an error flag (1: a radius is not positive, 2: a particle id is out of range).
struct ApjStage {
    double* f[14];
    int* box;
    int* ids;
    int* flag;
};

// APJ_DEBUG_LAUNCH=1: report the first kernel launch that fails, by name (launch errors are otherwise picked up by
// the next cudaGetLastError of the calling ABI function)
#include <cstdio>
#include <cstdlib>
inline void apj_check_launch(const char* what) {
    static const int on = getenv("APJ_DEBUG_LAUNCH") ? atoi(getenv("APJ_DEBUG_LAUNCH")) : 0;
    if (!on) return;
    const cudaError_t s = cudaPeekAtLastError();
    if (s != cudaSuccess) fprintf(stderr, "apj_b200: launch of %s failed: %s\n", what, cudaGetErrorString(s));
}

// Opt a kernel in to the device's full dynamic shared memory, once. The attribute is per FUNCTION, not per handle:
// setting it to what one handle needs would undercut another handle of the same process with a larger tile.
template <class K>
inline bool apj_allow_max_smem(K kernel) {
    int dev = 0, optin = 0;
    cudaFuncAttributes fa;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return false;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return false;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes) == cudaSuccess;
}

// host-side launchers (one per .cu)
struct ApjLaunch { cudaStream_t stream; long long* launch_counter; };
void apj_launch_pack(const DevState& st, cudaStream_t s, const ApjStage& sg, unsigned present, long long n, int slab);
void apj_launch_unpack(const DevState& st, cudaStream_t s, const ApjStage& sg, unsigned want, int by_id);
void apj_launch_checksum(const DevState& st, cudaStream_t s, unsigned long long* out);
// device-side initCells / topology / overlap hue (apj_setup.cu)
void apj_launch_radii_sums(cudaStream_t s, long long n, int n_sys, unsigned long long seed, double* d_part, double* d_sum);
void apj_launch_init_lattice(const DevState& st, cudaStream_t s, unsigned long long seed);
void apj_launch_box_table(const DevState& st, cudaStream_t s, int sys, double* d_centres, int* d_neighbors);
void apj_launch_overlap_hue(const DevState& st, cudaStream_t s, int* d_over_by_id);
void apj_launch_step(const DevState& st, const ApjLaunch& l, const double* noise_by_id, int always_full);
void apj_launch_step_parts(const DevState& st, const ApjLaunch& l, void (*between)(int, void*), void* arg);
void apj_launch_rebuild_chain(const DevState& st, const ApjLaunch& l, int max_nbox, int max_b);
int apj_rebuild_chain_launches(const DevState& st);
int apj_configure_kernels(DevState& st);   // also sets st.persist_grid
int apj_configure_rebuild(const DevState& st);
int apj_step_blocks_per_sm_limit(int tb);   // __launch_bounds__ of the step kernel
// ring kernel: large periodic single system, one lane per particle, 256-thread tiles, split tail
inline bool apj_ring_eligible(const DevState& st) { return st.want_ring && st.split_tail && st.n_sys == 1 && st.G == 1 && (st.tb == 256 || st.tb == 128) && !st.slab; }
// tile buffers of a ring block (5 for 256-particle tiles, 8 for 128-particle tiles) if they fit the 227 KB of a block next to the
// kernel's static shared memory, else 0 (classic kernel)
inline int apj_ring_buffers(int tb, int tile_cap) { const int nb = tb == 256 ? 5 : 8; return (size_t)nb * (tile_cap + 1) * 48 <= (size_t)(232448 - 2560) ? nb : 0; }
size_t apj_step_extra_smem(const DevState& st);   // dynamic shared memory of a step block beyond the tile
int apj_max_list_capacity();
int apj_scan_chunk_cells();
