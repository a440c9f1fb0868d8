// Rebuild chain: assignCellsToGrid + buildVerletLists (+ saveOldPositions when the skin test
// fired) of the reference (code/jam/jamming.cpp:527-548, :550-585, :825-835), re-designed as a
// counting sort that physically re-orders every particle array into cell order:
//
//   bin_count    exact nearest-box-centre binning + atomic histogram of cells
//   cell_scan    exclusive prefix scan of the histogram -> cell_start / cell_cursor (chunk sums,
//                scan of the chunk sums, in-chunk scan: three launches, any number of cells)
//   scatter      slot = atomicAdd(cursor[cell]) -> perm
//   cell_sort    sort each cell's slots by particle id (== Box::CellList ascending order;
//                makes the layout, and with it every reduction order, deterministic)
//   reorder      gather all SoA arrays through perm into the other ping-pong half
//   make_tiles   cut every cell column into work blocks of <= ppb particles (scan of the blocks per column),
//   fill_tiles   then record, one thread per block, the contiguous particle runs that cover all adjacent
//                cells (TileDesc)
//   verlet_build 3x3 cell sweep out of a shared-memory tile, d2 < rs2, full list stored as tile slots,
//                first coordination shell first
//   finish       flip parities, COM_old = COM, resetCounter++, clear `stale`
//
// Every kernel exits at once for systems whose ctl.stale is 0, so the chain can sit in the
// step graph unconditionally and costs a few empty launches when no rebuild is pending.
#include <algorithm>
#include "apj_device.cuh"

namespace {

constexpr int RB_BLOCK = 128;
constexpr int SCAN_BLOCK = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_CHUNK = SCAN_BLOCK * SCAN_ITEMS;   // cells scanned by one block
constexpr int MAX_S = 96;  // compile-time bound of DevState::S

// ---- assignCellsToGrid (jamming.cpp:527-548) ---------------------------------------------
// The reference scans ALL boxes for the nearest centre with strict '<' starting from
// r2 = lp*lp*0.25*NDIM; only the 3x3 boxes around the floor-binned one can win, and they are
// visited here in the same ascending box order (i + j*b), so ties resolve identically. If no
// box wins (particle exactly on a 4-corner) the particle keeps its previous box, as there.
__device__ __forceinline__ double box_centre(int i, double L, double Lh, int b) {
    const double mn = -Lh + i * L / b;         // grid[p].min (jamming.cpp:379-380)
    const double mx = -Lh + (i + 1) * L / b;   // grid[p].max (:381-382)
    return (mn + mx) / 2.;                     // :385
}

// The per-particle kernels of the chain run a fixed grid (bps blocks per system, grid-stride over
// the system's particles): an idle chain -- the common case, it sits in every step group -- then
// costs a few microseconds instead of retiring N / 128 empty blocks per kernel.
// returns the winning box as (column, row) of the GLOBAL grid; prev_x < 0: no previous box
__device__ __forceinline__ int2 nearest_box(const double2 me, const double L, const double Lh, const double lp, const int b,
                                            int prev_x, int prev_y) {
    const int gx = (int)floor((me.x + Lh) / lp), gy = (int)floor((me.y + Lh) / lp);
    double r2 = lp * lp * 0.25 * 2;
    int bx = prev_x, by = prev_y;
    // distances to the three candidate columns / rows, each centre evaluated once (the divisions dominate this kernel)
    double ddx[3], ddy[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double drx = me.x - box_centre(gx - 1 + k, L, Lh, b), dry = me.y - box_centre(gy - 1 + k, L, Lh, b);
        ddx[k] = drx * drx; ddy[k] = dry * dry;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
        const int qy = gy - 1 + ky;
        if (qy < 0 || qy >= b) continue;
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            const int qx = gx - 1 + kx;
            if (qx < 0 || qx >= b) continue;
            double d2 = 0.0;
            d2 += ddx[kx];
            d2 += ddy[ky];
            if (d2 < r2) { r2 = d2; bx = qx; by = qy; }
        }
    }
    if (bx < 0) {  // reference would index grid[-1]; clamp to the floor bin instead
        bx = min(max(gx, 0), b - 1); by = min(max(gy, 0), b - 1);
    }
    return make_int2(bx, by);
}

__global__ void __launch_bounds__(RB_BLOCK) apj_bin_count_kernel(const DevState st, const int bps) {
    const int sys = blockIdx.x / bps;
    SysCtl* __restrict__ ctl = st.ctl + sys;
    if (!ctl->stale) return;
    const int b = ctl->b, col0 = ctl->col0, ncols = ctl->ncols;
    const double L = ctl->L, Lh = ctl->Lover2, lp = ctl->lp;
    const int cur = ctl->cur, gen = ctl->gen;
    const int n = ctl->n_own;
    for (int i = (blockIdx.x - sys * bps) * RB_BLOCK + threadIdx.x; i < n; i += bps * RB_BLOCK) {
        const long long g = (long long)ctl->p0 + i;
        const double2 me = st.XY[cur][g];
        const int prev = st.BOX[gen][g];      // internal numbering cy + (cx - col0) * b: columns contiguous
        const int2 q = nearest_box(me, L, Lh, lp, b, prev < 0 ? -1 : prev / b + col0, prev < 0 ? -1 : prev % b);
        const int lc = q.x - col0;
        if (lc >= 0 && lc < ncols) {
            const int best = q.y + lc * b;
            st.boxnew[g] = best;
            atomicAdd(st.cell_count + ctl->cell_base + best, 1);
            continue;
        }
        // slab mode: the particle left this rank's columns -> hand the whole record to the neighbour
        // that owns its column (a peer atomic reserves the inbox slot, a peer store fills it)
        st.boxnew[g] = -1;
        int dst = -1;
        if (q.x == (col0 + b - 1) % b) dst = st.left;
        else if (q.x == (col0 + ncols) % b) dst = st.right;
        if (dst < 0) { ctl->slab_err |= 2; continue; }
        MigRec r;
        r.xy = me; r.cs = st.CS[cur][g]; r.xr = st.XR[cur][g]; r.rr = st.RR[gen][g]; r.x0 = st.X0[gen][g];
        r.xo = st.XO[gen][g]; r.v = st.V[gen][g]; r.phi = st.PHI[gen][g]; r.id = st.ID[gen][g]; r.pad = 0;
        const int k = atomicAdd_system(&apj_peer(st, dst, st.mail)->inbox_count, 1);
        if (k < st.mcap) apj_peer(st, dst, st.inbox)[k] = r;
        else ctl->slab_err |= 4;
    }
}

// ---- slab mode: rendezvous of all ranks of the box (one warp). Everything this rank wrote into peer
// memory in earlier kernels of its stream is complete at this kernel's start; the release store of
// the sequence number publishes it, the acquire loads of the peers' numbers make theirs visible.
__global__ void apj_slab_sync_kernel(const DevState st, const int ch) {
    SysCtl* __restrict__ ctl = st.ctl;
    if (!ctl->stale) return;
    const int lane = threadIdx.x;
    const unsigned long long seq = ctl->seq[ch] + 1;
    __syncwarp();
    __threadfence_system();
    if (lane < st.nranks) apj_st_release_sys(&apj_peer(st, lane, st.mail)->flag[ch][st.rank], seq);
    bool ok = true;
    if (lane < st.nranks && !ctl->slab_err) ok = apj_wait_flag(&st.mail->flag[ch][lane], seq, st.timeout_ns);
    const unsigned missing = __ballot_sync(0xffffffffu, !ok);
    if (lane == 0) {
        ctl->seq[ch] = seq;
        if (missing) {
            ctl->slab_err |= 1;
            if (!ctl->slab_diag) ctl->slab_diag = ((ch + 1) << 16) | (int)missing;
        }
    }
}

// particles in the arrays before the sort: owned ones + the migrants the neighbours delivered
__device__ __forceinline__ int slab_n_src(const DevState& st, const SysCtl* ctl) {
    if (!st.slab) return ctl->n_own;
    const int n_in = min(__ldcg(&st.mail->inbox_count), st.mcap);
    return min(ctl->n_own + n_in, st.cap);
}

// ---- slab mode: append the delivered migrants to the arrays and bin them ----
__global__ void __launch_bounds__(RB_BLOCK) apj_absorb_kernel(const DevState st) {
    SysCtl* __restrict__ ctl = st.ctl;
    if (!ctl->stale) return;
    const int b = ctl->b, col0 = ctl->col0, ncols = ctl->ncols;
    const double L = ctl->L, Lh = ctl->Lover2, lp = ctl->lp;
    const int cur = ctl->cur, gen = ctl->gen;
    const int n_in_all = __ldcg(&st.mail->inbox_count);
    const int n_own = ctl->n_own;
    const int n_in = slab_n_src(st, ctl) - n_own;
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n_in_all > st.mcap || n_own + n_in_all > st.cap)) ctl->slab_err |= 4;
    for (int k = blockIdx.x * RB_BLOCK + threadIdx.x; k < n_in; k += gridDim.x * RB_BLOCK) {
        const MigRec r = st.inbox[k];
        const long long g = n_own + k;
        st.XY[cur][g] = r.xy; st.CS[cur][g] = r.cs; st.XR[cur][g] = r.xr; st.RR[gen][g] = r.rr; st.X0[gen][g] = r.x0;
        st.XO[gen][g] = r.xo; st.V[gen][g] = r.v; st.PHI[gen][g] = r.phi; st.ID[gen][g] = r.id; st.BOX[gen][g] = -1;
        const int2 q = nearest_box(r.xy, L, Lh, lp, b, -1, -1);
        const int lc = q.x - col0;
        if (lc >= 0 && lc < ncols) {
            const int best = q.y + lc * b;
            st.boxnew[g] = best;
            atomicAdd(st.cell_count + ctl->cell_base + best, 1);
        } else {           // cannot happen: the sender binned it with the same arithmetic
            st.boxnew[g] = -1;
            ctl->slab_err |= 2;
        }
    }
}

// ---- exclusive scan of every stale system's cell histogram, three short kernels -------------
// chunk sums (SCAN_CHUNK cells per block) -> per-system scan of the chunk sums -> in-chunk scan.
// Block-wide exclusive scan of SCAN_ITEMS values per thread; returns the block total.
__device__ __forceinline__ int block_scan_exclusive(int (&v)[SCAN_ITEMS], int* s_warp) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) tsum += v[k];
    int inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    int run = (wid ? s_warp[wid - 1] : 0) + inc - tsum;
    const int total = s_warp[SCAN_BLOCK / 32 - 1];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { const int c = v[k]; v[k] = run; run += c; }
    return total;
}

__global__ void __launch_bounds__(SCAN_BLOCK) apj_scan_chunk_sums_kernel(const DevState st, const int chunks) {
    const int sys = blockIdx.y;
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    if (!ctl->stale) return;
    const int nbox = ctl->nbox;
    const int first = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    if (blockIdx.x * SCAN_CHUNK >= nbox) return;
    const int* __restrict__ count = st.cell_count + ctl->cell_base;
    int tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) tsum += (first + k < nbox) ? count[first + k] : 0;
    __shared__ int s_warp[SCAN_BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = tsum;
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = s_warp[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) st.chunk_sums[(long long)sys * chunks + blockIdx.x] = w;
    }
}

__global__ void __launch_bounds__(SCAN_BLOCK) apj_scan_chunk_offsets_kernel(const DevState st, const int chunks) {
    const int sys = blockIdx.x;
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    if (!ctl->stale) return;
    const int used = (ctl->nbox + SCAN_CHUNK - 1) / SCAN_CHUNK;
    int* __restrict__ sums = st.chunk_sums + (long long)sys * chunks;
    __shared__ int s_warp[SCAN_BLOCK / 32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = ctl->p0;      // absolute particle index of the system's first slot
    __syncthreads();
    for (int base = 0; base < used; base += SCAN_BLOCK * SCAN_ITEMS) {
        int v[SCAN_ITEMS];
        const int first = base + threadIdx.x * SCAN_ITEMS;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (first + k < used) ? sums[first + k] : 0;
        const int carry = s_carry;
        const int total = block_scan_exclusive(v, s_warp);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) if (first + k < used) sums[first + k] = carry + v[k];
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) st.ctl[sys].n_end = s_carry;   // one past the last binned particle
}

__global__ void __launch_bounds__(SCAN_BLOCK) apj_scan_cells_kernel(const DevState st, const int chunks) {
    const int sys = blockIdx.y;
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    if (!ctl->stale) return;
    const int nbox = ctl->nbox;
    if (blockIdx.x * SCAN_CHUNK >= nbox) return;
    int* __restrict__ count = st.cell_count + ctl->cell_base;
    int* __restrict__ start = st.cell_start + ctl->cell_base;
    int* __restrict__ cursor = st.cell_cursor + ctl->cell_base;
    __shared__ int s_warp[SCAN_BLOCK / 32];
    const int first = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (first + k < nbox) ? count[first + k] : 0;
    const int off = st.chunk_sums[(long long)sys * chunks + blockIdx.x];
    block_scan_exclusive(v, s_warp);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (first + k < nbox) {
            start[first + k] = off + v[k];
            cursor[first + k] = off + v[k];
            count[first + k] = 0;  // leave the histogram clean for the next rebuild
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) start[nbox] = ctl->n_end;   // every particle that stays sits in some cell
}

__global__ void __launch_bounds__(RB_BLOCK) apj_scatter_kernel(const DevState st, const int bps) {
    const int sys = blockIdx.x / bps;
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    if (!ctl->stale) return;
    const int n = slab_n_src(st, ctl);
    for (int i = (blockIdx.x - sys * bps) * RB_BLOCK + threadIdx.x; i < n; i += bps * RB_BLOCK) {
        const long long g = (long long)ctl->p0 + i;
        const int c = st.boxnew[g];
        if (c < 0) continue;               // slab mode: emigrated
        const int slot = atomicAdd(st.cell_cursor + ctl->cell_base + c, 1);
        st.perm[slot] = (int)g;
    }
}

// ---- per-cell sort of the slots by original particle id (CellList order, jamming.cpp:546) ----
__global__ void __launch_bounds__(RB_BLOCK) apj_cell_sort_kernel(const DevState st) {
    const int sys = blockIdx.y;
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    if (!ctl->stale) return;
    const int* __restrict__ idv = st.ID[ctl->gen];
    const int* __restrict__ start = st.cell_start + ctl->cell_base;
    for (int c = blockIdx.x * RB_BLOCK + threadIdx.x; c < ctl->nbox; c += gridDim.x * RB_BLOCK) {
        const int s0 = start[c], n = start[c + 1] - s0;
        int* __restrict__ p = st.perm + s0;
        for (int a = 1; a < n; a++) {   // insertion sort; n is ~9 (max ~14 at phi <= 1)
            const int pa = p[a];
            const int ka = idv[pa];
            int q = a - 1;
            while (q >= 0 && idv[p[q]] > ka) { p[q + 1] = p[q]; q--; }
            p[q + 1] = pa;
        }
    }
}

__global__ void __launch_bounds__(RB_BLOCK) apj_reorder_kernel(const DevState st, const int bps) {
    const int sys = blockIdx.x / bps;
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    if (!ctl->stale) return;
    const int cur = ctl->cur, gen = ctl->gen;
    const int n = ctl->n_end - ctl->p0;
    for (int i = (blockIdx.x - sys * bps) * RB_BLOCK + threadIdx.x; i < n; i += bps * RB_BLOCK) {
    const long long k = (long long)ctl->p0 + i;
    const int src = st.perm[k];
    const double2 p = st.XY[cur][src];
    st.XY[cur ^ 1][k] = p;
    st.CS[cur ^ 1][k] = st.CS[cur][src];
    st.XR[cur ^ 1][k] = st.XR[cur][src];
    st.RR[gen ^ 1][k] = st.RR[gen][src];
    st.X0[gen ^ 1][k] = st.X0[gen][src];
    st.XO[gen ^ 1][k] = ctl->save_old ? p : st.XO[gen][src];  // saveOldPositions (jamming.cpp:825-835)
    st.V[gen ^ 1][k] = st.V[gen][src];
    st.PHI[gen ^ 1][k] = st.PHI[gen][src];
    st.ID[gen ^ 1][k] = st.ID[gen][src];
    st.BOX[gen ^ 1][k] = st.boxnew[src];
    }
}

// ---- slab mode: the first / last owned column becomes the left / right neighbour's ghost column ----
// Full records of what a neighbour's sweep reads ({x,y}, {cos,sin}, {R,1/R}, id) and the row starts of
// the column go straight into the neighbour's arrays (peer stores); between rebuilds the step kernel
// refreshes {x,y} and {cos,sin} of the same slots from its epilogue.
__global__ void __launch_bounds__(RB_BLOCK) apj_push_ghosts_kernel(const DevState st) {
    SysCtl* __restrict__ ctl = st.ctl;
    if (!ctl->stale) return;
    const int b = ctl->b, w = ctl->ncols;
    const int* __restrict__ start = st.cell_start + ctl->cell_base;
    const int h = ctl->cur ^ 1, gn = ctl->gen ^ 1;       // the halves reorder just wrote
    for (int d = 0; d < 2; d++) {
        const int c_lo = d == 0 ? 0 : (w - 1) * b;
        const int dst = d == 0 ? st.left : st.right;
        const int side = d == 0 ? 1 : 0;                 // my first column is the left neighbour's RIGHT ghost
        const int s0 = start[c_lo];
        int n = start[c_lo + b] - s0;
        if (n > st.gcap) { if (blockIdx.x == 0 && threadIdx.x == 0) ctl->slab_err |= 4; n = st.gcap; }
        const long long base = (long long)st.cap + (long long)side * st.gcap;
        double2* __restrict__ pXY = apj_peer(st, dst, st.XY[h]) + base;
        double2* __restrict__ pCS = apj_peer(st, dst, st.CS[h]) + base;
        double2* __restrict__ pRR = apj_peer(st, dst, st.RR[gn]) + base;
        int* __restrict__ pID = apj_peer(st, dst, st.ID[gn]) + base;
        for (int i = blockIdx.x * RB_BLOCK + threadIdx.x; i < n; i += gridDim.x * RB_BLOCK) {
            pXY[i] = st.XY[h][s0 + i]; pCS[i] = st.CS[h][s0 + i]; pRR[i] = st.RR[gn][s0 + i]; pID[i] = st.ID[gn][s0 + i];
        }
        int* __restrict__ pgs = apj_peer(st, dst, st.gstart) + ((size_t)gn * 2 + side) * (b + 1);
        for (int r = blockIdx.x * RB_BLOCK + threadIdx.x; r <= b; r += gridDim.x * RB_BLOCK)
            pgs[r] = (int)base + min(start[c_lo + r] - s0, n);
        if (blockIdx.x == 0 && threadIdx.x == 0) apj_peer(st, dst, st.mail)->ghost_n[gn][side] = n;
    }
}

// ---- work decomposition: blocks of <= ppb particles inside one cell column ---------------
__device__ __forceinline__ void add_piece(TileDesc& d, int& npieces, int& slots, int p0, int p1) {
    if (p1 <= p0) return;
    d.pstart[npieces] = p0;
    d.plen[npieces] = p1 - p0;
    npieces++;
    slots += p1 - p0;
}

__global__ void __launch_bounds__(SCAN_BLOCK) apj_make_tiles_kernel(const DevState st) {
    const int sys = blockIdx.x;
    SysCtl* __restrict__ ctl = st.ctl + sys;
    if (!ctl->stale) return;
    const int b = ctl->b, w = ctl->ncols;
    const int* __restrict__ start = st.cell_start + ctl->cell_base;
    int* __restrict__ col_blk = st.col_blk + ctl->col_base;
    __shared__ int s_warp[SCAN_BLOCK / 32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ppb = st.ppb;
    // exclusive scan over columns of ceil(n_X / ppb)
    for (int base = 0; base < w; base += SCAN_BLOCK) {
        const int X = base + threadIdx.x;
        const int nb = (X < w) ? (start[(X + 1) * b] - start[X * b] + ppb - 1) / ppb : 0;
        int inc = nb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        if (X < w) col_blk[X] = carry + (wid ? s_warp[wid - 1] : 0) + inc - nb;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) s_carry = carry + s_warp[SCAN_BLOCK / 32 - 1];
        __syncthreads();
    }
    const int nblk = s_carry;
    if (threadIdx.x == 0) {
        col_blk[w] = nblk;
        ctl->nblk = nblk;
        ctl->last_col_start = start[(w - 1) * b];
        ctl->blk_shift = (st.slab && st.nranks > 1) ? col_blk[w - 1] : 0;
        ctl->tile_max = 0;
    }
}

// descriptors: one thread per work block (column found by bisection of the per-column block offsets)
__global__ void __launch_bounds__(RB_BLOCK) apj_fill_tiles_kernel(const DevState st) {
    const int sys = blockIdx.y;
    SysCtl* __restrict__ ctl = st.ctl + sys;
    if (!ctl->stale) return;
    const int blk = blockIdx.x * RB_BLOCK + threadIdx.x;
    const int b = ctl->b, w = ctl->ncols, col0 = ctl->col0;
    const int* __restrict__ col_blk = st.col_blk + ctl->col_base;
    const int nblk = col_blk[w];
    int slots = 0;
    if (blk < nblk) {
        const int* __restrict__ start = st.cell_start + ctl->cell_base;
        const int* __restrict__ gs = st.slab ? st.gstart + (size_t)(ctl->gen ^ 1) * 2 * (b + 1) : nullptr;   // ghost columns [side][b+1]
        const int* __restrict__ box = st.BOX[ctl->gen ^ 1];   // the half reorder just wrote
        const int ppb = st.ppb;
        int lo = 0, hi = w - 1;                            // last column X with col_blk[X] <= blk (empty columns share offsets)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (col_blk[mid] <= blk) lo = mid; else hi = mid - 1;
        }
        const int X = lo;
        const int c1 = start[(X + 1) * b];
        const int g0 = start[X * b] + (blk - col_blk[X]) * ppb;
        const int Xg = X + col0;                           // column of the global grid
        TileDesc d;
        d.g0 = g0; d.n = min(ppb, c1 - g0); d.own_slot = 0;
        for (int p = 0; p < APJ_MAX_PIECES; p++) { d.pstart[p] = 0; d.plen[p] = 0; }
        const int cy0 = box[g0] - X * b, cy1 = box[g0 + d.n - 1] - X * b;
        int npieces = 0;
        // minimum-image wraps can only fire in blocks that touch the periodic seam
        const int wraps = (b < 7 || Xg == 0 || Xg == b - 1 || cy0 == 0 || cy1 == b - 1) ? 1 : 0;
        for (int dx = -1; dx <= 1; dx++) {
            int Xd = X + dx;
            const int* __restrict__ cs;
            if (st.slab) {                             // beyond the slab: the neighbour's column, held as a ghost
                cs = Xd < 0 ? gs : (Xd >= w ? gs + (b + 1) : start + Xd * b);
            } else {
                if (Xd < 0) Xd += b; else if (Xd >= b) Xd -= b;
                cs = start + Xd * b;
            }
            const int lo_r = cy0 - 1, hi_r = cy1 + 1;
            const int before = slots;
            int piece_of_own = -1;
            if (hi_r - lo_r + 1 >= b) {                // rows cover the whole column
                piece_of_own = npieces; add_piece(d, npieces, slots, cs[0], cs[b]);
            } else if (lo_r < 0) {                     // wraps below: rows [0..hi] then [b-1]
                piece_of_own = npieces; add_piece(d, npieces, slots, cs[0], cs[hi_r + 1]);
                add_piece(d, npieces, slots, cs[lo_r + b], cs[b]);
            } else if (hi_r > b - 1) {                 // wraps above: rows [lo..b-1] then [0]
                piece_of_own = npieces; add_piece(d, npieces, slots, cs[lo_r], cs[b]);
                add_piece(d, npieces, slots, cs[0], cs[hi_r - b + 1]);
            } else {
                piece_of_own = npieces; add_piece(d, npieces, slots, cs[lo_r], cs[hi_r + 1]);
            }
            if (dx == 0)     // the block's own particles sit in the first piece of the middle column
                d.own_slot = 1 + before + (g0 - d.pstart[piece_of_own]);   // slots are 1-based (0 = sentinel)
        }
        d.info = npieces | (wraps ? APJ_INFO_WRAPS : 0);
        if (st.slab && X == 0) d.info |= APJ_INFO_PUSH_LEFT;
        if (st.slab && X == w - 1) d.info |= APJ_INFO_PUSH_RIGHT;
        st.tiles[(long long)sys * st.maxblk + blk] = d;
    }
    // largest tile of the decomposition / capacity check: one atomic per warp
    int m = slots;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) {
        atomicMax(&ctl->tile_max, m);
        if (m > st.tile_cap || m > 4094) atomicOr(&ctl->overflow, 2);   // 16-bit entries hold slot * 16
    }
}

// ---- buildVerletLists (jamming.cpp:550-585) as a full list in tile-slot form -----------------
// Same work blocks and shared-memory tile as the step kernel: one thread issues TMA bulk copies
// of the {x,y} pieces, then every thread walks the 3x3 cells around its particle (three runs of
// consecutive tile slots, or nine single rows where y wraps) straight out of shared memory --
// neighbouring lanes walk the same runs, so the reads are broadcasts. Two passes: count, then
// write. Entries are stored in APJ_CLASSES classes of build distance (class k: d < rn + (k+1) *
// skin / APJ_CLASSES), each class in slot order. Class 0 is the first coordination shell, so lanes
// of a warp agree on the d2 < rn2 branch at almost every list position; and a step only sweeps the
// classes that can have come within rn given the skin-test value D (apj_device.cuh, APJ_CLASSES):
// cntk[] holds the cumulative list length through each class. The pair SET is what the reference
// defines (d2 < rs2); the order inside a list is ours.

// One pass over the candidates, one thread per particle. Everything a hit triggers is predicated straight-line
// code (class tag, 16-bit store into the thread's column of a shared-memory staging array, two counters), so the
// lanes of a warp stay together for the whole 3x3 sweep; a second, short pass over the ~16 staged hits places
// them in class-major order at their final position in the block's [quad][thread] list array. (Round 1 walked the
// candidates twice with the placement inside a divergent branch: 333 warp-instructions per particle at 14 active
// lanes, 6.5 ms at N = 16M; see profiles/r02_a_apj_verlet_build_kernel_ncu_full.txt.)
#ifndef APJ_BUILD_BLOCKS
#define APJ_BUILD_BLOCKS 5   // 48 registers, no spills: 5 blocks per SM when the staging array leaves room (S <= 48)
#endif
template <bool WRAP>
__device__ __forceinline__ void build_lists(const DevState& st, SysCtl* __restrict__ ctl, const TileDesc& sd, const long long bg,
                                            const double2* __restrict__ sXY, unsigned short* __restrict__ stage, const int lgG) {
    const int t = threadIdx.x, TBn = blockDim.x;
    const int g = sd.g0 + t;
    const int gen = ctl->gen ^ 1;                      // the half reorder just wrote
    const int* __restrict__ start = st.cell_start + ctl->cell_base;
    const int b = ctl->b, w = ctl->ncols;
    const int* __restrict__ gs = st.slab ? st.gstart + (size_t)gen * 2 * (b + 1) : nullptr;
    const double L = ctl->L, Lh = ctl->Lover2, rs2 = st.rs2, rn2 = st.rn2, cinv = st.cls_inv;
    static_assert(APJ_CLASSES == 4, "class tags are two bits, cumulative counts four bytes");
    const unsigned base = apj_smem_addr(sXY), a_own = base + (unsigned)(sd.own_slot + t) * 16u;
    const double2 me = apj_lds_f64x2(a_own);
    const int c = st.BOX[gen][g];
    const int cx = c / b, cy = c - cx * b;
    const int S = st.S;
    const bool interior = cy >= 1 && cy <= b - 2;      // rows cy-1..cy+1 are one contiguous run of the column
    const int nruns = interior ? 1 : 3;
    int n = 0;                                         // hits so far (entries of the full list)
    unsigned short* __restrict__ sp = stage + t;       // this thread's column of the staging array, one row per hit
    for (int dcx = -1; dcx <= 1; dcx++) {
        int col = cx + dcx;
        const int* __restrict__ cs;
        if (gs) {                                      // slab mode: ghost columns beyond the slab
            cs = col < 0 ? gs : (col >= w ? gs + (b + 1) : start + col * b);
        } else {
            if (col < 0) col += b; else if (col >= b) col -= b;
            cs = start + col * b;
        }
        for (int r = 0; r < nruns; r++) {
            int a, e;
            if (interior) { a = cs[cy - 1]; e = cs[cy + 2]; }
            else {                                     // y wraps: row by row
                int row = cy - 1 + r;
                if (row < 0) row += b; else if (row >= b) row -= b;
                a = cs[row]; e = cs[row + 1];
            }
            if (e <= a) continue;
            // the loop runs on the shared-memory byte address of the candidate's slot: one induction variable
            const unsigned a0 = base + (unsigned)apj_slot_of(sd, a) * 16u, a1 = a0 + (unsigned)(e - a) * 16u;
            for (unsigned at = a0; at < a1; at += 16u) {
                const double2 q = apj_lds_f64x2(at);
                double dx = q.x - me.x, dy = q.y - me.y;
                if (WRAP) { dx = apj_wrap1(dx, L, Lh); dy = apj_wrap1(dy, L, Lh); }
                const double d2 = apj_d2(dx, dy);
                if (d2 < rs2 && at != a_own) {         // lanes of a warp hit at different candidates, so this body is issued for
                    *sp = (unsigned short)((at - base) >> 4);   // almost every candidate: it only records the slot (class: second pass)
                    n++;
                    if (n < S) sp += TBn;              // a list that outgrows S keeps overwriting its last row: it is discarded below
                }
            }
        }
    }
    const int total = n;
    const int nn = total <= S ? total : 0;             // a list that does not fit is not used at all (overflow is flagged below)
    // hits per class (one byte each) -> cumulative counts for cntk, exclusive ones = running write position of each class
    // the build-distance class of every hit, from the same d2 bits (recomputed: ~20 hits per thread against ~90 candidates)
    unsigned packed = 0;
    for (int e = 0; e < nn; e++) {
        const unsigned slot = stage[e * TBn + t];
        const double2 q = apj_lds_f64x2(base + slot * 16u);
        double dx = q.x - me.x, dy = q.y - me.y;
        if (WRAP) { dx = apj_wrap1(dx, L, Lh); dy = apj_wrap1(dy, L, Lh); }
        const unsigned cls = (unsigned)apj_entry_class(apj_d2(dx, dy), rn2, cinv);
        stage[e * TBn + t] = (unsigned short)((slot << 2) | cls);
        packed += 1u << (cls * 8u);
    }
    const unsigned k0 = packed & 0xffu, k1 = (packed >> 8) & 0xffu, k2 = (packed >> 16) & 0xffu;
    const unsigned cum0 = k0, cum1 = cum0 + k1, cum2 = cum1 + k2;
    unsigned cur = (cum0 << 8) | (cum1 << 16) | (cum2 << 24);
    unsigned short* __restrict__ out = reinterpret_cast<unsigned short*>(st.list32 + bg * (long long)st.max_quads * st.tb * 4);
    const int G = 1 << lgG;
    // entry e -> word e/2 -> lane (e/2) % G, that lane's word (e/2) / G, half e & 1
    auto put = [&](int e, unsigned v) {
        const int wd = e >> 1, sub = wd & (G - 1), kk = wd >> lgG;
        out[(((size_t)(kk >> 2) * st.tb + t * G + sub) * 4 + (kk & 3)) * 2 + (e & 1)] = (unsigned short)v;
    };
    for (int e = 0; e < nn; e++) {
        const unsigned v = stage[e * TBn + t];
        const unsigned sh = (v & 3u) * 8u;
        const int pos = (int)((cur >> sh) & 0xffu);
        cur += 1u << sh;
        put(pos, (v >> 2) << 4);
    }
    if (nn & 1) put(nn, 0u);                           // pad the last word with the sentinel
    st.cnt[g] = nn;
    st.cntk[g] = total <= S ? (cum0 | (cum1 << 8) | (cum2 << 16) | ((unsigned)nn << 24)) : 0u;
    if (total > S) ctl->overflow |= 1;                 // list capacity exceeded: the system stays stale, the host grows the storage
    if (total > ctl->list_max) atomicMax(&ctl->list_max, total);
}

__global__ void __launch_bounds__(APJ_TB_MAX, (APJ_TB_MAX > 256 ? 2 : APJ_BUILD_BLOCKS)) apj_verlet_build_kernel(const DevState st, const int lgG) {
    const int sys = blockIdx.x / st.maxblk;
    const int blk = blockIdx.x - sys * st.maxblk;
    SysCtl* __restrict__ ctl = st.ctl + sys;
    if (!ctl->stale || blk >= ctl->nblk || (ctl->overflow & 2)) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* __restrict__ sXY = reinterpret_cast<double2*>(smem_raw);   // slot 0 unused here (slots are 1-based)
    unsigned short* __restrict__ stage = reinterpret_cast<unsigned short*>(sXY + (st.tile_cap + 1));   // [S][blockDim.x]
    __shared__ TileDesc sd;
    __shared__ __align__(8) unsigned long long s_bar;
    const long long bg = (long long)sys * st.maxblk + blk;
    if (threadIdx.x < 16) reinterpret_cast<int*>(&sd)[threadIdx.x] = reinterpret_cast<const int*>(st.tiles + bg)[threadIdx.x];
    if (threadIdx.x == 0) apj_mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int npieces = sd.info & 0xff;
        const double2* __restrict__ P = st.XY[ctl->cur ^ 1];           // the half reorder just wrote
        int slots = 0;
        for (int p = 0; p < npieces; p++) slots += sd.plen[p];
        apj_mbar_expect_tx(&s_bar, (unsigned)slots * 16u);
        int off = 1;
        for (int p = 0; p < npieces; p++) {
            apj_bulk_g2s(sXY + off, P + sd.pstart[p], (unsigned)sd.plen[p] * 16u, &s_bar);
            off += sd.plen[p];
        }
    }
    unsigned dep = 0;
    apj_mbar_wait(&s_bar, 0, dep);
    if (threadIdx.x < sd.n) {
        if (sd.info & APJ_INFO_WRAPS) build_lists<true>(st, ctl, sd, bg, sXY, stage, lgG);
        else build_lists<false>(st, ctl, sd, bg, sXY, stage, lgG);
    }
}

__global__ void apj_finish_rebuild_kernel(const DevState st) {
    const int sys = blockIdx.x * blockDim.x + threadIdx.x;
    if (sys >= st.n_sys) return;
    SysCtl* ctl = st.ctl + sys;
    if (!ctl->stale) return;
    if (ctl->overflow & 2) return;   // tile larger than the launched shared-memory capacity: nothing was
                                     // committed, the system stays stale until the host re-launches bigger
    if ((ctl->overflow & 1) && !st.slab) return;   // a list did not fit: same -- no step may run on truncated lists;
                                     // the host grows the list storage and runs the chain again (slab ranks move in
                                     // lockstep and cannot stall alone: there the host reports the overflow)
    ctl->cur ^= 1;
    ctl->gen ^= 1;
    if (ctl->save_old) {   // newSkinList fired (jamming.cpp:611-615): resetCounter++, saveOldPositions
        ctl->COM_old[0] = ctl->COM[0];
        ctl->COM_old[1] = ctl->COM[1];
        ctl->reset_counter += 1;
    }
    ctl->n_rebuilds += 1;
    // skin-aware sweep length: valid only while x_old is the state the lists were built from
    ctl->trunc_ok = ctl->save_old ? 1 : 0;
    ctl->skinD = 0.0;
    ctl->skinBase = 0.0;
    ctl->kmin = 0;
    ctl->save_old = 0;
    ctl->stale = 0;
    ctl->n_own = ctl->n_end - ctl->p0;
    if (st.slab) st.mail->inbox_count = 0;   // neighbours push again only after the next step was exchanged
}

}  // namespace

size_t apj_build_smem_bytes(const DevState& st);
void apj_launch_rebuild_chain(const DevState& st, const ApjLaunch& l, int max_nbox, int max_b) {
    (void)max_b;
    // fixed grids: ~16 blocks of 128 threads per SM in total, shared out over the systems
    const int bps = std::max(1, std::min((st.cap + RB_BLOCK - 1) / RB_BLOCK, (148 * 16 + st.n_sys - 1) / st.n_sys));
    const int grid = st.n_sys * bps;
    apj_bin_count_kernel<<<grid, RB_BLOCK, 0, l.stream>>>(st, bps); apj_check_launch("apj_bin_count_kernel");
    if (st.slab) {
        apj_slab_sync_kernel<<<1, 32, 0, l.stream>>>(st, 1); apj_check_launch("apj_slab_sync_kernel");       // all migrants delivered
        apj_absorb_kernel<<<std::max(1, std::min((st.mcap + RB_BLOCK - 1) / RB_BLOCK, 148 * 4)), RB_BLOCK, 0, l.stream>>>(st); apj_check_launch("apj_absorb_kernel");
    }
    const int chunks = (max_nbox + SCAN_CHUNK - 1) / SCAN_CHUNK;       // == DevState::scan_chunks
    dim3 gc(chunks, st.n_sys);
    apj_scan_chunk_sums_kernel<<<gc, SCAN_BLOCK, 0, l.stream>>>(st, chunks); apj_check_launch("apj_scan_chunk_sums_kernel");
    apj_scan_chunk_offsets_kernel<<<st.n_sys, SCAN_BLOCK, 0, l.stream>>>(st, chunks); apj_check_launch("apj_scan_chunk_offsets_kernel");
    apj_scan_cells_kernel<<<gc, SCAN_BLOCK, 0, l.stream>>>(st, chunks); apj_check_launch("apj_scan_cells_kernel");
    apj_scatter_kernel<<<grid, RB_BLOCK, 0, l.stream>>>(st, bps); apj_check_launch("apj_scatter_kernel");
    dim3 gs(std::max(1, std::min((max_nbox + RB_BLOCK - 1) / RB_BLOCK, (148 * 16 + st.n_sys - 1) / st.n_sys)), st.n_sys);
    apj_cell_sort_kernel<<<gs, RB_BLOCK, 0, l.stream>>>(st); apj_check_launch("apj_cell_sort_kernel");
    apj_reorder_kernel<<<grid, RB_BLOCK, 0, l.stream>>>(st, bps); apj_check_launch("apj_reorder_kernel");
    if (st.slab) {
        apj_push_ghosts_kernel<<<std::max(1, std::min((st.gcap + RB_BLOCK - 1) / RB_BLOCK, 148 * 4)), RB_BLOCK, 0, l.stream>>>(st); apj_check_launch("apj_push_ghosts_kernel");
        apj_slab_sync_kernel<<<1, 32, 0, l.stream>>>(st, 2); apj_check_launch("apj_slab_sync_kernel");       // ghost columns in place on every rank
    }
    apj_make_tiles_kernel<<<st.n_sys, SCAN_BLOCK, 0, l.stream>>>(st); apj_check_launch("apj_make_tiles_kernel");
    apj_fill_tiles_kernel<<<dim3((st.maxblk + RB_BLOCK - 1) / RB_BLOCK, st.n_sys), RB_BLOCK, 0, l.stream>>>(st); apj_check_launch("apj_fill_tiles_kernel");
    int lgG = 0;
    while ((1 << lgG) < st.G) lgG++;
    apj_verlet_build_kernel<<<st.n_sys * st.maxblk, st.ppb, apj_build_smem_bytes(st), l.stream>>>(st, lgG); apj_check_launch("apj_verlet_build_kernel");
    apj_finish_rebuild_kernel<<<(st.n_sys + 63) / 64, 64, 0, l.stream>>>(st); apj_check_launch("apj_finish_rebuild_kernel");
    if (l.launch_counter) (*l.launch_counter) += apj_rebuild_chain_launches(st);
}
int apj_rebuild_chain_launches(const DevState& st) { return st.slab ? 15 : 11; }

int apj_max_list_capacity() { return MAX_S; }
int apj_scan_chunk_cells() { return SCAN_CHUNK; }
// dynamic shared memory of the list build: the {x,y} tile + the staging array of 16-bit hits, [S][particles of the block]
size_t apj_build_smem_bytes(const DevState& st) { return (size_t)(st.tile_cap + 1) * 16 + (size_t)st.S * st.ppb * 2; }
int apj_configure_rebuild(const DevState& st) {
    (void)st;
    return apj_allow_max_smem(apj_verlet_build_kernel) ? 0 : -1;
}
