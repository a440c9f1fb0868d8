// On-GPU observables. Reductions are two-level with a fixed order (block partials, then one
// block per system), so repeated calls on the same state return identical bits; histograms
// use integer atomics (exact); the pair-correlation sums use shared-memory fp64 atomics.
#include <algorithm>
#include <cmath>
#include "apj_observe.cuh"

namespace {

constexpr int OB = 256;

enum ObsMode { OBS_COM = 0, OBS_ORDER = 1, OBS_MSD = 2, OBS_FLUCT = 3 };

// Fluctuations::overlap, 2D branch (classes/Fluctuations.h:89-120): lens area of a disk of
// radius r whose centre is at distance d from the centre of a circle of radius R.
__device__ __forceinline__ double lens_overlap(double r, double R, double d) {
    if (R >= r + d) return APJ_PI * r * r;
    const double R12 = r * r, R22 = R * R;
    double x = (R12 - R22 + d * d) / (2.0 * d);
    double theta = acos(x / r);
    double A = R12 * theta - x * r * sin(theta);
    x = d - x;
    theta = acos(x / R);
    A += R22 * theta - x * R * sin(theta);
    return A;
}

template <int MODE>
__global__ void __launch_bounds__(OB) apj_obs_reduce_kernel(const DevState st, const int bps, const double* __restrict__ param,
                                                            double* __restrict__ part) {
    const int sys = blockIdx.x / bps, blk = blockIdx.x - sys * bps;
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    double a = 0.0, b = 0.0;
    for (int i = blk * OB + threadIdx.x; i < ctl->n_own; i += bps * OB) {
        const long long g = (long long)ctl->p0 + i;
        if (MODE == OBS_COM) {
            const double2 xr = st.XR[ctl->cur][g];
            a += xr.x; b += xr.y;
        } else if (MODE == OBS_ORDER) {   // jamming.cpp:780-786 / :795-801, Cell::get_speed (Cell.h:177-181)
            const double2 v = st.V[ctl->gen][g];
            const double inverseVel = 1.0 / sqrt(v.x * v.x + v.y * v.y);
            a += v.x * inverseVel; b += v.y * inverseVel;
        } else if (MODE == OBS_MSD) {     // jamming.cpp:811-820
            const double2 xr = st.XR[ctl->cur][g], x0 = st.X0[ctl->gen][g];
            const double dx = ((xr.x - x0.x) - ctl->COM[0]) + ctl->COM0[0];
            const double dy = ((xr.y - x0.y) - ctl->COM[1]) + ctl->COM0[1];
            a += dx * dx + dy * dy;
        } else {                          // Fluctuations::measureFluctuations inner loop (Fluctuations.h:62-76)
            const double radius = param[sys];
            const double2 xr = st.XR[ctl->cur][g];
            const double Ri = st.RR[ctl->gen][g].x;
            const double sumR = Ri + radius;
            const double dx = apj_delta_norm(xr.x - ctl->COM[0], ctl->L, ctl->Lover2);
            const double dy = apj_delta_norm(xr.y - ctl->COM[1], ctl->L, ctl->Lover2);
            const double d2 = dx * dx + dy * dy;
            if (d2 <= sumR * sumR) a += lens_overlap(Ri, radius, sqrt(d2));
        }
    }
    __shared__ double sa[OB / 32], sb[OB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < OB / 32; w++) { a += sa[w]; b += sb[w]; }
        part[2 * (size_t)blockIdx.x] = a; part[2 * (size_t)blockIdx.x + 1] = b;
    }
}

__global__ void __launch_bounds__(OB) apj_obs_final_kernel(const int bps, const double* __restrict__ part, double* __restrict__ out) {
    const int sys = blockIdx.x;
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < bps; k += OB) { a += part[2 * ((size_t)sys * bps + k)]; b += part[2 * ((size_t)sys * bps + k) + 1]; }
    __shared__ double sa[OB / 32], sb[OB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < OB / 32; w++) { a += sa[w]; b += sb[w]; }
        out[2 * sys] = a; out[2 * sys + 1] = b;
    }
}

// Correlations::velDist (Correlations.h:179-187)
__global__ void __launch_bounds__(OB) apj_velhist_kernel(const DevState st, const int bps, const double* __restrict__ dv,
                                                         unsigned long long* __restrict__ hist) {
    const int sys = blockIdx.x / bps, blk = blockIdx.x - sys * bps;
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    __shared__ unsigned sh[100];
    for (int k = threadIdx.x; k < 100; k += OB) sh[k] = 0;
    __syncthreads();
    for (int i = blk * OB + threadIdx.x; i < ctl->n_own; i += bps * OB) {
        const double2 v = st.V[ctl->gen][(long long)ctl->p0 + i];
        const double sp = sqrt(v.x * v.x + v.y * v.y);
        const double q = floor(sp / dv[sys]);
        if (q < 100.0) { const int bin = (int)q; if (bin >= 0) atomicAdd(&sh[bin], 1u); }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 100; k += OB) if (sh[k]) atomicAdd(&hist[(size_t)sys * 128 + k], (unsigned long long)sh[k]);
}

// Fluctuations::density_distribution (Fluctuations.h:122-139): occupancy of every box
__global__ void __launch_bounds__(OB) apj_occupancy_kernel(const DevState st, unsigned long long* __restrict__ hist) {
    const int sys = blockIdx.y;
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    __shared__ unsigned sh[50];
    for (int k = threadIdx.x; k < 50; k += OB) sh[k] = 0;
    __syncthreads();
    const int* __restrict__ start = st.cell_start + ctl->cell_base;
    for (int c = blockIdx.x * OB + threadIdx.x; c < ctl->nbox; c += gridDim.x * OB) {
        const int n = start[c + 1] - start[c];
        if (n < 50) atomicAdd(&sh[n], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 50; k += OB) if (sh[k]) atomicAdd(&hist[(size_t)sys * 128 + k], (unsigned long long)sh[k]);
}

// Correlations::spatialCorrelations accumulation (Correlations.h:85-152). Each unordered pair
// whose separation can fall inside a bin is visited once (from the particle with the lower
// position in cell order), over the distinct cells within `reach` of the particle's cell.
// On a slab handle (several ranks) the columns are the rank's own: no wrap in x, pairs that cross the slab's right
// edge are the business of apj_spatial_ext_kernel below.
// One pair. The three correlation histograms have few, wide bins (dr_c = 2: 10 bins at the local cutoff 20), so a
// block's threads would all hammer the same few shared-memory words: with PRIV every THREAD owns its copy of those
// bins, laid out [bin][thread] (bank = thread: conflict-free, no atomics), folded over the block at the end. The
// g(r) histogram has 20 x as many bins and stays one shared copy with atomics. Without PRIV (cutoff 140: 70 bins
// per thread would not fit) everything is the shared copy.
template <bool PRIV>
__device__ __forceinline__ void corr_pair(double* sh, double* priv, const int nc, const int np, const double r, const double2 mcs, const double2 cj,
                                          const double2 vi, const double spi, const double2 vj) {
    const double dr_c = 2.0, dr_p = 0.1;   // Correlations.h:52-53
    const double qp = floor(r / dr_p), qc = floor(r / dr_c);
    if (qp < (double)np) atomicAdd(&sh[3 * nc + (int)qp], 1.0 / r);
    if (qc < (double)nc) {
        const int bin = (int)qc;
        const double o = mcs.x * cj.x + mcs.y * cj.y;
        const double v = (vi.x * vj.x + vi.y * vj.y) / (spi * sqrt(vj.x * vj.x + vj.y * vj.y));
        if (PRIV) {
            double* q = priv + (size_t)(3 * bin) * OB;
            q[0] += 1.0; q[OB] += o; q[2 * OB] += v;
        } else {
            atomicAdd(&sh[bin], 1.0);
            atomicAdd(&sh[nc + bin], o);
            atomicAdd(&sh[2 * nc + bin], v);
        }
    }
}
// fold the per-thread copies into the block's shared copy (fixed order: lanes in a shuffle tree, warps by atomics of
// whole-warp sums -- the cross-block order is that of the global atomics either way)
__device__ __forceinline__ void corr_fold_private(double* sh, const double* priv_base, const int nc) {
    for (int k = 0; k < 3 * nc; k++) {
        double a = priv_base[(size_t)k * OB + threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        const int bin = k / 3, which = k - 3 * bin;
        if ((threadIdx.x & 31) == 0 && a != 0.0) atomicAdd(&sh[which * nc + bin], a);
    }
}
template <bool PRIV>
__global__ void __launch_bounds__(OB) apj_spatial_kernel(const DevState st, const int bps, const int nc, const int np,
                                                         const int reach_scale, double* __restrict__ acc) {
    extern __shared__ double sh[];  // counts[nc] | ori[nc] | vel[nc] | pair[np] | PRIV: [3 nc][OB] per-thread copies
    const int sys = blockIdx.x / bps, blk = blockIdx.x - sys * bps;
    const SysCtl* __restrict__ ctl = st.ctl + sys;
    const int nb = 3 * nc + np;
    for (int k = threadIdx.x; k < nb; k += OB) sh[k] = 0.0;
    double* priv_base = sh + ((nb + 1) & ~1);
    double* priv = priv_base + threadIdx.x;
    if (PRIV) for (int k = 0; k < 3 * nc; k++) priv[(size_t)k * OB] = 0.0;
    __syncthreads();
    const double dr_c = 2.0, dr_p = 0.1;
    const int b = ctl->b, w = ctl->ncols;
    const bool clip = st.slab && st.nranks > 1;                  // own columns only, no wrap in x
    const double L = ctl->L, Lh = ctl->Lover2;
    const double rmax = fmax(nc * dr_c, np * dr_p);
    int reach = (int)floor((rmax + st.skin) / ctl->lp) + 1;   // cells are those of the last rebuild: particles may have moved by up to the skin since
    const int span = (2 * reach + 1 >= b) ? b : 2 * reach + 1;   // distinct columns / rows to visit
    const double2* __restrict__ P = st.XY[ctl->cur];
    const double2* __restrict__ CSv = st.CS[ctl->cur];
    const double2* __restrict__ V = st.V[ctl->gen];
    const int* __restrict__ start = st.cell_start + ctl->cell_base;
    const int* __restrict__ box = st.BOX[ctl->gen];
    const int i = blk * OB + threadIdx.x;
    if (i < ctl->n_own) {
        const long long g = (long long)ctl->p0 + i;
        const double2 me = P[g];
        const double2 mcs = CSv[g];
        const double2 vi = V[g];
        const double spi = sqrt(vi.x * vi.x + vi.y * vi.y);
        const int c = box[g];
        const int cx = c / b, cy = c - cx * b;
        const int xspan = clip ? 2 * reach + 1 : span;
        for (int ox = 0; ox < xspan; ox++) {
            int col = (!clip && span == b) ? ox : cx - reach + ox;
            if (clip) { if (col < 0 || col >= w) continue; }
            else { col %= b; if (col < 0) col += b; }
            for (int oy = 0; oy < span; oy++) {
                int row = (span == b) ? oy : cy - reach + oy;
                row %= b; if (row < 0) row += b;
                const int cc = col * b + row;
                const int j1 = start[cc + 1];
                for (int j = max(start[cc], (int)g + 1); j < j1; j++) {
                    const double2 pj = P[j];
                    const double dx = apj_wrap1(pj.x - me.x, L, Lh), dy = apj_wrap1(pj.y - me.y, L, Lh);
                    corr_pair<PRIV>(sh, priv, nc, np, sqrt(apj_d2(dx, dy)), mcs, CSv[j], vi, spi, V[j]);
                }
            }
        }
    }
    (void)reach_scale;
    __syncthreads();
    if (PRIV) { corr_fold_private(sh, priv_base, nc); __syncthreads(); }
    for (int k = threadIdx.x; k < nb; k += OB) if (sh[k] != 0.0) atomicAdd(&acc[(size_t)sys * nb + k], sh[k]);
}

// Slab mode: pairs between an owned particle and an EXTERNAL one -- the particles of the rank(s) to the right that lie
// within the cutoff of this slab's right edge, handed in by the caller sorted by cell row (ext_row[b+1]); planes
// x | y | cos | sin | vx | vy of n_ext values each. Every such pair is counted here and nowhere else.
__global__ void __launch_bounds__(OB) apj_spatial_ext_kernel(const DevState st, const int nc, const int np, const long long n_ext,
                                                             const double* __restrict__ ext, const int* __restrict__ ext_row,
                                                             double* __restrict__ acc) {
    extern __shared__ double sh[];
    const SysCtl* __restrict__ ctl = st.ctl;
    const int nb = 3 * nc + np;
    for (int k = threadIdx.x; k < nb; k += OB) sh[k] = 0.0;
    __syncthreads();
    const int b = ctl->b;
    const double L = ctl->L, Lh = ctl->Lover2, lp = ctl->lp;
    const double rmax = fmax(nc * 2.0, np * 0.1);
    const int reach = (int)floor((rmax + st.skin) / lp) + 1;
    const int span = (2 * reach + 1 >= b) ? b : 2 * reach + 1;
    const double x_edge = -Lh + (double)(ctl->col0 + ctl->ncols) * lp;     // right edge of the slab
    const int i = blockIdx.x * OB + threadIdx.x;
    if (i < ctl->n_own) {
        const long long g = (long long)ctl->p0 + i;
        const double2 me = st.XY[ctl->cur][g];
        // near enough to the edge to reach across it (minimum image: an owned particle that drifted over the periodic seam
        // since the last rebuild has been wrapped to the other side of the box)
        if (apj_wrap1(x_edge - me.x, L, Lh) < rmax + lp) {
            const double2 mcs = st.CS[ctl->cur][g];
            const double2 vi = st.V[ctl->gen][g];
            const double spi = sqrt(vi.x * vi.x + vi.y * vi.y);
            const int c = st.BOX[ctl->gen][g];
            const int cy = c - (c / b) * b;
            const double *ex = ext, *ey = ext + n_ext, *ec = ext + 2 * n_ext, *es = ext + 3 * n_ext, *evx = ext + 4 * n_ext, *evy = ext + 5 * n_ext;
            for (int oy = 0; oy < span; oy++) {
                int row = (span == b) ? oy : cy - reach + oy;
                row %= b; if (row < 0) row += b;
                for (int j = ext_row[row]; j < ext_row[row + 1]; j++) {
                    const double dx = apj_wrap1(ex[j] - me.x, L, Lh), dy = apj_wrap1(ey[j] - me.y, L, Lh);
                    corr_pair<false>(sh, nullptr, nc, np, sqrt(apj_d2(dx, dy)), mcs, make_double2(ec[j], es[j]), vi, spi, make_double2(evx[j], evy[j]));
                }
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += OB) if (sh[k] != 0.0) atomicAdd(&acc[k], sh[k]);
}

// owned particles within `width` of the slab's LEFT edge, compacted (order irrelevant): what the left neighbour
// needs as its external set. out = 6 planes of `cap` values; *count receives the number found (may exceed cap).
__global__ void __launch_bounds__(OB) apj_export_edge_kernel(const DevState st, const double width, const long long cap, double* __restrict__ out,
                                                             int* __restrict__ count) {
    const SysCtl* __restrict__ ctl = st.ctl;
    const double x_left = -ctl->Lover2 + (double)ctl->col0 * ctl->lp;
    for (int i = blockIdx.x * OB + threadIdx.x; i < ctl->n_own; i += gridDim.x * OB) {
        const long long g = (long long)ctl->p0 + i;
        const double2 p = st.XY[ctl->cur][g];
        if (apj_wrap1(p.x - x_left, ctl->L, ctl->Lover2) < width) {   // minimum image: drifted over the periodic seam -> wrapped
            const int k = atomicAdd(count, 1);
            if (k < cap) {
                const double2 cs = st.CS[ctl->cur][g], v = st.V[ctl->gen][g];
                out[k] = p.x; out[cap + k] = p.y; out[2 * cap + k] = cs.x; out[3 * cap + k] = cs.y; out[4 * cap + k] = v.x; out[5 * cap + k] = v.y;
            }
        }
    }
}

// queue one two-level reduction on the stream; its 2 * n_sys raw sums land in d_dst (no host synchronisation)
int enqueue2(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, int mode, const double* h_param, double* d_dst) {
    const int bps = o->blocks_per_sys;
    if (h_param) if (cudaMemcpyAsync(o->d_param, h_param, sizeof(double) * st.n_sys, cudaMemcpyHostToDevice, s) != cudaSuccess) return APJ_E_CUDA_OBS;
    const int grid = st.n_sys * bps;
    switch (mode) {
        case OBS_COM: apj_obs_reduce_kernel<OBS_COM><<<grid, OB, 0, s>>>(st, bps, o->d_param, o->d_part); break;
        case OBS_ORDER: apj_obs_reduce_kernel<OBS_ORDER><<<grid, OB, 0, s>>>(st, bps, o->d_param, o->d_part); break;
        case OBS_MSD: apj_obs_reduce_kernel<OBS_MSD><<<grid, OB, 0, s>>>(st, bps, o->d_param, o->d_part); break;
        default: apj_obs_reduce_kernel<OBS_FLUCT><<<grid, OB, 0, s>>>(st, bps, o->d_param, o->d_part); break;
    }
    apj_obs_final_kernel<<<st.n_sys, OB, 0, s>>>(bps, o->d_part, d_dst);
    if (launches) *launches += 2;
    return cudaGetLastError() == cudaSuccess ? 0 : APJ_E_CUDA_OBS;
}
int reduce2(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, int mode, const double* h_param, double* h_out2) {
    if (int rc = enqueue2(o, st, s, launches, mode, h_param, o->d_out)) return rc;
    if (cudaMemcpyAsync(h_out2, o->d_out, sizeof(double) * 2 * st.n_sys, cudaMemcpyDeviceToHost, s) != cudaSuccess) return APJ_E_CUDA_OBS;
    if (cudaStreamSynchronize(s) != cudaSuccess) return APJ_E_CUDA_OBS;
    return 0;
}

}  // namespace

int apj_obs_alloc(ApjObsScratch* o, const DevState& st, cudaStream_t stream, std::vector<void*>& allocs) {
    o->blocks_per_sys = std::max(1, std::min((st.cap + OB - 1) / OB, 1024));
    auto A = [&](void** p, size_t bytes) {
        if (cudaMalloc(p, bytes) != cudaSuccess) return APJ_E_CUDA_OBS;
        cudaMemsetAsync(*p, 0, bytes, stream);
        allocs.push_back(*p);
        return 0;
    };
    if (A((void**)&o->d_part, sizeof(double) * 2 * (size_t)st.n_sys * o->blocks_per_sys)) return APJ_E_CUDA_OBS;
    if (A((void**)&o->d_out, sizeof(double) * 2 * (size_t)st.n_sys)) return APJ_E_CUDA_OBS;
    if (A((void**)&o->d_param, sizeof(double) * (size_t)st.n_sys)) return APJ_E_CUDA_OBS;
    if (A((void**)&o->d_hist, sizeof(unsigned long long) * 128 * (size_t)st.n_sys)) return APJ_E_CUDA_OBS;
    o->ring_cap = APJ_OBS_RING;
    if (A((void**)&o->d_ring, sizeof(double) * 2 * (size_t)st.n_sys * o->ring_cap)) return APJ_E_CUDA_OBS;
    return 0;
}

// Queued observables: the reduction is placed on the stream behind the steps already queued and its raw sums go
// to slot (ticket % ring) of a device-side ring; apj_obs_fetch_ring reads any run of tickets back in one copy.
// The reference's measurement cadence (fluctuations every 10 steps, jamming.cpp:211-216) then costs no host
// round trip per measurement.
int apj_obs_enqueue_ring(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, int kind, const double* h_param, long long* ticket) {
    if (kind < OBS_COM || kind > OBS_FLUCT) return -1;
    if (kind == OBS_FLUCT && !h_param) return -1;
    const long long tk = o->next_ticket;
    if (int rc = enqueue2(o, st, s, launches, kind, kind == OBS_FLUCT ? h_param : nullptr, o->d_ring + 2 * (size_t)st.n_sys * (size_t)(tk % o->ring_cap))) return rc;
    o->next_ticket = tk + 1;
    if (ticket) *ticket = tk;
    return 0;
}
int apj_obs_fetch_ring(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long first, long long count, double* h_out) {
    if (count < 0 || first < 0 || first + count > o->next_ticket || count > o->ring_cap || first < o->next_ticket - o->ring_cap) return -1;
    const size_t row = 2 * (size_t)st.n_sys;
    for (long long done = 0; done < count;) {          // at most two pieces (ring wrap)
        const long long slot = (first + done) % o->ring_cap;
        const long long n = std::min<long long>(count - done, o->ring_cap - slot);
        if (cudaMemcpyAsync(h_out + row * done, o->d_ring + row * slot, sizeof(double) * row * n, cudaMemcpyDeviceToHost, s) != cudaSuccess) return APJ_E_CUDA_OBS;
        done += n;
    }
    if (cudaStreamSynchronize(s) != cudaSuccess) return APJ_E_CUDA_OBS;
    return 0;
}

int apj_obs_com(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, double* com2) {
    if (int rc = reduce2(o, st, s, launches, OBS_COM, nullptr, com2)) return rc;
    for (int k = 0; k < 2 * st.n_sys; k++) com2[k] /= st.N;
    return 0;
}
int apj_obs_order(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, double* order, double* orient2) {
    std::vector<double> t(2 * st.n_sys);
    if (int rc = reduce2(o, st, s, launches, OBS_ORDER, nullptr, t.data())) return rc;
    for (int k = 0; k < st.n_sys; k++) {
        const double ox = t[2 * k], oy = t[2 * k + 1];
        if (order) order[k] = std::sqrt(ox * ox + oy * oy + 0.0 * 0.0) / (double)st.N;   // jamming.cpp:788
        if (orient2) { orient2[2 * k] = ox / (double)st.N; orient2[2 * k + 1] = oy / (double)st.N; }   // :803
    }
    return 0;
}
int apj_obs_msd(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, double* msd) {
    std::vector<double> t(2 * st.n_sys);
    if (int rc = reduce2(o, st, s, launches, OBS_MSD, nullptr, t.data())) return rc;
    for (int k = 0; k < st.n_sys; k++) msd[k] = t[2 * k] / st.N;   // jamming.cpp:822
    return 0;
}
int apj_obs_fluct(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, const double* radius, double* area) {
    std::vector<double> t(2 * st.n_sys);
    if (int rc = reduce2(o, st, s, launches, OBS_FLUCT, radius, t.data())) return rc;
    for (int k = 0; k < st.n_sys; k++) area[k] = t[2 * k];
    return 0;
}
int apj_obs_velhist(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, const double* dv, int64_t* hist100) {
    if (cudaMemcpyAsync(o->d_param, dv, sizeof(double) * st.n_sys, cudaMemcpyHostToDevice, s) != cudaSuccess) return APJ_E_CUDA_OBS;
    cudaMemsetAsync(o->d_hist, 0, sizeof(unsigned long long) * 128 * st.n_sys, s);
    apj_velhist_kernel<<<st.n_sys * o->blocks_per_sys, OB, 0, s>>>(st, o->blocks_per_sys, o->d_param, o->d_hist);
    if (launches) *launches += 1;
    std::vector<unsigned long long> h(128 * (size_t)st.n_sys);
    if (cudaMemcpyAsync(h.data(), o->d_hist, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, s) != cudaSuccess) return APJ_E_CUDA_OBS;
    if (cudaStreamSynchronize(s) != cudaSuccess) return APJ_E_CUDA_OBS;
    for (int k = 0; k < st.n_sys; k++) for (int q = 0; q < 100; q++) hist100[100 * k + q] = (int64_t)h[128 * (size_t)k + q];
    return 0;
}
int apj_obs_occupancy(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, int64_t* hist50) {
    cudaMemsetAsync(o->d_hist, 0, sizeof(unsigned long long) * 128 * st.n_sys, s);
    dim3 grid(64, st.n_sys);
    apj_occupancy_kernel<<<grid, OB, 0, s>>>(st, o->d_hist);
    if (launches) *launches += 1;
    std::vector<unsigned long long> h(128 * (size_t)st.n_sys);
    if (cudaMemcpyAsync(h.data(), o->d_hist, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, s) != cudaSuccess) return APJ_E_CUDA_OBS;
    if (cudaStreamSynchronize(s) != cudaSuccess) return APJ_E_CUDA_OBS;
    for (int k = 0; k < st.n_sys; k++) for (int q = 0; q < 50; q++) hist50[50 * k + q] = (int64_t)h[128 * (size_t)k + q];
    return 0;
}
int apj_obs_export_edge(const DevState& st, cudaStream_t s, long long* launches, double width, long long cap, double* h_out6, long long* n_found) {
    double* d = nullptr; int* dc = nullptr;
    if (cudaMalloc(&d, sizeof(double) * 6 * (size_t)std::max<long long>(cap, 1)) != cudaSuccess) return APJ_E_CUDA_OBS;
    if (cudaMalloc(&dc, sizeof(int)) != cudaSuccess) { cudaFree(d); return APJ_E_CUDA_OBS; }
    cudaMemsetAsync(dc, 0, sizeof(int), s);
    apj_export_edge_kernel<<<148 * 4, OB, 0, s>>>(st, width, cap, d, dc);
    if (launches) *launches += 1;
    int n = 0;
    bool ok = cudaMemcpyAsync(&n, dc, sizeof(int), cudaMemcpyDeviceToHost, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess;
    if (ok && n <= cap && n > 0 && h_out6) {
        for (int k = 0; k < 6 && ok; k++)          // compact the planes from stride cap to stride n
            ok = cudaMemcpyAsync(h_out6 + (size_t)k * n, d + (size_t)k * cap, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
    }
    cudaFree(d); cudaFree(dc);
    *n_found = n;
    return ok ? 0 : APJ_E_CUDA_OBS;
}

int apj_obs_spatial_ext(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, double cutoff, long long n_ext,
                        const double* h_ext6, const int* h_ext_row, int b, double* d_acc_row) {
    (void)o;
    if (n_ext <= 0) return 0;
    const int np = (int)std::ceil(cutoff / 0.1), nc = (int)std::ceil(cutoff / 2.0);
    const size_t nb = 3 * (size_t)nc + np;
    double* d = nullptr; int* dr = nullptr;
    if (cudaMalloc(&d, sizeof(double) * 6 * (size_t)n_ext) != cudaSuccess) return APJ_E_CUDA_OBS;
    if (cudaMalloc(&dr, sizeof(int) * (size_t)(b + 1)) != cudaSuccess) { cudaFree(d); return APJ_E_CUDA_OBS; }
    bool ok = cudaMemcpyAsync(d, h_ext6, sizeof(double) * 6 * (size_t)n_ext, cudaMemcpyHostToDevice, s) == cudaSuccess &&
              cudaMemcpyAsync(dr, h_ext_row, sizeof(int) * (size_t)(b + 1), cudaMemcpyHostToDevice, s) == cudaSuccess;
    if (ok) {
        const size_t smem = nb * sizeof(double);
        if (smem > 48 * 1024) apj_allow_max_smem(apj_spatial_ext_kernel);
        apj_spatial_ext_kernel<<<(unsigned)((st.cap + OB - 1) / OB), OB, smem, s>>>(st, nc, np, n_ext, d, dr, d_acc_row);
        if (launches) *launches += 1;
        ok = cudaStreamSynchronize(s) == cudaSuccess;
    }
    cudaFree(d); cudaFree(dr);
    return ok ? 0 : APJ_E_CUDA_OBS;
}

int apj_obs_spatial(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, const SysCtl* hctl, double cutoff,
                    double* counts, double* ori, double* vel, double* pair, std::vector<void*>& allocs,
                    long long n_ext, const double* h_ext6, const int* h_ext_row) {
    (void)hctl;
    const int np = (int)std::ceil(cutoff / 0.1), nc = (int)std::ceil(cutoff / 2.0);   // Correlations.h:55-56
    const size_t nb = 3 * (size_t)nc + np, need = nb * st.n_sys;
    if (nb * sizeof(double) > 200 * 1024) return APJ_E_CUDA_OBS;
    if (need > o->corr_cap) {
        void* p = nullptr;
        if (cudaMalloc(&p, need * sizeof(double)) != cudaSuccess) return APJ_E_CUDA_OBS;
        if (o->d_corr) {   // replace the smaller buffer (work queued on it has to drain first)
            cudaStreamSynchronize(s);
            allocs.erase(std::remove(allocs.begin(), allocs.end(), (void*)o->d_corr), allocs.end());
            cudaFree(o->d_corr);
        }
        allocs.push_back(p);
        o->d_corr = (double*)p; o->corr_cap = need;
    }
    cudaMemsetAsync(o->d_corr, 0, need * sizeof(double), s);
    const int bps = (st.cap + OB - 1) / OB;
    const bool priv = nc <= 16;                               // per-thread copies of the wide bins fit shared memory
    const size_t smem = (((nb + 1) & ~(size_t)1) + (priv ? (size_t)3 * nc * OB : 0)) * sizeof(double);
    if (priv) {
        if (smem > 48 * 1024) apj_allow_max_smem(apj_spatial_kernel<true>);
        apj_spatial_kernel<true><<<st.n_sys * bps, OB, smem, s>>>(st, bps, nc, np, 0, o->d_corr);
    } else {
        if (smem > 48 * 1024) apj_allow_max_smem(apj_spatial_kernel<false>);
        apj_spatial_kernel<false><<<st.n_sys * bps, OB, smem, s>>>(st, bps, nc, np, 0, o->d_corr);
    }
    if (launches) *launches += 1;
    if (n_ext > 0) if (int rc = apj_obs_spatial_ext(o, st, s, launches, cutoff, n_ext, h_ext6, h_ext_row, hctl[0].b, o->d_corr)) return rc;
    std::vector<double> h(need);
    if (cudaMemcpyAsync(h.data(), o->d_corr, need * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess) return APJ_E_CUDA_OBS;
    if (cudaStreamSynchronize(s) != cudaSuccess) return APJ_E_CUDA_OBS;
    for (int k = 0; k < st.n_sys; k++) {
        const double* r = h.data() + nb * k;
        for (int q = 0; q < nc; q++) { counts[(size_t)k * nc + q] = r[q]; ori[(size_t)k * nc + q] = r[nc + q]; vel[(size_t)k * nc + q] = r[2 * nc + q]; }
        for (int q = 0; q < np; q++) pair[(size_t)k * np + q] = r[3 * nc + q];
    }
    return 0;
}
