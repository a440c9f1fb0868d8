// Host side of the C ABI (include/apj_b200.h): owns device memory, the stream, the step graph
// and the per-system control blocks. No physics runs on the host; the only host arithmetic is
// setup that the reference also does once outside the hot loop (Rinv, cos/sin of the uploaded
// phi, the initial COM sum in index order -- jamming.cpp:298, :332-333, :761-774).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>
#include <vector>

#include "../../include/apj_b200.h"
#include "apj_device.cuh"
#include "apj_observe.cuh"

static thread_local std::string g_create_error;

struct apj_engine {
    apj_config cfg{};
    DevState st{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaGraphExec_t group_exec = nullptr;
    int m = 16;                 // step launches per group
    int kernels_per_group = 0;
    // graphs of the SHORTER groups a caller keeps asking for (a driver stepping 10 at a time between two
    // measurements, jamming.cpp:573): captured the second time the same remainder is seen
    std::map<int, cudaGraphExec_t> part_exec;
    std::map<int, int> part_seen, part_kernels;
    int max_nbox = 0;
    int max_b = 0;
    long long total_cols = 0;   // sum of b+1
    long long total_cells = 0;  // sum of nbox+1
    long long launches = 0;
    bool have_state = false;
    std::vector<SysCtl> hctl;
    bool ctl_clean = false;      // hctl equals the device control blocks (nothing that writes them was launched since the last copy)
    SysCtl* pin_ctl = nullptr;   // pinned bounce buffer of hctl: control-block copies never take the driver's pageable staging path
    std::vector<void*> allocs;
    double* d_noise = nullptr;
    // host <-> device state hand-over (apj_state_io.cu): staging planes in HBM, grown on demand and kept
    void* stage_mem = nullptr;
    size_t stage_n = 0;
    ApjStage stage{};
    int* pin_small = nullptr;    // pinned scratch for small read-backs (flags, checksums)
    ApjObsScratch obs{};
    std::string err;
    // slab mode
    size_t arena_bytes = 0;
    bool peer_set[APJ_MAX_RANKS] = {};
    bool slab_ready = false;
    std::vector<void*> ipc_opened;
};

#define APJ_CUDA(e, call)                                                                   \
    do {                                                                                    \
        cudaError_t _s = (call);                                                            \
        if (_s != cudaSuccess) {                                                            \
            char _b[512];                                                                   \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_s), __FILE__, __LINE__); \
            (e)->err = _b;                                                                  \
            return APJ_E_CUDA;                                                              \
        }                                                                                   \
    } while (0)

static int fail(apj_engine* e, int code, const std::string& msg) {
    if (e) e->err = msg; else g_create_error = msg;
    return code;
}

extern "C" const char* apj_version(void) { return "apj_b200 0.1 (sm_100a)"; }
extern "C" const char* apj_last_error(const apj_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

template <class T>
static int dev_alloc(apj_engine* e, T** p, size_t count) {
    void* q = nullptr;
    APJ_CUDA(e, cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    APJ_CUDA(e, cudaMemsetAsync(q, 0, std::max<size_t>(count, 1) * sizeof(T), e->stream));
    e->allocs.push_back(q);
    *p = (T*)q;
    return APJ_OK;
}

// Control-block transfers go through PINNED host memory. A pageable cudaMemcpyAsync is staged by the
// driver and can wait on work of OTHER streams; with several slab ranks on one device (tests) a rank
// whose kernel spins on a peer would then block that peer's control copy: a deadlock until the timeout.
static int pull_ctl(apj_engine* e) {
    const size_t bytes = sizeof(SysCtl) * e->hctl.size();
    if (!e->pin_ctl) APJ_CUDA(e, cudaMallocHost(&e->pin_ctl, bytes));
    APJ_CUDA(e, cudaMemcpyAsync(e->pin_ctl, e->st.ctl, bytes, cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    memcpy(e->hctl.data(), e->pin_ctl, bytes);
    e->ctl_clean = true;
    return APJ_OK;
}
// Read-only getters: no device round trip when nothing that writes the control blocks was launched since the
// last copy (a driver that reads COM and counters of every replica after apj_step pays for ONE copy, the one
// apj_step made). Slab ranks always copy: peers write their flags.
static int read_ctl(apj_engine* e) {
    if (e->ctl_clean && !e->st.slab) return APJ_OK;
    return pull_ctl(e);
}
static int push_ctl(apj_engine* e) {
    const size_t bytes = sizeof(SysCtl) * e->hctl.size();
    if (!e->pin_ctl) APJ_CUDA(e, cudaMallocHost(&e->pin_ctl, bytes));
    memcpy(e->pin_ctl, e->hctl.data(), bytes);
    e->ctl_clean = false;
    APJ_CUDA(e, cudaMemcpyAsync(e->st.ctl, e->pin_ctl, bytes, cudaMemcpyHostToDevice, e->stream));
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    e->ctl_clean = true;
    return APJ_OK;
}

__global__ void apj_add_target_kernel(SysCtl* ctl, int n_sys, long long n) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_sys) ctl[s].target = ctl[s].step + n;
}

static ApjLaunch launcher(apj_engine* e, bool count) { e->ctl_clean = false; return ApjLaunch{e->stream, count ? &e->launches : nullptr}; }

static void drop_group_graphs(apj_engine* e) {
    if (e->group_exec) { cudaGraphExecDestroy(e->group_exec); e->group_exec = nullptr; }
    for (auto& kv : e->part_exec) cudaGraphExecDestroy(kv.second);
    e->part_exec.clear(); e->part_kernels.clear();
}

// `nsteps` speculative steps followed by the rebuild chain, as one graph
static int capture_group(apj_engine* e, int nsteps, cudaGraphExec_t* exec, int* kernels) {
    cudaGraph_t graph = nullptr;
    long long dummy = 0;
    ApjLaunch l{e->stream, &dummy};
    APJ_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    for (int k = 0; k < nsteps; k++) apj_launch_step(e->st, l, nullptr, 0);
    apj_launch_rebuild_chain(e->st, l, e->max_nbox, e->max_b);
    APJ_CUDA(e, cudaStreamEndCapture(e->stream, &graph));
    APJ_CUDA(e, cudaGraphInstantiate(exec, graph, 0));
    APJ_CUDA(e, cudaGraphDestroy(graph));
    *kernels = (int)dummy;
    return APJ_OK;
}

static int build_group_graph(apj_engine* e) {
    drop_group_graphs(e);
    if (e->cfg.flags & APJ_FLAG_NO_GRAPH) return APJ_OK;
    if (e->st.slab && e->st.nranks > 1 && !e->slab_ready) return APJ_OK;   // peer addresses not known yet
    return capture_group(e, e->m, &e->group_exec, &e->kernels_per_group);
}

// a group of `rem` < m steps: direct launches the first time, a graph of its own from the second time on
static int launch_part(apj_engine* e, int rem) {
    e->ctl_clean = false;
    // slab ranks launch directly: capturing and instantiating a graph in the middle of a run can wait on the device,
    // and ranks that share ONE device (tests) would then stall the peers spinning on them
    if (e->group_exec && rem > 0 && !(e->st.slab && e->st.nranks > 1)) {
        auto it = e->part_exec.find(rem);
        if (it == e->part_exec.end() && e->part_seen[rem]++ >= 1) {
            cudaGraphExec_t x = nullptr;
            int kn = 0;
            if (int rc = capture_group(e, rem, &x, &kn)) return rc;
            e->part_kernels[rem] = kn;
            it = e->part_exec.emplace(rem, x).first;
        }
        if (it != e->part_exec.end()) {
            APJ_CUDA(e, cudaGraphLaunch(it->second, e->stream));
            e->launches += e->part_kernels[rem];
            return APJ_OK;
        }
    }
    ApjLaunch l = launcher(e, true);
    for (int k = 0; k < rem; k++) apj_launch_step(e->st, l, nullptr, 0);
    apj_launch_rebuild_chain(e->st, l, e->max_nbox, e->max_b);
    APJ_CUDA(e, cudaGetLastError());
    return APJ_OK;
}

static int launch_group(apj_engine* e) {
    e->ctl_clean = false;
    if (e->group_exec) {
        APJ_CUDA(e, cudaGraphLaunch(e->group_exec, e->stream));
        e->launches += e->kernels_per_group;
    } else {
        ApjLaunch l = launcher(e, true);
        for (int k = 0; k < e->m; k++) apj_launch_step(e->st, l, nullptr, 0);
        apj_launch_rebuild_chain(e->st, l, e->max_nbox, e->max_b);
        APJ_CUDA(e, cudaGetLastError());
    }
    return APJ_OK;
}

static int build_group_graph(apj_engine* e);

// Shared-memory tile capacity follows the decomposition. The capacity decides how many work
// blocks fit one SM (227 KB of shared memory, 1 KB reserved per block), so it is set to the LARGEST
// value that still reaches the block count the current tiles allow: full occupancy, maximum slack
// before a denser tile forces a re-launch. The launch configuration (and the graph) is rebuilt
// when a rebuild produced a tile larger than the capacity (nothing was committed, the system is
// still stale) or when a smaller capacity would fit one more block per SM.
static int blocks_per_sm(const DevState& st, int cap) {
    if (apj_ring_eligible(st) && apj_ring_buffers(st.tb, cap) > 0) return apj_ring_buffers(st.tb, cap);   // ring kernel: tile buffers of the one block of an SM
    // dynamic + static shared memory of a step block + the 1 KB the driver reserves per resident block
    const size_t per_block = (size_t)(cap + 1) * 48 + (st.G > 1 ? (size_t)st.ppb * 32 : 0) + apj_step_extra_smem(st) + 768 + 1024;
    const int by_smem = (int)((size_t)233472 / per_block);
    return std::max(1, std::min(by_smem, apj_step_blocks_per_sm_limit(st.tb)));   // register file: __launch_bounds__ of the step kernel
}
static int best_tile_cap(const DevState& st, int need) {
    // slack before a denser tile forces a re-launch: 3 % + 8 slots (APJ_TILE_SLACK=0, tuning runs: none)
    static const int slack = getenv("APJ_TILE_SLACK") ? atoi(getenv("APJ_TILE_SLACK")) : 1;
    if (apj_ring_eligible(st)) {                                      // ring kernel: the largest capacity its tile buffers allow, if the tiles fit it
        const int nb = st.tb == 256 ? 5 : 8;
        const int ring_cap = (int)((size_t)(232448 - 2560) / ((size_t)nb * 48)) - 1;
        if (need <= ring_cap) return ring_cap;
    }
    need = std::min(std::max(slack ? need + need / 32 + 8 : need, 64), 4094);
    const int target = blocks_per_sm(st, need);
    int lo = need, hi = 4094;                                         // largest cap with the same block count
    while (lo < hi) {
        const int mid = (lo + hi + 1) / 2;
        if (blocks_per_sm(st, mid) >= target) lo = mid; else hi = mid - 1;
    }
    return lo;
}
static int set_tile_cap(apj_engine* e, int need) {
    if (need > 4094) return fail(e, APJ_E_OVERFLOW, "a work block needs more than 4094 shared-memory slots (density too inhomogeneous for the tile)");
    e->st.tile_cap = e->cfg.tile_slots > 0 ? std::max(e->cfg.tile_slots, need) : best_tile_cap(e->st, need);
    if (apj_configure_kernels(e->st) != 0 || apj_configure_rebuild(e->st) != 0) return fail(e, APJ_E_CUDA, "cannot reserve shared memory for the tile");
    if ((e->cfg.flags & APJ_FLAG_TINY_GRID) && e->st.persist_grid > 3) e->st.persist_grid = 3;
    return build_group_graph(e);
}
// after pull_ctl: returns 1 if a tile overflow was repaired (caller must run the chain again)
static int repair_tile_overflow(apj_engine* e, int* repaired) {
    *repaired = 0;
    if (e->st.slab) return APJ_OK;   // ranks launch in lockstep: no unilateral re-launch; reported by check_overflow
    int need = 0;
    for (auto& c : e->hctl) if (c.overflow & 2) need = std::max(need, c.tile_max);
    if (!need) return APJ_OK;
    if (int rc = set_tile_cap(e, need)) return rc;
    for (auto& c : e->hctl) c.overflow &= ~2;
    *repaired = 1;
    return push_ctl(e);
}
// A Verlet list longer than the capacity S leaves the system stale (apj_finish_rebuild_kernel commits nothing), so
// no step ever runs on truncated lists: the list storage is re-allocated for the longest list seen (+25 %) and the
// chain is run again. The reference's std::vector lists are unbounded (jamming.cpp:550-585); the bound here is
// the 96 entries the 16-bit tile-slot encoding and the register-held quads are built for. Periodic handles only:
// slab ranks launch in lockstep, there the overflow is reported (check_overflow) and the caller re-creates.
static void set_list_capacity(DevState& st, int S) {
    st.S = S;
    st.max_rounds = (st.S / 2 + st.G - 1) / st.G;
    st.max_quads = (st.max_rounds + 3) / 4;
}
static int repair_list_overflow(apj_engine* e, int* repaired) {
    if (e->st.slab) return APJ_OK;
    int need = 0;
    for (auto& c : e->hctl) if (c.overflow & 1) need = std::max(need, c.list_max);
    if (!need) return APJ_OK;
    if (need > apj_max_list_capacity()) return APJ_OK;   // beyond what the layout holds: check_overflow reports it
    DevState& st = e->st;
    const int S = std::min(apj_max_list_capacity(), (need + need / 4 + 3) / 2 * 2);
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    set_list_capacity(st, S);
    unsigned* fresh = nullptr;
    const size_t words = (size_t)st.n_sys * st.maxblk * st.max_quads * 4 * st.tb;
    APJ_CUDA(e, cudaMalloc(&fresh, words * sizeof(unsigned)));
    APJ_CUDA(e, cudaMemsetAsync(fresh, 0, words * sizeof(unsigned), e->stream));
    e->allocs.erase(std::remove(e->allocs.begin(), e->allocs.end(), (void*)st.list32), e->allocs.end());
    cudaFree(st.list32);
    st.list32 = fresh;
    e->allocs.push_back(fresh);
    for (auto& c : e->hctl) { c.overflow &= ~1; c.list_max = 0; }
    if (int rc = push_ctl(e)) return rc;
    *repaired = 1;
    return build_group_graph(e);                          // kernels take DevState (S, max_quads, list32) by value
}
static int maybe_shrink_tile_cap(apj_engine* e) {
    if (e->cfg.tile_slots > 0) return APJ_OK;
    int need = 0;
    for (auto& c : e->hctl) need = std::max(need, c.tile_max);
    if (need > 0 && blocks_per_sm(e->st, best_tile_cap(e->st, need)) > blocks_per_sm(e->st, e->st.tile_cap)) return set_tile_cap(e, need);
    return APJ_OK;
}

static int check_overflow(apj_engine* e) {
    if (e->st.slab && e->hctl[0].slab_err) {
        const int f = e->hctl[0].slab_err;
        char b[480], w[160] = "";
        const int dg = e->hctl[0].slab_diag;
        if (f & 1) snprintf(w, sizeof w, " (first wait that timed out: %s, missing ranks mask 0x%x)",
                            (dg >> 16) == 1 ? "step partials" : ((dg >> 16) == 2 ? "rebuild, migrants delivered" : "rebuild, ghost columns"), dg & 0xffff);
        snprintf(b, sizeof b, "slab rank %d/%d:%s%s%s%s", e->st.rank, e->st.nranks,
                 (f & 1) ? " a peer rank did not arrive within the timeout" : "", w,
                 (f & 2) ? " a particle crossed more than one slab between rebuilds;" : "",
                 (f & 4) ? " particle / ghost-column / migration capacity exceeded (recreate with a larger capacity);" : "");
        return fail(e, (f & 4) ? APJ_E_OVERFLOW : APJ_E_STATE, b);
    }
    if (e->st.slab && (e->hctl[0].overflow & 2))
        return fail(e, APJ_E_OVERFLOW, "slab mode: a work block needs a larger shared-memory tile than configured (set tile_slots)");
    for (size_t s = 0; s < e->hctl.size(); s++)
        if (e->hctl[s].overflow & 1) {
            char b[256];
            snprintf(b, sizeof b, "Verlet list overflow in system %zu: %d neighbours within rs > max_neighbors=%d; recreate with a larger max_neighbors",
                     s, e->hctl[s].list_max, e->st.S);
            return fail(e, APJ_E_OVERFLOW, b);
        }
    return APJ_OK;
}

struct SlabSpec { int on = 0, rank = 0, nranks = 1; long long cap = 0; };

// carve `count` objects out of the arena (256-byte aligned); with base == nullptr only the size is computed
template <class T>
static void arena_take(char* base, size_t& off, T** p, size_t count) {
    off = (off + 255) / 256 * 256;
    if (base && p) *p = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
}
static size_t arena_layout(DevState& st, char* base, int b) {
    size_t off = 0;
    const size_t n = (size_t)st.cap + 2 * (size_t)st.gcap;
    for (int h = 0; h < 2; h++) {
        arena_take(base, off, &st.XY[h], n); arena_take(base, off, &st.CS[h], n);
        arena_take(base, off, &st.RR[h], n); arena_take(base, off, &st.ID[h], n);
    }
    arena_take(base, off, &st.gstart, (size_t)4 * (b + 1));
    arena_take(base, off, &st.mail, 1);
    arena_take(base, off, &st.inbox, (size_t)st.mcap);
    return (off + 255) / 256 * 256;
}

static int create_impl(const apj_config* cfg, const double* L, const SlabSpec& slab, apj_engine** out) {
    if (!cfg || !L || !out) return fail(nullptr, APJ_E_INVALID, "apj_create: null argument");
    if (cfg->n < 1 || cfg->n_systems < 1) return fail(nullptr, APJ_E_INVALID, "apj_create: n and n_systems must be >= 1");
    if ((long long)cfg->n * cfg->n_systems > 0x7fffffffLL) return fail(nullptr, APJ_E_INVALID, "apj_create: n*n_systems exceeds 2^31-1");
    if (slab.on && (cfg->n_systems != 1 || slab.nranks < 1 || slab.nranks > APJ_MAX_RANKS || slab.rank < 0 || slab.rank >= slab.nranks))
        return fail(nullptr, APJ_E_INVALID, "apj_slab_create: one system, 1 <= nranks <= 8, 0 <= rank < nranks");
    apj_engine* e = new apj_engine;
    e->cfg = *cfg;
    auto bail = [&](int code) { g_create_error = e->err; apj_destroy(e); return code; };
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        e->err = "apj_create: no CUDA device (this library has no CPU fallback)";
        return bail(APJ_E_CUDA);
    }
    if (cudaSetDevice(cfg->device) != cudaSuccess) { e->err = "apj_create: cudaSetDevice failed"; return bail(APJ_E_CUDA); }
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { e->err = "apj_create: stream"; return bail(APJ_E_CUDA); }
    cudaEventCreate(&e->ev0); cudaEventCreate(&e->ev1);

    DevState& st = e->st;
    st.n_sys = cfg->n_systems; st.N = (int)cfg->n;
    st.slab = slab.on; st.rank = slab.rank; st.nranks = slab.nranks;
    st.left = (slab.rank + slab.nranks - 1) % slab.nranks; st.right = (slab.rank + 1) % slab.nranks;
    st.cap = st.N;
    st.timeout_ns = 20000000000ull;
    if (slab.on) {
        const double rn0 = cfg->rn > 0 ? cfg->rn : 2.8;
        const int b0 = std::max(3, (int)floor(L[0] / (2 * rn0)));
        const long long col = st.N / b0 + 1;                       // mean population of one cell column
        // capacity of this rank: its share + 25 % + four columns; ghost columns / inbox: 1.5 columns + slack
        long long cap = slab.cap > 0 ? slab.cap : (slab.nranks == 1 ? (long long)st.N : (long long)(1.25 * st.N / slab.nranks) + 4 * col + 1024);
        st.cap = (int)std::min<long long>(cap, st.N);
        st.gcap = (int)(col + col / 2 + 256);
        st.mcap = st.gcap;
    }
    st.ntot = (long long)st.n_sys * st.cap;
    st.S = cfg->max_neighbors > 0 ? (cfg->max_neighbors + 1) / 2 * 2 : 48;   // grown on demand (repair_list_overflow)
    if (st.S > apj_max_list_capacity()) { e->err = "apj_create: max_neighbors too large (<= 96)"; return bail(APJ_E_INVALID); }
    // lanes per particle: spread small systems over enough warps to hide latency (measured on B200)
    st.G = cfg->lanes_per_particle;
    if (const char* le = getenv("APJ_LANES")) if (st.G == 0) st.G = atoi(le);   // tuning runs
    // measured on B200 (ms per step, lanes 1 / 2 / 4): N = 16 384: .0129 / .0128 / .0111; 65 536: .0158 / .0139 / .0188;
    // 131 072: .0221 / .0209 / .0287; 262 144: .0359 / .0432 / .0569; 524 288: .0470 / .0648 / .0992
    if (st.G == 0) st.G = st.ntot < 40000 ? 4 : (st.ntot < 200000 ? 2 : 1);
    if (st.G != 1 && st.G != 2 && st.G != 4 && st.G != 8) { e->err = "apj_create: lanes_per_particle must be 1, 2, 4 or 8"; return bail(APJ_E_INVALID); }
    st.tb = st.G == 1 ? APJ_TB_G1 : 128;
    if (const char* te = getenv("APJ_TB")) if (st.G == 1 && (atoi(te) == 128 || atoi(te) == 192 || atoi(te) == 256)) st.tb = atoi(te);   // tuning runs
    st.ppb = st.tb / st.G;
    set_list_capacity(st, st.S);
    // Engine constants exactly as the reference derives them (jamming.cpp:57, :112-115, :611)
    const double dt = cfg->dt > 0 ? cfg->dt : 0.1;
    const double rn = cfg->rn > 0 ? cfg->rn : 2.8;
    const double rs = (cfg->rs_factor > 0 ? cfg->rs_factor : 1.5) * rn;
    st.dt = dt; st.rn2 = rn * rn; st.rs2 = rs * rs; st.skin = rs - rn;
    st.rn = rn; st.cls_inv = APJ_CLASSES / (st.rs2 - st.rn2);
#ifdef APJ_NO_TRUNCATE
    st.truncate = 0;
#else
    st.truncate = 1;
#endif
    st.seed = cfg->seed;
    e->m = cfg->steps_per_launch > 0 ? cfg->steps_per_launch : 16;
    if (const char* me = getenv("APJ_GROUP")) if (cfg->steps_per_launch <= 0 && atoi(me) > 0) e->m = atoi(me);   // tuning runs

    e->hctl.assign(st.n_sys, SysCtl{});
    long long cells = 0, cols = 0;
    for (int s = 0; s < st.n_sys; s++) {
        SysCtl& c = e->hctl[s];
        c.L = L[s]; c.Lover2 = L[s] / 2.0;                  // jamming.cpp:306
        double lp = 2 * rn;                                  // Engine::topology, jamming.cpp:361-365
        c.b = static_cast<int>(floor(c.L / lp));
        if (c.b < 3) { e->err = "apj_create: box too small (b = floor(L/(2 rn)) < 3 aliases neighbour cells, jamming.cpp:359)"; return bail(APJ_E_INVALID); }
        if ((long long)c.b * c.b > 0x7ffffff0LL) { e->err = "apj_create: too many cells"; return bail(APJ_E_INVALID); }
        c.nbox = c.b * c.b;
        c.lp = c.L / floor(c.L / lp);
        c.p0 = s * st.N; c.n_own = st.N; c.col0 = 0; c.ncols = c.b;
        if (slab.on) {   // whole cell columns per rank: [rank*b/nranks, (rank+1)*b/nranks)
            c.col0 = (int)((long long)slab.rank * c.b / slab.nranks);
            c.ncols = (int)((long long)(slab.rank + 1) * c.b / slab.nranks) - c.col0;
            if (c.ncols < 1) { e->err = "apj_slab_create: fewer cell columns than ranks"; return bail(APJ_E_INVALID); }
            c.nbox = c.ncols * c.b;     // cells this rank owns
            c.n_own = 0;
        }
        c.cell_base = (int)cells;
        cells += c.nbox + 1;
        c.col_base = (int)cols;
        cols += c.b + 1;
        e->max_nbox = std::max(e->max_nbox, c.nbox);
        e->max_b = std::max(e->max_b, c.b);
        c.stale = 0;
    }
    e->total_cells = cells;
    e->total_cols = cols;
    st.maxblk = st.cap / st.ppb + e->max_b + 1;
    st.maxgrp = (st.maxblk + 31) / 32;
    // Large systems end the step with a separate fold + commit kernel (apj_step.cu); small ones (phase-diagram
    // replicas, tests) are launch-bound and keep the commit fused into the step kernel's last block.
#ifndef APJ_SPLIT_MIN_BLOCKS
#define APJ_SPLIT_MIN_BLOCKS 4096
#endif
    st.split_tail = (st.G == 1 && (long long)st.n_sys * st.maxblk >= APJ_SPLIT_MIN_BLOCKS && st.maxblk <= 1024 * 1024) ? 1 : 0;
    if (st.G == 1 && (cfg->flags & APJ_FLAG_SPLIT_TAIL) && st.maxblk <= 1024 * 1024) st.split_tail = 1;
    if (cfg->flags & APJ_FLAG_FUSED_TAIL) st.split_tail = 0;
    // launch-bound sizes: groups of 32 halve the share of the idle rebuild chain that ends every group (measured on one box, 16 / 32 / 64:
    // jam1k 9.84e7 / 1.011e8 / 1.021e8, jam65k 4.80e9 / 4.90e9 / 4.89e9, sweep512 7.09e9 / 7.42e9 / 7.37e9); large systems
    // pay for every launch that idles after a rebuild fired and stay at 16
    if (cfg->steps_per_launch <= 0 && !getenv("APJ_GROUP") && (long long)st.n_sys * st.maxblk < APJ_SPLIT_MIN_BLOCKS) e->m = 32;
    st.want_persist = (cfg->flags & APJ_FLAG_PERSIST) ? 1 : 0;
    if (const char* pe = getenv("APJ_STEP_PIPE")) st.want_persist = atoi(pe) ? 1 : 0;   // tuning runs: force the pipelined kernel on / off
    st.want_ring = 0;
    if (const char* re = getenv("APJ_STEP_RING")) st.want_ring = atoi(re) ? 1 : 0;      // ring kernel (large periodic single systems)
    {   // tile capacity: three columns x (block rows + 2 halo rows), sized from the mean cell occupancy
        double ppc = 0;
        for (int s = 0; s < st.n_sys; s++) ppc = std::max(ppc, (double)st.N / ((double)e->hctl[s].b * e->hctl[s].b));
        const double mean = 3.0 * (st.ppb / ppc + 3.0) * ppc;   // three columns x (block rows + halo rows)
        const int want = std::min((int)(mean + 4.0 * std::sqrt(mean)) + 16, 4094);
        st.tile_cap = cfg->tile_slots > 0 ? std::min(cfg->tile_slots, 4094) : best_tile_cap(st, want);   // follows ctl.tile_max later
    }

    int rc = APJ_OK;
#define A(x) if (rc == APJ_OK) rc = (x)
    A(dev_alloc(e, &st.ctl, (size_t)st.n_sys));
    if (slab.on) {   // everything a peer writes: one allocation, same layout on every rank
        e->arena_bytes = arena_layout(st, nullptr, e->max_b);
        A(dev_alloc(e, &st.arena, e->arena_bytes));
        if (rc == APJ_OK) arena_layout(st, st.arena, e->max_b);
        st.peer_arena[st.rank] = st.arena;
        e->peer_set[st.rank] = true;
    }
    for (int h = 0; h < 2; h++) {
        if (!slab.on) {
            A(dev_alloc(e, &st.XY[h], (size_t)st.ntot)); A(dev_alloc(e, &st.CS[h], (size_t)st.ntot));
            A(dev_alloc(e, &st.RR[h], (size_t)st.ntot)); A(dev_alloc(e, &st.ID[h], (size_t)st.ntot));
        }
        A(dev_alloc(e, &st.XR[h], (size_t)st.ntot));
        A(dev_alloc(e, &st.X0[h], (size_t)st.ntot));
        A(dev_alloc(e, &st.XO[h], (size_t)st.ntot)); A(dev_alloc(e, &st.V[h], (size_t)st.ntot));
        A(dev_alloc(e, &st.PHI[h], (size_t)st.ntot));
        A(dev_alloc(e, &st.BOX[h], (size_t)st.ntot));
    }
    A(dev_alloc(e, &st.tiles, (size_t)st.n_sys * st.maxblk));
    A(dev_alloc(e, &st.list32, (size_t)st.n_sys * st.maxblk * st.max_quads * 4 * st.tb));
    A(dev_alloc(e, &st.col_blk, (size_t)cols));
    A(dev_alloc(e, &st.chunk_sums, (size_t)st.n_sys * ((e->max_nbox + apj_scan_chunk_cells() - 1) / apj_scan_chunk_cells())));
    A(dev_alloc(e, &st.cnt, (size_t)st.ntot));
    A(dev_alloc(e, &st.cntk, (size_t)st.ntot));
    A(dev_alloc(e, &st.boxnew, (size_t)st.ntot));
    A(dev_alloc(e, &st.perm, (size_t)st.ntot));
    A(dev_alloc(e, &st.cell_count, (size_t)cells));
    A(dev_alloc(e, &st.cell_start, (size_t)cells));
    A(dev_alloc(e, &st.cell_cursor, (size_t)cells));
    A(dev_alloc(e, &st.partials, (size_t)st.n_sys * st.maxblk));
    A(dev_alloc(e, &st.gpartials, (size_t)st.n_sys * st.maxgrp));
    st.wpartials = nullptr;
    if (st.want_ring) A(dev_alloc(e, &st.wpartials, (size_t)st.n_sys * st.maxblk * 8));
    A(dev_alloc(e, &st.gticket, (size_t)st.n_sys * st.maxgrp));
    A(dev_alloc(e, &e->d_noise, (size_t)st.n_sys * st.N));
    A(apj_obs_alloc(&e->obs, st, e->stream, e->allocs));
#undef A
    if (rc != APJ_OK) return bail(rc);
    if (apj_configure_kernels(st) != 0 || apj_configure_rebuild(st) != 0) { e->err = "apj_create: cannot reserve shared memory for the tile (tile_slots too large?)"; return bail(APJ_E_CUDA); }
    if ((cfg->flags & APJ_FLAG_TINY_GRID) && st.persist_grid > 3) st.persist_grid = 3;
    if (push_ctl(e) != APJ_OK) return bail(APJ_E_CUDA);
    if (!slab.on || slab.nranks == 1) {      // slab mode: the graph embeds peer addresses, built by apj_slab_ready
        if (build_group_graph(e) != APJ_OK) return bail(APJ_E_CUDA);
    }
    *out = e;
    return APJ_OK;
}

extern "C" int apj_create(const apj_config* cfg, const double* L, apj_engine** out) {
    return create_impl(cfg, L, SlabSpec{}, out);
}

extern "C" int apj_destroy(apj_engine* e) {
    if (!e) return APJ_OK;
    if (e->stream) cudaStreamSynchronize(e->stream);
    drop_group_graphs(e);
    for (void* p : e->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void* p : e->allocs) cudaFree(p);
    if (e->pin_ctl) cudaFreeHost(e->pin_ctl);
    if (e->pin_small) cudaFreeHost(e->pin_small);
    if (e->stage_mem) cudaFree(e->stage_mem);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return APJ_OK;
}

extern "C" int apj_set_activity(apj_engine* e, const double* CFself, const double* CTnoise) {
    if (!e) return APJ_E_INVALID;
    if (int rc = pull_ctl(e)) return rc;
    for (int s = 0; s < e->st.n_sys; s++) {
        if (CFself) e->hctl[s].CFself = CFself[s];
        if (CTnoise) e->hctl[s].CTnoise = CTnoise[s];
    }
    return push_ctl(e);
}

extern "C" int apj_set_ramp(apj_engine* e, int64_t tthermalize) {
    if (!e || tthermalize < 0) return APJ_E_INVALID;
    if (int rc = pull_ctl(e)) return rc;
    for (auto& c : e->hctl) { c.ramp_len = tthermalize; c.ramp_t0 = c.step; }
    return push_ctl(e);
}

extern "C" int apj_skip_self_term_once(apj_engine* e, int32_t on) {
    if (!e) return APJ_E_INVALID;
    if (int rc = pull_ctl(e)) return rc;
    for (auto& c : e->hctl) c.no_self_once = on ? 1 : 0;
    return push_ctl(e);
}

static int run_chain_now(apj_engine* e) {
    for (int attempt = 0; attempt < 8; attempt++) {
        apj_launch_rebuild_chain(e->st, launcher(e, true), e->max_nbox, e->max_b);
        APJ_CUDA(e, cudaGetLastError());
        if (int rc = pull_ctl(e)) return rc;
        int repaired = 0;
        if (int rc = repair_tile_overflow(e, &repaired)) return rc;
        if (int rc = repair_list_overflow(e, &repaired)) return rc;
        if (!repaired) break;
    }
    return check_overflow(e);
}

// ---- state hand-over: staging planes + pack / unpack kernels (apj_state_io.cu) ---------------------
static int ensure_stage(apj_engine* e, size_t n) {
    if (!e->pin_small) APJ_CUDA(e, cudaMallocHost(&e->pin_small, 256));
    n = std::max<size_t>(n, 1);
    if (e->stage_n >= n) return APJ_OK;
    if (e->stage_mem) { APJ_CUDA(e, cudaStreamSynchronize(e->stream)); cudaFree(e->stage_mem); e->stage_mem = nullptr; e->stage_n = 0; }
    const size_t plane = (n * sizeof(double) + 255) / 256 * 256, iplane = (n * sizeof(int) + 255) / 256 * 256;
    APJ_CUDA(e, cudaMalloc(&e->stage_mem, 14 * plane + 2 * iplane + 256));
    char* p = static_cast<char*>(e->stage_mem);
    for (int k = 0; k < 14; k++) e->stage.f[k] = reinterpret_cast<double*>(p + k * plane);
    e->stage.box = reinterpret_cast<int*>(p + 14 * plane);
    e->stage.ids = reinterpret_cast<int*>(p + 14 * plane + iplane);
    e->stage.flag = reinterpret_cast<int*>(p + 14 * plane + 2 * iplane);
    e->stage_n = n;
    return APJ_OK;
}
static void state_planes(const apj_state* h, double* (&f)[14]) {
    double* t[14] = {h->x, h->y, h->x_real, h->y_real, h->x0, h->y0, h->x_old, h->y_old, h->R, h->phi, h->cosp, h->sinp, h->vx, h->vy};
    for (int k = 0; k < 14; k++) f[k] = t[k];
}
// H2D of every present field into the staging planes; returns the `present` mask of the pack kernel
// (paired fields count only when both halves are given, as the documented defaults require)
static int stage_in(apj_engine* e, const apj_state* h, size_t n, unsigned* present) {
    double* f[14];
    state_planes(h, f);
    for (int k = 2; k < 14; k += 2) if (k != 8 && !(f[k] && f[k + 1])) f[k] = f[k + 1] = nullptr;
    unsigned m = 0;
    APJ_CUDA(e, cudaMemsetAsync(e->stage.flag, 0, sizeof(int), e->stream));
    for (int k = 0; k < 14; k++)
        if (f[k]) { m |= 1u << k; APJ_CUDA(e, cudaMemcpyAsync(e->stage.f[k], f[k], n * sizeof(double), cudaMemcpyHostToDevice, e->stream)); }
    *present = m;
    return APJ_OK;
}
static int stage_flag(apj_engine* e, int* flag) {
    APJ_CUDA(e, cudaMemcpyAsync(e->pin_small, e->stage.flag, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    *flag = e->pin_small[0];
    return APJ_OK;
}

extern "C" int apj_upload_state(apj_engine* e, const apj_state* h) {
    if (!e || !h) return APJ_E_INVALID;
    if (!h->x || !h->y || !h->R) return fail(e, APJ_E_INVALID, "apj_upload_state: x, y and R are required");
    if (!h->phi && !(h->cosp && h->sinp)) return fail(e, APJ_E_INVALID, "apj_upload_state: phi or (cosp, sinp) required");
    if (e->st.slab) return fail(e, APJ_E_STATE, "apj_upload_state: slab handle, use apj_slab_upload");
    DevState& st = e->st;
    const size_t n = (size_t)st.ntot;
    if (int rc = pull_ctl(e)) return rc;
    if (int rc = ensure_stage(e, n)) return rc;
    unsigned present = 0;
    if (int rc = stage_in(e, h, n, &present)) return rc;
    if (h->box) {
        present |= 1u << 14;
        APJ_CUDA(e, cudaMemcpyAsync(e->stage.box, h->box, n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    }
    // calculate_COM in index order (jamming.cpp:761-774) while the copies are in flight
    const double* sx = (h->x_real && h->y_real) ? h->x_real : h->x;
    const double* sy = (h->x_real && h->y_real) ? h->y_real : h->y;
    for (int s = 0; s < st.n_sys; s++) {
        SysCtl& c = e->hctl[s];
        const double* px = sx + (size_t)s * st.N;
        const double* py = sy + (size_t)s * st.N;
        double cx = 0.0, cy = 0.0;
        for (int i = 0; i < st.N; i++) { cx += px[i]; cy += py[i]; }
        c.COM[0] = cx / st.N; c.COM[1] = cy / st.N;
        c.COM_old[0] = c.COM0[0] = c.COM[0]; c.COM_old[1] = c.COM0[1] = c.COM[1];
        c.cur = 0; c.gen = 0; c.stale = 1; c.save_old = 0; c.ticket = 0; c.overflow = 0; c.list_max = 0;
        c.target = c.step;
    }
    apj_launch_pack(st, e->stream, e->stage, present, (long long)n, 0);
    e->launches++;
    APJ_CUDA(e, cudaMemsetAsync(st.cell_count, 0, sizeof(int) * e->total_cells, e->stream));
    int flag = 0;
    if (int rc = stage_flag(e, &flag)) return rc;
    if (flag & 1) return fail(e, APJ_E_INVALID, "apj_upload_state: radii must be positive");
    if (int rc = push_ctl(e)) return rc;
    e->have_state = true;
    return run_chain_now(e);   // assignCellsToGrid + buildVerletLists (start() :184-185)
}

// D2H of the wanted planes after apj_unpack_kernel filled them
static int stage_out(apj_engine* e, const apj_state* h, size_t n, int by_id) {
    double* f[14];
    state_planes(h, f);
    unsigned want = 0;
    for (int k = 0; k < 14; k++) if (f[k]) want |= 1u << k;
    if (h->box) want |= 1u << 14;
    if (!want) return APJ_OK;
    apj_launch_unpack(e->st, e->stream, e->stage, want, by_id);
    e->launches++;
    for (int k = 0; k < 14; k++)
        if (f[k]) APJ_CUDA(e, cudaMemcpyAsync(f[k], e->stage.f[k], n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (h->box) APJ_CUDA(e, cudaMemcpyAsync(h->box, e->stage.box, n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    return APJ_OK;
}

extern "C" int apj_download_state(apj_engine* e, apj_state* h) {
    if (!e || !h) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_download_state: no state uploaded");
    if (e->st.slab) return fail(e, APJ_E_STATE, "apj_download_state: slab handle, use apj_slab_download");
    const size_t n = (size_t)e->st.ntot;
    if (int rc = ensure_stage(e, n)) return rc;
    if (int rc = stage_out(e, h, n, 1)) return rc;       // scatter by id: original particle order
    return pull_ctl(e);                                   // synchronises the stream; refreshes hctl (checkpoint writer)
}

// ---- Engine::initCells / Engine::topology on the device (apj_setup.cu) -------------------------------------
extern "C" int apj_lattice_box_length(int64_t n, int32_t n_systems, uint64_t seed, const double* dens, int32_t device, double* L_out) {
    if (n < 1 || n_systems < 1 || !dens || !L_out) return fail(nullptr, APJ_E_INVALID, "apj_lattice_box_length: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, APJ_E_CUDA, "apj_lattice_box_length: no CUDA device (this library has no CPU fallback)");
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, APJ_E_CUDA, "apj_lattice_box_length: cudaSetDevice failed");
    double *d_part = nullptr, *d_sum = nullptr;
    std::vector<double> sums(n_systems);
    bool ok = cudaMalloc(&d_part, sizeof(double) * 1024 * (size_t)n_systems) == cudaSuccess && cudaMalloc(&d_sum, sizeof(double) * (size_t)n_systems) == cudaSuccess;
    if (ok) {
        apj_launch_radii_sums(nullptr, n, n_systems, seed, d_part, d_sum);
        ok = cudaMemcpy(sums.data(), d_sum, sizeof(double) * (size_t)n_systems, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    cudaFree(d_part); cudaFree(d_sum);
    if (!ok) return fail(nullptr, APJ_E_CUDA, "apj_lattice_box_length: CUDA error");
    for (int s = 0; s < n_systems; s++) L_out[s] = std::sqrt(APJ_PI * sums[s] / dens[s]);   // jamming.cpp:305
    return APJ_OK;
}

extern "C" int apj_init_lattice(apj_engine* e, uint64_t seed) {
    if (!e) return APJ_E_INVALID;
    if (e->st.slab) return fail(e, APJ_E_STATE, "apj_init_lattice: periodic handles only (upload the slabs of a box initialised elsewhere)");
    DevState& st = e->st;
    if (int rc = pull_ctl(e)) return rc;
    apj_launch_init_lattice(st, e->stream, seed);
    e->launches++;
    for (auto& c : e->hctl) {
        c.cur = 0; c.gen = 0; c.stale = 1; c.save_old = 0; c.ticket = 0; c.overflow = 0; c.list_max = 0;
        c.target = c.step; c.trunc_ok = 0;
    }
    APJ_CUDA(e, cudaMemsetAsync(st.cell_count, 0, sizeof(int) * e->total_cells, e->stream));
    if (int rc = push_ctl(e)) return rc;
    e->have_state = true;
    std::vector<double> com(2 * st.n_sys);                        // COM = mean(x_real), fixed-order device reduction
    if (int rc = apj_obs_com(&e->obs, st, e->stream, &e->launches, com.data())) return fail(e, rc, "apj_init_lattice: COM reduction failed");
    for (int s = 0; s < st.n_sys; s++) {
        SysCtl& c = e->hctl[s];
        c.COM[0] = com[2 * s]; c.COM[1] = com[2 * s + 1];
        c.COM0[0] = c.COM_old[0] = c.COM[0]; c.COM0[1] = c.COM_old[1] = c.COM[1];
    }
    if (int rc = push_ctl(e)) return rc;
    return run_chain_now(e);                                      // assignCellsToGrid + buildVerletLists (start() :184-185)
}

extern "C" int apj_get_box_table(apj_engine* e, int32_t s, double* centres, int32_t* neighbors) {
    if (!e || s < 0 || s >= e->st.n_sys || !centres || !neighbors) return APJ_E_INVALID;
    if (e->st.slab) return fail(e, APJ_E_STATE, "apj_get_box_table: periodic handles only");
    const size_t nbox = (size_t)e->hctl[s].nbox;
    double* dc = nullptr; int* dn = nullptr;
    APJ_CUDA(e, cudaMalloc(&dc, sizeof(double) * 2 * nbox));
    if (cudaMalloc(&dn, sizeof(int) * 9 * nbox) != cudaSuccess) { cudaFree(dc); return fail(e, APJ_E_CUDA, "apj_get_box_table: out of device memory"); }
    apj_launch_box_table(e->st, e->stream, s, dc, dn);
    e->launches++;
    cudaError_t a = cudaMemcpyAsync(centres, dc, sizeof(double) * 2 * nbox, cudaMemcpyDeviceToHost, e->stream);
    cudaError_t b = cudaMemcpyAsync(neighbors, dn, sizeof(int) * 9 * nbox, cudaMemcpyDeviceToHost, e->stream);
    cudaError_t c = cudaStreamSynchronize(e->stream);
    cudaFree(dc); cudaFree(dn);
    if (a != cudaSuccess || b != cudaSuccess || c != cudaSuccess) return fail(e, APJ_E_CUDA, "apj_get_box_table: CUDA error");
    return APJ_OK;
}

extern "C" int apj_overlap_hue(apj_engine* e, int32_t* over) {
    if (!e || !over) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_overlap_hue: no state uploaded");
    if (e->st.slab) return fail(e, APJ_E_STATE, "apj_overlap_hue: periodic handles only");
    const size_t n = (size_t)e->st.ntot;
    if (int rc = ensure_stage(e, n)) return rc;
    apj_launch_overlap_hue(e->st, e->stream, e->stage.box);       // staging plane reused: n ints by particle id
    e->launches++;
    APJ_CUDA(e, cudaMemcpyAsync(over, e->stage.box, n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    return APJ_OK;
}

// 64-bit fingerprint of {id, x, y, cos, sin} summed over the owned particles: independent of particle order
// and of the slab decomposition (ranks add their shares modulo 2^64).
extern "C" int apj_state_checksum(apj_engine* e, uint64_t* out) {
    if (!e || !out) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_state_checksum: no state uploaded");
    if (int rc = ensure_stage(e, 1)) return rc;
    unsigned long long* d = reinterpret_cast<unsigned long long*>(e->obs.d_hist);
    APJ_CUDA(e, cudaMemsetAsync(d, 0, sizeof(unsigned long long), e->stream));
    apj_launch_checksum(e->st, e->stream, d);
    e->launches++;
    APJ_CUDA(e, cudaMemcpyAsync(e->pin_small, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    memcpy(out, e->pin_small, sizeof(uint64_t));
    return APJ_OK;
}

extern "C" int apj_set_com(apj_engine* e, int32_t s, const double* com, const double* com0, const double* com_old) {
    if (!e || s < 0 || s >= e->st.n_sys) return APJ_E_INVALID;
    if (int rc = pull_ctl(e)) return rc;
    SysCtl& c = e->hctl[s];
    if (com) { c.COM[0] = com[0]; c.COM[1] = com[1]; }
    if (com0) { c.COM0[0] = com0[0]; c.COM0[1] = com0[1]; }
    if (com_old) { c.COM_old[0] = com_old[0]; c.COM_old[1] = com_old[1]; }
    c.trunc_ok = 0;   // the skin-test value no longer bounds the motion since the list build: sweep full lists
    return push_ctl(e);
}
extern "C" int apj_get_com(apj_engine* e, int32_t s, double* com, double* com0, double* com_old) {
    if (!e || s < 0 || s >= e->st.n_sys) return APJ_E_INVALID;
    if (int rc = read_ctl(e)) return rc;
    const SysCtl& c = e->hctl[s];
    if (com) { com[0] = c.COM[0]; com[1] = c.COM[1]; }
    if (com0) { com0[0] = c.COM0[0]; com0[1] = c.COM0[1]; }
    if (com_old) { com_old[0] = c.COM_old[0]; com_old[1] = c.COM_old[1]; }
    return APJ_OK;
}

__global__ void __launch_bounds__(256) apj_mark_origin_kernel(const DevState st) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int sys = (int)(min(g, st.ntot - 1) / st.cap);
    const SysCtl* ctl = st.ctl + sys;
    unsigned long long bits = 0ull;
    if (g < st.ntot && g - ctl->p0 < ctl->n_own) {
        const double2 x = st.XY[ctl->cur][g];
        // how far this particle has moved since the lists were built (COM drift removed, as newSkinList measures it):
        // the largest value over the system feeds SysCtl::skinBase
        const double2 xo = st.XO[ctl->gen][g];
        const double ddx = apj_delta_norm(((x.x - xo.x) - ctl->COM[0]) + ctl->COM_old[0], ctl->L, ctl->Lover2);
        const double ddy = apj_delta_norm(((x.y - xo.y) - ctl->COM[1]) + ctl->COM_old[1], ctl->L, ctl->Lover2);
        bits = (unsigned long long)__double_as_longlong(apj_d2(ddx, ddy));   // >= 0: bit patterns order like the values
        st.XR[ctl->cur][g] = x;     // x_real = x   (jamming.cpp:193-196)
        st.X0[ctl->gen][g] = x;     // x0 = x
        st.XO[ctl->gen][g] = x;     // saveOldPositions (:203)
    }
    // one atomic per warp (a warp that straddles two replicas -- N not a multiple of 32 -- lets every lane speak for itself)
    const bool uniform = __match_any_sync(0xffffffffu, sys) == 0xffffffffu;
    if (uniform) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long b = __shfl_xor_sync(0xffffffffu, bits, o); bits = b > bits ? b : bits; }
        if ((threadIdx.x & 31) != 0) bits = 0ull;
    }
    if (bits) atomicMax(const_cast<unsigned long long*>(&st.ctl[sys].mark_d2), bits);
}

extern "C" int apj_mark_origin(apj_engine* e) {
    if (!e) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_mark_origin: no state uploaded");
    DevState& st = e->st;
    if (int rc = pull_ctl(e)) return rc;
    for (auto& c : e->hctl) c.mark_d2 = 0ull;
    if (int rc = push_ctl(e)) return rc;
    e->ctl_clean = false;
    apj_mark_origin_kernel<<<(unsigned)((st.ntot + 255) / 256), 256, 0, e->stream>>>(st);
    e->launches++;
    // COM = mean(x_real) in the deterministic two-level order of the step kernel
    std::vector<double> com(2 * st.n_sys);
    if (int rc = apj_obs_com(&e->obs, st, e->stream, &e->launches, com.data())) return fail(e, rc, "apj_mark_origin: COM reduction failed");
    if (int rc = pull_ctl(e)) return rc;
    for (int s = 0; s < st.n_sys; s++) {
        SysCtl& c = e->hctl[s];
        c.COM[0] = com[2 * s]; c.COM[1] = com[2 * s + 1];
        c.COM0[0] = c.COM_old[0] = c.COM[0]; c.COM0[1] = c.COM_old[1] = c.COM[1];
        // x_old moved without a list build. Periodic handles keep the skin-aware sweep length alive: the pair displacement
        // between the build and this reset is bounded by twice the largest displacement just measured (skinBase), the
        // skin test restarts from zero. Slab ranks would need the maximum over all ranks to stay in lockstep: full
        // lists there until the next skin-triggered rebuild.
        double d2max;
        memcpy(&d2max, &c.mark_d2, sizeof d2max);
        if (st.slab) c.trunc_ok = 0;
        else c.skinBase += 2.0 * std::sqrt(d2max) * (1.0 + 1e-12);
        c.skinD = 0.0; c.skinDD = 0.0;
    }
    return push_ctl(e);
}

// ---- checkpoint / restart (the reference has none: a killed run is lost) --------------
// One binary file: header, per-system scalars (step, resetCounter, COM / COM0 / COM_old, activity, ramp, L),
// then the 14 fp64 fields of apj_state and the box ids, each n_systems * n values in original particle
// order. Loading re-bins and rebuilds the lists from the stored positions (assignCellsToGrid +
// buildVerletLists without the skin bookkeeping), keeps x_old / COM_old, and continues the Philox stream
// at the stored step. The continuation equals the uninterrupted run to rounding (first step <= 1e-12;
// list order, hence summation order, differs after the fresh rebuild) -- not bit for bit.
extern "C" int apj_slab_upload(apj_engine* e, const apj_state* h, const int32_t* ids, int64_t n_local);
extern "C" int apj_slab_download(apj_engine* e, apj_state* h, int32_t* ids, int64_t cap, int64_t* n_local);
namespace {
struct CkptHeader { char magic[8]; int64_t n; int32_t n_sys, version; uint64_t seed; double dt, rn2, rs2; };
// slab handles: one file per rank -- the rank's owned particles in device order with their ids. version = 2,
// n_sys carries nranks, `rank` and `n_local` follow the common header.
struct CkptSlab { int32_t rank, nranks; int64_t n_local; };
struct CkptSys { int64_t step, reset_counter, ramp_len, ramp_t0; double L, CFself, CTnoise, COM[2], COM0[2], COM_old[2]; };
const char CKPT_MAGIC[8] = {'A', 'P', 'J', 'C', 'K', 'P', 'T', '1'};
}

extern "C" int apj_save_checkpoint(apj_engine* e, const char* path) {
    if (!e || !path) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_save_checkpoint: no state uploaded");
    const DevState& st = e->st;
    if (st.slab) {   // this rank's share: every rank writes its own file (the caller names them apart)
        int64_t nl = 0;
        apj_state q = {};
        if (int rc = apj_slab_download(e, &q, nullptr, 0, &nl)) return rc;
        const size_t T = (size_t)std::max<int64_t>(nl, 1);
        std::vector<double> f[14];
        for (auto& v : f) v.resize(T);
        std::vector<int32_t> ids(T);
        apj_state h = {};
        h.x = f[0].data(); h.y = f[1].data(); h.x_real = f[2].data(); h.y_real = f[3].data(); h.x0 = f[4].data(); h.y0 = f[5].data();
        h.x_old = f[6].data(); h.y_old = f[7].data(); h.R = f[8].data(); h.phi = f[9].data(); h.cosp = f[10].data(); h.sinp = f[11].data();
        h.vx = f[12].data(); h.vy = f[13].data();
        if (int rc = apj_slab_download(e, &h, ids.data(), (int64_t)T, &nl)) return rc;   // also refreshes hctl
        FILE* fp = fopen(path, "wb");
        if (!fp) return fail(e, APJ_E_INVALID, "apj_save_checkpoint: cannot open the file for writing");
        CkptHeader hd = {};
        memcpy(hd.magic, CKPT_MAGIC, 8);
        hd.n = st.N; hd.n_sys = st.nranks; hd.version = 2; hd.seed = st.seed; hd.dt = st.dt; hd.rn2 = st.rn2; hd.rs2 = st.rs2;
        CkptSlab sl = {st.rank, st.nranks, nl};
        const SysCtl& c = e->hctl[0];
        CkptSys cs = {};
        cs.step = c.step; cs.reset_counter = c.reset_counter; cs.ramp_len = c.ramp_len; cs.ramp_t0 = c.ramp_t0;
        cs.L = c.L; cs.CFself = c.CFself; cs.CTnoise = c.CTnoise;
        for (int k = 0; k < 2; k++) { cs.COM[k] = c.COM[k]; cs.COM0[k] = c.COM0[k]; cs.COM_old[k] = c.COM_old[k]; }
        const size_t n = (size_t)nl;
        bool ok = fwrite(&hd, sizeof hd, 1, fp) == 1 && fwrite(&sl, sizeof sl, 1, fp) == 1 && fwrite(&cs, sizeof cs, 1, fp) == 1;
        ok = ok && fwrite(ids.data(), sizeof(int32_t), n, fp) == n;
        for (auto& v : f) ok = ok && fwrite(v.data(), sizeof(double), n, fp) == n;
        ok = (fclose(fp) == 0) && ok;
        return ok ? APJ_OK : fail(e, APJ_E_INVALID, "apj_save_checkpoint: short write");
    }
    const size_t T = (size_t)st.n_sys * st.N;
    std::vector<double> f[14];
    for (auto& v : f) v.resize(T);
    std::vector<int32_t> box(T);
    apj_state h = {};
    h.x = f[0].data(); h.y = f[1].data(); h.x_real = f[2].data(); h.y_real = f[3].data(); h.x0 = f[4].data(); h.y0 = f[5].data();
    h.x_old = f[6].data(); h.y_old = f[7].data(); h.R = f[8].data(); h.phi = f[9].data(); h.cosp = f[10].data(); h.sinp = f[11].data();
    h.vx = f[12].data(); h.vy = f[13].data(); h.box = box.data();
    if (int rc = apj_download_state(e, &h)) return rc;          // also refreshes hctl
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(e, APJ_E_INVALID, "apj_save_checkpoint: cannot open the file for writing");
    CkptHeader hd = {};
    memcpy(hd.magic, CKPT_MAGIC, 8);
    hd.n = st.N; hd.n_sys = st.n_sys; hd.version = 1; hd.seed = st.seed; hd.dt = st.dt; hd.rn2 = st.rn2; hd.rs2 = st.rs2;
    bool ok = fwrite(&hd, sizeof hd, 1, fp) == 1;
    for (int s = 0; s < st.n_sys && ok; s++) {
        const SysCtl& c = e->hctl[s];
        CkptSys cs = {};
        cs.step = c.step; cs.reset_counter = c.reset_counter; cs.ramp_len = c.ramp_len; cs.ramp_t0 = c.ramp_t0;
        cs.L = c.L; cs.CFself = c.CFself; cs.CTnoise = c.CTnoise;
        for (int k = 0; k < 2; k++) { cs.COM[k] = c.COM[k]; cs.COM0[k] = c.COM0[k]; cs.COM_old[k] = c.COM_old[k]; }
        ok = fwrite(&cs, sizeof cs, 1, fp) == 1;
    }
    for (auto& v : f) ok = ok && fwrite(v.data(), sizeof(double), T, fp) == T;
    ok = ok && fwrite(box.data(), sizeof(int32_t), T, fp) == T;
    ok = (fclose(fp) == 0) && ok;
    return ok ? APJ_OK : fail(e, APJ_E_INVALID, "apj_save_checkpoint: short write");
}

extern "C" int apj_load_checkpoint(apj_engine* e, const char* path) {
    if (!e || !path) return APJ_E_INVALID;
    DevState& st = e->st;
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(e, APJ_E_INVALID, "apj_load_checkpoint: cannot open the file");
    CkptHeader hd;
    if (fread(&hd, sizeof hd, 1, fp) != 1 || memcmp(hd.magic, CKPT_MAGIC, 8) != 0 || hd.version != (st.slab ? 2 : 1)) {
        fclose(fp);
        return fail(e, APJ_E_INVALID, st.slab ? "apj_load_checkpoint: not a slab-rank APJCKPT1 file" : "apj_load_checkpoint: not an APJCKPT1 file of a periodic box");
    }
    if (st.slab) {   // [collective] every rank loads the file its rank wrote
        CkptSlab sl;
        CkptSys cs;
        bool ok = fread(&sl, sizeof sl, 1, fp) == 1 && fread(&cs, sizeof cs, 1, fp) == 1;
        if (!ok || hd.n != st.N || hd.n_sys != st.nranks || sl.rank != st.rank || sl.nranks != st.nranks || hd.dt != st.dt || hd.rn2 != st.rn2 ||
            hd.rs2 != st.rs2 || cs.L != e->hctl[0].L || sl.n_local < 0 || sl.n_local > st.cap) {
            fclose(fp);
            return fail(e, APJ_E_INVALID, "apj_load_checkpoint: the file was written by another rank / decomposition / shape");
        }
        const size_t n = (size_t)sl.n_local, T = std::max<size_t>(n, 1);
        std::vector<int32_t> ids(T);
        std::vector<double> f[14];
        ok = fread(ids.data(), sizeof(int32_t), n, fp) == n;
        for (auto& v : f) { v.resize(T); ok = ok && fread(v.data(), sizeof(double), n, fp) == n; }
        fclose(fp);
        if (!ok) return fail(e, APJ_E_INVALID, "apj_load_checkpoint: truncated file");
        apj_state h = {};
        h.x = f[0].data(); h.y = f[1].data(); h.x_real = f[2].data(); h.y_real = f[3].data(); h.x0 = f[4].data(); h.y0 = f[5].data();
        h.x_old = f[6].data(); h.y_old = f[7].data(); h.R = f[8].data(); h.phi = f[9].data(); h.cosp = f[10].data(); h.sinp = f[11].data();
        h.vx = f[12].data(); h.vy = f[13].data();
        if (int rc = apj_slab_upload(e, &h, ids.data(), sl.n_local)) return rc;     // collective: ghost columns, lists
        if (int rc = pull_ctl(e)) return rc;
        SysCtl& c = e->hctl[0];
        c.step = c.target = cs.step; c.reset_counter = cs.reset_counter; c.ramp_len = cs.ramp_len; c.ramp_t0 = cs.ramp_t0;
        c.CFself = cs.CFself; c.CTnoise = cs.CTnoise;
        for (int k = 0; k < 2; k++) { c.COM[k] = cs.COM[k]; c.COM0[k] = cs.COM0[k]; c.COM_old[k] = cs.COM_old[k]; }
        c.no_self_once = 0;
        c.trunc_ok = 0;
        e->st.seed = hd.seed;
        if (int rc = push_ctl(e)) return rc;
        return build_group_graph(e);
    }
    if (hd.n != st.N || hd.n_sys != st.n_sys || hd.dt != st.dt || hd.rn2 != st.rn2 || hd.rs2 != st.rs2) {
        fclose(fp);
        return fail(e, APJ_E_INVALID, "apj_load_checkpoint: the file was written for another shape (n, n_systems, dt, rn, rs)");
    }
    std::vector<CkptSys> cs(st.n_sys);
    bool ok = fread(cs.data(), sizeof(CkptSys), st.n_sys, fp) == (size_t)st.n_sys;
    for (int s = 0; s < st.n_sys && ok; s++)
        if (cs[s].L != e->hctl[s].L) { fclose(fp); return fail(e, APJ_E_INVALID, "apj_load_checkpoint: box length differs from the handle's"); }
    const size_t T = (size_t)st.n_sys * st.N;
    std::vector<double> f[14];
    for (auto& v : f) { v.resize(T); ok = ok && fread(v.data(), sizeof(double), T, fp) == T; }
    std::vector<int32_t> box(T);
    ok = ok && fread(box.data(), sizeof(int32_t), T, fp) == T;
    fclose(fp);
    if (!ok) return fail(e, APJ_E_INVALID, "apj_load_checkpoint: truncated file");
    apj_state h = {};
    h.x = f[0].data(); h.y = f[1].data(); h.x_real = f[2].data(); h.y_real = f[3].data(); h.x0 = f[4].data(); h.y0 = f[5].data();
    h.x_old = f[6].data(); h.y_old = f[7].data(); h.R = f[8].data(); h.phi = f[9].data(); h.cosp = f[10].data(); h.sinp = f[11].data();
    h.vx = f[12].data(); h.vy = f[13].data(); h.box = box.data();
    if (int rc = apj_upload_state(e, &h)) return rc;            // re-bins, rebuilds the lists
    if (int rc = pull_ctl(e)) return rc;
    for (int s = 0; s < st.n_sys; s++) {
        SysCtl& c = e->hctl[s];
        c.step = c.target = cs[s].step; c.reset_counter = cs[s].reset_counter; c.ramp_len = cs[s].ramp_len; c.ramp_t0 = cs[s].ramp_t0;
        c.CFself = cs[s].CFself; c.CTnoise = cs[s].CTnoise;
        for (int k = 0; k < 2; k++) { c.COM[k] = cs[s].COM[k]; c.COM0[k] = cs[s].COM0[k]; c.COM_old[k] = cs[s].COM_old[k]; }
        c.no_self_once = 0;
        c.trunc_ok = 0;
    }
    e->st.seed = hd.seed;                                       // the Philox key travels with the run
    if (int rc = push_ctl(e)) return rc;
    return build_group_graph(e);                                // kernels take DevState (seed) by value
}

extern "C" int apj_sync(apj_engine* e) {
    if (!e) return APJ_E_INVALID;
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    return APJ_OK;
}

static bool all_done(const apj_engine* e, long long* remaining) {
    long long r = 0;
    for (const auto& c : e->hctl) r = std::max(r, c.target - c.step);
    *remaining = r;
    return r == 0;
}

extern "C" int apj_step(apj_engine* e, int64_t n_steps) {
    if (!e || n_steps < 0) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_step: no state uploaded");
    if (n_steps == 0) return APJ_OK;
    if (int rc = maybe_shrink_tile_cap(e)) return rc;
    DevState& st = e->st;
    e->ctl_clean = false;
    apj_add_target_kernel<<<(st.n_sys + 63) / 64, 64, 0, e->stream>>>(st.ctl, st.n_sys, n_steps);
    e->launches++;
    long long remaining = n_steps;
    for (int guard = 0; guard < 1 << 20; guard++) {
        // whole groups through the graph, the remainder as direct launches: no launch is issued that cannot
        // commit a step unless a rebuild fires (the speculative step is dropped and the rest of ITS group idles)
        const long long full = remaining / e->m, rem = remaining % e->m;
        for (long long k = 0; k < full; k++)
            if (int rc = launch_group(e)) return rc;
        if (rem || !full)
            if (int rc = launch_part(e, (int)rem)) return rc;
        if (int rc = pull_ctl(e)) return rc;
        int repaired = 0;
        if (int rc = repair_tile_overflow(e, &repaired)) return rc;
        if (int rc = repair_list_overflow(e, &repaired)) return rc;
        if (int rc = check_overflow(e)) return rc;
        if (all_done(e, &remaining)) return APJ_OK;
    }
    return fail(e, APJ_E_STATE, "apj_step: no progress");
}

extern "C" int apj_step_injected(apj_engine* e, const double* noise) {
    if (!e || !noise) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_step_injected: no state uploaded");
    DevState& st = e->st;
    APJ_CUDA(e, cudaMemcpyAsync(e->d_noise, noise, sizeof(double) * st.n_sys * st.N, cudaMemcpyHostToDevice, e->stream));
    e->ctl_clean = false;
    apj_add_target_kernel<<<(st.n_sys + 63) / 64, 64, 0, e->stream>>>(st.ctl, st.n_sys, 1);
    e->launches++;
    ApjLaunch l = launcher(e, true);
    for (int attempt = 0; attempt < 12; attempt++) {
        apj_launch_step(st, l, e->d_noise, 1);
        apj_launch_rebuild_chain(st, l, e->max_nbox, e->max_b);
        APJ_CUDA(e, cudaGetLastError());
        if (int rc = pull_ctl(e)) return rc;
        int repaired = 0;
        if (int rc = repair_tile_overflow(e, &repaired)) return rc;
        if (int rc = repair_list_overflow(e, &repaired)) return rc;
        if (int rc = check_overflow(e)) return rc;
        long long remaining;
        if (all_done(e, &remaining)) return APJ_OK;
    }
    return fail(e, APJ_E_STATE, "apj_step_injected: step did not commit");
}

extern "C" int apj_force_rebuild(apj_engine* e) {
    if (!e) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_force_rebuild: no state uploaded");
    if (int rc = pull_ctl(e)) return rc;
    for (auto& c : e->hctl) { c.stale = 1; c.save_old = 0; }
    if (int rc = push_ctl(e)) return rc;
    return run_chain_now(e);
}

extern "C" int apj_get_counters(apj_engine* e, int32_t s, int64_t* o) {
    if (!e || !o || s < 0 || s >= e->st.n_sys) return APJ_E_INVALID;
    if (int rc = read_ctl(e)) return rc;
    const SysCtl& c = e->hctl[s];
    o[0] = c.step; o[1] = c.reset_counter; o[2] = c.n_rebuilds; o[3] = c.list_max; o[4] = c.overflow;
    o[5] = e->launches; o[6] = c.n_discarded; o[7] = c.nbox;
    return APJ_OK;
}
static_assert(APJ_CLASSES == 4, "apj_get_sweep_stats reports four classes");
extern "C" int apj_get_sweep_stats(apj_engine* e, int32_t s, double* o) {
    if (!e || !o || s < 0 || s >= e->st.n_sys) return APJ_E_INVALID;
    if (int rc = read_ctl(e)) return rc;
    const SysCtl& c = e->hctl[s];
    o[0] = (double)c.n_retried; o[1] = (e->st.truncate && c.trunc_ok) ? 1.0 : 0.0; o[2] = c.skinD; o[3] = c.kmin;
    for (int k = 0; k < APJ_CLASSES; k++) o[4 + k] = (double)c.n_class[k];
    return APJ_OK;
}
extern "C" int apj_set_sweep_truncation(apj_engine* e, int32_t on) {
    if (!e) return APJ_E_INVALID;
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    e->st.truncate = on ? 1 : 0;
    return build_group_graph(e);   // kernels take DevState by value
}
extern "C" int apj_get_tuning(apj_engine* e, int32_t* o) {
    if (!e || !o) return APJ_E_INVALID;
    if (int rc = pull_ctl(e)) return rc;
    int tm = 0, nb = 0;
    for (auto& c : e->hctl) { tm = std::max(tm, c.tile_max); nb += c.nblk; }
    o[0] = e->st.G; o[1] = e->st.tb; o[2] = e->st.ppb; o[3] = e->st.tile_cap; o[4] = tm;
    o[5] = (int)((e->st.tile_cap + 1) * 48 + (e->st.G > 1 ? (size_t)e->st.ppb * 32 : 0)); o[6] = nb; o[7] = e->m;
    if (e->st.ring_nb > 0) o[5] *= e->st.ring_nb;          // ring kernel: all tile buffers of the one block of an SM
    return APJ_OK;
}
extern "C" int apj_set_reset_counter(apj_engine* e, int32_t s, int64_t v) {
    if (!e || s < 0 || s >= e->st.n_sys) return APJ_E_INVALID;
    if (int rc = pull_ctl(e)) return rc;
    e->hctl[s].reset_counter = v;
    return push_ctl(e);
}
extern "C" int apj_get_geometry(apj_engine* e, int32_t s, double* o) {
    if (!e || !o || s < 0 || s >= e->st.n_sys) return APJ_E_INVALID;
    const SysCtl& c = e->hctl[s];
    o[0] = c.L; o[1] = c.Lover2; o[2] = c.lp; o[3] = c.b; o[4] = c.nbox;
    return APJ_OK;
}

__global__ void apj_list_stats_kernel(const int* __restrict__ cnt, long long n, unsigned long long* __restrict__ out) {
    unsigned long long a = 0;
    int mx = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = cnt[i];
        a += (unsigned long long)c;
        mx = max(mx, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, a); atomicMax(out + 1, (unsigned long long)mx); }
}

extern "C" int apj_list_stats(apj_engine* e, int32_t s, int64_t* out2) {
    if (!e || !out2 || s < 0 || s >= e->st.n_sys) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_list_stats: no state uploaded");
    DevState& st = e->st;
    unsigned long long* d = reinterpret_cast<unsigned long long*>(e->obs.d_hist);
    APJ_CUDA(e, cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), e->stream));
    if (int rc = pull_ctl(e)) return rc;
    apj_list_stats_kernel<<<296, 256, 0, e->stream>>>(st.cnt + e->hctl[s].p0, e->hctl[s].n_own, d);
    e->launches++;
    unsigned long long h[2];
    APJ_CUDA(e, cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    out2[0] = (int64_t)h[0]; out2[1] = (int64_t)h[1];
    return APJ_OK;
}

extern "C" int apj_get_pair_list(apj_engine* e, int32_t s, int64_t* offsets, int32_t* idx, int64_t cap, int64_t* total) {
    if (!e || s < 0 || s >= e->st.n_sys || !total) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_get_pair_list: no state uploaded");
    if (e->st.slab) return fail(e, APJ_E_STATE, "apj_get_pair_list: slab handle, use apj_slab_get_pairs");
    DevState& st = e->st;
    if (int rc = pull_ctl(e)) return rc;
    const SysCtl& c = e->hctl[s];
    const long long o = (long long)s * st.N;
    const size_t N = st.N;
    const size_t words_per_blk = (size_t)st.max_quads * 4 * st.tb;
    std::vector<int> id(N), cnt(N);
    std::vector<TileDesc> tiles(c.nblk);
    std::vector<unsigned> lst(words_per_blk * c.nblk);
    APJ_CUDA(e, cudaMemcpyAsync(id.data(), st.ID[c.gen] + o, N * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaMemcpyAsync(cnt.data(), st.cnt + o, N * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaMemcpyAsync(tiles.data(), st.tiles + (size_t)s * st.maxblk, c.nblk * sizeof(TileDesc), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaMemcpyAsync(lst.data(), st.list32 + (size_t)s * st.maxblk * words_per_blk, lst.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    // decode tile slots -> particle index -> id; keep partners with larger id, ascending. Two passes over
    // the lists (count, fill) into CSR rows by id, so the 16M-particle box needs no per-particle containers.
    std::vector<int64_t> row(N + 1, 0);
    std::vector<int> buf;
    size_t covered = 0;
    for (int pass = 0; pass < 2; pass++) {
        covered = 0;
        for (int blk = 0; blk < c.nblk; blk++) {
            const TileDesc& d = tiles[blk];
            for (int t = 0; t < d.n; t++, covered++) {
                const long long g = d.g0 + t - o;
                if (g < 0 || g >= (long long)N) return fail(e, APJ_E_STATE, "apj_get_pair_list: tile outside its system");
                for (int k = 0; k < cnt[g]; k++) {
                    const int wi = k >> 1, sub = wi % st.G, kk = wi / st.G;
                    const unsigned w = lst[blk * words_per_blk + ((size_t)(kk >> 2) * st.tb + (size_t)t * st.G + sub) * 4 + (kk & 3)];
                    int slot = (int)(((k & 1) ? (w >> 16) : (w & 0xffffu)) >> 4) - 1;   // entries are (1-based slot) * 16
                    long long j = -1;
                    for (int p = 0; p < (d.info & 0xff); p++) {
                        if (slot < d.plen[p]) { j = (long long)d.pstart[p] + slot - o; break; }
                        slot -= d.plen[p];
                    }
                    if (j < 0 || j >= (long long)N) return fail(e, APJ_E_STATE, "apj_get_pair_list: list entry outside its tile");
                    if (id[j] > id[g]) {
                        if (pass == 0) row[id[g] + 1]++;
                        else buf[row[id[g]]++] = id[j];
                    }
                }
            }
        }
        if (covered != N) return fail(e, APJ_E_STATE, "apj_get_pair_list: work blocks do not cover the system");
        if (pass == 0) {
            for (size_t i = 0; i < N; i++) row[i + 1] += row[i];
            if (!idx || cap < row[N]) {              // size query (or a buffer that cannot hold the pairs)
                if (offsets) for (size_t i = 0; i <= N; i++) offsets[i] = row[i];
                *total = row[N];
                return APJ_OK;
            }
            buf.resize((size_t)row[N]);
        }
    }
    // the fill advanced row[i] to the end of row i: row i is [row[i-1], row[i])
    int64_t tot = 0;
    for (size_t i = 0; i < N; i++) {
        const int64_t r0 = i ? row[i - 1] : 0, r1 = row[i];
        std::sort(buf.begin() + r0, buf.begin() + r1);
        if (offsets) offsets[i] = r0;
        for (int64_t q = r0; q < r1; q++) idx[q] = buf[q];
        tot = r1;
    }
    if (offsets) offsets[N] = tot;
    *total = tot;
    return APJ_OK;
}

extern "C" int apj_get_cell_lists(apj_engine* e, int32_t s, int64_t* offsets, int32_t* idx) {
    if (!e || s < 0 || s >= e->st.n_sys || !offsets || !idx) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_get_cell_lists: no state uploaded");
    if (e->st.slab) return fail(e, APJ_E_STATE, "apj_get_cell_lists: not available on a slab handle");
    DevState& st = e->st;
    if (int rc = pull_ctl(e)) return rc;
    const SysCtl& c = e->hctl[s];
    const long long o = (long long)s * st.N;
    const size_t N = st.N;
    std::vector<int> id(N), box(N);
    APJ_CUDA(e, cudaMemcpyAsync(id.data(), st.ID[c.gen] + o, N * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaMemcpyAsync(box.data(), st.BOX[c.gen] + o, N * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    std::vector<int> ref_box(N);
    for (size_t g = 0; g < N; g++) ref_box[id[g]] = (box[g] / c.b) + (box[g] % c.b) * c.b;   // -> i + j*b
    std::vector<int64_t> count(c.nbox + 1, 0);
    for (size_t i = 0; i < N; i++) count[ref_box[i] + 1]++;
    for (int p = 0; p < c.nbox; p++) count[p + 1] += count[p];
    for (int p = 0; p <= c.nbox; p++) offsets[p] = count[p];
    std::vector<int64_t> fill(count.begin(), count.end() - 1);
    for (size_t i = 0; i < N; i++) idx[fill[ref_box[i]]++] = (int)i;
    return APJ_OK;
}

// ---- observables: thin wrappers over apj_observe.cu ---------------------------------------
#define APJ_NEED_STATE(name) \
    if (!e) return APJ_E_INVALID; \
    if (!e->have_state) return fail(e, APJ_E_STATE, name ": no state uploaded")
#define APJ_OBS(call) do { int _rc = (call); if (_rc) return fail(e, _rc, cudaGetErrorString(cudaGetLastError())); return APJ_OK; } while (0)

extern "C" int apj_order_orientation(apj_engine* e, double* order, double* orient) {
    APJ_NEED_STATE("apj_order_orientation");
    APJ_OBS(apj_obs_order(&e->obs, e->st, e->stream, &e->launches, order, orient));
}
extern "C" int apj_msd(apj_engine* e, double* msd) {
    APJ_NEED_STATE("apj_msd");
    APJ_OBS(apj_obs_msd(&e->obs, e->st, e->stream, &e->launches, msd));
}
extern "C" int apj_fluct_area(apj_engine* e, const double* radius, double* area) {
    APJ_NEED_STATE("apj_fluct_area");
    if (!radius || !area) return APJ_E_INVALID;
    APJ_OBS(apj_obs_fluct(&e->obs, e->st, e->stream, &e->launches, radius, area));
}
extern "C" int apj_obs_enqueue(apj_engine* e, int32_t kind, const double* param, int64_t* ticket) {
    APJ_NEED_STATE("apj_obs_enqueue");
    long long tk = 0;
    const int rc = apj_obs_enqueue_ring(&e->obs, e->st, e->stream, &e->launches, kind, param, &tk);
    if (rc == -1) return fail(e, APJ_E_INVALID, "apj_obs_enqueue: kind must be APJ_OBS_COM..APJ_OBS_FLUCT (FLUCT needs the radii)");
    if (rc) return fail(e, rc, cudaGetErrorString(cudaGetLastError()));
    if (ticket) *ticket = tk;
    return APJ_OK;
}
extern "C" int apj_obs_fetch(apj_engine* e, int64_t first_ticket, int64_t count, double* out) {
    APJ_NEED_STATE("apj_obs_fetch");
    if (!out && count > 0) return APJ_E_INVALID;
    const int rc = apj_obs_fetch_ring(&e->obs, e->st, e->stream, first_ticket, count, out);
    if (rc == -1) return fail(e, APJ_E_INVALID, "apj_obs_fetch: tickets not issued yet, or already overwritten (the ring holds the last 4096)");
    if (rc) return fail(e, rc, cudaGetErrorString(cudaGetLastError()));
    return APJ_OK;
}
extern "C" int apj_spatial_correlations(apj_engine* e, double cutoff, double* counts, double* ori, double* vel, double* pair) {
    APJ_NEED_STATE("apj_spatial_correlations");
    if (!(cutoff > 0) || !counts || !ori || !vel || !pair) return APJ_E_INVALID;
    if (e->st.slab && e->st.nranks > 1) return fail(e, APJ_E_STATE, "apj_spatial_correlations: slab handle, use apj_slab_spatial_correlations (pairs across the slab edges need the neighbour's edge particles)");
    // nc*dr_c must not exceed the cutoff, or the reference's boxPairs filter (jamming.cpp:460-479)
    // would decide membership of the last bin; its own settings (20, 140) are multiples of 2.
    if (std::fmod(cutoff, 2.0) != 0.0) return fail(e, APJ_E_INVALID, "apj_spatial_correlations: cutoff must be a multiple of dr_c = 2");
    APJ_OBS(apj_obs_spatial(&e->obs, e->st, e->stream, &e->launches, e->hctl.data(), cutoff, counts, ori, vel, pair, e->allocs));
}
// ---- Correlations::spatialCorrelations on a decomposed box (halo of the cutoff width) ----
extern "C" int apj_slab_export_edge(apj_engine* e, double width, int64_t cap, double* out6, int64_t* n) {
    APJ_NEED_STATE("apj_slab_export_edge");
    if (!e->st.slab || !(width > 0) || !n || cap < 0) return APJ_E_INVALID;
    long long found = 0;
    if (int rc = apj_obs_export_edge(e->st, e->stream, &e->launches, width, cap, out6, &found)) return fail(e, rc, cudaGetErrorString(cudaGetLastError()));
    *n = found;
    return APJ_OK;
}
extern "C" int apj_slab_spatial_correlations(apj_engine* e, double cutoff, int64_t n_ext, const double* ext6, const int32_t* ext_row,
                                             double* counts, double* ori, double* vel, double* pair) {
    APJ_NEED_STATE("apj_slab_spatial_correlations");
    if (!e->st.slab || !(cutoff > 0) || !counts || !ori || !vel || !pair || n_ext < 0 || (n_ext > 0 && (!ext6 || !ext_row))) return APJ_E_INVALID;
    if (std::fmod(cutoff, 2.0) != 0.0) return fail(e, APJ_E_INVALID, "apj_slab_spatial_correlations: cutoff must be a multiple of dr_c = 2");
    if (int rc = pull_ctl(e)) return rc;
    const SysCtl& c = e->hctl[0];
    if (e->st.nranks > 1) {
        // a pair must be seen by exactly one rank: from its left particle, across ONE slab edge
        const double width = c.ncols * c.lp, need = (e->st.nranks == 2 ? 2.0 : 1.0) * cutoff + c.lp;
        if (width < need) return fail(e, APJ_E_INVALID, "apj_slab_spatial_correlations: slabs narrower than the cutoff (+ one cell; twice the cutoff on 2 ranks): "
                                                        "gather the box into a periodic handle for this cutoff");
    }
    APJ_OBS(apj_obs_spatial(&e->obs, e->st, e->stream, &e->launches, e->hctl.data(), cutoff, counts, ori, vel, pair, e->allocs,
                            e->st.nranks > 1 ? n_ext : 0, ext6, reinterpret_cast<const int*>(ext_row)));
}
extern "C" int apj_vel_hist(apj_engine* e, const double* dv, int64_t* hist100) {
    APJ_NEED_STATE("apj_vel_hist");
    if (!dv || !hist100) return APJ_E_INVALID;
    APJ_OBS(apj_obs_velhist(&e->obs, e->st, e->stream, &e->launches, dv, hist100));
}
extern "C" int apj_occupancy_hist(apj_engine* e, int64_t* hist50) {
    APJ_NEED_STATE("apj_occupancy_hist");
    if (!hist50) return APJ_E_INVALID;
    APJ_OBS(apj_obs_occupancy(&e->obs, e->st, e->stream, &e->launches, hist50));
}

extern "C" int apj_timer_begin(apj_engine* e) {
    if (!e) return APJ_E_INVALID;
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    APJ_CUDA(e, cudaEventRecord(e->ev0, e->stream));
    return APJ_OK;
}
extern "C" int apj_timer_end(apj_engine* e, float* ms) {
    if (!e || !ms) return APJ_E_INVALID;
    APJ_CUDA(e, cudaEventRecord(e->ev1, e->stream));
    APJ_CUDA(e, cudaEventSynchronize(e->ev1));
    APJ_CUDA(e, cudaEventElapsedTime(ms, e->ev0, e->ev1));
    return APJ_OK;
}

extern "C" int apj_time_step_kernel(apj_engine* e, int64_t n, float* mean_ms, int64_t* committed) {
    if (!e || n < 1 || n > 4096 || !mean_ms) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_time_step_kernel: no state uploaded");
    DevState& st = e->st;
    if (int rc = pull_ctl(e)) return rc;
    const long long step0 = e->hctl[0].step;
    struct Events {   // destroyed on every return path
        std::vector<cudaEvent_t> v;
        ~Events() { for (auto x : v) if (x) cudaEventDestroy(x); }
    } evs;
    evs.v.assign(2 * n, nullptr);
    std::vector<cudaEvent_t>& ev = evs.v;
    for (auto& x : ev) APJ_CUDA(e, cudaEventCreate(&x));
    ApjLaunch l = launcher(e, true);
    // one target for the whole series (as apj_step does): only the very last step takes the
    // "store the observables' fields" path, every other launch is the steady-state kernel
    e->ctl_clean = false;
    apj_add_target_kernel<<<(st.n_sys + 63) / 64, 64, 0, e->stream>>>(st.ctl, st.n_sys, n);
    e->launches++;
    for (int64_t k = 0; k < n; k++) {
        APJ_CUDA(e, cudaEventRecord(ev[2 * k], e->stream));
        apj_launch_step(st, l, nullptr, 0);
        APJ_CUDA(e, cudaEventRecord(ev[2 * k + 1], e->stream));
        apj_launch_rebuild_chain(st, l, e->max_nbox, e->max_b);
    }
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    double tot = 0;
    for (int64_t k = 0; k < n; k++) { float ms = 0; APJ_CUDA(e, cudaEventElapsedTime(&ms, ev[2 * k], ev[2 * k + 1])); tot += ms; }
    *mean_ms = (float)(tot / n);
    if (int rc = pull_ctl(e)) return rc;
    if (committed) *committed = e->hctl[0].step - step0;
    // launches that were dropped (rebuild fired, sweep too short) did not advance the step: finish the series
    long long remaining;
    if (!all_done(e, &remaining)) {
        for (int a = 0; a < 256 && !all_done(e, &remaining); a++) {
            apj_launch_step(st, l, nullptr, 0);
            apj_launch_rebuild_chain(st, l, e->max_nbox, e->max_b);
            if (int rc = pull_ctl(e)) return rc;
        }
    }
    return check_overflow(e);
}

// Per-part timing of a step: event after the step kernel, after the fold + commit kernel and after the slab commit
// (the all-rank rendezvous), n single steps; out3 = mean ms of the three parts (zeros where a part does not exist).
extern "C" int apj_time_step_parts(apj_engine* e, int64_t n, float* out3) {
    if (!e || n < 1 || n > 1024 || !out3) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_time_step_parts: no state uploaded");
    DevState& st = e->st;
    struct Events {
        std::vector<cudaEvent_t> v;
        cudaStream_t s; int k;
        ~Events() { for (auto x : v) if (x) cudaEventDestroy(x); }
    } evs;
    evs.v.assign(4 * n, nullptr);
    evs.s = e->stream; evs.k = 0;
    for (auto& x : evs.v) APJ_CUDA(e, cudaEventCreate(&x));
    ApjLaunch l = launcher(e, true);
    e->ctl_clean = false;
    apj_add_target_kernel<<<(st.n_sys + 63) / 64, 64, 0, e->stream>>>(st.ctl, st.n_sys, n);
    e->launches++;
    for (int64_t k = 0; k < n; k++) {
        evs.k = (int)(4 * k);
        APJ_CUDA(e, cudaEventRecord(evs.v[4 * k], e->stream));
        apj_launch_step_parts(st, l, [](int part, void* a) { Events* ev = static_cast<Events*>(a); cudaEventRecord(ev->v[ev->k + 1 + part], ev->s); }, &evs);
        apj_launch_rebuild_chain(st, l, e->max_nbox, e->max_b);
    }
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    double tot[3] = {0, 0, 0};
    for (int64_t k = 0; k < n; k++)
        for (int p = 0; p < 3; p++) { float ms = 0; APJ_CUDA(e, cudaEventElapsedTime(&ms, evs.v[4 * k + p], evs.v[4 * k + p + 1])); tot[p] += ms; }
    for (int p = 0; p < 3; p++) out3[p] = (float)(tot[p] / n);
    if (int rc = pull_ctl(e)) return rc;
    long long remaining;
    for (int a = 0; a < 256 && !all_done(e, &remaining); a++) {   // launches dropped by a rebuild: finish the series
        apj_launch_step(st, l, nullptr, 0);
        apj_launch_rebuild_chain(st, l, e->max_nbox, e->max_b);
        if (int rc = pull_ctl(e)) return rc;
    }
    return check_overflow(e);
}

// ---- slab mode: one global periodic box over the GPUs of a node (BASELINE config 4) ----
extern "C" int apj_slab_create(const apj_config* cfg, double L, int32_t rank, int32_t nranks, int64_t capacity, apj_engine** out) {
    SlabSpec sp;
    sp.on = 1; sp.rank = rank; sp.nranks = nranks; sp.cap = capacity;
    return create_impl(cfg, &L, sp, out);
}

extern "C" int apj_slab_info(apj_engine* e, int64_t* o) {
    if (!e || !o || !e->st.slab) return APJ_E_INVALID;
    if (int rc = pull_ctl(e)) return rc;
    const SysCtl& c = e->hctl[0];
    o[0] = e->st.rank; o[1] = e->st.nranks; o[2] = c.col0; o[3] = c.ncols; o[4] = e->st.cap; o[5] = e->st.gcap;
    o[6] = c.n_own; o[7] = (int64_t)e->arena_bytes;
    return APJ_OK;
}

extern "C" int apj_slab_export(apj_engine* e, void* handle64, uint64_t* local_ptr) {
    if (!e || !e->st.slab) return APJ_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (handle64) {
        cudaIpcMemHandle_t h;
        APJ_CUDA(e, cudaIpcGetMemHandle(&h, e->st.arena));
        memcpy(handle64, &h, sizeof h);
    }
    if (local_ptr) *local_ptr = (uint64_t)(uintptr_t)e->st.arena;
    return APJ_OK;
}

extern "C" int apj_slab_connect(apj_engine* e, int32_t peer, const void* handle64, uint64_t same_process_ptr, int32_t peer_device) {
    if (!e || !e->st.slab || peer < 0 || peer >= e->st.nranks) return APJ_E_INVALID;
    if (peer == e->st.rank) return APJ_OK;
    APJ_CUDA(e, cudaSetDevice(e->cfg.device));
    if (same_process_ptr) {          // peer handle lives in this process (tests: same device; or another device of the node)
        if (peer_device >= 0 && peer_device != e->cfg.device) {
            int can = 0;
            APJ_CUDA(e, cudaDeviceCanAccessPeer(&can, e->cfg.device, peer_device));
            if (!can) return fail(e, APJ_E_CUDA, "apj_slab_connect: no peer access between the two devices");
            cudaError_t s = cudaDeviceEnablePeerAccess(peer_device, 0);
            if (s != cudaSuccess && s != cudaErrorPeerAccessAlreadyEnabled) APJ_CUDA(e, s);
            cudaGetLastError();
        }
        e->st.peer_arena[peer] = reinterpret_cast<char*>((uintptr_t)same_process_ptr);
    } else {
        if (!handle64) return APJ_E_INVALID;
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, sizeof h);
        void* p = nullptr;
        APJ_CUDA(e, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        e->ipc_opened.push_back(p);
        e->st.peer_arena[peer] = reinterpret_cast<char*>(p);
    }
    e->peer_set[peer] = true;
    return APJ_OK;
}

extern "C" int apj_slab_set_timeout(apj_engine* e, double seconds) {
    if (!e || !e->st.slab || !(seconds > 0)) return APJ_E_INVALID;
    e->st.timeout_ns = (unsigned long long)(seconds * 1e9);
    return e->slab_ready ? build_group_graph(e) : APJ_OK;   // kernels take DevState by value
}

extern "C" int apj_slab_ready(apj_engine* e) {
    if (!e || !e->st.slab) return APJ_E_INVALID;
    for (int r = 0; r < e->st.nranks; r++)
        if (!e->peer_set[r]) return fail(e, APJ_E_STATE, "apj_slab_ready: not every peer rank is connected");
    e->slab_ready = true;
    return build_group_graph(e);
}

extern "C" int apj_slab_upload(apj_engine* e, const apj_state* h, const int32_t* ids, int64_t n_local) {
    if (!e || !h || !e->st.slab || n_local < 0 || (n_local > 0 && !ids)) return APJ_E_INVALID;
    if (e->st.nranks > 1 && !e->slab_ready) return fail(e, APJ_E_STATE, "apj_slab_upload: call apj_slab_ready first");
    if (n_local > 0 && (!h->x || !h->y || !h->R)) return fail(e, APJ_E_INVALID, "apj_slab_upload: x, y and R are required");
    if (n_local > 0 && !h->phi && !(h->cosp && h->sinp)) return fail(e, APJ_E_INVALID, "apj_slab_upload: phi or (cosp, sinp) required");
    DevState& st = e->st;
    if (n_local > st.cap) return fail(e, APJ_E_OVERFLOW, "apj_slab_upload: more particles than this rank's capacity");
    if (int rc = pull_ctl(e)) return rc;
    const size_t n = (size_t)n_local;
    if (int rc = ensure_stage(e, (size_t)st.cap)) return rc;
    unsigned present = 0;
    if (n > 0) {
        if (int rc = stage_in(e, h, n, &present)) return rc;
        APJ_CUDA(e, cudaMemcpyAsync(e->stage.ids, ids, n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
        apj_launch_pack(st, e->stream, e->stage, present, (long long)n, 1);
        e->launches++;
        int flag = 0;
        if (int rc = stage_flag(e, &flag)) return rc;
        // NOTE: the other ranks are already inside the collective rebuild; they report the missing peer after the timeout
        if (flag & 1) return fail(e, APJ_E_INVALID, "apj_slab_upload: radii must be positive");
        if (flag & 2) return fail(e, APJ_E_INVALID, "apj_slab_upload: particle id outside [0, N)");
    }
    SysCtl& c = e->hctl[0];
    c.cur = 0; c.gen = 0; c.stale = 1; c.save_old = 0; c.ticket = 0; c.overflow = 0; c.list_max = 0;
    c.target = c.step; c.n_own = (int)n_local; c.p0 = 0;
    APJ_CUDA(e, cudaMemsetAsync(st.cell_count, 0, sizeof(int) * e->total_cells, e->stream));
    if (int rc = push_ctl(e)) return rc;
    e->have_state = true;
    return run_chain_now(e);   // collective: migration of misplaced particles, ghost columns, lists
}

extern "C" int apj_slab_download(apj_engine* e, apj_state* h, int32_t* ids, int64_t cap, int64_t* n_local) {
    if (!e || !h || !e->st.slab || !n_local) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_slab_download: no state uploaded");
    DevState& st = e->st;
    if (int rc = pull_ctl(e)) return rc;
    const SysCtl& c = e->hctl[0];
    const size_t n = (size_t)c.n_own;
    *n_local = (int64_t)n;
    if (!ids || cap < (int64_t)n) return cap < (int64_t)n && ids ? fail(e, APJ_E_INVALID, "apj_slab_download: buffers too small") : APJ_OK;
    if (n == 0) return APJ_OK;
    if (int rc = ensure_stage(e, (size_t)st.cap)) return rc;
    if (int rc = stage_out(e, h, n, 0)) return rc;       // device (cell) order; the ids say who is who
    APJ_CUDA(e, cudaMemcpyAsync(ids, st.ID[c.gen], n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    return APJ_OK;
}

extern "C" int apj_slab_get_pairs(apj_engine* e, int32_t* pairs, int64_t cap_pairs, int64_t* total) {
    if (!e || !e->st.slab || !total) return APJ_E_INVALID;
    if (!e->have_state) return fail(e, APJ_E_STATE, "apj_slab_get_pairs: no state uploaded");
    DevState& st = e->st;
    if (int rc = pull_ctl(e)) return rc;
    const SysCtl& c = e->hctl[0];
    const size_t nall = (size_t)st.cap + 2 * (size_t)st.gcap;
    const size_t words_per_blk = (size_t)st.max_quads * 4 * st.tb;
    std::vector<int> id(nall), cnt(st.cap);
    std::vector<TileDesc> tiles(c.nblk);
    std::vector<unsigned> lst(words_per_blk * c.nblk);
    APJ_CUDA(e, cudaMemcpyAsync(id.data(), st.ID[c.gen], nall * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaMemcpyAsync(cnt.data(), st.cnt, (size_t)st.cap * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaMemcpyAsync(tiles.data(), st.tiles, c.nblk * sizeof(TileDesc), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaMemcpyAsync(lst.data(), st.list32, lst.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream));
    APJ_CUDA(e, cudaStreamSynchronize(e->stream));
    int64_t tot = 0;
    size_t covered = 0;
    for (int blk = 0; blk < c.nblk; blk++) {
        const TileDesc& d = tiles[blk];
        for (int t = 0; t < d.n; t++, covered++) {
            const long long g = d.g0 + t;
            if (g < 0 || g >= c.n_own) return fail(e, APJ_E_STATE, "apj_slab_get_pairs: tile outside the owned range");
            for (int k = 0; k < cnt[g]; k++) {
                const int wi = k >> 1, sub = wi % st.G, kk = wi / st.G;
                const unsigned w = lst[blk * words_per_blk + ((size_t)(kk >> 2) * st.tb + (size_t)t * st.G + sub) * 4 + (kk & 3)];
                int slot = (int)(((k & 1) ? (w >> 16) : (w & 0xffffu)) >> 4) - 1;
                long long j = -1;
                for (int p = 0; p < (d.info & 0xff); p++) {
                    if (slot < d.plen[p]) { j = (long long)d.pstart[p] + slot; break; }
                    slot -= d.plen[p];
                }
                if (j < 0 || j >= (long long)nall || (j >= c.n_own && j < st.cap)) return fail(e, APJ_E_STATE, "apj_slab_get_pairs: list entry outside its tile");
                if (id[j] > id[g]) {
                    if (pairs && tot < cap_pairs) { pairs[2 * tot] = id[g]; pairs[2 * tot + 1] = id[j]; }
                    tot++;
                }
            }
        }
    }
    if (covered != (size_t)c.n_own) return fail(e, APJ_E_STATE, "apj_slab_get_pairs: work blocks do not cover the owned particles");
    *total = tot;
    return APJ_OK;
}
