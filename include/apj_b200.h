/* apj_b200.h -- C ABI of the B200-native active-particle-jamming hot path.
 *
 * The reference (danielmccusker/active-particle-jamming) has no FFI layer: its hot path is a set
 * of `Engine` methods in one translation unit (code/jam/jamming.cpp) over header-only structs
 * (the headers of code/classes). This header is the boundary a host `Engine` binds instead of those
 * methods; every entry point names the reference code it replaces (file:line relative to the
 * reference root). INTEGRATION.md shows the reference-side binding.
 *
 * Conventions: plain C, no exceptions; every function returns 0 on success or a negative
 * APJ_E_* code, with a human-readable message from apj_last_error(); all device memory is owned
 * by the handle; all buffers passed in are caller-owned HOST memory laid out one value per
 * particle in ORIGINAL PARTICLE INDEX order, system-major (index = system * N + id); one CUDA
 * stream per handle; calls on one handle must come from one host thread. There is NO CPU
 * fallback: without a usable CUDA device apj_create fails with APJ_E_CUDA.
 */
#ifndef APJ_B200_H
#define APJ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APJ_OK 0
#define APJ_E_INVALID (-1)  /* bad argument */
#define APJ_E_CUDA (-2)     /* CUDA runtime / driver error (message has the CUDA string) */
#define APJ_E_OVERFLOW (-3) /* a Verlet list longer than the 96 entries the layout holds, or a slab capacity exceeded */
#define APJ_E_STATE (-4)    /* call not valid in the current state (e.g. no state uploaded) */
/* (-5 is unassigned: the slab data path runs on device-side peer stores over NVLink, not on NCCL -- see "slab mode") */

typedef struct apj_engine apj_engine;

/* Engine::Engine + the constants it fixes (jamming.cpp:57, :112-115, :119-165). */
typedef struct apj_config {
    int64_t n;             /* particles per system (Engine::N) */
    int32_t n_systems;     /* independent replicas batched in one handle (phase-diagram sweep); >= 1 */
    int32_t device;        /* CUDA device ordinal */
    double dt;             /* 0 -> 0.1  (jamming.cpp:57) */
    double rn;             /* 0 -> 2.8  (jamming.cpp:112) */
    double rs_factor;      /* 0 -> 1.5  (rs = 1.5*rn, jamming.cpp:113) */
    uint64_t seed;         /* Philox4x32-10 key */
    int32_t max_neighbors; /* INITIAL capacity of one full Verlet list; 0 -> 48. Grown on demand up to 96 (periodic handles) */
    int32_t steps_per_launch; /* speculative steps between rebuild checks; 0 -> 16 (32 for systems of fewer than 4096 work blocks) */
    int32_t flags;         /* APJ_FLAG_* */
    int32_t tile_slots;    /* shared-memory tile capacity of a work block, in particles; 0 -> adaptive (follows the largest tile) */
    int32_t lanes_per_particle; /* threads cooperating on one neighbour sweep: 1, 2, 4, 8; 0 -> by system size */
    int32_t reserved;
} apj_config;

#define APJ_FLAG_NO_GRAPH 1  /* launch kernels directly instead of through a CUDA graph */
#define APJ_FLAG_SPLIT_TAIL 2  /* end every step with the separate fold + commit kernel (default: systems of >= 4096 work blocks) */
#define APJ_FLAG_FUSED_TAIL 4  /* ... or always fold + commit in the step kernel's last block (default: small systems) */
#define APJ_FLAG_PERSIST 16    /* pipelined persistent form of the split-tail step kernel (one system): one resident wave of blocks, each
                                * walking tiles blk, blk + grid, ...; the last warp to finish a sweep refills the tile buffer (TMA) under
                                * the epilogue. Bit-identical positions; measured on a par with one block per tile on B200 at N = 16M
                                * (DESIGN.md section 7), hence opt-in (env APJ_STEP_PIPE=1 forces it on for tuning runs) */
#define APJ_FLAG_TINY_GRID 8   /* tests, with APJ_FLAG_PERSIST: only 3 blocks, so every block walks many tiles */

/* Host view of the per-particle fields of `struct Cell` (classes/Cell.h:15-43), 2D. Any
 * pointer may be NULL: on upload a NULL field takes the documented default, on download it is
 * skipped. */
typedef struct apj_state {
    double *x, *y;           /* Cell::x          */
    double *x_real, *y_real; /* Cell::x_real     (upload default: x, y) */
    double *x0, *y0;         /* Cell::x0         (upload default: x_real) */
    double *x_old, *y_old;   /* Cell::x_old      (upload default: x, y) */
    double *R;               /* Cell::R          (Rinv = 1/R is derived, jamming.cpp:298) */
    double *phi;             /* Cell::phi */
    double *cosp, *sinp;     /* Cell::cosp/sinp  (upload default: cos(phi), sin(phi), jamming.cpp:332-333) */
    double *vx, *vy;         /* Cell::vx, vy     (upload default: 0) */
    int32_t *box;            /* Cell::box, reference numbering i + j*b (upload default: -1) */
} apj_state;

const char* apj_version(void);
const char* apj_last_error(const apj_engine* e); /* e may be NULL: message of the failed apj_create */

/* Engine::Engine + Engine::topology scalars (jamming.cpp:119, :361-365): L[s] is the box length
 * of system s; b = floor(L/(2 rn)), lp = L/b, nbox = b*b are derived exactly as the reference. */
int apj_create(const apj_config* cfg, const double* L, apj_engine** out);
int apj_destroy(apj_engine* e);

/* lambda_s / lambda_n per system (Engine::CFself, Engine::CTnoise; jamming.cpp:49-50). */
int apj_set_activity(apj_engine* e, const double* CFself, const double* CTnoise);
/* relax() thermalisation ramp (jamming.cpp:516-520): for the next calls of apj_step, step k
 * (k = 0.. from now) runs with CFself = CFself_final - (tthermalize - k)*CFself_final/tthermalize
 * while k < tthermalize, then CFself_final. tthermalize = 0 cancels the ramp. */
int apj_set_ramp(apj_engine* e, int64_t tthermalize);

/* Replaces the state that initCells builds in vector<Cell> (jamming.cpp:285-354). Computes COM
 * from x_real in index order (calculate_COM, :761-774), sets COM_old = COM0 = COM unless
 * apj_set_com is called afterwards, then bins and builds the lists (assignCellsToGrid +
 * buildVerletLists, start() :184-185) WITHOUT touching x_old. */
int apj_upload_state(apj_engine* e, const apj_state* host);
int apj_download_state(apj_engine* e, apj_state* host);
/* Engine::initCells on the device (jamming.cpp:285-354, 2D branch), for systems too large to build vector<Cell> on the
 * host: radii 1 + N(0,1)/10, jittered offset-row lattice (spacing L / floor(sqrt N), even rows shifted by 1), polarity
 * U[-PI, PI), x_real = x0 = x_old = x, v = 0, COM; then assignCellsToGrid + buildVerletLists. The variates come from
 * the Philox stream keyed by `seed` (Box-Muller normals); the reference's own stream is time-seeded, so parity of the
 * initial condition is distributional. Two calls, because the box length depends on the radii:
 *   apj_lattice_box_length  L[s] = sqrt(PI * sum R^2 / dens[s])  (:305)  for the radii `seed` will produce
 *   apj_create(.., L, ..), then apj_init_lattice(e, seed)         (the same seed)                                 */
int apj_lattice_box_length(int64_t n, int32_t n_systems, uint64_t seed, const double* dens, int32_t device, double* L_out);
int apj_init_lattice(apj_engine* e, uint64_t seed);
/* Engine::topology (jamming.cpp:356-410) as a table, computed on the device: centres[2*nbox] = Box::center, neighbors[9*nbox] =
 * Box::neighbors, reference numbering p = i + j*b. (The kernels never read such a table: they derive both on the fly.) */
int apj_get_box_table(apj_engine* e, int32_t system, double* centres, int32_t* neighbors);
/* Cell::over as neighborInteractions leaves it on a filmed step (jamming.cpp:653-656) for the CURRENT state and lists:
 * 240 minus 240 |overlap| per overlapping pair, accumulated into an int in the reference's update order (every update
 * truncates). over[n_systems * n] by particle id; call it before the step whose frame print_video writes (:855-870). */
int apj_overlap_hue(apj_engine* e, int32_t* over);
/* Both copy the caller's arrays field by field (cudaMemcpyAsync on the handle's stream: page-locked caller
 * buffers move at PCIe speed, pageable ones work but are staged by the driver) through staging planes in
 * HBM; the AoS <-> SoA interleave, Rinv, cos/sin(phi), box renumbering and the by-id scatter run in kernels.
 *
 * 64-bit fingerprint of the current state: sum over particles of a hash of (system, id, bits of x, y, cos,
 * sin). It does not depend on the particle order in memory nor, on slab handles, on the decomposition: the
 * value of a slab handle is this rank's share, shares add modulo 2^64 to the value of the periodic handle. */
int apj_state_checksum(apj_engine* e, uint64_t* out);
/* COM / COM0 / COM_old of one system (jamming.cpp:89-91); NULL pointers are skipped. */
int apj_set_com(apj_engine* e, int32_t system, const double* com, const double* com0, const double* com_old);
int apj_get_com(apj_engine* e, int32_t system, double* com, double* com0, double* com_old);
/* start() :191-203: x_real = x0 = x, COM, COM0 = COM, saveOldPositions(). */
int apj_mark_origin(apj_engine* e);
/* The reference's very first step has no self term in the alignment sum (Cell::x_new is 0
 * until the first Cell::update, classes/Cell.h:75,102). 1 = drop the self term for the next
 * committed step only. */
int apj_skip_self_term_once(apj_engine* e, int32_t on);

/* n x Engine::calculate_next_positions (jamming.cpp:837-853): skin test (newSkinList :587-617),
 * on-device rebuild when it fires (assignCellsToGrid :527-548, buildVerletLists :550-585,
 * saveOldPositions :825-835), fused pair sweep (neighborInteractions :623-669) + Cell::update
 * (classes/Cell.h:92-118,157-175) + calculate_COM (:761-774). Noise: Philox4x32-10, counter =
 * (particle id, step), key = seed, mapped like boost::uniform_real on mt19937 (u/2^32*2PI-PI). */
int apj_step(apj_engine* e, int64_t n_steps);
/* One step in which randuni() for particle id of system s returns noise[s*N + id]
 * (replaces the RNG draw at jamming.cpp:667; parity gate). */
int apj_step_injected(apj_engine* e, const double* noise);
/* assignCellsToGrid + buildVerletLists without the skin bookkeeping (start() :184-185, :245-246). */
int apj_force_rebuild(apj_engine* e);
int apj_sync(apj_engine* e);

/* Checkpoint / restart (an addition: the reference never serialises its state). One binary file with the
 * per-system scalars (step, resetCounter, COM / COM0 / COM_old, activity, ramp) and every per-particle field
 * in original particle order. apj_load_checkpoint needs a handle of the same shape (n, n_systems, L, dt, rn,
 * rs); it restores the Philox key and step, re-bins and rebuilds the lists from the stored positions and keeps
 * x_old / COM_old, so the continuation follows the uninterrupted run to rounding (not bit for bit: the fresh
 * lists are ordered differently). On a slab handle both calls are [collective] and work on THIS rank's share:
 * every rank writes / reads its own file (owned particles with their ids; no gather of the box), and a file
 * can only be loaded by the rank of the decomposition that wrote it. */
int apj_save_checkpoint(apj_engine* e, const char* path);
int apj_load_checkpoint(apj_engine* e, const char* path);

/* out[8] = {step index, resetCounter (jamming.cpp:64), rebuilds incl. forced, longest list,
 *           overflow flag, kernel launches so far, discarded speculative steps, b*b} */
int apj_get_counters(apj_engine* e, int32_t system, int64_t* out8);
int apj_set_reset_counter(apj_engine* e, int32_t system, int64_t value); /* relax() :524 */
/* out[8] = {lanes per particle, threads per block, particles per block, tile capacity (slots),
 *           largest tile in use, dynamic shared memory per block (bytes), work blocks, steps per launch} */
int apj_get_tuning(apj_engine* e, int32_t* out8);
/* Skin-aware sweep length (a step sweeps only the list classes that can have come within rn given the
 * skin-test value of newSkinList, jamming.cpp:596-611; verified at commit, results identical to the full
 * sweep). out[8] = {launches dropped and re-run because the chosen length was too short, 1 if the
 * shortening is currently active, current skin-test value sqrt(l1)+sqrt(l2), class floor,
 * committed steps that swept 1, 2, 3, all 4 of the four build-distance classes}.
 * apj_set_sweep_truncation(e, 0) switches it off (every step sweeps the full lists). */
int apj_get_sweep_stats(apj_engine* e, int32_t system, double* out8);
int apj_set_sweep_truncation(apj_engine* e, int32_t on);
/* out[5] = {L, Lover2, lp, b, nbox}  (jamming.cpp:99-103) */
int apj_get_geometry(apj_engine* e, int32_t system, double* out5);

/* out[2] = {sum over particles of the full Verlet-list length (entries within rs; twice the number
 *           of pairs of Cell::VerletList), longest list}: n_full = out[0] / N is the figure the
 *           algorithmic-bytes roofline uses. */
int apj_list_stats(apj_engine* e, int32_t system, int64_t* out2);

/* Neighbour pair set of one system as a half list by particle id (the parity object of
 * SURVEY Q1): offsets[N+1], idx[cap] (partners j > i, ascending). *total receives the pair
 * count; idx may be NULL to query it. Replaces reading Cell::VerletList. */
int apj_get_pair_list(apj_engine* e, int32_t system, int64_t* offsets, int32_t* idx, int64_t cap, int64_t* total);
/* Box::CellList of every box (classes/Box.h:13), reference box numbering, ascending particle id:
 * offsets[nbox+1], idx[N]. */
int apj_get_cell_lists(apj_engine* e, int32_t system, int64_t* offsets, int32_t* idx);

/* ---- observables (per system: out arrays have n_systems entries / rows) ---- */
/* calculateOrderParameter + calculateSystemOrientation (jamming.cpp:776-806): order[s],
 * orient[2*s..]. */
int apj_order_orientation(apj_engine* e, double* order, double* orient);
/* Engine::MSD (jamming.cpp:808-823). */
int apj_msd(apj_engine* e, double* msd);
/* Inner loop of Fluctuations::measureFluctuations (classes/Fluctuations.h:62-76): total lens
 * area between every disk and a circle of radius[s] centred on COM. */
int apj_fluct_area(apj_engine* e, const double* radius, double* area);
/* Queued form of the four reductions above: the measurement is placed on the handle's stream behind the steps
 * already issued and returns at once; its raw sums (two doubles per system) are kept in a device-side ring of the
 * last 4096 tickets and apj_obs_fetch reads any run of them back with ONE copy and ONE synchronisation. The
 * reference's cadence (measureFluctuations every 10 steps, order / MSD every 100; jamming.cpp:211-239) then runs
 * without a host round trip per measurement. Raw sums per kind:
 *   APJ_OBS_COM    {sum x_real, sum y_real}                 (divide by N: calculate_COM, :761-774)
 *   APJ_OBS_ORDER  {sum vx/|v|, sum vy/|v|}                  (order = |.|/N :788, orientation = ./N :803)
 *   APJ_OBS_MSD    {sum of squared COM-corrected displacements, 0}   (divide by N, :822)
 *   APJ_OBS_FLUCT  {lens-area sum for radius param[s], 0}    (Fluctuations.h:62-76)                              */
#define APJ_OBS_COM 0
#define APJ_OBS_ORDER 1
#define APJ_OBS_MSD 2
#define APJ_OBS_FLUCT 3
int apj_obs_enqueue(apj_engine* e, int32_t kind, const double* param, int64_t* ticket);
int apj_obs_fetch(apj_engine* e, int64_t first_ticket, int64_t count, double* out /* count x n_systems x 2 */);
/* Correlations::spatialCorrelations raw sums (classes/Correlations.h:71-152) for one call:
 * counts[nc], ori_sum[nc], vel_sum[nc], pair_sum[np] per system (rows of nc / np), with
 * nc = ceil(cutoff/2.0), np = ceil(cutoff/0.1). Normalisation (:154-166) is left to the host. */
int apj_spatial_correlations(apj_engine* e, double cutoff, double* counts, double* ori_sum, double* vel_sum, double* pair_sum);
/* Correlations::velDist (classes/Correlations.h:179-187): counts per speed bin floor(v/dv[s]), 100 bins. */
int apj_vel_hist(apj_engine* e, const double* dv, int64_t* hist100);
/* Fluctuations::density_distribution (classes/Fluctuations.h:122-139): boxes per occupancy, 50 bins. */
int apj_occupancy_hist(apj_engine* e, int64_t* hist50);

/* CUDA-event timing on the handle's stream (bench.py; torch events cannot see this stream). */
int apj_timer_begin(apj_engine* e);
int apj_timer_end(apj_engine* e, float* milliseconds);
/* Per-kernel timing of the fused step kernel: launches n single steps with an event pair around
 * each launch of the step (step kernel + its fold/commit kernel) and returns the mean duration (ms). */
int apj_time_step_kernel(apj_engine* e, int64_t n, float* mean_ms, int64_t* committed);
/* The same, split: out3 = mean duration (ms) of {step kernel, fold + commit kernel, slab commit = all-rank rendezvous}
 * over n single steps ([collective] on slab handles). */
int apj_time_step_parts(apj_engine* e, int64_t n, float* out3);

/* ---- slab mode: ONE periodic box over the GPUs of a node (BASELINE config 4) --------
 * The global b x b cell grid (Engine::topology, jamming.cpp:356-480) is cut along x into slabs of
 * whole cell columns, rank r owning columns [r*b/nranks, (r+1)*b/nranks). Each rank runs one handle
 * on its GPU (one process per GPU, or several handles in one process). All exchange happens on the
 * device over peer memory (NVLink): the step kernel stores the new {x,y},{cos,sin} of its boundary
 * columns straight into the neighbours' ghost slots and every rank's {sum x_real, top-2 displacement}
 * partial into every rank's mailbox; a one-warp commit kernel folds the partials in rank order, so all
 * ranks take the reference's rebuild decision (newSkinList, jamming.cpp:587-617) on the same step; at a
 * rebuild, particles that left the slab are handed to the owning neighbour with peer atomics and the
 * ghost columns are re-sent. Results do not depend on the decomposition (Philox counters and list
 * order are keyed by particle id / global cell), except for the rounding of COM.
 * Calls marked [collective] must be made by all ranks (one host thread per rank) together. */
/* cfg->n is the particle count of the GLOBAL box, cfg->n_systems must be 1; capacity = particle slots of
 * this rank (0 -> n/nranks * 1.25 + four columns). */
int apj_slab_create(const apj_config* cfg, double L, int32_t rank, int32_t nranks, int64_t capacity, apj_engine** out);
/* out[8] = {rank, nranks, first owned column, owned columns, capacity, ghost-column capacity, owned particles, arena bytes} */
int apj_slab_info(apj_engine* e, int64_t* out8);
/* The peer-visible arena of this rank: a 64-byte cudaIpcMemHandle_t for ranks in other processes and
 * the local device pointer for handles in the same process. Either output may be NULL. */
int apj_slab_export(apj_engine* e, void* handle64, uint64_t* local_ptr);
/* Make rank `peer`'s arena addressable: same_process_ptr != 0 -> use that pointer (peer_device = its CUDA
 * device, for cudaDeviceEnablePeerAccess); otherwise open the IPC handle. */
int apj_slab_connect(apj_engine* e, int32_t peer, const void* handle64, uint64_t same_process_ptr, int32_t peer_device);
/* Device-side waits for a peer give up after this many seconds (default 20) and the call returns
 * APJ_E_STATE instead of hanging. */
int apj_slab_set_timeout(apj_engine* e, double seconds);
/* After every peer is connected: builds the step graph. */
int apj_slab_ready(apj_engine* e);
/* [collective] n_local particles (any subset; ids[k] = original particle index in [0, N)) in caller order.
 * Particles not in this rank's columns are migrated to the neighbour (adjacent slabs only). COM is NOT
 * derived: set it on every rank with apj_set_com (global mean of x_real). */
int apj_slab_upload(apj_engine* e, const apj_state* host, const int32_t* ids, int64_t n_local);
/* Owned particles in device (cell) order with their ids; *n_local receives the count (query with ids == NULL). */
int apj_slab_download(apj_engine* e, apj_state* host, int32_t* ids, int64_t cap, int64_t* n_local);
/* Pairs (id_i, id_j), id_j > id_i, of the owned particles' Verlet lists including ghost partners: the
 * union over ranks is the reference's half-list pair set (each pair exactly once). pairs[2*cap]. */
int apj_slab_get_pairs(apj_engine* e, int32_t* pairs, int64_t cap_pairs, int64_t* total);
/* Correlations::spatialCorrelations (classes/Correlations.h:71-152) on the decomposed box. Pairs among a rank's own
 * particles are its own; a pair that crosses a slab edge is counted by the rank of its LEFT particle, which needs the
 * neighbour's particles within the cutoff of the shared edge:
 *   apj_slab_export_edge           the owned particles within `width` of this slab's LEFT edge: 6 planes x | y | cos |
 *                                  sin | vx | vy of *n values (call with cap = 0 to size the buffer). Ownership only changes
 *                                  at a rebuild: use width = cutoff + skin (1.4) or more, so that partners which drifted
 *                                  over the edge since the last rebuild are covered;
 *   apj_slab_spatial_correlations  this rank's additive share of the raw sums, given the edge set of the rank to its right
 *                                  sorted by cell row, ext_row[b+1] = first particle of each row (b = cells per side).
 * The host ships the edge sets between ranks (slab.py: torch.distributed plumbing, 10 x per run in the reference's
 * cadence). Slabs must be wider than the cutoff + one cell (twice the cutoff on 2 ranks). */
int apj_slab_export_edge(apj_engine* e, double width, int64_t cap, double* out6, int64_t* n);
int apj_slab_spatial_correlations(apj_engine* e, double cutoff, int64_t n_ext, const double* ext6, const int32_t* ext_row,
                                  double* counts, double* ori_sum, double* vel_sum, double* pair_sum);
/* On a slab handle apj_step, apj_step_injected (noise indexed by global id), apj_force_rebuild and
 * apj_mark_origin are [collective]; apj_order_orientation (orient only), apj_msd, apj_fluct_area,
 * apj_vel_hist, apj_occupancy_hist and the COM left by apj_mark_origin are this rank's ADDITIVE share
 * (already divided by the global N where the reference divides): sum them over ranks. */

#ifdef __cplusplus
}
#endif
#endif /* APJ_B200_H */
