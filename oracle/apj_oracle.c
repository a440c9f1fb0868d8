/* TEST INFRASTRUCTURE ONLY (oracle/). Not product code: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library, and only as the
 * checker. Nothing under active_particle_jamming_b200/ links or calls it.
 *
 * Plain-C, single-thread restatement of the reference's 2D hot path over SoA arrays, following
 * the reference's operation order line by line so that -- compiled with the same compiler flags
 * (-ffp-contract=off, no fast-math) -- it is BIT-IDENTICAL to the reference build in oracle/_ref
 * on the same inputs (tests/test_oracle_vs_reference.py pins that here, where /root/reference
 * exists; tests/golden/ carries vectors produced by the reference build for the GPU box).
 *
 * Parity status: PINNED against outputs of the reference itself (oracle/_ref). The reference
 * owns no tests or golden vectors. RNG streams are NOT pinned (reference seeds
 * Boost mt19937 from the wall clock, jamming.cpp:36-37): noise is injected, or drawn from the
 * Philox4x32-10 restated below (Salmon et al., SC'11; Random123 constants).
 *
 * Citations are relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* code/jam/jamming.cpp:3-5 -- truncated literals are part of the algorithm. */
#define ORC_PI 3.14159265
#define ORC_PI2 6.28318531
#define ORC_SQRT2 1.41421356

typedef struct orc_sim {
    long N;
    double L, Lover2, lp;
    int b, nbox;
    double dt, rn, rs, rn2, rs2; /* jamming.cpp:57,112-115 */
    double CFself, CTnoise, dens;
    double COM[2], COM0[2], COM_old[2];
    long resetCounter;
    /* per particle, id order (fields of struct Cell, classes/Cell.h:15-43) */
    double *x, *y, *xr, *yr, *x0, *y0, *xo, *yo, *R, *Rinv, *phi, *cosp, *sinp, *vx, *vy, *Fx, *Fy, *xnew, *ynew;
    int* box;
    /* grid (struct Box, classes/Box.h:4-14) */
    double *cx, *cy; /* box centres */
    int* nbr;        /* nbox x 9 neighbour ids */
    int *cl_start, *cl_fill, *cl_items; /* Box::CellList as CSR, ascending id inside a box */
    /* half Verlet lists as CSR */
    long* vl_off; int* vl_idx; long vl_cap;
} orc_sim;

enum { F_X, F_Y, F_XR, F_YR, F_X0, F_Y0, F_XO, F_YO, F_R, F_RINV, F_PHI, F_COSP, F_SINP, F_VX, F_VY, F_FX, F_FY, F_XNEW, F_YNEW, F_COUNT };

orc_sim* orc_new(long N, double L, double dens) {
    orc_sim* s = (orc_sim*)calloc(1, sizeof(orc_sim));
    s->N = N; s->L = L; s->Lover2 = L / 2.0; s->dens = dens;
    s->dt = 0.1; s->rn = 2.8; s->rs = 1.5 * s->rn; s->rn2 = s->rn * s->rn; s->rs2 = s->rs * s->rs;
    double** f[F_COUNT] = {&s->x, &s->y, &s->xr, &s->yr, &s->x0, &s->y0, &s->xo, &s->yo, &s->R, &s->Rinv, &s->phi,
                           &s->cosp, &s->sinp, &s->vx, &s->vy, &s->Fx, &s->Fy, &s->xnew, &s->ynew};
    for (int k = 0; k < F_COUNT; k++) *f[k] = (double*)calloc((size_t)N, sizeof(double));
    s->box = (int*)malloc(sizeof(int) * (size_t)N);
    for (long i = 0; i < N; i++) { /* Cell::Cell(), classes/Cell.h:46-84 */
        s->box[i] = -1;
        s->x[i] = s->y[i] = s->xr[i] = s->yr[i] = s->x0[i] = s->y0[i] = s->xo[i] = s->yo[i] = -100;
        s->R[i] = -1.0;
    }
    s->vl_off = (long*)calloc((size_t)N + 1, sizeof(long));
    s->vl_cap = 16 * N + 64; s->vl_idx = (int*)malloc(sizeof(int) * (size_t)s->vl_cap);
    return s;
}
void orc_free(orc_sim* s) {
    double* f[F_COUNT] = {s->x, s->y, s->xr, s->yr, s->x0, s->y0, s->xo, s->yo, s->R, s->Rinv, s->phi,
                          s->cosp, s->sinp, s->vx, s->vy, s->Fx, s->Fy, s->xnew, s->ynew};
    for (int k = 0; k < F_COUNT; k++) free(f[k]);
    free(s->box); free(s->cx); free(s->cy); free(s->nbr); free(s->cl_start); free(s->cl_fill); free(s->cl_items);
    free(s->vl_off); free(s->vl_idx); free(s);
}
double* orc_field(orc_sim* s, int id) {
    double* f[F_COUNT] = {s->x, s->y, s->xr, s->yr, s->x0, s->y0, s->xo, s->yo, s->R, s->Rinv, s->phi,
                          s->cosp, s->sinp, s->vx, s->vy, s->Fx, s->Fy, s->xnew, s->ynew};
    return (id >= 0 && id < F_COUNT) ? f[id] : 0;
}
int* orc_box(orc_sim* s) { return s->box; }
/* scalars: {L, Lover2, lp, b, nbox, COM[2], COM0[2], COM_old[2], resetCounter, -, CFself, CTnoise} */
void orc_get_scalars(orc_sim* s, double* o) {
    o[0] = s->L; o[1] = s->Lover2; o[2] = s->lp; o[3] = s->b; o[4] = s->nbox;
    o[5] = s->COM[0]; o[6] = s->COM[1]; o[7] = s->COM0[0]; o[8] = s->COM0[1];
    o[9] = s->COM_old[0]; o[10] = s->COM_old[1]; o[11] = (double)s->resetCounter; o[12] = 0; o[13] = s->CFself; o[14] = s->CTnoise;
}
void orc_set_com(orc_sim* s, const double* com, const double* com0, const double* com_old) {
    if (com) { s->COM[0] = com[0]; s->COM[1] = com[1]; }
    if (com0) { s->COM0[0] = com0[0]; s->COM0[1] = com0[1]; }
    if (com_old) { s->COM_old[0] = com_old[0]; s->COM_old[1] = com_old[1]; }
}
void orc_set_params(orc_sim* s, double CFself, double CTnoise) { s->CFself = CFself; s->CTnoise = CTnoise; }
void orc_set_reset_counter(orc_sim* s, long v) { s->resetCounter = v; }

/* jamming.cpp:305 -- L = sqrt(PI * sum R^2 / dens), summed in index order. */
double orc_box_length(const double* R, long N, double dens) {
    double volume = 0;
    for (long i = 0; i < N; i++) volume += R[i] * R[i];
    return sqrt(ORC_PI * volume / dens);
}

/* jamming.cpp:872-880 (Engine::delta_norm; identical copies in Fluctuations.h:146, Correlations.h:207) */
static inline double delta_norm(const orc_sim* s, double delta) {
    int k = -1;
    if (delta < -s->Lover2) k = 1;
    while (delta < -s->Lover2 || delta >= s->Lover2) delta += k * s->L;
    return delta;
}
double orc_delta_norm(const orc_sim* s, double d) { return delta_norm(s, d); }

/* Cell::periodicAngles / Cell::PBC, classes/Cell.h:160-175 (single wrap, truncated constants). */
static inline double wrap_angle(double phi) {
    if (phi >= ORC_PI) phi -= ORC_PI2;
    else if (phi < -ORC_PI) phi += ORC_PI2;
    return phi;
}
static inline double wrap_pos(const orc_sim* s, double x) {
    if (x >= s->Lover2) x -= s->L;
    else if (x < -s->Lover2) x += s->L;
    return x;
}
/* initCells tail (jamming.cpp:331-353): cosp/sinp from phi, then periodicAngles + PBC. */
void orc_finish_init(orc_sim* s) {
    for (long i = 0; i < s->N; i++) {
        s->Rinv[i] = 1.0 / s->R[i];
        s->cosp[i] = cos(s->phi[i]); s->sinp[i] = sin(s->phi[i]);
        s->phi[i] = wrap_angle(s->phi[i]);
        s->x[i] = wrap_pos(s, s->x[i]); s->y[i] = wrap_pos(s, s->y[i]);
    }
}

/* Engine::topology, 2D branch: jamming.cpp:361-365 (b, lp), :373-390 (centres), :391-408 (neighbours). */
void orc_topology(orc_sim* s) {
    double lp = 2 * s->rn;
    int b = (int)floor(s->L / lp);
    s->b = b; s->nbox = b * b;
    s->lp = s->L / floor(s->L / lp);
    free(s->cx); free(s->cy); free(s->nbr); free(s->cl_start); free(s->cl_fill); free(s->cl_items);
    s->cx = (double*)malloc(sizeof(double) * (size_t)s->nbox); s->cy = (double*)malloc(sizeof(double) * (size_t)s->nbox);
    s->nbr = (int*)malloc(sizeof(int) * 9 * (size_t)s->nbox);
    s->cl_start = (int*)calloc((size_t)s->nbox + 1, sizeof(int)); s->cl_fill = (int*)calloc((size_t)s->nbox, sizeof(int));
    s->cl_items = (int*)malloc(sizeof(int) * (size_t)s->N);
    for (int i = 0; i < b; i++)
        for (int j = 0; j < b; j++) {
            int p = i + (j * b);
            double min0 = -s->Lover2 + i * s->L / b, min1 = -s->Lover2 + j * s->L / b;
            double max0 = -s->Lover2 + (i + 1) * s->L / b, max1 = -s->Lover2 + (j + 1) * s->L / b;
            s->cx[p] = (min0 + max0) / 2.; s->cy[p] = (min1 + max1) / 2.;
        }
    for (int k = 0; k < s->nbox; k++) {
        int vi = k % b, vj = k / b;
        for (int j = 0; j < 3; j++)
            for (int i = 0; i < 3; i++) {
                int p = i, q = j;
                if (vi == 0 && i == 0) p = p + b;
                if (vi == b - 1 && i == 2) p = p - b;
                if (vj == 0 && j == 0) q = q + b;
                if (vj == b - 1 && j == 2) q = q - b;
                s->nbr[k * 9 + i + (j * 3)] = k + (p - 1) + (q - 1) * b;
            }
    }
}

static void fill_cell_lists(orc_sim* s) { /* CellList.push_back(i) in ascending i, jamming.cpp:546 */
    memset(s->cl_start, 0, sizeof(int) * ((size_t)s->nbox + 1));
    for (long i = 0; i < s->N; i++) s->cl_start[s->box[i] + 1]++;
    for (int p = 0; p < s->nbox; p++) s->cl_start[p + 1] += s->cl_start[p];
    memset(s->cl_fill, 0, sizeof(int) * (size_t)s->nbox);
    for (long i = 0; i < s->N; i++) { int p = s->box[i]; s->cl_items[s->cl_start[p] + s->cl_fill[p]++] = (int)i; }
}

/* Engine::assignCellsToGrid verbatim in structure: O(N*nbox) nearest-centre search with strict <
 * from r2 = lp^2 * 0.25 * NDIM (jamming.cpp:527-548). Small N only. */
void orc_assign_bruteforce(orc_sim* s) {
    for (long i = 0; i < s->N; i++) {
        double r2 = s->lp * s->lp * 0.25 * 2;
        for (int j = 0; j < s->nbox; j++) {
            double d2 = 0.0;
            double dr = s->x[i] - s->cx[j]; d2 += dr * dr;
            dr = s->y[i] - s->cy[j]; d2 += dr * dr;
            if (d2 < r2) { r2 = d2; s->box[i] = j; }
        }
    }
    fill_cell_lists(s);
}
/* Same result in O(N): only the 3x3 boxes around the floor-binned box can win the strict-<
 * search; they are visited in ascending box index like the full scan (ties -> lowest index). */
void orc_assign(orc_sim* s) {
    int b = s->b;
    for (long i = 0; i < s->N; i++) {
        int gx = (int)floor((s->x[i] + s->Lover2) / s->lp), gy = (int)floor((s->y[i] + s->Lover2) / s->lp);
        double r2 = s->lp * s->lp * 0.25 * 2;
        for (int qy = gy - 1; qy <= gy + 1; qy++) {
            if (qy < 0 || qy >= b) continue;
            for (int qx = gx - 1; qx <= gx + 1; qx++) {
                if (qx < 0 || qx >= b) continue;
                int j = qx + qy * b;
                double d2 = 0.0;
                double dr = s->x[i] - s->cx[j]; d2 += dr * dr;
                dr = s->y[i] - s->cy[j]; d2 += dr * dr;
                if (d2 < r2) { r2 = d2; s->box[i] = j; }
            }
        }
    }
    fill_cell_lists(s);
}

/* Engine::buildVerletLists, jamming.cpp:550-585. Net effect is the half list {j>i, d2<rs2} in
 * neighbour-box (m = 0..8) then CellList order. */
void orc_build_verlet(orc_sim* s) {
    long tot = 0;
    for (long i = 0; i < s->N; i++) {
        s->vl_off[i] = tot;
        for (int m = 0; m < 9; m++) {
            int p = s->nbr[s->box[i] * 9 + m];
            for (int k = s->cl_start[p]; k < s->cl_start[p + 1]; k++) {
                int j = s->cl_items[k];
                if (j > i) {
                    double d2 = 0.0;
                    double dx = delta_norm(s, s->x[j] - s->x[i]);
                    double dy = delta_norm(s, s->y[j] - s->y[i]);
                    d2 += dx * dx + dy * dy;
                    if (d2 < s->rs2) {
                        if (tot >= s->vl_cap) { s->vl_cap *= 2; s->vl_idx = (int*)realloc(s->vl_idx, sizeof(int) * (size_t)s->vl_cap); }
                        s->vl_idx[tot++] = j;
                    }
                }
            }
        }
    }
    s->vl_off[s->N] = tot;
}
long orc_get_verlet(orc_sim* s, long* offsets, int* idx, long cap) {
    if (offsets) memcpy(offsets, s->vl_off, sizeof(long) * ((size_t)s->N + 1));
    long tot = s->vl_off[s->N];
    if (idx) memcpy(idx, s->vl_idx, sizeof(int) * (size_t)(tot < cap ? tot : cap));
    return tot;
}
long orc_get_cell_lists(orc_sim* s, long* offsets, int* idx) {
    for (int p = 0; p <= s->nbox; p++) offsets[p] = s->cl_start[p];
    memcpy(idx, s->cl_items, sizeof(int) * (size_t)s->N);
    return s->N;
}
void orc_get_box_neighbors(orc_sim* s, int* out9) { memcpy(out9, s->nbr, sizeof(int) * 9 * (size_t)s->nbox); }

/* Engine::saveOldPositions, jamming.cpp:825-835 */
void orc_save_old(orc_sim* s) {
    s->COM_old[0] = s->COM[0]; s->COM_old[1] = s->COM[1];
    for (long i = 0; i < s->N; i++) { s->xo[i] = s->x[i]; s->yo[i] = s->y[i]; }
}
/* Engine::newSkinList, jamming.cpp:587-617 (COM-drift corrected, top two displacements; Q7). */
int orc_new_skin_list(orc_sim* s) {
    int refresh = 0;
    double largest2 = 0., second2 = 0.;
    for (long i = 0; i < s->N; i++) {
        double d2 = 0.0;
        double dx = delta_norm(s, s->x[i] - s->xo[i] - s->COM[0] + s->COM_old[0]);
        double dy = delta_norm(s, s->y[i] - s->yo[i] - s->COM[1] + s->COM_old[1]);
        d2 += dx * dx + dy * dy;
        if (d2 > largest2) { second2 = largest2; largest2 = d2; }
        else if (d2 > second2) { second2 = d2; }
    }
    if ((sqrt(largest2) + sqrt(second2)) > (s->rs - s->rn)) {
        s->resetCounter++;
        orc_save_old(s);
        refresh = 1;
    }
    return refresh;
}
/* Engine::neighborInteractions, 2D branch, jamming.cpp:623-669. noise[i] is the value randuni()
 * returns for particle i (one draw per particle in ascending index, :667). */
void orc_neighbor_interactions(orc_sim* s, const double* noise) {
    for (long i = 0; i < s->N; i++) {
        for (long k = s->vl_off[i]; k < s->vl_off[i + 1]; k++) {
            int j = s->vl_idx[k];
            if (j > i) {
                double dx = delta_norm(s, s->x[j] - s->x[i]);
                double dy = delta_norm(s, s->y[j] - s->y[i]);
                double d2 = dx * dx + dy * dy;
                if (d2 < s->rn2) {
                    double sumR = s->R[i] + s->R[j];
                    if (d2 < sumR * sumR) {
                        double overlap = sumR / sqrt(d2) - 1;
                        double fx = overlap * dx, fy = overlap * dy;
                        s->Fx[i] -= fx; s->Fx[j] += fx;
                        s->Fy[i] -= fy; s->Fy[j] += fy;
                    }
                    s->xnew[i] += s->cosp[j]; s->ynew[i] += s->sinp[j];
                    s->xnew[j] += s->cosp[i]; s->ynew[j] += s->sinp[i];
                }
            }
        }
        s->phi[i] = atan2(s->ynew[i], s->xnew[i]) + s->CTnoise * noise[i];
    }
}
/* Cell::update, 2D branch, classes/Cell.h:92-118 + PBC :157 (SURVEY Q6 order). */
/* Cell::over on a filmed step (jamming.cpp:653-656): int over -= 240*abs(overlap) for both partners of every
 * overlapping pair, in the order neighborInteractions visits the half lists; `over` is an int, so every update
 * truncates toward zero. Starts from 240 (print_video, :866). */
void orc_overlap_hue(const orc_sim* s, int* over) {
    for (long i = 0; i < s->N; i++) over[i] = 240;
    for (long i = 0; i < s->N; i++)
        for (long k = s->vl_off[i]; k < s->vl_off[i + 1]; k++) {
            long j = s->vl_idx[k];
            double dx = delta_norm(s, s->x[j] - s->x[i]);
            double dy = delta_norm(s, s->y[j] - s->y[i]);
            double d2 = dx * dx + dy * dy;
            if (d2 < s->rn2) {
                double sumR = s->R[i] + s->R[j];
                if (d2 < sumR * sumR) {
                    double overlap = sumR / sqrt(d2) - 1;
                    over[i] = (int)(over[i] - 240 * fabs(overlap));
                    over[j] = (int)(over[j] - 240 * fabs(overlap));
                }
            }
        }
}

void orc_update(orc_sim* s) {
    for (long i = 0; i < s->N; i++) {
        s->phi[i] = wrap_angle(s->phi[i]);
        s->cosp[i] = cos(s->phi[i]); s->sinp[i] = sin(s->phi[i]);
        s->Fx[i] += s->cosp[i] * s->CFself * s->R[i];
        s->Fy[i] += s->sinp[i] * s->CFself * s->R[i];
        s->xnew[i] = s->cosp[i]; s->ynew[i] = s->sinp[i];
        s->vx[i] = s->Fx[i] * s->Rinv[i]; s->vy[i] = s->Fy[i] * s->Rinv[i];
        double dx = s->vx[i] * s->dt, dy = s->vy[i] * s->dt;
        s->x[i] += dx; s->y[i] += dy;
        s->xr[i] += dx; s->yr[i] += dy;
        s->Fx[i] = 0.0; s->Fy[i] = 0.0;
        s->x[i] = wrap_pos(s, s->x[i]); s->y[i] = wrap_pos(s, s->y[i]);
    }
}
/* Engine::calculate_COM, jamming.cpp:761-774 */
void orc_calculate_com(orc_sim* s) {
    for (int k = 0; k < 2; k++) {
        const double* a = k ? s->yr : s->xr;
        s->COM[k] = 0.0;
        for (long i = 0; i < s->N; i++) s->COM[k] += a[i];
        s->COM[k] /= s->N;
    }
}
/* Engine::calculate_next_positions, jamming.cpp:837-853. fast_assign=0 uses the O(N*nbox) scan. */
int orc_step(orc_sim* s, const double* noise, int fast_assign) {
    int rebuilt = 0;
    if (orc_new_skin_list(s)) {
        if (fast_assign) orc_assign(s); else orc_assign_bruteforce(s);
        orc_build_verlet(s);
        rebuilt = 1;
    }
    orc_neighbor_interactions(s, noise);
    orc_update(s);
    orc_calculate_com(s);
    return rebuilt;
}
/* the three statements between relax() and the main loop, jamming.cpp:191-203 */
void orc_mark_origin(orc_sim* s) {
    for (long i = 0; i < s->N; i++) { s->xr[i] = s->x[i]; s->x0[i] = s->x[i]; s->yr[i] = s->y[i]; s->y0[i] = s->y[i]; }
    orc_calculate_com(s);
    s->COM0[0] = s->COM[0]; s->COM0[1] = s->COM[1];
    orc_save_old(s);
}

/* ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11; Random123 constants). The GPU path
 * keys noise as counter = (particle id, step_lo, step_hi, replica), key = (seed_lo, seed_hi)
 * and uses output word 0. Uniform mapping mirrors Boost 1.64 generate_uniform_real over a
 * 32-bit engine: u / 2^32 * (PI - (-PI)) + (-PI)  (SURVEY Q5). ---- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; r++) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
double orc_u32_to_randuni(uint32_t u) { return (double)u / 4294967296.0 * (ORC_PI - (-ORC_PI)) + (-ORC_PI); }
void orc_philox_noise(uint64_t seed, uint64_t step, uint32_t replica, long N, double* out) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (long i = 0; i < N; i++) {
        uint32_t ctr[4] = {(uint32_t)i, (uint32_t)step, (uint32_t)(step >> 32), replica}, o[4];
        orc_philox4x32_10(ctr, key, o);
        out[i] = orc_u32_to_randuni(o[0]);
    }
}
/* n steps with Philox noise; returns number of rebuilds */
long orc_run_philox(orc_sim* s, uint64_t seed, uint64_t first_step, uint32_t replica, long n, int fast_assign) {
    double* noise = (double*)malloc(sizeof(double) * (size_t)s->N);
    long rebuilds = 0;
    for (long k = 0; k < n; k++) {
        orc_philox_noise(seed, first_step + (uint64_t)k, replica, s->N, noise);
        rebuilds += orc_step(s, noise, fast_assign);
    }
    free(noise);
    return rebuilds;
}

/* ---- on-engine observables, jamming.cpp:776-823 ---- */
static inline double speed(const orc_sim* s, long i) { return sqrt(s->vx[i] * s->vx[i] + s->vy[i] * s->vy[i]); } /* Cell.h:177-181 */
void orc_orientation_sum(const orc_sim* s, double* o) { /* shared loop of :776-806 */
    o[0] = o[1] = 0.0;
    for (long i = 0; i < s->N; i++) {
        double inverseVel = 1.0 / speed(s, i);
        o[0] += s->vx[i] * inverseVel; o[1] += s->vy[i] * inverseVel;
    }
}
double orc_order(const orc_sim* s) { double o[2]; orc_orientation_sum(s, o); return sqrt(o[0] * o[0] + o[1] * o[1] + 0.0 * 0.0) / (double)s->N; }
void orc_orientation(const orc_sim* s, double* o) { orc_orientation_sum(s, o); o[0] /= (double)s->N; o[1] /= (double)s->N; }
double orc_msd(const orc_sim* s) { /* :808-823 */
    double MSD = 0.0;
    for (long i = 0; i < s->N; i++) {
        double dx = s->xr[i] - s->x0[i] - s->COM[0] + s->COM0[0];
        double dy = s->yr[i] - s->y0[i] - s->COM[1] + s->COM0[1];
        double dz = 0.0;
        MSD += dx * dx + dy * dy + dz * dz;
    }
    return MSD / s->N;
}

/* ---- Fluctuations, classes/Fluctuations.h ---- */
double orc_fluct_overlap(double r, double R, double d) { /* :89-120, 2D */
    if (R >= r + d) return ORC_PI * r * r;
    double R12 = r * r, R22 = R * R;
    double x = (R12 - R22 + d * d) / (2.0 * d);
    double theta = acos(x / r);
    double A = R12 * theta - x * r * sin(theta);
    x = d - x;
    theta = acos(x / R);
    A += R22 * theta - x * R * sin(theta);
    return A;
}
/* inner loop of measureFluctuations (:62-76): total intersecting area V for one radius */
double orc_fluct_area(const orc_sim* s, double current_radius) {
    double V = 0.0;
    for (long i = 0; i < s->N; i++) {
        double sumR = s->R[i] + current_radius;
        double d2 = 0.0;
        double dr = delta_norm(s, s->xr[i] - s->COM[0]); d2 += dr * dr;
        dr = delta_norm(s, s->yr[i] - s->COM[1]); d2 += dr * dr;
        if (d2 <= sumR * sumR) V += orc_fluct_overlap(s->R[i], current_radius, sqrt(d2));
    }
    return V;
}
typedef struct { double current_radius, current_value, rad_interval, time_interval, dens; int counter; } orc_fluct;
void orc_fluct_init(orc_fluct* f, double L, int totalSteps, int skip, double dens) { /* :34-49 */
    f->dens = dens; f->current_radius = 3.0;
    f->rad_interval = pow((L / 2.0) / 3.0, 1. / ((double)10 + 1.0));
    f->time_interval = (double)totalSteps / ((double)skip * (double)10);
    f->counter = 0; f->current_value = 0;
}
/* measureFluctuations state machine (:51-87). Returns 1 and fills out2 =
 * {expectedV, rms} when the call flushes a point to fluct.dat (Print.h:174-176). */
int orc_fluct_measure(orc_fluct* f, const orc_sim* s, double* out2) {
    double expectedV = f->dens * ORC_PI * f->current_radius * f->current_radius;
    int flushed = 0;
    if (f->counter < f->time_interval) {
        double V = orc_fluct_area(s, f->current_radius);
        f->current_value += (V - expectedV) * (V - expectedV);
    } else {
        f->current_value = sqrt(f->current_value / (double)f->counter);
        out2[0] = expectedV; out2[1] = f->current_value; flushed = 1;
        f->current_value = 0; f->counter = -1;
        f->current_radius = f->current_radius * f->rad_interval;
    }
    f->counter++;
    return flushed;
}
void orc_density_distribution(const orc_sim* s, double* dist50) { /* :122-139 (accumulates) */
    for (int j = 0; j < s->nbox; j++) {
        double count = s->cl_start[j + 1] - s->cl_start[j];
        int bin = (int)floor(count / 1);
        if (bin < 50) dist50[bin] += 1.0;
    }
}

/* ---- Correlations, classes/Correlations.h ---- */
/* spatialCorrelations (:71-167) over the boxPairs list of topology (jamming.cpp:460-479):
 * all p<=q with centre distance < cutoff + sqrt2*lp, visited p-major; accumulates into
 * vel_nc / ori_nc / pair_np exactly like the reference (0/0 -> NaN for empty bins, Q14). */
void orc_spatial_correlations(const orc_sim* s, double cutoff, double* vel_nc, double* ori_nc, double* pair_np) {
    double dr_c = 2.0, dr_p = 0.1;
    int np = (int)ceil(cutoff / dr_p), nc = (int)ceil(cutoff / dr_c);
    double norm = 2 * s->L * s->L / (2 * ORC_PI * dr_p * (double)(s->N * s->N));
    double* pairTemp = (double*)calloc((size_t)np, sizeof(double));
    double* corrTemp = (double*)calloc((size_t)nc, sizeof(double));
    double* velTemp = (double*)calloc((size_t)nc, sizeof(double));
    double* counts = (double*)calloc((size_t)nc, sizeof(double));
    double boxCutoff = cutoff + (ORC_SQRT2 * s->lp);
    for (int p = 0; p < s->nbox; p++)
        for (int q = p; q < s->nbox; q++) {
            double boxDist2 = 0.0;
            double dr = delta_norm(s, s->cx[p] - s->cx[q]); boxDist2 += dr * dr;
            dr = delta_norm(s, s->cy[p] - s->cy[q]); boxDist2 += dr * dr;
            if (!(boxDist2 < boxCutoff * boxCutoff)) continue;
            for (int m = s->cl_start[p]; m < s->cl_start[p + 1]; m++)
                for (int n = s->cl_start[q]; n < s->cl_start[q + 1]; n++) {
                    int i = s->cl_items[m], j = s->cl_items[n];
                    if (p != q || (p == q && j > i)) {
                        double r = 0.0;
                        double dk = delta_norm(s, s->x[j] - s->x[i]); r += dk * dk;
                        dk = delta_norm(s, s->y[j] - s->y[i]); r += dk * dk;
                        r = sqrt(r);
                        int binp = (int)floor(r / dr_p), binc = (int)floor(r / dr_c);
                        if (binp < np) pairTemp[binp] += 1.0 / r;
                        if (binc < nc) {
                            counts[binc] += 1.0;
                            corrTemp[binc] += s->cosp[i] * s->cosp[j] + s->sinp[i] * s->sinp[j];
                            velTemp[binc] += (s->vx[i] * s->vx[j] + s->vy[i] * s->vy[j]) / (speed(s, i) * speed(s, j));
                        }
                    }
                }
        }
    for (int k = 0; k < nc; k++) {
        corrTemp[k] /= counts[k]; velTemp[k] /= counts[k];
        ori_nc[k] += corrTemp[k]; vel_nc[k] += velTemp[k];
    }
    for (int k = 0; k < np; k++) { pairTemp[k] *= norm; pair_np[k] += pairTemp[k]; }
    free(pairTemp); free(corrTemp); free(velTemp); free(counts);
}
void orc_vel_dist(const orc_sim* s, double CFself_for_bins, double* out100) { /* :179-187, dv = CFself/50 (:54) */
    double dv = CFself_for_bins / 50.0;
    for (long i = 0; i < s->N; i++) {
        double v = speed(s, i);
        int binv = (int)floor(v / dv);
        if (binv < 100) out100[binv] += 1.0 / (double)s->N;
    }
}
