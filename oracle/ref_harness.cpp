// TEST INFRASTRUCTURE ONLY (oracle/). Not product code: nothing under
// active_particle_jamming_b200/ may include, link or call this.
//
// C-callable harness around the UNMODIFIED reference translation unit
// (/root/reference/code/jam/jamming.cpp + code/classes/*.h). oracle/Makefile compiles the
// reference sources where they lie, with exactly two mechanical substitutions made on the fly
// (nothing is copied into the repo):
//   * `#define NDIM 3` -> `#define NDIM 2`  (jamming.cpp:2; north_star is the 2D branch)
//   * `main` renamed on the command line (-Dmain=apj_reference_main) so the TU can live in a .so
// and the Boost.Random shim in oracle/shim/ (Boost is un-vendored; only RNG lives there).
// This file is appended to that TU via APJ_REF_TU so it can reach `Engine`, `randuni`, `randnorm`.
//
// Everything here only *drives* reference functions (Engine::topology, assignCellsToGrid,
// buildVerletLists, calculate_next_positions, ...) and copies state in/out.
#ifndef APJ_REF_TU
#error "compile through oracle/Makefile (APJ_REF_TU must point at the NDIM=2 reference TU)"
#endif
#include APJ_REF_TU

#include <cstring>
#include <cstdlib>

#if NDIM != 2
#error "oracle/_ref must be the 2D branch"
#endif

namespace {
struct RefBox {
    Engine* eng;
    Print* printer = nullptr;
    Fluctuations* fluct = nullptr;
    Correlations* corr = nullptr;
};
}  // namespace

extern "C" {

int apjref_ndim() { return NDIM; }
double apjref_PI() { return PI; }
double apjref_PI2() { return PI2; }

// Re-seed the two by-value RNG copies (jamming.cpp:40-41) so initCells()/relax() are reproducible.
void apjref_seed(unsigned seed) {
    randuni.oracle_reseed(seed);
    randnorm.oracle_reseed(seed);
    gen.seed(seed);
}
void apjref_inject_uniform(const double* values, long n) { randuni.oracle_inject(values, (size_t)n); }
void apjref_clear_injection() { randuni.oracle_clear_injection(); }
long apjref_injected_consumed() { return (long)randuni.oracle_injected_consumed(); }
double apjref_randuni() { return randuni(); }
double apjref_randnorm() { return randnorm(); }

void* apjref_new(long n, long steps, double l_s, double l_n, double rho) {
    RefBox* b = new RefBox;
    b->eng = new Engine("apjref", "run0", n, steps, l_s, l_n, rho);
    return b;
}
void apjref_delete(void* h) {
    RefBox* b = (RefBox*)h;
    delete b->corr; delete b->fluct; delete b->printer;
    delete b->eng;
    delete b;
}
static Engine& E(void* h) { return *((RefBox*)h)->eng; }

// Reference initial condition + grid (jamming.cpp:285-354, :356-480).
void apjref_init_cells(void* h) { E(h).initCells(); }
void apjref_topology(void* h) { E(h).topology(); }

// Build the particle vector from caller-supplied radii / positions / angles the way initCells
// does (index, R, Rinv=1/R, L, Lover2, dt; cosp/sinp from phi; jamming.cpp:291-334), with L from
// the reference formula (:305) unless L_override > 0.
void apjref_set_particles(void* h, const double* R, const double* x, const double* y,
                          const double* phi, double L_override) {
    Engine& e = E(h);
    e.cell.clear();
    double volume = 0;
    for (int i = 0; i < e.N; i++) {
        Cell c;
        e.cell.push_back(c);
        e.cell[i].index = i;
        e.cell[i].R = R[i];
        e.cell[i].Rinv = 1.0 / R[i];
        volume += R[i] * R[i];
    }
    e.L = L_override > 0 ? L_override : sqrt(PI * volume / e.dens);
    e.Lover2 = e.L / 2.0;
    for (int i = 0; i < e.N; i++) {
        Cell& c = e.cell[i];
        c.L = e.L; c.Lover2 = e.Lover2; c.dt = e.dt;
        c.x[0] = x[i]; c.x[1] = y[i];
        c.theta = PI / 2.0;
        c.phi = phi[i];
        c.cosp = cos(c.phi); c.sinp = sin(c.phi);
        c.periodicAngles();
        c.PBC();
    }
}

// Field-wise state access. Any pointer may be NULL. Layout: one value per particle, id order.
struct apjref_state {
    double *x, *y, *xr, *yr, *x0, *y0, *xo, *yo, *R, *phi, *cosp, *sinp, *vx, *vy, *xnew, *ynew, *Fx, *Fy;
    int* box;
};
void apjref_get_state(void* h, apjref_state* s) {
    Engine& e = E(h);
    for (int i = 0; i < e.N; i++) {
        Cell& c = e.cell[i];
        if (s->x) s->x[i] = c.x[0];       if (s->y) s->y[i] = c.x[1];
        if (s->xr) s->xr[i] = c.x_real[0]; if (s->yr) s->yr[i] = c.x_real[1];
        if (s->x0) s->x0[i] = c.x0[0];     if (s->y0) s->y0[i] = c.x0[1];
        if (s->xo) s->xo[i] = c.x_old[0];  if (s->yo) s->yo[i] = c.x_old[1];
        if (s->R) s->R[i] = c.R;           if (s->phi) s->phi[i] = c.phi;
        if (s->cosp) s->cosp[i] = c.cosp;  if (s->sinp) s->sinp[i] = c.sinp;
        if (s->vx) s->vx[i] = c.vx;        if (s->vy) s->vy[i] = c.vy;
        if (s->xnew) s->xnew[i] = c.x_new; if (s->ynew) s->ynew[i] = c.y_new;
        if (s->Fx) s->Fx[i] = c.Fx;        if (s->Fy) s->Fy[i] = c.Fy;
        if (s->box) s->box[i] = c.box;
    }
}
void apjref_set_state(void* h, const apjref_state* s) {
    Engine& e = E(h);
    for (int i = 0; i < e.N; i++) {
        Cell& c = e.cell[i];
        if (s->x) c.x[0] = s->x[i];        if (s->y) c.x[1] = s->y[i];
        if (s->xr) c.x_real[0] = s->xr[i]; if (s->yr) c.x_real[1] = s->yr[i];
        if (s->x0) c.x0[0] = s->x0[i];     if (s->y0) c.x0[1] = s->y0[i];
        if (s->xo) c.x_old[0] = s->xo[i];  if (s->yo) c.x_old[1] = s->yo[i];
        if (s->R) { c.R = s->R[i]; c.Rinv = 1.0 / c.R; }
        if (s->phi) c.phi = s->phi[i];
        if (s->cosp) c.cosp = s->cosp[i];  if (s->sinp) c.sinp = s->sinp[i];
        if (s->vx) c.vx = s->vx[i];        if (s->vy) c.vy = s->vy[i];
        if (s->xnew) c.x_new = s->xnew[i]; if (s->ynew) c.y_new = s->ynew[i];
        if (s->Fx) c.Fx = s->Fx[i];        if (s->Fy) c.Fy = s->Fy[i];
        if (s->box) c.box = s->box[i];
    }
}

// Scalars: out = {L, Lover2, lp, b, nbox, COM[2], COM0[2], COM_old[2], resetCounter, t, CFself, CTnoise}
void apjref_get_scalars(void* h, double* out) {
    Engine& e = E(h);
    out[0] = e.L; out[1] = e.Lover2; out[2] = e.lp; out[3] = e.b; out[4] = e.nbox;
    out[5] = e.COM[0]; out[6] = e.COM[1]; out[7] = e.COM0[0]; out[8] = e.COM0[1];
    out[9] = e.COM_old[0]; out[10] = e.COM_old[1]; out[11] = (double)e.resetCounter;
    out[12] = (double)e.t; out[13] = e.CFself; out[14] = e.CTnoise;
}
void apjref_set_com(void* h, const double* com, const double* com0, const double* com_old) {
    Engine& e = E(h);
    if (com) { e.COM[0] = com[0]; e.COM[1] = com[1]; }
    if (com0) { e.COM0[0] = com0[0]; e.COM0[1] = com0[1]; }
    if (com_old) { e.COM_old[0] = com_old[0]; e.COM_old[1] = com_old[1]; }
}
void apjref_set_params(void* h, double CFself, double CTnoise) { E(h).CFself = CFself; E(h).CTnoise = CTnoise; }
void apjref_set_reset_counter(void* h, long v) { E(h).resetCounter = v; }

// Hot-path pieces, exactly as the reference exposes them (jamming.cpp:71-87).
void apjref_assign(void* h) { E(h).assignCellsToGrid(); }
void apjref_build(void* h) { E(h).buildVerletLists(); }
int apjref_new_skin_list(void* h) { return E(h).newSkinList() ? 1 : 0; }
void apjref_neighbor_interactions(void* h) { E(h).neighborInteractions(); }
void apjref_calculate_com(void* h) { E(h).calculate_COM(); }
void apjref_save_old(void* h) { E(h).saveOldPositions(); }
void apjref_step(void* h) { E(h).calculate_next_positions(); }
// Cell::over (the Ovito hue, jamming.cpp:653-656): run neighborInteractions with the filming condition
// (countdown <= film && t % nSkip == 0) switched on, read the hue, restore the clock. Forces and alignment sums
// accumulate as in any call of neighborInteractions: the caller works on a scratch engine.
void apjref_overlap_hue(void* h, int* over) {
    Engine& e = E(h);
    const long t0 = e.t, c0 = e.countdown;
    for (long i = 0; i < e.N; i++) e.cell[i].over = 240;          // print_video leaves 240 behind (:866)
    e.t = 0; e.countdown = 1;
    e.neighborInteractions();
    e.t = t0; e.countdown = c0;
    for (long i = 0; i < e.N; i++) over[i] = e.cell[i].over;
}
void apjref_steps(void* h, long n) { for (long k = 0; k < n; k++) E(h).calculate_next_positions(); }
void apjref_relax(void* h) { E(h).relax(); }
double apjref_delta_norm(void* h, double d) { return E(h).delta_norm(d); }
// The three lines between relax() and the main loop (jamming.cpp:191-203).
void apjref_mark_origin(void* h) {
    Engine& e = E(h);
    for (int i = 0; i < e.N; i++)
        for (int k = 0; k < NDIM; k++) { e.cell[i].x_real[k] = e.cell[i].x[k]; e.cell[i].x0[k] = e.cell[i].x[k]; }
    e.calculate_COM();
    e.COM0 = e.COM;
    e.saveOldPositions();
}

// Verlet lists as stored by the reference (half lists): returns total entries;
// offsets has N+1 slots; idx may be NULL to query the size.
long apjref_get_verlet(void* h, long* offsets, int* idx, long cap) {
    Engine& e = E(h);
    long tot = 0;
    for (int i = 0; i < e.N; i++) {
        if (offsets) offsets[i] = tot;
        for (size_t k = 0; k < e.cell[i].VerletList.size(); k++) {
            if (idx && tot < cap) idx[tot] = e.cell[i].VerletList[k];
            tot++;
        }
    }
    if (offsets) offsets[e.N] = tot;
    return tot;
}
// Box occupancy lists (Box::CellList) flattened: offsets has nbox+1 slots.
long apjref_get_cell_lists(void* h, long* offsets, int* idx, long cap) {
    Engine& e = E(h);
    long tot = 0;
    for (int p = 0; p < e.nbox; p++) {
        if (offsets) offsets[p] = tot;
        for (size_t k = 0; k < e.grid[p].CellList.size(); k++) {
            if (idx && tot < cap) idx[tot] = e.grid[p].CellList[k];
            tot++;
        }
    }
    if (offsets) offsets[e.nbox] = tot;
    return tot;
}
void apjref_get_box_neighbors(void* h, int* out9) {  // nbox x 9
    Engine& e = E(h);
    for (int p = 0; p < e.nbox; p++) for (int m = 0; m < 9; m++) out9[p * 9 + m] = e.grid[p].neighbors[m];
}
long apjref_num_box_pairs(void* h) { return (long)E(h).boxPairs.size(); }

// On-engine observables (jamming.cpp:776-823).
double apjref_order(void* h) { return E(h).calculateOrderParameter(); }
void apjref_orientation(void* h, double* out2) { vector<double> o = E(h).calculateSystemOrientation(); out2[0] = o[0]; out2[1] = o[1]; }
double apjref_msd(void* h) { return E(h).MSD(); }

// Fluctuations / Correlations / Print (classes/*.h), driven exactly like Engine::start does.
int apjref_attach_observers(void* h, const char* location) {
    RefBox* b = (RefBox*)h; Engine& e = *b->eng;
    e.location = location;  // public member; replaces the author's hard-coded home directory
    b->printer = new Print(e.location, e.fullRun, e.run, e.N, remote);
    b->fluct = new Fluctuations(e.L, e.totalSteps, e.fluct_int, e.dens);
    b->corr = new Correlations(e.L, e.dens, e.cutoff, e.tCorrelation, e.N, e.CFself);
    return 0;
}
double apjref_fluct_overlap(void* h, double r, double R, double d) { return ((RefBox*)h)->fluct->overlap(r, R, d); }
// out = {current_radius, current_value, counter, time_interval, rad_interval}
void apjref_fluct_measure(void* h, double* out) {
    RefBox* b = (RefBox*)h;
    b->fluct->measureFluctuations(b->eng->cell, b->eng->COM, *b->printer);
    out[0] = b->fluct->current_radius; out[1] = b->fluct->current_value; out[2] = b->fluct->counter;
    out[3] = b->fluct->time_interval; out[4] = b->fluct->rad_interval;
}
void apjref_density_distribution(void* h, double* out50) {
    RefBox* b = (RefBox*)h;
    b->fluct->density_distribution(b->eng->cell, b->eng->grid);
    for (int k = 0; k < 50; k++) out50[k] = b->fluct->distribution[k];
}
void apjref_corr_dims(void* h, int* out) {  // {nc, np, noBins, correlation_time}
    Correlations& c = *((RefBox*)h)->corr;
    out[0] = c.nc; out[1] = c.np; out[2] = c.noBins; out[3] = c.correlation_time;
}
void apjref_spatial_correlations(void* h, double* vel_nc, double* ori_nc, double* pair_np) {
    RefBox* b = (RefBox*)h; Correlations& c = *b->corr;
    c.spatialCorrelations(b->eng->boxPairs, b->eng->grid, b->eng->cell);
    for (int k = 0; k < c.nc; k++) { vel_nc[k] = c.velocityCorrelation[k]; ori_nc[k] = c.orientationCorrelation[k]; }
    for (int k = 0; k < c.np; k++) pair_np[k] = c.pairCorrelationValues[k];
}
void apjref_vel_dist(void* h, double* out100) {
    RefBox* b = (RefBox*)h; Correlations& c = *b->corr;
    c.velDist(b->eng->cell);
    for (int k = 0; k < c.noBins; k++) out100[k] = c.velocityDistributionValues[k];
}
void apjref_autocorrelation(void* h, int t, double* out) {
    RefBox* b = (RefBox*)h; Correlations& c = *b->corr;
    vector<double> o = b->eng->calculateSystemOrientation();
    if (t == 0) c.orientation0 = o;
    c.autocorrelation(t, o);
    out[0] = c.autocorrelationValues[t];
}

// Whole reference driver (jamming.cpp:173-283) writing its .dat tree under `location`
// (location must contain local_output/). Returns wall seconds of Engine::start().
double apjref_run_start(void* h, const char* location) {
    Engine& e = E(h);
    e.location = location;
    high_resolution_clock::time_point a = high_resolution_clock::now();
    e.start();
    high_resolution_clock::time_point b = high_resolution_clock::now();
    return duration_cast<duration<double>>(b - a).count();
}
// Timed hot-path loop for the CPU baseline: n calls of calculate_next_positions().
double apjref_time_steps(void* h, long n) {
    Engine& e = E(h);
    high_resolution_clock::time_point a = high_resolution_clock::now();
    for (long k = 0; k < n; k++) e.calculate_next_positions();
    high_resolution_clock::time_point b = high_resolution_clock::now();
    return duration_cast<duration<double>>(b - a).count();
}

}  // extern "C"
