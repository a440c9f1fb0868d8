// jam driver of the B200 build: same command line, same `Engine` surface and same output files as
// the reference's code/jam/jamming.cpp (reference :43-117 Engine, :173-283 start(), :889-929 main),
// with the hot path -- Engine::calculate_next_positions() and the observables -- delegated to the
// CUDA library through the C ABI of include/apj_b200.h. No physics runs on the host: what stays
// here is what the reference also does once, outside the hot loop (initial lattice, grid
// geometry, the cadence of measurements, file output).
//
//   jam <fullRunID> <runID> <N> <steps> <lambda_s> <lambda_n> <rho>
//   jam --sweep <input.txt> [replicas per batch]
//
// The second form replaces the reference's PBS job array (code/jam/create-arrays.sh, jamming.sh: one
// single-core process per line of input.txt): consecutive lines with the same N and step count are
// run as replicas of ONE device handle (n_systems > 1, BASELINE config 5), each with its own Engine,
// observers and output tree local_output/<fullRunID>/<runID>/, stepped in lockstep (classes/Batch.h).
//
// Environment (additions; the reference hard-codes these at compile time, jamming.cpp:28,36,149):
//   APJ_OUTPUT_ROOT  root that holds local_output/ (default "./")
//   APJ_SEED         seed of the initial condition and of the Philox noise (default: wall clock)
//   APJ_DEVICE       CUDA device ordinal (default 0)
#define NDIM 2
#define PI 3.14159265
#define PI2 6.28318531
#define sqrt2 1.41421356

#include <vector>
#include <cmath>
#include <ctime>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <random>
#include <string>
#include <fstream>
#include <sstream>

using namespace std;
using namespace std::chrono;

#include "../classes/Cell.h"
#include "../classes/Box.h"
#include "../classes/Print.h"
#include "../classes/Batch.h"
#include "../classes/Fluctuations.h"
#include "../classes/Correlations.h"
#include "../../../include/apj_b200.h"

const bool remote = 0;      // 0: local settings (timeAvg 10, cutoff 20), 1: cluster settings (reference :28, :147-162)
const bool makevid = getenv("APJ_MAKEVID") && atoi(getenv("APJ_MAKEVID"));   // Ovito frames for the last `film` steps (reference :32: const 0; APJ_MAKEVID=1 films)

// Initial-condition randomness (reference :36-41 uses Boost mt19937 seeded from the clock; Boost is
// not a dependency here). Radii / lattice jitter ~ N(0,1), angles ~ U[-PI,PI) on the 2^-32 lattice.
static uint64_t g_seed = 0;
static std::mt19937 gen;
static std::normal_distribution<double> normdist(0, 1);
static double randnorm() { return normdist(gen); }
static double randuni() { return (double)gen()/4294967296.0*(PI - (-PI)) + (-PI); }

struct Engine
{
    Engine(string, string, long int, long int, double, double, double);
    ~Engine();

    string location;

    int N;
    double CFself;
    double CTnoise;
    double dens;
    string run;
    string fullRun;

    const double dt = 0.1;
    const int nSkip = 100;
    const int film  = nSkip*100;
    const int fluct_int = 10;

    long int totalSteps;
    long int countdown;
    long int t;
    long int resetCounter;

    int timeAvg;
    int tCorrelation;
    int cutoff;

    void start();
    void topology();
    void initCells();
    void assignCellsToGrid();
    void buildVerletLists();
    void relax();
    bool newSkinList();
    void calculate_next_positions();
    void neighborInteractions();
    void calculate_COM();
    void saveOldPositions();
    double calculateOrderParameter();
    vector<double> calculateSystemOrientation();
    void print_video(Print&);
    bool device_init = false;       // initCells ran on the device: `cell` is filled on demand only
    void film_hue();
    bool pulled_hue = false;
    double delta_norm(double);
    double MSD();

    vector<double> COM;
    vector<double> COM0;
    vector<double> COM_old;
    double orderAvg, order2Avg, order4Avg;
    double binder, variance;

    vector<Cell> cell;
    vector<Box> grid;
    vector<vector<int>> boxPairs;       // left empty: the device derives the cell ranges within `cutoff` itself

    double L;
    double Lover2;
    double lp;
    int b;
    int nbox;
    int nboxnb;

    const double rn = 2.8;
    const double rs = 1.5*rn;
    const double rn2 = rn*rn;
    const double rs2 = rs*rs;

    // ---- device side (additions) ----
    ApjBatch* batch = nullptr;          // the device handle, shared with the other replicas of a sweep (classes/Batch.h)
    bool own_batch = false;
    int sys = 0;                        // this run's system index in the batch
    apj_engine* dev = nullptr;          // == batch->dev
    void attach_device();               // single run: apj_create + upload of `cell`
    static ApjBatch* attach_batch(vector<Engine*>& runs);   // sweep: one handle for all of them
    static void relax_all(vector<Engine*>& runs);
    void flush();                       // run the steps queued by calculate_next_positions()
    void pull_cells();                  // refresh the host mirror `cell` from HBM
    void pull_cell_lists();             // refresh grid[].CellList (Box::CellList, ascending particle index)
    void pull_verlet_lists();           // refresh cell[].VerletList (half lists, partners j > i)

    // start() in pieces, so that a sweep can run many Engines in lockstep
    void setup();                       // initCells + topology
    void open_outputs();                // Print, Fluctuations, Correlations
    void bind_observers();
    void tick();                        // one iteration of the measurement loop (start() :207-268)
    void finish();                      // averages, correlation / density output, summary
    Print* printer = nullptr;
    Fluctuations* fluct = nullptr;
    Correlations* corr = nullptr;
    int corrCounter = 0;
    high_resolution_clock::time_point t_begin;

private:
    long int pending = 0;               // steps requested but not yet launched
    long long requested = 0;            // measurement steps requested so far (the batch runs max over its Engines)
    long long seen_version = -1;
    bool lists_fresh = false;
    long int rebuilds_seen = 0;
    void check(int rc, const char* what);
};

void Engine::check(int rc, const char* what)
{
    if (rc == APJ_OK) return;
    cout << what << " failed (" << rc << "): " << apj_last_error(dev) << endl;
    exit(120);
}

Engine::Engine(string dir, string ID, long int n, long int steps, double l_s, double l_n, double rho)
{
    fullRun = dir;
    run = ID;
    N = n;
    totalSteps = steps;
    countdown = steps + 1;              // the loop in start() runs steps+1 times (reference :126,207; Q16)
    CFself = l_s;
    CTnoise = l_n;
    dens = rho;
    t = 0;
    resetCounter = 0;
    orderAvg = order2Avg = order4Avg = 0.0;
    binder = variance = 0.0;
    COM.assign(NDIM, 0.0);
    COM0.assign(NDIM, 0.0);
    COM_old.assign(NDIM, 0.0);
    nboxnb = 9;
    L = Lover2 = lp = 0.0;
    b = nbox = 0;

    const char* root = getenv("APJ_OUTPUT_ROOT");
    location = root ? root : "./";
    if (!location.empty() && location[location.size() - 1] != '/') location += "/";
    if (remote == 0) { timeAvg = 10; tCorrelation = 10; cutoff = 20; }
    else { timeAvg = 100; tCorrelation = 100; cutoff = 140; }
}

Engine::~Engine()
{
    delete corr; delete fluct; delete printer;
    if (own_batch) delete batch;
}

// ---------------------------------------------------------------------------------------------
// Host-side setup (done once, as in the reference)

// Polydisperse radii 1 + N(0,1)/10, box from the packing fraction with the truncated PI, jittered
// offset-row lattice, random polarity (reference initCells :285-354, 2D branch).
void Engine::initCells()
{
    // APJ_DEVICE_INIT=1 (or N >= 4M): the lattice is drawn on the device (apj_init_lattice, csrc/apj_setup.cu) and
    // vector<Cell> stays empty until somebody asks for it (pull_cells): a 16M-particle run never builds ~5 GB of Cell
    // records on the host. Only the box length is needed here (topology() and apj_create depend on it).
    const char* di = getenv("APJ_DEVICE_INIT");
    device_init = di ? atoi(di) != 0 : N >= 4194304;
    if (device_init) {
        const char* d = getenv("APJ_DEVICE");
        if (apj_lattice_box_length(N, 1, g_seed, &dens, d ? atoi(d) : 0, &L) != APJ_OK) {
            cout << "apj_lattice_box_length failed: " << apj_last_error(NULL) << endl;
            exit(120);
        }
        Lover2 = L/2.0;
        cell.clear();
        return;
    }
    cell.assign(N, Cell());
    double area = 0;
    for (int i = 0; i < N; i++) {
        Cell& c = cell[i];
        c.index = i;
        c.R = 1. + randnorm()/10;
        c.Rinv = 1.0/c.R;
        area += c.R*c.R;
    }
    L = sqrt(PI*area/dens);
    Lover2 = L/2.0;

    const int rootN = sqrt(N);
    const double spacing = L/rootN;
    for (int i = 0; i < N; i++) {
        Cell& c = cell[i];
        c.L = L; c.Lover2 = Lover2; c.dt = dt;
        const int row = i/rootN;
        c.x[0] = -Lover2 + spacing*(i % rootN) + randnorm()/10.;
        c.x[1] = -Lover2 + spacing*row + randnorm()/10.;
        if (row % 2 == 0) c.x[0] += 1.0;
        c.theta = PI/2.0;
        c.phi = randuni();
        c.cosp = cos(c.phi);
        c.sinp = sin(c.phi);
        c.periodicAngles();
        c.PBC();
    }
}

// Grid of b x b boxes, b = floor(L / 2 rn), with centres and the 3x3 periodic neighbour table in
// the reference's numbering p = i + j*b (reference topology :356-410). The device derives the same
// b and lp from L (apj_create); the O(nbox^2) boxPairs table (:460-479) is not built.
void Engine::topology()
{
    lp = 2*rn;
    b = static_cast<int>(floor(L/lp));
    nbox = b*b;
    lp = L/floor(L/lp);
    grid.assign(nbox, Box());
    for (int j = 0; j < b; j++) {
        for (int i = 0; i < b; i++) {
            Box& g = grid[i + j*b];
            g.serial_index = i + j*b;
            g.vector_index[0] = i; g.vector_index[1] = j;
            g.min[0] = -Lover2 + i*L/b;       g.min[1] = -Lover2 + j*L/b;
            g.max[0] = -Lover2 + (i + 1)*L/b; g.max[1] = -Lover2 + (j + 1)*L/b;
            for (int m = 0; m < NDIM; m++) g.center[m] = (g.min[m] + g.max[m])/2.;
            for (int dj = -1; dj <= 1; dj++)
                for (int di = -1; di <= 1; di++) {
                    const int ii = (i + di + b) % b, jj = (j + dj + b) % b;
                    g.neighbors[(di + 1) + (dj + 1)*3] = ii + jj*b;
                }
        }
    }
}

void Engine::attach_device()
{
    vector<Engine*> one(1, this);
    attach_batch(one);
    own_batch = true;
}

// One handle for all runs of a batch: system s = runs[s]. Per-system box length and activity; the
// state of every run is uploaded in one call (system-major SoA).
ApjBatch* Engine::attach_batch(vector<Engine*>& runs)
{
    const int S = (int)runs.size(), N = runs[0]->N;
    ApjBatch* b = new ApjBatch(S);
    apj_config cfg = {};
    cfg.n = N;
    cfg.n_systems = S;
    const char* d = getenv("APJ_DEVICE");
    cfg.device = d ? atoi(d) : 0;
    cfg.dt = runs[0]->dt; cfg.rn = runs[0]->rn; cfg.rs_factor = runs[0]->rs/runs[0]->rn;
    cfg.seed = g_seed;
    const char* mn = getenv("APJ_MAX_NEIGHBORS");      // initial Verlet-list capacity; the library grows it on demand
    cfg.max_neighbors = mn ? atoi(mn) : 0;
    vector<double> Ls(S), CF(S), CT(S);
    for (int s = 0; s < S; s++) { Ls[s] = runs[s]->L; CF[s] = runs[s]->CFself; CT[s] = runs[s]->CTnoise; }
    int rc = apj_create(&cfg, Ls.data(), &b->dev);
    if (rc != APJ_OK) { cout << "apj_create failed (" << rc << "): " << apj_last_error(NULL) << endl; exit(120); }
    b->check(apj_set_activity(b->dev, CF.data(), CT.data()), "apj_set_activity");

    if (S == 1 && runs[0]->device_init) {      // Engine::initCells on the device: same seed as apj_lattice_box_length saw
        b->check(apj_init_lattice(b->dev, g_seed), "apj_init_lattice");
        double zero[2] = {0.0, 0.0};
        b->check(apj_set_com(b->dev, 0, zero, zero, zero), "apj_set_com");
        b->check(apj_skip_self_term_once(b->dev, 1), "apj_skip_self_term_once");
        b->touch();
        Engine* e = runs[0];
        e->batch = b; e->dev = b->dev; e->sys = 0; e->lists_fresh = true;
        if (e->fluct) e->bind_observers();
        return b;
    }
    for (int s = 0; s < S; s++)
        if (runs[s]->device_init) { cout << "APJ_DEVICE_INIT: batched sweeps initialise on the host" << endl; exit(120); }
    const size_t T = (size_t)S*N;
    vector<double> x(T), y(T), xr(T), yr(T), x0(T), y0(T), xo(T), yo(T), R(T), phi(T), cp(T), sp(T);
    vector<int32_t> box(T);
    for (int s = 0; s < S; s++)
        for (int i = 0; i < N; i++) {
            const Cell& c = runs[s]->cell[i];
            const size_t k = (size_t)s*N + i;
            x[k] = c.x[0]; y[k] = c.x[1]; xr[k] = c.x_real[0]; yr[k] = c.x_real[1];
            x0[k] = c.x0[0]; y0[k] = c.x0[1]; xo[k] = c.x_old[0]; yo[k] = c.x_old[1];
            R[k] = c.R; phi[k] = c.phi; cp[k] = c.cosp; sp[k] = c.sinp; box[k] = c.box;
        }
    apj_state st = {};
    st.x = x.data(); st.y = y.data(); st.x_real = xr.data(); st.y_real = yr.data(); st.x0 = x0.data(); st.y0 = y0.data();
    st.x_old = xo.data(); st.y_old = yo.data(); st.R = R.data(); st.phi = phi.data(); st.cosp = cp.data(); st.sinp = sp.data();
    st.box = box.data();
    b->check(apj_upload_state(b->dev, &st), "apj_upload_state");          // bins and builds the lists too
    // a fresh reference Engine starts with COM = COM_old = 0 and x_new = 0 (reference :139-141, Cell.h:75)
    double zero[2] = {0.0, 0.0};
    for (int s = 0; s < S; s++) b->check(apj_set_com(b->dev, s, zero, zero, zero), "apj_set_com");
    b->check(apj_skip_self_term_once(b->dev, 1), "apj_skip_self_term_once");
    b->touch();
    for (int s = 0; s < S; s++) {
        Engine* e = runs[s];
        e->batch = b; e->dev = b->dev; e->sys = s; e->lists_fresh = true;
        if (e->fluct) e->bind_observers();
    }
    return b;
}

// ---------------------------------------------------------------------------------------------
// Hot path: thin calls into the ABI

// One Euler step. Steps are queued and launched together the next time somebody looks at the
// state (flush), so the device runs uninterrupted between two measurements.
void Engine::calculate_next_positions()
{
    pending++;
    lists_fresh = false;
}

void Engine::flush()
{
    if (pending > 0) { requested += pending; pending = 0; }
    batch->step_to(requested);          // no-op when another replica of the batch already advanced the device
    if (seen_version != batch->version) {
        double com[2], com0[2], como[2];
        long long reset = 0;
        batch->com(sys, com, com0, como, &reset);
        rebuilds_seen = reset - resetCounter;
        resetCounter = reset;
        for (int k = 0; k < NDIM; k++) { COM[k] = com[k]; COM0[k] = com0[k]; COM_old[k] = como[k]; }
        seen_version = batch->version;
    }
}

// assignCellsToGrid + buildVerletLists are one on-device chain (counting sort + list build);
// the pair keeps the reference's two-call protocol (start() :184-185, :245-246). In a sweep the
// chain runs once for all replicas.
void Engine::assignCellsToGrid()
{
    flush();
    long long v;
    batch->force_rebuild_once(v);
    lists_fresh = true;
}

void Engine::buildVerletLists()
{
    flush();
    long long v;
    if (!lists_fresh) batch->force_rebuild_once(v);
    lists_fresh = true;
}

// The skin test, the pair sweep and the COM sum are fused into the step kernel; these three keep
// the reference's names for callers that only observe. newSkinList reports whether the steps run
// by the last flush rebuilt the lists.
bool Engine::newSkinList() { flush(); return rebuilds_seen > 0; }
void Engine::neighborInteractions() { /* part of apj_step(): the fused pair sweep (csrc/apj_step.cu) */ }
void Engine::calculate_COM() { flush(); }
void Engine::saveOldPositions() { /* done on the device at every rebuild and by apj_mark_origin() */ }

// Passive relaxation, then a linear ramp of the self-propulsion (reference relax :482-525).
void Engine::relax()
{
    vector<Engine*> one(1, this);
    relax_all(one);
}

void Engine::relax_all(vector<Engine*>& runs)
{
    const long int trelax = remote ? (long int)(1000.0/runs[0]->dt) : 2000;
    const long int tthermalize = remote ? 1000000 : 2000;
    ApjBatch* b = runs[0]->batch;
    const int S = (int)runs.size();
    vector<double> zero(S, 0.0), CF(S), CT(S);
    for (int s = 0; s < S; s++) { runs[s]->flush(); CF[s] = runs[s]->CFself; CT[s] = runs[s]->CTnoise; }
    b->check(apj_set_activity(b->dev, zero.data(), CT.data()), "apj_set_activity");
    b->check(apj_step(b->dev, trelax), "apj_step");
    b->check(apj_set_activity(b->dev, CF.data(), CT.data()), "apj_set_activity");
    b->check(apj_set_ramp(b->dev, tthermalize), "apj_set_ramp");
    b->check(apj_step(b->dev, tthermalize), "apj_step");
    b->check(apj_set_ramp(b->dev, 0), "apj_set_ramp");
    for (int s = 0; s < S; s++) {
        runs[s]->resetCounter = 0;
        b->check(apj_set_reset_counter(b->dev, s, 0), "apj_set_reset_counter");
    }
    b->touch();
}

double Engine::calculateOrderParameter()
{
    flush();
    return batch->order(sys);           // apj_order_orientation
}

vector<double> Engine::calculateSystemOrientation()
{
    flush();
    vector<double> o(NDIM, 0.0);
    batch->orientation(sys, o.data());
    return o;
}

double Engine::MSD()
{
    flush();
    return batch->msd(sys);             // apj_msd
}

double Engine::delta_norm(double delta)
{
    const double step = delta < -Lover2 ? L : -L;
    while (delta < -Lover2 || delta >= Lover2) delta += step;
    return delta;
}

// ---------------------------------------------------------------------------------------------
// Mirrors, on request

void Engine::pull_cells()
{
    flush();
    const size_t T = (size_t)batch->nsys*N, o = (size_t)sys*N;
    vector<double> f[14];
    for (auto& v : f) v.resize(T);
    vector<int32_t> box(T);
    apj_state s = {};
    s.x = f[0].data(); s.y = f[1].data(); s.x_real = f[2].data(); s.y_real = f[3].data(); s.x0 = f[4].data(); s.y0 = f[5].data();
    s.x_old = f[6].data(); s.y_old = f[7].data(); s.R = f[8].data(); s.phi = f[9].data(); s.cosp = f[10].data(); s.sinp = f[11].data();
    s.vx = f[12].data(); s.vy = f[13].data(); s.box = box.data();
    check(apj_download_state(dev, &s), "apj_download_state");
    if ((long)cell.size() != N) {                   // device-side initCells: the records exist from now on
        cell.assign(N, Cell());
        for (int i = 0; i < N; i++) { Cell& c = cell[i]; c.index = i; c.L = L; c.Lover2 = Lover2; c.dt = dt; c.theta = PI/2.0; c.over = 240; }
    }
    for (int i = 0; i < N; i++) {
        Cell& c = cell[i];
        const size_t k = o + i;
        c.R = f[8][k]; c.Rinv = 1.0/c.R;
        c.x[0] = f[0][k]; c.x[1] = f[1][k]; c.x_real[0] = f[2][k]; c.x_real[1] = f[3][k];
        c.x0[0] = f[4][k]; c.x0[1] = f[5][k]; c.x_old[0] = f[6][k]; c.x_old[1] = f[7][k];
        c.phi = f[9][k]; c.cosp = f[10][k]; c.sinp = f[11][k]; c.vx = f[12][k]; c.vy = f[13][k];
        c.x_new = c.cosp; c.y_new = c.sinp; c.box = box[k];
    }
}

void Engine::pull_cell_lists()
{
    flush();
    vector<int64_t> off(nbox + 1);
    vector<int32_t> idx(N);
    check(apj_get_cell_lists(dev, sys, off.data(), idx.data()), "apj_get_cell_lists");
    for (int p = 0; p < nbox; p++) grid[p].CellList.assign(idx.begin() + off[p], idx.begin() + off[p + 1]);
}

void Engine::pull_verlet_lists()
{
    flush();
    vector<int64_t> off(N + 1);
    int64_t total = 0;
    check(apj_get_pair_list(dev, sys, off.data(), NULL, 0, &total), "apj_get_pair_list");
    vector<int32_t> idx(total > 0 ? total : 1);
    check(apj_get_pair_list(dev, sys, off.data(), idx.data(), total, &total), "apj_get_pair_list");
    for (int i = 0; i < N; i++) cell[i].VerletList.assign(idx.begin() + off[i], idx.begin() + off[i + 1]);
}

void Engine::print_video(Print& printer)
{
    pull_cells();
    int first = 0;
    for (int i = 0; i < N; i++) {
        vector<double> velocity(3, 0.0);
        velocity[0] = cell[i].vx;
        velocity[1] = cell[i].vy;
        printer.print_Ovito(first, N, i, cell[i].R, cell[i].over, cell[i].x, velocity);   // over: film_hue() of this step
        cell[i].over = 240;             // :866
        first = 1;
    }
}

// Cell::over of a filmed step (reference :653-656): the hue is accumulated by the neighborInteractions of THE step whose
// frame is printed, from the positions that step starts from. All queued steps but the last are run, the device
// evaluates the hue for that state (apj_overlap_hue: the reference's integer, update order included), then the last
// step runs. (If that very step rebuilds the lists, the reference visits the pairs in the new lists' order; the hue
// can then differ by one unit where two truncations fall differently.)
void Engine::film_hue()
{
    if (pending < 1) return;
    pending--;
    flush();
    pending = 1;
    if ((long)cell.size() != N) {                   // device-side initCells: materialise the records (no stepping: pending is parked)
        const long int keep = pending;
        pending = 0; pull_cells(); pending = keep;
    }
    vector<int32_t> hue((size_t)batch->nsys*N);
    check(apj_overlap_hue(dev, hue.data()), "apj_overlap_hue");
    for (int i = 0; i < N; i++) cell[i].over = hue[(size_t)sys*N + i];
    pulled_hue = true;
}

// ---------------------------------------------------------------------------------------------
// Run loop: the reference's cadence (start() :173-283) -- fluctuations every 10 steps (twice on
// multiples of 100, Q12), order / orientation / COM / MSD every 100, correlations timeAvg times
// per run, autocorrelation for tCorrelation steps after each of those.
void Engine::setup()
{
    t_begin = high_resolution_clock::now();
    initCells();
    topology();
}

void Engine::open_outputs()
{
    printer = new Print(location, fullRun, run, N, remote);
    fluct = new Fluctuations(L, totalSteps, fluct_int, dens);
    corr = new Correlations(L, dens, cutoff, tCorrelation, N, CFself);
    if (batch) bind_observers();
}

void Engine::bind_observers()
{
    fluct->bind(batch, sys);
    corr->bind(batch, sys);
}

void Engine::tick()
{
    const long int corr_every = totalSteps/timeAvg;

    calculate_next_positions();
    if (makevid && countdown < film && t % nSkip == 0) film_hue();   // the frame print_video writes below

    if (t % fluct_int == 0) { flush(); fluct->measureFluctuations(cell, COM, *printer); }

    if (t % nSkip == 0) {
        flush();
        fluct->measureFluctuations(cell, COM, *printer);
        const double order = calculateOrderParameter();
        vector<double> orientation = calculateSystemOrientation();
        const double order2 = order*order;
        orderAvg += order;
        order2Avg += order2;
        order4Avg += order2*order2;
        printer->print_COM(t, COM);
        printer->print_order(t, order);
        printer->print_orientation(t, orientation);
        printer->print_MSD(t, MSD());
        if (countdown < film && makevid) print_video(*printer);
    }

    if (corr_every > 0 && t % corr_every == 0 && t != 0) {
        assignCellsToGrid();
        buildVerletLists();
        corr->orientation0 = calculateSystemOrientation();
        corr->spatialCorrelations(boxPairs, grid, cell);
        corr->velDist(cell);
        fluct->density_distribution(cell, grid);
        corrCounter = 0;
    }

    if (corrCounter < tCorrelation) {
        vector<double> orient = calculateSystemOrientation();
        corr->autocorrelation(corrCounter, orient);
        corrCounter++;
    }

    t++;
    countdown--;
}

void Engine::finish()
{
    flush();

    orderAvg  /= (double)totalSteps/(double)nSkip;
    order2Avg /= (double)totalSteps/(double)nSkip;
    order4Avg /= (double)totalSteps/(double)nSkip;
    binder = 1.0 - order4Avg/(3.0*order2Avg*order2Avg);
    variance = order2Avg - orderAvg*orderAvg;

    corr->printCorrelations(timeAvg, *printer);
    fluct->print_density_distribution(timeAvg, *printer);

    high_resolution_clock::time_point t_end = high_resolution_clock::now();
    auto duration = duration_cast<seconds>(t_end - t_begin).count();
    printer->print_summary(run, N, L, t, 1./dt, CFself, CTnoise, dens, duration, resetCounter, binder, orderAvg, variance);
}

// APJ_TIMING=1: wall-clock seconds of every phase of start() on stderr (where a short run spends its time).
static void lap(const char* what)
{
    static const bool on = getenv("APJ_TIMING") && atoi(getenv("APJ_TIMING"));
    static high_resolution_clock::time_point last = high_resolution_clock::now();
    if (!on) return;
    const high_resolution_clock::time_point now = high_resolution_clock::now();
    fprintf(stderr, "[jam timing] %-28s %.3f s\n", what, duration_cast<duration<double>>(now - last).count());
    last = now;
}

void Engine::start()
{
    lap("process start");
    setup();
    open_outputs();
    lap("initCells + output tree");
    attach_device();
    lap("apj_create + upload");

    assignCellsToGrid();
    buildVerletLists();

    relax();
    lap("relax");

    check(apj_mark_origin(dev), "apj_mark_origin");       // x_real = x0 = x, COM0 = COM, saveOldPositions (:191-203)
    batch->touch();
    pending = 0;

    while (countdown != 0) tick();
    lap("measurement loop");
    finish();
    lap("finish");
}

// ---------------------------------------------------------------------------------------------
// Sweep: the reference's job array (create-arrays.sh writes "ID runK N steps lambda_s lambda_n rho"
// lines into input.txt, jamming.sh starts one process per line) as batched replicas on one GPU.
struct RunSpec { string full, run; long int n, steps; double l_s, l_n, rho; int line; };

static int run_sweep(const char* path, int per_batch)
{
    ifstream in(path);
    if (!in) { cout << "cannot open " << path << endl; return 2; }
    vector<RunSpec> specs;
    string ln;
    int lineno = 0;
    while (getline(in, ln)) {
        lineno++;
        istringstream is(ln);
        RunSpec r;
        if (!(is >> r.full >> r.run >> r.n >> r.steps >> r.l_s >> r.l_n >> r.rho)) {
            if (ln.find_first_not_of(" \t\r") == string::npos) continue;     // blank line
            cout << path << ":" << lineno << ": expected <fullRunID> <runID> <N> <steps> <lambda_s> <lambda_n> <rho>" << endl;
            return 2;
        }
        r.line = lineno;
        specs.push_back(r);
    }
    if (per_batch < 1) per_batch = 64;
    int skipped = 0;
    size_t first = 0;
    while (first < specs.size()) {
        size_t last = first + 1;                          // consecutive lines of one shape form a batch
        while (last < specs.size() && last - first < (size_t)per_batch && specs[last].n == specs[first].n && specs[last].steps == specs[first].steps) last++;
        vector<Engine*> runs;
        for (size_t k = first; k < last; k++) {
            const RunSpec& r = specs[k];
            gen.seed((uint32_t)(g_seed + 7919u*(uint64_t)r.line));       // every line its own initial condition
            normdist.reset();
            // a bad line is reported and skipped on its own: it must not take the other replicas of its batch with it
            if (r.n < 1 || r.steps < 1 || !(r.rho > 0)) {
                cout << path << ":" << r.line << ": skipped (N, steps and rho must be positive)" << endl;
                skipped++;
                continue;
            }
            Engine* e = new Engine(r.full, r.run, r.n, r.steps, r.l_s, r.l_n, r.rho);
            e->setup();
            if (e->b < 3) {     // b = floor(L / 2 rn) < 3 aliases neighbour cells (reference :359): the device refuses such a box
                cout << path << ":" << r.line << ": skipped (box too small: L = " << e->L << " gives b = " << e->b << " < 3 cells per side)" << endl;
                delete e;
                skipped++;
                continue;
            }
            e->open_outputs();
            runs.push_back(e);
        }
        if (runs.empty()) { first = last; continue; }
        ApjBatch* b = Engine::attach_batch(runs);
        for (Engine* e : runs) { e->assignCellsToGrid(); e->buildVerletLists(); }
        Engine::relax_all(runs);
        b->check(apj_mark_origin(b->dev), "apj_mark_origin");
        b->touch();
        while (runs[0]->countdown != 0)
            for (Engine* e : runs) e->tick();                 // lockstep: the first replica advances the device, the others read
        for (Engine* e : runs) e->finish();
        cout << "batch of " << runs.size() << " runs (N = " << specs[first].n << ", " << specs[first].steps << " steps): "
             << specs[first].run << " .. " << specs[last - 1].run << endl;
        for (Engine* e : runs) delete e;
        delete b;
        first = last;
    }
    if (skipped) cout << skipped << " line(s) of " << path << " skipped" << endl;
    return skipped ? 3 : 0;
}

int main(int argc, char* argv[])
{
    const char* s = getenv("APJ_SEED");
    g_seed = s ? strtoull(s, NULL, 10)
               : (uint64_t)duration_cast<nanoseconds>(high_resolution_clock::now().time_since_epoch()).count();
    gen.seed((uint32_t)g_seed);

    if (argc >= 3 && string(argv[1]) == "--sweep") return run_sweep(argv[2], argc >= 4 ? atoi(argv[3]) : 64);

    if (argc != 8) {
        cout << "Incorrect number of arguments. Need: " << endl
             << "- full run ID" << endl
             << "- single run ID" << endl
             << "- number of cells" << endl
             << "- number of steps" << endl
             << "- \\lambda_s" << endl
             << "- \\lambda_n" << endl
             << "- \\rho" << endl
             << "Program exit status (1)" << endl;
        return 1;
    }

    Engine engine(argv[1], argv[2], atol(argv[3]), atol(argv[4]), atof(argv[5]), atof(argv[6]), atof(argv[7]));
    engine.start();
    return 0;
}
