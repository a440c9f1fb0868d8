// One device handle shared by the Engines of a batched phase-diagram sweep (BASELINE config 5;
// the reference runs one single-core process per (rho, lambda_n, lambda_s) point, code/jam/jamming.sh:2-15).
// The replicas of a batch live in ONE apj_engine with n_systems > 1, so every kernel launch steps all of
// them; each Engine / Fluctuations / Correlations object keeps its own bookkeeping (cadence counters,
// accumulators, output files) and asks the batch for ITS system's number. The first object to ask for a
// quantity after the device state changed triggers one batched ABI call for all systems; the others read
// the cached row. A single run is a batch of one.
#ifndef APJ_HOST_BATCH_H
#define APJ_HOST_BATCH_H

#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../../include/apj_b200.h"

struct ApjBatch
{
    apj_engine* dev = nullptr;
    int nsys = 1;
    long long version = 0;              // bumped whenever the device state changes (steps, rebuilds, origin)
    long long dev_steps = 0;            // Euler steps of the measurement phase executed so far

    // inputs that differ per system, registered by the observers
    std::vector<double*> radius_ptr;    // Fluctuations::current_radius of every system
    std::vector<double> dv;             // Correlations::dv of every system

    explicit ApjBatch(int n = 1) : nsys(n), radius_ptr(n, nullptr), dv(n, 1.0),
        v_order(-1), v_msd(-1), v_com(-1), v_fluct(-1), v_occ(-1), v_vel(-1), v_corr(-1), corr_cutoff(0), corr_nc(0), corr_np(0) {}
    ~ApjBatch() { if (dev) apj_destroy(dev); }

    void check(int rc, const char* what) const {
        if (rc == APJ_OK) return;
        fprintf(stdout, "%s failed (%d): %s\n", what, rc, apj_last_error(dev));
        exit(120);
    }
    void touch() { version++; }

    // advance the whole batch to `target` measurement steps (no-op if another Engine already did)
    void step_to(long long target) {
        if (target > dev_steps) { check(apj_step(dev, target - dev_steps), "apj_step"); dev_steps = target; touch(); }
    }
    void force_rebuild_once(long long& seen_version) {   // assignCellsToGrid + buildVerletLists, once per state
        if (v_rebuilt != version) { check(apj_force_rebuild(dev), "apj_force_rebuild"); touch(); v_rebuilt = version; }
        seen_version = version;
    }

    double order(int s) { refresh_order(); return c_order[s]; }
    void orientation(int s, double* o2) { refresh_order(); o2[0] = c_orient[2*s]; o2[1] = c_orient[2*s + 1]; }
    double msd(int s) {
        if (v_msd != version) { c_msd.resize(nsys); check(apj_msd(dev, c_msd.data()), "apj_msd"); v_msd = version; }
        return c_msd[s];
    }
    void com(int s, double* com, double* com0, double* com_old, long long* reset_counter) {
        if (v_com != version) {
            c_com.resize(6*nsys); c_reset.resize(nsys);
            for (int k = 0; k < nsys; k++) {
                check(apj_get_com(dev, k, &c_com[6*k], &c_com[6*k + 2], &c_com[6*k + 4]), "apj_get_com");
                int64_t c[8];
                check(apj_get_counters(dev, k, c), "apj_get_counters");
                c_reset[k] = c[1];
            }
            v_com = version;
        }
        for (int k = 0; k < 2; k++) { com[k] = c_com[6*s + k]; com0[k] = c_com[6*s + 2 + k]; com_old[k] = c_com[6*s + 4 + k]; }
        *reset_counter = c_reset[s];
    }
    // total disk area of system s inside the circle of `radius` around its COM (Fluctuations.h:62-76)
    double fluct_area(int s, double radius) {
        if (v_fluct != version || c_rad.empty() || c_rad[s] != radius) {
            c_rad.resize(nsys); c_area.resize(nsys);
            for (int k = 0; k < nsys; k++) c_rad[k] = radius_ptr[k] ? *radius_ptr[k] : radius;
            c_rad[s] = radius;
            check(apj_fluct_area(dev, c_rad.data(), c_area.data()), "apj_fluct_area");
            v_fluct = version;
        }
        return c_area[s];
    }
    const int64_t* occupancy_hist(int s) {
        if (v_occ != version) { c_occ.resize(50*(size_t)nsys); check(apj_occupancy_hist(dev, c_occ.data()), "apj_occupancy_hist"); v_occ = version; }
        return &c_occ[50*(size_t)s];
    }
    const int64_t* vel_hist(int s) {
        if (v_vel != version) { c_vel.resize(100*(size_t)nsys); check(apj_vel_hist(dev, dv.data(), c_vel.data()), "apj_vel_hist"); v_vel = version; }
        return &c_vel[100*(size_t)s];
    }
    // raw sums of Correlations::spatialCorrelations for system s: counts[nc], ori[nc], vel[nc], pair[np]
    void spatial_correlations(int s, double cutoff, int nc, int np, const double** counts, const double** ori, const double** vel, const double** pair) {
        if (v_corr != version || corr_cutoff != cutoff || corr_nc != nc || corr_np != np) {
            c_cnt.resize((size_t)nc*nsys); c_ori.resize((size_t)nc*nsys); c_velc.resize((size_t)nc*nsys); c_pair.resize((size_t)np*nsys);
            check(apj_spatial_correlations(dev, cutoff, c_cnt.data(), c_ori.data(), c_velc.data(), c_pair.data()), "apj_spatial_correlations");
            v_corr = version; corr_cutoff = cutoff; corr_nc = nc; corr_np = np;
        }
        *counts = &c_cnt[(size_t)nc*s]; *ori = &c_ori[(size_t)nc*s]; *vel = &c_velc[(size_t)nc*s]; *pair = &c_pair[(size_t)np*s];
    }

private:
    long long v_order, v_msd, v_com, v_fluct, v_occ, v_vel, v_corr, v_rebuilt = -1;
    double corr_cutoff; int corr_nc, corr_np;
    std::vector<double> c_order, c_orient, c_msd, c_com, c_rad, c_area, c_cnt, c_ori, c_velc, c_pair;
    std::vector<long long> c_reset;
    std::vector<int64_t> c_occ, c_vel;
    void refresh_order() {
        if (v_order != version) {
            c_order.resize(nsys); c_orient.resize(2*(size_t)nsys);
            check(apj_order_orientation(dev, c_order.data(), c_orient.data()), "apj_order_orientation");
            v_order = version;
        }
    }
};

#endif
