// On-GPU observables (reference: Engine::{calculateOrderParameter, calculateSystemOrientation,
// MSD} code/jam/jamming.cpp:776-823; Fluctuations.h:51-139; Correlations.h:71-187).
#pragma once
#include <vector>
#include "apj_device.cuh"

#define APJ_E_CUDA_OBS (-2)

struct ApjObsScratch {
    int blocks_per_sys = 0;
    double* d_part = nullptr;   // n_sys * blocks_per_sys * 2
    double* d_out = nullptr;    // n_sys * 2
    double* d_param = nullptr;  // n_sys
    unsigned long long* d_hist = nullptr;  // n_sys * 128
    double* d_corr = nullptr;   // lazily sized n_sys * (3*nc + np)
    size_t corr_cap = 0;
    double* d_ring = nullptr;   // queued observables: ring_cap slots of n_sys * 2 raw sums
    long long ring_cap = 0, next_ticket = 0;
};
#define APJ_OBS_RING 4096

int apj_obs_alloc(ApjObsScratch* o, const DevState& st, cudaStream_t stream, std::vector<void*>& allocs);
int apj_obs_com(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, double* com2);
int apj_obs_order(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, double* order, double* orient2);
int apj_obs_msd(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, double* msd);
int apj_obs_fluct(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, const double* radius, double* area);
int apj_obs_velhist(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, const double* dv, int64_t* hist100);
int apj_obs_occupancy(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, int64_t* hist50);
int apj_obs_enqueue_ring(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, int kind, const double* h_param, long long* ticket);
int apj_obs_fetch_ring(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long first, long long count, double* h_out);
int apj_obs_spatial(ApjObsScratch* o, const DevState& st, cudaStream_t s, long long* launches, const SysCtl* hctl, double cutoff,
                    double* counts, double* ori, double* vel, double* pair, std::vector<void*>& allocs,
                    long long n_ext = 0, const double* h_ext6 = nullptr, const int* h_ext_row = nullptr);
int apj_obs_export_edge(const DevState& st, cudaStream_t s, long long* launches, double width, long long cap, double* h_out6, long long* n_found);
