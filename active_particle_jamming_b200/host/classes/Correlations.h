// Static and time correlation functions with the reference's interface (reference
// code/classes/Correlations.h:7-38): velocity-direction and polarity correlations in shells of
// 2.0, g(r) in shells of 0.1, the speed histogram and the orientation autocorrelation.
//
// The pair sweep (O(N * density * pi * cutoff^2) distance tests, reference :85-152) and the speed
// histogram run on the GPU (apj_spatial_correlations, apj_vel_hist); normalisation and
// accumulation over the timeAvg calls stay here, exactly as the reference orders them
// (:154-166). The box-pair list argument is accepted and ignored: the device walks the cells
// within the cutoff directly instead of the reference's O(nbox^2) boxPairs table. No CPU path.
#ifndef APJ_HOST_CORRELATIONS_H
#define APJ_HOST_CORRELATIONS_H

#include "../../../include/apj_b200.h"
#include "Batch.h"

struct Correlations
{
    Correlations(double, double, double, double, long int, double);

    void spatialCorrelations(vector<vector<int>>&, vector<Box>&, vector<Cell>&);
    void autocorrelation(int, vector<double>&);
    void velDist(vector<Cell>&);
    void printCorrelations(int, Print&);
    double delta_norm(double);

    vector<double> orientation0;
    vector<double> orientationCorrelation;
    vector<double> velocityCorrelation;
    vector<double> pairCorrelationValues;
    vector<double> autocorrelationValues;
    vector<double> velocityDistributionValues;

    long int N;
    double L, Lover2, dens;
    double cutoff;
    int correlation_time;
    double dr_c, dr_p, dv;
    int np, nc, noBins;
    double norm;

    ApjBatch* batch = nullptr;         // set by Engine::start(): the device handle (shared by the replicas of a sweep)
    int system = 0;                    // this run's system index in the batch

    void bind(ApjBatch* b, int sys) { batch = b; system = sys; b->dv[sys] = dv; }

private:
    void need_device(const char* who) const {
        if (!batch || !batch->dev) { fprintf(stderr, "Correlations::%s: no device engine bound (this build has no CPU path)\n", who); exit(117); }
    }
};

inline Correlations::Correlations(double L_, double dens_, double cut, double time, long int N_, double CFself_)
    : N(N_), L(L_), Lover2(L_/2.0), dens(dens_), cutoff(cut), correlation_time((int)time),
      dr_c(2.0), dr_p(0.1), dv(CFself_/50.0), noBins(100)
{
    np = (int)ceil(cutoff/dr_p);
    nc = (int)ceil(cutoff/dr_c);
    norm = 2*L*L/(2*PI*dr_p*(double)(N*N));          // 2D shell normalisation of g(r) (reference :59)
    orientation0.assign(NDIM, 0);
    orientationCorrelation.assign(nc, 0);
    velocityCorrelation.assign(nc, 0);
    pairCorrelationValues.assign(np, 0);
    autocorrelationValues.assign(correlation_time, 0);
    velocityDistributionValues.assign(noBins, 0);
}

inline void Correlations::spatialCorrelations(vector<vector<int>>&, vector<Box>&, vector<Cell>&)
{
    need_device("spatialCorrelations");
    const double *counts, *ori, *vel, *pair;                      // apj_spatial_correlations, this system's rows
    batch->spatial_correlations(system, cutoff, nc, np, &counts, &ori, &vel, &pair);
    for (int k = 0; k < nc; k++) {                    // empty shells give 0/0 = NaN, as in the reference (Q14)
        orientationCorrelation[k] += ori[k]/counts[k];
        velocityCorrelation[k] += vel[k]/counts[k];
    }
    for (int k = 0; k < np; k++) pairCorrelationValues[k] += pair[k]*norm;
}

inline void Correlations::autocorrelation(int t, vector<double>& orientation)
{
    double dot = 0.0;
    for (int k = 0; k < NDIM; k++) dot += orientation0[k]*orientation[k];
    autocorrelationValues[t] += dot;
}

inline void Correlations::velDist(vector<Cell>&)
{
    need_device("velDist");
    const int64_t* h = batch->vel_hist(system);                   // apj_vel_hist
    // the reference adds 1/N per particle; adding count/N per bin differs only in the last bits
    for (int k = 0; k < noBins; k++) velocityDistributionValues[k] += (double)h[k]/(double)N;
}

// abscissae as the reference prints them, including its 0.01*k for g(r) (reference :189-205, Q14)
inline void Correlations::printCorrelations(int timeAvg, Print& print)
{
    for (size_t k = 0; k < velocityCorrelation.size(); k++) print.print_corr(dr_c*(k + 1), velocityCorrelation[k]/timeAvg);
    for (size_t k = 0; k < orientationCorrelation.size(); k++) print.print_orientationCorr(dr_c*(k + 1), orientationCorrelation[k]/timeAvg);
    for (size_t k = 0; k < pairCorrelationValues.size(); k++) print.print_pairCorr(0.01*k, pairCorrelationValues[k]/timeAvg);
    for (size_t t = 0; t < autocorrelationValues.size(); t++) print.print_autoCorr(t, autocorrelationValues[t]/timeAvg);
    for (size_t k = 0; k < velocityDistributionValues.size(); k++) print.print_velDist(k*dv, velocityDistributionValues[k]/timeAvg);
}

inline double Correlations::delta_norm(double delta)
{
    const double step = delta < -Lover2 ? L : -L;
    while (delta < -Lover2 || delta >= Lover2) delta += step;
    return delta;
}

#endif
