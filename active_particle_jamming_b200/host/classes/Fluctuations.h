// Density-fluctuation observables with the reference's interface (reference
// code/classes/Fluctuations.h:7-30): giant-number-fluctuation scaling from a growing circle
// centred on the centre of mass, and the histogram of cell occupancies.
//
// The O(N) work -- the lens-area sum over all disks for one radius (reference :62-76) and the
// occupancy histogram (:122-139) -- runs on the GPU (apj_fluct_area, apj_occupancy_hist); this
// class keeps what is sequential bookkeeping in the reference too: the radius / sample-count
// state machine and the accumulators. bind() must be called before the first
// measurement; the vector<Cell>/vector<Box> arguments are accepted for source compatibility
// and not read (the live state is in HBM). No CPU path.
#ifndef APJ_HOST_FLUCTUATIONS_H
#define APJ_HOST_FLUCTUATIONS_H

#include "../../../include/apj_b200.h"
#include "Batch.h"

struct Fluctuations
{
    Fluctuations(double, int, int, double);

    double overlap(double, double, double);
    void measureFluctuations(vector<Cell>&, vector<double>&, Print&);
    double delta_norm(double);
    void density_distribution(vector<Cell>&, vector<Box>&);
    void print_density_distribution(int, Print&);

    vector<double> distribution;

    int nPoints;
    double min_radius;
    double max_radius;
    double current_radius;
    double rad_interval;
    double time_interval;
    int counter;
    double current_value;
    double dens;

    double Lover2, L;

    ApjBatch* batch = nullptr;         // set by Engine::start(): the device handle (shared by the replicas of a sweep)
    int system = 0;                    // this run's system index in the batch

    void bind(ApjBatch* b, int sys) { batch = b; system = sys; b->radius_ptr[sys] = &current_radius; }

private:
    void need_device(const char* who) const {
        if (!batch || !batch->dev) { fprintf(stderr, "Fluctuations::%s: no device engine bound (this build has no CPU path)\n", who); exit(117); }
    }
};

// 10 measurement radii, geometric from 3 to L/2 in 11 steps; steps/(skip*10) samples per radius
// (reference :32-49).
inline Fluctuations::Fluctuations(double L_, int totalSteps, int skip, double dens_)
    : distribution(50, 0.0), nPoints(10), min_radius(3.0), max_radius(L_/2.0), current_radius(3.0),
      counter(0), current_value(0), dens(dens_), Lover2(L_/2.0), L(L_)
{
    rad_interval = pow(max_radius/min_radius, 1./((double)nPoints + 1.0));
    time_interval = (double)totalSteps/((double)skip*(double)nPoints);
}

inline void Fluctuations::measureFluctuations(vector<Cell>&, vector<double>&, Print& print)
{
    need_device("measureFluctuations");
    const double expectedV = dens*PI*current_radius*current_radius;
    if (counter < time_interval) {
        const double V = batch->fluct_area(system, current_radius);   // total disk area inside the circle around COM (apj_fluct_area)
        current_value += (V - expectedV)*(V - expectedV);
    } else {                           // quota reached: flush this radius, grow the circle; no sample on this call
        current_value = sqrt(current_value/(double)counter);
        print.print_fluct(current_radius, expectedV, current_value);
        current_value = 0;
        counter = -1;
        current_radius = current_radius*rad_interval;
    }
    counter++;
}

// Lens area of a disk (radius r) cut by a circle (radius R) whose centres are d apart -- kept as a
// host utility with the reference's signature (reference :89-120); the device evaluates the same
// expression per particle (csrc/apj_observe.cu lens_overlap).
inline double Fluctuations::overlap(double r, double R, double d)
{
    if (R >= r + d) return PI*r*r;
    const double r2 = r*r, Rsq = R*R;
    const double a = (r2 - Rsq + d*d)/(2.0*d);      // distance from the disk centre to the chord
    const double ta = acos(a/r);
    const double c = d - a;
    const double tc = acos(c/R);
    return (r2*ta - a*r*sin(ta)) + (Rsq*tc - c*R*sin(tc));
}

inline void Fluctuations::density_distribution(vector<Cell>&, vector<Box>&)
{
    need_device("density_distribution");
    const int64_t* h = batch->occupancy_hist(system);             // apj_occupancy_hist
    for (size_t k = 0; k < distribution.size(); k++) distribution[k] += (double)h[k];
}

inline void Fluctuations::print_density_distribution(int timeAvg, Print& print)
{
    for (size_t k = 0; k < distribution.size(); k++) print.print_dens(k, distribution[k]/timeAvg);
}

inline double Fluctuations::delta_norm(double delta)
{
    const double step = delta < -Lover2 ? L : -L;
    while (delta < -Lover2 || delta >= Lover2) delta += step;
    return delta;
}

#endif
