// Particle record of the B200 build. Same name, members and method signatures as the reference's
// `struct Cell` (reference code/classes/Cell.h:6-44), because Engine, Print, Fluctuations and
// Correlations -- and anything a user wrote against them -- address particles through these fields.
// Here it is a HOST MIRROR: the live state is the SoA copy in HBM owned by the apj_engine handle
// (include/apj_b200.h); Engine::pull_cells() refreshes the mirror when a caller wants to look.
//
// Like the reference header, this file expects the including translation unit to have defined
// NDIM, PI, PI2 and `using namespace std` (reference code/jam/jamming.cpp:1-4,15).
#ifndef APJ_HOST_CELL_H
#define APJ_HOST_CELL_H

#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

struct Cell
{
    Cell();

    void update(double&);
    void PBC();
    void periodicAngles();
    double get_speed();

    double L, Lover2, dt;

    double R;
    double Rinv;
    const double Zinv = (9*PI)/16;

    int over;
    int index;
    int box;

    vector<double> x;
    vector<double> x_real;
    vector<double> x0;
    vector<double> x_old;

    double vx, vy, vz;
    double Fx, Fy, Fz;

    double phi, theta;
    double cosp, sinp, cost, sint;
    double x_new, y_new, z_new;

    vector<int> VerletList;     // filled on request (Engine::pull_verlet_lists), partners j > index

    Cell(const Cell& o) { copy_from(o); }
    Cell& operator=(const Cell& o) { if (this != &o) copy_from(o); return *this; }
private:
    void copy_from(const Cell& o);
};

// Same defaults as a freshly constructed reference Cell (Cell.h:46-84): sizes unset (-1),
// positions parked at -100, 2D polar angle theta = PI/2.
inline Cell::Cell()
    : L(-1.0), Lover2(-1.0), dt(-1.0), R(-1.0), Rinv(0.0), over(0), index(-1), box(-1),
      x(NDIM, -100), x_real(NDIM, -100), x0(NDIM, -100), x_old(NDIM, -100),
      vx(0.0), vy(0.0), vz(0.0), Fx(0.0), Fy(0.0), Fz(0.0), phi(0.0), theta(PI/2.0),
      cosp(0.0), sinp(0.0), cost(0.0), sint(1.0), x_new(0.0), y_new(0.0), z_new(0.0)
{
    VerletList.reserve(20);
}

inline void Cell::copy_from(const Cell& o)
{
    L = o.L; Lover2 = o.Lover2; dt = o.dt; R = o.R; Rinv = o.Rinv; over = o.over; index = o.index; box = o.box;
    x = o.x; x_real = o.x_real; x0 = o.x0; x_old = o.x_old;
    vx = o.vx; vy = o.vy; vz = o.vz; Fx = o.Fx; Fy = o.Fy; Fz = o.Fz;
    phi = o.phi; theta = o.theta; cosp = o.cosp; sinp = o.sinp; cost = o.cost; sint = o.sint;
    x_new = o.x_new; y_new = o.y_new; z_new = o.z_new; VerletList = o.VerletList;
}

// The Euler update (reference Cell.h:86-158) is the epilogue of the fused step kernel
// (csrc/apj_step.cu); there is deliberately no host implementation of the hot path.
inline void Cell::update(double&)
{
    fprintf(stderr, "Cell::update: the per-particle update runs on the GPU inside apj_step(); "
                    "this build has no CPU path. Call Engine::calculate_next_positions().\n");
    exit(118);
}

// Host-side setup helpers used by Engine::initCells, same single-wrap semantics and truncated
// constants as the reference (Cell.h:160-175).
inline void Cell::periodicAngles()
{
    if (phi >= PI) phi -= PI2;
    else if (phi < -PI) phi += PI2;
}

inline void Cell::PBC()
{
    for (int k = 0; k < NDIM; k++) {
        if (x[k] >= Lover2) x[k] -= L;
        else if (x[k] < -Lover2) x[k] += L;
    }
}

inline double Cell::get_speed()
{
    return sqrt(vx*vx + vy*vy);
}

#endif
