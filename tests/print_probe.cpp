// Drives every Print::print_* once with fixed values. Compiled twice: against the reference's
// Print.h (tests/golden/make_golden.py, where /root/reference exists -> tests/golden/print_bytes.json)
// and against the host mirror (tests/test_host_driver.py); the files must be byte-identical.
#define NDIM 2
#define PI 3.14159265
#define PI2 6.28318531
#include <vector>
#include <cmath>
#include <cstdlib>
using namespace std;
#include "Print.h"

int main(int argc, char** argv) {
    if (argc != 2) return 2;
    Print p(argv[1], "probe", "run7", 64, false);
    vector<double> v2 = {1.0 / 3.0, -2.5e-7}, x = {12.345678901, -0.000123456}, vel = {0.05, -0.049999, 0.0};
    for (long t = 0; t <= 200; t += 100) {
        p.print_COM(t, v2);
        p.print_orientation(t, v2);
        p.print_order(t, 0.987654321 / (t + 1));
        p.print_MSD((int)t, 1e-5 * (t + 1));
        p.print_autoCorr((int)t / 100, 0.25 - t);
    }
    p.print_fluct(3.0, 25.4469, 1.2345678);
    p.print_fluct(4.5, 1e6 / 7.0, 0.0);
    for (int k = 0; k < 3; k++) {
        p.print_corr(2.0 * (k + 1), 0.5 / (k + 1));
        p.print_orientationCorr(2.0 * (k + 1), nan(""));
        p.print_pairCorr(0.01 * k, 1.0 + k * 1e-3);
        p.print_velDist(k * 0.001, 0.1 * k);
        p.print_dens(k, 12.5 * k);
    }
    int k0 = 0, k1 = 1, n = 2, i0 = 0, i1 = 1, over = 240;
    double R = 1.0375;
    p.print_Ovito(k0, n, i0, R, over, x, vel);
    p.print_Ovito(k1, n, i1, R, over, x, vel);
    p.print_summary("run7", 64, 15.0123456789, 10001, 10, 0.05, 0.5, 0.9, 42, 17, 0.6543210987, 0.123456789, 1e-4);
    return 0;
}
