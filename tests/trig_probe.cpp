// Accuracy probe of csrc/apj_trig.h (the sin / cos of the noise term in the step kernel epilogue), host build:
// prints the worst error against long double over |x| <= 3. Used by tests/test_epilogue_identity.py.
#include <cstdio>
#include <cmath>
#include <random>
#include "apj_trig.h"
int main() {
    std::mt19937_64 g(1);
    std::uniform_real_distribution<double> U(-3.0, 3.0);
    double worst_s = 0, worst_c = 0, ws = 0, wc = 0;
    long bad = 0;
    for (long i = 0; i < 4000000; i++) {
        double x = i < 1000 ? (-3.0 + 6.0 * i / 999.0) : U(g);
        if (i >= 1000 && i < 2000) x = (i - 1500) * 1e-9;
        if (i >= 2000 && i < 4000) x = ((i & 7) - 3.5) * 0.7853981633974483 + (i - 3000) * 1e-12;   // near odd multiples of pi/4 (k ties)
        if (std::fabs(x) > 3.0) continue;
        double s, c;
        apj_sincos_pm3(x, &s, &c);
        long double es = sinl((long double)x), ec = cosl((long double)x);
        double us = std::fabs((double)((long double)s - es)) / std::fmax(std::fabs((double)es) * 1.1102230246251565e-16, 1e-300);
        double uc = std::fabs((double)((long double)c - ec)) / std::fmax(std::fabs((double)ec) * 1.1102230246251565e-16, 1e-300);
        // ulp-ish: error relative to half-ulp scale of the exact value (2 = 1 ulp)
        if (us > worst_s) { worst_s = us; ws = x; }
        if (uc > worst_c) { worst_c = uc; wc = x; }
        double abs_s = std::fabs((double)((long double)s - es)), abs_c = std::fabs((double)((long double)c - ec));
        if (abs_s > 2.3e-16 || abs_c > 2.3e-16) bad++;
    }
    printf("worst sin err %.3f half-ulps at %.17g; worst cos err %.3f half-ulps at %.17g; abs>2.3e-16: %ld\n", worst_s, ws, worst_c, wc, bad);
    return 0;
}
