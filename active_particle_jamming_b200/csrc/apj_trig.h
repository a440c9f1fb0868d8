// sin and cos of the noise term of one step, |x| <= 3 (the fast epilogue of apj_step.cu takes the library path outside
// that range). One Cody-Waite reduction by k * pi/2 with |k| <= 2 (k * PIO2_HI is exact, so is x - k * PIO2_HI) and the two
// classical minimax kernels on [-pi/4, pi/4] (the fdlibm coefficients: error < 2^-58 before rounding), evaluated with
// explicit fma: ~35 double-precision instructions and no branch, against ~70 plus a slow-path test for the general
// sincos(). Accuracy over the range: <= 1 ulp for both (tests/test_epilogue_identity.py compiles this header for the
// host and compares with long double), i.e. the same as the library routine it replaces.
// Plain C++ (host + device): no CUDA header needed.
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define APJ_HD __host__ __device__ __forceinline__
#else
#define APJ_HD inline
#endif

APJ_HD void apj_sincos_pm3(const double x, double* sn, double* cs) {
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    const double PIO2_HI = 1.57079632673412561417e+00;   // first 33 bits of pi/2
    const double PIO2_LO = 6.07710050650619224932e-11;   // pi/2 - PIO2_HI
    const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52: adding it rounds to the nearest integer
    const double t = fma(x, TWO_OVER_PI, MAGIC);
    const double kf = t - MAGIC;                         // k = nearest integer to x * 2/pi, |k| <= 2
    double r = fma(-kf, PIO2_HI, x);
    r = fma(-kf, PIO2_LO, r);
    const double z = r * r;
    // sin(r) = r + r^3 (S1 + z (S2 + ... ))
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double s = fma(r * z, ps, r);
    // cos(r) = 1 - z/2 + z^2 (C1 + z (C2 + ... ))
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double hz = 0.5 * z;
    const double w = 1.0 - hz;
    const double c = w + (((1.0 - w) - hz) + z * z * pc);   // 1 - hz carried with its rounding error (fdlibm's k_cos form)
    // quadrant k mod 4 (the low bits of t's mantissa hold k in two's complement)
    const int q = (int)kf & 3;
    const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}
