"""The step kernel forms cos/sin of phi = atan2(sum sin, sum cos) + eta by the angle-addition identity and applies the
reference's truncated-constant wrap (Cell.h:160-166) as a rotation by PI2 - 2 pi (csrc/apj_step.cu, "Epilogue" in
DESIGN.md). This restates that arithmetic in numpy WITH THE CONSTANTS PARSED FROM THE KERNEL SOURCE and compares it with
the reference form cos(wrap(atan2 + eta)) on millions of samples: every wrap decision equal, values within 2e-15
(the parity gate is 1e-12)."""
import math
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "active_particle_jamming_b200", "csrc", "apj_step.cu")).read()
PI, PI2 = 3.14159265, 6.28318531            # reference jamming.cpp:3-4


def const(name):
    m = re.search(r"constexpr double %s = ([0-9.eE+-]+);" % name, SRC)
    assert m, name
    return float(m.group(1))


def test_constants_are_what_the_truncated_literals_imply():
    assert const("SIN_PI") == math.sin(PI)                       # sin evaluated on the double 3.14159265
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    two_pi = 2 * Decimal("3.14159265358979323846264338327950288419716939937510582")
    assert abs(const("DPI2") - float(Decimal(PI2) - two_pi)) < 1e-24


def fast_form(ax, ay, nz):
    sin_pi, dpi2 = const("SIN_PI"), const("DPI2")
    inv = 1.0 / np.sqrt(ax * ax + ay * ay)
    c0, s0 = ax * inv, ay * inv
    cn, sn = np.cos(nz), np.sin(nz)
    c, s = c0 * cn - s0 * sn, s0 * cn + c0 * sn
    hi = (nz > 0) & (ay >= 0) & ((s < 0) | ((c < 0) & (s <= sin_pi)))
    lo = (nz < 0) & (ay < 0) & ((s > 0) | ((c < 0) & (s > -sin_pi)))
    return np.where(hi, c + s * dpi2, np.where(lo, c - s * dpi2, c)), np.where(hi, s - c * dpi2, np.where(lo, s + c * dpi2, s)), hi | lo


def reference_form(ax, ay, nz):
    phi = np.arctan2(ay, ax) + nz
    wrapped = (phi >= PI) | (phi < -PI)
    phi = np.where(phi >= PI, phi - PI2, np.where(phi < -PI, phi + PI2, phi))
    return np.cos(phi), np.sin(phi), wrapped


def test_identity_matches_the_reference_form():
    rng = np.random.default_rng(3)
    n = 1_500_000
    for ctnoise in (0.05, 0.5, 0.95):
        k = rng.integers(1, 20, n)                               # alignment sum over 1..19 unit vectors
        ax, ay = np.zeros(n), np.zeros(n)
        for j in range(19):
            th = rng.uniform(-np.pi, np.pi, n)
            ax += np.where(k > j, np.cos(th), 0.0); ay += np.where(k > j, np.sin(th), 0.0)
        u = rng.integers(0, 2 ** 32, n).astype(np.float64) / 4294967296.0 * (PI - (-PI)) + (-PI)   # randuni lattice (Q5)
        nz = ctnoise * u
        ok = (np.abs(nz) >= 1e-7) & (np.abs(nz) <= 3.0)          # the kernel's fast-path window
        cf, sf, wf = fast_form(ax[ok], ay[ok], nz[ok])
        cr, sr, wr = reference_form(ax[ok], ay[ok], nz[ok])
        assert np.array_equal(wf, wr)
        assert np.max(np.abs(cf - cr)) <= 2e-15 and np.max(np.abs(sf - sr)) <= 2e-15
        assert ctnoise < 0.3 or wr.mean() > 0.05                 # wraps do occur in the sample


def test_identity_on_the_axes_and_near_the_wrap():
    eta = np.array([1e-7, 0.3, 1.5, 2.9, 3.0, -1e-7, -0.3, -1.5, -2.9, -3.0])
    for ax, ay in ((1.0, 0.0), (-1.0, 0.0), (0.0, 1.0), (0.0, -1.0), (-3.0, 1e-12), (-3.0, -1e-12), (2.0, 2.0), (-2.0, -2.0)):
        a, b = np.full_like(eta, ax), np.full_like(eta, ay)
        cf, sf, wf = fast_form(a, b, eta)
        cr, sr, wr = reference_form(a, b, eta)
        assert np.array_equal(wf, wr), (ax, ay)
        assert np.max(np.abs(cf - cr)) <= 2e-15 and np.max(np.abs(sf - sr)) <= 2e-15, (ax, ay)


def test_noise_sincos_kernel_is_good_to_an_ulp(tmp_path):
    """csrc/apj_trig.h (Cody-Waite reduction + minimax kernels, explicit fma) replaces the library sincos() for the
    noise term, |x| <= 3: compiled for the host and compared with long double on 4e6 arguments incl. the reduction
    ties near odd multiples of pi/4 and tiny arguments. Error <= 1 ulp of the result, <= 2.3e-16 absolute."""
    import re
    import subprocess
    exe = str(tmp_path / "trig_probe")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "active_particle_jamming_b200", "csrc"),
                    os.path.join(ROOT, "tests", "trig_probe.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    m = re.search(r"worst sin err ([0-9.]+) half-ulps.*worst cos err ([0-9.]+) half-ulps.*abs>2.3e-16: (\d+)", out)
    assert m, out
    assert float(m.group(1)) <= 2.0 and float(m.group(2)) <= 2.0 and int(m.group(3)) == 0, out
    assert "apj_sincos_pm3(nz, &sn_n, &cn_n)" in SRC                # ... and it is what the kernel calls
