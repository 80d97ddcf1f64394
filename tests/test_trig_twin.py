"""CPU twin of tp_sin / tp_cos (extendablefem.jl_b200/csrc/fastplan.cuh): the coefficient table is PARSED from the CUDA source and the
algorithm (three-constant Cody-Waite reduction, fdlibm kernels, select by quadrant) restated in numpy, then compared with libm over
the argument range the kernel serves before it falls back to the library function (|a| < 2^20).  Guards the constants; the GPU
side is tests/test_gpu_round2b.py::test_rhs_fast_trig_ranges."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _coefficients():
    src = open(os.path.join(ROOT, "extendablefem.jl_b200", "csrc", "fastplan.cuh")).read()
    body = re.search(r"__constant__ double c_tp_trig\[16\] = \{(.*?)\};", src, re.S).group(1)
    body = re.sub(r"//[^\n]*", "", body)
    vals = [float(t) for t in body.replace("\n", " ").split(",") if t.strip()]
    assert len(vals) == 16
    return np.array(vals)


def _twin(a, shift, c):
    q = np.rint(a * c[0])
    k = q.astype(np.int64) + shift
    r = ((a + q * c[1]) + q * c[2]) + q * c[3]
    z = r * r
    ps = c[4]
    for i in range(5, 10):
        ps = ps * z + c[i]
    sn = r + (z * r) * ps
    pc = c[10]
    for i in range(11, 16):
        pc = pc * z + c[i]
    cs = (1.0 - 0.5 * z) + (z * z) * pc
    v = np.where(k & 1, cs, sn)
    return np.where(k & 2, -v, v)


def test_coefficient_table_is_the_documented_one():
    c = _coefficients()
    assert c[0] == 2.0 / np.pi
    # pi/2 in three parts: the leading one is the double nearest to pi/2, the sum reproduces pi/2 far beyond double precision
    assert -c[1] == np.pi / 2 and abs(c[2]) < 2.0 ** -53 and abs(c[3]) < 2.0 ** -103
    from fractions import Fraction
    pio2 = Fraction(314159265358979323846264338327950288419716939937510, 10 ** 50) / 2
    assert abs(-(Fraction(c[1]) + Fraction(c[2]) + Fraction(c[3])) - pio2) < Fraction(1, 10 ** 45)
    # Taylor-like leading coefficients of the fdlibm kernels
    assert abs(c[9] + 1.0 / 6.0) < 1e-15 and abs(c[15] - 1.0 / 24.0) < 1e-15


def test_twin_matches_libm():
    c = _coefficients()
    rng = np.random.default_rng(0)
    for scale in (1.0, 10.0, 300.0, 5.0e4, 1.0e6):
        a = np.concatenate([rng.uniform(-scale, scale, 200000), np.linspace(-scale, scale, 100001)])
        a = a[np.abs(a) < 1048576.0]
        # without fused multiply-adds the reduction loses up to |q| ulp(pi/2) / 2: compare at that level (the device uses fma)
        tol = 4e-16 + np.abs(a) * 1.5e-16
        assert (np.abs(_twin(a, 0, c) - np.sin(a)) <= tol).all()
        assert (np.abs(_twin(a, 1, c) - np.cos(a)) <= tol).all()
    # on the reduced interval the kernels themselves are good to an ulp
    r = np.linspace(-np.pi / 4, np.pi / 4, 400001)
    assert np.abs(_twin(r, 0, c) - np.sin(r)).max() < 2.3e-16 and np.abs(_twin(r, 1, c) - np.cos(r)).max() < 2.3e-16
