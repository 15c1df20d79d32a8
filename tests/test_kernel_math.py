"""CPU checks of the branch-free radial functions of the gradient kernel (gpyreg_b200/csrc/cov.cuh: exp_neg,
sqrt_rsqrt_nonneg).  The gradient tolerance (1e-7) would not notice a mistyped polynomial coefficient, so the
constants are read from the source and the algorithms are restated here step by step -- every fused multiply-add
evaluated in extended precision and rounded once -- and held to a couple of ulps against the library functions."""
import math
import os
import re
from fractions import Fraction

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpyreg_b200", "csrc", "cov.cuh")
LD = np.longdouble


def _constants():
    src = open(SRC).read()
    body = re.search(r"__constant__ double EXP_C\[16\] = \{(.*?)\};", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body)
    return [float(v) for v in body.replace("\n", " ").split(",")]


def _fma(a, b, c):
    return np.float64(LD(a) * LD(b) + LD(c))


def _exp_neg(x, C):
    x = np.minimum(x, 708.0)
    t = _fma(x, -C[0], C[1])
    nf = t - C[1]
    r = _fma(nf, C[2], -x)
    r = _fma(nf, C[3], r)
    p = np.full_like(x, C[4])
    for q in range(5, 16):
        p = _fma(p, r, C[q])
    p = _fma(p, r, C[15])
    return np.ldexp(p, nf.astype(int))


def test_exp_constants_are_what_the_comments_say():
    C = _constants()
    assert len(C) == 16
    assert C[0] == 1.4426950408889634 and C[1] == 1.5 * 2.0 ** 52
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    assert abs(Decimal(-C[2]) + Decimal(-C[3]) - Decimal(2).ln()) < Decimal(10) ** -32   # hi + lo = ln 2 to 106 bits
    assert -C[2] == math.log(2)                                   # hi part = the double nearest ln 2
    for q, k in zip(range(4, 16), range(12, 0, -1)):
        assert C[q] == float(Fraction(1, math.factorial(k))), (q, k)


def test_exp_neg_is_accurate_to_two_ulps():
    C = _constants()
    rng = np.random.default_rng(0)
    x = np.concatenate([np.linspace(0.0, 40.0, 100001), rng.uniform(0.0, 700.0, 50000), [0.0, 1e-300, 707.9]])
    got = _exp_neg(x, C)
    ref = np.exp(-x.astype(LD))
    assert float(np.max(np.abs((got - ref) / ref))) <= 2 * np.finfo(np.float64).eps
    # beyond the clamp: a tiny positive number instead of a denormal / zero, never NaN or negative
    big = _exp_neg(np.array([708.0, 745.0, 1e6, 1e300]), C)
    assert np.all(big > 0) and np.all(big < 1e-300)


def test_goldschmidt_sqrt_is_accurate_to_two_ulps():
    rng = np.random.default_rng(1)
    x = np.concatenate([10.0 ** rng.uniform(-200, 200, 50000), rng.uniform(0, 100, 50000)])
    # hardware seed: relative error below 2^-22 (rsqrt.approx.ftz.f64); take the worst case on both sides
    for seed_err in (2.0 ** -22, -2.0 ** -22):
        y = (1.0 / np.sqrt(x)) * (1.0 + seed_err)
        g, h = x * y, 0.5 * y
        r = _fma(-g, h, 0.5)
        g, h = _fma(g, r, g), _fma(h, r, h)
        r = _fma(-g, h, 0.5)
        g, h = _fma(g, r, g), _fma(h, r, h)
        ref = np.sqrt(x.astype(LD))
        assert float(np.max(np.abs((g - ref) / ref))) <= 2 * np.finfo(np.float64).eps
        assert float(np.max(np.abs((2.0 * h * ref) - 1.0))) <= 4 * np.finfo(np.float64).eps
