"""Host restatement of the division-free sequences of the CUDA path, checked against exact
rational arithmetic: div6/div3 (mx_device.cuh) and div_h (mx_stream.cuh) must equal the
correctly rounded quotient, i.e. what the reference's IEEE division produces."""
from fractions import Fraction

import numpy as np
import pytest


def rn(fr: Fraction) -> float:
    """round-to-nearest-even of an exact rational to binary64 (Python's float(Fraction) is
    correctly rounded)."""
    return float(fr)


def fma(a: float, b: float, c: float) -> float:
    return rn(Fraction(a) * Fraction(b) + Fraction(c))


def div_markstein(x: float, h: float, steps: int) -> float:
    y = rn(Fraction(1) / Fraction(h))
    q = rn(Fraction(x) * Fraction(y))
    for _ in range(steps):
        r = fma(-h, q, x)
        q = fma(r, y, q)
    return q


def operands(n, seed):
    rng = np.random.default_rng(seed)
    xs = np.concatenate([
        rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n),
        np.ldexp(1.0 + rng.random(n), rng.integers(-200, 200, n)),
        np.nextafter(np.ldexp(1.0, rng.integers(-50, 50, n)), 0.0),  # all-ones mantissas
        [0.0, 1.0, 3.0, 6.0, 1e-300, 1e300],
    ])
    return [float(v) for v in xs]


@pytest.mark.parametrize("c", [6.0, 3.0])
def test_div_by_constant_one_step(c):
    # div6 / div3: one refinement step is exact for these divisors
    for x in operands(4000, 1):
        assert div_markstein(x, c, 1) == rn(Fraction(x) / Fraction(c)), (x, c)


@pytest.mark.parametrize("h", [1.0 / 384.0, 3.6e8 / 384.0, 5.0e8 / 512.0, 1.0 / 12.0, 0.1,
                               1.0 / 3.0, float(np.nextafter(2.0, 0.0)), 7.03125e5])
def test_div_h_two_steps(h):
    for x in operands(3000, 2):
        if x != 0.0 and not (1e-280 < abs(x) < 1e280):
            continue  # intermediate under/overflow is excluded by the contract
        assert div_markstein(x, h, 2) == rn(Fraction(x) / Fraction(h)), (x, h)
