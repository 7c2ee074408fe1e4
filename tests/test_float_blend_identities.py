"""The arithmetic identities behind the float blend modes' fast path (pixie_b200/csrc/cuda/common.cuh: round_bits,
div255f, fdiv_r), checked on the CPU with exact rational arithmetic — every float32 operation below is the exactly
rounded one (round to nearest even, or toward zero where the kernel says so), an FMA rounds once.  The kernels
themselves are held to the oracle in tests/test_gpu_blend_blur.py."""
import math
import random
from fractions import Fraction

import numpy as np

f = np.float32


def _round(fr, toward_zero=False):
    """Fraction -> float32, round to nearest even (or toward zero); normal range only."""
    if fr == 0:
        return f(0)
    s = 1 if fr > 0 else -1
    a = abs(fr)
    e = math.floor(math.log2(float(a)))
    while Fraction(2) ** e > a:
        e -= 1
    while Fraction(2) ** (e + 1) <= a:
        e += 1
    ulp = Fraction(2) ** (e - 23)
    q = a / ulp
    n = q.numerator // q.denominator
    rem = q - n
    if not toward_zero and (rem > Fraction(1, 2) or (rem == Fraction(1, 2) and n % 2 == 1)):
        n += 1
    return f(float(s * n * ulp))


def F(x):
    return Fraction(float(x))


def fma(a, b, c):
    return _round(F(a) * F(b) + F(c))


def test_div255f_is_the_ieee_quotient_for_every_byte():
    r = f(1.0) / f(255.0)
    for u in range(256):
        q = f(f(u) * r)
        got = fma(fma(-q, f(255.0), f(u)), r, q)
        assert got == f(u) / f(255.0), u


def test_round_bits_is_roundf():
    """floor(v + 0.5) through two round-toward-zero additions == roundf(v) for 0 <= v < 2^22."""
    rng = random.Random(5)
    vals = [f(k + 0.5) for k in range(0, 300)] + [f(0.49999997), f(0.5), f(254.5), f(255.0), f(65024.5), f(4194303.5)]
    vals += [f(rng.uniform(0, 70000)) for _ in range(20000)]
    vals += [_round(Fraction(k) + Fraction(1, 2) - Fraction(1, 2 ** 20)) for k in range(1, 2000, 7)]
    for v in vals:
        y = _round(F(v) + Fraction(1, 2), toward_zero=True)
        z = _round(F(y) + 8388608, toward_zero=True)
        n = int(np.float32(z).view(np.uint32)) - 0x4B000000
        assert n == int(math.floor(float(v) + 0.5)), v  # roundf: half away from zero == half up for v >= 0


def test_shared_reciprocal_division_is_correctly_rounded():
    """q = x * RN(1/d) refined twice with the exact FMA residual == RN(x / d)."""
    rng = random.Random(11)
    for i in range(4000):
        if i % 2:
            x = f(rng.uniform(0, 1.2))
            d = f(rng.randint(1, 65025) / 65025.0)
        else:
            x = f(rng.uniform(-2, 2) * 10.0 ** rng.randint(-9, 3))
            d = f(rng.uniform(1e-6, 4.0))
        r = _round(1 / F(d))
        q = f(x * r)
        q = fma(fma(-q, d, x), r, q)
        q = fma(fma(-q, d, x), r, q)
        assert q == _round(F(x) / F(d)), (x, d)
