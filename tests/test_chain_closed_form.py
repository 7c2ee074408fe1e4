"""The numerical claim behind draw.cu's build_chain, checked on the CPU in numpy float32: the reference's
`srcPos += dx` accumulation (images.nim:588) equals a piecewise closed form — within one binade the rounded sum
advances by a constant multiple of the binade's ulp (after at most one step that settles a round-half-even tie),
so position k is `base + (k - k0) * inc` with both operations exact.  The model below is the algorithm of
build_chain statement for statement; the GPU kernel itself is held to the oracle in tests/test_gpu_draw.py."""
import math
import struct

import numpy as np

f = np.float32


def _bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def _frombits(b):
    return f(struct.unpack("<f", struct.pack("<I", b))[0])


def build_chain(v0, d, n):
    segs, k, cur, d = [], 0, f(v0), f(d)
    while k < n:
        p1 = f(cur + d)
        p2 = f(p1 + d)
        p3 = f(p2 + d)
        b2 = _bits(p2)
        e1, e2, e3 = _bits(p1) >> 23, b2 >> 23, _bits(p3) >> 23
        run = e1 == e2 == e3 and (e2 & 0xFF) not in (0, 0xFF)
        segs.append((k, cur, f(0)))
        if not run:
            cur = p1
            k += 1
            continue
        segs.append((k + 1, p1, f(0)))
        inc = f(p3 - p2)
        J = n
        if inc != 0:
            lo = _frombits(b2 & 0x7F800000)
            u = float(lo) / 8388608.0
            a2 = abs(float(p2))
            ai = -float(inc) if p2 < 0 else float(inc)
            if ai > 0:
                hi = 2.0 * float(lo)
                J = int(math.floor((hi - a2) / ai))
                while J > 0 and a2 + J * ai >= hi:
                    J -= 1
                while a2 + (J + 1) * ai < hi:
                    J += 1
            else:
                lo1 = float(lo) + u
                J = int(math.floor((a2 - lo1) / -ai))
                while J > 0 and a2 + J * ai < lo1:
                    J -= 1
                while a2 + (J + 1) * ai >= lo1:
                    J += 1
            J = min(max(J, 1), n)
        segs.append((k + 2, p2, inc))
        cur = f(f(p2 + f(f(J) * inc)) + d)
        k = k + 2 + J + 1
    return segs


def closed_form(segs, n):
    out = np.zeros(n, np.float32)
    starts = [s[0] for s in segs] + [1 << 40]
    j = 0
    for k in range(n):
        while starts[j + 1] <= k:
            j += 1
        k0, base, inc = segs[j]
        out[k] = f(base + f(f(k - k0) * inc))
    return out


def sequential(v0, d, n):
    out = np.zeros(n, np.float32)
    cur, d = f(v0), f(d)
    for k in range(n):
        out[k] = cur
        cur = f(cur + d)
    return out


def test_closed_form_equals_sequential_accumulation():
    rng = np.random.default_rng(5)
    worst = 0
    for trial in range(160):
        kind = trial % 8
        n = int(rng.integers(1, 2500))
        if kind == 0:
            v0, d = rng.uniform(-5000, 5000), rng.uniform(-2, 2)
        elif kind == 1:
            v0, d = rng.uniform(-50, 50), rng.uniform(-1e-3, 1e-3)
        elif kind == 2:
            v0, d = 0.0, rng.uniform(-1, 1) * 10 ** rng.uniform(-8, 0)
        elif kind == 3:  # dyadic increments: exact rounding ties
            v0, d = rng.uniform(-4000, 4000), float(f(rng.integers(1, 64)) / f(64)) * rng.choice([-1, 1])
        elif kind == 4:
            v0, d = float(rng.integers(-3000, 3000)) + 0.5, 1.0
        elif kind == 5:  # crossing zero
            v0, d = rng.uniform(-1, 1) * 1e-3, rng.uniform(-1, 1)
        elif kind == 6:
            v0, d = rng.uniform(2000, 8192), -rng.uniform(0.3, 1.9)
        else:
            v0, d = float(f(rng.uniform(-100, 100))), float(f(1.5) * f(2.0 ** -int(rng.integers(0, 20))))
        segs = build_chain(v0, d, n)
        worst = max(worst, len(segs))
        a, b = closed_form(segs, n), sequential(v0, d, n)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (trial, kind, v0, d, n)
    assert worst <= 124  # the kernel's table holds 128 segments; longer chains fall back to the sequential form


def test_single_fma_div_65280_is_exact():
    """blur_mma.cu quantises an accumulator with one FFMA.RZ: floor(a * float(0x37808081)) must equal
    a div 256 div 255 (images.nim:332-338) for every integer a < 2^24.  The product is formed exactly by the FMA, so
    the check is integer arithmetic on the float's mantissa: 0x808081 * 2^-39."""
    a = np.arange(0, 1 << 24, dtype=np.uint64)
    q = (a * np.uint64(0x808081)) >> np.uint64(39)
    assert np.array_equal(q, (a // np.uint64(256)) // np.uint64(255))
    assert np.array_equal(q, a // np.uint64(65280))
