"""Python model of the pair arithmetic behind csrc/dkg_nsq.cuh: exponentiation modulo N^2 using only
Montgomery arithmetic modulo N.  An element x of Z_{N^2} is held as (a, b) with
x*R = a + b*N*R^-1 (mod N^2); squaring is a' = REDC(a^2) with quotient m, b' = REDC(2ab) - m;
multiplication a' = REDC(ac), b' = REDC(ad + bc) - m.  The model checks the identities and the
bounds the kernel relies on (R >= 8N: a < 2N, b < R, every REDC output < R)."""
from __future__ import annotations

import random


def redc_q(T: int, N: int, R: int, Ninv: int) -> tuple[int, int]:
    m = (T * Ninv) % R
    assert (T + m * N) % R == 0
    t = (T + m * N) // R
    assert T == t * R - m * N
    return t, m


def fixup(b2: int, m: int, N: int, R: int) -> int:
    """b'' - m brought back into [0, R) congruent modulo N, the way pair_fixup does it."""
    S = b2 + (R - m)
    if S >= R:
        return S - R
    U = S + (-R) % N
    return U - N if U >= R else U


def pair_sqr(a, b, N, R, Ninv):
    assert a < 2 * N and b < R
    t, m = redc_q(a * a, N, R, Ninv)
    s, _ = redc_q(2 * a * b, N, R, Ninv)
    assert t < 2 * N and s < R
    return t, fixup(s, m, N, R)


def pair_mul(a, b, c, d, N, R, Ninv):
    assert a < 2 * N and c < 2 * N and b < R and d < R
    t, m = redc_q(a * c, N, R, Ninv)
    s, _ = redc_q(a * d + b * c, N, R, Ninv)
    assert t < 2 * N and s < R
    return t, fixup(s, m, N, R)


def plain_pair(v, N, R):
    return v % N, (v // N) * R % N


def modexp_pair(c, e, N, R):
    N2 = N * N
    Ninv = (-pow(N, -1, R)) % R
    rho = pow(R, -1, N2)
    pR2, pR, p1 = plain_pair(pow(R, 2, N2), N, R), plain_pair(R % N2, N, R), (1, 0)
    x = pair_mul(*plain_pair(c, N, R), *pR2, N, R, Ninv)
    assert (x[0] + x[1] * N * rho) % N2 == c * R % N2
    acc = pR
    for bit in bin(e)[2:]:
        acc = pair_sqr(*acc, N, R, Ninv)
        if bit == "1":
            acc = pair_mul(*acc, *x, N, R, Ninv)
    a, b = pair_mul(*acc, *p1, N, R, Ninv)
    h, _ = redc_q(b, N, R, Ninv)
    return (a + (h % N) * N) % N2  # a is an integer < 2N: it must not be reduced without moving N into h


def test_pair_arithmetic_nilpotent_and_non_unit_bases():
    """Composite-radical cases (a component reaching N exactly) that once exposed a missing carry
    from a into h in the exit step."""
    N = 9 * 29
    R = 1 << 128
    for c in (41934, 87, 174 * 5, N * 7, 3 * 29 * 11):
        for e in (1, 2, 3, 4, 7, 64, 501409110107081):
            assert modexp_pair(c, e, N, R) == pow(c, e, N * N)


def test_pair_arithmetic_equals_pow():
    rng = random.Random(5)
    for bits in (20, 61, 130, 515, 2048, 2051):
        limbs = (bits + 3 + 31) // 32  # R >= 8N
        R = 1 << (32 * limbs)
        for _ in range(4):
            p = rng.getrandbits(bits // 2) | 1 | (1 << (bits // 2 - 1))
            q = rng.getrandbits(bits - bits // 2) | 1 | (1 << (bits - bits // 2 - 1))
            N = p * q
            assert 8 * N <= R
            for c in (0, 1, N, N * N - 1, rng.randrange(N * N)):
                e = rng.getrandbits(rng.choice([1, 8, 200]))
                assert modexp_pair(c, e, N, R) == pow(c, e, N * N)


def pair_value(a, b, N, R):
    """The element of Z_{N^2} a Montgomery pair stands for."""
    N2 = N * N
    rho = pow(R, -1, N2)
    return (a + b * N * rho) * rho % N2


def pair_inverse(a, b, N, R, Ninv):
    """Montgomery pair of x^-1 from the Montgomery pair (a, b) of x, the way the kernels do it
    (csrc/dkg_nsq.cuh, pair_invert): g = a^-1 mod N by a binary GCD on the a component only, the pair
    of y0 = g R (an inverse of x modulo N), one Newton step y0 (2 - x y0) in the pair domain."""
    N2 = N * N
    pR2 = plain_pair(pow(R, 2, N2), N, R)
    g = pow(a, -1, N)
    y0 = pair_mul(g, 0, *pR2, N, R, Ninv)
    y0 = pair_mul(*y0, *pR2, N, R, Ninv)                 # pair of g R: x y0 = 1 (mod N)
    assert pair_value(*y0, N, R) % N == pow(pair_value(a, b, N, R), -1, N)
    at, bt = pair_mul(*y0, a, b, N, R, Ninv)             # pair of x y0
    a2, b2 = plain_pair(2 * R % N2, N, R)                # pair of 2
    twoa, twob = a2 + 2 * N, (b2 - 2 * R) % N            # the same element with a larger a component
    ad = twoa - at                                       # > 0; up to 4N: allowed as the x operand of one product
    bd = fixup(twob, bt, N, R)
    assert 0 < ad < 4 * N and 0 <= bd < R
    assert pair_value(ad, bd, N, R) == (2 - pair_value(at, bt, N, R)) % N2
    # the product with the wide a component: same formulas, bounds checked here
    c, d = y0
    t, m = redc_q(ad * c, N, R, Ninv)
    s, _ = redc_q(ad * d + bd * c, N, R, Ninv)
    assert t < 2 * N and s < R
    return t, fixup(s, m, N, R)


def test_pair_inverse_newton_step():
    rng = random.Random(9)
    for bits in (20, 61, 130, 515, 2048, 2051):
        limbs = (bits + 3 + 31) // 32
        R = 1 << (32 * limbs)
        for _ in range(6):
            p = rng.getrandbits(bits // 2) | 1 | (1 << (bits // 2 - 1))
            q = rng.getrandbits(bits - bits // 2) | 1 | (1 << (bits - bits // 2 - 1))
            N = p * q
            N2 = N * N
            Ninv = (-pow(N, -1, R)) % R
            pR2 = plain_pair(pow(R, 2, N2), N, R)
            for x in (1, 2, N2 - 1, N + 1, rng.randrange(1, N2), rng.randrange(1, N2)):
                import math
                if math.gcd(x, N) != 1:
                    continue
                xa, xb = pair_mul(*plain_pair(x, N, R), *pR2, N, R, Ninv)
                assert pair_value(xa, xb, N, R) == x
                ya, yb = pair_inverse(xa, xb, N, R, Ninv)
                assert ya < 2 * N and yb < R
                assert pair_value(ya, yb, N, R) == pow(x, -1, N2)
