"""Python model of exponentiation modulo N^2 with pair arithmetic modulo N (validation of the math)."""
import random
rng = random.Random(5)

def setup(N, Lbits):
    R = 1 << Lbits
    assert N < R and N % 2 == 1
    Ninv = (-pow(N, -1, R)) % R
    return R, Ninv

def redc_q(T, N, R, Ninv):
    """Montgomery reduction of T < R*N... returns (t, m) with T = t*R - m*N exactly, 0 <= t < (T/R)+N."""
    m = (T * Ninv) % R
    t = (T + m * N) // R
    assert (T + m * N) % R == 0
    return t, m

def norm(v, N, R):
    """bring an integer v (any sign, |v| < few R) to [0,R) congruent mod N the way the kernel would"""
    return v % N  # model: any representative works; kernel picks one in [0,R)

def pair_sqr(a, b, N, R, Ninv):
    t, m = redc_q(a * a, N, R, Ninv)
    s, _ = redc_q(2 * a * b, N, R, Ninv)
    # almost-Montgomery: keep a' in [0,R): subtract N if t >= R, compensating b' += R
    comp = 0
    if t >= R:
        t -= N; comp = R
    return t, (s - m + comp) % N

def pair_mul(a, b, c, d, N, R, Ninv):
    t, m = redc_q(a * c, N, R, Ninv)
    s, _ = redc_q(a * d + b * c, N, R, Ninv)
    comp = 0
    if t >= R:
        t -= N; comp = R
    return t, (s - m + comp) % N

def pair_value(a, b, N, R):
    rho = pow(R, -1, N * N)
    return (a + b * N * rho) % (N * N)

def to_pair_plain(c, N, R):
    c0, c1 = c % N, c // N
    return c0, (c1 * R) % N     # represents c itself

def modexp_pair(c, e, N, R, Ninv):
    N2 = N * N
    # constants (host): pair of R^2 mod N^2 (to enter the Montgomery domain) and of 1 (to leave it)
    pR2 = to_pair_plain(pow(R, 2, N2), N, R)
    p1 = to_pair_plain(1, N, R)
    x = pair_mul(*to_pair_plain(c, N, R), *pR2, N, R, Ninv)           # c*R
    assert pair_value(*x, N, R) == (c * R) % N2
    acc = pair_mul(*p1, *pR2, N, R, Ninv)                               # R  (one)
    for bit in bin(e)[2:]:
        acc = pair_sqr(*acc, N, R, Ninv)
        if bit == '1':
            acc = pair_mul(*acc, *x, N, R, Ninv)
    y = pair_mul(*acc, *p1, N, R, Ninv)                                  # leave the domain: value y
    a, b = y
    h, _ = redc_q(b, N, R, Ninv)
    return (a % N + (h % N) * N) % N2, (a, b)

for bits, Lbits in [(60, 64), (64, 64), (130, 160), (2050, 2112), (2048, 2048)]:
    for _ in range(5):
        p = rng.getrandbits(bits // 2) | 1 | (1 << (bits // 2 - 1)); q = rng.getrandbits(bits - bits // 2) | 1 | (1 << (bits - bits // 2 - 1))
        N = p * q
        if N.bit_length() > Lbits: continue
        R, Ninv = setup(N, Lbits)
        c = rng.randrange(N * N); e = rng.getrandbits(300)
        got, _ = modexp_pair(c, e, N, R, Ninv)
        assert got == pow(c, e, N * N), (bits, Lbits)
    print(bits, Lbits, "ok")
