"""
Lane-level Python model of the cooperative (warp-per-operand) arithmetic in csrc/dkg_coop.cuh.

A number of nb blocks of K 32-bit limbs lives with block p on lane p of a 32-lane warp.  The model
mirrors, step by step, what each lane does -- the diagonal plan, the chunk combination, the carry
resolution with generate/propagate lookahead (what the kernel does with __shfl_up_sync and
__ballot_sync), the three-phase Montgomery product, the pair arithmetic on top, the almost-inverse
(Kaliski) and the Newton iteration for -N^-1 mod R -- on Python integers per lane, and asserts the
bounds the CUDA code relies on.  tests/test_coop_model.py runs it against CPython pow; the host
planner in dkg_engine.cu is a transcription of ``make_plan``.
"""
from __future__ import annotations

LANES = 32


def blocks_of(x: int, nb: int, K: int) -> list[int]:
    W = 1 << (32 * K)
    return [(x >> (32 * K * i)) & (W - 1) for i in range(nb)] + [0] * (LANES - nb)


def value_of(blocks: list[int], K: int, count: int | None = None) -> int:
    return sum(b << (32 * K * i) for i, b in enumerate(blocks[: count if count is not None else len(blocks)]))


# ---- plan: which lane computes which chunk of which block anti-diagonal ---------------------------
def make_plan(nb: int, ndiag: int, low: bool = False, square: bool = False):
    """Diagonal d (0 <= d < ndiag) of an nb x nb block product has tiles (i, d - i),
    max(0, d-nb+1) <= i <= min(d, nb-1).  Lane d is the primary of diagonal d; spare lanes take a
    second/third/fourth chunk of the longest diagonals.  Returns per-lane (d, i0, i1) and per
    primary lane the list of partner lanes (<= 3)."""
    assert ndiag <= LANES
    lo = [max(0, d - nb + 1) for d in range(ndiag)]
    hi = [min(d, nb - 1) + 1 for d in range(ndiag)]
    ln = [hi[d] - lo[d] for d in range(ndiag)]
    # smallest chunk length c such that sum_d ceil(len_d / c) lanes suffice (at most 4 chunks each)
    c = 1
    while sum(-(-n // c) for n in ln) > LANES or max(-(-n // c) for n in ln) > 4:
        c += 1
    chunks = [-(-n // c) for n in ln]
    lanes = [(0, 0, 0)] * LANES
    partners: list[list[int]] = [[] for _ in range(LANES)]
    nxt = ndiag
    for d in range(ndiag):
        per = -(-ln[d] // chunks[d])
        for c in range(chunks[d]):
            a = lo[d] + c * per
            b = min(lo[d] + (c + 1) * per, hi[d])
            if c == 0:
                lanes[d] = (d, a, b)
            else:
                lanes[nxt] = (d, a, max(a, b))
                partners[d].append(nxt)
                nxt += 1
    return lanes, partners


# ---- one product phase -----------------------------------------------------------------------------
def product(plan, X: list[int], Y: list[int], K: int, X2=None, Y2=None) -> list[int]:
    """Per-lane diagonal sums E_d (on the primary lane d) of X*Y (+ X2*Y2)."""
    lanes, partners = plan
    part = [0] * LANES
    for lane, (d, i0, i1) in enumerate(lanes):
        acc = 0
        for i in range(i0, i1):
            acc += X[i] * Y[d - i]
            if X2 is not None:
                acc += X2[i] * Y2[d - i]
        assert acc < 1 << (32 * (2 * K + 2))
        part[lane] = acc
    e = list(part)
    for d in range(LANES):
        for src in partners[d]:
            e[d] += part[src]                      # shfl from the partner lane + add
        assert e[d] < 1 << (32 * (2 * K + 2))
    # only primaries keep their value
    ndiag = max(d for (d, _, _) in lanes) + 1
    return [e[d] if d < ndiag else 0 for d in range(LANES)]


def lookahead(g_bits: int, p_bits: int) -> int:
    """carry into each lane from generate / propagate masks (what the kernel does with two ballots)."""
    a, b = g_bits | p_bits, g_bits
    return (((a + b) ^ a ^ b)) & 0xFFFFFFFF


def resolve(e: list[int], K: int, addend: list[int] | None = None, carry_in: int = 0) -> tuple[list[int], int]:
    """sum_d e[d] W^d (+ sum addend[d] W^d) -> blocks on lanes 0..31.  Returns (blocks, carry out of
    lane 31)."""
    W = 1 << (32 * K)
    mid = [0] + [(e[d] >> (32 * K)) & (W - 1) for d in range(LANES - 1)]          # shfl_up 1, K limbs
    high = [0, 0] + [e[d] >> (64 * K) for d in range(LANES - 2)]                  # shfl_up 2, 2 limbs
    s, c = [0] * LANES, [0] * LANES
    for p in range(LANES):
        t = (e[p] & (W - 1)) + mid[p] + high[p] + (addend[p] if addend else 0) + (carry_in if p == 0 else 0)
        s[p], c[p] = t & (W - 1), t >> (32 * K)
        assert c[p] <= 3
    cin = [0] + c[:-1]
    g = 0
    for p in range(LANES):
        t = s[p] + cin[p]
        s[p] = t & (W - 1)
        if t >> (32 * K):
            g |= 1 << p
    pm = sum(1 << p for p in range(LANES) if s[p] == W - 1)
    assert g & pm == 0
    ci = lookahead(g, pm)
    for p in range(LANES):
        if (ci >> p) & 1:
            s[p] = (s[p] + 1) & (W - 1)
    top = c[LANES - 1] + ((g >> (LANES - 1)) & 1)
    return s, top


def add(x: list[int], y: list[int], K: int, carry_in: int = 0) -> tuple[list[int], int]:
    W = 1 << (32 * K)
    s, g = [0] * LANES, 0
    for p in range(LANES):
        t = x[p] + y[p] + (carry_in if p == 0 else 0)
        s[p] = t & (W - 1)
        if t >> (32 * K):
            g |= 1 << p
    pm = sum(1 << p for p in range(LANES) if s[p] == W - 1)
    pm &= ~g
    ci = lookahead(g, pm)
    for p in range(LANES):
        if (ci >> p) & 1:
            s[p] = (s[p] + 1) & (W - 1)
    a, b = g | pm, g
    return s, ((a + b) >> LANES) & 1


def sub(x: list[int], y: list[int], K: int) -> tuple[list[int], int]:
    """x - y (mod W^32) and the borrow out."""
    W = 1 << (32 * K)
    return add(x, [(W - 1) ^ v for v in y], K, 1)[0], 1 - add(x, [(W - 1) ^ v for v in y], K, 1)[1]


class Mont:
    """Montgomery context of an odd modulus N with R = W^nb >= 4N (values stay below 2N)."""

    def __init__(self, N: int, nb: int, K: int) -> None:
        self.N, self.nb, self.K = N, nb, K
        self.R = 1 << (32 * K * nb)
        assert N % 2 == 1 and 4 * N <= self.R and 2 * nb <= LANES
        self.NI = (-pow(N, -1, self.R)) % self.R
        self.plan_full = make_plan(nb, 2 * nb - 1)
        self.plan_low = make_plan(nb, nb)
        self.Nb = blocks_of(N, nb, K)
        self.NIb = blocks_of(self.NI, nb, K)

    def redc_blocks(self, T: list[int]) -> tuple[list[int], list[int]]:
        """T: 2nb resolved blocks.  Returns ((T + qN)/R as nb blocks on lanes 0.., q blocks)."""
        nb, K = self.nb, self.K
        Tlo = T[:nb] + [0] * (LANES - nb)
        e = product(self.plan_low, Tlo, self.NIb, K)
        q, _ = resolve(e, K)
        q = q[:nb] + [0] * (LANES - nb)
        e = product(self.plan_full, q, self.Nb, K)
        u, top = resolve(e, K, addend=T)
        assert top == 0 and all(v == 0 for v in u[:nb]) and all(v == 0 for v in u[2 * nb :])
        return u[nb : 2 * nb] + [0] * (LANES - nb), q

    def mul(self, X: list[int], Y: list[int], X2=None, Y2=None) -> tuple[list[int], list[int]]:
        e = product(self.plan_full, X, Y, self.K, X2, Y2)
        T, top = resolve(e, self.K)
        assert top == 0
        return self.redc_blocks(T)


def test_selfcheck() -> None:
    import random

    rng = random.Random(1)
    for K, nb, bits in [(6, 11, 2052), (6, 11, 2048), (6, 1, 67), (6, 3, 515), (12, 11, 4100), (6, 16, 3000), (4, 5, 515)]:
        for _ in range(3):
            N = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
            m = Mont(N, nb, K)
            x, y, x2, y2 = (rng.randrange(2 * N) for _ in range(4))
            r, q = m.mul(blocks_of(x, nb, K), blocks_of(y, nb, K))
            v = value_of(r, K)
            assert v < 2 * N and (v * m.R - x * y) % N == 0
            assert v * m.R == x * y + value_of(q, K) * N
            r, q = m.mul(blocks_of(x, nb, K), blocks_of(y, nb, K), blocks_of(x2, nb, K), blocks_of(y2 % N, nb, K))
            assert value_of(r, K) * m.R == x * y + x2 * (y2 % N) + value_of(q, K) * N
            a, b = rng.randrange(m.R), rng.randrange(m.R)
            s, c = add(blocks_of(a, LANES, K)[:LANES], blocks_of(b, LANES, K)[:LANES], K)
        # plan sanity: every tile exactly once
        for ndiag in (2 * nb - 1, nb):
            lanes, partners = make_plan(nb, ndiag)
            seen = set()
            for d, i0, i1 in lanes:
                for i in range(i0, i1):
                    assert (i, d - i) not in seen
                    seen.add((i, d - i))
            want = {(i, j) for i in range(nb) for j in range(nb) if i + j < ndiag}
            assert seen == want, (nb, ndiag)
            print(K, nb, ndiag, "max tiles per lane", max(i1 - i0 for _, i0, i1 in lanes), "rounds", max(len(p) for p in partners))


if __name__ == "__main__":
    test_selfcheck()
    print("ok")
