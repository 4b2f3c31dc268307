#!/usr/bin/env python
"""Workload for compute-sanitizer (memcheck + racecheck on the shared-memory
carry / window code).  Small shapes, every kernel family, both routes (cooperative warp-per-operand
and thread-per-operand), every result checked against CPython ``pow``:

    compute-sanitizer --tool memcheck  python tests/tools/sanitize.py
    compute-sanitizer --tool racecheck python tests/tools/sanitize.py
"""
import math
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import protocols.distributed_keygen_b200 as eng  # noqa: E402
from protocols.distributed_keygen_b200 import _native  # noqa: E402
from protocols.distributed_keygen_b200 import distributed_keygen as dkg  # noqa: E402

rng = random.Random(1)
quick = "--quick" in sys.argv
for coop in (1, 0):
    _native.config_set("coop_max", 1 << 20 if coop else 0)
    _native.config_set("coop_grouped_max", 1 << 20 if coop else 0)
    # generic odd moduli (direct kernel + batched inversion)
    for bits in (67, 134) if quick else (67, 134, 515):
        n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        for e in (rng.getrandbits(40), -rng.getrandbits(40)):
            ctx = eng.ModexpContext(n, e)
            bases = [b for b in (rng.randrange(1, n) for _ in range(90)) if math.gcd(b, n) == 1][:70]
            assert ctx.modexp(bases) == [pow(b, e, n) for b in bases]
            ctx.close()
    # pair arithmetic (N^2 with root), both window policies, negative exponents
    for sliding in ("0", "1"):
        os.environ["DKG_SLIDING_WINDOW"] = sliding
        for bits in (40, 130) if quick else (40, 130, 515):
            p_ = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
            q_ = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
            root = p_ * q_
            m2 = root * root
            for e in (rng.getrandbits(90), -rng.getrandbits(90), (1 << 70) + 1):
                ctx = eng.ModexpContext(m2, e, root=root)
                bases = [b for b in (rng.randrange(1, m2) for _ in range(50)) if math.gcd(b, m2) == 1][:40]
                bases += [root * 3, 0] if e > 0 else []
                assert ctx.modexp(bases) == [pow(b, e, m2) for b in bases], (bits, e, coop)
                ctx.close()
    os.environ["DKG_SLIDING_WINDOW"] = "0"
    # encryption, combination, biprimality round
    p, q = 1000003, 999983
    n = p * q
    enc = eng.EncryptContext(n)
    ms = [rng.randrange(n) for _ in range(40)]
    rs = [rng.randrange(1, n) for _ in range(40)]
    assert enc.encrypt(ms, rs) == [((1 + m * n) * pow(r, n, n * n)) % (n * n) for m, r in zip(ms, rs)]
    enc.close()
    moduli = [rng.getrandbits(130) | 1 | (1 << 129) for _ in range(5)]
    exps = [rng.getrandbits(100) for _ in moduli]
    gs = [[rng.randrange(m) for _ in range(16)] for m in moduli]
    got = dkg.biprime_test_v_calculation_batch([(g, m, 2 * e, 2 * e) for g, m, e in zip(gs, moduli, exps)], 2, 4)
    assert all(len(v) <= 4 for v in got)
    assert eng.modexp_grouped(moduli, exps, gs) == [[pow(g, e, m) for g in row] for m, e, row in zip(moduli, exps, gs)]
    eng.small_prime_sieve(moduli, [3, 5, 7, 11])
    eng.jacobi_batch(moduli[:2], gs[:2])
    if not coop:
        # round-2 kernels, thread-per-operand route: 13-limb blocks in 14-limb slots (2048-bit N), b
        # component in global scratch (4096-bit N), the shared squaring chain with a negative party
        for bits in (2049, 4096):
            root = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
            m2 = root * root
            for e in (rng.getrandbits(45), -rng.getrandbits(45)):
                ctx = eng.ModexpContext(m2, e, root=root)
                bases = [b for b in (rng.randrange(1, m2) for _ in range(60)) if math.gcd(b, root) == 1][:37]
                assert ctx.modexp(bases) == [pow(b, e, m2) for b in bases], (bits, e)
                ctx.close()
        big = [rng.getrandbits(2050) | 1 | (1 << 2049) for _ in range(2)]
        bexp = [rng.getrandbits(50) for _ in big]
        bgs = [[rng.randrange(m) for _ in range(5)] for m in big]
        assert eng.modexp_grouped(big, bexp, bgs) == [[pow(g, e, m) for g in row] for m, e, row in zip(big, bexp, bgs)]
        for bits in (130, 2049):
            root = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
            m2 = root * root
            exps = {1: rng.getrandbits(60), 2: -rng.getrandbits(61), 3: rng.getrandbits(59)}
            tctx = eng.ThresholdContext(root, 1, exps, [0])
            cs = [b for b in (rng.randrange(1, m2) for _ in range(80)) if math.gcd(b, root) == 1][:45]
            from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints
            parts, status = tctx.partials_limbs(ints_to_limbs(cs, tctx.n2_limbs))
            tctx.close()
            assert not status.any()
            for pid, e in exps.items():
                assert limbs_to_ints(parts[pid - 1]) == [pow(c, e, m2) for c in cs], (bits, pid)
    print("sanitize workload ok, cooperative kernels" if coop else "sanitize workload ok, thread-per-operand kernels", flush=True)
