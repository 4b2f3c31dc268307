import sys, random
sys.path.insert(0, '.')
import protocols.distributed_keygen_b200 as eng
from protocols.distributed_keygen_b200 import distributed_keygen as dkg
rng = random.Random(1)
for bits in (67, 134, 515):
    n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
    for e in (rng.getrandbits(40), -rng.getrandbits(40)):
        ctx = eng.ModexpContext(n, e)
        import math
        bases = [b for b in (rng.randrange(1, n) for _ in range(90)) if math.gcd(b, n) == 1][:70]
        assert ctx.modexp(bases) == [pow(b, e, n) for b in bases]
        ctx.close()
p, q = 1000003, 999983
n = p * q
enc = eng.EncryptContext(n)
ms = [rng.randrange(n) for _ in range(40)]; rs = [rng.randrange(1, n) for _ in range(40)]
cts = enc.encrypt(ms, rs)
assert cts == [((1 + m * n) * pow(r, n, n * n)) % (n * n) for m, r in zip(ms, rs)]
comb = eng.CombineContext(n, pow(5, -1, n), 2)
moduli = [rng.getrandbits(130) | 1 | (1 << 129) for _ in range(5)]
exps = [rng.getrandbits(100) for _ in moduli]
gs = [[rng.randrange(m) for _ in range(16)] for m in moduli]
got = dkg.biprime_test_v_calculation_batch([(g, m, 2 * e, 2 * e) for g, m, e in zip(gs, moduli, exps)], 2, 4)
print("sanitize workload ok", len(got))
print(eng.small_prime_sieve(moduli, [3, 5, 7, 11]), eng.jacobi_batch(moduli[:2], gs[:2])[0][:4])
# pair arithmetic (N^2 with root), both window policies, negative exponents, wider roots
import os
for sliding in ("0", "1"):
    os.environ["DKG_SLIDING_WINDOW"] = sliding
    for bits in (40, 130, 515):
        p_ = rng.getrandbits(bits) | 1 | (1 << (bits - 1)); q_ = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        root = p_ * q_; m2 = root * root
        for e in (rng.getrandbits(90), -rng.getrandbits(90), (1 << 70) + 1):
            ctx = eng.ModexpContext(m2, e, root=root)
            bases = [b for b in (rng.randrange(1, m2) for _ in range(50)) if math.gcd(b, m2) == 1][:40]
            want = [pow(b, e, m2) for b in bases]
            assert ctx.modexp(bases) == want, (bits, e)
            ctx.close()
print("pair arithmetic sanitize workload ok")
