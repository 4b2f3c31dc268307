import sys, random, math
sys.path.insert(0, '.')
import protocols.distributed_keygen_b200 as eng
rng = random.Random(3)
for bits in [20, 61, 67, 130, 257, 515, 1030, 2048, 2051]:
    p = rng.getrandbits(bits // 2) | 1 | (1 << (bits // 2 - 1)); q = rng.getrandbits(bits - bits // 2) | 1 | (1 << (bits - bits // 2 - 1))
    n = p * q; n2 = n * n
    for e in [0, 1, 2, 65537, rng.getrandbits(bits + 50), -rng.getrandbits(bits + 50)]:
        ctx = eng.ModexpContext(n2, e, root=n)
        bases = [1, 2, n2 - 1, n, n + 1] + [rng.randrange(n2) for _ in range(60)]
        if e < 0: bases = [b for b in bases if math.gcd(b, n) == 1]
        got = ctx.modexp(bases)
        want = [pow(b, e, n2) for b in bases]
        bad = [i for i in range(len(bases)) if got[i] != want[i]]
        print(bits, 'e', e.bit_length() * (1 if e >= 0 else -1), ctx.info(), 'OK' if not bad else ('BAD %d/%d first %d' % (len(bad), len(bases), bad[0])), flush=True)
        ctx.close()
