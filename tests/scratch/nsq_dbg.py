import sys, random, math
sys.path.insert(0, '.')
import protocols.distributed_keygen_b200 as eng
rng = random.Random(404)
bits = 9
p = rng.getrandbits(bits // 2) | 1 | (1 << (bits // 2 - 1)); q = rng.getrandbits(bits - bits // 2) | 1 | (1 << (bits - bits // 2 - 1))
n, n2 = p * q, (p * q) ** 2
print('p,q,n', p, q, n)
for e in [0, 1, 3, rng.getrandbits(bits + 40), -rng.getrandbits(bits + 40)]:
    fast = eng.ModexpContext(n2, e, root=n)
    bases = [0, 1, n2 - 1, n, n + 1, n - 1] + [rng.randrange(n2) for _ in range(70)]
    if e < 0: bases = [b for b in bases if math.gcd(b, n) == 1]
    got = fast.modexp(bases); want = [pow(b, e, n2) for b in bases]
    bad = [(i, bases[i], bases[i] % n, math.gcd(bases[i], n), got[i], want[i]) for i in range(len(bases)) if got[i] != want[i]]
    print(e, bad)
    for (i, b, *_r) in bad:
        for ee in [1, 2, 3, 4, 6, 8, 12, 64]:
            c2 = eng.ModexpContext(n2, ee, root=n)
            print('   b', b, 'e', ee, c2.modexp([b]), pow(b, ee, n2), 'single; in-batch:', c2.modexp(bases)[i])
            c2.close()
    fast.close()
