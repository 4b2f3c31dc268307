"""One cooperative-kernel call for ncu: B ciphertexts, 2048-bit reference-shaped key."""
import os, random, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import protocols.distributed_keygen_b200 as eng
from protocols.distributed_keygen_b200 import _native
from protocols.distributed_keygen_b200.limbs import ints_to_limbs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
mode = sys.argv[2] if len(sys.argv) > 2 else "nsq"
rng = random.Random(1)
if mode == "nsq":
    n = ((1 << 1025) + 643) * ((1 << 1025) + 1113)
    e = rng.getrandbits(4196) | (1 << 4195)
    _native.config_set("coop_max", 1 << 30)
    ctx = eng.ModexpContext(n * n, e, root=n)
    vals = []
    while len(vals) < B:
        v = rng.randrange(1, n * n)
        if math.gcd(v, n) == 1:
            vals.append(v)
    rows = ints_to_limbs(vals, (2 * n.bit_length() + 31) // 32)
    ctx.modexp_limbs(rows)
else:
    _native.config_set("coop_grouped_max", 1 << 30)
    C = max(1, B // 40)
    moduli = [rng.getrandbits(2051) | 1 | (1 << 2050) for _ in range(C)]
    exps = [rng.getrandbits(2046) for _ in range(C)]
    gs = [[rng.randrange(m) for _ in range(40)] for m in moduli]
    eng.modexp_grouped(moduli, exps, gs)
print("done")
