import random, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import protocols.distributed_keygen_b200 as eng
from protocols.distributed_keygen_b200 import _native
from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints
rng = random.Random(23)
p = (1 << 1025) + 643
q = (1 << 1025) + 1113  # odd; units are filtered below
n = p * q; n2 = n * n
for B in (100, 148, 149, 300, 600):
    import math
    vals = []
    while len(vals) < B:
        v = rng.randrange(1, n2)
        if math.gcd(v, n) == 1: vals.append(v)
    rows = ints_to_limbs(vals, (n2.bit_length() + 31) // 32)
    for sign in (1, -1):
        e = sign * (rng.getrandbits(300) | 1)
        want = [pow(v, e, n2) for v in vals]
        for label, limit in (("coop", 1 << 30), ("wave", 0)):
            _native.config_set("coop_max", limit)
            ctx = eng.ModexpContext(n2, e, root=n)
            out, st = ctx.modexp_limbs(rows)
            got = limbs_to_ints(out)
            bad = [i for i in range(B) if got[i] != want[i]]
            print(B, sign, label, "status nonzero:", int(st.astype(bool).sum()), "wrong:", len(bad), bad[:10], flush=True)
            ctx.close()
