#!/usr/bin/env python
"""A/B of the grouped (per-candidate modulus) kernel shapes on the biprimality-test batch (BASELINE
config 5): 2050-bit candidates, 40 bases each, host buffers end to end, spot-checked against CPython pow.

    DKG_GROUPED_SHAPE=14,5 python tests/tools/grouped_ab.py [candidates]
"""
from __future__ import annotations

import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402
import protocols.distributed_keygen_b200 as eng  # noqa: E402
from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints  # noqa: E402


def main() -> None:
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    rng = random.Random(5)
    base_c = min(C, 64)
    moduli, exps = [], []
    for _ in range(base_c):
        ps = [(rng.getrandbits(1024) | (1 << 1023) | 3) if i == 0 else ((rng.getrandbits(1024) | (1 << 1023)) & ~3) for i in range(3)]
        qs = [(rng.getrandbits(1024) | (1 << 1023) | 3) if i == 0 else ((rng.getrandbits(1024) | (1 << 1023)) & ~3) for i in range(3)]
        n = sum(ps) * sum(qs)
        moduli.append(n)
        exps.append((n - ps[0] - qs[0] + 1) // 4)
    L = 65
    reps = -(-C // base_c)
    m_arr = np.tile(ints_to_limbs(moduli, L), (reps, 1))[:C]
    e_arr = np.tile(ints_to_limbs(exps, L), (reps, 1))[:C]
    bases = bench.random_units(C * 40, 1 << 2049, L, 11).reshape(C, 40, L)
    eng.modexp_grouped_limbs(m_arr[:1], e_arr[:1], bases[:1])
    eng.modexp_grouped_limbs(m_arr, e_arr, bases)
    t0 = time.perf_counter()
    res = eng.modexp_grouped_limbs(m_arr, e_arr, bases)
    secs = time.perf_counter() - t0
    for c, k in ((0, 0), (C // 2, 17), (C - 1, 39)):
        b0 = limbs_to_ints(bases[c, k:k + 1])[0]
        assert limbs_to_ints(res[c, k:k + 1])[0] == pow(b0, exps[c % base_c], moduli[c % base_c])
    print(json.dumps({"config": "cfg5 grouped modexp", "shape": os.environ.get("DKG_GROUPED_SHAPE", "default"), "candidates": C,
                      "modexps": 40 * C, "ms": round(secs * 1e3, 2), "modexps_per_s": round(40 * C / secs, 1)}), flush=True)


if __name__ == "__main__":
    main()
