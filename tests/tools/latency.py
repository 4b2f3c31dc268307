#!/usr/bin/env python
"""Latency of one call as a function of the batch size, through the host-buffer C ABI (copies
included): partial decryption at 2048-bit N (positive and negative exponent) and one
compute_modulus round's v calculation, cooperative kernels vs thread-per-operand kernels vs GMP
``mpz_powm`` on all host cores.  One JSON line per point.

    python tests/tools/latency.py [--quick] [--key-bits 2048]
"""
from __future__ import annotations

import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import protocols.distributed_keygen_b200 as eng  # noqa: E402
from protocols.distributed_keygen_b200 import _native  # noqa: E402
from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints  # noqa: E402


def timed(fn, reps):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--key-bits", type=int, default=2048)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    from oracle import gmp as ogmp   # CPU baseline = the checker's GMP harness

    cores = os.cpu_count() or 1
    rng = random.Random(99)
    import math

    half = args.key_bits // 2 + 1          # reference-shaped: N is key_length + 2..4 bits
    p = rng.getrandbits(half) | 1 | (1 << (half - 1))
    q = rng.getrandbits(half) | 1 | (1 << (half - 1))
    n = p * q                              # not a biprime (cost is identical): bases are drawn as units
    n2 = n * n
    ebits = 2 * args.key_bits + 100
    sizes = [1, 32, 1024] if args.quick else [1, 8, 32, 128, 512, 1024, 4096, 16384]
    for sign in (1, -1):
        e = sign * (rng.getrandbits(ebits) | (1 << (ebits - 1)))
        for B in sizes:
            vals = []
            while len(vals) < min(B, 64):
                v = rng.randrange(1, n2)
                if math.gcd(v, n) == 1:
                    vals.append(v)
            rows = ints_to_limbs((vals * (B // len(vals) + 1))[:B], (n2.bit_length() + 31) // 32)
            line = {"op": "partial_decrypt", "key_bits": args.key_bits, "exp_sign": sign, "batch": B}
            for label, limit in (("coop_ms", 1 << 30), ("wave_ms", 0)):
                if label == "coop_ms" and B > 16384:
                    continue
                _native.config_set("coop_max", limit)
                ctx = eng.ModexpContext(n2, e, root=n)
                out = [None]

                def call():
                    out[0] = ctx.modexp_limbs(rows)

                line[label] = round(timed(call, 2 if B >= 4096 else 3), 3)
                got = limbs_to_ints(out[0][0][:2])
                assert got == [pow(v, e, n2) for v in vals[:2]], label
                ctx.close()
            if not args.no_cpu:
                sample = rows[: min(B, 4 * cores)]
                L = rows.shape[1]
                _, secs = ogmp.powm_batch_threads(sample, ints_to_limbs([n2], L)[0], ints_to_limbs([abs(e)], (ebits + 31) // 32)[0],
                                                  sign < 0, cores)
                waves = -(-B // cores)
                per_wave = secs / -(-len(sample) // cores)
                line["gmp_ms"] = round(per_wave * waves * 1e3, 3)
                line["gmp_cores"] = cores
            print(json.dumps(line), flush=True)
    _native.config_set("coop_max", 8192)
    # biprimality: C candidates x 40 bases, party 1 exponent
    hb = args.key_bits // 2
    cands = [1, 16, 64] if args.quick else [1, 2, 4, 8, 16, 32, 64, 128, 256, 512]
    for C in cands:
        moduli, exps, gs = [], [], []
        for _ in range(C):
            ps = [rng.getrandbits(hb) | (1 << (hb - 1)) | 3 if i == 0 else (rng.getrandbits(hb) | (1 << (hb - 1))) & ~3 for i in range(3)]
            qs = [rng.getrandbits(hb) | (1 << (hb - 1)) | 3 if i == 0 else (rng.getrandbits(hb) | (1 << (hb - 1))) & ~3 for i in range(3)]
            N = sum(ps) * sum(qs)
            moduli.append(N)
            exps.append((N - ps[0] - qs[0] + 1) // 4)
            gs.append([rng.randrange(N) for _ in range(40)])
        line = {"op": "biprime_v", "key_bits": args.key_bits, "candidates": C, "modexps": 40 * C}
        for label, limit in (("coop_ms", 1 << 30), ("wave_ms", 0)):
            _native.config_set("coop_grouped_max", limit)
            out = [None]

            def call():
                out[0] = eng.modexp_grouped(moduli, exps, gs)

            line[label] = round(timed(call, 2), 3)
            assert out[0][0][:2] == [pow(g, exps[0], moduli[0]) for g in gs[0][:2]], label
        if not args.no_cpu:
            L = (max(moduli).bit_length() + 31) // 32
            Le = (max(exps).bit_length() + 31) // 32
            _, secs = ogmp.powm_grouped_threads(ints_to_limbs([g for row in gs for g in row], L).reshape(C, 40, L),
                                                ints_to_limbs(moduli, L), ints_to_limbs(exps, Le), cores)
            line["gmp_ms"] = round(secs * 1e3, 3)
            line["gmp_cores"] = cores
        print(json.dumps(line), flush=True)
    _native.config_set("coop_grouped_max", 16384)


if __name__ == "__main__":
    main()
