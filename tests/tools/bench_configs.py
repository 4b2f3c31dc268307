#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configurations that are not the bench.py headline:
cfg3 (5 parties, t=2), cfg4 (key_length 4096: partial decrypt and r^N), cfg5 (biprimality batch
sweep), plus the reference-shaped 2048-bit key.  One JSON line per measurement; device-resident
timing with CUDA events, results spot-checked against CPython pow.

    python tests/tools/bench_configs.py [--quick]
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
import torch  # noqa: E402

import protocols.distributed_keygen_b200 as eng  # noqa: E402
from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints  # noqa: E402
from oracle import keys as okeys  # noqa: E402


def canonical_macs(ebits: int, limbs: int) -> float:
    return float(ebits + (ebits + 4) // 5 + 32) * float(2 * limbs * limbs + limbs)


def rand_rows(count, limbs, top_bits, seed):
    rng = np.random.default_rng(seed)
    arr = rng.integers(0, 2**32, size=(count, limbs), dtype=np.uint32)
    arr[:, -1] &= np.uint32((1 << (top_bits - 1)) - 1)
    arr[:, 0] |= np.uint32(1)
    return arr


def time_modexp(name, modulus, exponent, waves=2, check=True, root=None):
    ctx = eng.ModexpContext(modulus, exponent, root=root)
    info = ctx.info()
    B = waves * info["ctas"] * info["warps_per_cta"] * 32
    L = ctx.limbs
    host = rand_rows(B, L, modulus.bit_length() - 32 * (L - 1), 7)
    d_in = torch.from_numpy(host.view(np.int32)).cuda()
    d_out = torch.empty_like(d_in)
    d_st = torch.empty(B, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ctx.modexp_device(d_in.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), min(B, 4096), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.modexp_device(d_in.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), B, s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ok = None
    if check:
        out = d_out.cpu().numpy().view(np.uint32)
        ok = all(
            limbs_to_ints(out[i : i + 1])[0] == pow(limbs_to_ints(host[i : i + 1])[0], exponent, modulus)
            for i in (0, B // 2, B - 1)
        )
    ebits = abs(exponent).bit_length()
    line = {
        "config": name, "op": "modexp", "count": B, "ms": ms, "per_s": B / ms * 1e3,
        "modulus_bits": modulus.bit_length(), "exponent_bits": ebits * (1 if exponent >= 0 else -1),
        "pair_arithmetic": bool(info.get("pair_arithmetic")), "kernel_shape": info, "canonical_Tmac_per_s": B * canonical_macs(ebits, L) / ms / 1e9, "ok": ok,
    }
    print(json.dumps(line), flush=True)
    ctx.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    with open(os.path.join(ROOT, "tests", "golden", "dealer_vectors.json")) as fh:
        dv = json.load(fh)["keys"]

    # cfg2 realistic (129-limb N^2) and cfg3 (5 parties t=2): one partial decryption per party
    for name in ["cfg2_k2048_p3_t1_real", "cfg3_k2048_p5_t2_exact"]:
        dk = okeys.dealer_key_from_json(dv[name]["key"])
        for pid in sorted(dk.keys)[: (2 if args.quick else 2 * dk.t + 1)]:
            k = dk.keys[pid]
            time_modexp(f"{name}/party{pid}", k.n_square, k.partial_decrypt_exponent(), waves=1, root=k.n)
    # cfg3 combine-only rate (5 partials)
    dk = okeys.dealer_key_from_json(dv["cfg3_k2048_p5_t2_exact"]["key"])
    key = dk.keys[1]
    shares = 2 * dk.t + 1
    comb = eng.CombineContext(dk.n, key.theta_inv, shares)
    B = 1 << 16
    parts = torch.from_numpy(rand_rows(shares * B, comb.n2_limbs, (dk.n * dk.n).bit_length() - 32 * (comb.n2_limbs - 1), 3).view(np.int32)).cuda()
    d_out = torch.empty((B, comb.n_limbs), dtype=torch.int32, device="cuda")
    d_st = torch.empty(B, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    comb.combine_device(parts.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), B, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); comb.combine_device(parts.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), B, s); e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"config": "cfg3 combine-only (5 partials, device-resident)", "op": "combine", "count": B,
                      "ms": e0.elapsed_time(e1), "per_s": B / e0.elapsed_time(e1) * 1e3}), flush=True)

    # cfg4: key_length 4096
    dk = okeys.dealer_key_from_json(dv["cfg4_k4096_p3_t1_exact"]["key"])
    k = dk.keys[1]
    time_modexp("cfg4_k4096/partial_decrypt", k.n_square, k.partial_decrypt_exponent(), waves=1, root=k.n)
    time_modexp("cfg4_k4096/partial_decrypt (direct kernel)", k.n_square, k.partial_decrypt_exponent(), waves=1)
    time_modexp("cfg4_k4096/r^N", k.n_square, dk.n, waves=1, root=dk.n)
    dk2 = okeys.dealer_key_from_json(dv["cfg2_k2048_p3_t1_exact"]["key"])
    time_modexp("cfg2_k2048/r^N", dk2.n * dk2.n, dk2.n, waves=1, root=dk2.n)

    # cfg5: biprimality batch sweep (party 1 exponents ~2046 bits, 40 bases per candidate)
    rng = random.Random(5)
    pl = 1024
    sizes = [1, 8, 64, 512, 2048] if args.quick else [1, 4, 16, 64, 256, 1024, 4096, 16384]
    for C in sizes:
        moduli, exps = [], []
        for _ in range(C):
            p_sh = [okeys.prime_candidate_share(i + 1, pl, rng) for i in range(3)]
            q_sh = [okeys.prime_candidate_share(i + 1, pl, rng) for i in range(3)]
            n = sum(p_sh) * sum(q_sh)
            moduli.append(n)
            exps.append((n - p_sh[0] - q_sh[0] + 1) // 4)
        L = 65
        m_arr = ints_to_limbs(moduli, L)
        e_arr = ints_to_limbs(exps, 65)
        bases = rand_rows(C * 40, L, 2050 - 32 * 64, 11).reshape(C, 40, L)
        bases[:, :, -1] = 0  # < 2^2048 <= N
        eng.modexp_grouped_limbs(m_arr[:1], e_arr[:1], bases[:1])
        t0 = time.perf_counter()
        out = eng.modexp_grouped_limbs(m_arr, e_arr, bases)
        secs = time.perf_counter() - t0
        b0 = limbs_to_ints(bases[C - 1, 39:40])[0]
        ok = limbs_to_ints(out[C - 1, 39:40])[0] == pow(b0, exps[-1], moduli[-1])
        print(json.dumps({"config": "cfg5 biprime v_1 batch (host buffers, end to end)", "op": "modexp_grouped",
                          "candidates": C, "modexps": C * 40, "ms": secs * 1e3, "modexps_per_s": C * 40 / secs,
                          "candidates_per_s": C / secs, "ok": ok}), flush=True)


if __name__ == "__main__":
    main()
