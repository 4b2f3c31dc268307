#!/usr/bin/env python
"""A/B of the wide-key pair kernels (key_length 4096, BASELINE config 4): b component in global
scratch (default, 12 warps/SM) against both components in shared memory (DKG_NSQ_BG=0, 6 warps/SM).
One wave per launch, device-resident, CUDA events; spot-checked against CPython pow.

    python tests/tools/wide_ab.py [--short]      # --short: 512-bit exponent instead of the key's 8200 bits
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import protocols.distributed_keygen_b200 as eng  # noqa: E402
from protocols.distributed_keygen_b200.limbs import limbs_to_ints  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--short", action="store_true")
    ap.add_argument("--key", default="cfg4_k4096_p3_t1_exact")
    args = ap.parse_args()
    with open(os.path.join(ROOT, "tests", "golden", "dealer_vectors.json")) as fh:
        dv = json.load(fh)["keys"]
    dk = bench.KeyData(dv[args.key]["key"])
    keys = bench.gpu_keys(dk)
    n2 = dk.n * dk.n
    stream = torch.cuda.current_stream().cuda_stream
    for label, exponent in (("partial decrypt party 1", keys[1].partial_decrypt_exponent()), ("r^N", dk.n)):
        if args.short:
            exponent = exponent >> (abs(exponent).bit_length() - 512) if exponent > 0 else -((-exponent) >> ((-exponent).bit_length() - 512))
        for bg in ("1", "0"):
            os.environ["DKG_NSQ_BG"] = bg
            ctx = eng.ModexpContext(n2, exponent, root=dk.n)
            info = ctx.info()
            B = info["ctas"] * info["pair_warps_per_cta"] * 32
            host = bench.random_units(B, n2, ctx.limbs, 7)
            d_in = torch.from_numpy(host.view(np.int32)).cuda()
            d_out = torch.empty_like(d_in)
            d_st = torch.empty(B, dtype=torch.uint8, device="cuda")
            ctx.modexp_device(d_in.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), B, stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.modexp_device(d_in.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), B, stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            got = limbs_to_ints(d_out[B - 1:].cpu().numpy().view(np.uint32))[0]
            assert got == pow(limbs_to_ints(host[B - 1:])[0], exponent, n2), (label, bg)
            print(json.dumps({"config": f"{args.key} {label}", "b_in_global_scratch": bg == "1", "warps_per_cta": info["pair_warps_per_cta"],
                              "shape": [info["pair_K"], info["pair_M"]], "count": B, "exponent_bits": abs(exponent).bit_length(),
                              "ms": round(ms, 2), "per_s": round(B / ms * 1e3, 1)}), flush=True)
            ctx.close()


if __name__ == "__main__":
    main()
