#!/usr/bin/env python
"""A/B of the two ways a negative exponent's inversion is done on the thread-per-ciphertext route: in
the pair kernel per instance (pair_invert: one 2048-bit binary GCD on the a component + a Newton
step; DKG_INKERNEL_INVERSE=1) or by the batched inversion kernel in front of it (Montgomery's trick
modulo N^2; DKG_INKERNEL_INVERSE=0).  One wave, device-resident, CUDA events, checked against pow."""
from __future__ import annotations

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

with open(os.path.join(ROOT, "tests", "golden", "dealer_vectors.json")) as fh:
    dv = json.load(fh)["keys"]
dk = bench.KeyData(dv["cfg2_k2048_p3_t1_real"]["key"])
keys = bench.gpu_keys(dk)
n2 = dk.n * dk.n
pid = next(p for p, k in keys.items() if k.partial_decrypt_exponent() < 0)
e = keys[pid].partial_decrypt_exponent()
stream = torch.cuda.current_stream().cuda_stream
for mode in ("1", "0"):
    os.environ["DKG_INKERNEL_INVERSE"] = mode
    ctx = eng.ModexpContext(n2, e, root=dk.n)
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
    assert int(d_st.max().item()) == 0
    got = limbs_to_ints(d_out[B - 1:].cpu().numpy().view(np.uint32))[0]
    assert got == pow(limbs_to_ints(host[B - 1:])[0], e, n2)
    print(json.dumps({"config": "cfg2 reference-shaped key, partial decrypt party %d (negative exponent)" % pid,
                      "inversion": "in the pair kernel" if mode == "1" else "batched inversion kernel", "count": B,
                      "ms": round(ms, 2), "per_s": round(B / ms * 1e3, 1)}), flush=True)
    ctx.close()
