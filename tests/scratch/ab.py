import sys, json, os
sys.path.insert(0, ".")
import numpy as np, torch
import protocols.distributed_keygen_b200 as eng
from oracle import keys as okeys
dv = json.load(open("tests/golden/dealer_vectors.json"))
name = os.environ.get("KEY", "cfg2_k2048_p3_t1_exact")
dk = okeys.dealer_key_from_json(dv["keys"][name]["key"])
pid = int(os.environ.get("PID", "1"))
key = dk.keys[pid]; e = key.partial_decrypt_exponent()
ctx = eng.ModexpContext(key.n_square, e, root=None if os.environ.get("NOROOT") else key.n)
info = ctx.info(); B = info["ctas"] * (info["pair_warps_per_cta"] if info["pair_arithmetic"] else info["warps_per_cta"]) * 32 * int(os.environ.get("WAVES", "1"))
host = np.random.default_rng(1).integers(0, 2**32, size=(B, ctx.limbs), dtype=np.uint32); host[:, -1] &= 0x3fffffff
d_in = torch.from_numpy(host.view(np.int32)).cuda(); d_out = torch.empty_like(d_in); d_st = torch.empty(B, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
ctx.modexp_device(d_in.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), 4096, s); torch.cuda.synchronize()
best = 1e9
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.modexp_device(d_in.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), B, s); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
from protocols.distributed_keygen_b200.limbs import limbs_to_ints
out = d_out.cpu().numpy().view(np.uint32)
ok = all(limbs_to_ints(out[i:i+1])[0] == pow(limbs_to_ints(host[i:i+1])[0], e, key.n_square) for i in (0, 77, B - 1))
print(os.environ.get("TAG", ""), "w", info["window_bits"], "nmul", info["windows"], name, "pid", pid, "neg" if e < 0 else "pos", "B", B, "%.1f ms %.0f /s" % (best, B / best * 1e3), "ok" if ok else "MISMATCH")
