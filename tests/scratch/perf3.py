import sys, json, os
sys.path.insert(0, '.')
import numpy as np, torch
import protocols.distributed_keygen_b200 as eng
from oracle import keys as okeys
dv = json.load(open('tests/golden/dealer_vectors.json'))
dk = okeys.dealer_key_from_json(dv['keys']['cfg2_k2048_p3_t1_exact']['key'])
key = dk.keys[1]; e = key.partial_decrypt_exponent()
ctx = eng.ModexpContext(key.n_square, e, root=key.n)
info = ctx.info(); B = info['ctas'] * info['warps_per_cta'] * 32
host = np.random.default_rng(1).integers(0, 2**32, size=(B, ctx.limbs), dtype=np.uint32); host[:, -1] &= 0x7fffffff
d_in = torch.from_numpy(host.view(np.int32)).cuda(); d_out = torch.empty_like(d_in); d_st = torch.empty(B, dtype=torch.uint8, device='cuda')
s = torch.cuda.current_stream().cuda_stream
ctx.modexp_device(d_in.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), B, s); torch.cuda.synchronize()
print(info)
