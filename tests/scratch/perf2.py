import sys, time, random, json
sys.path.insert(0, '.')
import numpy as np, torch
import protocols.distributed_keygen_b200 as eng
from oracle import keys as okeys
dv = json.load(open('tests/golden/dealer_vectors.json'))
dk = okeys.dealer_key_from_json(dv['keys']['cfg2_k2048_p3_t1_exact']['key'])
n2 = dk.n * dk.n
rng = random.Random(5)
exps = {'A_zero_digits': 1 << 4185, 'B_all63': (1 << 4186) - 1, 'C_random': rng.getrandbits(4186) | (1 << 4185),
        'D_short_4186sq_only_w1': None}
B = 33152
host = np.random.default_rng(1).integers(0, 2**32, size=(B, 128), dtype=np.uint32); host[:, -1] &= 0x7fffffff
d_in = torch.from_numpy(host.view(np.int32)).cuda(); d_out = torch.empty_like(d_in); d_st = torch.empty(B, dtype=torch.uint8, device='cuda')
s = torch.cuda.current_stream().cuda_stream
import os
for name, e in [('C_random', exps['C_random'])]:
    if e is None: continue
    ctx = eng.ModexpContext(n2, e)
    ctx.modexp_device(d_in.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), 1024, s); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); ctx.modexp_device(d_in.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), B, s); ev1.record(); torch.cuda.synchronize()
    print(name, ctx.info(), f"{ev0.elapsed_time(ev1):.1f} ms", flush=True)
    ctx.close()
