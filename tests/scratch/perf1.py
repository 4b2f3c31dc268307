import sys, time, random, ctypes
sys.path.insert(0, '.')
import numpy as np, torch
import protocols.distributed_keygen_b200 as eng
from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints
from oracle import keys as okeys
import json
dv = json.load(open('tests/golden/dealer_vectors.json'))
for name in sys.argv[1:] or ['cfg2_k2048_p3_t1_exact']:
    dk = okeys.dealer_key_from_json(dv['keys'][name]['key'])
    for pid, key in dk.keys.items():
        e = key.partial_decrypt_exponent()
        ctx = eng.ModexpContext(key.n_square, e, root=(key.n if not __import__('os').environ.get('NOROOT') else None))
        info = ctx.info()
        per_wave = info['ctas'] * info['warps_per_cta'] * 32
        for B in [per_wave, 4 * per_wave]:
            rng = np.random.default_rng(1)
            host = rng.integers(0, 2**32, size=(B, ctx.limbs), dtype=np.uint32)
            host[:, -1] &= (1 << ((key.n_square.bit_length() - 1) % 32)) - 1 if key.n_square.bit_length() % 32 else 0x7fffffff
            d_in = torch.from_numpy(host.view(np.int32)).cuda()
            d_out = torch.empty_like(d_in)
            d_st = torch.empty(B, dtype=torch.uint8, device='cuda')
            s = torch.cuda.current_stream().cuda_stream
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.modexp_device(d_in.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), min(B, 1024), s)
            torch.cuda.synchronize()
            ev0.record()
            ctx.modexp_device(d_in.data_ptr(), d_out.data_ptr(), d_st.data_ptr(), B, s)
            ev1.record(); torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            L = info['padded_limbs']; E = info['exponent_bits']
            macs = (E + (E + 4) // 5 + 32) * (2 * ctx.limbs * ctx.limbs + ctx.limbs)
            out = d_out.cpu().numpy().view(np.uint32)
            # spot check
            idx = [0, B // 2, B - 1]
            ok = all(limbs_to_ints(out[i:i+1])[0] == pow(limbs_to_ints(host[i:i+1])[0], e, key.n_square) for i in idx)
            print(f"{name} party {pid} sign={'-' if e<0 else '+'} {info} B={B} {ms:.1f} ms  {B/ms*1e3:.0f} modexp/s  canonical {B*macs/ms/1e9:.2f} Tmac/s ok={ok}", flush=True)
        ctx.close()
