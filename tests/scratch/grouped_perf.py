import sys, time, random, json
sys.path.insert(0, '.')
import numpy as np
import protocols.distributed_keygen_b200 as eng
from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints
from oracle import keys as okeys
rng = random.Random(5); pl = 1024
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
moduli, exps = [], []
for _ in range(C):
    p_sh = [okeys.prime_candidate_share(i + 1, pl, rng) for i in range(3)]; q_sh = [okeys.prime_candidate_share(i + 1, pl, rng) for i in range(3)]
    n = sum(p_sh) * sum(q_sh); moduli.append(n); exps.append((n - p_sh[0] - q_sh[0] + 1) // 4)
L = 65
m_arr = ints_to_limbs(moduli, L); e_arr = ints_to_limbs(exps, 65)
bases = np.random.default_rng(11).integers(0, 2**32, size=(C, 40, L), dtype=np.uint32); bases[:, :, -1] = 0
eng.modexp_grouped_limbs(m_arr[:8], e_arr[:8], bases[:8])
for rep in range(2):
    t0 = time.perf_counter(); out = eng.modexp_grouped_limbs(m_arr, e_arr, bases); secs = time.perf_counter() - t0
    ok = limbs_to_ints(out[C - 1, 39:40])[0] == pow(limbs_to_ints(bases[C - 1, 39:40])[0], exps[-1], moduli[-1])
    print(C, 'candidates', f'{secs*1e3:.1f} ms', f'{C*40/secs:.0f} modexp/s', ok, flush=True)
