"""GPU parity for the encryption path: r^N mod N^2 and (1 + m N) r^N mod N^2 against the oracle,
and encrypt -> threshold decrypt round trips through the CUDA kernels only."""
from __future__ import annotations

import base64
import random

import pytest

pytestmark = pytest.mark.gpu


def test_encrypt_matches_oracle(dealer_vectors, fixture_vectors):
    import protocols.distributed_keygen_b200 as eng
    from oracle import keys as okeys
    from oracle.paillier_oracle import encrypt_raw, randomness

    rng = random.Random(17)
    ns = [okeys.dealer_key_from_json(dealer_vectors["keys"][k]["key"]).n
          for k in ("small_k128_p3_t1", "cfg1_k512_p3_t1", "cfg2_k2048_p3_t1_real")]
    ns.append(okeys.key_from_blob(base64.b64decode(fixture_vectors["sets"][0]["keys"][0]["blob_b64"])).n)
    for n in ns:
        ctx = eng.EncryptContext(n)
        count = 70 if n.bit_length() < 1000 else 34
        ms = [0, 1, n - 1] + [rng.randrange(n) for _ in range(count - 3)]
        rs = [1, n - 1, 2] + [rng.randrange(1, n) for _ in range(count - 3)]
        assert ctx.encrypt(ms, rs) == [encrypt_raw(n, m, r) for m, r in zip(ms, rs)]
        assert ctx.randomness(rs) == [randomness(n, r) for r in rs]
        ctx.close()


def test_encrypt_decrypt_roundtrip_gpu_only(dealer_vectors):
    """decrypt(encrypt(m)) == m with every modular operation on the GPU
    (the reference's own test criterion, test_distributed_keygen.py:111-129)."""
    import protocols.distributed_keygen_b200 as eng
    from oracle import keys as okeys

    dk = okeys.dealer_key_from_json(dealer_vectors["keys"]["cfg1_k512_p3_t1"]["key"])
    rng = random.Random(23)
    ms = [1, -1, 2, -2, 3, -3, 150000000, -150000000, 4242424242, -4242424242] + [
        rng.randrange(-(2**60), 2**60) for _ in range(90)]
    enc = eng.EncryptContext(dk.n)
    cts = enc.encrypt(ms, [rng.randrange(1, dk.n) for _ in ms])
    keys = {}
    for pid, k in dk.keys.items():
        share = eng.IntegerShares(dict(k.share.shares), k.share.degree, k.share.scaling, k.share.number_of_parties)
        keys[pid] = eng.PaillierSharedKey(k.n, k.t, pid, share, k.theta)
    parts = {pid: key.partial_decrypt_batch(cts) for pid, key in keys.items()}
    got = keys[3].decrypt_batch([{pid: parts[pid][i] for pid in keys} for i in range(len(ms))])
    assert got == [m % dk.n for m in ms]
    enc.close()
    for key in keys.values():
        key.close()
