"""The one-call, several-GPUs path (C ABI ``dkg_threshold_*``): every party's
partial decryption plus the combination of ``_decrypt_sequence_raw`` (``distributed_keygen.py:463-466,
510-515``) in one call, sharded by index over the devices, against the values recorded from the
reference and against the per-party contexts."""
from __future__ import annotations

import base64
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _h(x: str) -> int:
    return int(x, 16)


def _devices():
    from protocols.distributed_keygen_b200 import _native

    n = _native.device_count()
    return [[0]] + ([list(range(n))] if n > 1 else [])


def _gpu_keys(okeys_by_pid):
    from protocols.distributed_keygen_b200 import IntegerShares, PaillierSharedKey

    out = {}
    for pid, k in okeys_by_pid.items():
        share = IntegerShares(dict(k.share.shares), k.share.degree, k.share.scaling, k.share.number_of_parties)
        out[pid] = PaillierSharedKey(k.n, k.t, k.player_id, share, k.theta)
    return out


@pytest.mark.parametrize("chunk", ["7", ""])
def test_threshold_context_matches_reference_vectors(fixture_vectors, dealer_vectors, monkeypatch, chunk):
    """Golden keys (the reference's 24 stored keys + dealer keys up to 2048 bit): plaintexts,
    partials and the tampered-partial status; tiny chunks exercise the double-buffered pipeline."""
    from oracle import keys as okeys
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    if chunk:
        monkeypatch.setenv("DKG_CHUNK_ROWS", chunk)
    sets = []
    for entry in fixture_vectors["sets"]:
        sets.append(({k["player_id"]: okeys.key_from_blob(base64.b64decode(k["blob_b64"])) for k in entry["keys"]}, entry["vectors"]))
    for name in ("small_k128_p3_t1", "cfg1_k512_p3_t1", "cfg2_k2048_p3_t1_real", "cfg3_k2048_p5_t2_exact"):
        item = dealer_vectors["keys"][name]
        sets.append((okeys.dealer_key_from_json(item["key"]).keys, item["vectors"]))
    for okeys_by_pid, vectors in sets:
        keys = _gpu_keys(okeys_by_pid)
        good = [v for v in vectors if "error" not in v]
        n = keys[1].n
        l2 = ((n * n).bit_length() + 31) // 32
        rows = ints_to_limbs([_h(v["c"]) for v in good], l2)
        for devices in _devices():
            ctx = dkg.threshold_context(keys, devices)
            plain, status, parts = ctx.decrypt_limbs(rows, want_partials=True)
            assert not status.any()
            assert limbs_to_ints(plain) == [_h(v["plaintext"]) for v in good]
            for p in range(ctx.shares):
                assert limbs_to_ints(parts[p]) == [_h(v["partials"][str(p + 1)]) for v in good]
                one, st = ctx.partial_decrypt_limbs(p + 1, rows)
                assert not st.any() and np.array_equal(one, parts[p])
            plain2, st2 = ctx.combine_limbs(parts)
            assert not st2.any() and np.array_equal(plain2, plain)
            if len(good) >= 3:
                parts[0, 1, 0] ^= 1
                plain3, st3 = ctx.combine_limbs(parts)
                assert st3[1] == 2 and not np.delete(st3, 1).any() and not plain3[1].any()
            ctx.close()
            assert limbs_to_ints(dkg.decrypt_sequence_limbs(keys, rows, devices)) == [_h(v["plaintext"]) for v in good]
        for key in keys.values():
            key.close()


def test_threshold_context_large_batch_roundtrip(dealer_vectors):
    """encrypt -> sharded threshold decrypt round trip on 20 000 ciphertexts (512-bit key): the
    thread-per-operand route, several chunks per device, per-element status of non-units."""
    import math

    import protocols.distributed_keygen_b200 as eng
    from oracle import keys as okeys
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    dk = okeys.dealer_key_from_json(dealer_vectors["keys"]["cfg1_k512_p3_t1"]["key"])
    keys = _gpu_keys(dk.keys)
    rng = random.Random(9)
    count = 20000
    ms = [rng.randrange(dk.n) for _ in range(count)]
    enc = eng.EncryptContext(dk.n)
    m_rows = ints_to_limbs(ms, enc.n_limbs)
    r_rows = ints_to_limbs([rng.randrange(1, dk.n) for _ in range(count)], enc.n_limbs)
    cts = enc.encrypt_limbs(r_rows, m_rows)
    enc.close()
    for devices in _devices():
        ctx = dkg.threshold_context(keys, devices)
        with eng.pinned(cts):
            plain, status, _ = ctx.decrypt_limbs(cts)
        ctx.close()
        assert not status.any()
        assert np.array_equal(plain, m_rows)
    # a non-unit ciphertext (multiple of a prime factor) is flagged per element when an exponent is negative
    exps = {pid: k.partial_decrypt_exponent() for pid, k in dk.keys.items()}
    p = sum(dk.p_shares)
    bad = cts.copy()
    bad[5] = ints_to_limbs([p * 12345], cts.shape[1])[0]
    ctx = dkg.threshold_context(keys, [0])
    plain, status, _ = ctx.decrypt_limbs(bad[:64])
    ctx.close()
    if any(e < 0 for e in exps.values()):
        assert status[5] == 1 and not np.delete(status, 5).any()
    else:
        assert status[5] in (0, 2)
    assert np.array_equal(np.delete(plain, 5, axis=0), np.delete(m_rows[:64], 5, axis=0))
    for key in keys.values():
        key.close()


def test_partials_only_call_equals_the_decrypt_call(dealer_vectors):
    """``dkg_threshold_partials_batch`` (loop 1 of ``_decrypt_sequence_raw`` for every in-process
    party, ``distributed_keygen.py:463-466``): the same partials as the full decrypt call returns,
    a status per party, on one device and sharded over all of them, small and large batches."""
    import protocols.distributed_keygen_b200 as eng
    from oracle import keys as okeys
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs

    dk = okeys.dealer_key_from_json(dealer_vectors["keys"]["cfg1_k512_p3_t1"]["key"])
    keys = _gpu_keys(dk.keys)
    n2 = dk.n * dk.n
    rng = random.Random(77)
    exps = {pid: k.partial_decrypt_exponent() for pid, k in dk.keys.items()}
    for count in (5, 7001):
        cs = [rng.randrange(1, n2) for _ in range(count)]
        cs[3] = sum(dk.p_shares) * 4711 % n2        # not a unit: flagged for the parties with a negative exponent only
        rows = ints_to_limbs(cs, (n2.bit_length() + 31) // 32)
        for devices in _devices():
            ctx = dkg.threshold_context(keys, devices)
            parts, status = ctx.partials_limbs(rows)
            _, _, want = ctx.decrypt_limbs(rows, want_partials=True)
            ctx.close()
            assert status.shape == (ctx.shares, count)
            for p in range(ctx.shares):
                neg = exps[p + 1] < 0
                assert status[p, 3] == (1 if neg else 0) and not np.delete(status[p], 3).any()
                keep = np.ones(count, dtype=bool)
                keep[3] = not neg
                assert np.array_equal(parts[p][keep], want[p][keep])
    for k in keys.values():
        k.close()
