"""All parties' partial decryptions of one ciphertext batch through ONE squaring chain
(``modexp_nsq_multi_kernel``, csrc/dkg_nsq.cuh; C ABI ``dkg_threshold_decrypt_batch[_device]``):
every party's partial must be the very value its own ``partial_decrypt`` gives
(``ref: paillier_shared_key.py:52-93``), hence the reference-recorded golden values and digests, the
per-party kernels, and CPython ``pow``."""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from digests import digest_inputs, sha  # noqa: E402

from conftest import load_golden  # noqa: E402

pytestmark = pytest.mark.gpu


def _h(x: str) -> int:
    return int(x, 16)


@pytest.fixture()
def wave_route():
    """Send every batch, however small, to the thread-per-ciphertext kernels."""
    from protocols.distributed_keygen_b200 import _native

    saved = _native.config_get("coop_max")
    _native.config_set("coop_max", 0)
    yield _native
    _native.config_set("coop_max", saved)


def _gpu_keys(okeys_by_pid):
    from protocols.distributed_keygen_b200 import IntegerShares, PaillierSharedKey

    out = {}
    for pid, k in okeys_by_pid.items():
        share = IntegerShares(dict(k.share.shares), k.share.degree, k.share.scaling, k.share.number_of_parties)
        out[pid] = PaillierSharedKey(k.n, k.t, k.player_id, share, k.theta)
    return out


def _info_ex(ctx):
    import ctypes

    from protocols.distributed_keygen_b200 import _native

    arr = (ctypes.c_int * 8)()
    _native.check(_native.lib.dkg_threshold_info_ex(ctx._h, ctypes.byref(arr)))
    return list(arr)


@pytest.mark.parametrize("name", ["small_k128_p3_t1", "cfg1_k512_p3_t1", "cfg2_k2048_p3_t1_exact", "cfg2_k2048_p3_t1_real",
                                  "cfg3_k2048_p5_t2_exact", "cfg4_k4096_p3_t1_exact"])
def test_shared_chain_reproduces_reference_partials(wave_route, dealer_vectors, name):
    """Golden vectors (values) and, where recorded, the 64-160 digests per key of the reference's own
    ``partial_decrypt`` / ``decrypt``: the shared chain must reproduce every party's partial."""
    from oracle import keys as okeys
    from oracle.paillier_oracle import encrypt_raw
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    item = dealer_vectors["keys"][name]
    dk = okeys.dealer_key_from_json(item["key"])
    keys = _gpu_keys(dk.keys)
    ctx = dkg.threshold_context(keys, [0])
    info = _info_ex(ctx)
    assert info[0] == 1, "shared squaring chain not selected"
    l2 = ctx.n2_limbs
    good = [v for v in item["vectors"] if "error" not in v]
    plain, status, parts = ctx.decrypt_limbs(ints_to_limbs([_h(v["c"]) for v in good], l2), want_partials=True)
    assert not status.any()
    assert limbs_to_ints(plain) == [_h(v["plaintext"]) for v in good]
    for p in range(ctx.shares):
        assert limbs_to_ints(parts[p]) == [_h(v["partials"][str(p + 1)]) for v in good], (name, p + 1)
    digests = load_golden("dealer_digests.json")["keys"].get(name)
    if digests is not None:
        rows = digests["vectors"]
        inputs = digest_inputs(dk.n, digests["seed"], len(rows))
        cs = [encrypt_raw(dk.n, m, r) for m, r in inputs]
        plain, status, parts = ctx.decrypt_limbs(ints_to_limbs(cs, l2), want_partials=True)
        assert not status.any()
        for p in range(ctx.shares):
            assert [sha(v) for v in limbs_to_ints(parts[p])] == [v["partials"][str(p + 1)] for v in rows], (name, p + 1)
        assert limbs_to_ints(plain) == [m for m, _ in inputs]
    ctx.close()
    for k in keys.values():
        k.close()


@pytest.mark.parametrize("window", ["1", "2", "3", "5", "7", "8"])
def test_every_window_width(wave_route, dealer_vectors, monkeypatch, window):
    """Bucket aggregation for every digit width (1: a single bucket, no folding; 8: 255 buckets),
    ragged batch (77 rows: two full groups and a partial one), random units against CPython pow."""
    from oracle import keys as okeys
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    monkeypatch.setenv("DKG_MULTI_WINDOW", window)
    dk = okeys.dealer_key_from_json(dealer_vectors["keys"]["small_k128_p3_t1"]["key"])
    keys = _gpu_keys(dk.keys)
    ctx = dkg.threshold_context(keys, [0])
    assert _info_ex(ctx)[:2] == [1, int(window)]
    n2 = dk.n * dk.n
    rng = random.Random(int(window))
    cs = [rng.randrange(1, n2) for _ in range(77)]
    _, _, parts = ctx.decrypt_limbs(ints_to_limbs(cs, ctx.n2_limbs), want_partials=True)
    ctx.close()
    for pid, k in dk.keys.items():
        e = k.partial_decrypt_exponent()
        want = [pow(pow(c, -1, n2), -e, n2) if e < 0 else pow(c, e, n2) for c in cs]
        assert limbs_to_ints(parts[pid - 1]) == want, (window, pid)
    for k in keys.values():
        k.close()


def test_shared_chain_equals_per_party_kernels_and_flags_bad_rows(wave_route, dealer_vectors, monkeypatch):
    """3000 rows (512-bit key) through the shared chain and, with DKG_SHARED_SQUARINGS=0, through one
    exponentiation per party: identical partials, plaintexts and status bytes -- including a
    non-unit (status 1 under a negative exponent: the predicated per-element redo) and a row >= N^2
    (status 3)."""
    from oracle import keys as okeys
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    dk = okeys.dealer_key_from_json(dealer_vectors["keys"]["cfg1_k512_p3_t1"]["key"])
    keys = _gpu_keys(dk.keys)
    n2 = dk.n * dk.n
    rng = random.Random(31)
    count = 3000
    cs = [rng.randrange(1, n2) for _ in range(count)]
    p = sum(dk.p_shares)
    cs[1234] = p * 98765 % n2          # not a unit
    rows = ints_to_limbs(cs, ((n2.bit_length() + 31) // 32))
    rows[2999] = 0xFFFFFFFF            # >= N^2
    results = []
    for shared in ("1", "0"):
        monkeypatch.setenv("DKG_SHARED_SQUARINGS", shared)
        ctx = dkg.threshold_context(keys, [0])
        assert _info_ex(ctx)[0] == int(shared)
        results.append(ctx.decrypt_limbs(rows, want_partials=True))
        ctx.close()
    (plain_a, st_a, parts_a), (plain_b, st_b, parts_b) = results
    assert np.array_equal(st_a, st_b)
    exps = {pid: k.partial_decrypt_exponent() for pid, k in dk.keys.items()}
    if any(e < 0 for e in exps.values()):
        assert st_a[1234] == 1
    assert st_a[2999] == 3
    ok = np.flatnonzero(st_a == 0)
    assert ok.size >= count - 2 - 0 and np.array_equal(parts_a[:, ok], parts_b[:, ok]) and np.array_equal(plain_a[ok], plain_b[ok])
    for pid, e in exps.items():
        for i in (0, 1, 1233, 1235, 2998):
            c = cs[i]
            want = pow(pow(c, -1, n2), -e, n2) if e < 0 else pow(c, e, n2)
            assert limbs_to_ints(parts_a[pid - 1, i : i + 1])[0] == want
    for k in keys.values():
        k.close()


def test_device_resident_call(dealer_vectors):
    """``dkg_threshold_decrypt_batch_device``: device pointers in and out on the caller's stream, a
    batch large enough for the wave route without any override (40 000 x 3 instances)."""
    import torch

    import protocols.distributed_keygen_b200 as eng
    from oracle import keys as okeys
    from protocols.distributed_keygen_b200 import _native
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs

    dk = okeys.dealer_key_from_json(dealer_vectors["keys"]["cfg1_k512_p3_t1"]["key"])
    keys = _gpu_keys(dk.keys)
    rng = random.Random(5)
    count = 40000
    enc = eng.EncryptContext(dk.n)
    m_rows = ints_to_limbs([rng.randrange(dk.n) for _ in range(count)], enc.n_limbs)
    r_rows = ints_to_limbs([rng.randrange(1, dk.n) for _ in range(count)], enc.n_limbs)
    cts = enc.encrypt_limbs(r_rows, m_rows)
    enc.close()
    ctx = dkg.threshold_context(keys, [0])
    S, l2, ln = ctx.shares, ctx.n2_limbs, ctx.n_limbs
    d_cts = torch.from_numpy(cts.view(np.int32)).cuda()
    d_parts = torch.empty((S, count, l2), dtype=torch.int32, device="cuda")
    d_plain = torch.empty((count, ln), dtype=torch.int32, device="cuda")
    d_st = torch.full(((S + 1) * count,), 7, dtype=torch.uint8, device="cuda")
    _native.check(_native.lib.dkg_threshold_decrypt_batch_device(
        ctx._h, d_cts.data_ptr(), d_plain.data_ptr(), d_parts.data_ptr(), d_st.data_ptr(), count,
        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert int(d_st.max().item()) == 0
    assert np.array_equal(d_plain.cpu().numpy().view(np.uint32), m_rows)
    host_parts = d_parts.cpu().numpy().view(np.uint32)
    for p in range(S):
        one, st = ctx.partial_decrypt_limbs(p + 1, cts[:2048])
        assert not st.any() and np.array_equal(one, host_parts[p, :2048])
    ctx.close()
    for k in keys.values():
        k.close()
