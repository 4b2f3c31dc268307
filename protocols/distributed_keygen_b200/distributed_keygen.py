"""
Host-side mirror of the batch drivers of the reference's ``DistributedPaillier``
(``distributed_keygen.py`` in tno.mpc.protocols.distributed_keygen v4.2.2) for the hot path only:

* the biprimality-test v calculation (``__biprime_test_v_calculation``, ``:1056-1108``) for one
  candidate and -- what ``compute_modulus``'s list comprehension (``:1313-1329``) becomes -- for a
  whole batch of candidates in one grouped GPU call;
* the biprimality verdict (``__biprime_test_with_v_i``, ``:1110-1175``);
* the two loops of ``_decrypt_sequence_raw`` (``:463-466`` and ``:510-515``) as batched calls, for
  the in-process (``distributed=False``) case where all parties' keys live in one process.

The interactive protocol (pools, message ids, Shamir exchanges, key storage) is out of scope: see
INTEGRATION.md for how these functions are patched into the reference.
"""
from __future__ import annotations

from typing import Iterable, Mapping, Sequence

import numpy as np

from .engine import biprime_v_batch_limbs, biprime_verdict, small_prime_sieve
from .limbs import ints_to_limbs, limbs_for_bits, limbs_to_ints
from .paillier_shared_key import PaillierSharedKey

JACOBI_CORRECTION_FACTOR = 4  # distributed_keygen.py:60


def biprime_exponent(index: int, modulus: int, p_i: int, q_i: int) -> int:
    """``:1092-1097``: party 1 uses (N - p_1 - q_1 + 1) // 4, the others (p_i + q_i) // 4."""
    if index == 1:
        return (modulus - p_i - q_i + 1) // 4
    return (p_i + q_i) // 4


def biprime_test_v_calculation_batch(
    candidates: Sequence[tuple[Sequence[int], int, int, int]],
    index: int,
    correct_param_biprime: int,
    device: int = 0,
) -> list[list[int]]:
    """All candidates of one ``compute_modulus`` round at once.  ``candidates`` holds
    ``(g_values, modulus, p_i, q_i)`` per surviving candidate N (the tuple the reference's list
    comprehension unpacks at ``:1321-1329``); returns this party's v values per candidate."""
    if not candidates:
        return []
    moduli = [c[1] for c in candidates]
    exps = [biprime_exponent(index, n, p_i, q_i) for (_, n, p_i, q_i) in candidates]
    if min(exps) < 0:
        raise ValueError("negative biprimality-test exponent (p_i + q_i or N - p_i - q_i + 1 must be >= 0)")
    # everything on the device: Jacobi filter (sympy.jacobi_symbol at :1089), selection of the first
    # `correct_param_biprime` usable g's, grouped modexp -- one call.  Ragged g lists are padded with
    # zeros: (0 / N) = 0 is never selected.
    ng = max(max(len(c[0]) for c in candidates), 1)
    limbs = limbs_for_bits(max(n.bit_length() for n in moduli))
    exp_limbs = limbs_for_bits(max(max(e.bit_length() for e in exps), 1))
    flat: list[int] = []
    for gs, n, _, _ in candidates:
        flat.extend(g % n for g in gs)
        flat.extend([0] * (ng - len(gs)))
    g_arr = ints_to_limbs(flat, limbs).reshape(len(candidates), ng, limbs)
    v, count = biprime_v_batch_limbs(
        ints_to_limbs(moduli, limbs), ints_to_limbs(exps, exp_limbs), g_arr,
        max(1, min(correct_param_biprime, ng)), device,
    )
    c_eff = v.shape[1]
    vals = limbs_to_ints(v.reshape(-1, limbs))
    return [vals[i * c_eff : i * c_eff + min(int(count[i]), correct_param_biprime)] for i in range(len(candidates))]


def biprime_test_v_calculation(
    g_values: Sequence[int], index: int, modulus: int, p_i: int, q_i: int, correct_param_biprime: int,
    device: int = 0,
) -> list[int]:
    """Single-candidate form with the reference's argument order (``:1056-1065``); returns the list
    the reference stores with ``batched_v_i.set_share(index, v_values)``."""
    return biprime_test_v_calculation_batch(
        [(g_values, modulus, p_i, q_i)], index, correct_param_biprime, device
    )[0]


def biprime_test_with_v_i(
    v_by_party: Mapping[int, Sequence[int]], modulus: int, correct_param_biprime: int
) -> bool:
    """``:1110-1175``: N passes iff the first ``correct_param_biprime`` tests all satisfy
    v_1 = +- prod_{i>1} v_i (mod N); running out of tests is a failure."""
    n_tests = min(len(v) for v in v_by_party.values())
    successful = 0
    for k in range(n_tests):
        product = 1
        for key, values in v_by_party.items():
            if key != 1:
                product *= values[k]
        value1 = v_by_party[1][k]
        if not (value1 % modulus == product % modulus or value1 % modulus == -product % modulus):
            return False
        successful += 1
        if successful >= correct_param_biprime:
            return True
    return False


def biprime_test_with_v_i_batch(
    v_by_party: Mapping[int, Sequence[Sequence[int]]], moduli: Sequence[int], correct_param_biprime: int,
    device: int = 0,
) -> list[bool]:
    """``__biprime_test_with_v_i`` for all candidates of a round on the GPU (in-process parties):
    ``v_by_party[i][g]`` = party i's v values for candidate g."""
    return biprime_verdict(moduli, dict(v_by_party), correct_param_biprime, device)


def small_prime_divisors_test_batch(prime_list: Sequence[int], moduli: Sequence[int], device: int = 0) -> list[bool]:
    """``__small_prime_divisors_test`` (``:1197-1209``) for all candidates of a round at once (the
    filter of ``:1288-1292``): ``True`` where N has a divisor in ``prime_list``."""
    return small_prime_sieve(moduli, prime_list, device)


def threshold_context(keys: Mapping[int, PaillierSharedKey], devices: Sequence[int] | None = None):
    """One multi-GPU context for all in-process parties of a key (``engine.ThresholdContext``): the
    exponents of parties 1..degree+1, theta^-1 and N."""
    from .engine import ThresholdContext

    any_key = next(iter(keys.values()))
    need = range(1, any_key.share.degree + 2)
    return ThresholdContext(any_key.n, any_key.theta_inv, {i: keys[i].partial_decrypt_exponent() for i in need}, devices)


def decrypt_sequence_limbs(
    keys: Mapping[int, PaillierSharedKey], ciphertext_rows: np.ndarray, devices: Sequence[int] | None = None
) -> np.ndarray:
    """``decrypt_sequence_local`` on limb rows, sharded over ``devices`` (default: every GPU of the
    box): [count][limbs(N^2)] -> [count][limbs(N)].  Same exceptions as the reference's per-element
    calls: ``ZeroDivisionError`` (a ciphertext that is not a unit under a negative exponent),
    ``ValueError`` (combined value minus one not divisible by N)."""
    ctx = threshold_context(keys, devices)
    try:
        plain, status, _ = ctx.decrypt_limbs(ciphertext_rows)
    finally:
        ctx.close()
    if (status == 1).any():
        raise ZeroDivisionError("ciphertext not invertible modulo N^2")
    if (status == 3).any():
        raise ValueError("ciphertext value is not below N^2")
    if status.any():
        raise ValueError(
            "Combined decryption minus one is not divisible by N. This might be caused by the "
            "fact that the ciphertext that is being decrypted, differs between the parties."
        )
    return plain


_THRESHOLD_CACHE: dict[tuple, object] = {}


def _cached_threshold_context(keys: Mapping[int, PaillierSharedKey], devices: Sequence[int] | None):
    any_key = next(iter(keys.values()))
    need = range(1, any_key.share.degree + 2)
    ident = (any_key.n, tuple(keys[i].partial_decrypt_exponent() for i in need), tuple(devices) if devices is not None else None)
    ctx = _THRESHOLD_CACHE.get(ident)
    if ctx is None:
        if len(_THRESHOLD_CACHE) >= 8:   # a handful of keys per process; drop the oldest
            _THRESHOLD_CACHE.pop(next(iter(_THRESHOLD_CACHE))).close()
        ctx = _THRESHOLD_CACHE[ident] = threshold_context(keys, devices)
    return ctx


def decrypt_sequence_local(
    keys: Mapping[int, PaillierSharedKey], ciphertexts: Sequence[object], combiner: int | None = None,
    devices: Sequence[int] | None = (0,),
) -> list[int]:
    """The arithmetic of ``_decrypt_sequence_raw`` (``:430-517``) when every party's key is in this
    process: loop 1 (``:463-466``, every party) and loop 2 (``:510-515``) in ONE engine call (the
    ciphertexts are converted to limbs and uploaded once, the partials stay on the device).
    Returns the raw plaintext integers (what the reference wraps in ``EncodedPlaintext``); raises
    like the reference's per-element calls (``ZeroDivisionError`` / ``ValueError``)."""
    any_key = keys[combiner if combiner is not None else min(keys)]
    values = [any_key._raw_value(c) % any_key.n_square for c in ciphertexts]
    ctx = _cached_threshold_context(keys, devices)
    plain, status, _ = ctx.decrypt_limbs(ints_to_limbs(values, ctx.n2_limbs))
    if (status == 1).any():
        raise ZeroDivisionError("ciphertext not invertible modulo N^2")
    if status.any():
        raise ValueError(
            "Combined decryption minus one is not divisible by N. This might be caused by the "
            "fact that the ciphertext that is being decrypted, differs between the parties."
        )
    return limbs_to_ints(plain)


def partial_decryption_message(key: PaillierSharedKey, ciphertext_rows: np.ndarray) -> bytes:
    """Loop 1 of ``_decrypt_sequence_raw`` plus the body of its broadcast (``:463-484``) with no
    Python int in between: ciphertext limb rows -> batched partial decryption -> message bytes
    (``wire.pack_partial_decryption_message``).  ``ZeroDivisionError`` if a ciphertext is not a
    unit and the exponent is negative, as ``mod_inv`` raises in the reference."""
    from . import wire

    rows, status = key.partial_decrypt_limbs(ciphertext_rows)
    if (status == 3).any():
        raise ValueError("ciphertext value is not below N^2")
    if status.any():
        raise ZeroDivisionError("ciphertext not invertible modulo N^2")
    return wire.pack_partial_decryption_message(rows)


def decrypt_from_messages(key: PaillierSharedKey, messages: Mapping[int, bytes]) -> np.ndarray:
    """The receive side (``:497-515``): one message body per party index -> plaintext limb rows
    [count][limbs(N)].  Indexing parties 1..degree+1 raises ``KeyError`` for a missing one
    (``paillier_shared_key.py:108-110``); a failed divisibility check raises the reference's
    ``ValueError`` (``:119-123``)."""
    from . import wire
    from .paillier_shared_key import n_square_limbs

    l2 = n_square_limbs(key.n)
    shares = key.share.degree + 1
    parts = [wire.unpack_partial_decryption_message(messages[i + 1], l2) for i in range(shares)]
    if len({p.shape[0] for p in parts}) != 1:
        raise ValueError("partial decryption messages of different lengths")
    out, status = key.decrypt_limbs(np.stack(parts))
    if status.any():
        raise ValueError(
            "Combined decryption minus one is not divisible by N. This might be caused by the "
            "fact that the ciphertext that is being decrypted, differs between the parties."
        )
    return out
