"""
Drop-in switch: make the reference's OWN classes (``tno.mpc.protocols.distributed_keygen`` v4.2.2)
run their hot path on the B200 engine, without editing the reference.

    import tno.mpc.protocols.distributed_keygen          # the unmodified reference
    from protocols.distributed_keygen_b200 import patch
    patch.install()            # or: DKG_B200=1 in the environment + patch.install_from_env()

What ``install()`` rebinds (reference file:line -> what runs instead):

* ``PaillierSharedKey.partial_decrypt`` (``paillier_shared_key.py:52-93``): same type / key checks
  (``:62-68``) and the same ``ciphertext.get_value()`` freshness bookkeeping (``:69``), then one
  GPU modexp with the per-key exponent (computed once per key, not per call).  Called by the
  reference's own ``_decrypt_raw`` (``distributed_keygen.py:346-350``) and, per element, by its
  own ``_decrypt_sequence_raw`` loop when ``batched_sequence=False``.
* ``PaillierSharedKey.decrypt`` (``:95-127``): GPU share combination, same ``KeyError`` /
  ``ValueError`` behaviour.
* new ``PaillierSharedKey.partial_decrypt_batch`` / ``decrypt_batch``: the two loops of
  ``_decrypt_sequence_raw`` (``distributed_keygen.py:463-466``, ``:510-515``) as one call each.
* ``DistributedPaillier._decrypt_sequence_raw`` (``:430-517``), when ``batched_sequence=True``
  (default): the same messages with the same ids and contents, but both loops go through the
  batched calls above.
* ``DistributedPaillier.__biprime_test_v_calculation`` (``:1056-1108``, name-mangled classmethod,
  patched the way ``scripts/bench_batch_size.py:94-103`` patches its siblings): Jacobi filter,
  first-``correct_param_biprime`` selection and the modexps in one GPU call; returns the same
  ``Batched[AdditiveVariable]`` object.
* new ``DistributedPaillier._b200_biprime_v_batch``: all candidates of a ``compute_modulus`` round
  (the list comprehension at ``:1313-1329``) in one call, for maintainers who edit that line.

In-process parties (``distributed=False``, the reference's tests and benchmark,
``distributed_keygen.py:203-226``): every party's coroutine calls ``partial_decrypt_batch`` on the SAME
ciphertexts.  When the keys of all parties 1..degree+1 of a modulus are live in this process and the
batch is large, the first such call computes every party's partial decryptions in one engine call
(one shared chain of squarings on the device) and the other parties' calls
pick theirs up; each party still gets exactly the values its own call would return.
``DKG_B200_SHARE=0`` turns this off, ``DKG_B200_SHARE_MIN`` sets the smallest batch (default 4096).

``uninstall()`` restores every attribute.  Nothing here computes on the host: without the CUDA
library / a device the patched methods raise (there is no CPU fallback).
"""
from __future__ import annotations

import hashlib
import importlib
import os
import weakref
from typing import Any, Iterable, Mapping, Sequence

from . import distributed_keygen as _dk
from .paillier_shared_key import IntegerShares as _Shares
from .paillier_shared_key import PaillierSharedKey as _GpuKey

_ORIGINALS: list[tuple[Any, str, Any, bool]] = []   # (owner, attribute, old value, existed)
_DEVICE = 0
_PEERS: dict[int, dict[int, Any]] = {}               # n -> {player_id: weak reference to the reference key}
_SHARED: dict[tuple[int, bytes], dict[str, Any]] = {}   # (n, digest of the batch) -> partials of all parties


def _ref_modules(ref_pkg: Any = None) -> tuple[Any, Any, Any]:
    pkg = ref_pkg if ref_pkg is not None else importlib.import_module("tno.mpc.protocols.distributed_keygen")
    name = pkg.__name__
    return (
        importlib.import_module(name + ".distributed_keygen"),
        importlib.import_module(name + ".paillier_shared_key"),
        importlib.import_module(name + ".utils"),
    )


def gpu_key(ref_key: Any) -> _GpuKey:
    """The engine-side twin of a reference ``PaillierSharedKey`` (contexts are created lazily and
    cached on the reference object, so the exponent / Montgomery constants are per key)."""
    twin = ref_key.__dict__.get("_b200_twin")
    if twin is None or twin.n != ref_key.n or twin.share.shares != dict(ref_key.share.shares):
        sh = ref_key.share
        twin = _GpuKey(
            ref_key.n, ref_key.t, ref_key.player_id,
            _Shares(dict(sh.shares), sh.degree, sh.scaling, sh.scheme.number_of_parties),
            ref_key.theta, device=_DEVICE,
        )
        ref_key.__dict__["_b200_twin"] = twin
    return twin


def _shared_partials(ref_key: Any, values: Sequence[int]) -> list[int] | None:
    """This party's partial decryptions out of ONE engine call for all in-process parties of the key,
    or None when that does not apply (see the module docstring)."""
    from .limbs import ints_to_limbs, limbs_to_ints

    if os.environ.get("DKG_B200_SHARE", "1") == "0" or not (
            int(os.environ.get("DKG_B200_SHARE_MIN", "4096")) <= len(values) <= (1 << 21)):
        return None
    twin = gpu_key(ref_key)
    n = twin.n
    try:
        _PEERS.setdefault(n, {})[twin.player_id] = weakref.ref(ref_key)
    except TypeError:          # a key class without weak-reference support: no sharing
        return None
    need = range(1, twin.share.degree + 2)
    live = {i: r() for i, r in _PEERS[n].items()}
    if twin.player_id not in need or any(live.get(i) is None for i in need):
        return None
    twins = {i: gpu_key(live[i]) for i in need}
    if any(k.t != twin.t or k.theta != twin.theta or k.share.degree != twin.share.degree for k in twins.values()):
        return None
    ctx = _dk._cached_threshold_context(twins, (_DEVICE,))
    rows = ints_to_limbs([v % twin.n_square for v in values], ctx.n2_limbs)
    ident = (n, hashlib.blake2b(rows.tobytes(), digest_size=16).digest())
    entry = _SHARED.get(ident)
    if entry is None or twin.player_id not in entry["left"]:
        parts, status = ctx.partials_limbs(rows)
        while len(_SHARED) >= 2:
            _SHARED.pop(next(iter(_SHARED)))
        entry = _SHARED[ident] = {"parts": parts, "status": status, "left": set(need)}
    entry["left"].discard(twin.player_id)
    mine, status = entry["parts"][twin.player_id - 1], entry["status"][twin.player_id - 1]
    if not entry["left"]:
        _SHARED.pop(ident, None)
    if status.any():
        raise ZeroDivisionError("base is not invertible for the given modulus")
    return limbs_to_ints(mine)


def _set(owner: Any, attr: str, value: Any) -> None:
    existed = attr in owner.__dict__
    _ORIGINALS.append((owner, attr, owner.__dict__.get(attr), existed))
    setattr(owner, attr, value)


def installed() -> bool:
    return bool(_ORIGINALS)


def uninstall() -> None:
    _PEERS.clear()
    _SHARED.clear()
    while _ORIGINALS:
        owner, attr, old, existed = _ORIGINALS.pop()
        if existed:
            setattr(owner, attr, old)
        else:
            delattr(owner, attr)


def install(ref_pkg: Any = None, device: int = 0, batched_sequence: bool = True) -> None:
    """Rebind the reference's hot-path methods to the B200 engine (see the module docstring)."""
    global _DEVICE
    if installed():
        uninstall()
    _DEVICE = device
    ref_dk, ref_psk, ref_utils = _ref_modules(ref_pkg)
    RefKey = ref_psk.PaillierSharedKey
    RefScheme = ref_dk.DistributedPaillier
    RefCiphertext = ref_psk.PaillierCiphertext
    EncodedPlaintext = ref_dk.EncodedPlaintext

    def checked_value(self: Any, ciphertext: Any) -> int:
        # paillier_shared_key.py:62-69
        if not isinstance(ciphertext, RefCiphertext):
            raise TypeError(f"Expected ciphertext to be a PaillierCiphertext not: {type(ciphertext)}")
        if self.n != ciphertext.scheme.public_key.n:
            raise ValueError("encrypted against a different key!")
        return int(ciphertext.get_value())

    def partial_decrypt_batch(self: Any, ciphertexts: Iterable[Any]) -> list[int]:
        values = [checked_value(self, c) for c in ciphertexts]
        shared = _shared_partials(self, values)
        return shared if shared is not None else gpu_key(self).partial_decrypt_batch(values)

    def partial_decrypt(self: Any, ciphertext: Any) -> int:
        return partial_decrypt_batch(self, [ciphertext])[0]

    def decrypt_batch(self: Any, partial_dicts: Sequence[Mapping[int, int]]) -> list[int]:
        return gpu_key(self).decrypt_batch(partial_dicts)

    def decrypt(self: Any, partial_dict: Mapping[int, int]) -> int:
        return decrypt_batch(self, [partial_dict])[0]

    _set(RefKey, "partial_decrypt", partial_decrypt)
    _set(RefKey, "decrypt", decrypt)
    _set(RefKey, "partial_decrypt_batch", partial_decrypt_batch)
    _set(RefKey, "decrypt_batch", decrypt_batch)

    async def decrypt_sequence_raw(self: Any, ciphertext_sequence: Iterable[Any],
                                   receivers: list[str] | None = None) -> list[Any] | None:
        """Batched ``_decrypt_sequence_raw`` (distributed_keygen.py:430-517): identical wire
        protocol (message id :469-475, content "partial_decryption_sequence" :476-484, receive
        :494-505); the two per-ciphertext loops (:463-466, :510-515) are one GPU call each."""
        ciphertexts = list(ciphertext_sequence)
        self_receive = receivers is None or "self" in receivers
        others = None if receivers is None else [r for r in receivers if r != "self"]
        key = self.secret_key
        mine = key.partial_decrypt_batch(ciphertexts)
        tag = bin(ciphertexts[0].peek_value()).zfill(32)[2:34] + f"{len(mine)}"
        message_id = f"distributed_decryption_session#{self.session_id}_hash#{tag}"
        if others is None or len(others) != 0:
            self.pool.async_broadcast(
                {"content": "partial_decryption_sequence", "value": mine},
                msg_id=message_id, handler_names=others,
            )
        if not self_receive:
            return None
        per_ciphertext: list[dict[int, int]] = [{self.index: share} for share in mine]
        for party, message in await self.pool.recv_all(msg_id=message_id):
            content = message["content"]
            assert content == "partial_decryption_sequence", (
                f"received a share for {content}, but expected partial_decryption_sequence")
            for shares, value in zip(per_ciphertext, message["value"]):
                shares[self.party_indices[party]] = value
        return [EncodedPlaintext(m, scheme=self) for m in key.decrypt_batch(per_ciphertext)]

    if batched_sequence:
        _set(RefScheme, "_decrypt_sequence_raw", decrypt_sequence_raw)

    Batched, AdditiveVariable = ref_utils.Batched, ref_utils.AdditiveVariable

    def v_batch(cls: Any, candidates: Sequence[tuple[Sequence[int], int, Any, int, int]], index: int,
                correct_param_biprime: int) -> list[Any]:
        """``candidates``: the tuples ``compute_modulus`` builds at distributed_keygen.py:1309-1312,
        ``(g_values, n, prime_candidate_q, p_additive, q_additive)``."""
        vs = _dk.biprime_test_v_calculation_batch(
            [(g, n, p_i, q_i) for (g, n, _, p_i, q_i) in candidates], index, correct_param_biprime, _DEVICE)
        out = []
        for (_, n, _, _, _), v in zip(candidates, vs):
            batched = Batched(AdditiveVariable(label="v", modulus=n), batch_size=correct_param_biprime)
            batched.set_share(index, v)          # :1107, raises like the reference when too few g's
            out.append(batched)
        return out

    def v_one(cls: Any, g_values: list[int], index: int, modulus: int, p_i: int, q_i: int,
              correct_param_biprime: int) -> Any:
        return v_batch(cls, [(g_values, modulus, None, p_i, q_i)], index, correct_param_biprime)[0]

    _set(RefScheme, "_b200_biprime_v_batch", classmethod(v_batch))
    _set(RefScheme, "_DistributedPaillier__biprime_test_v_calculation", classmethod(v_one))


def install_from_env(ref_pkg: Any = None) -> bool:
    """Opt-in switch with unchanged defaults: installs only when ``DKG_B200=1``
    (``DKG_B200_DEVICE`` picks the GPU, ``DKG_B200_SEQUENCE=loop`` keeps the reference's own
    per-ciphertext loops)."""
    if os.environ.get("DKG_B200", "0") != "1":
        return False
    install(ref_pkg, device=int(os.environ.get("DKG_B200_DEVICE", "0")),
            batched_sequence=os.environ.get("DKG_B200_SEQUENCE", "batched") != "loop")
    return True
