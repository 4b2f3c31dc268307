"""
Host-side mirror of the reference's ``PaillierSharedKey`` (``paillier_shared_key.py:25-127`` in
tno.mpc.protocols.distributed_keygen v4.2.2) with the two arithmetic methods re-routed to the
B200 engine, plus their batched forms -- the calls that replace the per-ciphertext loops of
``DistributedPaillier._decrypt_sequence_raw`` (``distributed_keygen.py:463-466`` and ``:510-515``).

Same names, argument meaning and error behaviour as the reference:

* ``partial_decrypt(ciphertext)`` -> int; ``TypeError`` if the argument is not a ciphertext object
  (anything with ``get_value()`` and ``scheme.public_key.n``), ``ValueError`` if it was encrypted
  under a different key; a bare ``int`` is accepted as the raw ciphertext value.
* ``decrypt(partial_dict)`` -> int; ``KeyError`` when a party in 1..degree+1 is missing,
  ``ValueError`` when the combined value minus one is not divisible by N.
* a non-invertible ciphertext under a negative exponent raises ``ZeroDivisionError`` (what the
  third-party ``mod_inv`` raises).

All modular arithmetic happens on the GPU; the only big-integer work left on the host is the
per-key exponent (``:70-85``, computed once per key instead of once per ciphertext) and
int <-> limb conversion.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Iterable, Mapping, Sequence

import numpy as np

from .engine import CombineContext, ModexpContext
from .limbs import ints_to_limbs, limbs_for_bits, limbs_to_ints


def mult_list(list_: Iterable[int], modulus: int | None = None) -> int:
    """Mirror of ``utils.mult_list`` (``utils.py:23-38``); only used on small per-key integers."""
    out = 1
    for element in list_:
        out = out * element if modulus is None else out * element % modulus
    return out


@dataclass
class IntegerShares:
    """The attributes of third-party ``IntegerShares`` that the key object reads."""

    shares: dict[int, int]
    degree: int
    scaling: int
    number_of_parties: int
    n_fac: int = field(init=False)

    def __post_init__(self) -> None:
        self.n_fac = math.factorial(self.number_of_parties)


class PaillierSharedKey:
    """Shared Paillier secret key whose arithmetic runs on the B200 engine."""

    def __init__(self, n: int, t: int, player_id: int, share: Any, theta: int, device: int = 0) -> None:
        # paillier_shared_key.py:43-50
        self.share = share
        self.n = n
        self.n_square = n * n
        self.t = t
        self.player_id = player_id
        self.theta = theta
        self.device = device
        try:
            self.theta_inv = pow(theta, -1, n)
        except ValueError as exc:
            raise ZeroDivisionError(str(exc)) from exc
        self._modexp: ModexpContext | None = None
        self._combine: CombineContext | None = None

    # -- per-key constants ---------------------------------------------------------------------
    def partial_decrypt_exponent(self) -> int:
        """Signed exponent of this party's partial decryption (``:70-85``): Lagrange coefficient
        for the reconstruction set {1..degree+1} folded into n! * s_i, with floor division."""
        n_fac = self.share.n_fac
        others = [i + 1 for i in range(self.share.degree + 1) if i + 1 != self.player_id]
        enumerator = mult_list(others)
        denominator = mult_list([(j - self.player_id) for j in others])
        return (n_fac * enumerator * self.share.shares[self.player_id]) // denominator

    def _modexp_ctx(self) -> ModexpContext:
        if self._modexp is None:
            self._modexp = ModexpContext(self.n_square, self.partial_decrypt_exponent(), self.device, root=self.n)
        return self._modexp

    def _combine_ctx(self) -> CombineContext:
        if self._combine is None:
            self._combine = CombineContext(self.n, self.theta_inv, self.share.degree + 1, self.device)
        return self._combine

    def close(self) -> None:
        for ctx in (self._modexp, self._combine):
            if ctx is not None:
                ctx.close()
        self._modexp = self._combine = None

    # -- reference-compatible scalar API ---------------------------------------------------------
    def _raw_value(self, ciphertext: Any) -> int:
        if isinstance(ciphertext, int):
            return ciphertext
        if not hasattr(ciphertext, "get_value") or not hasattr(ciphertext, "scheme"):
            raise TypeError(
                f"Expected ciphertext to be a PaillierCiphertext not: {type(ciphertext)}"
            )
        if self.n != ciphertext.scheme.public_key.n:
            raise ValueError("encrypted against a different key!")
        return int(ciphertext.get_value())

    def partial_decrypt(self, ciphertext: Any) -> int:
        return self.partial_decrypt_batch([ciphertext])[0]

    def decrypt(self, partial_dict: Mapping[int, int]) -> int:
        return self.decrypt_batch([partial_dict])[0]

    # -- batched API (what replaces the loops in _decrypt_sequence_raw) -------------------------
    def partial_decrypt_batch(self, ciphertexts: Sequence[Any]) -> list[int]:
        values = [self._raw_value(c) for c in ciphertexts]
        return self._modexp_ctx().modexp(values)

    def partial_decrypt_limbs(self, values: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        """uint32 [count, limbs(N^2)] in, (partials, status) out: no Python-int conversion."""
        return self._modexp_ctx().modexp_limbs(values)

    def decrypt_batch(self, partial_dicts: Sequence[Mapping[int, int]]) -> list[int]:
        shares = self.share.degree + 1
        ctx = self._combine_ctx()
        # paillier_shared_key.py:108-110: indexing raises KeyError for a missing party
        rows = [[d[i + 1] for d in partial_dicts] for i in range(shares)]
        arr = np.stack([ints_to_limbs(r, ctx.n2_limbs) for r in rows]) if partial_dicts else np.zeros(
            (shares, 0, ctx.n2_limbs), dtype=np.uint32
        )
        out, status = ctx.combine_limbs(arr)
        if status.any():
            raise ValueError(
                "Combined decryption minus one is not divisible by N. This might be caused by the "
                "fact that the ciphertext that is being decrypted, differs between the parties."
            )
        return limbs_to_ints(out)

    def decrypt_limbs(self, partials: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        """uint32 [degree+1, count, limbs(N^2)] in, (plaintexts [count, limbs(N)], status) out."""
        return self._combine_ctx().combine_limbs(partials)

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, PaillierSharedKey):
            raise TypeError(f"Expected comparison with another PaillierSharedKey, not {type(other)}")
        return (
            self.share == other.share and self.n == other.n and self.t == other.t
            and self.player_id == other.player_id and self.theta == other.theta
        )

    def __str__(self) -> str:
        return str({"priv_shared_key": {"n": self.n, "t": self.t, "player_id": self.player_id,
                                        "theta": self.theta, "share": self.share}})


def n_square_limbs(n: int) -> int:
    return limbs_for_bits((n * n).bit_length())
