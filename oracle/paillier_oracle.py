"""
CPU restatement of the reference's threshold-Paillier arithmetic.  TEST INFRASTRUCTURE ONLY
(see ``oracle/__init__.py``): never imported by the product package.

All ``file:line`` citations are relative to
``/root/reference/src/tno/mpc/protocols/distributed_keygen/``.

Ground truth for values is CPython ``pow(b, e, m)`` / ``pow(b, -1, m)``; the un-vendored
third-party ``pow_mod`` / ``mod_inv`` (``tno.mpc.encryption_schemes.utils ~=0.10``) return the same
canonical residues.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Iterable, Sequence


def mult_list(list_: Iterable[int], modulus: int | None = None) -> int:
    """Running product, optionally reduced at every step.  Follows ``utils.py:23-38``."""
    out = 1
    if modulus is None:
        for element in list_:
            out = out * element
    else:
        for element in list_:
            out = out * element % modulus
    return out


def pow_mod(base: int, exponent: int, modulus: int) -> int:
    """Third-party ``pow_mod`` (un-vendored): canonical residue of ``base**exponent mod modulus``.
    Negative exponents invert the base first (``ZeroDivisionError``/``ValueError`` if impossible)."""
    return pow(base, exponent, modulus)


def mod_inv(value: int, modulus: int) -> int:
    """Third-party ``mod_inv`` (un-vendored): canonical inverse; raises ``ZeroDivisionError`` when
    ``gcd(value, modulus) != 1`` (gmpy2.invert behaviour)."""
    try:
        return pow(value, -1, modulus)
    except ValueError as exc:  # CPython raises ValueError("base is not invertible ...")
        raise ZeroDivisionError(str(exc)) from exc


@dataclass
class IntegerSharesO:
    """The fields of third-party ``IntegerShares`` that the hot path reads
    (``paillier_shared_key.py:70-85``): ``shares``, ``degree``, ``n_fac``; ``scaling`` is carried
    for fidelity with the key blob (``distributed_keygen.py:943-951``)."""

    shares: dict[int, int]
    degree: int
    scaling: int
    number_of_parties: int
    kappa: int = 40
    max_int: int = 0
    n_fac: int = field(init=False)

    def __post_init__(self) -> None:
        self.n_fac = math.factorial(self.number_of_parties)


class SharedKeyOracle:
    """Restatement of ``PaillierSharedKey`` (``paillier_shared_key.py:25-127``)."""

    def __init__(self, n: int, t: int, player_id: int, share: IntegerSharesO, theta: int) -> None:
        # paillier_shared_key.py:43-50
        self.share = share
        self.n = n
        self.n_square = n * n
        self.t = t
        self.player_id = player_id
        self.theta = theta
        self.theta_inv = mod_inv(theta, n)

    def partial_decrypt_exponent(self) -> int:
        """The signed, per-key exponent of ``partial_decrypt`` (``paillier_shared_key.py:70-85``):
        ``n! * prod(j) * s_i // prod(j - i)`` over ``j in 1..degree+1, j != i`` (floor division)."""
        n_fac = self.share.n_fac
        other_honest_players = [
            i + 1 for i in range(self.share.degree + 1) if i + 1 != self.player_id
        ]
        enumerator = mult_list(other_honest_players)
        denominator = mult_list([(j - self.player_id) for j in other_honest_players])
        return (n_fac * enumerator * self.share.shares[self.player_id]) // denominator

    def partial_decrypt(self, ciphertext_value: int) -> int:
        """``paillier_shared_key.py:86-93`` on the raw ciphertext integer: a negative exponent
        inverts the ciphertext modulo N^2 first, then one modexp."""
        exp = self.partial_decrypt_exponent()
        if exp < 0:
            ciphertext_value = mod_inv(ciphertext_value, self.n_square)
            exp = -exp
        return pow_mod(ciphertext_value, exp, self.n_square)

    def decrypt(self, partial_dict: dict[int, int]) -> int:
        """Share combination, ``paillier_shared_key.py:95-127``: product of the partials of
        parties 1..degree+1 (unreduced, then one ``%``), divisibility check, L-function, times
        theta^-1 mod N."""
        partial_decryptions = [partial_dict[i + 1] for i in range(self.share.degree + 1)]
        if len(partial_decryptions) < self.share.degree + 1:
            raise ValueError("Not enough shares.")
        combined = mult_list(partial_decryptions[: self.share.degree + 1]) % self.n_square
        if (combined - 1) % self.n != 0:
            raise ValueError(
                "Combined decryption minus one is not divisible by N. This might be caused by the "
                "fact that the ciphertext that is being decrypted, differs between the parties."
            )
        return ((combined - 1) // self.n * self.theta_inv) % self.n


def encrypt_raw(n: int, m: int, r: int) -> int:
    """Third-party ``Paillier`` raw encryption + randomisation with g = n + 1
    (``distributed_keygen.py:712``): ``(1 + m n) * r^n mod n^2`` with the
    plaintext taken modulo n (negative plaintexts are ``n - |m|``)."""
    n2 = n * n
    return ((1 + (m % n) * n) % n2) * pow_mod(r, n, n2) % n2


def randomness(n: int, r: int) -> int:
    """The encryption randomness ``r^n mod n^2`` (third-party ``pow_mod(r, n, n_squared)``)."""
    return pow_mod(r, n, n * n)


def jacobi(a: int, n: int) -> int:
    """Jacobi symbol (a/n) for odd n > 0 (what ``sympy.jacobi_symbol`` returns at
    ``distributed_keygen.py:1089``).  Binary algorithm."""
    if n <= 0 or n % 2 == 0:
        raise ValueError("n must be a positive odd integer")
    a %= n
    result = 1
    while a:
        while a % 2 == 0:
            a //= 2
            if n % 8 in (3, 5):
                result = -result
        a, n = n, a
        if a % 4 == 3 and n % 4 == 3:
            result = -result
        a %= n
    return result if n == 1 else 0


def biprime_exponent(index: int, modulus: int, p_i: int, q_i: int) -> int:
    """Exponent of the biprimality-test modexp (``distributed_keygen.py:1092-1097``)."""
    if index == 1:
        return (modulus - p_i - q_i + 1) // 4
    return (p_i + q_i) // 4


def biprime_select_g(g_values: Sequence[int], modulus: int, correct_param_biprime: int) -> list[int]:
    """The g's actually used: first ``correct_param_biprime`` with Jacobi symbol +1
    (``distributed_keygen.py:1084-1091``)."""
    picked: list[int] = []
    for g in g_values:
        if len(picked) == correct_param_biprime:
            break
        if jacobi(g, modulus) != 1:
            continue
        picked.append(g)
    return picked


def biprime_v_calculation(
    g_values: Sequence[int],
    index: int,
    modulus: int,
    p_i: int,
    q_i: int,
    correct_param_biprime: int,
) -> list[int]:
    """``DistributedPaillier.__biprime_test_v_calculation`` (``distributed_keygen.py:1056-1108``):
    the list of v values this party contributes for one candidate modulus."""
    exponent = biprime_exponent(index, modulus, p_i, q_i)
    return [
        int(pow_mod(g, exponent, modulus))
        for g in biprime_select_g(g_values, modulus, correct_param_biprime)
    ]


def biprime_verdict(
    v_by_party: dict[int, Sequence[int]], modulus: int, correct_param_biprime: int
) -> bool:
    """``__biprime_test_with_v_i`` (``distributed_keygen.py:1110-1175``): every one of the first
    ``correct_param_biprime`` tests must satisfy ``v_1 == +-prod_{i>1} v_i (mod N)``; running out of
    tests is a failure."""
    successful = 0
    n_tests = min(len(v) for v in v_by_party.values())
    for k in range(n_tests):
        product = 1
        for key, values in v_by_party.items():
            if key != 1:
                product *= values[k]
        value1 = v_by_party[1][k]
        success = ((value1 % modulus) == (product % modulus)) or (
            (value1 % modulus) == (-product % modulus)
        )
        if not success:
            return False
        successful += 1
        if successful >= correct_param_biprime:
            return True
    return False


def small_prime_divisors_test(prime_list: Iterable[int], modulus: int) -> bool:
    """``__small_prime_divisors_test`` (``distributed_keygen.py:1197-1209``)."""
    for prime in prime_list:
        if modulus % prime == 0:
            return True
    return False
