"""
Key material for the oracle.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

* ``load_key_blob`` decodes the reference's on-disk key format (what ``store_private_key`` writes,
  ``distributed_keygen.py:1511-1537``; the 24 golden fixtures under ``test/test_data/*.obj``).
* ``dealer_keygen`` simulates, with a trusted dealer, the *shape* of a key the distributed protocol
  produces (``distributed_keygen.py:869-875, 1192-1195, 1418-1499``):
  additive prime shares, N = p q with p = q = 3 (mod 4), integer-Shamir sharings of lambda and beta
  with P! scaling, secret-key share = pointwise product (degree 2t, scaling (P!)^2), theta.
"""

from __future__ import annotations

import math
import random
from dataclasses import dataclass
from typing import Any

import msgpack

from .paillier_oracle import IntegerSharesO, SharedKeyOracle


def _decode(obj: Any) -> Any:
    """Undo the communication module's tagging: ``{"type": "int", "data": <LE two's complement>}``."""
    if isinstance(obj, dict):
        if obj.get("type") == "int" and isinstance(obj.get("data"), (bytes, bytearray)):
            return int.from_bytes(obj["data"], "little", signed=True)
        return {k: _decode(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_decode(v) for v in obj]
    return obj


def load_key_blob(blob: bytes) -> dict[str, Any]:
    """Decode one stored key (``distributed_keygen.py:1539-1586`` reads the same structure)."""
    raw = msgpack.unpackb(blob, raw=False, strict_map_key=False)
    return _decode(raw["object"])


def key_from_blob(blob: bytes) -> SharedKeyOracle:
    obj = load_key_blob(blob)
    priv = obj["priv_key"]["data"]
    sh = priv["share"]["data"]
    share = IntegerSharesO(
        shares={int(k): v for k, v in sh["shares"].items()},
        degree=sh["degree"],
        scaling=sh["scaling"],
        number_of_parties=sh["scheme"]["number_of_parties"],
        kappa=sh["scheme"]["kappa"],
        max_int=sh["scheme"]["max_int"],
    )
    return SharedKeyOracle(
        n=priv["n"], t=priv["t"], player_id=priv["player_id"], share=share, theta=priv["theta"]
    )


# ---------------------------------------------------------------------------------------------
# dealer-simulated, reference-shaped keys
# ---------------------------------------------------------------------------------------------


def _is_probable_prime(n: int, rng: random.Random, rounds: int = 24) -> bool:
    if n < 2:
        return False
    for sp in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71):
        if n % sp == 0:
            return n == sp
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for _ in range(rounds):
        a = rng.randrange(2, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def prime_candidate_share(index: int, prime_length: int, rng: random.Random) -> int:
    """``_generate_prime_candidate`` (``distributed_keygen.py:855-876``): party 1 draws 3 (mod 4),
    the others 0 (mod 4), all with the top bit set."""
    mod4 = 3 if index == 1 else 0
    return 2 ** (prime_length - 1) + (rng.getrandbits(prime_length - 3) << 2) + mod4


@dataclass
class DealerKey:
    n: int
    parties: int
    t: int
    p_shares: list[int]  # additive shares of p, party 1 first
    q_shares: list[int]
    keys: dict[int, SharedKeyOracle]  # player_id -> key
    theta: int

    @property
    def p(self) -> int:
        return sum(self.p_shares)

    @property
    def q(self) -> int:
        return sum(self.q_shares)


def _additive_prime(parties: int, prime_length: int, rng: random.Random) -> list[int]:
    while True:
        shares = [prime_candidate_share(i + 1, prime_length, rng) for i in range(parties)]
        if _is_probable_prime(sum(shares), rng):
            return shares


def _exact_prime(bits: int, parties: int, rng: random.Random) -> list[int]:
    """A prime of exactly ``bits`` bits, = 3 (mod 4), with the two top bits set (so that p*q has
    exactly 2*bits bits), split into additive shares with the reference's residues mod 4."""
    while True:
        p = (3 << (bits - 2)) | (rng.getrandbits(bits - 2) & ~3) | 3
        if _is_probable_prime(p, rng):
            break
    others = [rng.getrandbits(bits - 4) << 2 for _ in range(parties - 1)]
    return [p - sum(others)] + others


def _int_shamir_share(
    secret: int, parties: int, t: int, kappa: int, max_int: int, rng: random.Random
) -> dict[int, int]:
    """Integer Shamir sharing as the un-vendored ``ShamirSecretSharingIntegers`` produces it
    ([memory] for the exact coefficient range): f(X) = P!*secret +
    sum_k a_k X^k, a_k uniform in (-2^kappa (P!)^2 max_int, +2^kappa (P!)^2 max_int)."""
    n_fac = math.factorial(parties)
    bound = (2**kappa) * n_fac * n_fac * max_int
    coeffs = [n_fac * secret] + [rng.randrange(-bound + 1, bound) for _ in range(t)]
    return {i: sum(c * i**k for k, c in enumerate(coeffs)) for i in range(1, parties + 1)}


def dealer_keygen(
    key_length: int,
    parties: int,
    t: int,
    seed: int,
    kappa: int = 40,
    exact: bool = False,
) -> DealerKey:
    """Reference-shaped threshold key from a trusted dealer (all secrets in one place: only for
    tests and benchmarks).  ``exact=True`` gives an N of exactly ``key_length`` bits instead of the
    reference's ``key_length + 2*log2(P)`` bits."""
    rng = random.Random(seed)
    prime_length = key_length // 2
    if exact:
        p_shares = _exact_prime(prime_length, parties, rng)
        q_shares = _exact_prime(prime_length, parties, rng)
    else:
        p_shares = _additive_prime(parties, prime_length, rng)
        q_shares = _additive_prime(parties, prime_length, rng)
    p, q = sum(p_shares), sum(q_shares)
    n = p * q
    # distributed_keygen.py:1192-1195: lambda = N - p - q + 1 held additively
    lam = n - p - q + 1
    n_fac = math.factorial(parties)
    while True:
        # distributed_keygen.py:1449: every party draws beta_i = randbelow(N)
        beta = sum(rng.randrange(n) for _ in range(parties))
        # distributed_keygen.py:1486-1489
        theta = (lam * beta % n) * n_fac**3 % n
        if math.gcd(theta, n) == 1:
            break
    lam_shares = _int_shamir_share(lam, parties, t, kappa, n, rng)
    beta_shares = _int_shamir_share(beta, parties, t, kappa, n, rng)
    keys = {}
    for i in range(1, parties + 1):
        # distributed_keygen.py:1465: secret_key_sharing = lambda_ * beta (pointwise, degree 2t)
        share = IntegerSharesO(
            shares={i: lam_shares[i] * beta_shares[i]},
            degree=2 * t,
            scaling=n_fac * n_fac,
            number_of_parties=parties,
            kappa=kappa,
            max_int=n,
        )
        keys[i] = SharedKeyOracle(n=n, t=t, player_id=i, share=share, theta=theta)
    return DealerKey(
        n=n, parties=parties, t=t, p_shares=p_shares, q_shares=q_shares, keys=keys, theta=theta
    )


def dealer_key_to_json(dk: DealerKey) -> dict[str, Any]:
    return {
        "n": hex(dk.n),
        "parties": dk.parties,
        "t": dk.t,
        "theta": hex(dk.theta),
        "p_shares": [hex(x) for x in dk.p_shares],
        "q_shares": [hex(x) for x in dk.q_shares],
        "kappa": next(iter(dk.keys.values())).share.kappa,
        "shares": {str(i): hex(k.share.shares[i]) for i, k in dk.keys.items()},
    }


def _unhex(s: str) -> int:
    return int(s, 16)


def dealer_key_from_json(d: dict[str, Any]) -> DealerKey:
    n, parties, t = _unhex(d["n"]), d["parties"], d["t"]
    n_fac = math.factorial(parties)
    theta = _unhex(d["theta"])
    keys = {}
    for i_s, s in d["shares"].items():
        i = int(i_s)
        share = IntegerSharesO(
            shares={i: _unhex(s)},
            degree=2 * t,
            scaling=n_fac * n_fac,
            number_of_parties=parties,
            kappa=d.get("kappa", 40),
            max_int=n,
        )
        keys[i] = SharedKeyOracle(n=n, t=t, player_id=i, share=share, theta=theta)
    return DealerKey(
        n=n,
        parties=parties,
        t=t,
        p_shares=[_unhex(x) for x in d["p_shares"]],
        q_shares=[_unhex(x) for x in d["q_shares"]],
        keys=keys,
        theta=theta,
    )
