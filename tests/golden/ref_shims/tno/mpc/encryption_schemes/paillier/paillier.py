"""Shim for tno.mpc.encryption_schemes.paillier.paillier: data holders only."""
from typing import Any, TypedDict, Union

Plaintext = Union[int, float]


class PaillierPublicKey:
    def __init__(self, n: int, g: int) -> None:
        self.n = n
        self.g = g
        self.n_squared = n * n


class PaillierSecretKey:
    pass


class Paillier:
    class SerializedPaillier(TypedDict):
        """Name only: subclassed by the reference at distributed_keygen.py:1588."""

        prec: int

    def __init__(self, public_key: Any = None, secret_key: Any = None, precision: int = 0,
                 share_secret_key: bool = False, **_kwargs: Any) -> None:
        self.public_key = public_key
        self.secret_key = secret_key
        self.precision = precision
        self.share_secret_key = share_secret_key


class PaillierCiphertext:
    """Holds the raw integer; ``get_value`` is what partial_decrypt reads
    (paillier_shared_key.py:69)."""

    def __init__(self, raw_value: int, scheme: Any) -> None:
        self._raw_value = raw_value
        self.scheme = scheme

    def get_value(self) -> int:
        return self._raw_value

    def peek_value(self) -> int:
        return self._raw_value
