"""Shim for tno.mpc.encryption_schemes.shamir: the attributes the hot path reads."""
import math
from typing import Any


class ShamirSecretSharingScheme:
    def __init__(self, modulus: int = 0, number_of_parties: int = 0, polynomial_degree: int = 0) -> None:
        self.modulus = modulus
        self.number_of_parties = number_of_parties
        self.polynomial_degree = polynomial_degree


class ShamirSecretSharingIntegers:
    def __init__(self, kappa: int = 40, max_int: int = 0, number_of_parties: int = 0, polynomial_degree: int = 0) -> None:
        self.kappa = kappa
        self.max_int = max_int
        self.number_of_parties = number_of_parties
        self.polynomial_degree = polynomial_degree


class ShamirShares:
    pass


class IntegerShares:
    def __init__(self, scheme: Any, shares: dict, degree: int, scaling: int) -> None:
        self.scheme = scheme
        self.shares = shares
        self.degree = degree
        self.scaling = scaling
        self.n_fac = math.factorial(scheme.number_of_parties)
