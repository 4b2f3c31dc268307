"""Shim for tno.mpc.encryption_schemes.utils (~=0.10) without gmpy2: CPython pow."""


def pow_mod(base: int, exponent: int, modulus: int) -> int:
    return pow(base, exponent, modulus)


def mod_inv(value: int, modulus: int) -> int:
    try:
        return pow(value, -1, modulus)
    except ValueError as exc:
        raise ZeroDivisionError(str(exc)) from exc
