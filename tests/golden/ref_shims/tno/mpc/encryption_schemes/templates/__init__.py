"""Shim for tno.mpc.encryption_schemes.templates."""
from tno.mpc.encryption_schemes.templates.encryption_scheme import EncodedPlaintext  # noqa: F401


class SecretKey:
    def __init__(self) -> None:
        pass


class SerializationError(Exception):
    pass
