"""Shim for tno.mpc.communication: names only (no networking is exercised)."""


class RepetitionError(Exception):
    pass


class Serialization:
    @staticmethod
    def register_class(*_args, **_kwargs):
        return None


class SupportsSerialization:
    pass


class Pool:
    pass
