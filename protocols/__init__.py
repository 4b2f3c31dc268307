"""Namespace package for the B200-native reimplementation of the hot path of
``tno.mpc.protocols.distributed_keygen``; the code lives in ``protocols.distributed_keygen_b200``."""
