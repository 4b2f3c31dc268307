"""
B200-native batched modular-exponentiation engine for the data-parallel hot path of
``tno.mpc.protocols.distributed_keygen`` (threshold Paillier): partial decryption, share
combination, encryption randomness and the biprimality-test exponentiations.

The arithmetic runs in hand-written sm_100a CUDA kernels behind a C ABI
(``include/dkg_b200.h`` -> ``libdkg_b200.so``); this package is the Python host side that mirrors
the reference's interface for that path.  Importing it requires the built shared library; there is
no CPU fallback.
"""
from . import _native  # noqa: F401  (fails loudly if the CUDA engine is not built)
from .engine import (  # noqa: F401
    CombineContext, EncryptContext, ModexpContext, ThresholdContext, pinned, biprime_v_batch_limbs, biprime_verdict, jacobi_batch, launch_count,
    modexp_grouped, modexp_grouped_limbs, small_prime_sieve,
)
from . import distributed_keygen  # noqa: F401
from .paillier_shared_key import IntegerShares, PaillierSharedKey  # noqa: F401

__all__ = ["ThresholdContext", "pinned", "CombineContext", "EncryptContext", "ModexpContext", "PaillierSharedKey", "IntegerShares", "launch_count", "modexp_grouped",
           "modexp_grouped_limbs", "distributed_keygen"]
