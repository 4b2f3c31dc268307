"""
ctypes binding of ``libdkg_b200.so`` (C ABI declared in ``include/dkg_b200.h``).

There is deliberately no CPU fallback: if the shared library has not been built
(``python -c "import __graft_entry__ as g; g.build()"``) importing this module raises, and every
compute entry point returns ``DKG_ERR_CUDA`` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DKG_B200_LIB: another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("DKG_B200_LIB") or os.path.join(_HERE, "libdkg_b200.so")

DKG_OK = 0
DKG_ERR_INVALID = 1
DKG_ERR_CUDA = 2
DKG_ERR_UNSUPPORTED = 3
DKG_ERR_NOMEM = 4
DKG_ERR_NOT_IMPLEMENTED = 5

STATUS_OK = 0
STATUS_NOT_INVERTIBLE = 1
STATUS_NOT_DIVISIBLE = 2
STATUS_OUT_OF_RANGE = 3

# every symbol include/dkg_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "dkg_version", "dkg_last_error", "dkg_device_count", "dkg_launch_count",
    "dkg_measure_imad_peak", "dkg_config_set", "dkg_config_get", "dkg_kernel_times",
    "dkg_modexp_ctx_create", "dkg_modexp_ctx_create_nsq", "dkg_modexp_ctx_destroy", "dkg_modexp_ctx_info",
    "dkg_modexp_batch", "dkg_modexp_batch_device",
    "dkg_combine_ctx_create", "dkg_combine_ctx_destroy", "dkg_combine_n2_limbs",
    "dkg_combine_batch", "dkg_combine_batch_device",
    "dkg_threshold_ctx_create", "dkg_threshold_ctx_destroy", "dkg_threshold_info", "dkg_threshold_info_ex", "dkg_threshold_decrypt_batch",
    "dkg_threshold_decrypt_batch_device", "dkg_threshold_partials_batch",
    "dkg_threshold_partial_decrypt_batch", "dkg_threshold_combine_batch", "dkg_host_register", "dkg_host_unregister",
    "dkg_encrypt_batch", "dkg_modexp_grouped",
    "dkg_biprime_v_batch", "dkg_jacobi_batch", "dkg_small_prime_sieve", "dkg_biprime_verdict",
    "dkg_wire_encode_rows", "dkg_wire_decode_rows",
]


class DkgError(RuntimeError):
    def __init__(self, code: int, message: str) -> None:
        super().__init__(f"dkg_b200 error {code}: {message}")
        self.code = code


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA engine has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    c_u32p, c_u8p, c_void = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
    lib.dkg_version.restype = ctypes.c_int
    lib.dkg_last_error.restype = ctypes.c_char_p
    lib.dkg_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    lib.dkg_launch_count.restype = ctypes.c_ulonglong
    lib.dkg_measure_imad_peak.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.dkg_config_set.argtypes = [ctypes.c_char_p, ctypes.c_long]
    lib.dkg_config_get.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_long)]
    lib.dkg_modexp_ctx_create.argtypes = [ctypes.c_int, c_u32p, ctypes.c_int, c_u32p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_void)]
    lib.dkg_modexp_ctx_create_nsq.argtypes = [ctypes.c_int, c_u32p, ctypes.c_int, c_u32p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_void)]
    lib.dkg_modexp_ctx_destroy.argtypes = [c_void]
    lib.dkg_modexp_ctx_destroy.restype = None
    lib.dkg_modexp_ctx_info.argtypes = [c_void, ctypes.POINTER(ctypes.c_int * 12)]
    lib.dkg_modexp_batch.argtypes = [c_void, c_u32p, c_u32p, c_u8p, ctypes.c_size_t]
    lib.dkg_modexp_batch_device.argtypes = [c_void, c_u32p, c_u32p, c_u8p, ctypes.c_size_t, c_void]
    lib.dkg_combine_ctx_create.argtypes = [ctypes.c_int, c_u32p, ctypes.c_int, c_u32p, ctypes.c_int, ctypes.POINTER(c_void)]
    lib.dkg_combine_ctx_destroy.argtypes = [c_void]
    lib.dkg_combine_ctx_destroy.restype = None
    lib.dkg_combine_n2_limbs.argtypes = [c_void]
    lib.dkg_combine_batch.argtypes = [c_void, c_u32p, c_u32p, c_u8p, ctypes.c_size_t]
    lib.dkg_combine_batch_device.argtypes = [c_void, c_u32p, c_u32p, c_u8p, ctypes.c_size_t, c_void]
    lib.dkg_threshold_ctx_create.argtypes = [c_void, ctypes.c_int, c_u32p, ctypes.c_int, c_u32p, ctypes.c_int, c_u32p, ctypes.c_int,
                                             c_u8p, ctypes.POINTER(c_void)]
    lib.dkg_threshold_ctx_destroy.argtypes = [c_void]
    lib.dkg_threshold_ctx_destroy.restype = None
    lib.dkg_threshold_info.argtypes = [c_void, ctypes.POINTER(ctypes.c_int * 4)]
    lib.dkg_threshold_decrypt_batch.argtypes = [c_void, c_u32p, c_u32p, c_u32p, c_u8p, ctypes.c_size_t]
    lib.dkg_kernel_times.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double * 4096), ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.dkg_threshold_info_ex.argtypes = [c_void, ctypes.POINTER(ctypes.c_int * 8)]
    lib.dkg_threshold_decrypt_batch_device.argtypes = [c_void, c_void, c_void, c_void, c_void, ctypes.c_size_t, c_void]
    lib.dkg_threshold_partials_batch.argtypes = [c_void, c_u32p, c_u32p, c_u8p, ctypes.c_size_t]
    lib.dkg_threshold_partial_decrypt_batch.argtypes = [c_void, ctypes.c_int, c_u32p, c_u32p, c_u8p, ctypes.c_size_t]
    lib.dkg_threshold_combine_batch.argtypes = [c_void, c_u32p, c_u32p, c_u8p, ctypes.c_size_t]
    lib.dkg_host_register.argtypes = [c_void, ctypes.c_size_t]
    lib.dkg_host_unregister.argtypes = [c_void]
    lib.dkg_encrypt_batch.argtypes = [c_void, c_u32p, ctypes.c_int, c_u32p, c_u32p, c_u32p, ctypes.c_size_t]
    lib.dkg_modexp_grouped.argtypes = [ctypes.c_int, c_u32p, c_u32p, ctypes.c_int, c_u32p, c_u32p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
    lib.dkg_biprime_v_batch.argtypes = [ctypes.c_int, c_u32p, c_u32p, ctypes.c_int, c_u32p, ctypes.c_int, ctypes.c_int,
                                        c_u32p, c_void, ctypes.c_size_t, ctypes.c_int]
    lib.dkg_jacobi_batch.argtypes = [ctypes.c_int, c_u32p, c_u32p, ctypes.c_int, c_void, ctypes.c_size_t, ctypes.c_int]
    lib.dkg_small_prime_sieve.argtypes = [ctypes.c_int, c_u32p, c_u32p, ctypes.c_int, c_u8p, ctypes.c_size_t, ctypes.c_int]
    lib.dkg_biprime_verdict.argtypes = [ctypes.c_int, c_u32p, c_u32p, ctypes.c_int, ctypes.c_int, c_u8p, ctypes.c_size_t, ctypes.c_int]
    c_sizep = ctypes.POINTER(ctypes.c_size_t)
    lib.dkg_wire_encode_rows.argtypes = [c_u32p, ctypes.c_size_t, ctypes.c_int, c_u8p, ctypes.c_size_t, c_sizep]
    lib.dkg_wire_decode_rows.argtypes = [c_u8p, ctypes.c_size_t, ctypes.c_int, c_u32p, ctypes.c_size_t, c_sizep, c_sizep]
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != DKG_OK:
        msg = lib.dkg_last_error()
        raise DkgError(rc, msg.decode() if msg else "")


def config_set(key: str, value: int) -> None:
    check(lib.dkg_config_set(key.encode(), int(value)))


def config_get(key: str) -> int:
    v = ctypes.c_long(0)
    check(lib.dkg_config_get(key.encode(), ctypes.byref(v)))
    return int(v.value)


def kernel_times(device: int = 0, capacity: int = 4096) -> list[float]:
    """Durations (ms) of the exponentiation-kernel launches recorded since the last call
    (``config_set("time_kernels", 1)`` turns the recording on)."""
    capacity = 4096
    buf = (ctypes.c_double * capacity)()
    n = ctypes.c_int(0)
    check(lib.dkg_kernel_times(device, buf, capacity, ctypes.byref(n)))
    return [float(buf[i]) for i in range(n.value)]


def device_count() -> int:
    n = ctypes.c_int(0)
    rc = lib.dkg_device_count(ctypes.byref(n))
    return n.value if rc == DKG_OK else 0
