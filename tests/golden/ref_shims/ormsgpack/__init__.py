"""Shim: option flags read at import time by the reference (distributed_keygen.py:62-68)."""
OPT_PASSTHROUGH_BIG_INT = 1
OPT_PASSTHROUGH_TUPLE = 2
OPT_PASSTHROUGH_DATACLASS = 4
OPT_SERIALIZE_NUMPY = 8
OPT_NON_STR_KEYS = 16
