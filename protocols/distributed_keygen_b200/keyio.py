"""
Reader for the reference's stored-key format (what ``DistributedPaillier.store_private_key``
writes, ``distributed_keygen.py:1511-1537``, and ``load_private_key_from_bytes`` reads,
``:1539-1586``): an (or)msgpack map whose big integers are tagged
``{"type": "int", "data": <little-endian two's-complement bytes>}``.  The 24 golden fixtures of the
reference (``test/test_data/*.obj``) are in this format.  Only the key material is rebuilt; the
session/pool renegotiation of the reference's loader is protocol orchestration and out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import msgpack

from .paillier_shared_key import IntegerShares, PaillierSharedKey


def _decode(obj: Any) -> Any:
    if isinstance(obj, dict):
        if obj.get("type") == "int" and isinstance(obj.get("data"), (bytes, bytearray)):
            return int.from_bytes(obj["data"], "little", signed=True)
        return {k: _decode(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_decode(v) for v in obj]
    return obj


@dataclass
class StoredKey:
    n: int
    g: int
    precision: int
    index: int
    party_indices: dict[str, int]
    corruption_threshold: int
    secret_key: PaillierSharedKey


def load_private_key_from_bytes(blob: bytes, device: int = 0) -> StoredKey:
    raw = msgpack.unpackb(blob, raw=False, strict_map_key=False)
    obj = _decode(raw["object"])
    pub = obj["pub_key"]["data"]
    priv = obj["priv_key"]["data"]
    sh = priv["share"]["data"]
    share = IntegerShares(
        shares={int(k): v for k, v in sh["shares"].items()},
        degree=sh["degree"],
        scaling=sh["scaling"],
        number_of_parties=sh["scheme"]["number_of_parties"],
    )
    key = PaillierSharedKey(priv["n"], priv["t"], priv["player_id"], share, priv["theta"], device=device)
    return StoredKey(
        n=pub["n"], g=pub["g"], precision=obj["precision"], index=obj["index"],
        party_indices=dict(obj["party_indices"]), corruption_threshold=obj["corruption_threshold"],
        secret_key=key,
    )
