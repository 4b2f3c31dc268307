"""Wire format of the partial-decryption message: the vectorised limb-row
codec must produce exactly the bytes a msgpack serializer with the reference's big-integer tagging
produces, and that tagging is pinned on the bytes of the reference's own key fixtures."""
import base64
import json
import os
import random

import msgpack
import numpy as np
import pytest

from protocols.distributed_keygen_b200 import wire
from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

HERE = os.path.dirname(os.path.abspath(__file__))


def tagged(v: int):
    if v.bit_length() <= 64:   # native msgpack integers cover [0, 2^64) (uint64, 0xcf)
        return v
    return {"type": "int", "data": v.to_bytes((v.bit_length() + 8) // 8, "little", signed=True)}


def generic_message(values):
    return msgpack.packb({"content": "partial_decryption_sequence", "value": [tagged(v) for v in values]},
                         use_bin_type=True)


def test_int_tagging_matches_reference_fixture_bytes():
    fx = json.load(open(os.path.join(HERE, "golden", "fixture_vectors.json")))
    checked = 0
    for entry in [k for s in fx["sets"] for k in s["keys"]]:
        blob = base64.b64decode(entry["blob_b64"])
        raw = msgpack.unpackb(blob, raw=False, strict_map_key=False)

        def walk(o):
            nonlocal checked
            if isinstance(o, dict):
                if o.get("type") == "int":
                    v = int.from_bytes(o["data"], "little", signed=True)
                    if v >= 1 << 64:
                        limbs = (v.bit_length() + 31) // 32
                        enc = wire.encode_int_rows(ints_to_limbs([v], limbs))
                        assert enc[0] == 0x91
                        assert enc[1:] in blob, "tagged big integer bytes differ from the reference blob"
                        checked += 1
                    return
                for x in o.values():
                    walk(x)
            elif isinstance(o, list):
                for x in o:
                    walk(x)

        walk(raw)
    assert checked >= 24 * 3


@pytest.mark.parametrize("count", [0, 1, 15, 16, 300])
def test_message_bytes_equal_generic_msgpack(count):
    rng = random.Random(count)
    limbs = 128
    values = []
    for i in range(count):
        bits = rng.choice([64, 65, 71, 72, 2040, 2047, 2048, 4000, 4088, 4095, 4096])
        values.append(rng.getrandbits(bits) | (1 << (bits - 1)))
    rows = ints_to_limbs(values, limbs)
    body = wire.pack_partial_decryption_message(rows)
    assert body == generic_message(values)
    back = wire.unpack_partial_decryption_message(body, limbs)
    assert back.dtype == np.uint32 and back.shape == (count, limbs)
    assert limbs_to_ints(back) == values if count else back.size == 0


def test_small_values_take_native_integers():
    values = [0, 1, 127, 128, 2**32, 2**63 - 1, 2**63 + 5, 2**64 - 1, 2**64, 2**200 + 3]
    rows = ints_to_limbs(values, 8)
    body = wire.pack_partial_decryption_message(rows)
    assert body == generic_message(values)
    assert limbs_to_ints(wire.unpack_partial_decryption_message(body, 8)) == values


def test_array32_header_and_large_batch():
    rng = np.random.default_rng(3)
    rows = rng.integers(0, 2**32, size=(70000, 5), dtype=np.uint32)
    rows[:, -1] |= 1
    body = wire.pack_partial_decryption_message(rows)
    assert body[len(wire._MSG_HEAD)] == 0xDD
    assert np.array_equal(wire.unpack_partial_decryption_message(body, 5), rows)
    sample = limbs_to_ints(rows[:50])
    assert msgpack.unpackb(body, raw=False)["value"][7]["data"] == sample[7].to_bytes(
        (sample[7].bit_length() + 8) // 8, "little", signed=True)


def test_decode_errors():
    neg = msgpack.packb({"content": "partial_decryption_sequence",
                         "value": [{"type": "int", "data": (-(2**100)).to_bytes(14, "little", signed=True)}]},
                        use_bin_type=True)
    with pytest.raises(ValueError):
        wire.unpack_partial_decryption_message(neg, 8)
    wide = generic_message([2**300])
    with pytest.raises(ValueError):
        wire.unpack_partial_decryption_message(wide, 8)
    other = msgpack.packb({"content": "something_else", "value": []}, use_bin_type=True)
    with pytest.raises(AssertionError, match="expected partial_decryption_sequence"):
        wire.unpack_partial_decryption_message(other, 8)
    good = generic_message([2**100])
    with pytest.raises(ValueError):
        wire.unpack_partial_decryption_message(good + b"\x00", 8)
    # an array header that promises more elements than the message has bytes (peer input) must not
    # size an allocation: ValueError, not MemoryError
    with pytest.raises(ValueError):
        wire.decode_int_rows(b"\xdd\xff\xff\xff\xff", 8)
    with pytest.raises(ValueError):
        wire.decode_int_rows(b"\xdc\xff\xff\x01", 8)
