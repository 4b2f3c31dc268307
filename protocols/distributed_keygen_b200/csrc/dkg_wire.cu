// Host-side codec for the batched partial-decryption message: limb rows
// <-> the msgpack list of integers the reference broadcasts (distributed_keygen.py:476-484, read
// back at :497-505).  Integers that do not fit 64 bits are written the way the reference's
// serializer tags them, {"type": "int", "data": <little-endian two's complement, (bits+8)/8 bytes>};
// smaller ones as minimal native msgpack integers.  No device work here: this is the packing that
// otherwise costs one Python int per element on either side of the GPU call.
#include <stdint.h>
#include <string.h>

#include "../../../include/dkg_b200.h"
#include "dkg_host_bigint.h"

void dkg_set_error(const char* msg);  // dkg_engine.cu

namespace {

const uint8_t kIntTag[] = {0x82, 0xa4, 't', 'y', 'p', 'e', 0xa3, 'i', 'n', 't', 0xa4, 'd', 'a', 't', 'a'};
const size_t kIntTagLen = sizeof(kIntTag);

inline size_t array_header(uint8_t* p, size_t count) {
  if (count < 16) {
    if (p) p[0] = (uint8_t)(0x90 + count);
    return 1;
  }
  if (count < (1u << 16)) {
    if (p) {
      p[0] = 0xdc;
      p[1] = (uint8_t)(count >> 8);
      p[2] = (uint8_t)count;
    }
    return 3;
  }
  if (p) {
    p[0] = 0xdd;
    for (int k = 0; k < 4; ++k) p[1 + k] = (uint8_t)(count >> (8 * (3 - k)));
  }
  return 5;
}

inline size_t put_be(uint8_t* p, uint8_t tag, uint64_t v, int bytes) {
  p[0] = tag;
  for (int k = 0; k < bytes; ++k) p[1 + k] = (uint8_t)(v >> (8 * (bytes - 1 - k)));
  return 1 + (size_t)bytes;
}

// one integer; p may be null (size query)
inline size_t encode_row(uint8_t* p, const uint32_t* row, int limbs) {
  const int bits = dkg_host::bit_length(row, limbs);
  if (bits <= 64) {   // native msgpack integers cover the whole uint64 range (0xcf)
    const uint64_t v = (uint64_t)row[0] | (limbs > 1 ? (uint64_t)row[1] << 32 : 0);
    if (v < 128) {
      if (p) p[0] = (uint8_t)v;
      return 1;
    }
    const int bytes = v < (1u << 8) ? 1 : v < (1u << 16) ? 2 : v < (1ull << 32) ? 4 : 8;
    if (p) put_be(p, (uint8_t)(bytes == 1 ? 0xcc : bytes == 2 ? 0xcd : bytes == 4 ? 0xce : 0xcf), v, bytes);
    return 1 + (size_t)bytes;
  }
  const size_t nbytes = ((size_t)bits + 8) / 8;
  const size_t hdr = nbytes < 256 ? 2 : nbytes < 65536 ? 3 : 5;
  if (p) {
    memcpy(p, kIntTag, kIntTagLen);
    p += kIntTagLen;
    if (hdr == 2) put_be(p, 0xc4, nbytes, 1);
    else if (hdr == 3) put_be(p, 0xc5, nbytes, 2);
    else put_be(p, 0xc6, nbytes, 4);
    p += hdr;
    const size_t avail = 4 * (size_t)limbs;
    const size_t body = nbytes < avail ? nbytes : avail;
    memcpy(p, row, body);  // little-endian host, little-endian limbs
    if (nbytes > avail) p[avail] = 0;  // sign byte of a value with a full top limb
  }
  return kIntTagLen + hdr + nbytes;
}

}  // namespace

extern "C" int dkg_wire_encode_rows(const uint32_t* rows, size_t count, int limbs, uint8_t* out,
                                    size_t capacity, size_t* written) {
  if (!rows && count) { dkg_set_error("dkg_wire_encode_rows: null rows"); return DKG_ERR_INVALID; }
  if (limbs <= 0 || !written) { dkg_set_error("dkg_wire_encode_rows: bad arguments"); return DKG_ERR_INVALID; }
  size_t need = array_header(nullptr, count);
  for (size_t i = 0; i < count; ++i) need += encode_row(nullptr, rows + i * (size_t)limbs, limbs);
  *written = need;
  if (!out) return DKG_OK;  // size query
  if (capacity < need) { dkg_set_error("dkg_wire_encode_rows: output buffer too small"); return DKG_ERR_NOMEM; }
  uint8_t* p = out + array_header(out, count);
  for (size_t i = 0; i < count; ++i) p += encode_row(p, rows + i * (size_t)limbs, limbs);
  return DKG_OK;
}

extern "C" int dkg_wire_decode_rows(const uint8_t* buf, size_t len, int limbs, uint32_t* rows,
                                    size_t capacity_rows, size_t* count_out, size_t* consumed) {
  if (!buf || limbs <= 0 || !count_out) { dkg_set_error("dkg_wire_decode_rows: bad arguments"); return DKG_ERR_INVALID; }
  size_t pos = 0, count = 0;
  auto need = [&](size_t n) { return pos + n <= len; };
  if (!need(1)) { dkg_set_error("dkg_wire_decode_rows: truncated"); return DKG_ERR_INVALID; }
  const uint8_t b0 = buf[0];
  if (b0 >= 0x90 && b0 <= 0x9f) { count = b0 - 0x90; pos = 1; }
  else if (b0 == 0xdc && len >= 3) { count = ((size_t)buf[1] << 8) | buf[2]; pos = 3; }
  else if (b0 == 0xdd && len >= 5) { for (int k = 0; k < 4; ++k) count = (count << 8) | buf[1 + k]; pos = 5; }
  else { dkg_set_error("dkg_wire_decode_rows: not a msgpack array"); return DKG_ERR_INVALID; }
  // every element takes at least one byte: a header that promises more than the buffer holds is
  // malformed (and must not size an allocation on the caller's side)
  if (count > len - pos) { dkg_set_error("dkg_wire_decode_rows: array header longer than the message"); return DKG_ERR_INVALID; }
  *count_out = count;
  if (!rows) { if (consumed) *consumed = 0; return DKG_OK; }  // count query
  if (capacity_rows < count) { dkg_set_error("dkg_wire_decode_rows: row buffer too small"); return DKG_ERR_NOMEM; }
  const size_t width = 4 * (size_t)limbs;
  for (size_t i = 0; i < count; ++i) {
    uint8_t* dst = (uint8_t*)(rows + i * (size_t)limbs);
    memset(dst, 0, width);
    if (!need(1)) { dkg_set_error("dkg_wire_decode_rows: truncated"); return DKG_ERR_INVALID; }
    const uint8_t t = buf[pos];
    if (need(kIntTagLen + 2) && memcmp(buf + pos, kIntTag, kIntTagLen) == 0) {
      pos += kIntTagLen;
      const uint8_t bt = buf[pos];
      size_t n = 0;
      int hb = bt == 0xc4 ? 1 : bt == 0xc5 ? 2 : bt == 0xc6 ? 4 : 0;
      if (!hb || !need(1 + (size_t)hb)) { dkg_set_error("dkg_wire_decode_rows: tagged integer without bin payload"); return DKG_ERR_INVALID; }
      for (int k = 0; k < hb; ++k) n = (n << 8) | buf[pos + 1 + k];
      pos += 1 + (size_t)hb;
      if (n == 0 || !need(n)) { dkg_set_error("dkg_wire_decode_rows: truncated integer"); return DKG_ERR_INVALID; }
      if (buf[pos + n - 1] & 0x80) { dkg_set_error("negative value in partial decryption message"); return DKG_ERR_INVALID; }
      size_t body = n;
      while (body > width && buf[pos + body - 1] == 0) --body;  // sign / padding bytes
      if (body > width) { dkg_set_error("value does not fit the modulus width"); return DKG_ERR_INVALID; }
      memcpy(dst, buf + pos, body);
      pos += n;
    } else if (t <= 0x7f) {
      dst[0] = t;
      pos += 1;
    } else if (t >= 0xcc && t <= 0xcf) {
      const int bytes = 1 << (t - 0xcc);
      if (!need(1 + (size_t)bytes)) { dkg_set_error("dkg_wire_decode_rows: truncated"); return DKG_ERR_INVALID; }
      uint64_t v = 0;
      for (int k = 0; k < bytes; ++k) v = (v << 8) | buf[pos + 1 + k];
      if (width < 8 && (v >> (8 * width)) != 0) { dkg_set_error("value does not fit the modulus width"); return DKG_ERR_INVALID; }
      memcpy(dst, &v, width < 8 ? width : 8);
      pos += 1 + (size_t)bytes;
    } else if (t >= 0xe0 || (t >= 0xd0 && t <= 0xd3)) {
      // negative fixint / signed ints: non-negative signed encodings are not minimal, msgpack
      // writers emit them only for negative numbers
      dkg_set_error("negative value in partial decryption message");
      return DKG_ERR_INVALID;
    } else {
      dkg_set_error("dkg_wire_decode_rows: unexpected msgpack type in integer list");
      return DKG_ERR_INVALID;
    }
  }
  if (consumed) *consumed = pos;
  return DKG_OK;
}
