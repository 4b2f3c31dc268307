// Small helper kernels around the modexp engine (none of them is on the multiplier roofline).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dkg {

// dst[i][0..dst_limbs) = src[i][0..src_limbs) zero-extended
__global__ void pad_rows_kernel(const uint32_t* src, int src_limbs, uint32_t* dst, int dst_limbs,
                                unsigned long long count) {
  const unsigned long long total = count * (unsigned long long)dst_limbs;
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long row = idx / dst_limbs;
    const int l = (int)(idx % dst_limbs);
    dst[idx] = l < src_limbs ? src[row * src_limbs + l] : 0u;
  }
}

// Rows that are not below the modulus (the C ABI requires values < modulus; the reference's pow_mod
// would reduce them, the Montgomery kernels would silently compute something else): flag them with
// DKG_STATUS_OUT_OF_RANGE (3) and zero the result row.  One warp per row, run after the compute
// kernels (which are safe on any input that fits the limb width).
__global__ void range_check_kernel(const uint32_t* bases, const uint32_t* modulus, int limbs, unsigned long long count,
                                   uint32_t* out, uint8_t* status) {
  const unsigned long long row = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= count) return;
  const uint32_t* b = bases + row * (unsigned long long)limbs;
  // most significant differing limb decides
  int verdict = 0;   // +1: base > modulus, -1: base < modulus
  for (int hi = limbs - 1; hi >= 0 && verdict == 0; hi -= 32) {
    const int l = hi - lane;
    const uint32_t x = l >= 0 ? b[l] : 0u, m = l >= 0 ? modulus[l] : 0u;
    const unsigned gt = __ballot_sync(0xffffffffu, x > m), lt = __ballot_sync(0xffffffffu, x < m);
    if (gt | lt) verdict = (__ffs(gt) != 0 && (lt == 0 || __ffs(gt) < __ffs(lt))) ? 1 : -1;
  }
  if (verdict >= 0) {   // base >= modulus
    for (int l = lane; l < limbs; l += 32) out[row * (unsigned long long)limbs + l] = 0;
    if (status != nullptr && lane == 0) status[row] = 3;
  }
}

// out[i] = 1 + m[i] * N   (the Paillier plaintext factor (1 + N)^m mod N^2 for g = N + 1;
// third-party Paillier raw encryption, distributed_keygen.py:712).  m < N, so 1 + m N < N^2.
// One element per thread, schoolbook, operands read with a warp-friendly stride-free pattern is
// not attempted: 4k wide-MACs per element against 8e7 for the r^N that follows.
__global__ void one_plus_mn_kernel(const uint32_t* m, const uint32_t* n, int n_limbs, uint32_t* out,
                                   int out_limbs, unsigned long long count) {
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  const uint32_t* mi = m + idx * (unsigned long long)n_limbs;
  uint32_t* o = out + idx * (unsigned long long)out_limbs;
  for (int l = 0; l < out_limbs; ++l) o[l] = 0;
  for (int i = 0; i < n_limbs; ++i) {
    const uint32_t a = mi[i];
    uint64_t carry = 0;
    for (int j = 0; j < n_limbs && i + j < out_limbs; ++j) {
      const uint64_t t = (uint64_t)a * n[j] + o[i + j] + carry;
      o[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    for (int l = i + n_limbs; carry != 0 && l < out_limbs; ++l) {
      const uint64_t t = (uint64_t)o[l] + carry;
      o[l] = (uint32_t)t;
      carry = t >> 32;
    }
  }
  uint64_t carry = 1;
  for (int l = 0; carry != 0 && l < out_limbs; ++l) {
    const uint64_t t = (uint64_t)o[l] + carry;
    o[l] = (uint32_t)t;
    carry = t >> 32;
  }
}

}  // namespace dkg
