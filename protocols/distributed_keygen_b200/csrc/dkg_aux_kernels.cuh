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
