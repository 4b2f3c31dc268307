// The filters around the biprimality-test exponentiations, so that a whole
// compute_modulus round (distributed_keygen.py:1288-1329 of the reference) can stay on the device:
//   * small-prime trial division of the candidates      (__small_prime_divisors_test, :1197-1209)
//   * Jacobi symbol of the jointly drawn g's             (sympy.jacobi_symbol(g, N) != 1 -> skip, :1089)
//   * selection of the first `correct` g's with symbol +1 (:1084-1091) and gathering them as the
//     bases of the grouped modexp.
// Thread-per-item, generic limb counts, limb arrays in local memory: these are O(L^2) bit-serial
// loops next to O(L^2 * E) exponentiations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dkg_combine.cuh"
#include "dkg_grouped_params_fwd.h"

namespace dkg {

// out[g] = 1 if moduli[g] is divisible by any of primes[0..nprimes), else 0.
// One thread per (candidate, 32-prime slice); N mod p by Horner on 32-bit limbs.
__global__ void small_prime_sieve_kernel(const uint32_t* moduli, int limbs, unsigned long long groups,
                                         const uint32_t* primes, int nprimes, uint8_t* out) {
  const unsigned long long g = blockIdx.x;
  if (g >= groups) return;
  const uint32_t* n = moduli + g * (unsigned long long)limbs;
  __shared__ uint32_t hit;
  if (threadIdx.x == 0) hit = 0;
  __syncthreads();
  uint32_t mine = 0;
  for (int k = threadIdx.x; k < nprimes; k += blockDim.x) {
    const uint32_t p = primes[k];
    uint64_t r = 0;
    for (int l = limbs - 1; l >= 0; --l) r = ((r << 32) | n[l]) % p;
    if (r == 0) mine = 1;
  }
  if (mine) atomicOr(&hit, 1u);
  __syncthreads();
  if (threadIdx.x == 0) out[g] = (uint8_t)hit;
}

// Jacobi symbol (a / n), n odd: binary algorithm with subtraction steps (a <- a - n keeps the
// symbol), multi-bit shifts.  sym[g][k] in {-1, 0, +1}.  One thread per (candidate, g value).
__global__ void __launch_bounds__(128) jacobi_kernel(const uint32_t* moduli, const uint32_t* gvals, int limbs,
                                                     unsigned long long groups, int per_group, int8_t* sym) {
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= groups * (unsigned long long)per_group) return;
  const unsigned long long g = idx / (unsigned long long)per_group;
  uint32_t A[kGroupedMaxLimbs], Nn[kGroupedMaxLimbs];
  uint32_t* a = A;
  uint32_t* n = Nn;
  const uint32_t* src = gvals + idx * (unsigned long long)limbs;
  const uint32_t* msrc = moduli + g * (unsigned long long)limbs;
  uint32_t nz = 0;
  for (int l = 0; l < limbs; ++l) { a[l] = src[l]; n[l] = msrc[l]; nz |= a[l]; }
  // reduce a below n if needed (the reference's g is already reduced; keep it robust): a >= n -> a -= n
  // repeatedly is unbounded, so fall back to the subtraction loop below, which tolerates a >= n.
  int len = limbs;
  int t = 1;
  int guard = 64 * limbs + 64;
  while (nz != 0 && guard-- > 0) {
    // strip factors of two from a
    while ((a[0] & 1u) == 0) {
      const uint32_t a0 = a[0];
      const int tz = a0 ? __ffs(a0) - 1 : 32;
      uint32_t lo = a0;
      for (int l = 0; l < len; ++l) {
        const uint32_t hi = (l + 1 < len) ? a[l + 1] : 0u;
        a[l] = (uint32_t)((((uint64_t)hi << 32) | lo) >> tz);
        lo = hi;
      }
      const uint32_t n8 = n[0] & 7u;
      if ((tz & 1) && (n8 == 3u || n8 == 5u)) t = -t;
    }
    // both odd: make a >= n (reciprocity when they swap)
    bool lt = false;
    for (int l = len - 1; l >= 0; --l)
      if (a[l] != n[l]) { lt = a[l] < n[l]; break; }
    if (lt) {
      uint32_t* tmp = a; a = n; n = tmp;
      if ((a[0] & 3u) == 3u && (n[0] & 3u) == 3u) t = -t;
    }
    uint32_t borrow = 0;
    nz = 0;
    for (int l = 0; l < len; ++l) {
      const uint64_t d = (uint64_t)a[l] - n[l] - borrow;
      a[l] = (uint32_t)d;
      nz |= (uint32_t)d;
      borrow = (uint32_t)(d >> 63);
    }
    while (len > 1 && a[len - 1] == 0 && n[len - 1] == 0) --len;
  }
  // gcd is in n: symbol is t if gcd == 1 else 0
  uint32_t rest = n[0] ^ 1u;
  for (int l = 1; l < limbs; ++l) rest |= n[l];
  sym[idx] = (int8_t)(rest == 0 ? t : 0);
}

// Per candidate: the first `correct` g's with symbol +1, in order.  pick[g][c] = index or -1.
__global__ void select_g_kernel(const int8_t* sym, unsigned long long groups, int per_group, int correct,
                                int* pick, int* count) {
  const unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  const int8_t* s = sym + g * (unsigned long long)per_group;
  int* p = pick + g * (unsigned long long)correct;
  int c = 0;
  for (int k = 0; k < per_group && c < correct; ++k)
    if (s[k] == 1) p[c++] = k;
  count[g] = c;
  for (; c < correct; ++c) p[c] = -1;
}

// bases[g][c] = gvals[g][pick[g][c]]  (or 1 where nothing was picked)
__global__ void gather_g_kernel(const uint32_t* gvals, const int* pick, int limbs, unsigned long long groups,
                                int per_group, int correct, uint32_t* bases) {
  const unsigned long long total = groups * (unsigned long long)correct * (unsigned long long)limbs;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const int l = (int)(i % limbs);
    const unsigned long long gc = i / limbs;
    const unsigned long long g = gc / correct;
    const int k = pick[gc];
    bases[i] = k >= 0 ? gvals[(g * (unsigned long long)per_group + (unsigned long long)k) * limbs + l] : (l == 0 ? 1u : 0u);
  }
}

// zero the rows of candidates' unused slots (c >= count[g])
__global__ void clear_unused_kernel(uint32_t* out, const int* count, int limbs, unsigned long long groups, int correct) {
  const unsigned long long total = groups * (unsigned long long)correct * (unsigned long long)limbs;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long gc = i / limbs;
    if ((int)(gc % correct) >= count[gc / correct]) out[i] = 0u;
  }
}

// Biprimality verdict (__biprime_test_with_v_i, distributed_keygen.py:1110-1175) for all
// candidates and tests at once: test k of candidate g succeeds iff v_1 = +- prod_{i>1} v_i (mod N).
// v: [parties][groups][correct][limbs] (party 1 first), canonical residues; ok[g] must be
// pre-set to 1 and is cleared by any failing test.  The product is formed with word-serial
// Montgomery multiplications without ever converting into Montgomery form: after multiplying the
// P-1 factors the value carries R^-(P-2), and v_1 is scaled by the same power through
// multiplications by 1, so only -N^-1 mod 2^32 is needed per candidate.
__global__ void __launch_bounds__(64) biprime_verdict_kernel(const uint32_t* moduli, const uint32_t* v, int limbs,
                                                             unsigned long long groups, int parties, int correct,
                                                             uint32_t* ok) {
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= groups * (unsigned long long)correct) return;
  const unsigned long long g = idx / (unsigned long long)correct;
  const uint32_t* n = moduli + g * (unsigned long long)limbs;
  uint32_t n0inv = n[0];
  for (int i = 0; i < 5; ++i) n0inv *= 2u - n[0] * n0inv;
  n0inv = 0u - n0inv;
  const unsigned long long party_stride = groups * (unsigned long long)correct * (unsigned long long)limbs;
  const uint32_t* row = v + idx * (unsigned long long)limbs;
  uint32_t x[kGroupedMaxLimbs], t1[kGroupedMaxLimbs], tmp[kGroupedMaxLimbs + 2], one[kGroupedMaxLimbs];
  for (int l = 0; l < limbs; ++l) { t1[l] = row[l]; x[l] = parties > 1 ? row[party_stride + l] : (l == 0 ? 1u : 0u); one[l] = (l == 0) ? 1u : 0u; }
  for (int p = 2; p < parties; ++p) {
    gen_mont_mul(x, row + (unsigned long long)p * party_stride, 1, n, n0inv, limbs, tmp);
    gen_mont_mul(t1, one, 1, n, n0inv, limbs, tmp);
  }
  // success iff x == t1 or x + t1 == N (both canonical)
  uint32_t diff = 0;
  uint64_t carry = 0;
  uint32_t sumdiff = 0;
  for (int l = 0; l < limbs; ++l) {
    diff |= x[l] ^ t1[l];
    const uint64_t s = (uint64_t)x[l] + t1[l] + carry;
    sumdiff |= (uint32_t)s ^ n[l];
    carry = s >> 32;
  }
  sumdiff |= (uint32_t)carry;
  if (diff != 0 && sumdiff != 0) atomicAnd(&ok[g], 0u);
}

}  // namespace dkg
