// Share combination (PaillierSharedKey.decrypt, paillier_shared_key.py:108-125 of the reference):
//   x = prod_j partial_j mod N^2 ; (x - 1) % N == 0 else flag ; m = ((x - 1)/N * theta^-1) mod N
// One ciphertext per thread, generic limb counts, word-serial Montgomery (CIOS) on thread-local
// arrays.  This path is ~0.05 % of a threshold decryption's arithmetic (7.4e4 of 5e8 wide-MACs),
// so it is written for exactness and generality, not for the multiplier roofline.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dkg {

constexpr int kCombineMaxL2 = 272;  // limbs of N^2
constexpr int kCombineMaxL = 137;   // limbs of N

struct CombineParams {
  const uint32_t* partials;  // [shares][count][l2]
  uint32_t* out;             // [count][ln]
  uint8_t* status;           // [count]
  unsigned long long count;
  int shares;
  int l2, ln;
  // device constants: N2[l2] | RPOW[l2] | N[ln] | NINVPOS[ln] | THETAINV_R[ln]
  const uint32_t* consts;
  uint32_t n2_0inv;  // -N2^-1 mod 2^32
  uint32_t n_0inv;   // -N^-1 mod 2^32
};

// x <- x * b / 2^(32 L) mod n, all operands < n, result < n.  b is read with stride bstride.
__device__ inline void gen_mont_mul(uint32_t* x, const uint32_t* b, size_t bstride, const uint32_t* n,
                                    uint32_t n0inv, int L, uint32_t* t) {
  for (int j = 0; j < L + 2; ++j) t[j] = 0;
  for (int i = 0; i < L; ++i) {
    const uint32_t bi = b[(size_t)i * bstride];
    uint64_t carry = 0;
    for (int j = 0; j < L; ++j) {
      const uint64_t s = (uint64_t)x[j] * bi + t[j] + carry;
      t[j] = (uint32_t)s;
      carry = s >> 32;
    }
    uint64_t s = (uint64_t)t[L] + carry;
    t[L] = (uint32_t)s;
    t[L + 1] = (uint32_t)(s >> 32);
    const uint32_t m = t[0] * n0inv;
    carry = ((uint64_t)m * n[0] + t[0]) >> 32;
    for (int j = 1; j < L; ++j) {
      const uint64_t s2 = (uint64_t)m * n[j] + t[j] + carry;
      t[j - 1] = (uint32_t)s2;
      carry = s2 >> 32;
    }
    s = (uint64_t)t[L] + carry;
    t[L - 1] = (uint32_t)s;
    t[L] = t[L + 1] + (uint32_t)(s >> 32);
  }
  // conditional subtraction: t (L+1 limbs) >= n ?
  uint32_t borrow = 0;
  for (int j = 0; j < L; ++j) {
    const uint64_t d = (uint64_t)t[j] - n[j] - borrow;
    borrow = (uint32_t)(d >> 63);
  }
  const uint32_t ge = (t[L] != 0 || borrow == 0) ? 0xffffffffu : 0u;
  borrow = 0;
  for (int j = 0; j < L; ++j) {
    const uint64_t d = (uint64_t)t[j] - (n[j] & ge) - borrow;
    x[j] = (uint32_t)d;
    borrow = (uint32_t)(d >> 63);
  }
}

__global__ void __launch_bounds__(128) combine_kernel(const CombineParams p) {
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.count) return;
  const int L2 = p.l2, Ln = p.ln;
  const uint32_t* N2 = p.consts;
  const uint32_t* RPOW = N2 + L2;
  const uint32_t* N = RPOW + L2;
  const uint32_t* NINV = N + Ln;
  const uint32_t* THR = NINV + Ln;

  uint32_t x[kCombineMaxL2];
  uint32_t t[kCombineMaxL2 + 2];
  uint32_t u[kCombineMaxL];

  // x = partial_0 ; x = x * partial_j / R ... ; x = x * R^shares / R  => plain product mod N^2
  const uint32_t* row0 = p.partials + idx * (unsigned long long)L2;
  for (int j = 0; j < L2; ++j) x[j] = row0[j];
  for (int s = 1; s < p.shares; ++s) {
    const uint32_t* row = p.partials + ((unsigned long long)s * p.count + idx) * (unsigned long long)L2;
    gen_mont_mul(x, row, 1, N2, p.n2_0inv, L2, t);
  }
  gen_mont_mul(x, RPOW, 1, N2, p.n2_0inv, L2, t);

  // y = x - 1   (x == 0 gives "not divisible", as (0 - 1) % N != 0 in the reference)
  uint32_t nz = 0;
  for (int j = 0; j < L2; ++j) nz |= x[j];
  uint32_t bad = nz ? 0u : 1u;
  {
    uint32_t borrow = 1;
    for (int j = 0; j < L2; ++j) {
      const uint64_t d = (uint64_t)x[j] - borrow;
      x[j] = (uint32_t)d;
      borrow = (uint32_t)(d >> 63);
    }
  }
  // u = y * N^-1 mod 2^(32 Ln): the exact quotient if N | y
  for (int j = 0; j < Ln; ++j) u[j] = 0;
  for (int i = 0; i < Ln; ++i) {
    const uint32_t yi = x[i];
    uint64_t carry = 0;
    for (int j = 0; i + j < Ln; ++j) {
      const uint64_t s = (uint64_t)yi * NINV[j] + u[i + j] + carry;
      u[i + j] = (uint32_t)s;
      carry = s >> 32;
    }
  }
  // divisibility: u * N == y on all L2 limbs (column sums with a 96-bit accumulator)
  {
    uint32_t c0 = 0, c1 = 0, c2 = 0;
    for (int c = 0; c < L2; ++c) {
      const int lo = c - (Ln - 1) > 0 ? c - (Ln - 1) : 0;
      const int hi = c < Ln - 1 ? c : Ln - 1;
      for (int i = lo; i <= hi; ++i) {
        const uint64_t pr = (uint64_t)u[i] * N[c - i];
        const uint64_t s0 = (uint64_t)c0 + (uint32_t)pr;
        c0 = (uint32_t)s0;
        const uint64_t s1 = (uint64_t)c1 + (uint32_t)(pr >> 32) + (s0 >> 32);
        c1 = (uint32_t)s1;
        c2 += (uint32_t)(s1 >> 32);
      }
      bad |= (c0 ^ x[c]);
      c0 = c1; c1 = c2; c2 = 0;
    }
    bad |= c0 | c1;  // product wider than L2 limbs
  }
  // m = u * theta^-1 mod N   (u < N when divisible)
  gen_mont_mul(u, THR, 1, N, p.n_0inv, Ln, t);
  uint32_t* orow = p.out + idx * (unsigned long long)Ln;
  for (int j = 0; j < Ln; ++j) orow[j] = bad ? 0u : u[j];
  if (p.status) p.status[idx] = bad ? 2 : 0;
}

}  // namespace dkg
