// Entry / exit of the pair arithmetic (dkg_nsq.cuh): per element, generic limb counts, word-serial
// Montgomery on thread-local arrays.  ~0.3 % of an exponentiation's work.
//   entry: c (< N^2)  ->  (c mod N, (c div N) * R mod N)
//   exit : (a, b) with result y = a + N * (b R^-1 mod N)  ->  canonical y < N^2
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dkg_combine.cuh"

namespace dkg {

constexpr int kNsqMaxL = 144;

struct NsqIoParams {
  const uint32_t* in;       // entry: [count][in_limbs] values below N^2 ; exit: [count][2][Lp] pairs
  uint32_t* out;            // entry: [count][2][Lp] ; exit: [count][out_limbs]
  unsigned long long count;
  int io_limbs;             // limbs of a value modulo N^2 at the ABI (in_limbs / out_limbs)
  int Lp;                   // limbs of the pair components (R = 2^(32 Lp) >= 8 N)
  // device constants: N[Lp] | R2N[Lp] (R^2 mod N) | NINVPOS[Lp] (N^-1 mod R)
  const uint32_t* consts;
  uint32_t n0inv;           // -N^-1 mod 2^32
  // exit only, optional: plaintexts m (< N), [count][m_limbs]: the result is multiplied by (1 + m N)
  // (Paillier encryption with g = N + 1): y0 + N (y1 + y0 m mod N)
  const uint32_t* mrows;
  int m_limbs;
};

__global__ void __launch_bounds__(64) nsq_entry_kernel(const NsqIoParams p) {
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.count) return;
  const int Lp = p.Lp;
  const uint32_t* N = p.consts;
  const uint32_t* R2N = N + Lp;
  const uint32_t* NINVPOS = R2N + Lp;
  const uint32_t* c = p.in + idx * (unsigned long long)p.io_limbs;
  uint32_t x[kNsqMaxL], t[kNsqMaxL + 2], one[kNsqMaxL], c1[kNsqMaxL];
  auto climb = [&](int l) -> uint32_t { return l < p.io_limbs ? c[l] : 0u; };
  for (int l = 0; l < Lp; ++l) { x[l] = climb(l); one[l] = (l == 0) ? 1u : 0u; }
  // c mod N = ((c_lo R^-1 + c_hi) mod N) * R^2 * R^-1,  c = c_lo + c_hi R, c_hi < N/8
  gen_mont_mul(x, one, 1, N, p.n0inv, Lp, t);
  {
    uint64_t carry = 0;
    for (int l = 0; l < Lp; ++l) { const uint64_t s = (uint64_t)x[l] + climb(Lp + l) + carry; x[l] = (uint32_t)s; carry = s >> 32; }
    uint32_t borrow = 0;
    for (int l = 0; l < Lp; ++l) { const uint64_t d = (uint64_t)x[l] - N[l] - borrow; t[l] = (uint32_t)d; borrow = (uint32_t)(d >> 63); }
    if (borrow == 0) for (int l = 0; l < Lp; ++l) x[l] = t[l];
  }
  gen_mont_mul(x, R2N, 1, N, p.n0inv, Lp, t);                  // x = c mod N
  // c1 = (c - c0) / N exactly = (c - c0) * N^-1 mod R   (c1 < N < R)
  {
    uint32_t borrow = 0;
    for (int l = 0; l < Lp; ++l) { const uint64_t d = (uint64_t)climb(l) - x[l] - borrow; t[l] = (uint32_t)d; borrow = (uint32_t)(d >> 63); }
    for (int l = 0; l < Lp; ++l) c1[l] = 0;
    for (int i = 0; i < Lp; ++i) {
      const uint32_t yi = t[i];
      uint64_t carry = 0;
      for (int j = 0; i + j < Lp; ++j) {
        const uint64_t s = (uint64_t)yi * NINVPOS[j] + c1[i + j] + carry;
        c1[i + j] = (uint32_t)s;
        carry = s >> 32;
      }
    }
  }
  uint32_t* o = p.out + idx * (unsigned long long)(2 * Lp);
  for (int l = 0; l < Lp; ++l) o[l] = x[l];
  gen_mont_mul(c1, R2N, 1, N, p.n0inv, Lp, t);                 // (c div N) * R mod N
  for (int l = 0; l < Lp; ++l) o[Lp + l] = c1[l];
}

__global__ void __launch_bounds__(64) nsq_exit_kernel(const NsqIoParams p) {
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.count) return;
  const int Lp = p.Lp;
  const uint32_t* N = p.consts;
  const uint32_t* pr = p.in + idx * (unsigned long long)(2 * Lp);
  uint32_t a[kNsqMaxL], h[kNsqMaxL], t[kNsqMaxL + 2], one[kNsqMaxL];
  for (int l = 0; l < Lp; ++l) { a[l] = pr[l]; h[l] = pr[Lp + l]; one[l] = (l == 0) ? 1u : 0u; }
  gen_mont_mul(h, one, 1, N, p.n0inv, Lp, t);                  // h = b R^-1 mod N, canonical
  // a < 2N as an integer: a = a' + N moves one N into h (h <- h + 1 mod N)
  {
    uint32_t borrow = 0;
    for (int l = 0; l < Lp; ++l) { const uint64_t d = (uint64_t)a[l] - N[l] - borrow; t[l] = (uint32_t)d; borrow = (uint32_t)(d >> 63); }
    if (borrow == 0) {
      for (int l = 0; l < Lp; ++l) a[l] = t[l];
      uint64_t carry = 1;
      for (int l = 0; l < Lp; ++l) { const uint64_t s2 = (uint64_t)h[l] + carry; h[l] = (uint32_t)s2; carry = s2 >> 32; }
      uint32_t diff = 0;
      for (int l = 0; l < Lp; ++l) diff |= h[l] ^ N[l];
      if (diff == 0) for (int l = 0; l < Lp; ++l) h[l] = 0;
    }
  }
  if (p.mrows != nullptr) {
    // (a + N h)(1 + m N) = a + N (h + a m)  (mod N^2)
    const uint32_t* R2N = N + Lp;
    const uint32_t* mr = p.mrows + idx * (unsigned long long)p.m_limbs;
    uint32_t am[kNsqMaxL], mm[kNsqMaxL];
    for (int l = 0; l < Lp; ++l) { am[l] = a[l]; mm[l] = l < p.m_limbs ? mr[l] : 0u; }
    gen_mont_mul(am, mm, 1, N, p.n0inv, Lp, t);      // a m R^-1
    gen_mont_mul(am, R2N, 1, N, p.n0inv, Lp, t);     // a m mod N
    uint64_t carry = 0;
    for (int l = 0; l < Lp; ++l) { const uint64_t s2 = (uint64_t)h[l] + am[l] + carry; h[l] = (uint32_t)s2; carry = s2 >> 32; }
    uint32_t borrow = 0;
    for (int l = 0; l < Lp; ++l) { const uint64_t d = (uint64_t)h[l] - N[l] - borrow; t[l] = (uint32_t)d; borrow = (uint32_t)(d >> 63); }
    if (borrow == 0) for (int l = 0; l < Lp; ++l) h[l] = t[l];
  }
  // y = a + N * h  (< N^2), low io_limbs limbs
  uint32_t* o = p.out + idx * (unsigned long long)p.io_limbs;
  uint32_t c0 = 0, c1 = 0, c2 = 0;
  for (int col = 0; col < p.io_limbs; ++col) {
    if (col < Lp) { const uint64_t s = (uint64_t)c0 + a[col]; c0 = (uint32_t)s; const uint64_t s1 = (uint64_t)c1 + (s >> 32); c1 = (uint32_t)s1; c2 += (uint32_t)(s1 >> 32); }
    const int lo = col - (Lp - 1) > 0 ? col - (Lp - 1) : 0;
    const int hi = col < Lp - 1 ? col : Lp - 1;
    for (int i = lo; i <= hi; ++i) {
      const uint64_t pr2 = (uint64_t)N[i] * h[col - i];
      const uint64_t s0 = (uint64_t)c0 + (uint32_t)pr2;
      c0 = (uint32_t)s0;
      const uint64_t s1 = (uint64_t)c1 + (uint32_t)(pr2 >> 32) + (s0 >> 32);
      c1 = (uint32_t)s1;
      c2 += (uint32_t)(s1 >> 32);
    }
    o[col] = c0;
    c0 = c1; c1 = c2; c2 = 0;
  }
}

}  // namespace dkg
