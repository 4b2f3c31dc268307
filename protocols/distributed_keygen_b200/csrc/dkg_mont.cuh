// Block-column Montgomery multiplication, one big integer per thread.
//
// A value of L = K*M 32-bit limbs is M blocks of K limbs (K even).  The product is formed
// column by column over blocks (product scanning): column c sums X_i*Y_{c-i} and Q_i*N_{c-i}
// into a (2K+2)-limb register accumulator T; each K x K block product is a fully unrolled
// schoolbook on IMAD.WIDE.U32.X carry chains, split into an even-aligned (E) and an odd-aligned
// (O) accumulator so every 64-bit accumulate hits an aligned register pair.  The Montgomery
// quotient block Q_c = T_low * (-N^-1) mod 2^(32K) is one low-half block product; adding
// Q_c*N_0 zeroes the low block.  All operands other than T/E/O/two operand blocks live in the
// caller's storage (shared memory on the device) behind the IO policy, so code size is one block
// product per kernel, not M^2 of them.
//
// Values are kept in [0, R), R = 2^(32*K*M) ("almost Montgomery"): the result of a*b*R^-1 is
// < R + N, and N is subtracted once iff the carry limb is set.  Canonical reduction happens once,
// at the end of an exponentiation.
//
// Work per multiplication: 2*M^2 full block products + M low-half products
//   = 2*L^2 + L*(K+1)/2 wide multiply-accumulates (the canonical count is 2*L^2 + L).
#pragma once
#include "dkg_prims.cuh"

namespace dkg {

enum MontMode { MONT_MUL = 0, MONT_REDC = 1 };

// E + (O << 32) = x * y  (fresh product; E, O have 2K+2 limbs, the top ones end up zero)
template <int K>
DKG_HD void block_mul(uint32_t (&E)[2 * K + 2], uint32_t (&O)[2 * K + 2], const uint32_t (&x)[K],
                      const uint32_t (&y)[K]) {
  static_assert(K % 2 == 0 && K >= 2, "K must be even");
#pragma unroll
  for (int i = 0; i < 2 * K + 2; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
  for (int j = 0; j < K; j++) {
    // products x_i*y_j with i+j even land in E at limb i+j; with i+j odd in O at limb i+j-1
    if ((j & 1) == 0) {
      mad_cc(E[j], E[j + 1], x[0], y[j]);
#pragma unroll
      for (int i = 2; i < K; i += 2) madc_cc(E[i + j], E[i + j + 1], x[i], y[j]);
      addc(E[j + K], 0);
      mad_cc(O[j], O[j + 1], x[1], y[j]);
#pragma unroll
      for (int i = 3; i < K; i += 2) madc_cc(O[i + j - 1], O[i + j], x[i], y[j]);
      addc(O[j + K], 0);
    } else {
      mad_cc(E[j + 1], E[j + 2], x[1], y[j]);
#pragma unroll
      for (int i = 3; i < K; i += 2) madc_cc(E[i + j], E[i + j + 1], x[i], y[j]);
      addc(E[j + K + 1], 0);
      mad_cc(O[j - 1], O[j], x[0], y[j]);
#pragma unroll
      for (int i = 2; i < K; i += 2) madc_cc(O[i + j - 1], O[i + j], x[i], y[j]);
      addc(O[j + K - 1], 0);
    }
  }
}

// T += E + (O << 32)
template <int K>
DKG_HD void acc_add(uint32_t (&T)[2 * K + 2], const uint32_t (&E)[2 * K + 2],
                    const uint32_t (&O)[2 * K + 2]) {
  add_cc(T[0], E[0]);
#pragma unroll
  for (int p = 1; p <= 2 * K; p++) addc_cc(T[p], E[p]);
  addc(T[2 * K + 1], 0);
  add_cc(T[1], O[0]);
#pragma unroll
  for (int p = 1; p <= 2 * K - 1; p++) addc_cc(T[p + 1], O[p]);
  addc(T[2 * K + 1], 0);
}

// r = x[0..K) * y mod 2^(32K)   (x is the low block of a wider array)
template <int K, int XN>
DKG_HD void block_mul_lo(uint32_t (&r)[K], const uint32_t (&x)[XN], const uint32_t (&y)[K]) {
  static_assert(XN >= K, "x too short");
  uint32_t E[K], O[K];
#pragma unroll
  for (int i = 0; i < K; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
  for (int j = 0; j < K; j++) {
    // E chain: i = j (mod 2), limb position i+j (even) <= K-2, pair (i+j, i+j+1) always in range
    {
      const int i0 = j & 1;
      if (i0 + j <= K - 2) {
        mad_cc(E[i0 + j], E[i0 + j + 1], x[i0], y[j]);
#pragma unroll
        for (int i = i0 + 2; i + j <= K - 2; i += 2) madc_cc(E[i + j], E[i + j + 1], x[i], y[j]);
      }
    }
    // O chain: i != j (mod 2), limb position i+j (odd) stored at O[i+j-1]; position K-1 keeps
    // only its low half
    {
      const int i0 = (j & 1) ^ 1;
      if (i0 + j <= K - 3) {
        mad_cc(O[i0 + j - 1], O[i0 + j], x[i0], y[j]);
#pragma unroll
        for (int i = i0 + 2; i + j <= K - 3; i += 2) madc_cc(O[i + j - 1], O[i + j], x[i], y[j]);
        // the element at position K-1 (if this row reaches it) continues the chain, low half only
        if (((K - 1 - j) & 1) == i0 && K - 1 - j >= 0) madc_lo(O[K - 2], x[K - 1 - j], y[j]);
      } else if (i0 + j == K - 1) {
        mad_lo(O[K - 2], x[i0], y[j]);
      }
    }
  }
  r[0] = E[0];
  if (K == 2) {
    r[1] = E[1] + O[0];
  } else {
    r[1] = E[1];
    add_cc(r[1], O[0]);
#pragma unroll
    for (int p = 2; p < K - 1; p++) { r[p] = E[p]; addc_cc(r[p], O[p - 1]); }
    r[K - 1] = E[K - 1];
    addc(r[K - 1], O[K - 2]);
  }
}

// IO policy (all indices are block indices; r has K limbs):
//   load_x(i, r)  load_y(j, r)  load_q(i, r)  load_n(j, r)  load_ninv(r)
//   store_q(i, r) store_x(i, r)
//
// MONT_MUL : X <- X * Y * R^-1 mod N   (Y may alias X: squaring)
// MONT_REDC: X <- X * R^-1 mod N
// Result in [0, R); X is overwritten block by block (block c-M is dead when column c starts).
template <int K, int M, int MODE, class IO>
DKG_HD void mont_mul(const IO& io) {
  uint32_t T[2 * K + 2];
#pragma unroll
  for (int i = 0; i < 2 * K + 2; i++) T[i] = 0;

  for (int c = 0; c < 2 * M; ++c) {
    const int lo = c >= M ? c - M + 1 : 0;
    const int hi = c < M ? c : M - 1;
    const int nxy = (MODE == MONT_REDC) ? 0 : (hi - lo + 1);
    const int nq = (c < M ? c - 1 : M - 1) - lo + 1;
    const int total = nxy + nq + (c < M ? 1 : 0);

    if (MODE == MONT_REDC && c < M) {
      uint32_t xb[K];
      io.load_x(c, xb);
      add_cc(T[0], xb[0]);
#pragma unroll
      for (int p = 1; p < K; p++) addc_cc(T[p], xb[p]);
#pragma unroll
      for (int p = K; p <= 2 * K; p++) addc_cc(T[p], 0);
      addc(T[2 * K + 1], 0);
    }

    for (int t = 0; t < total; ++t) {
      uint32_t xb[K], yb[K];
      if (t < nxy) {
        const int i = lo + t;
        io.load_x(i, xb);
        io.load_y(c - i, yb);
      } else if (t < nxy + nq) {
        const int i = lo + (t - nxy);
        io.load_q(i, xb);
        io.load_n(c - i, yb);
      } else {
        // quotient block: Q_c = T_low * (-N^-1) mod 2^(32K); Q_c * N_0 then clears T_low
        io.load_ninv(yb);
        block_mul_lo<K>(xb, T, yb);
        io.store_q(c, xb);
        io.load_n(0, yb);
      }
      uint32_t E[2 * K + 2], O[2 * K + 2];
      block_mul<K>(E, O, xb, yb);
      acc_add<K>(T, E, O);
    }

    if (c >= M) {
      uint32_t ob[K];
#pragma unroll
      for (int p = 0; p < K; p++) ob[p] = T[p];
      io.store_x(c - M, ob);
    }
#pragma unroll
    for (int p = 0; p < K + 2; p++) T[p] = T[p + K];
#pragma unroll
    for (int p = K + 2; p < 2 * K + 2; p++) T[p] = 0;
  }

  // result = T[0]*R + X < R + N: subtract N once iff the carry limb is set
  const uint32_t mask = 0u - T[0];
  uint32_t borrow = 0;
  for (int b = 0; b < M; ++b) {
    uint32_t xb[K], nb[K];
    io.load_x(b, xb);
    io.load_n(b, nb);
#pragma unroll
    for (int p = 0; p < K; p++) {
      const uint64_t d = (uint64_t)xb[p] - (nb[p] & mask) - borrow;
      xb[p] = (uint32_t)d;
      borrow = (uint32_t)(d >> 63);
    }
    io.store_x(b, xb);
  }
}

// X >= N ?  (returns 1/0).  Scans all blocks, no early exit.
template <int K, int M, class IO>
DKG_HD uint32_t geq_n(const IO& io) {
  uint32_t borrow = 0;
  for (int b = 0; b < M; ++b) {
    uint32_t xb[K], nb[K];
    io.load_x(b, xb);
    io.load_n(b, nb);
#pragma unroll
    for (int p = 0; p < K; p++) {
      const uint64_t d = (uint64_t)xb[p] - nb[p] - borrow;
      borrow = (uint32_t)(d >> 63);
    }
  }
  return borrow ^ 1u;
}

// X <- X - (N & mask)
template <int K, int M, class IO>
DKG_HD void sub_n_masked(const IO& io, uint32_t mask) {
  uint32_t borrow = 0;
  for (int b = 0; b < M; ++b) {
    uint32_t xb[K], nb[K];
    io.load_x(b, xb);
    io.load_n(b, nb);
#pragma unroll
    for (int p = 0; p < K; p++) {
      const uint64_t d = (uint64_t)xb[p] - (nb[p] & mask) - borrow;
      xb[p] = (uint32_t)d;
      borrow = (uint32_t)(d >> 63);
    }
    io.store_x(b, xb);
  }
}

// bring X in [0, R) (congruent to the true value) to the canonical residue in [0, N).
// After a MONT_REDC the value is <= N, so one conditional subtraction suffices; `rounds` > 1 is
// for callers that canonicalise a raw [0, R) value with R < 2^rounds * N.
template <int K, int M, class IO>
DKG_HD void canonicalize(const IO& io, int rounds = 1) {
  for (int r = 0; r < rounds; ++r) {
    const uint32_t ge = geq_n<K, M>(io);
    sub_n_masked<K, M>(io, 0u - ge);
  }
}

}  // namespace dkg
