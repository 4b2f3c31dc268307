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
//   = 2*L^2 + L*(K+1)/2 wide multiply-accumulates (the canonical count is 2*L^2 + L);
// per squaring: M(M+1)/2 + M^2 block products + M low-half products (cross products doubled).
#pragma once
#include "dkg_prims.cuh"

namespace dkg {

// MONT_MUL2S / MONT_MULADD serve the pair arithmetic modulo N^2 (dkg_nsq.cuh):
//   MUL2S : X <- 2 * X * S * R^-1      (S = second shared-memory operand, products doubled)
//   MULADD: X <- (X * Y + S * Y2) * R^-1   (Y, Y2 in global memory)
enum MontMode { MONT_MUL = 0, MONT_REDC = 1, MONT_SQR = 2, MONT_MUL2S = 3, MONT_MULADD = 4 };

// Pipe-balance ballast.  ptxas decides per FUNCTION, from static instruction counts and blind to
// loop depth, on which pipe the register moves and single carry adds of the whole function go:
// ALU (MOV, IADD3.X) or the multiplier pipe (IMAD.MOV.U32, IMAD.X).  The merge/shift/subtract code
// around the block product is ALU-heavy, so without help it sends ~70 moves and ~15 carry adds per
// block product to the multiplier pipe -- the one pipe this kernel saturates (measured: 16 % of
// its cycles, profiles/r01_ncu_nsq_source_opcodes.txt).  A block of FFMAs that is never executed
// (guarded by io.never(), a run-time condition that is never true) tips the static balance, and every one of those instructions
// moves to the idle ALU pipe.  It costs code bytes that are never fetched.
#ifndef DKG_PIPE_BALLAST
#define DKG_PIPE_BALLAST 4096
#endif
DKG_HD uint32_t pipe_ballast(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  float f = __uint_as_float(a);
  const float g = __uint_as_float(b);
#pragma unroll
  for (int q = 0; q < DKG_PIPE_BALLAST; q++) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(g));
  return __float_as_uint(f);
#else
  return a ^ b;
#endif
}

// Column accumulator of the block product scan, kept in a carry-save form so that a K x K block
// multiply-accumulate is nothing but IMAD.WIDE carry chains:
//   value = E + (O << 32) + sum_k CE[k] << 32(K+2k) + sum_k CO[k] << 32(K+2k+1)
// E takes the partial products a_i*b_j with i+j even (64-bit aligned at even limbs), O those with
// i+j odd (aligned at odd limbs); the carry out of every row chain is counted in CE/CO instead of
// being rippled through the upper limbs, so rows stay independent and nothing is merged inside
// the inner loop.  `merge` folds everything back into E (once or twice per column).
// K may be odd (a 2048-bit-class N needs 65 limbs: 5 blocks of 13 instead of 5 of 14 saves 14 % of
// the multiplications).  Operand blocks then sit in slots of KP = K + 1 limbs (whole 64-bit vectors),
// the pad limb zero and never multiplied.  Carry counters by limb position: CE[k] weighs limb
// KP + 2k (the even positions from K up), CO[k] limb KO + 2k (the odd ones), KO = K + 1 - (K & 1).
template <int K> constexpr int kpad = K + (K & 1);          // slot size in limbs
template <int K> constexpr int kodd0 = K + 1 - (K & 1);     // first odd limb position >= K
template <int K> constexpr int kncE = (2 * K - kpad<K>) / 2 + 1;
template <int K> constexpr int kncO = (2 * K - 1 - kodd0<K>) / 2 + 1;

template <int K>
struct ColAcc {
  uint64_t E[K + 1];   // limbs (2p, 2p+1)
  uint64_t O[K - 1];   // limbs (2p+1, 2p+2)
  uint32_t CE[kncE<K>];
  uint32_t CO[kncO<K>];
};

template <int K>
DKG_HD void acc_clear_side(ColAcc<K>& a) {
#pragma unroll
  for (int i = 0; i < K - 1; i++) a.O[i] = 0;
#pragma unroll
  for (int i = 0; i < kncE<K>; i++) a.CE[i] = 0;
#pragma unroll
  for (int i = 0; i < kncO<K>; i++) a.CO[i] = 0;
}

// Operand kinds of a block product, also the source selector of the prefetch:
//   PAIR_XY: x = X_i (shared), y = Y_j (global, multiplication operand / table entry)
//   PAIR_XX: x = X_i, y = X_j (shared; squaring)
//   PAIR_NQ: x = N_i (shared, CTA-uniform), y = Q_j (global scratch, written earlier by this thread)
//   PAIR_QC: the quotient step (x = Q_c fresh in registers, y = N_0)
//   PAIR_SY2: x = S_i, y = Y2_j (second global operand)
//   PAIR_XX2: x = block i of 2*X, y = X_j (cross product of a squaring, i < j)
//   PAIR_SX2: x = block i of 2*S, y = X_j (doubled product)
// "block i of 2*V" = (V_i << 1 | top bit of V_{i-1}) mod 2^(32K): the factor 2 of the squaring
// modes is applied to an operand as it is loaded, not to the accumulated products.
enum PairKind { PAIR_XY = 0, PAIR_XX = 1, PAIR_NQ = 2, PAIR_QC = 3, PAIR_NONE = 4, PAIR_SY2 = 6,
                PAIR_XX2 = 7, PAIR_SX2 = 8 };

struct PairDesc {
  int kind, xi, yi;
};

// acc += x * y.  y[j] is dead once row j is done, so while this product runs, y is refilled behind
// the scan with the y operand of the NEXT block product (descriptor pf from io.prefetch_desc):
// vector v is fetched right after the rows that used vector v.  Register-level double buffering:
// the L2/global latency of Q blocks and table entries hides under the ~1000 multiplier cycles of
// this product.
template <int K, class IO>
DKG_HD void block_mac(ColAcc<K>& a, const uint32_t (&x)[kpad<K>], uint32_t (&y)[kpad<K>], const IO& io,
                      const typename IO::Prefetch& pf) {
  static_assert(K >= 4, "K must be >= 4");
  constexpr int VW = IO::VW;
#pragma unroll
  for (int j = 0; j < K; j++) {
    // even positions i + j: chain into E, its carry out counted at limb (last i) + j + 2
    {
      const int i0 = j & 1;
      const int last = ((K - 1 - i0) & 1) ? K - 2 : K - 1;
      mad_cc64(a.E[(i0 + j) / 2], x[i0], y[j]);
#pragma unroll
      for (int i = i0 + 2; i < last; i += 2) madc_cc64(a.E[(i + j) / 2], x[i], y[j]);
      madc_cc64_count(a.E[(last + j) / 2], x[last], y[j], a.CE[(last + j + 2 - kpad<K>) / 2]);
    }
    // odd positions: chain into O (O[q] holds limbs (2q+1, 2q+2))
    {
      const int i0 = (j & 1) ^ 1;
      const int last = ((K - 1 - i0) & 1) ? K - 2 : K - 1;
      mad_cc64(a.O[(i0 + j - 1) / 2], x[i0], y[j]);
#pragma unroll
      for (int i = i0 + 2; i < last; i += 2) madc_cc64(a.O[(i + j - 1) / 2], x[i], y[j]);
      madc_cc64_count(a.O[(last + j - 1) / 2], x[last], y[j], a.CO[(last + j + 2 - kodd0<K>) / 2]);
    }
    // one predicated load, no branch: the block product stays a single basic block
    if ((j + 1) % VW == 0 || j == K - 1) io.prefetch_load(pf, j / VW, y);
  }
}

// e <- everything folded together (2K+2 limbs); O, CE, CO cleared.  a.E is left stale: the caller
// edits e and hands it back with acc_load.
template <int K>
DKG_HD void acc_merge(ColAcc<K>& a, uint32_t (&e)[2 * K + 2]) {
  uint32_t o[2 * K - 2];
#pragma unroll
  for (int p = 0; p < K + 1; p++) unpack64(a.E[p], e[2 * p], e[2 * p + 1]);
#pragma unroll
  for (int p = 0; p < K - 1; p++) unpack64(a.O[p], o[2 * p], o[2 * p + 1]);
  add_cc(e[1], o[0]);
#pragma unroll
  for (int p = 1; p < 2 * K - 2; p++) addc_cc(e[p + 1], o[p]);
  addc_cc(e[2 * K - 1], 0);
  addc_cc(e[2 * K], 0);
  addc(e[2 * K + 1], 0);
  // the carry counters, one chain over the limbs K .. 2K (even positions: CE, odd ones: CO)
#pragma unroll
  for (int p = K; p <= 2 * K; p++) {
    const bool even = ((p - kpad<K>) & 1) == 0;
    const uint32_t cnt = even ? a.CE[(p - kpad<K>) / 2] : a.CO[(p - kodd0<K>) / 2];
    if (p == K) add_cc(e[p], cnt); else addc_cc(e[p], cnt);
  }
  addc(e[2 * K + 1], 0);
  acc_clear_side<K>(a);
}

// low block of the accumulated value: (E + (O << 32)) mod 2^(32K); a is not modified
template <int K>
DKG_HD void acc_low(const ColAcc<K>& a, uint32_t (&tl)[K]) {
  uint32_t t[kpad<K>], o[kpad<K>];  // O limbs 0..K-2 (O limb q sits at limb position q+1)
#pragma unroll
  for (int p = 0; p < (K + 1) / 2; p++) unpack64(a.E[p], t[2 * p], t[2 * p + 1]);
#pragma unroll
  for (int p = 0; p < K / 2; p++) unpack64(a.O[p], o[2 * p], o[2 * p + 1]);
  add_cc(t[1], o[0]);
#pragma unroll
  for (int p = 2; p < K; p++) {
    if (p < K - 1) addc_cc(t[p], o[p - 1]); else addc(t[p], o[p - 1]);
  }
#pragma unroll
  for (int p = 0; p < K; p++) tl[p] = t[p];
}

template <int K>
DKG_HD void acc_load(ColAcc<K>& a, const uint32_t (&e)[2 * K + 2]) {
#pragma unroll
  for (int p = 0; p < K + 1; p++) a.E[p] = pack64(e[2 * p], e[2 * p + 1]);
}

// r = x[0..K) * y mod 2^(32K)   (x is the low block of a wider array; pad limb of r cleared)
template <int K, int XN>
DKG_HD void block_mul_lo(uint32_t (&r)[kpad<K>], const uint32_t (&x)[XN], const uint32_t (&y)[kpad<K>]) {
  static_assert(XN >= K, "x too short");
  // E[p] = limb p as seen by the even-position chains, O[q] = limb q + 1 as seen by the odd ones; a
  // product at position i + j <= K - 2 is a full 64-bit accumulate, at position K - 1 its low half
  // only (it ends its chain: the carry out would leave the block)
  uint32_t E[K + 1], O[K + 1];
#pragma unroll
  for (int i = 0; i < K + 1; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
  for (int j = 0; j < K; j++) {
    {
      const int i0 = j & 1;
#pragma unroll
      for (int i = i0; i + j <= K - 1; i += 2) {
        const int pos = i + j;
        if (pos <= K - 2) { if (i == i0) mad_cc(E[pos], E[pos + 1], x[i], y[j]); else madc_cc(E[pos], E[pos + 1], x[i], y[j]); }
        else { if (i == i0) mad_lo(E[pos], x[i], y[j]); else madc_lo(E[pos], x[i], y[j]); }
      }
    }
    {
      const int i0 = (j & 1) ^ 1;
#pragma unroll
      for (int i = i0; i + j <= K - 1; i += 2) {
        const int q = i + j - 1;
        if (q + 1 <= K - 2) { if (i == i0) mad_cc(O[q], O[q + 1], x[i], y[j]); else madc_cc(O[q], O[q + 1], x[i], y[j]); }
        else { if (i == i0) mad_lo(O[q], x[i], y[j]); else madc_lo(O[q], x[i], y[j]); }
      }
    }
  }
  r[0] = E[0];
  r[1] = E[1];
  add_cc(r[1], O[0]);
#pragma unroll
  for (int p = 2; p < K - 1; p++) { r[p] = E[p]; addc_cc(r[p], O[p - 1]); }
  r[K - 1] = E[K - 1];
  addc(r[K - 1], O[K - 2]);
  if (K & 1) r[K] = 0;
}

// Pair schedule of column c of the block product scan: N*Q products, operand products, quotient
// step, in that order for every mode.
//
// Squaring: sum_{i<j} 2 X_i X_j W^(i+j) (W = 2^(32K)) is formed with the doubled operand,
//   sum_{i<j} D_i X_j W^(i+j) + sum_{j>=1} t_{j-1} X_j W^(2j),   D = blocks of 2X mod W^M,
// where t_{j-1} is the top bit of block j-1: doubling the low j blocks of X carries that bit out
// to block position j, and the triangular sum has no pair (j, j) to absorb it.  The second sum is
// the "carry-bit correction" mont_mul adds at the start of column 2j.  The top block is never
// doubled (i <= M-2), so this holds for every X < R.
// Doubled product (MUL2S): 2*X*S = X * (2S) block by block; rectangular, no correction, but 2S must
// fit: S < R/2 (the pair arithmetic's a < 2N <= R/4).
template <int M>
struct ColPlan {
  int lo = 0, span = 0, nxy = 0, nq = 0, total = 0, MODE = 0;
  DKG_HD constexpr ColPlan(int c, int mode) {
    MODE = mode;
    lo = c >= M ? c - M + 1 : 0;
    const int hi = c < M ? c : M - 1;
    span = hi - lo + 1;                               // X_i * Y_{c-i}, i in [lo, hi]
    // squaring: pairs i < c-i (doubled operand) and the middle i == c-i
    nxy = (MODE == MONT_MUL || MODE == MONT_MUL2S) ? span
          : (MODE == MONT_SQR ? span / 2 + (span & 1) : (MODE == MONT_MULADD ? 2 * span : 0));
    nq = (c < M ? c - 1 : M - 1) - lo + 1;            // Q_i * N_{c-i}, i in [lo, ..]
    total = nxy + nq + (c < M ? 1 : 0);               // + the quotient step
  }
  DKG_HD constexpr PairDesc at(int c, int t) const {
    PairDesc d{PAIR_NONE, 0, 0};
    if (t < nq) { d.kind = PAIR_NQ; d.yi = lo + t; d.xi = c - d.yi; }
    else if (t < nq + nxy) {
      const int u = t - nq;
      if (MODE == MONT_MULADD && u >= span) { d.kind = PAIR_SY2; d.xi = lo + (u - span); d.yi = c - d.xi; }
      else {
        d.xi = lo + u; d.yi = c - d.xi;
        d.kind = MODE == MONT_SQR ? (d.xi < d.yi ? PAIR_XX2 : PAIR_XX) : (MODE == MONT_MUL2S ? PAIR_SX2 : PAIR_XY);
      }
    } else d.kind = PAIR_QC;
    return d;
  }
};

// ---- tabulated schedule -------------------------------------------------------------------------
// The pair sequence of one Montgomery product depends only on (M, mode).  Computing "what comes
// next" from ColPlan inside the block-product loop costs ~70 uniform-datapath instructions and 9
// branches per block product; an IO policy with SCHED = true instead reads it from a table
// (one word per pair, all pairs of a mode in execution order, then a PAIR_NONE terminator) that the
// kernel fills once at start with the very same ColPlan.
DKG_HD uint32_t pack_desc(const PairDesc& d) { return (uint32_t)d.kind | ((uint32_t)d.xi << 8) | ((uint32_t)d.yi << 16); }
DKG_HD PairDesc unpack_desc(uint32_t w) {
  PairDesc d;
  d.kind = (int)(w & 0xffu); d.xi = (int)((w >> 8) & 0xffu); d.yi = (int)((w >> 16) & 0xffu);
  return d;
}
constexpr int kSchedModes = 5;
// words of mode `mode`'s list (pairs + terminator) / offset of its first word
template <int M>
DKG_HD constexpr int sched_words(int mode) {
  int n = 1;
  for (int c = 0; c < 2 * M; ++c) n += ColPlan<M>(c, mode).total;
  return n;
}
template <int M>
DKG_HD constexpr int sched_offset(int mode) {
  int off = 0;
  for (int m = 0; m < mode; ++m) off += sched_words<M>(m);
  return off;
}
// entry `i` of the whole table (all modes back to back), for the fill loop
template <int M>
DKG_HD uint32_t sched_entry(int i) {
  for (int mode = 0; mode < kSchedModes; ++mode) {
    const int n = sched_words<M>(mode);
    if (i >= n) { i -= n; continue; }
    for (int c = 0; c < 2 * M; ++c) {
      const ColPlan<M> plan(c, mode);
      if (i < plan.total) return pack_desc(plan.at(c, i));
      i -= plan.total;
    }
    break;
  }
  PairDesc none;
  none.kind = PAIR_NONE; none.xi = 0; none.yi = 0;
  return pack_desc(none);
}

// IO policy (all indices are block indices; r has K limbs; VW = limbs per vector):
//   load_x(i, r)  load_xs(from_s, i, r)  load_xs2(from_s, i, r)  x_limb(l)  load_y(j, r)  load_n(j, r)  load_ninv(r)
//   prefetch_desc(kind, block) -> IO::Prefetch ; prefetch_load(desc, v, r)   (vector v of the block)
//   store_q(i, r) store_x(i, r)
//
// MONT_MUL : X <- X * Y * R^-1 mod N
// MONT_SQR : X <- X * X * R^-1 mod N   (cross block products computed once and doubled)
// MONT_REDC: X <- X * R^-1 mod N
// Result in [0, R); X is overwritten block by block (block c-M is dead when column c starts).
// `MODE` is a run-time (warp-uniform) argument on purpose: all modes share ONE instance of the
// unrolled block product, so the hot loop of an exponentiation (squarings, multiplications and, in
// the pair arithmetic, the two mixed products) stays inside the instruction cache.
template <int K, int M, class IO>
DKG_HD void mont_mul(const IO& io, const int MODE) {
  using Plan = ColPlan<M>;
  ColAcc<K> a;
  uint32_t e[2 * K + 2];  // 32-bit view of the accumulator at column boundaries and events
  uint32_t Tc[K + 2];     // carry-in from the previous column (merged)
#pragma unroll
  for (int i = 0; i < K + 2; i++) Tc[i] = 0;
  acc_clear_side<K>(a);

  // position of the NEXT pair in the tabulated schedule (SCHED IO policies)
  auto sched = io.sched_begin(IO::SCHED ? sched_offset<M>(MODE) + 1 : 0);
  // the schedule word is read ONE block product ahead: its shared-memory latency (load, move to the
  // uniform datapath, decode) hides under the running block product instead of delaying the next
  uint32_t sched_ahead = 0;
  if constexpr (IO::SCHED) sched_ahead = io.sched_word(sched, 0);

  constexpr int KP = kpad<K>;
  uint32_t xb[KP], yb[KP];
  // operands of the very first block product (everything after it is prefetched)
  {
    const PairDesc d = Plan(0, MODE).at(0, 0);
    if (d.kind == PAIR_XY) { io.load_x(d.xi, xb); io.load_y(d.yi, yb); }
    else if (d.kind == PAIR_XX) { io.load_x(d.xi, xb); io.load_x(d.yi, yb); }
    else if (d.kind == PAIR_SX2) { io.load_xs2(true, d.xi, xb); io.load_x(d.yi, yb); }
  }

  for (int c = 0; c < 2 * M; ++c) {
    const Plan plan(c, MODE);

    // seed the accumulator with the carry-in
#pragma unroll
    for (int p = 0; p < K + 2; p++) e[p] = Tc[p];
#pragma unroll
    for (int p = K + 2; p < 2 * K + 2; p++) e[p] = 0;

    if (MODE == MONT_REDC && c < M) {
      uint32_t tb[KP];
      io.load_x(c, tb);
      add_cc(e[0], tb[0]);
#pragma unroll
      for (int p = 1; p < K; p++) addc_cc(e[p], tb[p]);
#pragma unroll
      for (int p = K; p <= 2 * K; p++) addc_cc(e[p], 0);
      addc(e[2 * K + 1], 0);
    }
    if (MODE == MONT_SQR && (c & 1) == 0 && c >= 2) {
      // carry-bit correction of the doubled operand (see ColPlan): + t_{j-1} * X_j at column 2j
      const int j = c >> 1;
      const uint32_t mask = 0u - (io.x_top_limb(j - 1) >> 31);
      uint32_t tb[KP];
      io.load_x(j, tb);
      add_cc(e[0], tb[0] & mask);
#pragma unroll
      for (int p = 1; p < K; p++) addc_cc(e[p], tb[p] & mask);
#pragma unroll
      for (int p = K; p <= 2 * K; p++) addc_cc(e[p], 0);
      addc(e[2 * K + 1], 0);
    }
    // The column's block products run in two stretches around the quotient step, through the SAME
    // inner loop, whose body is nothing but the block product and its operand traffic; the
    // accumulator enters it once per column from the 32-bit view e (acc_load) and is folded back
    // once (acc_merge) -- the quotient step only reads its low block.
    const int t_quot = c < M ? plan.total - 1 : -1;
    int t = 0;
    while (t < plan.total) {
      if (t == 0) acc_load<K>(a, e);
      if (t == t_quot) {
        // quotient block: Q_c = T_low * (-N^-1) mod 2^(32K); Q_c * N_0 then clears T_low.  T_low
        // needs only the low blocks of E and O (the counters weigh 2^(32K) and more), so the
        // accumulator itself stays as it is, in carry-save form.
        uint32_t tl[K];
        acc_low<K>(a, tl);
        io.load_ninv(yb);
        block_mul_lo<K>(xb, tl, yb);
        io.store_q(c, xb);
        io.load_n(0, yb);
      }
      int t_end = plan.total;
      if (t < t_quot) t_end = t_quot;
      for (; t < t_end; ++t) {
        // what comes next (possibly in the next column): its y operand is prefetched behind this
        // block product, its x operand loaded right after
        PairDesc nx;
        if constexpr (IO::SCHED) {
          nx = unpack_desc(sched_ahead);
          sched = io.sched_next(sched);
          sched_ahead = io.sched_word(sched, 0);   // (one word past the list's terminator at the very end: unused)
        } else {
          nx.kind = PAIR_NONE; nx.xi = 0; nx.yi = 0;
          if (t + 1 < plan.total) nx = plan.at(c, t + 1);
          else if (c + 1 < 2 * M) {
            const Plan np(c + 1, MODE);   // only the very last column can be empty
            if (np.total > 0) nx = np.at(c + 1, 0);
          }
        }
        block_mac<K>(a, xb, yb, io, io.prefetch_desc(nx.kind, nx.yi));
        if (nx.kind == PAIR_NQ) io.load_n(nx.xi, xb);
        else if (nx.kind == PAIR_XX2 || nx.kind == PAIR_SX2) io.load_xs2(nx.kind == PAIR_SX2, nx.xi, xb);
        else if (nx.kind != PAIR_QC && nx.kind != PAIR_NONE) io.load_xs(nx.kind == PAIR_SY2, nx.xi, xb);
      }
    }
    if (plan.total > 0) acc_merge<K>(a, e);

    if (c >= M) {
      uint32_t ob[KP];
#pragma unroll
      for (int p = 0; p < K; p++) ob[p] = e[p];
      if (K & 1) ob[K] = 0;
      io.store_x(c - M, ob);
    }
#pragma unroll
    for (int p = 0; p < K + 2; p++) Tc[p] = e[p + K];
  }

  if (io.never()) Tc[0] = pipe_ballast(Tc[0], Tc[1]);
  // result = Tc[0]*R + X < R + N: subtract N once iff the carry limb is set.  Skipped when no lane of
  // the warp has it set -- always in the pair arithmetic, whose products stay below 5N <= R
  //  -- instead of a masked pass over X and N.
  if (!io.any_lane(Tc[0])) return;
  const uint32_t mask = 0u - Tc[0];
  uint32_t borrow = 0;
  for (int b = 0; b < M; ++b) {
    uint32_t tb[kpad<K>], nb[kpad<K>];
    io.load_x(b, tb);
    io.load_n(b, nb);
#pragma unroll
    for (int p = 0; p < K; p++) {
      const uint64_t d = (uint64_t)tb[p] - (nb[p] & mask) - borrow;
      tb[p] = (uint32_t)d;
      borrow = (uint32_t)(d >> 63);
    }
    io.store_x(b, tb);
  }
}

// X >= N ?  (returns 1/0).  Scans all blocks, no early exit.
template <int K, int M, class IO>
DKG_HD uint32_t geq_n(const IO& io) {
  uint32_t borrow = 0;
  for (int b = 0; b < M; ++b) {
    uint32_t xb[kpad<K>], nb[kpad<K>];
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
    uint32_t xb[kpad<K>], nb[kpad<K>];
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
