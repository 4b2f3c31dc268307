// Exponentiation modulo N^2 with arithmetic modulo N only ("pair arithmetic").
//
// The modulus of threshold-Paillier partial decryption and of the encryption randomness is a
// perfect square whose root N is public.  An element x of Z_{N^2} is held as a pair (a, b) of
// L-limb integers (L = limbs of N plus >= 3 spare bits, R = 2^(32L)) with
//        x * R  =  a + b * N * R^-1      (mod N^2),        0 <= a < 2N,  0 <= b < R.
// Because (N R^-1)^2 = 0 (mod N^2):
//   square  : a' = REDC(a^2)       with Montgomery quotient m  (a^2 = a' R - m N exactly)
//             b' = REDC(2 a b) - m                             (mod N)
//   multiply: a' = REDC(a c)       with quotient m
//             b' = REDC(a d + b c) - m                         (mod N)
// where REDC is Montgomery reduction modulo N.  A squaring costs one L-limb Montgomery squaring
// and one L-limb Montgomery multiplication instead of one 2L-limb Montgomery squaring:
// 15.9k instead of 26.7k wide multiply-accumulates at 2048-bit N (1.68x fewer), a multiplication
// 21.5k instead of 33.9k.  Results are bit-identical to direct arithmetic modulo N^2 (the exit
// step returns the canonical residue).  Derivation and bounds: DESIGN.md section 2.8; Python model:
// tests/test_pair_model.py.
//
// "- m" is applied after the reduction as b'' + (R - m), then brought back into [0, R) with at
// most one masked addition of (-R mod N) and one masked subtraction of N; b only matters mod N.
#pragma once
#include "dkg_modexp.cuh"
#include "dkg_nsq_params_fwd.h"

namespace dkg {


// b <- b - m (mod N), kept in [0, R).  Bv: this lane's vector 0 of B in shared memory, mq: the
// quotient blocks just written by the a-component's reduction (global, same vector layout),
// N / Dneg: shared, CTA-uniform.  Whole vectors per access: conflict-free like the block loads.
template <int K, int M>
__device__ __forceinline__ void pair_fixup(typename VecSel<kpad<K>>::T* Bv, const typename VecSel<kpad<K>>::T* mq,
                                           const uint32_t* Ns, const uint32_t* Dneg) {
  constexpr int KP = kpad<K>;
  using V = typename VecSel<KP>::T;
  constexpr int VW = VecSel<KP>::VW;
  constexpr int LV = KP * M / VW;
  // odd K: the last limb of every slot is a pad (zero in b, N and Dneg); the carry steps over it
  auto is_pad = [](int v, int k) -> bool { return (K & 1) && (v * VW + k) % KP == K; };
  // S = b + (R - m) = b + ~m + 1
  uint32_t carry = 1;
  for (int v = 0; v < LV; ++v) {
    uint32_t bb[VW], mm[VW];
    unpack(Bv[v * 32], bb);
    unpack(mq[(size_t)v * 32], mm);
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      if (is_pad(v, k)) { bb[k] = 0; continue; }
      const uint64_t s = (uint64_t)bb[k] + (uint32_t)~mm[k] + carry;
      bb[k] = (uint32_t)s;
      carry = (uint32_t)(s >> 32);
    }
    V o; pack(o, bb); Bv[v * 32] = o;
  }
  // carry set: S >= R, dropping R leaves b - m >= 0.  Otherwise S = b - m + R: add (-R mod N),
  // and if that passes R take N off once.
  const uint32_t need = carry ^ 1u;
  if (__any_sync(0xffffffffu, need)) {
    const uint32_t mask = 0u - need;
    uint32_t c2 = 0;
    for (int v = 0; v < LV; ++v) {
      uint32_t bb[VW];
      unpack(Bv[v * 32], bb);
#pragma unroll
      for (int k = 0; k < VW; ++k) {
        if (is_pad(v, k)) continue;
        const uint64_t s = (uint64_t)bb[k] + (Dneg[v * VW + k] & mask) + c2;
        bb[k] = (uint32_t)s;
        c2 = (uint32_t)(s >> 32);
      }
      V o; pack(o, bb); Bv[v * 32] = o;
    }
    if (__any_sync(0xffffffffu, c2)) {
      const uint32_t mask2 = 0u - c2;
      uint32_t borrow = 0;
      for (int v = 0; v < LV; ++v) {
        uint32_t bb[VW];
        unpack(Bv[v * 32], bb);
#pragma unroll
        for (int k = 0; k < VW; ++k) {
          if (is_pad(v, k)) continue;
          const uint64_t d = (uint64_t)bb[k] - (Ns[v * VW + k] & mask2) - borrow;
          bb[k] = (uint32_t)d;
          borrow = (uint32_t)(d >> 63);
        }
        V o; pack(o, bb); Bv[v * 32] = o;
      }
    }
  }
}

// ---- inversion in the pair domain ------------------------------------------------------------------
// (A, B) <- Montgomery pair of x^-1 from the Montgomery pair of x (negative exponents: the reference
// inverts with mod_inv, paillier_shared_key.py:89-91).  g = a^-1 mod N by ONE binary GCD on the a
// component only (half the limbs of N^2: a quarter of the work of inverting modulo N^2), then the
// pair of y0 = g R -- an inverse of x modulo N -- and one Newton step y0 (2 - x y0) in the pair
// domain (precision N -> N^2).  Per instance, so the "not invertible" status is exact per element.
// Integer model with the bounds: tests/test_pair_model.py::pair_inverse.
//   work: global scratch, one pair slot (the GCD arrays u, v) then 3 pair slots (x, y0, a temporary),
//   lane offset NOT applied; the other two GCD arrays sit where A and B are (shared or global).
// Returns 1 if x is not a unit (the pair is then unspecified), else 0.
template <int K, int M, class PairMul>
__device__ __forceinline__ uint32_t pair_invert(typename VecSel<kpad<K>>::T* Aw, typename VecSel<kpad<K>>::T* Bw,
                                                uint32_t* work, const uint32_t* U32, const uint32_t* Nd,
                                                const uint32_t* two_ab, const typename VecSel<kpad<K>>::T* R2Ar,
                                                const typename VecSel<kpad<K>>::T* R2Br, int lane, PairMul&& pair_mul) {
  constexpr int KP = kpad<K>;
  using V = typename VecSel<KP>::T;
  constexpr int VW = VecSel<KP>::VW;
  constexpr int Lp = KP * M, La = K * M, LV = Lp / VW, KVB = KP / VW;
  uint32_t* Aw32 = reinterpret_cast<uint32_t*>(Aw);
  uint32_t* Bw32 = reinterpret_cast<uint32_t*>(Bw);
  uint32_t* pu = work + lane;
  uint32_t* pv = pu + La * 32;
  V* slots = reinterpret_cast<V*>(work + 2 * Lp * 32);   // (one pair slot for u and v: 2 La <= 2 Lp)
  auto slot_a = [&](int s) -> V* { return slots + (size_t)s * 2 * LV * 32 + lane; };
  auto slot_b = [&](int s) -> V* { return slots + ((size_t)s * 2 + 1) * LV * 32 + lane; };
  auto sidx_of = [&](int l) -> int { const int sl = (l / K) * KP + l % K; return ((sl / VW) * 32 + lane) * VW + sl % VW; };
  // x -> slot 0; a (dense) -> u
  for (int v = 0; v < LV; ++v) { slot_a(0)[(size_t)v * 32] = Aw[v * 32 + lane]; slot_b(0)[(size_t)v * 32] = Bw[v * 32 + lane]; }
  for (int l = 0; l < La; ++l) pu[l * 32] = Aw32[sidx_of(l)];
  uint32_t bad = 0;
  const uint32_t* g = mod_inverse_arrays(pu, pv, Aw32 + lane, Bw32 + lane, Nd, U32[Lp], La, &bad);
  // g sits in one of the four arrays: through u (free by now) into the plain pair (g, 0)
  if (g != pu) for (int l = 0; l < La; ++l) pu[l * 32] = g[l * 32];
  {
    const V zero = V();
    for (int v = 0; v < LV; ++v) { Aw[v * 32 + lane] = zero; Bw[v * 32 + lane] = zero; }
  }
  for (int l = 0; l < La; ++l) Aw32[sidx_of(l)] = pu[l * 32];
  pair_mul(R2Ar, R2Br);
  pair_mul(R2Ar, R2Br);                                   // pair of g R = y0
  for (int v = 0; v < LV; ++v) { slot_a(1)[(size_t)v * 32] = Aw[v * 32 + lane]; slot_b(1)[(size_t)v * 32] = Bw[v * 32 + lane]; }
  pair_mul(slot_a(0), slot_b(0));                         // pair of x y0 = 1 (mod N)
  // (A, B) <- 2 - (A, B):  A <- TWOA - A (no borrow out: TWOA >= 2N > A),  B <- TWOB - B (mod N)
  const uint32_t* twoa = two_ab;
  const uint32_t* twob = two_ab + Lp;
  {
    uint32_t borrow = 0;
    for (int v = 0, q = 0; v < LV; ++v, q = (q + 1 == KVB ? 0 : q + 1)) {
      uint32_t aa[VW], bb[VW], tb[VW];
      unpack(Aw[v * 32 + lane], aa);
      unpack(Bw[v * 32 + lane], bb);
#pragma unroll
      for (int k = 0; k < VW; ++k) {
        tb[k] = twob[v * VW + k];
        if ((K & 1) && q == KVB - 1 && k == VW - 1) { aa[k] = 0; continue; }
        const uint64_t d = (uint64_t)twoa[v * VW + k] - aa[k] - borrow;
        aa[k] = (uint32_t)d;
        borrow = (uint32_t)(d >> 63);
      }
      V oa, ob, ot; pack(oa, aa); pack(ob, bb); pack(ot, tb);
      Aw[v * 32 + lane] = oa;
      slot_b(2)[(size_t)v * 32] = ob;                     // the old b: what pair_fixup subtracts
      Bw[v * 32 + lane] = ot;
    }
  }
  pair_fixup<K, M>(Bw + lane, slot_b(2), U32, U32 + Lp + KP);
  pair_mul(slot_a(1), slot_b(1));                         // y0 (2 - x y0)
  return bad;
}

template <int M>
constexpr int kNsqSchedWords = sched_offset<M>(kSchedModes);

// BG: the b component of the running pair lives in the warp's GLOBAL scratch (after the quotient
// blocks, L2 resident) instead of shared memory, which then holds a only: twice as many warps per SM
// for wide keys (6 -> 12 at key_length 4096), where shared memory, not registers, caps the occupancy.
template <int K, int M, bool BG = false>
__global__ void __launch_bounds__(DKG_MAX_THREADS, 1) modexp_nsq_kernel(const NsqParams p) {
  static_assert(kNsqSchedWords<M> == sched_total_words_closed(M), "schedule table size: host formula out of sync with ColPlan");
  constexpr int KP = kpad<K>;          // slot of one block in memory (odd K: one zero pad limb)
  using V = typename VecSel<KP>::T;
  constexpr int VW = VecSel<KP>::VW;
  constexpr int Lp = KP * M;           // limbs of a number in memory
  constexpr int La = K * M;            // limbs of a number in the pairs_in / pairs_out rows (R = 2^(32 La))
  constexpr int LV = Lp / VW;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* U32 = reinterpret_cast<uint32_t*>(smem_raw);   // N[Lp] | NINV[KP] | DNEG[Lp]  (slot layout)
  constexpr int UNI0 = ((2 * Lp + KP + La) * 4 + 15) / 16 * 16;   // ... | N once more, dense (the GCD of pair_invert)
  // ... | schedule table (kNsqSchedWords<M> words, see fill_schedule)
  constexpr int UNI = UNI0 + (kNsqSchedWords<M> * 4 + 15) / 16 * 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * Lp + KP; i += blockDim.x) U32[i] = p.consts[i];
  for (int i = threadIdx.x; i < La; i += blockDim.x) U32[2 * Lp + KP + i] = p.consts[(i / K) * KP + i % K];
  fill_schedule<M>(reinterpret_cast<uint32_t*>(smem_raw + UNI0));
  __syncthreads();
  const uint32_t* Ns32 = U32;
  const uint32_t* Dneg = U32 + Lp + KP;
  const uint32_t* Nd = U32 + 2 * Lp + KP;

  V* Aw = reinterpret_cast<V*>(smem_raw + UNI) + (size_t)warp * (BG ? 1 : 2) * LV * 32;
  uint32_t* Aw32 = reinterpret_cast<uint32_t*>(Aw);
  const V* Cg = reinterpret_cast<const V*>(p.consts + 2 * Lp + KP);
  const V* ONEA = Cg + 2 * LV, *ONEB = Cg + 3 * LV;
  // the six constants again, lane-replicated ([v][lane]) so that they can be multiplication operands
  const V* Crep = Cg + 6 * LV + lane;
  const V* R2Ar = Crep, *R2Br = Crep + (size_t)LV * 32, *ONEAr = Crep + (size_t)2 * LV * 32,
          *ONEBr = Crep + (size_t)3 * LV * 32,
          *PLAIN1r = Crep + (size_t)4 * LV * 32, *ZEROr = Crep + (size_t)5 * LV * 32;
  const uint32_t* two_ab = p.consts + (2 * Lp + KP) + 6 * Lp + 6 * Lp * 32;   // TWOA | TWOB, slot layout

  const unsigned gwarp = blockIdx.x * nwarps + warp;
  uint32_t* scratch32 = p.scratch + (size_t)gwarp * p.scratch_per_warp;
  V* tab = reinterpret_cast<V*>(scratch32);   // entry d (1-based): a at ((d-1)*2)*LV*32, b right after
  V* Qg = reinterpret_cast<V*>(scratch32 + p.scratch_q_offset);
  V* Bw = BG ? Qg + LV * 32 : Aw + LV * 32;
  uint32_t* Bw32 = reinterpret_cast<uint32_t*>(Bw);

  const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(Aw + lane);
  const uint32_t b_s = BG ? 0u : (uint32_t)__cvta_generic_to_shared(Bw + lane);
  WarpIO<K, M, false, true, BG> io;
  io.sched_s = (uint32_t)__cvta_generic_to_shared(smem_raw + UNI0);
  io.ns = (uint32_t)__cvta_generic_to_shared(Ns32);
  io.nis = (uint32_t)__cvta_generic_to_shared(Ns32 + Lp);
  io.Qg = Qg + lane;

  // (A, B) <- (A, B) * (c, d): c, d in global memory, lane layout
  auto pair_mul = [&](const V* c, const V* d) {
    io.xs = b_s; io.ss = a_s; io.Y = c; io.Y2 = d;
    if constexpr (BG) io.xg = Bw + lane;
    mont_call<K, M, MONT_MULADD, false, true, BG>(io);             // B <- REDC(B c + A d)
    io.xs = a_s;
    if constexpr (BG) io.xg = nullptr;
    mont_call<K, M, MONT_MUL, false, true, BG>(io);                // A <- REDC(A c), quotient m in Q
    pair_fixup<K, M>(Bw + lane, Qg + lane, Ns32, Dneg);
  };
  auto pair_sqr = [&]() {
    io.xs = b_s; io.ss = a_s;
    if constexpr (BG) io.xg = Bw + lane;
    mont_call<K, M, MONT_MUL2S, false, true, BG>(io);              // B <- REDC(2 B A)
    io.xs = a_s;
    if constexpr (BG) io.xg = nullptr;
    mont_call<K, M, MONT_SQR, false, true, BG>(io);                // A <- REDC(A^2), quotient m in Q
    pair_fixup<K, M>(Bw + lane, Qg + lane, Ns32, Dneg);
  };

  const unsigned long long ngroups = (p.count + 31ull) / 32ull;
  for (;;) {
    unsigned int g = 0;
    if (lane == 0) g = atomicAdd(p.counter, 1u);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= ngroups) break;
    const unsigned long long first = (unsigned long long)g * 32ull;
    const int cnt = (int)((p.count - first) < 32ull ? (p.count - first) : 32ull);

    // rows of 2*Lp limbs: a then b; idle lanes get the pair of 1
    for (int r = 0; r < 32; ++r) {
      const uint32_t* row = p.pairs_in + (first + (unsigned long long)r) * (unsigned long long)(2 * La);
      for (int l = lane; l < Lp; l += 32) {          // l: limb in slot layout, src: the dense limb (or a pad)
        const int i = ((l / VW) * 32 + r) * VW + (l % VW);
        const int src = (l / KP) * K + (l % KP);
        const bool pad = (l % KP) >= K;
        Aw32[i] = pad ? 0u : ((r < cnt) ? row[src] : (l == 0 ? 1u : 0u));
        Bw32[i] = (pad || r >= cnt) ? 0u : row[La + src];
      }
    }
    __syncwarp();

    pair_mul(R2Ar, R2Br);   // into the Montgomery domain: value c * R
    uint32_t st = 0;
    if (p.negative)
      st = pair_invert<K, M>(Aw, Bw, scratch32 + p.inv_slot * (unsigned long long)(2 * Lp * 32), U32, Nd, two_ab, R2Ar, R2Br, lane, pair_mul);
    if (p.nops == 0) {
      for (int v = 0; v < LV; ++v) { Aw[v * 32 + lane] = ONEA[v]; Bw[v * 32 + lane] = ONEB[v]; }
    } else {
      // window table: entry k = c^(k+1) (or the odd powers c^(2k+1), c^2 after the last), a at
      // (2k)*LV*32, b right after
      const int tn = p.tab_entries;
      auto tab_a = [&](int k) -> V* { return tab + (size_t)k * 2 * LV * 32 + lane; };
      auto tab_b = [&](int k) -> V* { return tab + ((size_t)k * 2 + 1) * LV * 32 + lane; };
      for (int v = 0; v < LV; ++v) { tab_a(0)[(size_t)v * 32] = Aw[v * 32 + lane]; tab_b(0)[(size_t)v * 32] = Bw[v * 32 + lane]; }
      if (tn > 1) {
        V* sa = tab_a(0); V* sb = tab_b(0);   // multiply by c ...
        if (p.table_odd) {                    // ... or by c^2
          pair_sqr();
          sa = tab_a(tn); sb = tab_b(tn);
          for (int v = 0; v < LV; ++v) {
            sa[(size_t)v * 32] = Aw[v * 32 + lane]; sb[(size_t)v * 32] = Bw[v * 32 + lane];
            Aw[v * 32 + lane] = tab_a(0)[(size_t)v * 32]; Bw[v * 32 + lane] = tab_b(0)[(size_t)v * 32];
          }
        }
        for (int k = 1; k < tn; ++k) {
          pair_mul(sa, sb);
          V* da = tab_a(k); V* db = tab_b(k);
          for (int v = 0; v < LV; ++v) { da[(size_t)v * 32] = Aw[v * 32 + lane]; db[(size_t)v * 32] = Bw[v * 32 + lane]; }
        }
      }
      // Constant-time table access (p.ct_table, fixed windows only): instead of reading entry
      // `idx`, scan ALL entries (and the Montgomery one, digit 0) in a fixed order and keep the
      // wanted one under a mask, into the spare slot after the table: the addresses touched no
      // longer depend on the exponent's digits.
      auto ct_select = [&](uint32_t idx) {
        V* da = tab_a(tn); V* db = tab_b(tn);
        const uint32_t m0 = 0u - (uint32_t)(idx == 0xfeu);
        for (int v = 0; v < LV; ++v) {
          uint32_t aa[VW], bb[VW], tt[VW];
          unpack(ONEAr[(size_t)v * 32], aa); unpack(ONEBr[(size_t)v * 32], bb);
#pragma unroll
          for (int q = 0; q < VW; ++q) { aa[q] &= m0; bb[q] &= m0; }
          for (int k = 0; k < tn; ++k) {
            const uint32_t mk = 0u - (uint32_t)((uint32_t)k == idx);
            unpack(tab_a(k)[(size_t)v * 32], tt);
#pragma unroll
            for (int q = 0; q < VW; ++q) aa[q] |= tt[q] & mk;
            unpack(tab_b(k)[(size_t)v * 32], tt);
#pragma unroll
            for (int q = 0; q < VW; ++q) bb[q] |= tt[q] & mk;
          }
          V oa, ob; pack(oa, aa); pack(ob, bb);
          da[(size_t)v * 32] = oa; db[(size_t)v * 32] = ob;
        }
      };
      {
        int k0 = (int)(p.ops[0] & 0xffu);
        if (p.ct_table) { ct_select((uint32_t)k0); k0 = tn; }
        const V* sa = tab_a(k0); const V* sb = tab_b(k0);
        for (int v = 0; v < LV; ++v) { Aw[v * 32 + lane] = sa[(size_t)v * 32]; Bw[v * 32 + lane] = sb[(size_t)v * 32]; }
      }
      for (int t = 1; t < p.nops; ++t) {
        const uint32_t op = p.ops[t];
        for (uint32_t s = op >> 8; s > 0; --s) pair_sqr();
        const uint32_t idx = op & 0xffu;
        if (p.ct_table) { ct_select(idx); pair_mul(tab_a(tn), tab_b(tn)); }
        else if (idx == 0xfeu) pair_mul(ONEAr, ONEBr);
        else if (idx != 0xffu) pair_mul(tab_a((int)idx), tab_b((int)idx));
      }
    }
    pair_mul(PLAIN1r, ZEROr);   // out of the Montgomery domain: the pair now stands for the result itself
    if (p.negative) {
      if (st) {                 // not a unit: zero row, status 1 (what mod_inv's ZeroDivisionError becomes)
        const V zero = V();
        for (int v = 0; v < LV; ++v) { Aw[v * 32 + lane] = zero; Bw[v * 32 + lane] = zero; }
      }
      if (p.status != nullptr && lane < cnt) p.status[first + lane] = (uint8_t)st;
    }
    __syncwarp();

    for (int r = 0; r < cnt; ++r) {
      uint32_t* row = p.pairs_out + (first + (unsigned long long)r) * (unsigned long long)(2 * La);
      for (int l = lane; l < Lp; l += 32) {
        if ((l % KP) >= K) continue;
        const int i = ((l / VW) * 32 + r) * VW + (l % VW);
        const int dst = (l / KP) * K + (l % KP);
        row[dst] = Aw32[i];
        row[La + dst] = Bw32[i];
      }
    }
    __syncwarp();
  }
}

// ---- several exponents, one base: shared squaring chain ------------------------------------------
// When the shares of several parties sit on one device (the reference's in-process parties,
// distributed_keygen.py:203-226; the threshold context of this library), their partial decryptions
// c^(e_1), ..., c^(e_P) of the SAME ciphertext need only ONE chain of squarings.  Right to left in
// base D = 2^w:  c^(e_p) = prod_k (c^(D^k))^(digit_pk) = prod_{d=1}^{D-1} (A_pd)^d  with the buckets
// A_pd = prod_{k: digit_pk = d} c^(D^k).  Per window: w squarings of the running power (shared by
// all parties) and ONE multiplication per party into the bucket its digit names (digit 0 goes to a
// dummy bucket: the operation sequence does not depend on the exponents).  The buckets are folded
// with the running-product trick, T <- T * A_d, S <- S * T for d = D-2 .. 1: 2 (D - 2)
// multiplications per party.  At 4190-bit exponents and w = 6: 4182 squarings + P * 822
// multiplications instead of P * (4186 squarings + 724 multiplications): 2.05x fewer wide
// multiply-accumulates for P = 3, 2.6x for P = 5.  Every party's result is the same canonical
// residue as its own exponentiation would give (negative exponents: the caller inverts the RESULT,
// (c^-1)^|e| = (c^|e|)^-1).
//
// The Montgomery product overwrites its shared-memory operand, so the running power is parked in
// the warp's global scratch around every bucket multiplication (2 P + 1 copies of 2 Lp limbs per
// window against 6 squarings + P multiplications: < 1 % of the instructions).
template <int K, int M, bool BG = false>
__global__ void __launch_bounds__(DKG_MAX_THREADS, 1) modexp_nsq_multi_kernel(const NsqMultiParams p) {
  constexpr int KP = kpad<K>;          // slot of one block in memory (odd K: one zero pad limb)
  using V = typename VecSel<KP>::T;
  constexpr int VW = VecSel<KP>::VW;
  constexpr int Lp = KP * M;           // limbs of a number in memory
  constexpr int La = K * M;            // limbs of a number in the pairs_in / pairs_out rows (R = 2^(32 La))
  constexpr int LV = Lp / VW;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* U32 = reinterpret_cast<uint32_t*>(smem_raw);   // same uniform area as modexp_nsq_kernel
  constexpr int UNI0 = ((2 * Lp + KP + La) * 4 + 15) / 16 * 16;   // ... | N once more, dense (the GCD of pair_invert)
  constexpr int UNI = UNI0 + (kNsqSchedWords<M> * 4 + 15) / 16 * 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * Lp + KP; i += blockDim.x) U32[i] = p.consts[i];
  for (int i = threadIdx.x; i < La; i += blockDim.x) U32[2 * Lp + KP + i] = p.consts[(i / K) * KP + i % K];
  fill_schedule<M>(reinterpret_cast<uint32_t*>(smem_raw + UNI0));
  __syncthreads();
  const uint32_t* Ns32 = U32;
  const uint32_t* Dneg = U32 + Lp + KP;
  const uint32_t* Nd = U32 + 2 * Lp + KP;

  V* Aw = reinterpret_cast<V*>(smem_raw + UNI) + (size_t)warp * (BG ? 1 : 2) * LV * 32;
  uint32_t* Aw32 = reinterpret_cast<uint32_t*>(Aw);
  const V* Cg = reinterpret_cast<const V*>(p.consts + 2 * Lp + KP);
  const V* Crep = Cg + 6 * LV + lane;
  const V* R2Ar = Crep, *R2Br = Crep + (size_t)LV * 32, *ONEAr = Crep + (size_t)2 * LV * 32,
          *ONEBr = Crep + (size_t)3 * LV * 32,
          *PLAIN1r = Crep + (size_t)4 * LV * 32, *ZEROr = Crep + (size_t)5 * LV * 32;
  const uint32_t* two_ab = p.consts + (2 * Lp + KP) + 6 * Lp + 6 * Lp * 32;   // TWOA | TWOB, slot layout

  const unsigned gwarp = blockIdx.x * nwarps + warp;
  uint32_t* scratch32 = p.scratch + (size_t)gwarp * p.scratch_per_warp;
  V* slots = reinterpret_cast<V*>(scratch32);   // pair s: a at (2s)*LV*32, b right after
  V* Qg = reinterpret_cast<V*>(scratch32 + p.scratch_q_offset);
  V* Bw = BG ? Qg + LV * 32 : Aw + LV * 32;
  uint32_t* Bw32 = reinterpret_cast<uint32_t*>(Bw);
  const int D = 1 << p.wbits;
  auto slot_a = [&](int s) -> V* { return slots + (size_t)s * 2 * LV * 32 + lane; };
  auto slot_b = [&](int s) -> V* { return slots + ((size_t)s * 2 + 1) * LV * 32 + lane; };
  const int CUR = p.nparties * D, ACC = CUR + 1;   // parked running power / bucket accumulator S

  const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(Aw + lane);
  const uint32_t b_s = BG ? 0u : (uint32_t)__cvta_generic_to_shared(Bw + lane);
  WarpIO<K, M, false, true, BG> io;
  io.sched_s = (uint32_t)__cvta_generic_to_shared(smem_raw + UNI0);
  io.ns = (uint32_t)__cvta_generic_to_shared(Ns32);
  io.nis = (uint32_t)__cvta_generic_to_shared(Ns32 + Lp);
  io.Qg = Qg + lane;

  auto pair_mul = [&](const V* c, const V* d) {
    io.xs = b_s; io.ss = a_s; io.Y = c; io.Y2 = d;
    if constexpr (BG) io.xg = Bw + lane;
    mont_call<K, M, MONT_MULADD, false, true, BG>(io);
    io.xs = a_s;
    if constexpr (BG) io.xg = nullptr;
    mont_call<K, M, MONT_MUL, false, true, BG>(io);
    pair_fixup<K, M>(Bw + lane, Qg + lane, Ns32, Dneg);
  };
  auto pair_sqr = [&]() {
    io.xs = b_s; io.ss = a_s;
    if constexpr (BG) io.xg = Bw + lane;
    mont_call<K, M, MONT_MUL2S, false, true, BG>(io);
    io.xs = a_s;
    if constexpr (BG) io.xg = nullptr;
    mont_call<K, M, MONT_SQR, false, true, BG>(io);
    pair_fixup<K, M>(Bw + lane, Qg + lane, Ns32, Dneg);
  };
  auto park = [&](int s) {      // slot s <- (A, B)
    V* da = slot_a(s); V* db = slot_b(s);
#pragma unroll 5
    for (int v = 0; v < LV; ++v) { da[(size_t)v * 32] = Aw[v * 32 + lane]; db[(size_t)v * 32] = Bw[v * 32 + lane]; }
  };
  auto fetch = [&](int s) {     // (A, B) <- slot s
    const V* sa = slot_a(s); const V* sb = slot_b(s);
#pragma unroll 5
    for (int v = 0; v < LV; ++v) { Aw[v * 32 + lane] = sa[(size_t)v * 32]; Bw[v * 32 + lane] = sb[(size_t)v * 32]; }
  };

  const unsigned long long ngroups = (p.count + 31ull) / 32ull;
  for (;;) {
    unsigned int g = 0;
    if (lane == 0) g = atomicAdd(p.counter, 1u);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= ngroups) break;
    const unsigned long long first = (unsigned long long)g * 32ull;
    const int cnt = (int)((p.count - first) < 32ull ? (p.count - first) : 32ull);

    for (int r = 0; r < 32; ++r) {
      const uint32_t* row = p.pairs_in + (first + (unsigned long long)r) * (unsigned long long)(2 * La);
      for (int l = lane; l < Lp; l += 32) {          // l: limb in slot layout, src: the dense limb (or a pad)
        const int i = ((l / VW) * 32 + r) * VW + (l % VW);
        const int src = (l / KP) * K + (l % KP);
        const bool pad = (l % KP) >= K;
        Aw32[i] = pad ? 0u : ((r < cnt) ? row[src] : (l == 0 ? 1u : 0u));
        Bw32[i] = (pad || r >= cnt) ? 0u : row[La + src];
      }
    }
    __syncwarp();
    pair_mul(R2Ar, R2Br);   // into the Montgomery domain

    // every bucket starts as the Montgomery one
    for (int s = 0; s < p.nparties * D; ++s) {
      V* da = slot_a(s); V* db = slot_b(s);
#pragma unroll 5
      for (int v = 0; v < LV; ++v) { da[(size_t)v * 32] = ONEAr[(size_t)v * 32]; db[(size_t)v * 32] = ONEBr[(size_t)v * 32]; }
    }

    for (int k = 0; k < p.nwin; ++k) {
      if (k > 0) for (int s = 0; s < p.wbits; ++s) pair_sqr();
      park(CUR);
      for (int q = 0; q < p.nparties; ++q) {
        const int s = q * D + (int)p.digits[(size_t)q * p.nwin + k];
        pair_mul(slot_a(s), slot_b(s));
        park(s);
        if (q + 1 < p.nparties || k + 1 < p.nwin) fetch(CUR);
      }
    }

    for (int q = 0; q < p.nparties; ++q) {
      fetch(q * D + D - 1);
      for (int d = D - 2; d >= 1; --d) {
        if (d == D - 2) park(ACC);          // S = T = A_{D-1}
        pair_mul(slot_a(q * D + d), slot_b(q * D + d));   // T <- T * A_d
        park(CUR);
        pair_mul(slot_a(ACC), slot_b(ACC));               // S * T
        park(ACC);
        if (d > 1) fetch(CUR);
      }
      // (D = 2: the single bucket is the result and is already in (A, B); otherwise S is)
      uint32_t st = 0;
      const bool neg = (p.negative_mask >> q) & 1u;
      if (neg)   // (c^-1)^|e| = (c^|e|)^-1: the RESULT is inverted, per instance
        st = pair_invert<K, M>(Aw, Bw, scratch32 + (unsigned long long)(p.nparties * D + 2) * (unsigned long long)(2 * Lp * 32), U32, Nd,
                               two_ab, R2Ar, R2Br, lane, pair_mul);
      pair_mul(PLAIN1r, ZEROr);   // out of the Montgomery domain
      if (neg) {
        if (st) {
          const V zero = V();
          for (int v = 0; v < LV; ++v) { Aw[v * 32 + lane] = zero; Bw[v * 32 + lane] = zero; }
        }
        if (p.status != nullptr && lane < cnt) p.status[(size_t)q * p.count + first + lane] = (uint8_t)st;
      }
      __syncwarp();
      uint32_t* outp = p.pairs_out + (size_t)q * p.count * (size_t)(2 * La);
      for (int r = 0; r < cnt; ++r) {
        uint32_t* row = outp + (first + (unsigned long long)r) * (unsigned long long)(2 * La);
        for (int l = lane; l < Lp; l += 32) {
          if ((l % KP) >= K) continue;
          const int i = ((l / VW) * 32 + r) * VW + (l % VW);
          const int dst = (l / KP) * K + (l % KP);
          row[dst] = Aw32[i];
          row[La + dst] = Bw32[i];
        }
      }
      __syncwarp();
    }
  }
}

}  // namespace dkg
