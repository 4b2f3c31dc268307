// Fixed-modulus, fixed-exponent batched modular exponentiation: one big integer per thread,
// one persistent CTA per SM, every warp owns 32 independent instances at a time.
//
// Data layout (see DESIGN.md "Data layout"):
//  * shared memory, per warp: X (running value) and Q (Montgomery quotient blocks), each stored
//    limb-vector-major / lane-minor: vector v (VW = 4 or 2 limbs) of lane l at [(v*32 + l)] so that
//    one LDS.128/LDS.64 of a warp touches 32 consecutive vectors (conflict-free);
//  * shared memory, per CTA: the modulus N, its block inverse -N^-1 mod 2^(32K) (uniform:
//    broadcast loads);
//  * global scratch, per warp: the window table c^1..c^(2^w-1) in Montgomery form, same
//    vector-major / lane-minor layout, so a warp's table reads are 512-byte coalesced;
//  * bases / results in HBM are row-major [count][limbs] as the C ABI hands them over; the warp
//    transposes through shared memory on the way in and out.
#pragma once
#include <cuda_runtime.h>
#include "dkg_mont.cuh"
#include "dkg_modexp_params.h"

namespace dkg {

template <int K> struct VecSel { using T = uint2; static constexpr int VW = 2; };
#define DKG_VEC4(K_) template <> struct VecSel<K_> { using T = uint4; static constexpr int VW = 4; };
DKG_VEC4(4) DKG_VEC4(8) DKG_VEC4(12) DKG_VEC4(16) DKG_VEC4(20) DKG_VEC4(24) DKG_VEC4(28) DKG_VEC4(32)
#undef DKG_VEC4

__device__ __forceinline__ void unpack(const uint4& v, uint32_t* r) { r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; }
__device__ __forceinline__ void unpack(const uint2& v, uint32_t* r) { r[0] = v.x; r[1] = v.y; }
__device__ __forceinline__ void pack(uint4& v, const uint32_t* r) { v = make_uint4(r[0], r[1], r[2], r[3]); }
__device__ __forceinline__ void pack(uint2& v, const uint32_t* r) { v = make_uint2(r[0], r[1]); }

// IO policy of dkg::mont_mul for one thread of a warp.  X/Q point at this lane's vector 0.
// YMODE 0: Y aliases X (squaring); 1: Y is in global memory at Y[v * ystride].
template <int K, int M, int YMODE>
struct WarpIO {
  using V = typename VecSel<K>::T;
  static constexpr int VW = VecSel<K>::VW;
  static constexpr int KV = K / VW;
  V* X;
  V* Q;
  const V* Ns;   // shared, uniform
  const V* NIs;  // shared, uniform
  const V* Y;    // global
  int ystride;

  __device__ __forceinline__ void load_x(int i, uint32_t (&r)[K]) const {
#pragma unroll
    for (int q = 0; q < KV; q++) unpack(X[(i * KV + q) * 32], &r[q * VW]);
  }
  __device__ __forceinline__ void load_y(int j, uint32_t (&r)[K]) const {
    if (YMODE == 0) {
      load_x(j, r);
    } else {
#pragma unroll
      for (int q = 0; q < KV; q++) unpack(Y[(size_t)(j * KV + q) * ystride], &r[q * VW]);
    }
  }
  __device__ __forceinline__ void load_q(int i, uint32_t (&r)[K]) const {
#pragma unroll
    for (int q = 0; q < KV; q++) unpack(Q[(i * KV + q) * 32], &r[q * VW]);
  }
  __device__ __forceinline__ void load_n(int j, uint32_t (&r)[K]) const {
#pragma unroll
    for (int q = 0; q < KV; q++) unpack(Ns[j * KV + q], &r[q * VW]);
  }
  __device__ __forceinline__ void load_ninv(uint32_t (&r)[K]) const {
#pragma unroll
    for (int q = 0; q < KV; q++) unpack(NIs[q], &r[q * VW]);
  }
  __device__ __forceinline__ void store_q(int i, const uint32_t (&r)[K]) const {
#pragma unroll
    for (int q = 0; q < KV; q++) pack(Q[(i * KV + q) * 32], &r[q * VW]);
  }
  __device__ __forceinline__ void store_x(int i, const uint32_t (&r)[K]) const {
#pragma unroll
    for (int q = 0; q < KV; q++) pack(X[(i * KV + q) * 32], &r[q * VW]);
  }
};


// limb l of a lane-private big integer stored vector-major in shared memory (base = lane's vector 0)
template <int VW>
__device__ __forceinline__ int sidx(int l) { return (l / VW) * 32 * VW + (l % VW); }

// Modular inverse of the lane's value in X (mod N), binary extended GCD with multi-bit shifts.
// u (X region) and v (Q region) live in shared memory, the cofactors x1, x2 in the warp's global
// scratch (stride 32).  Invariants: x1*a = u, x2*a = v (mod N).  Returns 0 if invertible (X <-
// a^-1 mod N), 1 otherwise (X unspecified).
template <int K, int M>
__device__ uint32_t mod_inverse_lane(uint32_t* U, uint32_t* Vv, const uint32_t* Ns, uint32_t n0inv,
                                     uint32_t* g1, uint32_t* g2) {
  constexpr int VW = VecSel<K>::VW;
  constexpr int Lp = K * M;
  uint32_t nz = 0;
  for (int l = 0; l < Lp; ++l) {
    Vv[sidx<VW>(l)] = Ns[l];
    g1[l * 32] = (l == 0) ? 1u : 0u;
    g2[l * 32] = 0u;
    nz |= U[sidx<VW>(l)];
  }
  uint32_t* pu = U; uint32_t* pv = Vv; uint32_t* px1 = g1; uint32_t* px2 = g2;
  int len = Lp;
  int guard = 64 * Lp + 64;
  while (nz != 0 && guard-- > 0) {
    // 1. make u odd: shift out up to 32 zero bits at a time, dividing x1 by the same power of 2
    uint32_t u0 = pu[sidx<VW>(0)];
    while ((u0 & 1u) == 0) {
      const int tz = u0 ? __ffs(u0) - 1 : 32;
      uint32_t lo = u0;
      for (int l = 0; l < len; ++l) {
        const uint32_t hi = (l + 1 < len) ? pu[sidx<VW>(l + 1)] : 0u;
        pu[sidx<VW>(l)] = (uint32_t)((((uint64_t)hi << 32) | lo) >> tz);
        lo = hi;
      }
      const uint32_t mask = tz == 32 ? 0xffffffffu : ((1u << tz) - 1u);
      const uint32_t m = (px1[0] * n0inv) & mask;
      uint64_t carry = 0;
      uint32_t prev = 0;
      for (int l = 0; l < Lp; ++l) {
        const uint64_t t = (uint64_t)px1[l * 32] + (uint64_t)m * Ns[l] + carry;
        const uint32_t cur = (uint32_t)t;
        carry = t >> 32;
        if (l > 0) px1[(l - 1) * 32] = (uint32_t)((((uint64_t)cur << 32) | prev) >> tz);
        prev = cur;
      }
      px1[(Lp - 1) * 32] = (uint32_t)(((carry << 32) | prev) >> tz);
      u0 = pu[sidx<VW>(0)];
    }
    // 2. order: u >= v
    bool lt = false;
    for (int l = len - 1; l >= 0; --l) {
      const uint32_t a = pu[sidx<VW>(l)], b = pv[sidx<VW>(l)];
      if (a != b) { lt = a < b; break; }
    }
    if (lt) {
      uint32_t* t = pu; pu = pv; pv = t;
      t = px1; px1 = px2; px2 = t;
    }
    // 3. u -= v ; x1 = x1 - x2 (mod N)
    uint32_t borrow = 0;
    nz = 0;
    for (int l = 0; l < len; ++l) {
      const uint64_t d = (uint64_t)pu[sidx<VW>(l)] - pv[sidx<VW>(l)] - borrow;
      pu[sidx<VW>(l)] = (uint32_t)d;
      nz |= (uint32_t)d;
      borrow = (uint32_t)(d >> 63);
    }
    bool xlt = false;
    for (int l = Lp - 1; l >= 0; --l) {
      const uint32_t a = px1[l * 32], b = px2[l * 32];
      if (a != b) { xlt = a < b; break; }
    }
    const uint32_t addmask = xlt ? 0xffffffffu : 0u;
    int64_t c = 0;
    for (int l = 0; l < Lp; ++l) {
      const int64_t t = (int64_t)px1[l * 32] - (int64_t)px2[l * 32] + (int64_t)(Ns[l] & addmask) + c;
      px1[l * 32] = (uint32_t)t;
      c = t >> 32;
    }
    while (len > 1 && pu[sidx<VW>(len - 1)] == 0 && pv[sidx<VW>(len - 1)] == 0) --len;
  }
  // gcd is in v; invertible iff v == 1
  uint32_t bad = pv[sidx<VW>(0)] ^ 1u;
  for (int l = 1; l < Lp; ++l) bad |= pv[sidx<VW>(l)];
  // the inverse is x2; both u and v are dead now
  for (int l = 0; l < Lp; ++l) U[sidx<VW>(l)] = px2[l * 32];
  return bad ? 1u : 0u;
}

// Out-of-line instances of the Montgomery product: the kernel body calls these (three function
// bodies per shape: square, multiply-by-global-operand, reduce) instead of inlining seven copies,
// which keeps the hot loop inside the instruction cache.
template <int K, int M, int YMODE, int MODE>
__device__ __noinline__ void mont_call(const WarpIO<K, M, YMODE> io) {
  mont_mul<K, M, MODE>(io);
}

template <int K, int M>
__global__ void __launch_bounds__(256, 1) modexp_fixed_kernel(const ModexpParams p) {
  using V = typename VecSel<K>::T;
  constexpr int VW = VecSel<K>::VW;
  constexpr int Lp = K * M;
  constexpr int LV = Lp / VW;  // vectors per big integer

  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* Ns32 = reinterpret_cast<uint32_t*>(smem_raw);  // N[Lp] then NINV[K]
  constexpr int UNI = ((Lp + K) * 4 + 15) / 16 * 16;       // bytes of the uniform area
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

  for (int i = threadIdx.x; i < Lp + K; i += blockDim.x) Ns32[i] = p.consts[i];
  __syncthreads();

  V* Xw = reinterpret_cast<V*>(smem_raw + UNI) + (size_t)warp * 2 * LV * 32;
  V* Qw = Xw + LV * 32;
  uint32_t* Xw32 = reinterpret_cast<uint32_t*>(Xw);
  const V* Ns = reinterpret_cast<const V*>(Ns32);
  const V* NIs = reinterpret_cast<const V*>(Ns32 + Lp);
  const V* R2g = reinterpret_cast<const V*>(p.consts + Lp + K);
  const V* ONEg = reinterpret_cast<const V*>(p.consts + Lp + K + Lp);

  const unsigned gwarp = blockIdx.x * nwarps + warp;
  uint32_t* scratch32 = p.scratch + (size_t)gwarp * p.scratch_per_warp;
  V* tab = reinterpret_cast<V*>(scratch32);  // entry d (1-based) at tab[((d-1)*LV + v)*32 + lane]

  WarpIO<K, M, 0> io_sqr{Xw + lane, Qw + lane, Ns, NIs, nullptr, 0};
  WarpIO<K, M, 1> io_mul{Xw + lane, Qw + lane, Ns, NIs, nullptr, 0};

  const unsigned long long ngroups = (p.count + 31ull) / 32ull;
  for (;;) {
    unsigned int g = 0;
    if (lane == 0) g = atomicAdd(p.counter, 1u);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= ngroups) break;
    const unsigned long long first = (unsigned long long)g * 32ull;
    const int cnt = (int)((p.count - first) < 32ull ? (p.count - first) : 32ull);

    // ---- load 32 rows, transposed into X (zero padded; idle lanes get the value 1) ----------
    for (int r = 0; r < 32; ++r) {
      const uint32_t* row = p.bases + (first + (unsigned long long)r) * (unsigned long long)p.in_limbs;
      for (int l = lane; l < Lp; l += 32) {
        uint32_t v = (r < cnt) ? (l < p.in_limbs ? row[l] : 0u) : (l == 0 ? 1u : 0u);
        Xw32[((l / VW) * 32 + r) * VW + (l % VW)] = v;
      }
    }
    __syncwarp();

    uint32_t st = 0;
    if (p.negative) {
      st = mod_inverse_lane<K, M>(Xw32 + lane * VW, reinterpret_cast<uint32_t*>(Qw) + lane * VW, Ns32,
                                  p.n0inv, scratch32 + lane, scratch32 + Lp * 32 + lane);
      __syncwarp();
    }

    // ---- to Montgomery form: X <- X * R^2 / R ------------------------------------------------
    io_mul.Y = R2g; io_mul.ystride = 1;
    mont_call<K, M, 1, MONT_MUL>(io_mul);

    if (p.ndigits == 0) {
      for (int v = 0; v < LV; ++v) Xw[v * 32 + lane] = ONEg[v];
    } else {
      // ---- window table: tab[d] = c^d (Montgomery form), d = 1 .. 2^w - 1 --------------------
      const int tsize = (1 << p.wbits) - 1;
      for (int v = 0; v < LV; ++v) tab[(size_t)v * 32 + lane] = Xw[v * 32 + lane];
      for (int d = 2; d <= tsize; ++d) {
        io_mul.Y = tab + lane; io_mul.ystride = 32;
        mont_call<K, M, 1, MONT_MUL>(io_mul);
        V* dst = tab + (size_t)(d - 1) * LV * 32 + lane;
        for (int v = 0; v < LV; ++v) dst[(size_t)v * 32] = Xw[v * 32 + lane];
      }
      // ---- left-to-right fixed windows; every window multiplies (digit 0 by R mod N) ---------
      {
        const int d0 = p.digits[0];
        const V* src = tab + (size_t)(d0 - 1) * LV * 32 + lane;
        for (int v = 0; v < LV; ++v) Xw[v * 32 + lane] = src[(size_t)v * 32];
      }
      for (int t = 1; t < p.ndigits; ++t) {
        for (int s = 0; s < p.wbits; ++s) mont_call<K, M, 0, MONT_MUL>(io_sqr);
        const int d = p.digits[t];
        if (d == 0) { io_mul.Y = ONEg; io_mul.ystride = 1; }
        else { io_mul.Y = tab + (size_t)(d - 1) * LV * 32 + lane; io_mul.ystride = 32; }
        mont_call<K, M, 1, MONT_MUL>(io_mul);
      }
    }

    // ---- optional final plain multiplier (encryption epilogue), else leave Montgomery form ---
    if (p.final_mul != nullptr) {
      // stage the multipliers through the table area (entry 1) in lane layout
      __syncwarp();
      uint32_t* stage = scratch32;
      for (int r = 0; r < 32; ++r) {
        const uint32_t* row = p.final_mul + (first + (unsigned long long)r) * (unsigned long long)p.in_limbs;
        for (int l = lane; l < Lp; l += 32) {
          uint32_t v = (r < cnt && l < p.in_limbs) ? row[l] : 0u;
          stage[((l / VW) * 32 + r) * VW + (l % VW)] = v;
        }
      }
      __syncwarp();
      io_mul.Y = tab + lane; io_mul.ystride = 32;
      mont_call<K, M, 1, MONT_MUL>(io_mul);   // (x R) * y / R = x y
    } else {
      mont_call<K, M, 0, MONT_REDC>(io_sqr);
    }
    canonicalize<K, M>(io_sqr, p.final_mul != nullptr ? 2 : 1);
    __syncwarp();

    // ---- store rows ---------------------------------------------------------------------------
    for (int r = 0; r < cnt; ++r) {
      const uint32_t st_r = __shfl_sync(0xffffffffu, st, r);
      uint32_t* row = p.out + (first + (unsigned long long)r) * (unsigned long long)p.in_limbs;
      for (int l = lane; l < p.in_limbs; l += 32) {
        const uint32_t v = Xw32[((l / VW) * 32 + r) * VW + (l % VW)];
        row[l] = st_r ? 0u : v;
      }
    }
    if (p.status != nullptr && lane < cnt) p.status[first + lane] = (uint8_t)st;
    __syncwarp();
  }
}

}  // namespace dkg
