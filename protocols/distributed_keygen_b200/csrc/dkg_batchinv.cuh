// Batched modular inversion (negative exponents: the reference inverts every ciphertext with
// mod_inv, paillier_shared_key.py:89-91).  Montgomery's trick along chains: one chain per lane,
// G elements per chain, so a batch of B inversions costs B/G binary-GCD inversions plus ~4
// Montgomery products per element.  A chain that contains a non-unit is flagged as a whole; the
// modexp kernel then inverts the elements of that chain one by one to pin the status per element.
//
// One warp per 32 chains.  Shared memory per warp: X (running prefix product / running inverse),
// X2 (the element being folded in) and the four work arrays of the binary GCD.
#pragma once
#include "dkg_modexp.cuh"

namespace dkg {

template <int K, int M>
__global__ void __launch_bounds__(64, 1) batchinv_kernel(const BatchInvParams p) {
  using V = typename VecSel<K>::T;
  constexpr int VW = VecSel<K>::VW;
  constexpr int Lp = K * M;
  constexpr int LV = Lp / VW;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* Ns32 = reinterpret_cast<uint32_t*>(smem_raw);
  constexpr int UNI = ((Lp + K) * 4 + 15) / 16 * 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < Lp + K; i += blockDim.x) Ns32[i] = p.consts[i];
  __syncthreads();

  const int w = blockIdx.x * nwarps + warp;
  if (w >= p.nchain_warps) return;

  // per warp: X | X2 | 4 work arrays, each Lp*32 words
  uint32_t* base32 = reinterpret_cast<uint32_t*>(smem_raw + UNI) + (size_t)warp * 6 * Lp * 32;
  V* Xw = reinterpret_cast<V*>(base32);
  V* X2w = reinterpret_cast<V*>(base32 + Lp * 32);
  uint32_t* work = base32 + 2 * Lp * 32;
  const V* Ns = reinterpret_cast<const V*>(Ns32);
  const V* NIs = reinterpret_cast<const V*>(Ns32 + Lp);
  // lane-replicated R^2, R, R^3 ([v][lane]) follow the uniform constants
  const V* R2rep = reinterpret_cast<const V*>(p.consts + Lp + K + 3 * Lp) + lane;
  const V* R3rep = R2rep + (size_t)2 * LV * 32;
  V* Qg = reinterpret_cast<V*>(p.scratch + (size_t)w * Lp * 32);

  WarpIO<K, M> ioX;
  ioX.xs = (uint32_t)__cvta_generic_to_shared(Xw + lane);
  ioX.ss = ioX.xs;   // (no second operand here: the prefetch's dummy reads go to X)
  ioX.ns = (uint32_t)__cvta_generic_to_shared(Ns);
  ioX.nis = (uint32_t)__cvta_generic_to_shared(NIs);
  ioX.Qg = Qg + lane; ioX.Y = nullptr;
  WarpIO<K, M> ioX2 = ioX;
  ioX2.xs = (uint32_t)__cvta_generic_to_shared(X2w + lane);
  ioX2.ss = ioX2.xs;

  const unsigned long long ngroups = (p.count + 31ull) / 32ull;
  auto group_of = [&](int k) -> unsigned long long { return (unsigned long long)w + (unsigned long long)k * p.nchain_warps; };
  auto S = [&](unsigned long long g) -> V* { return reinterpret_cast<V*>(p.chain_s + g * (size_t)Lp * 32) + lane; };
  auto P = [&](unsigned long long g) -> V* { return reinterpret_cast<V*>(p.chain_p + g * (size_t)Lp * 32) + lane; };

  // ---- phase A: prefix products P_k = prod_{j<=k} c_j (Montgomery form) ------------------------
  int glen = 0;
  uint32_t* X2w32 = reinterpret_cast<uint32_t*>(X2w);
  for (int k = 0; k < p.chain_len; ++k) {
    const unsigned long long g = group_of(k);
    if (g >= ngroups) break;
    glen = k + 1;
    const unsigned long long first = g * 32ull;
    const int cnt = (int)((p.count - first) < 32ull ? (p.count - first) : 32ull);
    for (int r = 0; r < 32; ++r) {
      const uint32_t* row = p.bases + (first + (unsigned long long)r) * (unsigned long long)p.in_limbs;
      for (int l = lane; l < Lp; l += 32) {
        uint32_t v = (r < cnt) ? (l < p.in_limbs ? row[l] : 0u) : (l == 0 ? 1u : 0u);
        X2w32[((l / VW) * 32 + r) * VW + (l % VW)] = v;
      }
    }
    __syncwarp();
    ioX2.Y = R2rep;
    mont_call<K, M, MONT_MUL>(ioX2);                     // c_k * R
    V* s = S(g);
    for (int v = 0; v < LV; ++v) s[(size_t)v * 32] = X2w[v * 32 + lane];
    if (k == 0) {
      for (int v = 0; v < LV; ++v) Xw[v * 32 + lane] = X2w[v * 32 + lane];
    } else {
      ioX.Y = s;
      mont_call<K, M, MONT_MUL>(ioX);                    // P_k = P_{k-1} * c_k
      V* pp = P(g);
      for (int v = 0; v < LV; ++v) pp[(size_t)v * 32] = Xw[v * 32 + lane];
    }
    __syncwarp();
  }
  if (glen == 0) return;

  // ---- phase B: invert the chain product --------------------------------------------------------
  // X = P * R (Montgomery form); (X)^-1 = P^-1 R^-1; times R^3 / R -> P^-1 * R
  uint32_t bad = mod_inverse_lane<K, M>(reinterpret_cast<uint32_t*>(Xw) + lane * VW, Ns32, p.n0inv, work + lane);
  __syncwarp();
  ioX.Y = R3rep;
  mont_call<K, M, MONT_MUL>(ioX);
  p.chain_status[(size_t)w * 32 + lane] = bad;
  if (bad && p.any_bad != nullptr) atomicOr(p.any_bad, 1u);

  // optional: canonical plain inverse of the group held in `io`'s X, written as rows
  auto store_plain = [&](WarpIO<K, M>& io, uint32_t* x32, unsigned long long g) {
    mont_call<K, M, MONT_REDC>(io);
    canonicalize<K, M>(io, 1);
    __syncwarp();
    const unsigned long long first = g * 32ull;
    const int cnt = (int)((p.count - first) < 32ull ? (p.count - first) : 32ull);
    for (int r = 0; r < cnt; ++r) {
      uint32_t* row = p.plain_out + (first + (unsigned long long)r) * (unsigned long long)p.in_limbs;
      for (int l = lane; l < p.in_limbs; l += 32) row[l] = x32[((l / VW) * 32 + r) * VW + (l % VW)];
    }
    __syncwarp();
  };

  // ---- phase C: peel the chain from the back ------------------------------------------------------
  // inv = (c_0 .. c_k)^-1 ; c_k^-1 = inv * P_{k-1} ; inv <- inv * c_k
  for (int k = glen - 1; k >= 1; --k) {
    const unsigned long long g = group_of(k), gp = group_of(k - 1);
    for (int v = 0; v < LV; ++v) X2w[v * 32 + lane] = Xw[v * 32 + lane];
    ioX2.Y = (k - 1 == 0) ? S(gp) : P(gp);
    mont_call<K, M, MONT_MUL>(ioX2);                     // c_k^-1 * R
    V* s = S(g);
    ioX.Y = s;
    mont_call<K, M, MONT_MUL>(ioX);                      // inv for the next step
    for (int v = 0; v < LV; ++v) s[(size_t)v * 32] = X2w[v * 32 + lane];
    if (p.plain_out != nullptr) store_plain(ioX2, X2w32, g);
  }
  {
    V* s = S(group_of(0));
    for (int v = 0; v < LV; ++v) s[(size_t)v * 32] = Xw[v * 32 + lane];
    if (p.plain_out != nullptr) store_plain(ioX, reinterpret_cast<uint32_t*>(Xw), group_of(0));
  }
}

}  // namespace dkg
