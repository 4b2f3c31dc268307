// Launch parameters of the fixed-modulus modexp kernel (shared by the kernels and the host dispatch).
#pragma once
#include <stdint.h>

// Threads per CTA the modexp kernels are compiled for (__launch_bounds__): 12 warps = 3 per SM
// sub-partition, which caps the kernel at 168 registers per thread.
#ifndef DKG_MAX_THREADS
#define DKG_MAX_THREADS 384
#endif

namespace dkg {

// Words of the tabulated pair schedule (all five modes of the Montgomery product, each list with a
// terminator; dkg_mont.cuh "tabulated schedule") for M blocks: the host sizes the kernel's shared
// memory with it, the kernel static_asserts that it equals what ColPlan generates.
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr int sched_total_words_closed(int M) {
  int half = 0;  // sum over columns of ceil(span / 2): the operand pairs of a squaring
  for (int c = 0; c < 2 * M; ++c) {
    const int lo = c >= M ? c - M + 1 : 0, hi = c < M ? c : M - 1, span = hi - lo + 1;
    half += span / 2 + (span & 1);
  }
  const int mm = M * M;
  // MUL 2M^2, REDC M^2, SQR half + M^2, MUL2S 2M^2, MULADD 3M^2, plus five terminators
  return 2 * mm + mm + (half + mm) + 2 * mm + 3 * mm + 5;
}

struct ModexpParams {
  const uint32_t* bases;   // [count][in_limbs]
  uint32_t* out;           // [count][in_limbs]
  uint8_t* status;         // [count] or null
  unsigned long long count;
  int in_limbs;
  // device constants: N[Lp] | NINV[K] | R2[Lp] | ONER[Lp] | R3[Lp]   (Lp = K*M)
  const uint32_t* consts;
  // operation list shared by the whole batch (built by the host, dkg_engine.cu):
  // op = (nsq << 8) | idx: square nsq times, then multiply by table[idx]
  // (idx 0xff: no multiplication, 0xfe: multiply by the Montgomery one); op 0 has nsq = 0 and
  // starts from table[idx].  table_odd = 0: table[k] = c^(k+1) (fixed windows, the default);
  // table_odd = 1: table[k] = c^(2k+1) (sliding windows)
  const uint32_t* ops;
  int nops;
  int tab_entries;
  int table_odd;
  int negative;            // invert the base first
  uint32_t n0inv;          // -N^-1 mod 2^32
  uint32_t* scratch;       // per-warp table scratch
  unsigned long long scratch_per_warp;  // in uint32
  unsigned long long scratch_q_offset;  // offset (uint32) of the quotient-block area inside a warp's scratch
  unsigned int* counter;   // work-group ticket
  // optional: run only if *run_if != 0 (the device-side decision "the batched inversion of the
  // pair path met a non-unit: redo the call here, with exact per-element status"); null = always
  const unsigned int* run_if;
  // optional per-element plain multiplier applied at the end (encryption: 1 + m N), or null
  const uint32_t* final_mul;  // [count][in_limbs]
  // batched inversion results (negative exponents): c^-1 * R per group in lane layout, and one
  // flag per chain lane (non-zero: the chain was not invertible as a whole); null = invert in-kernel
  const uint32_t* inv_mont;       // [groups][Lp * 32]
  const uint32_t* chain_status;   // [nchain_warps * 32]
  int nchain_warps;
};

// Batched modular inversion by Montgomery's trick (dkg_batchinv.cuh): chain warp w owns groups
// w, w + nchain_warps, w + 2 nchain_warps, ... (a group = 32 consecutive rows; lane l of the warp
// chains element l of each of its groups).
struct BatchInvParams {
  const uint32_t* bases;  // [count][in_limbs]
  unsigned long long count;
  int in_limbs;
  const uint32_t* consts;
  uint32_t n0inv;
  uint32_t* chain_s;       // [groups][Lp*32]  c*R, later overwritten with c^-1 * R
  uint32_t* chain_p;       // [groups][Lp*32]  prefix products
  uint32_t* chain_status;  // [nchain_warps*32]
  uint32_t* scratch;       // per-warp quotient-block scratch, [warps][Lp*32]
  int nchain_warps;
  int chain_len;           // groups per chain warp (upper bound)
  uint32_t* plain_out;     // optional [count][in_limbs]: canonical c^-1 mod N (rows), else null
  unsigned int* any_bad;   // set to 1 if any chain was not invertible
};

}  // namespace dkg
