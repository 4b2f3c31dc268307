// Launch parameters of the cooperative (warp-per-operand) kernels, dkg_coop.cuh (shared by the
// kernels and the host dispatch).
#pragma once
#include <stdint.h>

namespace dkg {

constexpr unsigned kCoopFull = 0xffffffffu;
constexpr int kCoopMaxBlocks = 16;

// Host-built lane plan of one product phase (dkg_engine.cu: make_coop_plan).
struct CoopPlanTable {
  uint8_t d[32], i0[32], i1[32];
  int8_t partner[3][32];  // lanes whose partial sum this (primary) lane adds; -1 = none
  int32_t rounds;         // max number of partners of any lane
  int32_t ndiag;          // lanes 0..ndiag-1 are primaries
};

constexpr int kCoopMaxParties = 8;

struct CoopNsqParams {
  const uint32_t* pairs_in;   // [count][2][Lc] plain pairs (c mod N, (c div N) * R mod N), R = 2^(32 Lc)
  uint32_t* pairs_out;        // [nparties][count][2][Lc] plain pairs of the results
  uint8_t* status;            // [nparties][count] or null: 1 = base not invertible (negative exponent only)
  unsigned long long count;   // ciphertexts; instance j = (party j / count, ciphertext j % count)
  int nb;                     // blocks per component; Lc = nb * K
  // N | NI (-N^-1 mod R) | DNEG (-R mod N) | R2A | R2B | ONEA | ONEB | TWOA | TWOB | PLAIN1 | ZERO
  const uint32_t* consts;
  // per party (all parties of one key share N; a plain context is the one-party case): the
  // operation list of its exponent (see ModexpParams), sign, table size
  int nparties;
  const uint32_t* ops[kCoopMaxParties];
  int nops[kCoopMaxParties], tab_entries[kCoopMaxParties], table_odd[kCoopMaxParties], negative[kCoopMaxParties];
  int ct_table;               // masked scan of the whole table per multiplication (fixed windows only)
  uint32_t* scratch;          // per-warp window table, (max tab_entries + 1) * 2 * Lc words
  unsigned long long scratch_per_warp;
  unsigned int* counter;
  CoopPlanTable full, low;
};
constexpr int kCoopNsqConsts = 11;
constexpr int kCoopNsqWarpBufs = 14;   // numbers of nb*K limbs of shared memory per warp

// Share combination (PaillierSharedKey.decrypt, paillier_shared_key.py:108-125), one ciphertext per warp.
struct CoopCombineParams {
  const uint32_t* partials;   // [shares][count][l2]
  uint32_t* out;              // [count][ln]
  uint8_t* status;            // [count] or null: 2 = (x - 1) not divisible by N
  unsigned long long count;
  int shares, l2, ln;
  int nb;                     // Lc = nb * K limbs, R = 2^(32 Lc) >= 4 N^2
  // N2 | NI2 (-N2^-1 mod R) | RPOW (R^shares mod N^2) | N | NIN (-N^-1 mod R) | THR (theta^-1 * R mod N); Lc limbs each
  const uint32_t* consts;
  unsigned int* counter;
  CoopPlanTable full, low;
};
constexpr int kCoopCombineConsts = 6;
constexpr int kCoopCombineWarpBufs = 5;   // X Y T Q U

struct CoopGroupedParams {
  const uint32_t* moduli;   // [groups][limbs]
  const uint32_t* exps;     // [groups][exp_limbs]
  const uint32_t* bases;    // [groups * per_group][limbs], each below its modulus
  uint32_t* out;
  unsigned long long groups;
  int per_group, limbs, exp_limbs;
  int nb;
  int wbits, ndigits;
  uint32_t* scratch;        // per-warp window table, (2^wbits - 1) * Lc words
  unsigned long long scratch_per_warp;
  unsigned int* counter;
  CoopPlanTable full, low;
};

}  // namespace dkg
