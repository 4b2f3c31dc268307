// Launch parameters of the grouped (per-group modulus) modexp kernels.
#pragma once
#include <stdint.h>

namespace dkg {

constexpr int kGroupedMaxLimbs = 144;

struct GroupedParams {
  const uint32_t* moduli;   // [groups][limbs]
  const uint32_t* exps;     // [groups][exp_limbs]
  const uint32_t* bases;    // [groups*per_group][limbs]
  uint32_t* out;            // same shape as bases
  unsigned long long groups;
  int per_group, limbs, exp_limbs;
  int K, Lp;                // memory layout: slot of one block (even), limbs of a number = K * blocks
  int Ka;                   // arithmetic block size (Ka = K, or K - 1 with a zero pad limb per slot): R = 2^(32 Ka blocks)
  int wbits, ndigits;
  uint32_t* gconsts;        // [groups][Lp + K + Lp + Lp]: N | NINV | R2 | ONER  (slot layout)
  uint8_t* digits;          // [groups][ndigits], most significant first
  uint32_t* scratch;
  unsigned long long scratch_per_warp, scratch_q_offset;
  unsigned int* counter;
};

}  // namespace dkg
