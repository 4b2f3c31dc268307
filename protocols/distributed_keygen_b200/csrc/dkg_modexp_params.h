// Launch parameters of the fixed-modulus modexp kernel (shared by the kernels and the host dispatch).
#pragma once
#include <stdint.h>

namespace dkg {

struct ModexpParams {
  const uint32_t* bases;   // [count][in_limbs]
  uint32_t* out;           // [count][in_limbs]
  uint8_t* status;         // [count] or null
  unsigned long long count;
  int in_limbs;
  // device constants: N[Lp] | NINV[K] | R2[Lp] | ONER[Lp]   (Lp = K*M)
  const uint32_t* consts;
  const uint8_t* digits;   // window digits, most significant first
  int ndigits;
  int wbits;
  int negative;            // invert the base first
  uint32_t n0inv;          // -N^-1 mod 2^32
  uint32_t* scratch;       // per-warp table scratch
  unsigned long long scratch_per_warp;  // in uint32
  unsigned int* counter;   // work-group ticket
  // optional per-element plain multiplier applied at the end (encryption: 1 + m N), or null
  const uint32_t* final_mul;  // [count][in_limbs]
};

}  // namespace dkg
