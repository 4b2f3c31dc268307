// Launch parameters of the pair-arithmetic (modulus N^2) exponentiation kernel.
#pragma once
#include <stdint.h>
#include "dkg_modexp_params.h"

namespace dkg {

struct NsqParams {
  const uint32_t* pairs_in;   // [count][2][Lp]  plain pairs (c mod N, (c div N) * R mod N)
  uint32_t* pairs_out;        // [count][2][Lp]  plain pairs of the result
  unsigned long long count;
  // device constants, Lp = K*M limbs each unless noted:
  //   N | NINV[K] | DNEG (= -R mod N) | R2A | R2B (pair of R^2 mod N^2) | ONEA | ONEB (pair of R mod N^2)
  //   | PLAIN1 (= 1) | ZERO
  const uint32_t* consts;
  const uint32_t* ops;        // operation list, see ModexpParams
  int nops, tab_entries, table_odd;
  int ct_table;               // masked scan of the whole table per multiplication (fixed windows only)
  uint32_t* scratch;
  unsigned long long scratch_per_warp;   // in uint32
  unsigned long long scratch_q_offset;
  unsigned int* counter;
};

}  // namespace dkg
