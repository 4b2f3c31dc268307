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
  //   | PLAIN1 (= 1) | ZERO | the six pairs / constants again, lane-replicated | TWOA | TWOB
  //   (TWOA, TWOB: the pair of 2 with 2N added to its a component, for the Newton step of the
  //   in-kernel inversion; slot layout like N)
  const uint32_t* consts;
  const uint32_t* ops;        // operation list, see ModexpParams
  int nops, tab_entries, table_odd;
  int ct_table;               // masked scan of the whole table per multiplication (fixed windows only)
  int negative;               // invert the base first, in the kernel (pair_invert); status per element
  uint8_t* status;            // [count] or null: written when `negative` (1 = not invertible)
  unsigned long long inv_slot;   // first of 3 spare pair slots (+ GCD work space) in a warp's scratch, in pair slots
  uint32_t* scratch;
  unsigned long long scratch_per_warp;   // in uint32
  unsigned long long scratch_q_offset;
  unsigned int* counter;
};

// Several exponentiations of the SAME bases (all parties' partial decryptions of one ciphertext
// batch when their shares sit on one device): one shared squaring chain, per-party digit buckets
// (modexp_nsq_multi_kernel in dkg_nsq.cuh).
constexpr int kNsqMultiMaxParties = 8;
struct NsqMultiParams {
  const uint32_t* pairs_in;   // [count][2][Lp]
  uint32_t* pairs_out;        // [nparties][count][2][Lp]
  unsigned long long count;
  const uint32_t* consts;     // as NsqParams::consts
  const uint8_t* digits;      // [nparties][nwin]: |exponent| of party p in base 2^wbits, least significant digit first
  int nparties, nwin, wbits;
  unsigned int negative_mask; // bit p: party p's exponent is negative: its RESULT is inverted in the kernel
  uint8_t* status;            // [nparties][count] or null: written for the parties of negative_mask
  uint32_t* scratch;
  unsigned long long scratch_per_warp;   // in uint32: (nparties * 2^wbits + 2 + 4) pairs, then the quotient blocks
  unsigned long long scratch_q_offset;
  unsigned int* counter;
};

}  // namespace dkg
