// Carry-chain primitives for 32-bit-limb big-integer arithmetic on sm_100a.
//
// On the device every helper is one or two PTX instructions on the CC.CF carry flag; ptxas fuses
// each mad.lo.cc/madc.hi.cc pair into ONE `IMAD.WIDE.U32[.X] Rd, P, Ra, Rb, Rc, P` (64-bit
// multiply-accumulate with predicate carry-in/out) and renames CC.CF onto P0..P6, so several
// chains are in flight at once (checked with cuobjdump; see DESIGN.md "SASS evidence").
// `asm volatile` keeps the statements of one chain in program order.
//
// The host versions (plain C on a thread-local carry bit) exist only so that the *same* templates
// in dkg_mont.cuh can be unit-tested on a machine without a GPU (tests/host/); they are never
// compiled into the product library's compute path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DKG_HD __host__ __device__ __forceinline__
#else
#define DKG_HD inline
#endif

namespace dkg {

#if defined(__CUDA_ARCH__)

// {hi,lo} += a*b                      (starts a chain: carry-out only)
DKG_HD void mad_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// {hi,lo} += a*b + CF                 (continues a chain)
DKG_HD void madc_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// Same on a 64-bit typed accumulator (the aligned register pair IMAD.WIDE works on).  Typing the
// pair as ONE value matters to ptxas: with two 32-bit registers per pair it let the halves drift
// apart across loop back edges and re-paired ~70 registers with moves around every block product.
DKG_HD void mad_cc64(uint64_t& acc, uint32_t a, uint32_t b) {
  asm volatile("{\n\t.reg .u32 l, h;\n\tmov.b64 {l, h}, %0;\n\tmad.lo.cc.u32 l, %1, %2, l;\n\t"
               "madc.hi.cc.u32 h, %1, %2, h;\n\tmov.b64 %0, {l, h};\n\t}" : "+l"(acc) : "r"(a), "r"(b));
}
DKG_HD void madc_cc64(uint64_t& acc, uint32_t a, uint32_t b) {
  asm volatile("{\n\t.reg .u32 l, h;\n\tmov.b64 {l, h}, %0;\n\tmadc.lo.cc.u32 l, %1, %2, l;\n\t"
               "madc.hi.cc.u32 h, %1, %2, h;\n\tmov.b64 %0, {l, h};\n\t}" : "+l"(acc) : "r"(a), "r"(b));
}
// last element of a chain together with the count of its carry out (one statement: ptxas keeps the
// carry add next to the chain instead of parking the carry predicates of a whole block product in a
// mask register, which it did for odd block sizes: 66 LOP3 per block product)
DKG_HD void madc_cc64_count(uint64_t& acc, uint32_t a, uint32_t b, uint32_t& cnt) {
  asm volatile("{\n\t.reg .u32 l, h;\n\tmov.b64 {l, h}, %0;\n\tmadc.lo.cc.u32 l, %2, %3, l;\n\t"
               "madc.hi.cc.u32 h, %2, %3, h;\n\taddc.u32 %1, %1, 0;\n\tmov.b64 %0, {l, h};\n\t}"
               : "+l"(acc), "+r"(cnt) : "r"(a), "r"(b));
}
// Explicit pair <-> halves moves.  Written as (volatile) mov.b64 rather than shifts and ors: ptxas
// then keeps the halves where the pair lives; with the C++ form it re-paired ~30 more registers
// per block product.
DKG_HD uint64_t pack64(uint32_t lo, uint32_t hi) {
  uint64_t v;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
  return v;
}
DKG_HD void unpack64(uint64_t v, uint32_t& lo, uint32_t& hi) {
  asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
// lo += lo32(a*b)                     (no flags)
DKG_HD void mad_lo(uint32_t& lo, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a), "r"(b));
}
// lo += lo32(a*b) + CF                (ends a chain, carry-out dropped)
DKG_HD void madc_lo(uint32_t& lo, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a), "r"(b));
}
DKG_HD void add_cc(uint32_t& x, uint32_t y) { asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x) : "r"(y)); }
DKG_HD void addc_cc(uint32_t& x, uint32_t y) { asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x) : "r"(y)); }
DKG_HD void addc(uint32_t& x, uint32_t y) { asm volatile("addc.u32 %0, %0, %1;" : "+r"(x) : "r"(y)); }

#else  // host emulation (unit tests only)

namespace detail {
inline uint32_t& cf() {
  static thread_local uint32_t flag = 0;
  return flag;
}
inline void add3(uint32_t& x, uint32_t y, uint32_t cin, bool set) {
  uint64_t s = (uint64_t)x + y + cin;
  x = (uint32_t)s;
  if (set) cf() = (uint32_t)(s >> 32);
}
}  // namespace detail

DKG_HD void mad_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  detail::add3(lo, (uint32_t)p, 0, true);
  detail::add3(hi, (uint32_t)(p >> 32), detail::cf(), true);
}
DKG_HD void madc_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  detail::add3(lo, (uint32_t)p, detail::cf(), true);
  detail::add3(hi, (uint32_t)(p >> 32), detail::cf(), true);
}
DKG_HD void mad_cc64(uint64_t& acc, uint32_t a, uint32_t b) {
  uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);
  mad_cc(lo, hi, a, b);
  acc = ((uint64_t)hi << 32) | lo;
}
DKG_HD void madc_cc64(uint64_t& acc, uint32_t a, uint32_t b) {
  uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);
  madc_cc(lo, hi, a, b);
  acc = ((uint64_t)hi << 32) | lo;
}
DKG_HD void madc_cc64_count(uint64_t& acc, uint32_t a, uint32_t b, uint32_t& cnt) {
  madc_cc64(acc, a, b);
  detail::add3(cnt, 0, detail::cf(), false);
}
DKG_HD uint64_t pack64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
DKG_HD void unpack64(uint64_t v, uint32_t& lo, uint32_t& hi) { lo = (uint32_t)v; hi = (uint32_t)(v >> 32); }
DKG_HD void mad_lo(uint32_t& lo, uint32_t a, uint32_t b) { lo += a * b; }
DKG_HD void madc_lo(uint32_t& lo, uint32_t a, uint32_t b) { lo += a * b + detail::cf(); }
DKG_HD void add_cc(uint32_t& x, uint32_t y) { detail::add3(x, y, 0, true); }
DKG_HD void addc_cc(uint32_t& x, uint32_t y) { detail::add3(x, y, detail::cf(), true); }
DKG_HD void addc(uint32_t& x, uint32_t y) { detail::add3(x, y, detail::cf(), false); }

#endif

}  // namespace dkg
