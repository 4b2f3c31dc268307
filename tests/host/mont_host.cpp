// Host build of the SAME block-Montgomery templates the CUDA kernels instantiate
// (protocols/distributed_keygen_b200/csrc/dkg_mont.cuh), with the PTX carry primitives replaced by their C
// emulation.  Test scaffolding only: lets `pytest -m "not gpu"` check the index logic, bounds and
// carry handling of the kernels' arithmetic against Python big integers without a GPU.
#include <cstring>
#include "dkg_mont.cuh"

namespace {
template <int K, int M>
struct HostIO {
  uint32_t* X; const uint32_t* Y; uint32_t* Q; const uint32_t* N; const uint32_t* NI;
  const uint32_t* S = nullptr; const uint32_t* Y2 = nullptr;
  bool never() const { return false; }
  bool any_lane(uint32_t v) const { return v != 0; }
  static constexpr bool SCHED = false;   // the host harness recomputes the schedule from ColPlan
  int sched_begin(int) const { return 0; }
  int sched_next(int p) const { return p; }
  uint32_t sched_word(int, int) const { return 0; }
  // blocks are dense in memory (K limbs); the kernels' operand arrays are slots of KP = K + (K & 1)
  // limbs whose pad limb is zero
  static constexpr int KP = dkg::kpad<K>;
  static void get(uint32_t (&r)[KP], const uint32_t* src) { std::memset(r, 0, sizeof(r)); std::memcpy(r, src, K * 4); }
  void load_s(int i, uint32_t (&r)[KP]) const { get(r, S + i * K); }
  void load_x(int i, uint32_t (&r)[KP]) const { get(r, X + i * K); }
  uint32_t x_top_limb(int b) const { return X[b * K + K - 1]; }
  void load_xs2(bool from_s, int i, uint32_t (&r)[KP]) const {
    const uint32_t* v = from_s ? S : X;
    std::memset(r, 0, sizeof(r));
    for (int p = 0; p < K; ++p) {
      const int l = i * K + p;
      r[p] = (v[l] << 1) | (l > 0 ? v[l - 1] >> 31 : 0u);
    }
  }
  void load_xs(bool from_s, int i, uint32_t (&r)[KP]) const { get(r, (from_s ? S : X) + i * K); }
  void load_y(int j, uint32_t (&r)[KP]) const { get(r, Y + j * K); }
  void load_q(int i, uint32_t (&r)[KP]) const { get(r, Q + i * K); }
  void load_n(int j, uint32_t (&r)[KP]) const { get(r, N + j * K); }
  void load_ninv(uint32_t (&r)[KP]) const { get(r, NI); }
  static constexpr int VW = (KP % 4 == 0) ? 4 : 2;
  struct Prefetch { const uint32_t* base; };
  Prefetch prefetch_desc(int kind, int blk) const {
    if (kind == dkg::PAIR_XY) return Prefetch{Y + blk * K};
    if (kind == dkg::PAIR_XX || kind == dkg::PAIR_XX2 || kind == dkg::PAIR_SX2) return Prefetch{X + blk * K};
    if (kind == dkg::PAIR_NQ) return Prefetch{Q + blk * K};
    if (kind == dkg::PAIR_SY2) return Prefetch{Y2 + blk * K};
    return Prefetch{nullptr};
  }
  void prefetch_load(const Prefetch& pf, int v, uint32_t (&r)[KP]) const {
    if (!pf.base) return;
    for (int e = v * VW; e < v * VW + VW; ++e) r[e] = e < K ? pf.base[e] : 0u;
  }
  void store_q(int i, const uint32_t (&r)[KP]) const { std::memcpy(Q + i * K, r, K * 4); }
  void store_x(int i, const uint32_t (&r)[KP]) const { std::memcpy(X + i * K, r, K * 4); }
};

template <int K, int M>
int run(int mode, uint32_t* x, const uint32_t* y, const uint32_t* n, const uint32_t* ninv, int canon,
        const uint32_t* s_op, const uint32_t* y2) {
  uint32_t q[K * M];
  std::memset(q, 0, sizeof(q));
  HostIO<K, M> io{x, y, q, n, ninv};
  if (mode == 0) dkg::mont_mul<K, M>(io, dkg::MONT_MUL);
  else if (mode == 1) { io.Y = x; dkg::mont_mul<K, M>(io, dkg::MONT_MUL); }
  else if (mode == 3) dkg::mont_mul<K, M>(io, dkg::MONT_SQR);
  else if (mode == 4) { io.S = y; dkg::mont_mul<K, M>(io, dkg::MONT_MUL2S); }                 // x <- 2*x*y/R
  else if (mode == 5) { io.S = s_op; io.Y2 = y2; dkg::mont_mul<K, M>(io, dkg::MONT_MULADD); }  // x <- (x*y + s*y2)/R
  else dkg::mont_mul<K, M>(io, dkg::MONT_REDC);
  if (canon) dkg::canonicalize<K, M>(io, canon);
  return 0;
}
}  // namespace

#define CASE(K_, M_) if (K == K_ && M == M_) return run<K_, M_>(mode, x, y, n, ninv, canon, s_op, y2);

// mode 0: x <- x*y/R mod n; 1: x <- x*x/R (y aliased to x, in place); 2: x <- x/R;
// 3: x <- x*x/R through the dedicated squaring path; 4: x <- 2*x*y/R (y as the second shared
// operand); 5: x <- (x*y + s*y2)/R.
extern "C" int host_mont2(int K, int M, int mode, uint32_t* x, const uint32_t* y, const uint32_t* n,
                          const uint32_t* ninv, int canon, const uint32_t* s_op, const uint32_t* y2) {
  CASE(4, 1) CASE(4, 3) CASE(4, 2) CASE(4, 5) CASE(6, 3) CASE(8, 4) CASE(12, 3) CASE(16, 2)
  CASE(16, 8) CASE(12, 11) CASE(22, 3) CASE(22, 6) CASE(16, 16) CASE(14, 5) CASE(12, 6) CASE(14, 7) CASE(13, 5) CASE(13, 10) CASE(5, 3) CASE(7, 2) CASE(9, 1)
  return -1;
}

extern "C" int host_mont(int K, int M, int mode, uint32_t* x, const uint32_t* y, const uint32_t* n,
                         const uint32_t* ninv, int canon) {
  return host_mont2(K, M, mode, x, y, n, ninv, canon, nullptr, nullptr);
}
