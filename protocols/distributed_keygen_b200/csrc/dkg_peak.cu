// Integer-multiplier roofline probe: register-resident multiply-accumulate loops with no memory
// traffic, timed with CUDA events.  Two flavours, because they issue at different rates on sm_100a
// (measured, profiles/r01_imad_microbench_*.txt):
//   plain : mad.wide.u32 d, a, b, d            -> IMAD.WIDE.U32 (64-bit accumulate, no carry), or
//                                                 IMAD.WIDE.U32(RZ) + IADD3/IADD3.X where ptxas splits it
//   carry : mad.lo.cc / madc.hi.cc chains      -> IMAD.WIDE.U32.X      (carry in/out via predicate)
// ptxas schedules every IMAD.WIDE at a 4-cycle issue interval per SM sub-partition (32 lanes / 4
// cycles x 4 sub-partitions = 32 wide-MAC/clk/SM, 9.3 T/s at 1965 MHz); both probes measure how
// close register-resident code gets to that.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../../include/dkg_b200.h"

namespace {

__global__ void peak_plain_kernel(uint32_t* out, int iters, uint32_t seed) {
  // 64 distinct products a[i]*b[j] per iteration into 16 accumulators; the operands change every
  // iteration so that nothing is loop-invariant (an earlier version with constant operands was
  // folded by ptxas into a handful of instructions and reported a fictitious 2x rate).
  uint64_t acc[16];
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = threadIdx.x * 2654435761u + seed + i * 97u;
    b[i] = blockIdx.x * 40503u + 12345u + seed * (i + 3);
  }
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = (uint64_t)(a[i & 7] + i) << 17;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[(i + j) & 15]) : "r"(a[i]), "r"(b[j]));
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] ^= (uint32_t)it; b[i] += (uint32_t)it; }
  }
  uint64_t r = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) r ^= acc[i];
  if (r == 0x123456789ull) out[0] = (uint32_t)r;
}

__global__ void peak_carry_kernel(uint32_t* out, int iters, uint32_t seed) {
  // four independent chains of eight 64-bit accumulate steps each
  uint32_t acc[4][18];
  uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 12345u + seed;
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int i = 0; i < 18; i++) acc[c][i] = a + c * 18 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
      for (int c = 0; c < 4; c++) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                     : "+r"(acc[c][0]), "+r"(acc[c][1]) : "r"(a), "r"(b));
#pragma unroll
        for (int i = 2; i < 16; i += 2)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                       : "+r"(acc[c][i]), "+r"(acc[c][i + 1]) : "r"(a), "r"(b));
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[c][16]));
      }
  }
  uint32_t r = 0;
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int i = 0; i < 18; i++) r ^= acc[c][i];
  if (r == 0x12345678u) out[0] = r;
}

template <typename F>
cudaError_t time_best(F launch, int reps, float* best_ms) {
  cudaEvent_t e0, e1;
  cudaError_t e = cudaEventCreate(&e0);
  if (e != cudaSuccess) return e;
  e = cudaEventCreate(&e1);
  if (e != cudaSuccess) return e;
  *best_ms = 1e30f;
  for (int r = 0; r < reps + 1; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < *best_ms) *best_ms = ms;  // first run is warm-up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return e;
}

}  // namespace

extern "C" int dkg_measure_imad_peak(int device, double* plain_wide_mac_per_s, double* carry_wide_mac_per_s) {
  cudaError_t e = cudaSetDevice(device);
  int sms = 0;
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  uint32_t* out = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&out, 256);
  if (e != cudaSuccess) return DKG_ERR_CUDA;
  const int threads = 512, iters = 20000;
  float ms_plain = 0, ms_carry = 0;
  e = time_best([&] { peak_plain_kernel<<<sms, threads>>>(out, iters, 1u); }, 3, &ms_plain);
  if (e == cudaSuccess) e = time_best([&] { peak_carry_kernel<<<sms, threads>>>(out, iters, 1u); }, 3, &ms_carry);
  cudaFree(out);
  if (e != cudaSuccess) return DKG_ERR_CUDA;
  const double total_threads = (double)sms * threads;
  if (plain_wide_mac_per_s) *plain_wide_mac_per_s = total_threads * iters * 64.0 / (ms_plain * 1e-3);
  if (carry_wide_mac_per_s) *carry_wide_mac_per_s = total_threads * iters * 64.0 / (ms_carry * 1e-3);
  return DKG_OK;
}
