// Scratch microbenchmark: which IMAD flavour / dependency shape saturates the sm_100a multiplier.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_bench imad_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// V0: independent mad.wide.u32, 8 accumulators
__global__ void v_wide_indep(uint32_t* out, int iters, uint32_t s) {
  uint64_t acc[8];
  uint32_t a = threadIdx.x * 2654435761u + s, b = blockIdx.x * 40503u + 12345u + s;
  for (int i = 0; i < 8; i++) acc[i] = (uint64_t)(a + i) << 17;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a), "r"(b));
  }
  uint64_t r = 0;
  for (int i = 0; i < 8; i++) r ^= acc[i];
  if (r == 0x123456789ull) out[0] = (uint32_t)r;
}
// V1: independent mad.lo.u32
__global__ void v_lo_indep(uint32_t* out, int iters, uint32_t s) {
  uint32_t acc[8];
  uint32_t a = threadIdx.x * 2654435761u + s, b = blockIdx.x * 40503u + 12345u + s;
  for (int i = 0; i < 8; i++) acc[i] = (a + i) << 3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
  }
  uint32_t r = 0;
  for (int i = 0; i < 8; i++) r ^= acc[i];
  if (r == 0x12345678u) out[0] = r;
}
// V2: independent mad.hi.u32
__global__ void v_hi_indep(uint32_t* out, int iters, uint32_t s) {
  uint32_t acc[8];
  uint32_t a = threadIdx.x * 2654435761u + s, b = blockIdx.x * 40503u + 12345u + s;
  for (int i = 0; i < 8; i++) acc[i] = (a + i) << 3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
  }
  uint32_t r = 0;
  for (int i = 0; i < 8; i++) r ^= acc[i];
  if (r == 0x12345678u) out[0] = r;
}

// V3: row-chain schoolbook KxK fresh product into even/odd accumulators (the shape the engine uses)
template <int K>
__device__ __forceinline__ void block_mul_rows(uint32_t (&E)[2 * K + 2], uint32_t (&O)[2 * K + 2],
                                               const uint32_t (&a)[K], const uint32_t (&b)[K]) {
#pragma unroll
  for (int i = 0; i < 2 * K + 2; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
  for (int j = 0; j < K; j++) {
    // products a_i*b_j with (i+j) even go to E at position i+j, odd go to O at position i+j-1
    if ((j & 1) == 0) {
      // even i -> E[i+j], odd i -> O[i+j-1]
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(E[j]), "+r"(E[j + 1]) : "r"(a[0]), "r"(b[j]));
#pragma unroll
      for (int i = 2; i < K; i += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(E[i + j]), "+r"(E[i + j + 1]) : "r"(a[i]), "r"(b[j]));
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(E[j + K]));
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(O[j]), "+r"(O[j + 1]) : "r"(a[1]), "r"(b[j]));
#pragma unroll
      for (int i = 3; i < K; i += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(O[i + j - 1]), "+r"(O[i + j]) : "r"(a[i]), "r"(b[j]));
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(O[j + K]));
    } else {
      // odd i -> E[i+j], even i -> O[i+j-1]
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(E[j + 1]), "+r"(E[j + 2]) : "r"(a[1]), "r"(b[j]));
#pragma unroll
      for (int i = 3; i < K; i += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(E[i + j]), "+r"(E[i + j + 1]) : "r"(a[i]), "r"(b[j]));
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(E[j + K + 1]));
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(O[j - 1]), "+r"(O[j]) : "r"(a[0]), "r"(b[j]));
#pragma unroll
      for (int i = 2; i < K; i += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(O[i + j - 1]), "+r"(O[i + j]) : "r"(a[i]), "r"(b[j]));
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(O[j + K - 1]));
    }
  }
}

template <int K>
__global__ void v_rows(uint32_t* out, int iters, uint32_t s) {
  uint32_t a[K], b[K], E[2 * K + 2], O[2 * K + 2];
  for (int i = 0; i < K; i++) { a[i] = threadIdx.x * 2654435761u + s + i; b[i] = blockIdx.x * 40503u + i * 7 + s; }
  for (int it = 0; it < iters; it++) {
    block_mul_rows<K>(E, O, a, b);
#pragma unroll
    for (int i = 0; i < K; i++) { a[i] ^= E[i] + O[i + K]; b[i] += E[i + K] ^ O[i]; }
  }
  uint32_t r = 0;
  for (int i = 0; i < K; i++) r ^= a[i] ^ b[i];
  if (r == 0x12345678u) out[0] = r;
}

// V4: Comba (column-wise) KxK product with NS independent partial sums per column
template <int K, int NS>
__global__ void v_comba(uint32_t* out, int iters, uint32_t s) {
  uint32_t a[K], b[K], r_[2 * K];
  for (int i = 0; i < K; i++) { a[i] = threadIdx.x * 2654435761u + s + i; b[i] = blockIdx.x * 40503u + i * 7 + s; }
  for (int it = 0; it < iters; it++) {
    uint32_t c0[NS], c1[NS], c2[NS];
#pragma unroll
    for (int q = 0; q < NS; q++) { c0[q] = 0; c1[q] = 0; c2[q] = 0; }
#pragma unroll
    for (int p = 0; p < 2 * K - 1; p++) {
      int n = 0;
#pragma unroll
      for (int i = 0; i < K; i++) {
        int j = p - i;
        if (j < 0 || j >= K) continue;
        int q = n % NS; n++;
        asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                     : "+r"(c0[q]), "+r"(c1[q]), "+r"(c2[q]) : "r"(a[i]), "r"(b[j]));
      }
      // merge partial sums into slot 0
#pragma unroll
      for (int q = 1; q < NS; q++) {
        asm volatile("add.cc.u32 %0, %0, %3; addc.cc.u32 %1, %1, %4; addc.u32 %2, %2, %5;"
                     : "+r"(c0[0]), "+r"(c1[0]), "+r"(c2[0]) : "r"(c0[q]), "r"(c1[q]), "r"(c2[q]));
        c0[q] = 0; c1[q] = 0; c2[q] = 0;
      }
      r_[p] = c0[0]; c0[0] = c1[0]; c1[0] = c2[0]; c2[0] = 0;
    }
    r_[2 * K - 1] = c0[0];
#pragma unroll
    for (int i = 0; i < K; i++) { a[i] ^= r_[i]; b[i] += r_[i + K]; }
  }
  uint32_t r = 0;
  for (int i = 0; i < K; i++) r ^= a[i] ^ b[i];
  if (r == 0x12345678u) out[0] = r;
}

template <typename F>
static int run(const char* name, F launch, double macs_per_thread_iter, int iters) {
  int cfgs[][2] = {{148, 128}, {148, 256}, {148, 512}, {148, 1024}, {296, 128}, {296, 224}, {148, 224}, {148, 32}, {148, 64}};
  for (auto& c : cfgs) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(c[0], c[1], iters / 4);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(e0));
      launch(c[0], c[1], iters);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    double macs = macs_per_thread_iter * (double)iters * c[0] * c[1];
    printf("%-18s grid=%4d block=%4d  %8.3f ms  %8.3f Tmac/s  (%.2f mac/clk/SM @1.965GHz)\n", name, c[0], c[1], best,
           macs / best / 1e9, macs / (best * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}

int main() {
  uint32_t* out; CK(cudaMalloc(&out, 4096));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  run("wide_indep", [&](int g, int b, int it) { v_wide_indep<<<g, b>>>(out, it, 1); }, 64, 4000);
  run("lo_indep", [&](int g, int b, int it) { v_lo_indep<<<g, b>>>(out, it, 1); }, 64, 4000);
  run("hi_indep", [&](int g, int b, int it) { v_hi_indep<<<g, b>>>(out, it, 1); }, 64, 4000);
  run("rows_k16", [&](int g, int b, int it) { v_rows<16><<<g, b>>>(out, it, 1); }, 256, 1000);
  run("rows_k8", [&](int g, int b, int it) { v_rows<8><<<g, b>>>(out, it, 1); }, 64, 4000);
  run("comba_k16_ns1", [&](int g, int b, int it) { v_comba<16, 1><<<g, b>>>(out, it, 1); }, 256, 1000);
  run("comba_k16_ns2", [&](int g, int b, int it) { v_comba<16, 2><<<g, b>>>(out, it, 1); }, 256, 1000);
  run("comba_k16_ns4", [&](int g, int b, int it) { v_comba<16, 4><<<g, b>>>(out, it, 1); }, 256, 1000);
  return 0;
}
