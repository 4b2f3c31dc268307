#include <cstdint>
// realistic lazy-carry block MAC: persistent 64-bit column accumulators, operands from shared memory
template <int K>
__global__ void __launch_bounds__(256,1) lazy_kernel(uint32_t* out, const uint32_t* in, int iters) {
  extern __shared__ uint32_t sm[];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int q = 0; q < 16 * K; q++) sm[q * nt + tid] = in[q * nt + tid] & 0x0fffffffu;
  __syncthreads();
  uint64_t col[2 * K];
#pragma unroll
  for (int i = 0; i < 2 * K; i++) col[i] = in[i] + 1;
  uint32_t mix = 0;
  for (int it = 0; it < iters; it++) {
    uint32_t x[K], y[K];
    const int bx = it & 7, by = (it * 5 + 3) & 7;
#pragma unroll
    for (int i = 0; i < K; i++) { x[i] = sm[(bx * K + i) * nt + tid]; y[i] = sm[((8 + by) * K + i) * nt + tid]; }
#pragma unroll
    for (int i = 0; i < K; i++)
#pragma unroll
      for (int j = 0; j < K; j++)
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(col[i + j]) : "r"(x[i]), "r"(y[j]));
    if ((it & 3) == 3) {
      // normalise: 28-bit digits
#pragma unroll
      for (int p = 0; p < 2 * K - 1; p++) { col[p + 1] += col[p] >> 28; col[p] &= 0x0fffffffull; }
#pragma unroll
      for (int p = 0; p < K; p++) mix ^= (uint32_t)col[p];
#pragma unroll
      for (int p = 0; p < K; p++) { col[p] = col[p + K]; col[p + K] = 0; }
    }
  }
  uint64_t r = mix;
#pragma unroll
  for (int i = 0; i < 2 * K; i++) r ^= col[i];
  out[blockIdx.x * nt + tid] = (uint32_t)r ^ (uint32_t)(r >> 32);
}
template __global__ void lazy_kernel<16>(uint32_t*, const uint32_t*, int);
template __global__ void lazy_kernel<20>(uint32_t*, const uint32_t*, int);
