#include <cstdint>
// check how ptxas fuses mad.lo.cc/madc.hi.cc
__global__ void k(uint32_t* out, const uint32_t* in) {
  uint32_t a[8], acc[10];
  for (int i=0;i<8;i++) a[i]=in[threadIdx.x*8+i];
  for (int i=0;i<10;i++) acc[i]=in[100+threadIdx.x*10+i];
  uint32_t b = in[999];
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(b));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[2]), "+r"(acc[3]) : "r"(a[2]), "r"(b));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[4]), "+r"(acc[5]) : "r"(a[4]), "r"(b));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[6]), "+r"(acc[7]) : "r"(a[6]), "r"(b));
  asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[8]));
  for (int i=0;i<10;i++) out[threadIdx.x*10+i]=acc[i];
}
