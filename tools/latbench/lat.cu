// dependent-chain latency of a few sm_100a integer / DPX instructions (cycles per op)
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int W> __global__ void k(long long* out, unsigned a, unsigned b) {
  unsigned x = a + threadIdx.x, y = b;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) {
    if (W == 0) x = x + y;                                   // IADD
    else if (W == 1) x = (unsigned)__vimax3_s32((int)x, (int)y, (int)(x >> 1));
    else if (W == 2) x = __vadd2(x, y);
    else if (W == 3) x = __vimax3_s16x2_relu(x, y, y + 1);
    else if (W == 4) x = __viaddmin_u16x2(x, y, 0x00010001u) + y;  // + 1 IADD
    else if (W == 5) { bool ph, pl; x = __vibmax_s16x2(x, y, &ph, &pl); y += pl ? 1u : 0u; }
    else if (W == 6) x = __vminu2(x, y) + y;
    else if (W == 7) x = x * y + 3u;                         // IMAD
    else if (W == 8) x = (unsigned)__viaddmax_s32((int)x, (int)y, 7);
    else if (W == 9) { bool ph, pl; x = __vibmax_s16x2(x, y, &ph, &pl); y = pl ? y : x; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[2 * W] = t1 - t0; out[2 * W + 1] = x + y; }
}
int main() {
  long long* d; cudaMalloc(&d, 8 * 64);
  k<0><<<1, 32>>>(d, 1, 2); k<1><<<1, 32>>>(d, 1, 2); k<2><<<1, 32>>>(d, 1, 2); k<3><<<1, 32>>>(d, 1, 2);
  k<4><<<1, 32>>>(d, 1, 2); k<5><<<1, 32>>>(d, 1, 2); k<6><<<1, 32>>>(d, 1, 2); k<7><<<1, 32>>>(d, 1, 2);
  k<8><<<1, 32>>>(d, 1, 2); k<9><<<1, 32>>>(d, 1, 2);
  long long h[64]; cudaMemcpy(h, d, 8 * 64, cudaMemcpyDeviceToHost);
  const char* nm[] = {"IADD", "VIMNMX3.s32", "VIADD.16x2", "VIMNMX3.S16x2.RELU", "VIADDMNMX.U16x2 (+IADD)", "VIMNMX.S16x2 pred (+SEL/IADD)", "VIMNMX.U16x2 (+IADD)", "IMAD", "VIADDMNMX.s32", "VIMNMX.S16x2 pred -> SEL"};
  for (int w = 0; w < 10; w++) printf("%-32s %.2f cycles/iter\n", nm[w], (double)h[2 * w] / N);
  return 0;
}
