// write-bandwidth of the fill kernels' pointer-store pattern: groups of 8 lanes each stream
// 128 B rows into a private slice; footprint = nslices * slice_bytes, recycled `reps` times
#include <cstdio>
#include <cuda_runtime.h>
__global__ void wr(uint4* base, long long slice_rows, long long nslices, int rows_per_pair, int pairs) {
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 8;
  const int l = threadIdx.x & 7;
  for (int p = 0; p < pairs; p++) {
    const long long slice = (gid + (long long)p * (gridDim.x * (blockDim.x / 8))) % nslices;
    uint4* dst = base + slice * slice_rows * 8 + l;
    for (int x = 0; x < rows_per_pair; x++) {
      uint4 v = make_uint4(x, p, l, (unsigned)gid);
      dst[(long long)x * 8] = v;
      // a little ALU work between stores, like the fill
      unsigned a = v.x;
#pragma unroll
      for (int i = 0; i < 64; i++) a = a * 1664525u + 1013904223u;
      if (a == 0x12345678u) dst[0] = v;
    }
  }
}
int main() {
  const int sms = 148, blocks = sms * 4, threads = 128, groups = blocks * threads / 8;
  const int rows = 151, pairs = 8;
  uint4* d; size_t cap = (size_t)6 << 30; cudaMalloc(&d, cap);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  long long foot_mb[] = {16, 64, 100, 160, 366, 1000, 4000};
  for (long long f : foot_mb) {
    long long nslices = f * 1000000 / (rows * 128);
    if (nslices < 1) nslices = 1;
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      wr<<<blocks, threads>>>(d, rows, nslices, rows, pairs);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double bytes = (double)groups * pairs * rows * 128;
      if (rep) printf("footprint %5lld MB (%lld slices): %.3f ms  %.0f GB/s written\n", f, nslices, ms, bytes / ms / 1e6);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
