// bsq_microbench.cu -- pipe-rate microbenchmarks (roofline denominators).
// Each kernel runs NCHAIN independent dependency chains per thread, unrolled,
// so that the pipe under test is the only limiter.
#include <string>
#include "../../include/bsqgpu.h"
#include "bsq_common.h"

#ifndef BSQ_SIMT_EMU

#define MB_ITERS 512
#define MB_UNROLL 8

template <int WHICH>
__global__ void __launch_bounds__(256) bsq_mb_kernel(double* sink, int iters, double c0, int ci) {
  __shared__ double sh[256 * 2];
  sh[threadIdx.x] = c0 + threadIdx.x;
  sh[threadIdx.x + 256] = c0 - threadIdx.x;
  __syncthreads();
  double d[8];
  int v[8], w[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { d[k] = c0 * (k + 1) + threadIdx.x; v[k] = ci * (k + 1) + threadIdx.x; w[k] = ci - k; }
  const double dc = c0 * 0.5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < MB_UNROLL; u++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (WHICH == BSQ_MB_DADD) {
          asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[k]) : "d"(dc));
        } else if (WHICH == BSQ_MB_DMAX) {
          asm volatile("{ .reg .pred p; setp.gt.f64 p, %0, %1; selp.f64 %0, %0, %1, p; }" : "+d"(d[k]) : "d"(d[(k + 1) & 7]));
        } else if (WHICH == BSQ_MB_DSETP) {
          asm volatile("{ .reg .pred p; setp.eq.f64 p, %1, %2; @p or.b32 %0, %0, 2; }" : "+r"(v[k]) : "d"(d[k]), "d"(dc));
        } else if (WHICH == BSQ_MB_IADD) {
          asm volatile("add.s32 %0, %0, %1;" : "+r"(v[k]) : "r"(w[k]));
          asm volatile("add.s32 %0, %0, %1;" : "+r"(w[k]) : "r"(v[k]));
        } else if (WHICH == BSQ_MB_VIMAX3) {
          v[k] = __vimax3_s32(v[k], w[k], v[(k + 1) & 7] + it);
        } else if (WHICH == BSQ_MB_VIADDMAX) {
          v[k] = __viaddmax_s32(v[k], w[k], v[(k + 3) & 7]);
        } else if (WHICH == BSQ_MB_VADD2) {
          v[k] = (int)__vadd2((unsigned)v[k], (unsigned)w[k]);
          w[k] = (int)__vadd2((unsigned)w[k], (unsigned)v[k]);
        } else if (WHICH == BSQ_MB_SHFL) {
          v[k] = __shfl_up_sync(0xffffffffu, v[k], 1);
        } else if (WHICH == BSQ_MB_LDS64) {
          d[k] += sh[(threadIdx.x + (v[k] & 256)) & 511];
          v[k] ^= 256;
        }
      }
    }
  }
  double acc = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) acc += d[k] + v[k] + w[k];
  if (acc == 123.456) sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int WHICH>
static int run_mb(int sms, cudaStream_t st, double* ops_per_sec, std::string* err) {
  double* sink = nullptr;
  const int grid = sms * 8, block = 256;
  if (cudaMalloc(&sink, sizeof(double) * grid * block) != cudaSuccess) { *err = "cudaMalloc failed"; return BSQ_ERR_NOMEM; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0, st);
    bsq_mb_kernel<WHICH><<<grid, block, 0, st>>>(sink, MB_ITERS, 1.000001, 3);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (e != cudaSuccess) { *err = cudaGetErrorString(e); return BSQ_ERR_CUDA; }
  double per_thread = (double)MB_ITERS * MB_UNROLL * 8;
  if (WHICH == BSQ_MB_IADD) per_thread *= 2;
  if (WHICH == BSQ_MB_VADD2) per_thread *= 4;  // two instructions, two halfword adds each
  *ops_per_sec = per_thread * grid * block / (best * 1e-3);
  return BSQ_SUCCESS;
}

int bsq_microbench_impl(int sms, cudaStream_t st, int which, double* out, std::string* err) {
  switch (which) {
    case BSQ_MB_DADD: return run_mb<BSQ_MB_DADD>(sms, st, out, err);
    case BSQ_MB_DMAX: return run_mb<BSQ_MB_DMAX>(sms, st, out, err);
    case BSQ_MB_DSETP: return run_mb<BSQ_MB_DSETP>(sms, st, out, err);
    case BSQ_MB_IADD: return run_mb<BSQ_MB_IADD>(sms, st, out, err);
    case BSQ_MB_VIMAX3: return run_mb<BSQ_MB_VIMAX3>(sms, st, out, err);
    case BSQ_MB_VIADDMAX: return run_mb<BSQ_MB_VIADDMAX>(sms, st, out, err);
    case BSQ_MB_VADD2: return run_mb<BSQ_MB_VADD2>(sms, st, out, err);
    case BSQ_MB_SHFL: return run_mb<BSQ_MB_SHFL>(sms, st, out, err);
    case BSQ_MB_LDS64: return run_mb<BSQ_MB_LDS64>(sms, st, out, err);
  }
  *err = "unknown microbenchmark";
  return BSQ_ERR_ARG;
}

#else  // the emulator has no pipes to measure

int bsq_microbench_impl(int, cudaStream_t, int, double* out, std::string* err) {
  *out = 0;
  *err = "microbenchmarks need a GPU";
  return BSQ_ERR_CUDA;
}

#endif
