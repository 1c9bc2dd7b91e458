// Probe: MUFU.EX2 issue rate per SM sub-partition (cycles per warp instruction) with 1, 2, 4 warps per sub-partition,
// alone and interleaved with FFMA2-like FMA work.  nvcc -arch=sm_100a -o /tmp/mufu tools/mufu_probe.cu && /tmp/mufu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, long long* cyc, int iters, int with_fma) {
  float x[8];
  for (int j = 0; j < 8; ++j) x[j] = -0.001f * (threadIdx.x + j);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[j]));
      if (with_fma) { acc = fmaf(y, 1.0001f, acc); x[j] = fmaf(x[j], 0.9999f, -1e-6f); } else x[j] = y - 1.0f;
    }
  }
  const long long t1 = clock64();
  float s = acc;
  for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int with_fma = 0; with_fma < 2; ++with_fma)
    for (int warps : {4, 8, 16, 32}) {
      k<<<148, warps * 32>>>(out, cyc, iters, with_fma);
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double per_smsp_instr = (double)iters * 8 * (warps / 4);
      printf("fma=%d warps/SM=%2d (%d per sub-partition): %.2f cycles per MUFU warp-instruction per sub-partition\n", with_fma, warps, warps / 4,
             (double)c / per_smsp_instr);
    }
  return 0;
}
