// Micro-benchmark: issue rate of the legacy warp-level mma.sync.m16n8k16 (f16 x f16 -> f32) on sm_100a, per SM.
// Used to decide whether the decoder mat-vecs may run on it (decoder section).
#include <cstdio>
#include <cuda_fp16.h>
__global__ void __launch_bounds__(256) k(float* out, int iters, int chains) {
  unsigned a0 = threadIdx.x, a1 = 0, a2 = threadIdx.x * 3, a3 = 0, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < chains)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 2 * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int chains : {1, 2, 4, 8}) {
    for (int ctas : {148, 296}) {
      const int iters = 20000;
      k<<<ctas, 256>>>(out, 100, chains);
      cudaEventRecord(e0);
      k<<<ctas, 256>>>(out, iters, chains);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double mma_per_sm = (double)iters * chains * 8 * (ctas / 148);
      const double cyc = ms * 1e-3 * 1.965e9;
      printf("chains %d ctas/SM %d: %.3f ms, %.2f cycles per warp-MMA per SM (%.1f per sub-partition), %.0f dense TFLOP/s chip\n", chains,
             ctas / 148, ms, cyc / mma_per_sm, 4 * cyc / mma_per_sm, 148 * mma_per_sm * 4096 / (ms * 1e-3) / 1e12);
    }
  }
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
