// Self-test of the tcgen05 building blocks used by encoder_tc.cu: one 128 x N x K tile product
// D = A * B^T with A written to shared memory by the threads (generic proxy + fence.proxy.async, as the
// encoder's aggregation / epilogue do) and B delivered as a pre-swizzled tile image by a bulk async copy
// (as the encoder's weight stream).  tests/test_gpu_tc.py compares D with a float64 product.
#include "common.cuh"
#include "tc_ptx.cuh"

using namespace tc;

template <typename T>
__global__ void __launch_bounds__(128, 1)
selftest_umma_kernel(const T* __restrict__ A, const uint8_t* __restrict__ Bimg, float* __restrict__ D, int K, int N) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nkb = K / 64;
  uint8_t* sA = smem;                              // nkb tiles [128][128 B]
  uint8_t* sB = sA + (size_t)nkb * 16384;          // nkb tiles [N][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)nkb * N * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bars[0], 1);   // B landed
    mbar_init(&bars[1], 1);   // MMA done
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    const uint32_t bytes = (uint32_t)nkb * N * 128;
    mbar_arrive_expect_tx(&bars[0], bytes);
    for (uint32_t off = 0; off < bytes; off += 16384)
      bulk_g2s(sB + off, Bimg + off, min(16384u, bytes - off), &bars[0]);
  }
  // A: row-major global -> swizzled tiles, 16-byte chunks
  for (int e = tid; e < 128 * (K / 8); e += 128) {
    const int r = e / (K / 8), ch = e % (K / 8);
    const int kb = ch >> 3, c = ch & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * K + ch * 8);
    *reinterpret_cast<uint4*>(sA + (size_t)kb * 16384 + sw128_off(r, c)) = v;
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc(128, (uint32_t)N, Elem<T>::fmt);
    for (int kb = 0; kb < nkb; ++kb)
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = umma_desc_sw128(smem_u32(sA + (size_t)kb * 16384) + k * 32);
        const uint64_t db = umma_desc_sw128(smem_u32(sB + (size_t)kb * N * 128) + k * 32);
        umma_f16(tmem, da, db, idesc, (kb | k) ? 1u : 0u);
      }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) D[(size_t)row * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// A (128, K) row-major, Bimg = pre-swizzled image of B (N, K) (packing.sw128_image), D (128, N) fp32.
extern "C" int jampr_selftest_umma(const void* A, const void* Bimg, float* D, int32_t K, int32_t N, int32_t precision,
                                   void* stream) {
  if (!A || !Bimg || !D || K <= 0 || K % 64 || N < 16 || N > 256 || N % 16) return JAMPR_E_BADARG;
  const size_t smem = (size_t)(K / 64) * (16384 + (size_t)N * 128) + 64 + 1024;
  if (smem > 227 * 1024) return JAMPR_E_TOOLARGE;
  cudaStream_t s = (cudaStream_t)stream;
  if (precision == JAMPR_PREC_F16) {
    CUDA_RET(cudaFuncSetAttribute(selftest_umma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selftest_umma_kernel<__half><<<1, 128, smem, s>>>((const __half*)A, (const uint8_t*)Bimg, D, K, N);
  } else if (precision == JAMPR_PREC_BF16) {
    CUDA_RET(cudaFuncSetAttribute(selftest_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    selftest_umma_kernel<__nv_bfloat16><<<1, 128, smem, s>>>((const __nv_bfloat16*)A, (const uint8_t*)Bimg, D, K, N);
  } else {
    return JAMPR_E_BADARG;
  }
  return launch_status();
}
