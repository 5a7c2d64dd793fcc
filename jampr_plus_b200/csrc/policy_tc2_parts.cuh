// Device pieces shared by the per-rollout GNN kernels (policy_tc2.cu: one rollout per CTA, two CTAs per SM;
// policy_duo.cu: two rollouts per CTA with shared weight stages): weight-stage loads, max aggregation into the operand
// tiles, layer epilogue, shared-memory sizing.  `tid` is the thread index inside the 256-thread group that owns the
// rollout (the whole CTA in policy_tc2.cu).
#pragma once
#include "policy_tc_common.cuh"

#define P2_THREADS 256
#define P2_WARPS (P2_THREADS / 32)
#define P2_GROUPS ((P2_THREADS - 32) / 8)    // customer rows aggregated concurrently (warps 1..7, 8 threads per row)
#define NE16_STRIDE 528                       // bytes per parked node-embedding row: 256 x 16 bit + 16 (bank shift 4 words)
#define QP2_STRIDE 260

template <typename T>
__device__ __forceinline__ void load_w2(uint8_t* sW, int slot, const void* w16, int img, uint64_t* full) {
  mbar_arrive_expect_tx(&full[slot], 16384u);
  bulk_g2s(sW + slot * 16384, reinterpret_cast<const uint8_t*>(w16) + ((size_t)(W16_ENC + img * W16_STAGE)) * 2, 16384u,
           &full[slot]);
}

template <typename T>
__device__ __forceinline__ void agg_nbh_tc2(uint8_t* sA, uint32_t ts, uint32_t tab_saddr, int k4, uint32_t dtab_saddr,
                                            int nd, int N, int tid) {
  using T2 = typename Elem<T>::T2;
  const uint32_t h_saddr = smem_u32(sA) + 2u * ts;
  if (tid < 32) {
    const int cc = tid & 15, half = tid >> 4;
    const uint32_t kboff = (uint32_t)(cc >> 3) * ts, c = cc & 7;
    uint4 r = agg_run<T>(dtab_saddr, half * (nd >> 1), (half + 1) * (nd >> 1), h_saddr + kboff, c << 4);
    r.x = as_u32(__hmax2(from_u32<T2>(r.x), from_u32<T2>(__shfl_xor_sync(0xffffffffu, r.x, 16))));
    r.y = as_u32(__hmax2(from_u32<T2>(r.y), from_u32<T2>(__shfl_xor_sync(0xffffffffu, r.y, 16))));
    r.z = as_u32(__hmax2(from_u32<T2>(r.z), from_u32<T2>(__shfl_xor_sync(0xffffffffu, r.z, 16))));
    r.w = as_u32(__hmax2(from_u32<T2>(r.w), from_u32<T2>(__shfl_xor_sync(0xffffffffu, r.w, 16))));
    if (half == 0) *reinterpret_cast<uint4*>(sA + kboff + sw128_off(0, c)) = r;
  } else {
    // 8 threads per customer row, two adjacent 16-byte chunks each (chunk pair cp of K-block cp >> 2)
    const int g = (tid - 32) >> 3, cp = (tid - 32) & 7;
    // chunks (c, c ^ 4); the K-block-1 threads start with the upper chunk so that the eight threads of a row hit
    // eight distinct 16-byte bank groups in each of their two loads
    const uint32_t kboff = (uint32_t)(cp >> 2) * ts, cA = (uint32_t)(cp & 3) + (uint32_t)(cp >> 2) * 4u, cB = cA ^ 4u;
    for (int u = 1 + g; u < N; u += P2_GROUPS) {
      uint4 ra, rb;
      agg_run2<T>(tab_saddr + (uint32_t)(u * k4) * 4u, 0, k4, h_saddr + kboff, cA << 4, cB << 4, ra, rb);
      *reinterpret_cast<uint4*>(sA + kboff + sw128_off(u, cA)) = ra;
      *reinterpret_cast<uint4*>(sA + kboff + sw128_off(u, cB)) = rb;
    }
  }
}

template <typename T>
__device__ __forceinline__ void agg_tour_tc2(uint8_t* sA, uint32_t ts, const uint32_t* s_ltab, uint32_t etab_saddr,
                                             int n_e4, int N, int tid) {
  using T2 = typename Elem<T>::T2;
  const uint32_t h_saddr = smem_u32(sA) + 2u * ts;
  if (tid < 32) {
    if (tid < 16) {
      const uint32_t kboff = (uint32_t)(tid >> 3) * ts, c = tid & 7;
      uint4 r = make_uint4(0u, 0u, 0u, 0u);
      if (n_e4 > 0) r = agg_run<T>(etab_saddr, 0, n_e4, h_saddr + kboff, c << 4);
      *reinterpret_cast<uint4*>(sA + kboff + sw128_off(0, c)) = r;
    }
  } else {
    const int g = (tid - 32) >> 3, cp = (tid - 32) & 7;
    // chunks (c, c ^ 4); the K-block-1 threads start with the upper chunk so that the eight threads of a row hit
    // eight distinct 16-byte bank groups in each of their two loads
    const uint32_t kboff = (uint32_t)(cp >> 2) * ts, cA = (uint32_t)(cp & 3) + (uint32_t)(cp >> 2) * 4u, cB = cA ^ 4u;
    for (int u = 1 + g; u < N; u += P2_GROUPS) {
      T2 m[8];
      agg_edge2<T>(s_ltab[u], h_saddr + kboff, cA << 4, cB << 4, m, true);
      *reinterpret_cast<uint4*>(sA + kboff + sw128_off(u, cA)) = make_uint4(as_u32(m[0]), as_u32(m[1]), as_u32(m[2]), as_u32(m[3]));
      *reinterpret_cast<uint4*>(sA + kboff + sw128_off(u, cB)) = make_uint4(as_u32(m[4]), as_u32(m[5]), as_u32(m[6]), as_u32(m[7]));
    }
  }
}

// y = GELU(acc + b) + skip for 32 columns; accumulates the packed LayerNorm partials
__device__ __forceinline__ void epi_pass1(uint32_t (&acc)[32], const uint32_t (&sk)[32], uint32_t pb, float2& s2, float2& q2) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b4 = lds_f32x4(pb + i * 16);
    float2 y0 = gelu2(__fadd2_rn(make_float2(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1])), make_float2(b4.x, b4.y)));
    float2 y1 = gelu2(__fadd2_rn(make_float2(__uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3])), make_float2(b4.z, b4.w)));
    y0 = __fadd2_rn(y0, make_float2(__uint_as_float(sk[4 * i]), __uint_as_float(sk[4 * i + 1])));
    y1 = __fadd2_rn(y1, make_float2(__uint_as_float(sk[4 * i + 2]), __uint_as_float(sk[4 * i + 3])));
    s2 = __fadd2_rn(s2, __fadd2_rn(y0, y1));
    q2 = __ffma2_rn(y0, y0, q2);
    q2 = __ffma2_rn(y1, y1, q2);
    acc[4 * i] = __float_as_uint(y0.x); acc[4 * i + 1] = __float_as_uint(y0.y);
    acc[4 * i + 2] = __float_as_uint(y1.x); acc[4 * i + 3] = __float_as_uint(y1.y);
  }
}

// Layer epilogue, two passes.  Thread (q, lane, ch): row 32q + lane, accumulator columns 64ch .. 64ch+63.
// BAR0 = first of the four named barriers on which the warp pairs (q, q + 4) of this thread group meet.
template <typename T, int BAR0>
__device__ __forceinline__ void layer_epilogue2(uint32_t tmem, uint8_t* sA, uint32_t ts, float2* s_red, uint32_t par_saddr,
                                                int N, float* __restrict__ hfin, int tid) {
  const int warp = tid >> 5, lane = tid & 31, q = warp & 3, ch = warp >> 2, r = q * 32 + lane;
  const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16);
  float2 s2 = make_float2(0.0f, 0.0f), q2 = make_float2(0.0f, 0.0f);
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    const int col0 = ch * 64 + c * 32;
    uint32_t acc[32], sk[32];
    tmem_ld32_nowait(ta + col0, acc);
    tmem_ld32_nowait(ta + 128 + col0, sk);
    tmem_wait_ld();
    epi_pass1(acc, sk, par_saddr + (uint32_t)col0 * 4u, s2, q2);
    tmem_st32(ta + col0, acc);
  }
  tmem_wait_st();
  s_red[ch * 128 + r] = make_float2(s2.x + s2.y, q2.x + q2.y);
  // the two halves of a row live in warps q and q + 4: only that pair has to meet (named barrier 1 + q, 64 threads)
  if (q == 0) asm volatile("bar.sync %0, 64;" ::"n"(BAR0) : "memory");            // literal ids: no blanket reservation
  else if (q == 1) asm volatile("bar.sync %0, 64;" ::"n"(BAR0 + 1) : "memory");
  else if (q == 2) asm volatile("bar.sync %0, 64;" ::"n"(BAR0 + 2) : "memory");
  else asm volatile("bar.sync %0, 64;" ::"n"(BAR0 + 3) : "memory");
  const float2 v0 = s_red[r], v1 = s_red[128 + r];
  const float mean = (v0.x + v1.x) * (1.0f / 128.0f);
  const float var = fmaxf(fmaf(-mean, mean, (v0.y + v1.y) * (1.0f / 128.0f)), 0.0f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float2 rstd2 = f2(rstd), shift2 = f2(-mean * rstd);      // (y - mean) rstd = y rstd + shift: one FFMA2
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    const int col0 = ch * 64 + c * 32;
    const uint32_t pb = par_saddr + (uint32_t)col0 * 4u;
    uint32_t acc[32];
    tmem_ld32_nowait(ta + col0, acc);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 g4 = lds_f32x4(pb + 512 + i * 16), e4 = lds_f32x4(pb + 1024 + i * 16);
      const float2 d0 = __ffma2_rn(make_float2(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1])), rstd2, shift2);
      const float2 d1 = __ffma2_rn(make_float2(__uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3])), rstd2, shift2);
      const float2 o0 = __ffma2_rn(d0, make_float2(g4.x, g4.y), make_float2(e4.x, e4.y));
      const float2 o1 = __ffma2_rn(d1, make_float2(g4.z, g4.w), make_float2(e4.z, e4.w));
      acc[4 * i] = __float_as_uint(o0.x); acc[4 * i + 1] = __float_as_uint(o0.y);
      acc[4 * i + 2] = __float_as_uint(o1.x); acc[4 * i + 3] = __float_as_uint(o1.y);
    }
    tmem_st32(ta + 128 + col0, acc);
    if (r < N) {
      uint8_t* tile = sA + (2 + ch) * ts;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          pk[e] = as_u32(Elem<T>::from2(__uint_as_float(acc[i * 8 + 2 * e]), __uint_as_float(acc[i * 8 + 2 * e + 1])));
        *reinterpret_cast<uint4*>(tile + sw128_off(r, c * 4 + i)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      if (hfin) {
        float4* dst = reinterpret_cast<float4*>(hfin + (size_t)r * 128 + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                               __uint_as_float(acc[4 * i + 3]));
      }
    }
  }
  tmem_wait_st();
}

__host__ __device__ inline uint32_t p2_al(uint32_t b) { return (b + 15u) & ~15u; }
__host__ __device__ inline uint32_t p2_tile_stride(int N) { return (uint32_t)((N + 7) / 8) * 1024u; }
// bytes of the region shared by the edge tables (GNN phase) and decoder scratch (decoder phase)
__host__ __device__ inline uint32_t p2_r2_bytes(int N, int k, int A) {
  const uint32_t k4 = (uint32_t)(k + 3) & ~3u, nd = (uint32_t)(N + 7) & ~7u;
  const uint32_t tabs = p2_al((uint32_t)N * k4 * 4) + p2_al(nd * 4) + 2 * p2_al((uint32_t)N * 4) + 2 * p2_al((uint32_t)(N + 4) * 4);
  const uint32_t dec = 2048 + 4096 + 2048 + 512 + 2048 + p2_al(16u * A) + p2_al(4u * A);
  return tabs > dec ? tabs : dec;
}
// bytes of the region shared by the weight stages (GNN phase) and decoder scratch (decoder phase)
#define P2_R1_BYTES 32768u
// accumulator row owned by a thread in the epilogue / h_0 phases: TMEM lane quarter = warp & 3
__device__ __forceinline__ int q4_row(int tid) { return ((tid >> 5) & 3) * 32 + (tid & 31); }

