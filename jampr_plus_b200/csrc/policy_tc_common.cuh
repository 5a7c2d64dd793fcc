// Shared device helpers of the tensor-core policy-step kernels (policy_tc.cu, policy_tc2.cu).
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

using namespace tc;

// 16-bit weight blob (elements), mirrors packing.pack_tc
#define W16_STAGE 8192                                  // one 16 KB stage = [128 rows][64 K] elements
#define W16_ENC 0                                       // 7 x 4 stages: layers 0..5, then node_emb_output_proj
#define W16_TO2 (W16_ENC + 28 * W16_STAGE)              // [128][256]  tour output_proj, per-tour-mean half, transposed
#define W16_CTXT (W16_TO2 + 128 * 256)                  // [512][256]  ctxt_proj^T
#define W16_GK (W16_CTXT + 512 * 256)                   // [256][256]  set_proj rows 0..255 as stored
#define W16_GV (W16_GK + 256 * 256)                     // [256][256]  set_proj rows 256..511, transposed
#define W16_M1 (W16_GV + 256 * 256)                     // [256][256]  (W_lk^T W_out)^T
#define W16_NE (W16_M1 + 256 * 256)                     // 20 stages: node-encoder layers 0..2, output_proj halves, TE input proj
// split-decoder GEMM operands (decoder_tc.cu), B tiles [rows][64 K] in the 128-byte swizzle, a hi part followed by a lo
// part (w = hi + lo, both 16-bit): ctxt_proj (256 x 512) as 8 K-block tiles; per head B[c][d] = set_proj[64h + d][c]
// (256 x 64); per head 4 K-block tiles of set_proj[256 + 64h .. +64] (64 x 256); M1 = W_lk^T W_out (256 x 256) as 4 tiles;
// tour_encoder.output_proj[:, 128:] (256 x 128, the per-tour-mean half) as 2 tiles + the folded feature half as a third
#define WD_TILE 16384                                   // elements of a [256 rows][64 K] tile (32 KB)
#define WD_C 0
#define WD_GK (WD_C + 8 * WD_TILE)
#define WD_GV (WD_GK + 4 * WD_TILE)
#define WD_M1 (WD_GV + 4 * WD_TILE)
#define WD_TO (WD_M1 + 4 * WD_TILE)                     // tour output_proj: per-tour-mean half (256 x 128) as 2 K-block tiles, then
                                                        // one tile whose K rows 0..4 are the folded tour-feature matrix
                                                        // (OFF_TOF_WT) and K row 5 the folded bias (OFF_TOF_B), rest 0
#define WD_PART (WD_TO + 3 * WD_TILE)                   // elements of one part (hi or lo)
#define W16_DEC (W16_NE + 20 * W16_STAGE)
#define W16_END (W16_DEC + 2 * WD_PART)

struct PtArgs {
  jampr_dims_t d;
  jampr_instances_t p;
  jampr_state_t st;
  const float* w;        // fp32 blob (biases, LayerNorm, folded tour-feature matrix)
  const void* w16;       // 16-bit blob
  const float* uniforms;
  uint64_t seed;
  int step_no;
  int decode_type;
  int graph_fresh;
  int store_cache;       // write h_fin / ne / ne_stats to HBM (needed when later steps reuse them)
  int split;             // 1 = stop after the query / tour embeddings and hand over to decoder_tc.cu (st.ne16, dq, dte)
  int ablate;            // timing experiments only (JAMPR_TC_ABLATE): 1 skip kNN aggregation, 2 skip layer epilogue
                         // math, 4 skip the decoder mat-vec chain, 8 skip the MMAs, 16 skip the whole decoder, 32 no wait for weight stages
};

template <typename T>
__device__ __forceinline__ T to_T(float v);
template <>
__device__ __forceinline__ __half to_T<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 to_T<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <typename T>
__device__ __forceinline__ float from_T(T v);
template <>
__device__ __forceinline__ float from_T<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float from_T<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ uint16_t bits_T(float v) {
  T t = to_T<T>(v);
  return *reinterpret_cast<uint16_t*>(&t);
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }

// GELU(erf) on two values with the packed fp32 pipe and ONE special-function op per value.
//   gelu(x) = x Phi(x) = relu(x) - |x| Phi(-|x|),   Phi(-u) = 2^q(u),  q = log2(erfc(u / sqrt 2) / 2)
// q is smooth (no pole, ~ -0.72 u^2 for large u), so a degree-5 polynomial fitted for minimax absolute error of
// Phi on [0, 6.5] (weights = Phi itself; tools/fit_gelu.py) reaches 2.9e-7 in Phi and 1.3e-6 in gelu — the error
// level of ex2.approx itself and 200x below the 16-bit operand rounding of this path.  The leading coefficient is
// negative: beyond the fitted range q keeps falling and 2^q underflows to 0 (gelu -> relu), no clamp needed.
// The previous Abramowitz-Stegun 7.1.26 form needed rcp + ex2 per value and made the layer epilogue bound by the
// special-function pipe (16 lanes per SM on sm_100a): 2 x 98k MUFU per rollout-step.
__device__ __forceinline__ float2 gelu2(float2 x) {
  // polynomial in nu = -|x|: odd coefficients carry the sign
  const float2 nu = make_float2(-fabsf(x.x), -fabsf(x.y));
  float2 q = __ffma2_rn(nu, f2(0.0005188681534491479f), f2(0.007387245539575815f));
  q = __ffma2_rn(nu, q, f2(0.05253808572888374f));
  q = __ffma2_rn(nu, q, f2(-0.45927664637565613f));
  q = __ffma2_rn(nu, q, f2(1.1510831117630005f));
  q = __ffma2_rn(nu, q, f2(-1.0000008344650269f));
  const float2 h = make_float2(ex2_approx(q.x), ex2_approx(q.y));      // Phi(-|x|)
  return __ffma2_rn(nu, h, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// ---- one K-block (64 of K) of a 128 x 128 tile product: 4 tcgen05.mma of K = 16
__device__ __forceinline__ void mma_kblock(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool first) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_f16(tmem_d, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, (first && k == 0) ? 0u : 1u);
}

// ---- max aggregation into the agg tiles (K-blocks 0,1), packed 16-bit math.
// An edge is one 32-bit table entry: low half = byte offset of the source row inside an h tile with the row's
// swizzle phase folded in (row * 128 | (row & 7) << 4), high half = the edge weight in the operand type.  The
// 16-byte chunk c of that row then sits at (entry & 0xffff) ^ (c << 4).  Rows are padded to a multiple of four
// entries by repeating the last edge (max is idempotent), so the table is read with 128-bit loads.
__device__ __forceinline__ uint32_t agg_entry(uint32_t row, uint16_t wbits) {
  return (row << 7) | ((row & 7u) << 4) | ((uint32_t)wbits << 16);
}

template <typename T>
__device__ __forceinline__ void agg_edge(uint32_t e, uint32_t hb, uint32_t c4, typename Elem<T>::T2 (&m)[4], bool first) {
  using T2 = typename Elem<T>::T2;
  const T2 w2 = from_u32<T2>(__byte_perm(e, e, 0x3232));
  const uint4 v = lds128(hb + ((e & 0xffffu) ^ c4));
  const T2 p0 = __hmul2(from_u32<T2>(v.x), w2), p1 = __hmul2(from_u32<T2>(v.y), w2);
  const T2 p2 = __hmul2(from_u32<T2>(v.z), w2), p3 = __hmul2(from_u32<T2>(v.w), w2);
  if (first) { m[0] = p0; m[1] = p1; m[2] = p2; m[3] = p3; }
  else { m[0] = __hmax2(m[0], p0); m[1] = __hmax2(m[1], p1); m[2] = __hmax2(m[2], p2); m[3] = __hmax2(m[3], p3); }
}

// max over entries [e0, e1) (both multiples of 4, e1 > e0) of weight * h[source] for one 16-byte chunk
template <typename T>
__device__ __forceinline__ uint4 agg_run(uint32_t tab_saddr, int e0, int e1, uint32_t hb, uint32_t c4) {
  using T2 = typename Elem<T>::T2;
  T2 m[4];
  for (int j = e0; j < e1; j += 4) {
    const uint4 q = lds128(tab_saddr + (uint32_t)j * 4u);
    agg_edge<T>(q.x, hb, c4, m, j == e0);
    agg_edge<T>(q.y, hb, c4, m, false);
    agg_edge<T>(q.z, hb, c4, m, false);
    agg_edge<T>(q.w, hb, c4, m, false);
  }
  return make_uint4(as_u32(m[0]), as_u32(m[1]), as_u32(m[2]), as_u32(m[3]));
}

// two adjacent 16-byte chunks (cA, cB = chunk index << 4) of the same rows per thread: the edge decode (offset,
// weight broadcast, table load) is paid once for 32 bytes of features
template <typename T>
__device__ __forceinline__ void agg_edge2(uint32_t e, uint32_t hb, uint32_t cA, uint32_t cB, typename Elem<T>::T2 (&m)[8],
                                          bool first) {
  using T2 = typename Elem<T>::T2;
  const T2 w2 = from_u32<T2>(__byte_perm(e, e, 0x3232));
  // cB == cA ^ 0x40 at every call site and hb is 1 KB aligned: the second address is one XOR away from the first
  (void)cB;
  const uint32_t aA = hb + ((e & 0xffffu) ^ cA);
  const uint4 va = lds128(aA), vb = lds128(aA ^ 0x40u);
  const T2 p0 = __hmul2(from_u32<T2>(va.x), w2), p1 = __hmul2(from_u32<T2>(va.y), w2);
  const T2 p2 = __hmul2(from_u32<T2>(va.z), w2), p3 = __hmul2(from_u32<T2>(va.w), w2);
  const T2 p4 = __hmul2(from_u32<T2>(vb.x), w2), p5 = __hmul2(from_u32<T2>(vb.y), w2);
  const T2 p6 = __hmul2(from_u32<T2>(vb.z), w2), p7 = __hmul2(from_u32<T2>(vb.w), w2);
  if (first) { m[0] = p0; m[1] = p1; m[2] = p2; m[3] = p3; m[4] = p4; m[5] = p5; m[6] = p6; m[7] = p7; }
  else {
    m[0] = __hmax2(m[0], p0); m[1] = __hmax2(m[1], p1); m[2] = __hmax2(m[2], p2); m[3] = __hmax2(m[3], p3);
    m[4] = __hmax2(m[4], p4); m[5] = __hmax2(m[5], p5); m[6] = __hmax2(m[6], p6); m[7] = __hmax2(m[7], p7);
  }
}

template <typename T>
__device__ __forceinline__ void agg_run2(uint32_t tab_saddr, int e0, int e1, uint32_t hb, uint32_t cA, uint32_t cB,
                                         uint4& ra, uint4& rb) {
  using T2 = typename Elem<T>::T2;
  T2 m[8];
  for (int j = e0; j < e1; j += 4) {
    const uint4 q = lds128(tab_saddr + (uint32_t)j * 4u);
    agg_edge2<T>(q.x, hb, cA, cB, m, j == e0);
    agg_edge2<T>(q.y, hb, cA, cB, m, false);
    agg_edge2<T>(q.z, hb, cA, cB, m, false);
    agg_edge2<T>(q.w, hb, cA, cB, m, false);
  }
  ra = make_uint4(as_u32(m[0]), as_u32(m[1]), as_u32(m[2]), as_u32(m[3]));
  rb = make_uint4(as_u32(m[4]), as_u32(m[5]), as_u32(m[6]), as_u32(m[7]));
}

// sum over the warp of four values at once: lane 16*(h>>1) + 8*(h&1) ends up with the total of value h
__device__ __forceinline__ float warp_sum4(float d0, float d1, float d2, float d3, int lane) {
  const bool up4 = (lane & 16) != 0, up3 = (lane & 8) != 0;
  float x0 = up4 ? d2 : d0, x1 = up4 ? d3 : d1;
  const float y0 = up4 ? d0 : d2, y1 = up4 ? d1 : d3;
  x0 += __shfl_xor_sync(0xffffffffu, y0, 16);
  x1 += __shfl_xor_sync(0xffffffffu, y1, 16);
  float z = up3 ? x1 : x0;
  const float sn = up3 ? x0 : x1;
  z += __shfl_xor_sync(0xffffffffu, sn, 8);
  z += __shfl_xor_sync(0xffffffffu, z, 4);
  z += __shfl_xor_sync(0xffffffffu, z, 2);
  z += __shfl_xor_sync(0xffffffffu, z, 1);
  return z;
}

