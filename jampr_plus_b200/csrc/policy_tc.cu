// Tensor-core policy step (sm_100a): the per-step tour-graph GNN (tour_encoder.py:159-220), the node-embedding
// projection (:217-218), the per-tour embeddings (:202-213, :226-243) and the pointer decoder with action
// selection (policy.py:170-227, attn_decoder.py:62-98) fused into ONE kernel, one CTA per rollout.
//
//   * The six GraphConv blocks are [agg | h] (128 x 256) x W_l (256 x 128) products on the 5th-gen tensor cores:
//     tcgen05.mma (M=128, N=128, K=16 per instruction, 16-bit operands, fp32 accumulation in TMEM) issued by one
//     thread; A = the CTA's own shared-memory tiles (K-major, 128-byte swizzle) that the aggregation phase and the
//     previous layer's epilogue write in place; B = pre-swizzled weight tile images streamed from L2 by the bulk
//     async-copy engine (cp.async.bulk + mbarrier complete_tx) into four 16 KB stages, one layer ahead.
//   * lin_root(h) (K-blocks 2,3) is issued before the max-aggregation runs, so that half of every layer's tensor
//     work overlaps the CUDA-core aggregation; lin_rel(agg) (K-blocks 0,1) follows it.
//   * Epilogue per layer (bias, exact-erf GELU to 1.5e-7, skip, LayerNorm) runs on the accumulator rows read back
//     with tcgen05.ld; the fp32 skip operand lives in TMEM too (columns 128..255, tcgen05.st), so the only
//     16-bit roundings on the path are the GEMM operands themselves.
//   * node_emb = h W_o^T + b + static_emb is a 128 x 256 x 128 product into TMEM columns 0..255; its fp32 rows are
//     parked in shared memory (over the dead operand tiles) where the decoder reads them: the action embeddings
//     are never materialised and nothing but the chosen action goes back to HBM (unless the embedding cache is
//     needed for tour_graph_update_step > 1 or a caller asks for it).
//   * Decoder algebra: as in decoder.cu the set_proj thirds are moved onto the query side; in addition
//     out_proj and the logit-key third are pre-multiplied on the host (M1 = W_lk^T W_out) and the tour-feature
//     branch of tour_encoder.output_proj is folded into a 5 x 256 matrix (packing.py).
#include "policy_tc_common.cuh"

#define PT_THREADS 512
#define NE_STRIDE 260                 // fp32 row stride of the parked node embedding (260 % 32 = 4: conflict-free rows)
#define OVERLAY_BYTES (128 * NE_STRIDE * 4)   // 133120 >= operand tiles (65536) + weight stages (65536)
#define QP_STRIDE 260

template <typename T>
__device__ __forceinline__ void load_w_stage(uint8_t* sW, int stage, const void* w16, int layer, uint64_t* full) {
  mbar_arrive_expect_tx(&full[stage], 16384u);
  bulk_g2s(sW + stage * 16384, reinterpret_cast<const uint8_t*>(w16) + ((size_t)(W16_ENC + (layer * 4 + stage) * W16_STAGE)) * 2,
           16384u, &full[stage]);
}

// Static neighbourhood graph (graph_utils.py:127-141).  Warp 0: the depot row (receives from every node, weight =
// time to depot), node range split over the two half-warps; warps 1..15: 30 customers at a time, 16 threads
// (16-byte chunks of the 128 features) per customer.
template <typename T>
__device__ __forceinline__ void agg_nbh_tc(uint8_t* sA, uint32_t tab_saddr, int k4, uint32_t dtab_saddr, int nd, int N) {
  using T2 = typename Elem<T>::T2;
  const int tid = threadIdx.x;
  const uint32_t h_saddr = smem_u32(sA) + 2u * 16384u;
  if (tid < 32) {
    const int cc = tid & 15, half = tid >> 4;
    const uint32_t kboff = (uint32_t)(cc >> 3) * 16384u, c = cc & 7;
    uint4 r = agg_run<T>(dtab_saddr, half * (nd >> 1), (half + 1) * (nd >> 1), h_saddr + kboff, c << 4);
    r.x = as_u32(__hmax2(from_u32<T2>(r.x), from_u32<T2>(__shfl_xor_sync(0xffffffffu, r.x, 16))));
    r.y = as_u32(__hmax2(from_u32<T2>(r.y), from_u32<T2>(__shfl_xor_sync(0xffffffffu, r.y, 16))));
    r.z = as_u32(__hmax2(from_u32<T2>(r.z), from_u32<T2>(__shfl_xor_sync(0xffffffffu, r.z, 16))));
    r.w = as_u32(__hmax2(from_u32<T2>(r.w), from_u32<T2>(__shfl_xor_sync(0xffffffffu, r.w, 16))));
    if (half == 0) *reinterpret_cast<uint4*>(sA + kboff + sw128_off(0, c)) = r;
  } else {
    const int g = (tid - 32) >> 4, cc = (tid - 32) & 15;
    const uint32_t kboff = (uint32_t)(cc >> 3) * 16384u, c = cc & 7;
    for (int u = 1 + g; u < N; u += 30) {
      const uint4 r = agg_run<T>(tab_saddr + (uint32_t)(u * k4) * 4u, 0, k4, h_saddr + kboff, c << 4);
      *reinterpret_cast<uint4*>(sA + kboff + sw128_off(u, c)) = r;
    }
  }
}

// Tour graph (env.py:804-835): customer u receives from its predecessor (successor in the reversed graph) when it
// is on a tour — ltab[u] is that single edge, or the zero-weight edge 0 otherwise; the depot receives from the
// tour ends (etab, n_e4 entries padded to a multiple of 4, 0 = no tour yet -> zeros).
template <typename T>
__device__ __forceinline__ void agg_tour_tc(uint8_t* sA, const uint32_t* s_ltab, uint32_t etab_saddr, int n_e4, int N) {
  using T2 = typename Elem<T>::T2;
  const int tid = threadIdx.x;
  const uint32_t h_saddr = smem_u32(sA) + 2u * 16384u;
  if (tid < 32) {
    if (tid < 16) {
      const uint32_t kboff = (uint32_t)(tid >> 3) * 16384u, c = tid & 7;
      uint4 r = make_uint4(0u, 0u, 0u, 0u);
      if (n_e4 > 0) r = agg_run<T>(etab_saddr, 0, n_e4, h_saddr + kboff, c << 4);
      *reinterpret_cast<uint4*>(sA + kboff + sw128_off(0, c)) = r;
    }
  } else {
    const int g = (tid - 32) >> 4, cc = (tid - 32) & 15;
    const uint32_t kboff = (uint32_t)(cc >> 3) * 16384u, c = cc & 7;
    for (int u = 1 + g; u < N; u += 30) {
      T2 m[4];
      agg_edge<T>(s_ltab[u], h_saddr + kboff, c << 4, m, true);
      *reinterpret_cast<uint4*>(sA + kboff + sw128_off(u, c)) = make_uint4(as_u32(m[0]), as_u32(m[1]), as_u32(m[2]), as_u32(m[3]));
    }
  }
}

// ---- layer epilogue: y = GELU(acc + b) + skip -> LayerNorm -> new h (fp32 into the TMEM skip columns, 16-bit into
// the h tiles, optionally fp32 to HBM).  Thread (q, lane, cg): row 32q+lane, columns 32cg..32cg+31.  par_saddr:
// shared address of this layer's [bias 128 | gamma 128 | beta 128]; all math on the packed fp32 pipe.
template <typename T>
__device__ __forceinline__ void layer_epilogue(uint32_t tmem, uint8_t* sA, float2* s_red, uint32_t par_saddr, int N,
                                               float* __restrict__ hfin) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q = warp & 3, cg = warp >> 2, r = q * 32 + lane;
  const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16);
  uint32_t acc[32], sk[32];
  tmem_ld32_nowait(ta + cg * 32, acc);
  tmem_ld32_nowait(ta + 128 + cg * 32, sk);
  tmem_wait_ld();
  float2 s2 = make_float2(0.0f, 0.0f), q2 = make_float2(0.0f, 0.0f);
  const uint32_t pb = par_saddr + (uint32_t)cg * 128u;
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
  s_red[cg * 128 + r] = make_float2(s2.x + s2.y, q2.x + q2.y);
  __syncthreads();
  float ts = 0.0f, tq = 0.0f;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float2 v = s_red[g * 128 + r];
    ts += v.x;
    tq += v.y;
  }
  const float mean = ts * (1.0f / 128.0f);
  const float var = fmaxf(fmaf(-mean, mean, tq * (1.0f / 128.0f)), 0.0f);
  const float2 rstd2 = f2(rsqrtf(var + 1e-5f)), nmean2 = f2(-mean);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 g4 = lds_f32x4(pb + 512 + i * 16), e4 = lds_f32x4(pb + 1024 + i * 16);
    const float2 d0 = __fadd2_rn(make_float2(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1])), nmean2);
    const float2 d1 = __fadd2_rn(make_float2(__uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3])), nmean2);
    const float2 o0 = __ffma2_rn(d0, __fmul2_rn(make_float2(g4.x, g4.y), rstd2), make_float2(e4.x, e4.y));
    const float2 o1 = __ffma2_rn(d1, __fmul2_rn(make_float2(g4.z, g4.w), rstd2), make_float2(e4.z, e4.w));
    acc[4 * i] = __float_as_uint(o0.x); acc[4 * i + 1] = __float_as_uint(o0.y);
    acc[4 * i + 2] = __float_as_uint(o1.x); acc[4 * i + 3] = __float_as_uint(o1.y);
  }
  tmem_st32(ta + 128 + cg * 32, acc);
  if (r < N) {
    uint8_t* tile = sA + (2 + (cg >> 1)) * 16384;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        pk[e] = as_u32(Elem<T>::from2(__uint_as_float(acc[i * 8 + 2 * e]), __uint_as_float(acc[i * 8 + 2 * e + 1])));
      *reinterpret_cast<uint4*>(tile + sw128_off(r, (cg & 1) * 4 + i)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    if (hfin) {
      float4* dst = reinterpret_cast<float4*>(hfin + (size_t)r * 128 + cg * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        dst[i] = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                             __uint_as_float(acc[4 * i + 3]));
    }
  }
  tmem_wait_st();
}

// ---- decoder mat-vec partials: y[o] = sum_c x[c] W[c][o] over 256 outputs, W 16-bit [K][256].  A thread owns OPT
// adjacent outputs and one of 512*OPT/256 K-slices; partial sums go to s_part[(slice * NV + v) * 256 + o].
// x_saddr: shared address of this thread's input vector(s), vector v at + v * xstride floats.
template <typename T, int OPT, int NV>
__device__ __forceinline__ void matvec_partial(const T* __restrict__ W, int K, uint32_t x_saddr, int xstride, float* s_part) {
  using T2 = typename Elem<T>::T2;
  constexpr int NG = 256 / OPT, NS = PT_THREADS / NG;
  const int og = threadIdx.x % NG, ks = threadIdx.x / NG;
  const int kn = K / NS, k0 = ks * kn;
  float acc[NV][OPT];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < OPT; ++j) acc[v][j] = 0.0f;
  const T* wp = W + (size_t)k0 * 256 + og * OPT;
  const uint32_t xa = x_saddr + (uint32_t)k0 * 4u;
#pragma unroll 8
  for (int c = 0; c < kn; ++c) {
    float wv[OPT];
    if constexpr (OPT == 8) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(wp + (size_t)c * 256));
      const float2 f0 = Elem<T>::to2(from_u32<T2>(raw.x)), f1 = Elem<T>::to2(from_u32<T2>(raw.y));
      const float2 f2_ = Elem<T>::to2(from_u32<T2>(raw.z)), f3 = Elem<T>::to2(from_u32<T2>(raw.w));
      wv[0] = f0.x; wv[1] = f0.y; wv[2] = f1.x; wv[3] = f1.y; wv[4] = f2_.x; wv[5] = f2_.y; wv[6] = f3.x; wv[7] = f3.y;
    } else {
      static_assert(OPT == 2, "OPT must be 8 or 2");
      const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(wp + (size_t)c * 256));
      const float2 f0 = Elem<T>::to2(from_u32<T2>(raw));
      wv[0] = f0.x; wv[1] = f0.y;
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float xv = lds_f32(xa + (uint32_t)(v * xstride + c) * 4u);
#pragma unroll
      for (int j = 0; j < OPT; j += 2) {
        const float2 a2 = __ffma2_rn(make_float2(wv[j], wv[j + 1]), f2(xv), make_float2(acc[v][j], acc[v][j + 1]));
        acc[v][j] = a2.x;
        acc[v][j + 1] = a2.y;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < OPT; ++j) s_part[(ks * NV + v) * 256 + og * OPT + j] = acc[v][j];
}

// fp32-weight variant (8 outputs per thread, one input vector) for ctxt_proj and the glimpse keys: their 16-bit rounding
// dominated the log-probability error of the path (see policy_tc2.cu mvf_prefetch)
__device__ __forceinline__ void matvec_partial_f32(const float* __restrict__ W, int K, uint32_t x_saddr, float* s_part) {
  constexpr int NG = 32, NS = PT_THREADS / NG;
  const int og = threadIdx.x % NG, ks = threadIdx.x / NG;
  const int kn = K / NS, k0 = ks * kn;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  const float* wp = W + (size_t)k0 * 256 + og * 8;
  const uint32_t xa = x_saddr + (uint32_t)k0 * 4u;
#pragma unroll 8
  for (int c = 0; c < kn; ++c) {
    const float4 wa = __ldg(reinterpret_cast<const float4*>(wp + (size_t)c * 256));
    const float4 wb = __ldg(reinterpret_cast<const float4*>(wp + (size_t)c * 256) + 1);
    const float2 xv = f2(lds_f32(xa + (uint32_t)c * 4u));
    const float2 a0 = __ffma2_rn(make_float2(wa.x, wa.y), xv, make_float2(acc[0], acc[1]));
    const float2 a1 = __ffma2_rn(make_float2(wa.z, wa.w), xv, make_float2(acc[2], acc[3]));
    const float2 a2 = __ffma2_rn(make_float2(wb.x, wb.y), xv, make_float2(acc[4], acc[5]));
    const float2 a3 = __ffma2_rn(make_float2(wb.z, wb.w), xv, make_float2(acc[6], acc[7]));
    acc[0] = a0.x; acc[1] = a0.y; acc[2] = a1.x; acc[3] = a1.y; acc[4] = a2.x; acc[5] = a2.y; acc[6] = a3.x; acc[7] = a3.y;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s_part[ks * 256 + og * 8 + j] = acc[j];
}

template <typename T>
__global__ void __launch_bounds__(PT_THREADS, 1) policy_step_tc_kernel(const PtArgs a) {
  using T2 = typename Elem<T>::T2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const jampr_dims_t& d = a.d;
  const jampr_state_t& st = a.st;
  const jampr_instances_t& p = a.p;
  if (!jampr_step_active(st.ctl, a.step_no)) return;
  const int N = d.n_nodes, K = d.n_slots, k = d.k_nbh, A = K * k, L = d.plan_len;
  const int k4 = (k + 3) & ~3, nd = (N + 7) & ~7;
  const int b = blockIdx.x, ins = b / d.n_samples;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* __restrict__ w = a.w;
  const T* __restrict__ w16 = reinterpret_cast<const T*>(a.w16);
  const bool fresh = a.graph_fresh != 0;

  // ---------------------------------------------------------------- shared memory carve-up (32-bit offsets)
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;               // 4 operand tiles: agg (K-blocks 0,1), h (K-blocks 2,3)
  uint8_t* sW = base + 65536;       // 4 weight stages
  float* s_ne = reinterpret_cast<float*>(base);   // node embedding rows, parked over sA/sW once the GNN is done
  uint32_t off = OVERLAY_BYTES;
  auto take = [&](uint32_t bytes) {
    uint8_t* r = base + off;
    off += (bytes + 15u) & ~15u;
    return r;
  };
  uint64_t* bars = reinterpret_cast<uint64_t*>(take(64));     // full[0..3], root, acc
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(take(16));
  float2* s_red = reinterpret_cast<float2*>(take(4 * 128 * sizeof(float2)));
  float* s_part = reinterpret_cast<float*>(take(16 * 256 * 4));
  float* s_par = reinterpret_cast<float*>(take(6 * 384 * 4));          // per layer: bias | gamma | beta
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(take((uint32_t)N * k4 * 4));
  uint32_t* s_dtab = reinterpret_cast<uint32_t*>(take((uint32_t)nd * 4));
  uint32_t* s_lf = reinterpret_cast<uint32_t*>(take((uint32_t)N * 4));
  uint32_t* s_lr = reinterpret_cast<uint32_t*>(take((uint32_t)N * 4));
  uint32_t* s_ef = reinterpret_cast<uint32_t*>(take((uint32_t)(N + 4) * 4));
  uint32_t* s_er = reinterpret_cast<uint32_t*>(take((uint32_t)(N + 4) * 4));
  int* s_cnt = reinterpret_cast<int*>(take(16));
  int16_t* s_plan = reinterpret_cast<int16_t*>(take((uint32_t)K * L * 2));
  float* s_mean = reinterpret_cast<float*>(take(4 * 128 * 4));
  float* s_te = reinterpret_cast<float*>(take((uint32_t)K * 256 * 4));
  float* s_q = reinterpret_cast<float*>(take(512 * 4));
  float* s_ctxt = reinterpret_cast<float*>(take(256 * 4));
  float* s_qp = reinterpret_cast<float*>(take(4 * QP_STRIDE * 4));
  float* s_gbar = reinterpret_cast<float*>(take(4 * QP_STRIDE * 4));
  float* s_g = reinterpret_cast<float*>(take(256 * 4));
  float* s_lp = reinterpret_cast<float*>(take(256 * 4));
  float4* s_sn4 = reinterpret_cast<float4*>(take(128 * 16));           // per node: glimpse score of the 4 heads
  float* s_ln = reinterpret_cast<float*>(take(128 * 4));
  float4* s_pn4 = reinterpret_cast<float4*>(take(128 * 16));           // per node: glimpse weight of the 4 heads
  float* s_qt = reinterpret_cast<float*>(take(4 * JAMPR_MAX_SLOTS * 4));
  float* s_lt = reinterpret_cast<float*>(take(JAMPR_MAX_SLOTS * 4));
  float* s_P = reinterpret_cast<float*>(take(4 * JAMPR_MAX_SLOTS * 4));
  float* s_sc = reinterpret_cast<float*>(take((uint32_t)4 * A * 4));
  float* s_logit = reinterpret_cast<float*>(take((uint32_t)A * 4));
  float* s_feat = reinterpret_cast<float*>(take((uint32_t)K * 8 * 4));
  int* s_nbh = reinterpret_cast<int*>(take((uint32_t)A * 4));
  int8_t* s_pos = reinterpret_cast<int8_t*>(take((uint32_t)K * N));    // slot t, node n -> candidate index j or -1
  uint8_t* s_mask = take(A);
  int* s_idx = reinterpret_cast<int*>(take(JAMPR_MAX_SLOTS * 4));
  uint64_t* full = bars;
  uint64_t* bar_root = bars + 4;
  uint64_t* bar_acc = bars + 5;

  // ---------------------------------------------------------------- prologue
  if (fresh) {
    if (tid == 0) {
      for (int i = 0; i < 4; ++i) mbar_init(&full[i], 1);
      mbar_init(bar_root, 1);
      mbar_init(bar_acc, 1);
      mbar_fence_init();
      for (int s = 0; s < 4; ++s) load_w_stage<T>(sW, s, a.w16, 0, full);
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    // zero the operand tiles: rows >= N stay zero for the whole kernel
    for (int e = tid; e < 65536 / 16; e += PT_THREADS) reinterpret_cast<uint4*>(sA)[e] = make_uint4(0u, 0u, 0u, 0u);
    for (int e = tid; e < 6 * 96; e += PT_THREADS) {
      const int l = e / 96, i = e - l * 96;
      reinterpret_cast<float4*>(s_par)[e] = __ldg(reinterpret_cast<const float4*>(w + OFF_TE_BLK + (size_t)l * BLK_FLOATS + BLK_B) + i);
    }
  }
  if (tid < 4) s_cnt[tid] = 0;
  for (int e = tid; e < K * N; e += PT_THREADS) s_pos[e] = (int8_t)-1;
  // tour features (env.py:1061-1068), tour buffer rows of the active vehicles and their lengths
  if (tid < K) {
    const size_t bk = (size_t)b * K + tid;
    const int row = st.row[bk];
    s_feat[tid * 8 + 0] = (float)st.cur[bk];
    s_feat[tid * 8 + 1] = st.cap[bk];
    s_feat[tid * 8 + 2] = st.time[bk];
    s_feat[tid * 8 + 3] = st.cttd[bk];
    s_feat[tid * 8 + 4] = (float)(d.k_max - row) / (float)d.k_max;
  }
  for (int e = tid; e < K * L; e += PT_THREADS) {
    const int t = e / L, q = e - t * L;
    s_plan[e] = st.plan[((size_t)b * d.k_max + st.row[(size_t)b * K + t]) * L + q];
  }
  for (int e = tid; e < A; e += PT_THREADS) {
    s_nbh[e] = st.nbh[(size_t)b * A + e];
    s_mask[e] = st.mask[(size_t)b * A + e];
  }
  __syncthreads();
  for (int e = tid; e < A; e += PT_THREADS) s_pos[(e / k) * N + s_nbh[e]] = (int8_t)(e % k);
  if (fresh) {
    {
      const uint4* gt = reinterpret_cast<const uint4*>(p.tc_tab + (size_t)ins * N * k4);
      for (int e = tid; e < (N * k4) / 4; e += PT_THREADS) reinterpret_cast<uint4*>(s_tab)[e] = __ldg(gt + e);
      const uint4* gd = reinterpret_cast<const uint4*>(p.tc_dtab + (size_t)ins * nd);
      for (int e = tid; e < nd / 4; e += PT_THREADS) reinterpret_cast<uint4*>(s_dtab)[e] = __ldg(gd + e);
    }
    const float* dist = p.dist + (size_t)ins * N * N;
    for (int u = tid; u < N; u += PT_THREADS) {
      const int pr = st.pred[(size_t)b * N + u], su = st.succ[(size_t)b * N + u];
      const bool vs = u > 0 && st.visited[(size_t)b * N + u] != 0;
      s_lf[u] = vs ? agg_entry((uint32_t)pr, bits_T<T>(dist[(size_t)u * N + pr])) : 0u;
      s_lr[u] = vs ? agg_entry((uint32_t)su, bits_T<T>(dist[(size_t)u * N + su])) : 0u;
      if (vs) {
        const uint32_t en = agg_entry((uint32_t)u, bits_T<T>(p.ttd[(size_t)ins * N + u]));
        if (su == 0) s_ef[atomicAdd(&s_cnt[0], 1)] = en;
        if (pr == 0) s_er[atomicAdd(&s_cnt[1], 1)] = en;
      }
    }
  }
  if (tid < K) {
    int idx = 0;
    while (idx < L && s_plan[tid * L + idx] != 0) ++idx;
    s_idx[tid] = (idx >= L) ? 0 : idx;        // argmin over an all-nonzero row is position 0 (tour_encoder.py:236)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int q4 = warp & 3, cg = warp >> 2, r = q4 * 32 + lane;      // epilogue role: row r, columns 32cg..32cg+31

  if (fresh) {
    const uint32_t tmem = *tmem_slot;
    const uint32_t ta = tmem + ((uint32_t)(q4 * 32) << 16);
    const uint32_t idesc = umma_idesc(128, 128, Elem<T>::fmt);
    const uint32_t a_addr = smem_u32(sA), w_addr = smem_u32(sW);
    // pad the tour-end tables to a multiple of four entries (repeat the last one)
    const int n_endf = s_cnt[0], n_endr = s_cnt[1];
    const int n_ef4 = (n_endf + 3) & ~3, n_er4 = (n_endr + 3) & ~3;
    if (tid < 3 && n_endf > 0 && n_endf + tid < n_ef4) s_ef[n_endf + tid] = s_ef[n_endf - 1];
    if (tid >= 32 && tid < 35 && n_endr > 0 && n_endr + (tid - 32) < n_er4) s_er[n_endr + tid - 32] = s_er[n_endr - 1];
    // ---- h_0 = node_emb_input_proj(static_emb) (hoisted, per instance): fp32 -> TMEM skip columns, 16-bit -> h tiles
    {
      uint32_t hv[32];
      if (r < N) {
        // column-chunk-major copy (include/jampr_b200.h h0_t): contiguous across the warp's nodes
        const float4* src = reinterpret_cast<const float4*>(p.h0_t) + ((size_t)ins * 32 + cg * 8) * N + r;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = __ldg(src + (size_t)i * N);
          hv[4 * i] = __float_as_uint(v.x); hv[4 * i + 1] = __float_as_uint(v.y);
          hv[4 * i + 2] = __float_as_uint(v.z); hv[4 * i + 3] = __float_as_uint(v.w);
        }
        uint8_t* tile = sA + (2 + (cg >> 1)) * 16384;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            pk[e] = as_u32(Elem<T>::from2(__uint_as_float(hv[i * 8 + 2 * e]), __uint_as_float(hv[i * 8 + 2 * e + 1])));
          *reinterpret_cast<uint4*>(tile + sw128_off(r, (cg & 1) * 4 + i)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) hv[i] = 0u;
      }
      tmem_st32(ta + 128 + cg * 32, hv);
      tmem_wait_st();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const uint32_t tab_saddr = smem_u32(s_tab), dtab_saddr = smem_u32(s_dtab), par_saddr = smem_u32(s_par);
    const uint32_t ef_saddr = smem_u32(s_ef), er_saddr = smem_u32(s_er);
    for (int l = 0; l < 6; ++l) {
      const uint32_t ph = (uint32_t)(l & 1);
      if (tid == 0) {
        // lin_root(h): K-blocks 2,3 — overlaps the aggregation below
        mbar_wait(&full[2], ph);
        mbar_wait(&full[3], ph);
        tc_fence_after();
        mma_kblock(tmem, a_addr + 2 * 16384, w_addr + 2 * 16384, idesc, true);
        mma_kblock(tmem, a_addr + 3 * 16384, w_addr + 3 * 16384, idesc, false);
        umma_commit(bar_root);
      }
      if (l == 2 || l == 5) agg_nbh_tc<T>(sA, tab_saddr, k4, dtab_saddr, nd, N);
      else if (l < 2) agg_tour_tc<T>(sA, s_lf, ef_saddr, n_ef4, N);
      else agg_tour_tc<T>(sA, s_lr, er_saddr, n_er4, N);
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        mbar_wait(bar_root, ph);                       // stages 2,3 are free: stream the next layer into them
        load_w_stage<T>(sW, 2, a.w16, l + 1, full);
        load_w_stage<T>(sW, 3, a.w16, l + 1, full);
        mbar_wait(&full[0], ph);
        mbar_wait(&full[1], ph);
        tc_fence_after();
        mma_kblock(tmem, a_addr, w_addr, idesc, false);
        mma_kblock(tmem, a_addr + 16384, w_addr + 16384, idesc, false);
        umma_commit(bar_acc);
      }
      mbar_wait(bar_acc, ph);
      tc_fence_after();
      if (tid == 0) {
        load_w_stage<T>(sW, 0, a.w16, l + 1, full);
        load_w_stage<T>(sW, 1, a.w16, l + 1, full);
      }
      layer_epilogue<T>(tmem, sA, s_red, par_saddr + (uint32_t)l * 1536u, N,
                        (l == 5 && a.store_cache) ? st.h_fin + (size_t)b * N * 128 : nullptr);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
    // ---- node_emb_output_proj: D[:, 0:128] = h W_o[0:128]^T (stages 0,1), D[:, 128:256] = h W_o[128:256]^T (stages 2,3)
    if (tid == 0) {
      for (int s = 0; s < 4; ++s) mbar_wait(&full[s], 0u);
      tc_fence_after();
      mma_kblock(tmem, a_addr + 2 * 16384, w_addr, idesc, true);
      mma_kblock(tmem, a_addr + 3 * 16384, w_addr + 16384, idesc, false);
      mma_kblock(tmem + 128, a_addr + 2 * 16384, w_addr + 2 * 16384, idesc, true);
      mma_kblock(tmem + 128, a_addr + 3 * 16384, w_addr + 3 * 16384, idesc, false);
      umma_commit(bar_acc);
    }
  } else {
    // cached tour-graph embedding (tour_encoder.py:163-166): h_fin -> h tiles (for the per-tour means below)
    for (int e = tid; e < N * 16; e += PT_THREADS) {
      const int u = e >> 4, cc = e & 15;
      const float4* src = reinterpret_cast<const float4*>(st.h_fin + ((size_t)b * N + u) * 128 + cc * 8);
      const float4 v0 = src[0], v1 = src[1];
      *reinterpret_cast<uint4*>(sA + (2 + (cc >> 3)) * 16384 + sw128_off(u, cc & 7)) =
          make_uint4(as_u32(Elem<T>::from2(v0.x, v0.y)), as_u32(Elem<T>::from2(v0.z, v0.w)),
                     as_u32(Elem<T>::from2(v1.x, v1.y)), as_u32(Elem<T>::from2(v1.z, v1.w)));
    }
    __syncthreads();
  }

  // ---------------------------------------------------------------- per-tour mean of h over the buffer positions
  // 0..idx (includes one depot entry, tour_encoder.py:236-243); overlaps the projection MMAs
  for (int e = tid; e < K * 64; e += PT_THREADS) {
    const int t = e >> 6, c2 = e & 63;              // feature pair 2*c2, 2*c2+1
    const uint32_t toff = (uint32_t)(2 + (c2 >> 5)) * 16384u;
    const uint32_t ch = (uint32_t)(c2 & 31) >> 2, within = (uint32_t)(c2 & 3) * 4u;
    const int idx = s_idx[t];
    float sx = 0.0f, sy = 0.0f;
    for (int q = 0; q <= idx; ++q) {
      const uint32_t u = (uint32_t)s_plan[t * L + q];
      const float2 v = Elem<T>::to2(*reinterpret_cast<const T2*>(sA + toff + sw128_off(u, ch) + within));
      sx += v.x;
      sy += v.y;
    }
    const float inv = 1.0f / (float)(idx + 1);
    s_mean[t * 128 + 2 * c2] = sx * inv;
    s_mean[t * 128 + 2 * c2 + 1] = sy * inv;
  }
  __syncthreads();

  // ---------------------------------------------------------------- node embedding rows -> shared memory (fp32)
  if (fresh) {
    const uint32_t tmem = *tmem_slot;
    const uint32_t ta = tmem + ((uint32_t)(q4 * 32) << 16);
    mbar_wait(bar_acc, 0u);
    tc_fence_after();
    uint32_t acc0[32], acc1[32];
    tmem_ld32_nowait(ta + cg * 32, acc0);
    tmem_ld32_nowait(ta + 128 + cg * 32, acc1);
    tmem_wait_ld();
    if (r < N) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int c0 = hf * 128 + cg * 32;
        const float4* se = reinterpret_cast<const float4*>(p.static_emb_t) + ((size_t)ins * 64 + (c0 >> 2)) * N + r;   // [chunk][node]
        const float4* bo = reinterpret_cast<const float4*>(w + OFF_TE_OUT_B + c0);
        float4* dst = reinterpret_cast<float4*>(s_ne + (size_t)r * NE_STRIDE + c0);
        float4* gdst = a.store_cache ? reinterpret_cast<float4*>(st.ne + ((size_t)b * N + r) * 256 + c0) : nullptr;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 s4 = __ldg(se + (size_t)i * N), b4 = __ldg(bo + i);
          const uint32_t* ac = hf ? acc1 : acc0;
          const float4 v = make_float4(__uint_as_float(ac[4 * i]) + b4.x + s4.x, __uint_as_float(ac[4 * i + 1]) + b4.y + s4.y,
                                       __uint_as_float(ac[4 * i + 2]) + b4.z + s4.z, __uint_as_float(ac[4 * i + 3]) + b4.w + s4.w);
          dst[i] = v;
          if (gdst) gdst[i] = v;
        }
      }
    }
    tc_fence_before();
  } else {
    for (int e = tid; e < N * 64; e += PT_THREADS) {
      const int u = e >> 6, c4 = (e & 63) * 4;
      *reinterpret_cast<float4*>(s_ne + (size_t)u * NE_STRIDE + c4) =
          *reinterpret_cast<const float4*>(st.ne + ((size_t)b * N + u) * 256 + c4);
    }
  }
  __syncthreads();
  if (fresh && warp == 0) {
    tc_fence_after();
    tmem_dealloc(*tmem_slot, 256);
  }

  // ---------------------------------------------------------------- decoder
  const uint32_t ne_saddr = smem_u32(s_ne), te_saddr = smem_u32(s_te);
  // (b) tour_emb[t] = folded tour-feature branch + W_to2 mean_h[t]          (tour_encoder.py:202-213)
  matvec_partial<T, 2, 4>(w16 + W16_TO2, 128, smem_u32(s_mean), 128, s_part);   // NV = 4 >= K slots (K <= 4 here)
  // (a) column mean / max of node_emb over the nodes (policy.py:179-181) — independent of (b)
  float nmean = 0.0f, nmax = -INFINITY;
  if (tid < 256) {
    float s0 = 0.0f, s1 = 0.0f, m0 = -INFINITY, m1 = -INFINITY;
    int u = 0;
    for (; u + 1 < N; u += 2) {
      const float v0 = lds_f32(ne_saddr + (uint32_t)(u * NE_STRIDE + tid) * 4u);
      const float v1 = lds_f32(ne_saddr + (uint32_t)((u + 1) * NE_STRIDE + tid) * 4u);
      s0 += v0; s1 += v1;
      m0 = fmaxf(m0, v0); m1 = fmaxf(m1, v1);
    }
    if (u < N) {
      const float v0 = lds_f32(ne_saddr + (uint32_t)(u * NE_STRIDE + tid) * 4u);
      s0 += v0;
      m0 = fmaxf(m0, v0);
    }
    nmean = (s0 + s1) / (float)N;
    nmax = fmaxf(m0, m1);
    if (a.store_cache && fresh) {
      st.ne_stats[(size_t)b * 512 + tid] = nmean;
      st.ne_stats[(size_t)b * 512 + 256 + tid] = nmax;
    }
  }
  __syncthreads();
  if (tid < 256) {
    float sm = 0.0f, mx = -INFINITY;
    for (int t = 0; t < K; ++t) {
      float v = __ldg(w + OFF_TOF_B + tid);
#pragma unroll
      for (int i = 0; i < 5; ++i) v = fmaf(s_feat[t * 8 + i], __ldg(w + OFF_TOF_WT + i * 256 + tid), v);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) v += s_part[(ks * 4 + t) * 256 + tid];
      s_te[t * 256 + tid] = v;
      sm += v;
      mx = fmaxf(mx, v);
    }
    s_q[tid] = nmean + sm / (float)K;
    s_q[256 + tid] = nmax + mx;
  }
  __syncthreads();
  // (c) ctxt = ctxt_proj(query)
  matvec_partial_f32(w + OFF_CTXT_WT, 512, smem_u32(s_q), s_part);
  __syncthreads();
  if (tid < 256) {
    float v = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) v += s_part[ks * 256 + tid];
    s_ctxt[tid] = v;
  }
  __syncthreads();
  // (d) qp[h] = W_gk,h^T ctxt_h : K-slice ks (16 rows of set_proj) belongs to head ks / 4
  matvec_partial_f32(w + OFF_GK_W, 256, smem_u32(s_ctxt), s_part);
  __syncthreads();
  for (int e = tid; e < 1024; e += PT_THREADS) {
    const int h = e >> 8, c = e & 255;
    s_qp[h * QP_STRIDE + c] = s_part[(4 * h) * 256 + c] + s_part[(4 * h + 1) * 256 + c] + s_part[(4 * h + 2) * 256 + c] +
                              s_part[(4 * h + 3) * 256 + c];
  }
  __syncthreads();
  // (e) per-node and per-tour glimpse scores qp_h . node_emb[n], qp_h . tour_emb[t]: one warp per row, the lane's
  // eight columns of the four query vectors stay in registers
  {
    float qv[4][8];
#pragma unroll
    for (int h = 0; h < 4; ++h)
#pragma unroll
      for (int i = 0; i < 8; ++i) qv[h][i] = s_qp[h * QP_STRIDE + lane + 32 * i];
    for (int n = warp; n < N + K; n += PT_THREADS / 32) {
      const uint32_t ra = ((n < N) ? (ne_saddr + (uint32_t)n * NE_STRIDE * 4u) : (te_saddr + (uint32_t)(n - N) * 1024u)) + lane * 4u;
      float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = lds_f32(ra + i * 128u);
        d0 = fmaf(v, qv[0][i], d0);
        d1 = fmaf(v, qv[1][i], d1);
        d2 = fmaf(v, qv[2][i], d2);
        d3 = fmaf(v, qv[3][i], d3);
      }
      const float z = warp_sum4(d0, d1, d2, d3, lane);
      if ((lane & 7) == 0) {
        const int h = ((lane >> 4) << 1) | ((lane >> 3) & 1);
        if (n < N) reinterpret_cast<float*>(s_sn4)[n * 4 + h] = z;
        else s_qt[h * 8 + (n - N)] = z;
      }
    }
  }
  __syncthreads();
  // (f) glimpse scores per action, softmax per head
  for (int e = tid; e < 4 * A; e += PT_THREADS) {
    const int h = e / A, aa = e - h * A;
    const float sc = (reinterpret_cast<const float*>(s_sn4)[s_nbh[aa] * 4 + h] + s_qt[h * 8 + aa / k]) * 0.125f;
    s_sc[e] = s_mask[aa] ? -INFINITY : sc;
  }
  __syncthreads();
  if (warp < JAMPR_HEADS) {
    float* sc = s_sc + warp * A;
    float mx = -INFINITY;
    for (int e = lane; e < A; e += 32) mx = fmaxf(mx, sc[e]);
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int e = lane; e < A; e += 32) {
      const float ex = expf(sc[e] - mx);
      sc[e] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    for (int e = lane; e < A; e += 32) sc[e] = sc[e] / sum;
  }
  __syncthreads();
  // glimpse weights folded onto nodes and tours (fixed summation order over the slots: deterministic)
  for (int e = tid; e < 4 * N; e += PT_THREADS) {
    const int n = e >> 2, h = e & 3;
    float acc = 0.0f;
    for (int t = 0; t < K; ++t) {
      const int j = s_pos[t * N + n];
      if (j >= 0) acc += s_sc[h * A + t * k + j];
    }
    reinterpret_cast<float*>(s_pn4)[e] = acc;
  }
  if (tid >= 480 && tid < 480 + 4 * K) {
    const int e = tid - 480, h = e / K, t = e - h * K;
    float acc = 0.0f;
    for (int j = 0; j < k; ++j) acc += s_sc[h * A + t * k + j];
    s_P[h * 8 + t] = acc;
  }
  __syncthreads();
  // (g) gbar[h] = sum_a p[h][a] act[a] = sum_t P[h][t] tour_emb[t] + sum_n pn[h][n] node_emb[n]
  {
    const int c = tid & 255, half = tid >> 8;
    const uint32_t pn_saddr = smem_u32(s_pn4);
    float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, g3 = 0.0f;
    for (int n = half; n < N; n += 2) {
      const float v = lds_f32(ne_saddr + (uint32_t)(n * NE_STRIDE + c) * 4u);
      const float4 pw = lds_f32x4(pn_saddr + (uint32_t)n * 16u);
      g0 = fmaf(pw.x, v, g0);
      g1 = fmaf(pw.y, v, g1);
      g2 = fmaf(pw.z, v, g2);
      g3 = fmaf(pw.w, v, g3);
    }
    if (half == 0) {
      for (int t = 0; t < K; ++t) {
        const float v = s_te[t * 256 + c];
        g0 = fmaf(s_P[t], v, g0);
        g1 = fmaf(s_P[8 + t], v, g1);
        g2 = fmaf(s_P[16 + t], v, g2);
        g3 = fmaf(s_P[24 + t], v, g3);
      }
    }
    s_part[(half * 4 + 0) * 256 + c] = g0;
    s_part[(half * 4 + 1) * 256 + c] = g1;
    s_part[(half * 4 + 2) * 256 + c] = g2;
    s_part[(half * 4 + 3) * 256 + c] = g3;
  }
  __syncthreads();
  for (int e = tid; e < 1024; e += PT_THREADS) {
    const int h = e >> 8, c = e & 255;
    s_gbar[h * QP_STRIDE + c] = s_part[h * 256 + c] + s_part[(4 + h) * 256 + c];
  }
  __syncthreads();
  // (h) glimpse g[o] = W_gv[o] . gbar[head(o)]
  matvec_partial<T, 8, 1>(w16 + W16_GV, 256, smem_u32(s_gbar) + (uint32_t)((tid & 31) >> 3) * QP_STRIDE * 4u, 0, s_part);
  __syncthreads();
  if (tid < 256) {
    float v = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) v += s_part[ks * 256 + tid];
    s_g[tid] = v;
  }
  __syncthreads();
  // (i) lp = W_lk^T out_proj(g) = M1 g
  matvec_partial<T, 8, 1>(w16 + W16_M1, 256, smem_u32(s_g), 0, s_part);
  __syncthreads();
  if (tid < 256) {
    float v = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) v += s_part[ks * 256 + tid];
    s_lp[tid] = v;
  }
  __syncthreads();
  // (j) pointer logits: 10 tanh((lp . node_emb[n] + lp . tour_emb[t]) / 16)
  {
    float lv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) lv[i] = s_lp[lane + 32 * i];
    for (int n = warp; n < N + K; n += PT_THREADS / 32) {
      const uint32_t ra = ((n < N) ? (ne_saddr + (uint32_t)n * NE_STRIDE * 4u) : (te_saddr + (uint32_t)(n - N) * 1024u)) + lane * 4u;
      float dd = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) dd = fmaf(lds_f32(ra + i * 128u), lv[i], dd);
      dd = warp_sum(dd);
      if (lane == 0) {
        if (n < N) s_ln[n] = dd;
        else s_lt[n - N] = dd;
      }
    }
  }
  __syncthreads();
  for (int aa = tid; aa < A; aa += PT_THREADS)
    s_logit[aa] = s_mask[aa] ? -INFINITY : 10.0f * tanhf((s_ln[s_nbh[aa]] + s_lt[aa / k]) * 0.0625f);
  __syncthreads();
  // ---- log-softmax + selection (warp 0), same rules as decoder.cu
  if (warp == 0) {
    float mx = -INFINITY;
    bool nan = false;
    for (int e = lane; e < A; e += 32) {
      const float v = s_logit[e];
      nan |= (v != v);
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int e = lane; e < A; e += 32) sum += expf(s_logit[e] - mx);
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    float ent = 0.0f, best = -INFINITY;
    int besti = 0x7fffffff;
    for (int e = lane; e < A; e += 32) {
      const float lpv = s_logit[e] - lse;
      nan |= (lpv != lpv);
      s_logit[e] = lpv;
      if (st.logp) st.logp[(size_t)b * A + e] = lpv;
      if (lpv >= -10000.0f) ent -= lpv * expf(lpv);
      if (lpv > best) { best = lpv; besti = e; }
    }
    nan = __any_sync(0xffffffffu, nan);
    ent = warp_sum(ent);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    int choice = besti;
    if (a.decode_type == JAMPR_DECODE_SAMPLING) {
      const float u = a.uniforms ? a.uniforms[b] : jampr_uniform(a.seed, (uint32_t)b, (uint32_t)a.step_no);
      float run = 0.0f;
      int hit = -1, lastf = -1;
      for (int bs0 = 0; bs0 < A; bs0 += 32) {
        const int e = bs0 + lane;
        const float pe = (e < A) ? expf(s_logit[e]) : 0.0f;
        float inc = pe;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const float cum = run + inc;
        const unsigned hm = __ballot_sync(0xffffffffu, e < A && pe > 0.0f && cum > u);
        const unsigned fm = __ballot_sync(0xffffffffu, e < A && pe > 0.0f);
        if (hit < 0 && hm) hit = bs0 + __ffs(hm) - 1;
        if (fm) lastf = bs0 + 31 - __clz(fm);
        run = __shfl_sync(0xffffffffu, cum, 31);
      }
      choice = (hit >= 0) ? hit : lastf;
      if (choice < 0) choice = 0;
    }
    if (lane == 0) {
      if (choice >= A) choice = 0;
      st.action[2 * (size_t)b] = choice / k;
      st.action[2 * (size_t)b + 1] = s_nbh[choice];
      st.ll[b] = s_logit[choice];
      st.entropy[b] = ent;
      if (nan) atomicOr(st.status + b, JAMPR_ST_NAN_LOGITS);
    }
  }
}

static size_t pt_smem_bytes(const jampr_dims_t* d) {
  const size_t N = d->n_nodes, k = d->k_nbh, K = d->n_slots, A = K * k, L = d->plan_len;
  const size_t k4 = (k + 3) & ~size_t(3), nd = (N + 7) & ~size_t(7);
  auto al = [](size_t b) { return (b + 15) & ~size_t(15); };
  size_t s = 1024 + OVERLAY_BYTES;   // alignment slack + operand tiles / weight stages / parked node embedding
  s += al(64) + al(16) + al(4 * 128 * 8) + al(16 * 256 * 4) + al(6 * 384 * 4);          // bars, tmem slot, s_red, s_part, s_par
  s += al(N * k4 * 4) + al(nd * 4) + 2 * al(N * 4) + 2 * al((N + 4) * 4) + al(16);       // edge tables, counters
  s += al(K * L * 2) + al(4 * 128 * 4) + al(K * 256 * 4) + al(512 * 4) + al(256 * 4);    // plan, mean, te, q, ctxt
  s += 2 * al(4 * QP_STRIDE * 4) + 2 * al(256 * 4) + al(128 * 16) + al(128 * 4) + al(128 * 16);   // qp, gbar, g, lp, sn4, ln, pn4
  s += al(4 * JAMPR_MAX_SLOTS * 4) + al(JAMPR_MAX_SLOTS * 4) + al(4 * JAMPR_MAX_SLOTS * 4);       // qt, lt, P
  s += al(4 * A * 4) + al(A * 4) + al(K * 8 * 4) + al(A * 4) + al(K * N) + al(A) + al(JAMPR_MAX_SLOTS * 4);
  return s;
}

int64_t jampr_w16_elems(void) { return (int64_t)W16_END; }

template <typename T>
static int launch_pt(const PtArgs& a, size_t smem, cudaStream_t s) {
  CUDA_RET(cudaFuncSetAttribute(policy_step_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  policy_step_tc_kernel<T><<<a.d.n_inst * a.d.n_samples, PT_THREADS, smem, s>>>(a);
  return launch_status();
}

// One fused policy step on the tensor-core path.  store_cache: also write h_fin / ne / ne_stats (fp32) to HBM.
int jampr_policy_step_tc(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                         const jampr_state_t* st, int graph_fresh, int step_no, int decode_type, uint64_t seed,
                         const float* uniforms, int store_cache, cudaStream_t s) {
  if (d->n_nodes > 128 || d->n_slots > 4 || d->k_nbh > 127) return JAMPR_E_TOOLARGE;
  if (!w->blob_bf16 || w->n_bf16 < (int64_t)W16_END) return JAMPR_E_BADARG;
  if (!inst->h0_t || !inst->static_emb_t) return JAMPR_E_BADARG;      // column-chunk-major tables (jampr_node_encoder)
  const size_t smem = pt_smem_bytes(d);
  if (smem > 227 * 1024) return JAMPR_E_TOOLARGE;
  PtArgs a;
  a.d = *d;
  a.p = *inst;
  a.st = *st;
  a.w = w->blob;
  a.w16 = w->blob_bf16;
  a.uniforms = uniforms;
  a.seed = seed;
  a.step_no = step_no;
  a.decode_type = decode_type;
  a.graph_fresh = graph_fresh;
  a.store_cache = store_cache;
  a.ablate = 0;
  if (w->precision == JAMPR_PREC_F16) return launch_pt<__half>(a, smem, s);
  return launch_pt<__nv_bfloat16>(a, smem, s);
}
