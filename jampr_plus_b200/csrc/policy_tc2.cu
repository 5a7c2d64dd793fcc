// Tensor-core policy step, two CTAs per SM (sm_100a).  Same math as policy_tc.cu (tour-graph GNN on tcgen05, node
// embedding projection, per-tour embeddings, pointer decoder, selection: tour_encoder.py:159-243, policy.py:170-227,
// attn_decoder.py:62-98), re-tiled so that TWO rollouts are resident per SM and hide each other's latencies:
//   * 256 threads per CTA (8 warps: TMEM lane quarter = warp & 3, accumulator column half = warp >> 2), <= 128
//     registers, 256 TMEM columns, <= 113 KB shared memory:
//       - operand tiles cut to ceil(N/8) row groups (13 KB instead of 16 KB per tile at N = 101; the MMA still reads
//         128 rows, the overhang lands in mapped shared memory and only feeds accumulator rows >= N, never read);
//       - two 16 KB weight stages, each filled twice per layer (lin_root pair, then lin_rel pair);
//       - the node embedding is parked in the operand type (16-bit) over the dead operand tiles;
//       - decoder scratch overlays the weight stages and the edge tables.
//   * layer epilogue in two passes over the thread's 64 accumulator columns (pass 1 parks GELU + skip back into the
//     accumulator columns of TMEM and reduces the LayerNorm statistics, pass 2 normalises).
// One CTA = one rollout = one step.  See policy_tc.cu for the single-CTA-per-SM variant with fp32 parked embeddings.
#include <stdlib.h>
#include "policy_tc_common.cuh"

// timing experiments (tools/ablate.sh): build with `make EXTRA=-DJAMPR_ABLATE_BUILD`, select with JAMPR_TC_ABLATE
#ifdef JAMPR_ABLATE_BUILD
#define ABL(bit) ((a.ablate & (bit)) != 0)
#else
#define ABL(bit) false
#endif

// per-phase time line of sampled CTAs (tools/timeline.py): build with `make EXTRA=-DJAMPR_TIMELINE`
#ifdef JAMPR_TIMELINE
#define TL_N 40
__device__ uint32_t g_timeline[1024][TL_N];
#define TL(k) do { if (threadIdx.x == 0) s_tl[k] = (uint32_t)clock64(); } while (0)
#else
#define TL(k) do { } while (0)
#endif

#include "policy_tc2_parts.cuh"

// decoder mat-vec partials, 256 threads: y[o] = sum_c x[c] W[c][o]; thread = OPT adjacent outputs x one K-slice;
// partial sums at s_part[(slice * NV + v) * 256 + o]
template <typename T, int OPT, int NV>
__device__ __forceinline__ void matvec2(const T* __restrict__ W, int K, uint32_t x_saddr, int xstride, float* s_part) {
  using T2 = typename Elem<T>::T2;
  constexpr int NG = 256 / OPT, NS = P2_THREADS / NG;
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

// Per-instance packed edge tables (once per episode, from jampr_node_encoder): kNN rows padded to a multiple of
// four entries by repeating the last edge, and the node -> depot edges padded to a multiple of eight.
template <typename T>
__global__ void tc_prepare_kernel(jampr_dims_t d, jampr_instances_t p) {
  const int N = d.n_nodes, k = d.k_nbh, k4 = (k + 3) & ~3, nd = (N + 7) & ~7, ins = blockIdx.x;
  const int16_t* gk = p.knn + (size_t)ins * N * k;
  const float* gw = p.knn_w + (size_t)ins * N * k;
  for (int e = threadIdx.x; e < N * k4; e += blockDim.x) {
    const int u = e / k4, j = min(e - u * k4, k - 1);
    p.tc_tab[(size_t)ins * N * k4 + e] = agg_entry((uint32_t)(uint16_t)gk[u * k + j], bits_T<T>(gw[u * k + j]));
  }
  for (int u = threadIdx.x; u < nd; u += blockDim.x) {
    const int uu = min(u, N - 1);
    p.tc_dtab[(size_t)ins * nd + u] = agg_entry((uint32_t)uu, bits_T<T>(p.ttd[(size_t)ins * N + uu]));
  }
}

int jampr_tc_prepare(const jampr_dims_t* d, const jampr_instances_t* inst, int precision, cudaStream_t s) {
  if (!inst->tc_tab || !inst->tc_dtab) return JAMPR_E_BADARG;
  if (precision == JAMPR_PREC_F16) tc_prepare_kernel<__half><<<d->n_inst, 256, 0, s>>>(*d, *inst);
  else tc_prepare_kernel<__nv_bfloat16><<<d->n_inst, 256, 0, s>>>(*d, *inst);
  return launch_status();
}

// Column-chunk-major copies of h0 / static_emb (include/jampr_b200.h: h0_t, static_emb_t), once per episode after the node
// encoder.  The step kernel reads these tables with thread = node (accumulator row): row-major that is 32 different
// 128-byte lines per warp load (LSU-bound: 2.1 us for h0, 4.2 us for static_emb per rollout-step), here 512 contiguous bytes.
__global__ void tc_transpose_kernel(int N, int D4, const float4* __restrict__ src, float4* __restrict__ dst) {
  const size_t ins = blockIdx.x;
  const float4* s = src + ins * (size_t)N * D4;
  float4* t = dst + ins * (size_t)N * D4;
  for (int e = threadIdx.x; e < N * D4; e += blockDim.x) {
    const int c = e / N, n = e - c * N;
    t[e] = s[(size_t)n * D4 + c];
  }
}

int jampr_tc_finish(const jampr_dims_t* d, const jampr_instances_t* inst, cudaStream_t s) {
  if (!inst->h0_t || !inst->static_emb_t) return JAMPR_E_BADARG;
  tc_transpose_kernel<<<d->n_inst, 256, 0, s>>>(d->n_nodes, 32, reinterpret_cast<const float4*>(inst->h0),
                                                reinterpret_cast<float4*>(inst->h0_t));
  tc_transpose_kernel<<<d->n_inst, 256, 0, s>>>(d->n_nodes, 64, reinterpret_cast<const float4*>(inst->static_emb),
                                                reinterpret_cast<float4*>(inst->static_emb_t));
  return launch_status();
}

// The same mat-vec (8 outputs per thread, one input vector) software-pipelined over batches of eight weight rows:
// batch i+1 is in flight while batch i is consumed, and the first batch (`pre`) was requested by the caller long
// before — the weights depend on nothing, only the input vector does.  The chain of five dependent decoder stages
// is latency-bound on these L2 loads (profiles/r1_policy_step_tc2_ablation.txt).
template <typename T>
__device__ __forceinline__ void mv_prefetch(const T* __restrict__ W, int K, uint4 (&pre)[8]) {
  const int og = threadIdx.x & 31, ks = threadIdx.x >> 5;
  const T* wp = W + (size_t)(ks * (K / 8)) * 256 + og * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) pre[i] = __ldg(reinterpret_cast<const uint4*>(wp + (size_t)i * 256));
}

template <typename T>
__device__ __forceinline__ void mv_consume8(const uint4 (&raw)[8], uint32_t xa, int c0, float (&acc)[8]) {
  using T2 = typename Elem<T>::T2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 f0 = Elem<T>::to2(from_u32<T2>(raw[i].x)), f1 = Elem<T>::to2(from_u32<T2>(raw[i].y));
    const float2 f2_ = Elem<T>::to2(from_u32<T2>(raw[i].z)), f3 = Elem<T>::to2(from_u32<T2>(raw[i].w));
    const float2 xv = f2(lds_f32(xa + (uint32_t)(c0 + i) * 4u));
    float2 a0 = __ffma2_rn(f0, xv, make_float2(acc[0], acc[1]));
    float2 a1 = __ffma2_rn(f1, xv, make_float2(acc[2], acc[3]));
    float2 a2 = __ffma2_rn(f2_, xv, make_float2(acc[4], acc[5]));
    float2 a3 = __ffma2_rn(f3, xv, make_float2(acc[6], acc[7]));
    acc[0] = a0.x; acc[1] = a0.y; acc[2] = a1.x; acc[3] = a1.y; acc[4] = a2.x; acc[5] = a2.y; acc[6] = a3.x; acc[7] = a3.y;
  }
}

template <typename T>
__device__ __forceinline__ void matvec2_pipe(const T* __restrict__ W, int K, uint32_t x_saddr, float* s_part, uint4 (&pre)[8]) {
  const int og = threadIdx.x & 31, ks = threadIdx.x >> 5;
  const int kn = K / 8, k0 = ks * kn;
  const T* wp = W + (size_t)k0 * 256 + og * 8;
  const uint32_t xa = x_saddr + (uint32_t)k0 * 4u;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
#pragma unroll 1
  for (int c0 = 0; c0 < kn; c0 += 16) {
    uint4 nxt[8];
    if (c0 + 8 < kn) {
#pragma unroll
      for (int i = 0; i < 8; ++i) nxt[i] = __ldg(reinterpret_cast<const uint4*>(wp + (size_t)(c0 + 8 + i) * 256));
    }
    mv_consume8<T>(pre, xa, c0, acc);
    if (c0 + 16 < kn) {
#pragma unroll
      for (int i = 0; i < 8; ++i) pre[i] = __ldg(reinterpret_cast<const uint4*>(wp + (size_t)(c0 + 16 + i) * 256));
    }
    if (c0 + 8 < kn) mv_consume8<T>(nxt, xa, c0 + 8, acc);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s_part[ks * 256 + og * 8 + j] = acc[j];
}

// fp32-weight variant of the pipelined mat-vec (batches of four weight rows, 8 outputs per thread).  Used for the two
// decoder stages whose 16-bit weights dominated the log-probability error (tools/emulate_tc_precision.py: glimpse keys
// 9.3e-3, ctxt_proj 2.8e-3 of the 1.0e-2 total at N = 101; the other three stages < 6e-4 each): the glimpse scores
// qp_h . act[a] are O(10) before the softmax, so a 2^-11 relative error of qp moves the glimpse weights directly.
__device__ __forceinline__ void mvf_prefetch(const float* __restrict__ W, int K, float4 (&pre)[8]) {
  const int og = threadIdx.x & 31, ks = threadIdx.x >> 5;
  const float* wp = W + (size_t)(ks * (K / 8)) * 256 + og * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    pre[2 * i] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)i * 256));
    pre[2 * i + 1] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)i * 256) + 1);
  }
}

__device__ __forceinline__ void mvf_consume4(const float4 (&raw)[8], uint32_t xa, int c0, float (&acc)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 xv = f2(lds_f32(xa + (uint32_t)(c0 + i) * 4u));
    const float4 wa = raw[2 * i], wb = raw[2 * i + 1];
    float2 a0 = __ffma2_rn(make_float2(wa.x, wa.y), xv, make_float2(acc[0], acc[1]));
    float2 a1 = __ffma2_rn(make_float2(wa.z, wa.w), xv, make_float2(acc[2], acc[3]));
    float2 a2 = __ffma2_rn(make_float2(wb.x, wb.y), xv, make_float2(acc[4], acc[5]));
    float2 a3 = __ffma2_rn(make_float2(wb.z, wb.w), xv, make_float2(acc[6], acc[7]));
    acc[0] = a0.x; acc[1] = a0.y; acc[2] = a1.x; acc[3] = a1.y; acc[4] = a2.x; acc[5] = a2.y; acc[6] = a3.x; acc[7] = a3.y;
  }
}

__device__ __forceinline__ void matvec2_pipe_f32(const float* __restrict__ W, int K, uint32_t x_saddr, float* s_part,
                                                 float4 (&pre)[8]) {
  const int og = threadIdx.x & 31, ks = threadIdx.x >> 5;
  const int kn = K / 8, k0 = ks * kn;                  // kn is a multiple of 8 (K = 256, 512)
  const float* wp = W + (size_t)k0 * 256 + og * 8;
  const uint32_t xa = x_saddr + (uint32_t)k0 * 4u;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
#pragma unroll 1
  for (int c0 = 0; c0 < kn; c0 += 8) {
    float4 nxt[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      nxt[2 * i] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(c0 + 4 + i) * 256));
      nxt[2 * i + 1] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(c0 + 4 + i) * 256) + 1);
    }
    mvf_consume4(pre, xa, c0, acc);
    if (c0 + 8 < kn) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pre[2 * i] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(c0 + 8 + i) * 256));
        pre[2 * i + 1] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(c0 + 8 + i) * 256) + 1);
      }
    }
    mvf_consume4(nxt, xa, c0 + 4, acc);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s_part[ks * 256 + og * 8 + j] = acc[j];
}

// SPLIT = true: stop after the query / tour embeddings and hand the decoder over to decoder_tc.cu
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(P2_THREADS, 2) policy_step_tc2_kernel(const PtArgs a) {
  using T2 = typename Elem<T>::T2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const jampr_dims_t& d = a.d;
  const jampr_state_t& st = a.st;
  const jampr_instances_t& p = a.p;
  const uint8_t allvis_b = st.logp ? (uint8_t)0 : st.allvis[blockIdx.x];     // issued together with the control word
  if (!jampr_step_active(st.ctl, a.step_no)) return;
  const int N = d.n_nodes, K = d.n_slots, k = d.k_nbh, A = K * k, L = d.plan_len;
  const int k4 = (k + 3) & ~3, nd = (N + 7) & ~7;
  const int b = blockIdx.x, ins = b / d.n_samples;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (jampr_rollout_idle2(st, b, K, allvis_b)) {
    if (tid == 0) {
      st.action[2 * (size_t)b] = 0;
      st.action[2 * (size_t)b + 1] = 0;
      st.ll[b] = 0.0f;
      st.entropy[b] = 0.0f;
    }
    return;
  }
#ifdef JAMPR_TIMELINE
  __shared__ uint32_t s_tl[TL_N];
  if (tid < TL_N) s_tl[tid] = 0u;
  __syncthreads();
#endif
  TL(0);
  const float* __restrict__ w = a.w;
  const T* __restrict__ w16 = reinterpret_cast<const T*>(a.w16);
  const bool fresh = a.graph_fresh != 0;
  const uint32_t ts = p2_tile_stride(N);

  // ---------------------------------------------------------------- shared memory carve-up
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;                       // 4 operand tiles of `ts` bytes: agg (K-blocks 0,1), h (K-blocks 2,3)
  uint8_t* sW = base + 4 * ts;              // 2 weight stages (region R1)
  uint8_t* s_ne16 = base;                   // parked node embedding rows (16-bit), over sA (+ the head of sW)
  uint8_t* r1 = sW + 1024;                  // decoder scratch over the weight stages (first KB: tail of s_ne16)
  uint32_t off = 4 * ts + P2_R1_BYTES;
  uint8_t* r2 = base + off;                 // edge tables | decoder scratch
  off += p2_r2_bytes(N, k, A);
  auto take = [&](uint32_t bytes) {
    uint8_t* rr = base + off;
    off += p2_al(bytes);
    return rr;
  };
  uint64_t* bars = reinterpret_cast<uint64_t*>(take(32));     // full[0..1], root, acc
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(take(16));
  float2* s_red = reinterpret_cast<float2*>(take(2 * 128 * sizeof(float2)));
  float* s_par = reinterpret_cast<float*>(take(384 * 4));              // current layer: bias | gamma | beta
  int* s_cnt = reinterpret_cast<int*>(take(16));
  int16_t* s_plan = reinterpret_cast<int16_t*>(take((uint32_t)K * L * 2));
  float* s_qt = reinterpret_cast<float*>(take(4 * JAMPR_MAX_SLOTS * 4));
  float* s_lt = reinterpret_cast<float*>(take(JAMPR_MAX_SLOTS * 4));
  float* s_P = reinterpret_cast<float*>(take(4 * JAMPR_MAX_SLOTS * 4));
  float* s_feat = reinterpret_cast<float*>(take((uint32_t)K * 8 * 4));
  int* s_nbh = reinterpret_cast<int*>(take((uint32_t)A * 4));
  int8_t* s_pos = reinterpret_cast<int8_t*>(take((uint32_t)K * N));    // slot t, node n -> candidate index j or -1
  int16_t* s_cand = reinterpret_cast<int16_t*>(take((uint32_t)N * 2));  // nodes that are a candidate of some slot, ascending
  uint8_t* s_mask = take(A);
  int* s_idx = reinterpret_cast<int*>(take(JAMPR_MAX_SLOTS * 4));
  // region R2 during the GNN: edge tables
  uint32_t o2 = 0;
  auto take2 = [&](uint32_t bytes) {
    uint8_t* rr = r2 + o2;
    o2 += p2_al(bytes);
    return rr;
  };
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(take2((uint32_t)N * k4 * 4));
  uint32_t* s_dtab = reinterpret_cast<uint32_t*>(take2((uint32_t)nd * 4));
  uint32_t* s_lf = reinterpret_cast<uint32_t*>(take2((uint32_t)N * 4));
  uint32_t* s_lr = reinterpret_cast<uint32_t*>(take2((uint32_t)N * 4));
  uint32_t* s_ef = reinterpret_cast<uint32_t*>(take2((uint32_t)(N + 4) * 4));
  uint32_t* s_er = reinterpret_cast<uint32_t*>(take2((uint32_t)(N + 4) * 4));
  // region R2 during the decoder (tables are dead after the last aggregation)
  float* s_mean = reinterpret_cast<float*>(r2);                        // [4][128]
  float* s_stat = reinterpret_cast<float*>(r2 + 2048);                 // [4][256]: sum half0/1, max half0/1
  float4* s_sn4 = reinterpret_cast<float4*>(r2 + 6144);                // per node: glimpse score of the 4 heads
  float* s_ln = reinterpret_cast<float*>(r2 + 8192);                   // per node: logit dot product
  float4* s_pn4 = reinterpret_cast<float4*>(r2 + 8704);                // per node: glimpse weight of the 4 heads
  float* s_sc = reinterpret_cast<float*>(r2 + 10752);                  // [4][A]
  float* s_logit = reinterpret_cast<float*>(r2 + 10752 + p2_al(16u * A));
  // region R1 during the decoder (weight stages are dead after the projection MMAs)
  float* s_part = reinterpret_cast<float*>(r1);                        // 8 KB
  float* s_te = reinterpret_cast<float*>(r1 + 8192);                   // [4][256]
  float* s_q = reinterpret_cast<float*>(r1 + 12288);                   // [512]
  float* s_ctxt = reinterpret_cast<float*>(r1 + 14336);                // [256]
  float* s_qp = reinterpret_cast<float*>(r1 + 15360);                  // [4][260]
  float* s_gbar = reinterpret_cast<float*>(r1 + 19520);                // [4][260]
  float* s_g = reinterpret_cast<float*>(r1 + 23680);                   // [256]
  float* s_lp = reinterpret_cast<float*>(r1 + 24704);                  // [256]  (ends at 25728 <= 31744)
  uint64_t* full = bars;
  uint64_t* bar_root = bars + 2;
  uint64_t* bar_acc = bars + 3;

  // ---------------------------------------------------------------- prologue
  if (fresh) {
    if (tid == 0) {
      mbar_init(&full[0], 1);
      mbar_init(&full[1], 1);
      mbar_init(bar_root, 1);
      mbar_init(bar_acc, 1);
      mbar_fence_init();
      load_w2<T>(sW, 0, a.w16, 2, full);       // layer 0, lin_root pair (K-blocks 2,3)
      load_w2<T>(sW, 1, a.w16, 3, full);
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    // operand rows >= N are never initialised: the MMA reads them, but they only feed accumulator rows >= N, which
    // nothing reads (rows are independent in every later phase)
  }
  // The per-rollout state comes from HBM (10 GB of it per launch: nothing stays in L2), so the prologue is a chain of
  // dependent round trips.  It is written as two links: everything addressed by (b, tid) alone is requested first
  // (registers), then the two dependent gathers (tour rows by st.row, link distances by pred / succ).
  const int rdx = min(q4_row(tid), N - 1);
  // ---- link 1
  int row_of = 0, nb0 = 0, pr = 0, su = 0;
  uint8_t mk0 = 0;
  bool vs = false;
  float ttd_u = 0.0f, f_cur = 0.0f, f_cap = 0.0f, f_time = 0.0f, f_cttd = 0.0f;
  int f_row = 0;
  if (tid < K * L) row_of = st.row[(size_t)b * K + tid / L];
  if (tid < A) {
    nb0 = st.nbh[(size_t)b * A + tid];
    mk0 = st.mask[(size_t)b * A + tid];
  }
  if (fresh && tid < N) {
    pr = st.pred[(size_t)b * N + tid];
    su = st.succ[(size_t)b * N + tid];
    vs = tid > 0 && st.visited[(size_t)b * N + tid] != 0;
    ttd_u = p.ttd[(size_t)ins * N + tid];
  }
  if (tid < K) {
    const size_t bk = (size_t)b * K + tid;
    f_row = st.row[bk];
    f_cur = (float)st.cur[bk];
    f_cap = st.cap[bk];
    f_time = st.time[bk];
    f_cttd = st.cttd[bk];
  }
  const int n_tab4 = (N * k4) / 4;
  const uint4* gt = reinterpret_cast<const uint4*>(p.tc_tab + (size_t)ins * N * k4);
  uint4 tabv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    tabv[i] = make_uint4(0u, 0u, 0u, 0u);
    if (fresh && tid + i * P2_THREADS < n_tab4) tabv[i] = __ldg(gt + tid + i * P2_THREADS);
  }
  // h_0 row chunk 0 (per instance, L2): requested here, used after the prologue barriers
  float4 h0v[8];
  if (fresh) {
    const float4* src = reinterpret_cast<const float4*>(p.h0_t) + ((size_t)ins * 32 + (tid >> 7) * 16) * N + rdx;   // [chunk][node]
#pragma unroll
    for (int i = 0; i < 8; ++i) h0v[i] = __ldg(src + (size_t)i * N);
  }
  if (tid < 4) s_cnt[tid] = 0;
  for (int e = tid; e < K * N; e += P2_THREADS) s_pos[e] = (int8_t)-1;
  // ---- link 2
  const float* dist = p.dist + (size_t)ins * N * N;
  float d_pr = 0.0f, d_su = 0.0f;
  if (vs) {
    d_pr = dist[(size_t)tid * N + pr];
    d_su = dist[(size_t)tid * N + su];
  }
  if (tid < K * L) s_plan[tid] = st.plan[((size_t)b * d.k_max + row_of) * L + (tid - (tid / L) * L)];
  // ---- park link 1
  if (tid < A) {
    s_nbh[tid] = nb0;
    s_mask[tid] = mk0;
  }
  for (int e = tid + P2_THREADS; e < A; e += P2_THREADS) {
    s_nbh[e] = st.nbh[(size_t)b * A + e];
    s_mask[e] = st.mask[(size_t)b * A + e];
  }
  if (tid < K) {
    s_feat[tid * 8 + 0] = f_cur;
    s_feat[tid * 8 + 1] = f_cap;
    s_feat[tid * 8 + 2] = f_time;
    s_feat[tid * 8 + 3] = f_cttd;
    s_feat[tid * 8 + 4] = (float)(d.k_max - f_row) / (float)d.k_max;
  }
  if (fresh) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (tid + i * P2_THREADS < n_tab4) reinterpret_cast<uint4*>(s_tab)[tid + i * P2_THREADS] = tabv[i];
    for (int e = tid + 3 * P2_THREADS; e < n_tab4; e += P2_THREADS) reinterpret_cast<uint4*>(s_tab)[e] = __ldg(gt + e);
    const uint4* gd = reinterpret_cast<const uint4*>(p.tc_dtab + (size_t)ins * nd);
    for (int e = tid; e < nd / 4; e += P2_THREADS) reinterpret_cast<uint4*>(s_dtab)[e] = __ldg(gd + e);
  }
  __syncthreads();
  for (int e = tid; e < A; e += P2_THREADS) s_pos[(e / k) * N + s_nbh[e]] = (int8_t)(e % k);
  if (fresh && tid < N) {
    s_lf[tid] = vs ? agg_entry((uint32_t)pr, bits_T<T>(d_pr)) : 0u;
    s_lr[tid] = vs ? agg_entry((uint32_t)su, bits_T<T>(d_su)) : 0u;
    if (vs) {
      const uint32_t en = agg_entry((uint32_t)tid, bits_T<T>(ttd_u));
      if (su == 0) s_ef[atomicAdd(&s_cnt[0], 1)] = en;
      if (pr == 0) s_er[atomicAdd(&s_cnt[1], 1)] = en;
    }
  }
  if (tid < K) {
    int idx = 0;
    while (idx < L && s_plan[tid * L + idx] != 0) ++idx;
    s_idx[tid] = (idx >= L) ? 0 : idx;        // argmin over an all-nonzero row is position 0 (tour_encoder.py:236)
  }
  tc_fence_before();
  __syncthreads();
  TL(1);
  tc_fence_after();

  const int q4 = warp & 3, ch = warp >> 2, r = q4 * 32 + lane;      // epilogue role: row r, columns 64ch..64ch+63
  if (warp == 1) {
    // compact list of the distinct candidate nodes in ascending order (deterministic): the decoder's per-node dot
    // products and weighted sums only visit these (typically 50..70 of the 101 nodes)
    int cnt = 0;
    for (int n0 = 0; n0 < N; n0 += 32) {
      const int n = n0 + lane;
      bool is = false;
      if (n < N)
        for (int t = 0; t < K; ++t) is |= s_pos[t * N + n] >= 0;
      const unsigned m = __ballot_sync(0xffffffffu, is);
      if (is) s_cand[cnt + __popc(m & ((1u << lane) - 1u))] = (int16_t)n;
      cnt += __popc(m);
    }
    if (lane == 0) s_cnt[2] = cnt;
  }

  if (fresh) {
    const uint32_t tmem = *tmem_slot;
    const uint32_t ta = tmem + ((uint32_t)(q4 * 32) << 16);
    const uint32_t idesc = umma_idesc(128, 128, Elem<T>::fmt);
    const uint32_t a_addr = smem_u32(sA), w_addr = smem_u32(sW);
    const int n_endf = s_cnt[0], n_endr = s_cnt[1];
    const int n_ef4 = (n_endf + 3) & ~3, n_er4 = (n_endr + 3) & ~3;
    if (tid < 3 && n_endf > 0 && n_endf + tid < n_ef4) s_ef[n_endf + tid] = s_ef[n_endf - 1];
    if (tid >= 32 && tid < 35 && n_endr > 0 && n_endr + (tid - 32) < n_er4) s_er[n_endr + tid - 32] = s_er[n_endr - 1];
    // ---- h_0 (hoisted node_emb_input_proj, per instance): fp32 -> TMEM skip columns, 16-bit -> h tiles
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col0 = ch * 64 + c * 32;
      uint32_t hv[32];
      if (r < N) {
        const float4* src = reinterpret_cast<const float4*>(p.h0_t) + ((size_t)ins * 32 + (col0 >> 2)) * N + r;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = (c == 0) ? h0v[i] : __ldg(src + (size_t)i * N);
          hv[4 * i] = __float_as_uint(v.x); hv[4 * i + 1] = __float_as_uint(v.y);
          hv[4 * i + 2] = __float_as_uint(v.z); hv[4 * i + 3] = __float_as_uint(v.w);
        }
        uint8_t* tile = sA + (2 + ch) * ts;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            pk[e] = as_u32(Elem<T>::from2(__uint_as_float(hv[i * 8 + 2 * e]), __uint_as_float(hv[i * 8 + 2 * e + 1])));
          *reinterpret_cast<uint4*>(tile + sw128_off(r, c * 4 + i)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) hv[i] = 0u;
      }
      tmem_st32(ta + 128 + col0, hv);
    }
    tmem_wait_st();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
  TL(2);
    tc_fence_after();

    const uint32_t tab_saddr = smem_u32(s_tab), dtab_saddr = smem_u32(s_dtab), par_saddr = smem_u32(s_par);
    const uint32_t ef_saddr = smem_u32(s_ef), er_saddr = smem_u32(s_er);
    // layers 0..5 = GraphConv blocks; "layer" 6 = node_emb_output_proj (two 128-column halves, no aggregation)
#pragma unroll 1
    for (int l = 0; l < 7; ++l) {
      const uint32_t ph = (uint32_t)(l & 1);
      const bool proj = (l == 6);
      if (tid == 0) {
        // first weight pair: lin_root(h) (K-blocks 2,3) — or W_o rows 0..127 — overlaps the aggregation below
        if (!ABL(32)) {
          mbar_wait(&full[0], 0u);
          mbar_wait(&full[1], 0u);
        }
        tc_fence_after();
        if (!(ABL(8))) {
          mma_kblock(tmem, a_addr + 2 * ts, w_addr, idesc, true);
          mma_kblock(tmem, a_addr + 3 * ts, w_addr + 16384, idesc, false);
        }
        umma_commit(bar_root);
      }
      if (!proj) {
        if (tid >= 32 && tid < 128)
          reinterpret_cast<float4*>(s_par)[tid - 32] =
              __ldg(reinterpret_cast<const float4*>(w + OFF_TE_BLK + (size_t)l * BLK_FLOATS + BLK_B) + (tid - 32));
        if (l == 2 || l == 5) { if (!(ABL(1))) agg_nbh_tc2<T>(sA, ts, tab_saddr, k4, dtab_saddr, nd, N, tid); }
        else if (l < 2) agg_tour_tc2<T>(sA, ts, s_lf, ef_saddr, n_ef4, N, tid);
        else agg_tour_tc2<T>(sA, ts, s_lr, er_saddr, n_er4, N, tid);
        fence_proxy_async();
      }
      if (tid == 0) {
        // warp 0's share of the aggregation ran while the lin_root MMAs executed; both stages are free once they are
        // complete: stream the second pair
        mbar_wait(bar_root, ph);
        load_w2<T>(sW, 0, a.w16, l * 4 + (proj ? 2 : 0), full);
        load_w2<T>(sW, 1, a.w16, l * 4 + (proj ? 3 : 1), full);
      }
      __syncthreads();
      TL(3 + 3 * l);
      if (proj) {
        // per-tour mean of h over buffer positions 0..idx (one depot entry included, tour_encoder.py:236-243);
        // overlaps the projection MMAs.  The edge tables are dead: s_mean overlays them (hence the barrier above).
        for (int e = tid; e < K * 64; e += P2_THREADS) {
          const int t = e >> 6, c2 = e & 63;
          const uint32_t toff = (uint32_t)(2 + (c2 >> 5)) * ts;
          const uint32_t chk = (uint32_t)(c2 & 31) >> 2, within = (uint32_t)(c2 & 3) * 4u;
          const int idx = s_idx[t];
          float sx = 0.0f, sy = 0.0f;
          for (int q = 0; q <= idx; ++q) {
            const uint32_t u = (uint32_t)s_plan[t * L + q];
            const float2 v = Elem<T>::to2(*reinterpret_cast<const T2*>(sA + toff + sw128_off(u, chk) + within));
            sx += v.x;
            sy += v.y;
          }
          const float inv = 1.0f / (float)(idx + 1);
          s_mean[t * 128 + 2 * c2] = sx * inv;
          s_mean[t * 128 + 2 * c2 + 1] = sy * inv;
        }
      }
      if (tid == 0) {
        if (!ABL(32)) {
          mbar_wait(&full[0], 1u);
          mbar_wait(&full[1], 1u);
        }
        tc_fence_after();
        if (ABL(8)) {
        } else if (!proj) {
          mma_kblock(tmem, a_addr, w_addr, idesc, false);                   // lin_rel(agg): K-blocks 0,1
          mma_kblock(tmem, a_addr + ts, w_addr + 16384, idesc, false);
        } else {
          mma_kblock(tmem + 128, a_addr + 2 * ts, w_addr, idesc, true);    // W_o rows 128..255
          mma_kblock(tmem + 128, a_addr + 3 * ts, w_addr + 16384, idesc, false);
        }
        umma_commit(bar_acc);
      }
      if (proj) break;
      mbar_wait(bar_acc, ph);                         // every thread waits for the accumulator itself (no block barrier hop)
      TL(4 + 3 * l);
      if (tid == 0) {
        const int nl = l + 1;                         // next layer's first pair: both stages are free again
        load_w2<T>(sW, 0, a.w16, nl * 4 + (nl == 6 ? 0 : 2), full);
        load_w2<T>(sW, 1, a.w16, nl * 4 + (nl == 6 ? 1 : 3), full);
      }
      tc_fence_after();
      if (!(ABL(2)))
        layer_epilogue2<T, 1>(tmem, sA, ts, s_red, par_saddr, N,
                              (l == 5 && a.store_cache) ? st.h_fin + (size_t)b * N * 128 : nullptr, tid);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      TL(5 + 3 * l);
    }
    // ---- node embedding rows: ne = acc + b + static_emb -> parked 16-bit rows (+ fp32 to HBM when cached).
    // The static_emb rows come from L2: chunk c+1 is in flight while chunk c is combined; chunk 0 is requested
    // before the wait for the projection MMAs.
    // chunk i (columns 128 ch + 4 i ..) of this thread's node: [chunk][node] layout, i-th chunk at se_row[i * N]
    const float4* se_row = reinterpret_cast<const float4*>(p.static_emb_t) + ((size_t)ins * 64 + ch * 32) * N + min(r, N - 1);
    float4 sv0[8], sv1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sv0[i] = __ldg(se_row + (size_t)i * N);
#pragma unroll
    for (int i = 0; i < 8; ++i) sv1[i] = __ldg(se_row + (size_t)(8 + i) * N);
    if (tid == 0) mbar_wait(bar_acc, 0u);
    __syncthreads();          // projection complete and per-tour means done: the h tiles may be overwritten
  TL(21);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col0 = ch * 128 + c * 32;
      uint32_t acc[32];
      tmem_ld32_nowait(ta + col0, acc);
      tmem_wait_ld();
      float4 (&cur)[8] = (c & 1) ? sv1 : sv0;
      if (r < N) {
        const float4* bo = reinterpret_cast<const float4*>(w + OFF_TE_OUT_B + col0);
        float4* gdst = a.store_cache ? reinterpret_cast<float4*>(st.ne + ((size_t)b * N + r) * 256 + col0) : nullptr;
        uint8_t* dst = s_ne16 + (size_t)r * NE16_STRIDE + col0 * 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 sa = cur[2 * i], sb = cur[2 * i + 1], ba = __ldg(bo + 2 * i), bb = __ldg(bo + 2 * i + 1);
          const float2 v0 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(acc[8 * i]), __uint_as_float(acc[8 * i + 1])), make_float2(ba.x, ba.y)), make_float2(sa.x, sa.y));
          const float2 v1 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(acc[8 * i + 2]), __uint_as_float(acc[8 * i + 3])), make_float2(ba.z, ba.w)), make_float2(sa.z, sa.w));
          const float2 v2 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(acc[8 * i + 4]), __uint_as_float(acc[8 * i + 5])), make_float2(bb.x, bb.y)), make_float2(sb.x, sb.y));
          const float2 v3 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(acc[8 * i + 6]), __uint_as_float(acc[8 * i + 7])), make_float2(bb.z, bb.w)), make_float2(sb.z, sb.w));
          const float4 va = make_float4(v0.x, v0.y, v1.x, v1.y), vb = make_float4(v2.x, v2.y, v3.x, v3.y);
          const uint4 pk = make_uint4(as_u32(Elem<T>::from2(va.x, va.y)), as_u32(Elem<T>::from2(va.z, va.w)),
                                      as_u32(Elem<T>::from2(vb.x, vb.y)), as_u32(Elem<T>::from2(vb.z, vb.w)));
          *reinterpret_cast<uint4*>(dst + i * 16) = pk;
          if (gdst) { gdst[2 * i] = va; gdst[2 * i + 1] = vb; }
        }
      }
      if (c < 2) {          // refill this buffer with the chunk two ahead: in flight during the next chunk
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = __ldg(se_row + (size_t)((c + 2) * 8 + i) * N);
      }
    }
    tc_fence_before();
  } else {
    // cached tour-graph embedding (tour_encoder.py:163-166): per-tour means from h_fin, node embedding from ne
    for (int e = tid; e < K * 128; e += P2_THREADS) {
      const int t = e >> 7, c = e & 127;
      const int idx = s_idx[t];
      float sx = 0.0f;
      for (int q = 0; q <= idx; ++q)
        sx += from_T<T>(to_T<T>(st.h_fin[((size_t)b * N + s_plan[t * L + q]) * 128 + c]));
      s_mean[t * 128 + c] = sx / (float)(idx + 1);
    }
    for (int e = tid; e < N * 32; e += P2_THREADS) {
      const int u = e >> 5, c8 = (e & 31) * 8;
      const float4* src = reinterpret_cast<const float4*>(st.ne + ((size_t)b * N + u) * 256 + c8);
      const float4 va = src[0], vb = src[1];
      *reinterpret_cast<uint4*>(s_ne16 + (size_t)u * NE16_STRIDE + c8 * 2) =
          make_uint4(as_u32(Elem<T>::from2(va.x, va.y)), as_u32(Elem<T>::from2(va.z, va.w)),
                     as_u32(Elem<T>::from2(vb.x, vb.y)), as_u32(Elem<T>::from2(vb.z, vb.w)));
    }
  }
  __syncthreads();
  TL(22);
  if (fresh && warp == 0) {
    tc_fence_after();
    tmem_dealloc(*tmem_slot, 256);
  }
  if (SPLIT && fresh && warp == 1) {
    // split decoder: the parked 16-bit rows also go to HBM (decoder_tc.cu gathers the candidate rows from there), one
    // 512-byte bulk copy per row by the copy engine while the CTA computes the column statistics below
    fence_proxy_async();
    uint8_t* g = reinterpret_cast<uint8_t*>(st.ne16) + (size_t)b * N * 512;
    for (int rr = lane; rr < N; rr += 32) bulk_s2g(g + (size_t)rr * 512, s_ne16 + (size_t)rr * NE16_STRIDE, 512u);
    bulk_commit();
  }

  // ---------------------------------------------------------------- decoder
  const uint32_t ne_saddr = smem_u32(s_ne16), te_saddr = smem_u32(s_te);
  // (b) tour_emb[t] = folded tour-feature branch + W_to2 mean_h[t]          (tour_encoder.py:202-213)
  const bool mv = !(ABL(4));
  if (ABL(16)) {      // timing experiment: first feasible action, nothing else
    if (tid == 0) {
      int choice = 0;
      for (int e = 0; e < A; ++e) if (!s_mask[e]) { choice = e; break; }
      st.action[2 * (size_t)b] = choice / k;
      st.action[2 * (size_t)b + 1] = s_nbh[choice];
      st.ll[b] = 0.0f;
      st.entropy[b] = 0.0f;
    }
    return;
  }
  uint4 pre[8];
  float4 pref[8];
  if (!SPLIT) mvf_prefetch(w + OFF_CTXT_WT, 512, pref);   // first weight batch of stage (c): no dependency, hides its latency
  if (mv && !SPLIT) matvec2<T, 2, 4>(w16 + W16_TO2, 128, smem_u32(s_mean), 128, s_part);     // NV = 4 >= K slots (K <= 4 here)
  // (a) column mean / max of node_emb over the nodes (policy.py:179-181): thread = column pair x row parity
  {
    const int c2 = tid & 127, half = tid >> 7;
    // the maximum commutes with the 16-bit rounding: it is taken on the packed operand type, the sum in fp32
    T2 mxh = *reinterpret_cast<const T2*>(s_ne16 + (size_t)half * NE16_STRIDE + c2 * 4);
    float2 sm = Elem<T>::to2(mxh);
    for (int n = half + 2; n < N; n += 2) {
      const T2 raw = *reinterpret_cast<const T2*>(s_ne16 + (size_t)n * NE16_STRIDE + c2 * 4);
      mxh = __hmax2(mxh, raw);
      sm = __fadd2_rn(sm, Elem<T>::to2(raw));
    }
    const float2 mx = Elem<T>::to2(mxh);
    *reinterpret_cast<float2*>(s_stat + half * 256 + 2 * c2) = sm;
    *reinterpret_cast<float2*>(s_stat + (2 + half) * 256 + 2 * c2) = mx;
  }
  __syncthreads();
  {
    const float nmean = (s_stat[tid] + s_stat[256 + tid]) / (float)N;
    const float nmax = fmaxf(s_stat[512 + tid], s_stat[768 + tid]);
    if (a.store_cache && fresh) {
      st.ne_stats[(size_t)b * 512 + tid] = nmean;
      st.ne_stats[(size_t)b * 512 + 256 + tid] = nmax;
    }
    if (SPLIT) {
      // split decoder: the column statistics, the per-tour means of h and the tour features go to HBM; the tour
      // embeddings (tour_encoder.py:202-213), the query (policy.py:179-182) and everything after them are computed by
      // decoder_tc.cu for 128 rollouts at a time (dq / dqp are its scratch rows of this rollout)
      st.dq[(size_t)b * 512 + tid] = nmean;
      st.dq[(size_t)b * 512 + 256 + tid] = nmax;
      for (int e = tid; e < K * 128; e += P2_THREADS) st.dqp[((size_t)b * 4 + (e >> 7)) * 256 + (e & 127)] = s_mean[e];
      if (tid < K * 8) st.dqp[((size_t)b * 4 + (tid >> 3)) * 256 + 128 + (tid & 7)] = s_feat[tid];
    } else {
      float sm = 0.0f, mx = -INFINITY;
      for (int t = 0; t < K; ++t) {
        float v = __ldg(w + OFF_TOF_B + tid);
#pragma unroll
        for (int i = 0; i < 5; ++i) v = fmaf(s_feat[t * 8 + i], __ldg(w + OFF_TOF_WT + i * 256 + tid), v);
        v += s_part[t * 256 + tid] + s_part[(4 + t) * 256 + tid];
        s_te[t * 256 + tid] = v;
        sm += v;
        mx = fmaxf(mx, v);
      }
      s_q[tid] = nmean + sm / (float)K;
      s_q[256 + tid] = nmax + mx;
    }
  }
  if (SPLIT) {
    if (fresh && warp == 1) bulk_wait_all();       // the copy engine has read the parked rows: shared memory may go
#ifdef JAMPR_TIMELINE
    if (tid == 0) {
      s_tl[23] = (uint32_t)clock64();
      for (int q = 24; q <= 32; ++q) s_tl[q] = s_tl[23];
      if ((b & 63) == 0 && (b >> 6) < 1024)
        for (int q = 0; q < TL_N; ++q) g_timeline[b >> 6][q] = s_tl[q];
    }
#endif
    return;
  }
  __syncthreads();
  TL(23);
  // (c) ctxt = ctxt_proj(query)
  if (mv) matvec2_pipe_f32(w + OFF_CTXT_WT, 512, smem_u32(s_q), s_part, pref);      // fp32 weights (see mvf_prefetch)
  mvf_prefetch(w + OFF_GK_W, 256, pref);
  __syncthreads();
  {
    float v = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) v += s_part[ks * 256 + tid];
    s_ctxt[tid] = v;
  }
  __syncthreads();
  TL(24);
  // (d) qp[h] = W_gk,h^T ctxt_h : K-slice ks (32 rows of set_proj) belongs to head ks / 2
  if (mv) matvec2_pipe_f32(w + OFF_GK_W, 256, smem_u32(s_ctxt), s_part, pref);
  mv_prefetch<T>(w16 + W16_GV, 256, pre);            // consumed after the glimpse softmax: fully hidden
  __syncthreads();
  for (int e = tid; e < 1024; e += P2_THREADS) {
    const int h = e >> 8, c = e & 255;
    s_qp[h * QP2_STRIDE + c] = s_part[(2 * h) * 256 + c] + s_part[(2 * h + 1) * 256 + c];
  }
  __syncthreads();
  TL(25);
  // (e) per-node and per-tour glimpse scores qp_h . node_emb[n], qp_h . tour_emb[t]: one warp per row; lane owns the
  // column pairs 2 (lane + 32 i), i < 4, of the four query vectors in registers
  {
    float2 qv[4][4];
#pragma unroll
    for (int h = 0; h < 4; ++h)
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[h][i] = *reinterpret_cast<const float2*>(s_qp + h * QP2_STRIDE + 2 * (lane + 32 * i));
    const int nc = s_cnt[2];
    for (int i = warp; i < nc + K; i += P2_WARPS) {
      const int n = i < nc ? (int)s_cand[i] : N + (i - nc);
      float2 e0 = make_float2(0.0f, 0.0f), e1 = e0, e2 = e0, e3 = e0;      // even / odd columns accumulate separately
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 v;
        if (n < N) v = Elem<T>::to2(*reinterpret_cast<const T2*>(s_ne16 + (size_t)n * NE16_STRIDE + (lane + 32 * i) * 4));
        else v = *reinterpret_cast<const float2*>(s_te + (n - N) * 256 + 2 * (lane + 32 * i));
        e0 = __ffma2_rn(v, qv[0][i], e0);
        e1 = __ffma2_rn(v, qv[1][i], e1);
        e2 = __ffma2_rn(v, qv[2][i], e2);
        e3 = __ffma2_rn(v, qv[3][i], e3);
      }
      const float z = warp_sum4(e0.x + e0.y, e1.x + e1.y, e2.x + e2.y, e3.x + e3.y, lane);
      if ((lane & 7) == 0) {
        const int h = ((lane >> 4) << 1) | ((lane >> 3) & 1);
        if (n < N) reinterpret_cast<float*>(s_sn4)[n * 4 + h] = z;
        else s_qt[h * 8 + (n - N)] = z;
      }
    }
  }
  __syncthreads();
  TL(26);
  // (f) glimpse scores per action, softmax per head
  for (int e = tid; e < 4 * A; e += P2_THREADS) {
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
  for (int e = tid; e < 4 * N; e += P2_THREADS) {
    const int n = e >> 2, h = e & 3;
    float acc = 0.0f;
    for (int t = 0; t < K; ++t) {
      const int j = s_pos[t * N + n];
      if (j >= 0) acc += s_sc[h * A + t * k + j];
    }
    reinterpret_cast<float*>(s_pn4)[e] = acc;
  }
  if (tid >= P2_THREADS - 32 && tid < P2_THREADS - 32 + 4 * K) {
    const int e = tid - (P2_THREADS - 32), h = e / K, t = e - h * K;
    float acc = 0.0f;
    for (int j = 0; j < k; ++j) acc += s_sc[h * A + t * k + j];
    s_P[h * 8 + t] = acc;
  }
  __syncthreads();
  TL(27);
  // (g) gbar[h] = sum_a p[h][a] act[a] = sum_t P[h][t] tour_emb[t] + sum_n pn[h][n] node_emb[n]
  {
    const int c2 = tid & 127, half = tid >> 7;
    const uint32_t pn_saddr = smem_u32(s_pn4);
    float2 g0 = make_float2(0.0f, 0.0f), g1 = g0, g2 = g0, g3 = g0;
    const int nc = s_cnt[2];
    for (int i = half; i < nc; i += 2) {
      const int n = s_cand[i];
      const float2 v = Elem<T>::to2(*reinterpret_cast<const T2*>(s_ne16 + (size_t)n * NE16_STRIDE + c2 * 4));
      const float4 pw = lds_f32x4(pn_saddr + (uint32_t)n * 16u);
      g0 = __ffma2_rn(f2(pw.x), v, g0);
      g1 = __ffma2_rn(f2(pw.y), v, g1);
      g2 = __ffma2_rn(f2(pw.z), v, g2);
      g3 = __ffma2_rn(f2(pw.w), v, g3);
    }
    if (half == 0) {
      for (int t = 0; t < K; ++t) {
        const float2 v = *reinterpret_cast<const float2*>(s_te + t * 256 + 2 * c2);
        g0 = __ffma2_rn(f2(s_P[t]), v, g0);
        g1 = __ffma2_rn(f2(s_P[8 + t]), v, g1);
        g2 = __ffma2_rn(f2(s_P[16 + t]), v, g2);
        g3 = __ffma2_rn(f2(s_P[24 + t]), v, g3);
      }
    }
    *reinterpret_cast<float2*>(s_part + (half * 4 + 0) * 256 + 2 * c2) = g0;
    *reinterpret_cast<float2*>(s_part + (half * 4 + 1) * 256 + 2 * c2) = g1;
    *reinterpret_cast<float2*>(s_part + (half * 4 + 2) * 256 + 2 * c2) = g2;
    *reinterpret_cast<float2*>(s_part + (half * 4 + 3) * 256 + 2 * c2) = g3;
  }
  __syncthreads();
  for (int e = tid; e < 1024; e += P2_THREADS) {
    const int h = e >> 8, c = e & 255;
    s_gbar[h * QP2_STRIDE + c] = s_part[h * 256 + c] + s_part[(4 + h) * 256 + c];
  }
  __syncthreads();
  TL(28);
  // (h) glimpse g[o] = W_gv[o] . gbar[head(o)]
  if (mv) matvec2_pipe<T>(w16 + W16_GV, 256, smem_u32(s_gbar) + (uint32_t)((tid & 31) >> 3) * QP2_STRIDE * 4u, s_part, pre);
  mv_prefetch<T>(w16 + W16_M1, 256, pre);
  __syncthreads();
  {
    float v = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) v += s_part[ks * 256 + tid];
    s_g[tid] = v;
  }
  __syncthreads();
  TL(29);
  // (i) lp = W_lk^T out_proj(g) = M1 g
  if (mv) matvec2_pipe<T>(w16 + W16_M1, 256, smem_u32(s_g), s_part, pre);
  __syncthreads();
  {
    float v = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) v += s_part[ks * 256 + tid];
    s_lp[tid] = v;
  }
  __syncthreads();
  TL(30);
  // (j) pointer logits: 10 tanh((lp . node_emb[n] + lp . tour_emb[t]) / 16)
  {
    float2 lv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) lv[i] = *reinterpret_cast<const float2*>(s_lp + 2 * (lane + 32 * i));
    const int nc = s_cnt[2];
    for (int i = warp; i < nc + K; i += P2_WARPS) {
      const int n = i < nc ? (int)s_cand[i] : N + (i - nc);
      float2 d2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 v;
        if (n < N) v = Elem<T>::to2(*reinterpret_cast<const T2*>(s_ne16 + (size_t)n * NE16_STRIDE + (lane + 32 * i) * 4));
        else v = *reinterpret_cast<const float2*>(s_te + (n - N) * 256 + 2 * (lane + 32 * i));
        d2 = __ffma2_rn(v, lv[i], d2);
      }
      float dd = warp_sum(d2.x + d2.y);
      if (lane == 0) {
        if (n < N) s_ln[n] = dd;
        else s_lt[n - N] = dd;
      }
    }
  }
  __syncthreads();
  for (int aa = tid; aa < A; aa += P2_THREADS)
    s_logit[aa] = s_mask[aa] ? -INFINITY : 10.0f * tanhf((s_ln[s_nbh[aa]] + s_lt[aa / k]) * 0.0625f);
  __syncthreads();
  TL(31);
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
#ifdef JAMPR_TIMELINE
      s_tl[32] = (uint32_t)clock64();
      if ((b & 63) == 0 && (b >> 6) < 1024)
        for (int q = 0; q < TL_N; ++q) g_timeline[b >> 6][q] = s_tl[q];
#endif
    }
  }
}

#ifdef JAMPR_TIMELINE
extern "C" __attribute__((visibility("default"))) int jampr_debug_timeline(uint32_t* host_out) {
  CUDA_RET(cudaDeviceSynchronize());
  CUDA_RET(cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(g_timeline)));
  return 0;
}
#endif

static size_t p2_smem_bytes(const jampr_dims_t* d) {
  const int N = d->n_nodes, k = d->k_nbh, K = d->n_slots, A = K * k, L = d->plan_len;
  size_t s = 1024 + 4 * (size_t)p2_tile_stride(N) + P2_R1_BYTES + p2_r2_bytes(N, k, A);
  s += p2_al(32) + p2_al(16) + p2_al(2 * 128 * 8) + p2_al(384 * 4) + p2_al(16) + p2_al(K * L * 2);
  s += 2 * p2_al(4 * JAMPR_MAX_SLOTS * 4) + p2_al(JAMPR_MAX_SLOTS * 4) + p2_al(K * 8 * 4) + p2_al(A * 4) + p2_al(K * N) + p2_al(A) +
       p2_al(JAMPR_MAX_SLOTS * 4) + p2_al(N * 2);
  return s;
}

template <typename T, bool SPLIT>
static int launch_p2(const PtArgs& a, size_t smem, cudaStream_t s) {
  CUDA_RET(cudaFuncSetAttribute(policy_step_tc2_kernel<T, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  policy_step_tc2_kernel<T, SPLIT><<<a.d.n_inst * a.d.n_samples, P2_THREADS, smem, s>>>(a);
  return launch_status();
}

// Returns JAMPR_E_TOOLARGE when the configuration does not fit this tiling (caller falls back to policy_tc.cu).
int jampr_policy_step_tc2(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                          const jampr_state_t* st, int graph_fresh, int step_no, int decode_type, uint64_t seed,
                          const float* uniforms, int store_cache, int split, cudaStream_t s) {
  if (d->n_nodes > 128 || d->n_slots > 4 || d->k_nbh > 127) return JAMPR_E_TOOLARGE;
  if (!w->blob_bf16 || w->n_bf16 < (int64_t)W16_END) return JAMPR_E_BADARG;
  if (!inst->h0_t || !inst->static_emb_t) return JAMPR_E_BADARG;      // column-chunk-major tables (jampr_node_encoder)
  const int N = d->n_nodes, A = d->n_slots * d->k_nbh;
  // the parked embedding (N rows of 528 B) must fit below the decoder scratch that overlays the weight stages
  if ((size_t)N * NE16_STRIDE > 4 * (size_t)p2_tile_stride(N) + 1024) return JAMPR_E_TOOLARGE;
  if (10752 + p2_al(16u * A) + p2_al(4u * A) > p2_r2_bytes(N, d->k_nbh, A)) return JAMPR_E_TOOLARGE;
  const size_t smem = p2_smem_bytes(d);
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
  a.split = split;
#ifdef JAMPR_ABLATE_BUILD
  static const int ablate = [] {
    const char* e = getenv("JAMPR_TC_ABLATE");
    return e ? atoi(e) : 0;
  }();
  a.ablate = ablate;
#else
  a.ablate = 0;
#endif
  if (w->precision == JAMPR_PREC_F16) return split ? launch_p2<__half, true>(a, smem, s) : launch_p2<__half, false>(a, smem, s);
  return split ? launch_p2<__nv_bfloat16, true>(a, smem, s) : launch_p2<__nv_bfloat16, false>(a, smem, s);
}
