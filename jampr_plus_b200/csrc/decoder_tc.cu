// Split decoder of the tensor-core policy step (sm_100a): attn_decoder.py:62-98 + policy.py:203-227 for 128 rollouts
// per CTA.  policy_tc2.cu (split mode) leaves per rollout the 16-bit node-embedding rows (st.ne16), their column mean /
// max (st.dq), the per-tour means of h and the tour features (st.dqp) in HBM; this kernel runs
//
//   stage T   tour_emb[t] (per-tour-mean half) = mean_h[t] W_to2^T      K x GEMM 128 x 256 x 128   (tour_encoder.py:202-213)
//             + per rollout: feature half, query q = [mean + mean, max + max]                   (policy.py:179-182)
//   stage A   ctxt = q W_c^T                       GEMM 128 x 256 x 512          (attn_decoder.py:73)
//   stage B   qp_h = W_gk,h^T ctxt_h  (4 heads)    4 GEMMs 128 x 256 x 64        (set_proj moved to the query side)
//   phase C   per rollout, one warp: glimpse scores qp_h . (node_emb[n_a] + tour_emb[t_a]) over the FEASIBLE actions,
//             masked softmax per head, gbar_h = sum_a p_h[a] act[a]
//   stage D   g[64h..] = W_gv,h gbar_h  (4 heads)  4 GEMMs 128 x 64 x 256        (attn_decoder.py:86-88)
//   stage E   lp = (W_lk^T W_out) g                GEMM 128 x 256 x 256          (out_proj + logit keys, pre-multiplied)
//   phase F   per rollout, one warp: logits 10 tanh(lp . act[a] / 16), masked log-softmax, greedy argmax or
//             inverse-CDF sampling (policy.py:203-227)
//
// The dense stages are tcgen05.mma (M = 128 rollouts, fp32 accumulators in TMEM).  Both operands are SPLIT into a hi and
// a lo 16-bit part (x = hi + lo) and every K step issues hi*hi + lo*hi + hi*lo: ~21 significand bits, i.e. the stages
// are as accurate as fp32 mat-vecs.  That matters: the glimpse scores are O(10) before the softmax, and rounding the
// glimpse-key weights alone to fp16 costs 9e-3 in the log-probabilities (tools/emulate_tc_precision.py).  The tensor
// pipe has the room: 0.85 MFLOP per rollout x 3 is 5 % of the GNN's MMAs.  Weights are read once per 128 rollouts
// (the fused per-rollout decoder streamed 704 KB of them per rollout).
// A operand tiles are written by the threads (from HBM vectors or straight from the previous accumulator in TMEM),
// B tiles are pre-swizzled images streamed by the bulk-copy engine into a double buffer, one item ahead.
#include <stdlib.h>
#include "policy_tc_common.cuh"

#define DT_THREADS 512
#define DT_WARPS (DT_THREADS / 32)
#define DT_M 128
#define DT_BATCH 8             // node-embedding rows in flight per lane in the per-rollout passes
#define DT_BATCH3 16           // same in the logit pass, which keeps few other registers live
#define DT_ITEMS 32            // B tiles consumed after stage T (2 per vehicle slot): 8 (stage A) + 4 (B) + 16 (D) + 4 (E)

struct DtArgs {
  jampr_dims_t d;
  jampr_state_t st;
  const float* w;        // fp32 blob: folded tour-feature branch (OFF_TOF_*)
  const void* w16;
  const float* uniforms;
  uint64_t seed;
  int step_no;
  int decode_type;
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// 16 fp32 values of row r, 16-column group g (0..3) of a K-block -> hi / lo operand tiles (two 16-byte chunks each)
template <typename T>
__device__ __forceinline__ void write_a16(uint8_t* sAhi, uint8_t* sAlo, int r, int g, const float (&v)[16]) {
  using T2 = typename Elem<T>::T2;
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const T2 hh = Elem<T>::from2(v[2 * i], v[2 * i + 1]);
    const float2 hf = Elem<T>::to2(hh);
    h[i] = as_u32(hh);
    l[i] = as_u32(Elem<T>::from2(v[2 * i] - hf.x, v[2 * i + 1] - hf.y));
  }
  const uint32_t o0 = sw128_off((uint32_t)r, (uint32_t)(2 * g)), o1 = sw128_off((uint32_t)r, (uint32_t)(2 * g + 1));
  *reinterpret_cast<uint4*>(sAhi + o0) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(sAhi + o1) = make_uint4(h[4], h[5], h[6], h[7]);
  *reinterpret_cast<uint4*>(sAlo + o0) = make_uint4(l[0], l[1], l[2], l[3]);
  *reinterpret_cast<uint4*>(sAlo + o1) = make_uint4(l[4], l[5], l[6], l[7]);
}

// B tile of item i inside one part of the decoder blob: element offset and element count
__device__ __forceinline__ void dt_item_src(int i, int n_t, uint32_t& off, uint32_t& elems) {
  if (i < n_t) { off = WD_TO + (uint32_t)(i & 1) * WD_TILE; elems = WD_TILE; return; }      // stage T: (slot, K-block) items
  i -= n_t;
  if (i < 8) { off = WD_C + (uint32_t)i * WD_TILE; elems = WD_TILE; }
  else if (i < 12) { off = WD_GK + (uint32_t)(i - 8) * WD_TILE; elems = WD_TILE; }
  else if (i < 28) { off = WD_GV + (uint32_t)(i - 12) * (WD_TILE / 4); elems = WD_TILE / 4; }
  else { off = WD_M1 + (uint32_t)(i - 28) * WD_TILE; elems = WD_TILE; }
}

__host__ __device__ inline uint32_t dt_a32(int A) { return (uint32_t)(A + 31) & ~31u; }
// per-warp scratch: probabilities [A32][4] f32 | logits [A32] f32 | feasible list [A32] i16 | nodes [A32] i16 | qt [16] f32
__host__ __device__ inline uint32_t dt_warp_bytes(int A) { return dt_a32(A) * (16 + 4 + 2 + 2) + 64; }

template <typename T>
__global__ void __launch_bounds__(DT_THREADS, 1) decoder_tc_kernel(const DtArgs a) {
  using T2 = typename Elem<T>::T2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const jampr_dims_t& d = a.d;
  const jampr_state_t& st = a.st;
  if (!jampr_step_active(st.ctl, a.step_no)) return;
  const int N = d.n_nodes, K = d.n_slots, k = d.k_nbh, A = K * k;
  const int bs = d.n_inst * d.n_samples;
  const int b0 = blockIdx.x * DT_M;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t A32 = dt_a32(A);

  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sAhi = base;                    // [128 rows][64 K] hi
  uint8_t* sAlo = base + 16384;            // lo
  uint8_t* sB = base + 32768;              // 2 buffers x (hi 32 KB | lo 32 KB)
  uint8_t* scratch = base + 32768 + 131072;
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch);          // full[0], full[1], mma
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(scratch + 32);
  int* s_any = reinterpret_cast<int*>(scratch + 48);
  uint8_t* wbase = scratch + 64 + (size_t)warp * dt_warp_bytes(A);
  float* s_p = reinterpret_cast<float*>(wbase);                               // [A32][4]
  float* s_logit = reinterpret_cast<float*>(wbase + A32 * 16);                // [A32]
  int16_t* s_list = reinterpret_cast<int16_t*>(wbase + A32 * 20);             // feasible action indices
  int16_t* s_nb = reinterpret_cast<int16_t*>(wbase + A32 * 22);               // node of action a
  float* s_qt = reinterpret_cast<float*>(wbase + A32 * 24);                   // [4 heads][4 slots]
  uint64_t* full = bars;
  uint64_t* bar_mma = bars + 2;

  // ---- does any rollout of this tile need a decision?  (late in the episode whole tiles are idle)
  if (tid == 0) *s_any = 0;
  __syncthreads();
  {
    bool need = false;
    if (tid < DT_M && b0 + tid < bs) need = !jampr_rollout_idle(st, b0 + tid, K);
    if (need) *s_any = 1;
  }
  __syncthreads();
  if (*s_any == 0) {
    if (tid < DT_M && b0 + tid < bs) {
      const size_t b = (size_t)(b0 + tid);
      st.action[2 * b] = 0;
      st.action[2 * b + 1] = 0;
      st.ll[b] = 0.0f;
      st.entropy[b] = 0.0f;
    }
    return;
  }

  const int n_t = 2 * K, n_items = n_t + DT_ITEMS;      // stage T: two K-block items per vehicle slot
  const uint8_t* wdec = reinterpret_cast<const uint8_t*>(a.w16) + (size_t)W16_DEC * 2;
  auto issue_b = [&](int i) {          // one thread
    uint32_t off, elems;
    dt_item_src(i, n_t, off, elems);
    const int buf = i & 1;
    mbar_arrive_expect_tx(&full[buf], elems * 4u);
    bulk_g2s(sB + buf * 65536, wdec + (size_t)off * 2, elems * 2u, &full[buf]);
    bulk_g2s(sB + buf * 65536 + 32768, wdec + ((size_t)WD_PART + off) * 2, elems * 2u, &full[buf]);
  };
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
    issue_b(0);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int q4 = warp & 3, cg = warp >> 2;            // TMEM lane quarter, column group of this warp
  const int r_t = q4 * 32 + lane;                     // accumulator row (= rollout of the tile) owned in TMEM accesses
  const uint32_t ta = tmem + ((uint32_t)(q4 * 32) << 16);
  const uint32_t a_hi = smem_u32(sAhi), a_lo = smem_u32(sAlo), b_addr = smem_u32(sB);
  const uint32_t idesc256 = umma_idesc(128, 256, Elem<T>::fmt), idesc64 = umma_idesc(128, 64, Elem<T>::fmt);

  // A tile from a row-major fp32 matrix in HBM: thread = (row tid / 4, 16-column group tid % 4).  Two halves: the loads
  // of the NEXT item are issued before the wait for the current item's MMAs, the tile is written after it
  auto load_global = [&](float (&v)[16], const float* src, size_t row_stride, int col0) {
    const int r = tid >> 2, g = tid & 3;
    if (b0 + r < bs) {
      const float4* p = reinterpret_cast<const float4*>(src + (size_t)(b0 + r) * row_stride + col0 + g * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 x = p[i];
        v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.0f;
    }
  };
  auto store_tile = [&](const float (&v)[16]) { write_a16<T>(sAhi, sAlo, tid >> 2, tid & 3, v); };
  // A tile from 64 accumulator columns in TMEM: warp = (lane quarter, 16-column group)
  auto fill_from_tmem = [&](uint32_t col0) {
    uint32_t raw[16];
    tmem_ld16_nowait(ta + col0 + (uint32_t)cg * 16u, raw);
    tmem_wait_ld();
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
    write_a16<T>(sAhi, sAlo, r_t, cg, v);
  };
  // accumulator columns [col0, col0 + 256) -> row-major fp32 rows in HBM (dst row stride in floats)
  auto drain_to_global = [&](uint32_t col0, float* dst, size_t row_stride) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t raw[32];
      tmem_ld32_nowait(ta + col0 + (uint32_t)(cg * 64 + c * 32), raw);
      tmem_wait_ld();
      if (b0 + r_t < bs) {
        float4* o = reinterpret_cast<float4*>(dst + (size_t)(b0 + r_t) * row_stride + cg * 64 + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o[i] = make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]), __uint_as_float(raw[4 * i + 2]),
                             __uint_as_float(raw[4 * i + 3]));
      }
    }
  };
  // one item, first half: the A tile is filled; issue the next B load, wait for this one, 12 MMAs (4 K steps x hi*hi,
  // lo*hi, hi*lo), commit.  Second half: every thread waits for the MMAs (A tile and accumulator free / valid).
  auto issue_item = [&](int i, uint32_t d_col, uint32_t idesc, bool first) {
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(&full[i & 1], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t bh = b_addr + (uint32_t)(i & 1) * 65536u, bl = bh + 32768u;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t dah = umma_desc_sw128(a_hi + ks * 32), dal = umma_desc_sw128(a_lo + ks * 32);
        const uint64_t dbh = umma_desc_sw128(bh + ks * 32), dbl = umma_desc_sw128(bl + ks * 32);
        umma_f16(tmem + d_col, dah, dbh, idesc, (first && ks == 0) ? 0u : 1u);
        umma_f16(tmem + d_col, dal, dbh, idesc, 1u);
        umma_f16(tmem + d_col, dah, dbl, idesc, 1u);
      }
      umma_commit(bar_mma);
      if (i + 1 < n_items) issue_b(i + 1);           // off the critical path; its buffer was last read by item i - 1 (complete)
    }
  };
  auto wait_item = [&](int i) {
    mbar_wait(bar_mma, (uint32_t)(i & 1));
    tc_fence_after();
  };
  auto run_item = [&](int i, uint32_t d_col, uint32_t idesc, bool first) {
    issue_item(i, d_col, idesc, first);
    wait_item(i);
  };

  // ================================================================ stage T: per-tour-mean half of the tour embeddings
  // tour_encoder.py:202-213: tour_emb[t] = W_to [input_proj(features_t) | mean_h[t]] + b.  The mean half is a GEMM over
  // the 128 rollouts of the tile per vehicle slot (A = the per-tour means the GNN kernel left in st.dqp[:, t, 0..127]);
  // its raw result goes to st.dte, the feature half (folded 5 x 256 matrix) is added in the pass below.
  {
    float v[16];
    load_global(v, st.dqp, 1024, 0);
#pragma unroll 1
    for (int e = 0; e < n_t; ++e) {
      const int t = e >> 1, kb = e & 1;
      store_tile(v);
      issue_item(e, 0u, idesc256, kb == 0);
      if (e + 1 < n_t) load_global(v, st.dqp + (size_t)((e + 1) >> 1) * 256, 1024, ((e + 1) & 1) * 64);
      wait_item(e);
      if (kb == 1) {
        drain_to_global(0u, st.dte + (size_t)t * 256, 1024);
        tc_fence_before();
      }
    }
  }
  __threadfence_block();
  __syncthreads();
  // ---- per rollout (one warp, lane = 8 columns): tour embeddings and the decoder query (policy.py:179-182)
  // q = [mean_n(node_emb) + mean_t(tour_emb), max_n(node_emb) + max_t(tour_emb)]; st.dq arrives holding the node statistics
  {
  const int c0 = lane * 8;
  float bias[8], wt[5][8];                 // folded tour-feature branch for this lane's columns: the same for every rollout
  {
    const float4* pb = reinterpret_cast<const float4*>(a.w + OFF_TOF_B + c0);
    const float4 x = __ldg(pb), y = __ldg(pb + 1);
    bias[0] = x.x; bias[1] = x.y; bias[2] = x.z; bias[3] = x.w; bias[4] = y.x; bias[5] = y.y; bias[6] = y.z; bias[7] = y.w;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float4* pw = reinterpret_cast<const float4*>(a.w + OFF_TOF_WT + i * 256 + c0);
      const float4 u = __ldg(pw), z = __ldg(pw + 1);
      wt[i][0] = u.x; wt[i][1] = u.y; wt[i][2] = u.z; wt[i][3] = u.w; wt[i][4] = z.x; wt[i][5] = z.y; wt[i][6] = z.z; wt[i][7] = z.w;
    }
  }
#pragma unroll 1
  for (int rr = warp; rr < DT_M; rr += DT_WARPS) {
    const int b = b0 + rr;
    if (b >= bs) break;
    if (jampr_rollout_idle(st, b, K)) continue;
    float sum[8], mx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sum[j] = 0.0f; mx[j] = -INFINITY; }
    for (int t = 0; t < K; ++t) {
      float* pt = st.dte + ((size_t)b * 4 + t) * 256 + c0;
      const float* pf = st.dqp + ((size_t)b * 4 + t) * 256 + 128;
      const float4 x = *reinterpret_cast<const float4*>(pt), y = *reinterpret_cast<const float4*>(pt + 4);
      const float gm[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
      float f[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) f[i] = pf[i];
      float te[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = bias[j];
#pragma unroll
        for (int i = 0; i < 5; ++i) v = fmaf(f[i], wt[i][j], v);
        v += gm[j];
        te[j] = v;
        sum[j] += v;
        mx[j] = fmaxf(mx[j], v);
      }
      *reinterpret_cast<float4*>(pt) = make_float4(te[0], te[1], te[2], te[3]);
      *reinterpret_cast<float4*>(pt + 4) = make_float4(te[4], te[5], te[6], te[7]);
    }
    float* pq = st.dq + (size_t)b * 512 + c0;
    const float4 m0 = *reinterpret_cast<const float4*>(pq), m1 = *reinterpret_cast<const float4*>(pq + 4);
    const float4 x0 = *reinterpret_cast<const float4*>(pq + 256), x1 = *reinterpret_cast<const float4*>(pq + 260);
    // same operation order as the fused kernel: nmean + sum / K, nmax + max
    *reinterpret_cast<float4*>(pq) = make_float4(m0.x + sum[0] / (float)K, m0.y + sum[1] / (float)K, m0.z + sum[2] / (float)K, m0.w + sum[3] / (float)K);
    *reinterpret_cast<float4*>(pq + 4) = make_float4(m1.x + sum[4] / (float)K, m1.y + sum[5] / (float)K, m1.z + sum[6] / (float)K, m1.w + sum[7] / (float)K);
    *reinterpret_cast<float4*>(pq + 256) = make_float4(x0.x + mx[0], x0.y + mx[1], x0.z + mx[2], x0.w + mx[3]);
    *reinterpret_cast<float4*>(pq + 260) = make_float4(x1.x + mx[4], x1.y + mx[5], x1.z + mx[6], x1.w + mx[7]);
  }
  }
  __threadfence_block();
  __syncthreads();

  // ================================================================ stage A: ctxt (TMEM columns 0..255)
  {
    float v[16];
    load_global(v, st.dq, 512, 0);
#pragma unroll 1
    for (int kb = 0; kb < 8; ++kb) {
      store_tile(v);
      issue_item(n_t + kb, 0u, idesc256, kb == 0);
      if (kb + 1 < 8) load_global(v, st.dq, 512, (kb + 1) * 64);
      wait_item(n_t + kb);
    }
  }
  // ================================================================ stage B: qp_h (TMEM columns 256..511) -> st.dqp
#pragma unroll 1
  for (int h = 0; h < 4; ++h) {
    fill_from_tmem((uint32_t)h * 64u);
    run_item(n_t + 8 + h, 256u, idesc256, true);
    drain_to_global(256u, st.dqp + (size_t)h * 256, 1024);
    tc_fence_before();           // the next item's MMAs overwrite these columns after the block barrier in run_item
  }
  __threadfence_block();
  __syncthreads();               // st.dqp rows of this tile are visible to every warp of the CTA (same SM, L1 coherent
                                 // for data written by the CTA itself; reads below use plain loads)

  // ================================================================ phase C: glimpse per rollout (one warp each)
  const T* ne16 = reinterpret_cast<const T*>(st.ne16);
#pragma unroll 1
  for (int rr = warp; rr < DT_M; rr += DT_WARPS) {
    const int b = b0 + rr;
    if (b >= bs) break;
    if (jampr_rollout_idle(st, b, K)) continue;
    // feasible actions in ascending order
    int cnt = 0;
    for (int e0 = 0; e0 < A; e0 += 32) {
      const int e = e0 + lane;
      bool ok = false;
      if (e < A) {
        s_nb[e] = st.nbh[(size_t)b * A + e];
        ok = st.mask[(size_t)b * A + e] == 0;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) s_list[cnt + __popc(m & ((1u << lane) - 1u))] = (int16_t)(e | ((e / k) << 10));      // action | slot << 10
      cnt += __popc(m);
    }
    // per-head glimpse query: lane owns columns 8 lane .. 8 lane + 7
    float2 qv[4][4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float4* p = reinterpret_cast<const float4*>(st.dqp + ((size_t)b * 4 + h) * 256 + lane * 8);
      const float4 x = p[0], y = p[1];
      qv[h][0] = make_float2(x.x, x.y); qv[h][1] = make_float2(x.z, x.w);
      qv[h][2] = make_float2(y.x, y.y); qv[h][3] = make_float2(y.z, y.w);
    }
    for (int t = 0; t < K; ++t) {
      const float4* p = reinterpret_cast<const float4*>(st.dte + ((size_t)b * 4 + t) * 256 + lane * 8);
      const float4 x = p[0], y = p[1];
      const float2 v0 = make_float2(x.x, x.y), v1 = make_float2(x.z, x.w), v2 = make_float2(y.x, y.y), v3 = make_float2(y.z, y.w);
      float dd[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        float2 e = __ffma2_rn(v0, qv[h][0], make_float2(0.0f, 0.0f));
        e = __ffma2_rn(v1, qv[h][1], e);
        e = __ffma2_rn(v2, qv[h][2], e);
        e = __ffma2_rn(v3, qv[h][3], e);
        dd[h] = e.x + e.y;
      }
      const float z = warp_sum4(dd[0], dd[1], dd[2], dd[3], lane);
      if ((lane & 7) == 0) s_qt[(((lane >> 4) << 1) | ((lane >> 3) & 1)) * 4 + t] = z;
    }
    __syncwarp();
    // pass 1: scores of the feasible actions.  The rows come from HBM / L2 (~1 us): DT_BATCH independent 16-byte loads
    // per lane are issued back to back before the first is consumed (Little: ~30 KB in flight per SM are needed)
    const T* rows = ne16 + (size_t)b * N * 256 + lane * 8;
#pragma unroll 1
    for (int i0 = 0; i0 < cnt; i0 += DT_BATCH) {
      uint4 raw[DT_BATCH];
      int ee[DT_BATCH];
#pragma unroll
      for (int j = 0; j < DT_BATCH; ++j) {
        const int i = min(i0 + j, cnt - 1);
        ee[j] = s_list[i];
        raw[j] = *reinterpret_cast<const uint4*>(rows + (size_t)s_nb[ee[j] & 1023] * 256);
      }
#pragma unroll
      for (int j = 0; j < DT_BATCH; ++j) {
        const float2 v0 = Elem<T>::to2(from_u32<T2>(raw[j].x)), v1 = Elem<T>::to2(from_u32<T2>(raw[j].y));
        const float2 v2 = Elem<T>::to2(from_u32<T2>(raw[j].z)), v3 = Elem<T>::to2(from_u32<T2>(raw[j].w));
        float dd[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float2 acc = __ffma2_rn(v0, qv[h][0], make_float2(0.0f, 0.0f));
          acc = __ffma2_rn(v1, qv[h][1], acc);
          acc = __ffma2_rn(v2, qv[h][2], acc);
          acc = __ffma2_rn(v3, qv[h][3], acc);
          dd[h] = acc.x + acc.y;
        }
        const float z = warp_sum4(dd[0], dd[1], dd[2], dd[3], lane);
        if ((lane & 7) == 0 && i0 + j < cnt) {
          const int h = ((lane >> 4) << 1) | ((lane >> 3) & 1);
          s_p[(i0 + j) * 4 + h] = (z + s_qt[h * 4 + (ee[j] >> 10)]) * 0.125f;
        }
      }
    }
    __syncwarp();
    // softmax per head over the feasible actions (masked actions have weight 0)
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float mx = -INFINITY;
      for (int i = lane; i < cnt; i += 32) mx = fmaxf(mx, s_p[i * 4 + h]);
      mx = warp_max(mx);
      float sum = 0.0f;
      for (int i = lane; i < cnt; i += 32) {
        const float ex = expf(s_p[i * 4 + h] - mx);
        s_p[i * 4 + h] = ex;
        sum += ex;
      }
      sum = warp_sum(sum);
      for (int i = lane; i < cnt; i += 32) s_p[i * 4 + h] = s_p[i * 4 + h] / sum;
    }
    __syncwarp();
    // pass 2: gbar_h = sum_a p_h[a] (node_emb[n_a] + tour_emb[t_a])
    float2 g[4][4];
#pragma unroll
    for (int h = 0; h < 4; ++h)
#pragma unroll
      for (int j = 0; j < 4; ++j) g[h][j] = make_float2(0.0f, 0.0f);
    const uint32_t p_saddr = smem_u32(s_p);
    // tour-embedding part first: gbar_h += P_h[t] tour_emb[t], P_h[t] = sum of p_h over the actions of slot t
    for (int t = 0; t < K; ++t) {
      float ps[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int i = lane; i < cnt; i += 32) {
        if ((s_list[i] >> 10) == t) {
          const float4 pw = lds_f32x4(p_saddr + (uint32_t)i * 16u);
          ps[0] += pw.x; ps[1] += pw.y; ps[2] += pw.z; ps[3] += pw.w;
        }
      }
      const float4* p = reinterpret_cast<const float4*>(st.dte + ((size_t)b * 4 + t) * 256 + lane * 8);
      const float4 x = p[0], y = p[1];
      const float2 t0 = make_float2(x.x, x.y), t1 = make_float2(x.z, x.w), t2 = make_float2(y.x, y.y), t3 = make_float2(y.z, y.w);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 p2 = f2(warp_sum(ps[h]));
        g[h][0] = __ffma2_rn(p2, t0, g[h][0]);
        g[h][1] = __ffma2_rn(p2, t1, g[h][1]);
        g[h][2] = __ffma2_rn(p2, t2, g[h][2]);
        g[h][3] = __ffma2_rn(p2, t3, g[h][3]);
      }
    }
#pragma unroll 1
    for (int i0 = 0; i0 < cnt; i0 += DT_BATCH) {
      uint4 raw[DT_BATCH];
#pragma unroll
      for (int j = 0; j < DT_BATCH; ++j) {
        const int i = min(i0 + j, cnt - 1);
        raw[j] = *reinterpret_cast<const uint4*>(rows + (size_t)s_nb[s_list[i] & 1023] * 256);
      }
#pragma unroll
      for (int j = 0; j < DT_BATCH; ++j) {
        if (i0 + j < cnt) {        // warp-uniform
          const float2 v0 = Elem<T>::to2(from_u32<T2>(raw[j].x)), v1 = Elem<T>::to2(from_u32<T2>(raw[j].y));
          const float2 v2 = Elem<T>::to2(from_u32<T2>(raw[j].z)), v3 = Elem<T>::to2(from_u32<T2>(raw[j].w));
          const float4 pw = lds_f32x4(p_saddr + (uint32_t)(i0 + j) * 16u);
          const float ph[4] = {pw.x, pw.y, pw.z, pw.w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 p2 = f2(ph[h]);
            g[h][0] = __ffma2_rn(p2, v0, g[h][0]);
            g[h][1] = __ffma2_rn(p2, v1, g[h][1]);
            g[h][2] = __ffma2_rn(p2, v2, g[h][2]);
            g[h][3] = __ffma2_rn(p2, v3, g[h][3]);
          }
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float4* o = reinterpret_cast<float4*>(st.dgbar + ((size_t)b * 4 + h) * 256 + lane * 8);
      o[0] = make_float4(g[h][0].x, g[h][0].y, g[h][1].x, g[h][1].y);
      o[1] = make_float4(g[h][2].x, g[h][2].y, g[h][3].x, g[h][3].y);
    }
    __syncwarp();
  }
  __threadfence_block();
  __syncthreads();

  // ================================================================ stage D: g (TMEM columns 0..255, 64 per head)
  {
    float v[16];
    load_global(v, st.dgbar, 1024, 0);
#pragma unroll 1
    for (int e = 0; e < 16; ++e) {
      const int h = e >> 2, kb = e & 3;
      store_tile(v);
      issue_item(n_t + 12 + e, (uint32_t)h * 64u, idesc64, kb == 0);
      if (e + 1 < 16) load_global(v, st.dgbar + (size_t)((e + 1) >> 2) * 256, 1024, ((e + 1) & 3) * 64);
      wait_item(n_t + 12 + e);
    }
  }
  // ================================================================ stage E: lp (TMEM columns 256..511) -> st.dlp
#pragma unroll 1
  for (int kb = 0; kb < 4; ++kb) {
    fill_from_tmem((uint32_t)kb * 64u);
    run_item(n_t + 28 + kb, 256u, idesc256, kb == 0);
  }
  drain_to_global(256u, st.dlp, 256);
  tc_fence_before();
  __threadfence_block();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }

  // ================================================================ phase F: logits + selection per rollout
#pragma unroll 1
  for (int rr = warp; rr < DT_M; rr += DT_WARPS) {
    const int b = b0 + rr;
    if (b >= bs) break;
    if (jampr_rollout_idle(st, b, K)) {
      if (lane == 0) {
        st.action[2 * (size_t)b] = 0;
        st.action[2 * (size_t)b + 1] = 0;
        st.ll[b] = 0.0f;
        st.entropy[b] = 0.0f;
      }
      continue;
    }
    int cnt = 0;
    for (int e0 = 0; e0 < A; e0 += 32) {
      const int e = e0 + lane;
      bool ok = false;
      if (e < A) {
        s_nb[e] = st.nbh[(size_t)b * A + e];
        ok = st.mask[(size_t)b * A + e] == 0;
        s_logit[e] = -INFINITY;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) s_list[cnt + __popc(m & ((1u << lane) - 1u))] = (int16_t)(e | ((e / k) << 10));
      cnt += __popc(m);
    }
    float2 lv[4];
    {
      const float4* p = reinterpret_cast<const float4*>(st.dlp + (size_t)b * 256 + lane * 8);
      const float4 x = p[0], y = p[1];
      lv[0] = make_float2(x.x, x.y); lv[1] = make_float2(x.z, x.w); lv[2] = make_float2(y.x, y.y); lv[3] = make_float2(y.z, y.w);
    }
    float lt[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int t = 0; t < K; ++t) {
      const float4* p = reinterpret_cast<const float4*>(st.dte + ((size_t)b * 4 + t) * 256 + lane * 8);
      const float4 x = p[0], y = p[1];
      float2 e = __ffma2_rn(make_float2(x.x, x.y), lv[0], make_float2(0.0f, 0.0f));
      e = __ffma2_rn(make_float2(x.z, x.w), lv[1], e);
      e = __ffma2_rn(make_float2(y.x, y.y), lv[2], e);
      e = __ffma2_rn(make_float2(y.z, y.w), lv[3], e);
      const float z = warp_sum(e.x + e.y);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j == t) lt[j] = z;
    }
    __syncwarp();
    const T* rows = ne16 + (size_t)b * N * 256 + lane * 8;
#pragma unroll 1
    for (int i0 = 0; i0 < cnt; i0 += DT_BATCH3) {
      uint4 raw[DT_BATCH3];
      int ee[DT_BATCH3];
#pragma unroll
      for (int j = 0; j < DT_BATCH3; ++j) {
        const int i = min(i0 + j, cnt - 1);
        ee[j] = s_list[i];
        raw[j] = *reinterpret_cast<const uint4*>(rows + (size_t)s_nb[ee[j] & 1023] * 256);
      }
#pragma unroll
      for (int j = 0; j < DT_BATCH3; ++j) {
        float2 acc = __ffma2_rn(Elem<T>::to2(from_u32<T2>(raw[j].x)), lv[0], make_float2(0.0f, 0.0f));
        acc = __ffma2_rn(Elem<T>::to2(from_u32<T2>(raw[j].y)), lv[1], acc);
        acc = __ffma2_rn(Elem<T>::to2(from_u32<T2>(raw[j].z)), lv[2], acc);
        acc = __ffma2_rn(Elem<T>::to2(from_u32<T2>(raw[j].w)), lv[3], acc);
        const float z = warp_sum(acc.x + acc.y);
        if (lane == 0 && i0 + j < cnt) s_logit[ee[j] & 1023] = z;          // raw dot product; the clip follows, vectorised
      }
    }
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) {
      const int v = s_list[i], e = v & 1023, t = v >> 10;
      const float ltv = (t == 0) ? lt[0] : (t == 1) ? lt[1] : (t == 2) ? lt[2] : lt[3];
      s_logit[e] = 10.0f * tanhf((s_logit[e] + ltv) * 0.0625f);
    }
    __syncwarp();
    // ---- masked log-softmax + selection, same rules as decoder.cu / policy_tc2.cu
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
    __syncwarp();
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
          const float tt = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += tt;
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
      st.action[2 * (size_t)b + 1] = s_nb[choice];
      st.ll[b] = s_logit[choice];
      st.entropy[b] = ent;
      if (nan) atomicOr(st.status + b, JAMPR_ST_NAN_LOGITS);
    }
    __syncwarp();
  }
}

static size_t dt_smem_bytes(const jampr_dims_t* d) {
  return 1024 + 32768 + 131072 + 64 + (size_t)DT_WARPS * dt_warp_bytes(d->n_slots * d->k_nbh);
}

template <typename T>
static int launch_dt(const DtArgs& a, size_t smem, cudaStream_t s) {
  CUDA_RET(cudaFuncSetAttribute(decoder_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int bs = a.d.n_inst * a.d.n_samples;
  decoder_tc_kernel<T><<<(bs + DT_M - 1) / DT_M, DT_THREADS, smem, s>>>(a);
  return launch_status();
}

// Returns JAMPR_E_TOOLARGE when the configuration does not fit (caller keeps the fused decoder).
int jampr_decoder_tc(const jampr_dims_t* d, const jampr_weights_t* w, const jampr_state_t* st, int step_no,
                     int decode_type, uint64_t seed, const float* uniforms, cudaStream_t s) {
  if (d->n_slots > 4 || d->n_slots * d->k_nbh > 1023) return JAMPR_E_TOOLARGE;
  if (!st->ne16 || !st->dq || !st->dte || !st->dqp || !st->dgbar || !st->dlp) return JAMPR_E_BADARG;
  if (!w->blob_bf16 || w->n_bf16 < (int64_t)W16_END) return JAMPR_E_BADARG;
  const size_t smem = dt_smem_bytes(d);
  if (smem > 227 * 1024) return JAMPR_E_TOOLARGE;
  DtArgs a;
  a.d = *d;
  a.st = *st;
  a.w = w->blob;
  a.w16 = w->blob_bf16;
  a.uniforms = uniforms;
  a.seed = seed;
  a.step_no = step_no;
  a.decode_type = decode_type;
  if (w->precision == JAMPR_PREC_F16) return launch_dt<__half>(a, smem, s);
  return launch_dt<__nv_bfloat16>(a, smem, s);
}

int jampr_decoder_tc_fits(const jampr_dims_t* d) {
  return d->n_slots <= 4 && d->n_slots * d->k_nbh <= 1023 && dt_smem_bytes(d) <= 227 * 1024;      // action index: 10 bits
}
