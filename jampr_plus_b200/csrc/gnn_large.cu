// Large-graph tensor-core path (N > 128, e.g. the 1000-customer Gehring-Homberger instances of BASELINE config 4).
//
// A graph no longer fits one 128-row UMMA tile, so the GNN runs layer by layer with the node embedding in HBM/L2:
// one launch per GraphConv block (kernel boundary = the grid-wide dependency between layers), grid = (row tiles,
// graphs), one CTA per 128-row tile.  Per CTA the block is the same as in policy_tc2.cu — lin_root(h) on tcgen05
// while the max-aggregation runs, lin_rel(agg) after it, weights streamed as pre-swizzled tile images by bulk
// copies through two 16 KB stages, two-pass GELU / skip / LayerNorm epilogue on the packed fp32 pipe — except that
// neighbour rows are gathered from the 16-bit embedding of the previous layer in global memory (ping-pong buffers)
// and the fp32 skip operand lives in a global buffer.  The same kernel serves the node encoder (graph = instance,
// node_encoder.py:128-155) and the per-step tour-graph encoder (graph = rollout, tour_encoder.py:159-198).
//   lg_proj_kernel      node_emb = h W_o^T + b + static_emb (tour_encoder.py:217-218) or static_emb = h W^T + b
//                       (node_encoder.py:147), fp32 to HBM, + per-tile column sums / maxima
//   lg_h0_kernel        h0 = node_emb_input_proj(static_emb) (tour_encoder.py:170-172), K = 256
// The decoder of this path is the fp32 decoder.cu kernel (it reads h_fin / ne / ne_stats from HBM).
#include "policy_tc_common.cuh"

#define LG_THREADS 256
#define IMG_NE_BASE 72     // first node-encoder weight stage in the 16-bit blob (W16_NE / W16_STAGE)

struct LgArgs {
  jampr_dims_t d;
  jampr_instances_t p;
  jampr_state_t st;
  const float* w;        // fp32 blob
  const void* w16;       // 16-bit blob
  const void* h_src;     // (G, Npad, 128) 16-bit, previous layer
  void* h_dst;           // (G, Npad, 128) 16-bit, this layer's output
  float* hskip;          // (G, Npad, 128) fp32 skip operand, updated in place
  float* hfin;           // (G, N, 128) fp32 copy of the output, or NULL
  int img;               // first weight stage of this layer (4 stages: K-blocks 0..3)
  int par_off;           // fp32 blob offset of [bias 128 | gamma 128 | beta 128]
  int kind;              // 0 static kNN graph, 1 tour graph forward (pred -> node), 2 reversed (succ -> node)
  int per_instance;      // 1: graph index = instance (node encoder); 0: graph index = rollout
  int step_no;
  int mode;              // lg_proj: 0 node_emb (+ static_emb, stats), 1 static_emb
  int bias_off;          // lg_proj / lg_h0: fp32 blob offset of the output bias
  int tab_shift;         // kNN edge entries: source row = (entry & 0xffff) >> tab_shift (7: the single-tile encoding of
                         // policy_tc_common.cuh agg_entry, shared with the step kernel; 0: plain row index, N > 128)
};

template <typename T>
__device__ __forceinline__ void lg_load_w(uint8_t* sW, int slot, const void* w16, int img, uint64_t* full) {
  mbar_arrive_expect_tx(&full[slot], 16384u);
  bulk_g2s(sW + slot * 16384, reinterpret_cast<const uint8_t*>(w16) + (size_t)img * W16_STAGE * 2, 16384u, &full[slot]);
}

template <typename T>
__device__ __forceinline__ typename Elem<T>::T2 t2_lowest();
template <>
__device__ __forceinline__ __half2 t2_lowest<__half>() { return from_u32<__half2>(0xFBFFFBFFu); }
template <>
__device__ __forceinline__ __nv_bfloat162 t2_lowest<__nv_bfloat16>() { return from_u32<__nv_bfloat162>(0xFF7FFF7Fu); }

// one edge: weight * 32 bytes (chunk pair) of source row `src` of the global 16-bit embedding
template <typename T>
__device__ __forceinline__ void lg_edge(const T* __restrict__ hs, uint32_t src, typename Elem<T>::T2 w2, int foff,
                                        typename Elem<T>::T2 (&m)[8], bool first) {
  using T2 = typename Elem<T>::T2;
  const uint4* gp = reinterpret_cast<const uint4*>(hs + (size_t)src * 128 + foff);
  const uint4 va = __ldg(gp), vb = __ldg(gp + 1);
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
__device__ __forceinline__ typename Elem<T>::T2 t2_bcast_bits(uint16_t bits) {
  return from_u32<typename Elem<T>::T2>((uint32_t)bits | ((uint32_t)bits << 16));
}

// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LG_THREADS, 2) lg_layer_kernel(const LgArgs a) {
  using T2 = typename Elem<T>::T2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const jampr_dims_t& d = a.d;
  const jampr_instances_t& p = a.p;
  const jampr_state_t& st = a.st;
  if (!a.per_instance) {
    if (!jampr_step_active(st.ctl, a.step_no)) return;
    if (jampr_rollout_idle(st, blockIdx.y, d.n_slots)) return;
  }
  const int N = d.n_nodes, k = d.k_nbh, k4 = (k + 3) & ~3;
  const int tile = blockIdx.x, g = blockIdx.y, npad = gridDim.x * 128;
  const int ins = a.per_instance ? g : g / d.n_samples;
  const int row0 = tile * 128, nrows = min(128, N - row0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* __restrict__ hs = reinterpret_cast<const T*>(a.h_src) + (size_t)g * npad * 128;
  T* __restrict__ hd = reinterpret_cast<T*>(a.h_dst) + (size_t)g * npad * 128;
  float* __restrict__ sk = a.hskip + (size_t)g * npad * 128;

  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;                    // tiles: agg K-blocks 0,1 | own h K-blocks 2,3 (16 KB each)
  uint8_t* sW = base + 65536;            // two weight stages
  uint8_t* s_dep = base + 98304;         // depot-row partial maxima: [32 groups][8 chunk pairs][32 B]
  float* s_par = reinterpret_cast<float*>(base + 106496);     // bias | gamma | beta
  float2* s_red = reinterpret_cast<float2*>(base + 108032);    // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 110080);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + 110112);
  uint64_t* full = bars;
  uint64_t* bar_root = bars + 2;
  uint64_t* bar_acc = bars + 3;

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(bar_root, 1);
    mbar_init(bar_acc, 1);
    mbar_fence_init();
    lg_load_w<T>(sW, 0, a.w16, a.img + 2, full);      // lin_root pair
    lg_load_w<T>(sW, 1, a.w16, a.img + 3, full);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  if (tid >= 32 && tid < 128)
    reinterpret_cast<float4*>(s_par)[tid - 32] = __ldg(reinterpret_cast<const float4*>(a.w + a.par_off) + (tid - 32));
  // own rows of the previous embedding -> h tiles (K-blocks 2,3); rows >= nrows and the agg tiles start as zero
  for (int e = tid; e < 128 * 16; e += LG_THREADS) {
    const int r = e >> 4, cc = e & 15;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < nrows) v = __ldg(reinterpret_cast<const uint4*>(hs + (size_t)(row0 + r) * 128 + cc * 8));
    *reinterpret_cast<uint4*>(sA + (2 + (cc >> 3)) * 16384 + sw128_off(r, cc & 7)) = v;
  }
  for (int e = tid; e < 32768 / 16; e += LG_THREADS) reinterpret_cast<uint4*>(sA)[e] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = umma_idesc(128, 128, Elem<T>::fmt);
  const uint32_t a_addr = smem_u32(sA), w_addr = smem_u32(sW);
  if (tid == 0) {
    mbar_wait(&full[0], 0u);
    mbar_wait(&full[1], 0u);
    tc_fence_after();
    mma_kblock(tmem, a_addr + 2 * 16384, w_addr, idesc, true);
    mma_kblock(tmem, a_addr + 3 * 16384, w_addr + 16384, idesc, false);
    umma_commit(bar_root);
    mbar_wait(bar_root, 0u);
    lg_load_w<T>(sW, 0, a.w16, a.img + 0, full);      // lin_rel pair
    lg_load_w<T>(sW, 1, a.w16, a.img + 1, full);
  }

  // ---------------------------------------------------------------- max aggregation (sources gathered from L2)
  const int grp = tid >> 3, cp = tid & 7;             // 32 row groups x 8 chunk pairs (32 bytes of features each)
  const int foff = cp * 16;                           // first feature of the pair
  const uint32_t kb_off = (uint32_t)(cp >> 2) * 16384u, cA = (uint32_t)(cp & 3) * 2u;
  const size_t gN = (size_t)g * N, iN = (size_t)ins * N;
  const int16_t* lnk = a.kind == 1 ? st.pred : st.succ;
  const int16_t* oth = a.kind == 1 ? st.succ : st.pred;
  for (int r = grp; r < nrows; r += 32) {
    const int u = row0 + r;
    if (u == 0) continue;                             // the depot row is reduced cooperatively below
    T2 m[8];
    if (a.kind == 0) {
      const uint4* et = reinterpret_cast<const uint4*>(p.tc_tab + (iN + u) * k4);
      for (int j4 = 0; j4 < k4 / 4; ++j4) {
        const uint4 q = __ldg(et + j4);
        lg_edge<T>(hs, (q.x & 0xffffu) >> a.tab_shift, from_u32<T2>(__byte_perm(q.x, q.x, 0x3232)), foff, m, j4 == 0);
        lg_edge<T>(hs, (q.y & 0xffffu) >> a.tab_shift, from_u32<T2>(__byte_perm(q.y, q.y, 0x3232)), foff, m, false);
        lg_edge<T>(hs, (q.z & 0xffffu) >> a.tab_shift, from_u32<T2>(__byte_perm(q.z, q.z, 0x3232)), foff, m, false);
        lg_edge<T>(hs, (q.w & 0xffffu) >> a.tab_shift, from_u32<T2>(__byte_perm(q.w, q.w, 0x3232)), foff, m, false);
      }
    } else {
      if (st.visited[gN + u]) {
        const int src = lnk[gN + u];
        lg_edge<T>(hs, (uint32_t)src, t2_bcast_bits<T>(bits_T<T>(p.dist[(iN + u) * N + src])), foff, m, true);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = Elem<T>::bcast(0.0f);
      }
    }
    *reinterpret_cast<uint4*>(sA + kb_off + sw128_off(r, cA)) = make_uint4(as_u32(m[0]), as_u32(m[1]), as_u32(m[2]), as_u32(m[3]));
    *reinterpret_cast<uint4*>(sA + kb_off + sw128_off(r, cA + 1)) = make_uint4(as_u32(m[4]), as_u32(m[5]), as_u32(m[6]), as_u32(m[7]));
  }
  if (tile == 0) {
    // depot row: receives from every node (kNN graph, weight = time to depot) or from the tour ends (tour graph)
    T2 m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = t2_lowest<T>();
    int have = 0;
    for (int u = grp; u < N; u += 32) {
      bool take = a.kind == 0;
      if (!take) take = u > 0 && st.visited[gN + u] && oth[gN + u] == 0;
      if (take) {
        lg_edge<T>(hs, (uint32_t)u, t2_bcast_bits<T>(bits_T<T>(p.ttd[iN + u])), foff, m, false);
        have = 1;
      }
    }
    uint4* dp = reinterpret_cast<uint4*>(s_dep + (grp * 8 + cp) * 32);
    dp[0] = make_uint4(as_u32(m[0]), as_u32(m[1]), as_u32(m[2]), as_u32(m[3]));
    dp[1] = make_uint4(as_u32(m[4]), as_u32(m[5]), as_u32(m[6]), as_u32(m[7]));
    const int any = __syncthreads_or(have);
    if (tid < 8) {
      T2 r8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) r8[i] = t2_lowest<T>();
      for (int gq = 0; gq < 32; ++gq) {
        const uint4* q = reinterpret_cast<const uint4*>(s_dep + (gq * 8 + tid) * 32);
        const uint4 x = q[0], y = q[1];
        r8[0] = __hmax2(r8[0], from_u32<T2>(x.x)); r8[1] = __hmax2(r8[1], from_u32<T2>(x.y));
        r8[2] = __hmax2(r8[2], from_u32<T2>(x.z)); r8[3] = __hmax2(r8[3], from_u32<T2>(x.w));
        r8[4] = __hmax2(r8[4], from_u32<T2>(y.x)); r8[5] = __hmax2(r8[5], from_u32<T2>(y.y));
        r8[6] = __hmax2(r8[6], from_u32<T2>(y.z)); r8[7] = __hmax2(r8[7], from_u32<T2>(y.w));
      }
      uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
      if (any) {
        o0 = make_uint4(as_u32(r8[0]), as_u32(r8[1]), as_u32(r8[2]), as_u32(r8[3]));
        o1 = make_uint4(as_u32(r8[4]), as_u32(r8[5]), as_u32(r8[6]), as_u32(r8[7]));
      }
      const uint32_t kb2 = (uint32_t)(tid >> 2) * 16384u, c2 = (uint32_t)(tid & 3) * 2u;
      *reinterpret_cast<uint4*>(sA + kb2 + sw128_off(0, c2)) = o0;
      *reinterpret_cast<uint4*>(sA + kb2 + sw128_off(0, c2 + 1)) = o1;
    }
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(&full[0], 1u);
    mbar_wait(&full[1], 1u);
    tc_fence_after();
    mma_kblock(tmem, a_addr, w_addr, idesc, false);
    mma_kblock(tmem, a_addr + 16384, w_addr + 16384, idesc, false);
    umma_commit(bar_acc);
    mbar_wait(bar_acc, 0u);
  }
  __syncthreads();
  tc_fence_after();

  // ---------------------------------------------------------------- epilogue: GELU + skip -> LayerNorm -> HBM
  {
    const int q = warp & 3, ch = warp >> 2, r = q * 32 + lane;
    const bool live = r < nrows;
    const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t par_saddr = smem_u32(s_par);
    // fp32 skip operand in tile-chunk-row order: float4 ((tile * 32 + chunk) * 128 + row): thread = row accesses are
    // contiguous across a warp (row-major they were 32 different lines per warp instruction)
    float4* sk4 = reinterpret_cast<float4*>(sk) + (size_t)tile * 32 * 128 + r;
    float2 s2 = make_float2(0.0f, 0.0f), q2 = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col0 = ch * 64 + c * 32;
      uint32_t acc[32];
      tmem_ld32_nowait(ta + col0, acc);
      float4 sv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) sv[i] = live ? sk4[(size_t)((col0 >> 2) + i) * 128] : make_float4(0.f, 0.f, 0.f, 0.f);
      tmem_wait_ld();
      const uint32_t pb = par_saddr + (uint32_t)col0 * 4u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = lds_f32x4(pb + i * 16);
        float2 y0 = gelu2(__fadd2_rn(make_float2(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1])), make_float2(b4.x, b4.y)));
        float2 y1 = gelu2(__fadd2_rn(make_float2(__uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3])), make_float2(b4.z, b4.w)));
        y0 = __fadd2_rn(y0, make_float2(sv[i].x, sv[i].y));
        y1 = __fadd2_rn(y1, make_float2(sv[i].z, sv[i].w));
        s2 = __fadd2_rn(s2, __fadd2_rn(y0, y1));
        q2 = __ffma2_rn(y0, y0, q2);
        q2 = __ffma2_rn(y1, y1, q2);
        acc[4 * i] = __float_as_uint(y0.x); acc[4 * i + 1] = __float_as_uint(y0.y);
        acc[4 * i + 2] = __float_as_uint(y1.x); acc[4 * i + 3] = __float_as_uint(y1.y);
      }
      tmem_st32(ta + col0, acc);
    }
    tmem_wait_st();
    s_red[ch * 128 + r] = make_float2(s2.x + s2.y, q2.x + q2.y);
    __syncthreads();
    const float2 v0 = s_red[r], v1 = s_red[128 + r];
    const float mean = (v0.x + v1.x) * (1.0f / 128.0f);
    const float var = fmaxf(fmaf(-mean, mean, (v0.y + v1.y) * (1.0f / 128.0f)), 0.0f);
    const float2 rstd2 = f2(rsqrtf(var + 1e-5f)), nmean2 = f2(-mean);
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
        const float2 d0 = __fadd2_rn(make_float2(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1])), nmean2);
        const float2 d1 = __fadd2_rn(make_float2(__uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3])), nmean2);
        const float2 o0 = __ffma2_rn(d0, __fmul2_rn(make_float2(g4.x, g4.y), rstd2), make_float2(e4.x, e4.y));
        const float2 o1 = __ffma2_rn(d1, __fmul2_rn(make_float2(g4.z, g4.w), rstd2), make_float2(e4.z, e4.w));
        acc[4 * i] = __float_as_uint(o0.x); acc[4 * i + 1] = __float_as_uint(o0.y);
        acc[4 * i + 2] = __float_as_uint(o1.x); acc[4 * i + 3] = __float_as_uint(o1.y);
      }
      if (live) {
        const size_t ro = (size_t)(row0 + r) * 128 + col0;
        uint4* d16 = reinterpret_cast<uint4*>(hd + ro);
        float4* dfin = a.hfin ? reinterpret_cast<float4*>(a.hfin + ((size_t)g * N + row0 + r) * 128 + col0) : nullptr;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            pk[e] = as_u32(Elem<T>::from2(__uint_as_float(acc[i * 8 + 2 * e]), __uint_as_float(acc[i * 8 + 2 * e + 1])));
          d16[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                                       __uint_as_float(acc[4 * i + 3]));
          sk4[(size_t)((col0 >> 2) + i) * 128] = v;
          if (dfin) dfin[i] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// out (N x 256) = h W^T + bias (+ static_emb): A = own h rows (K-blocks 2,3 of the tile layout), B = four stages
// (rows 0..127 x K-blocks 0,1, then rows 128..255).  mode 0 also reduces column sum / max over the tile's rows.
template <typename T>
__global__ void __launch_bounds__(LG_THREADS, 2) lg_proj_kernel(const LgArgs a, float* __restrict__ out, float* __restrict__ stats_part,
                                                                T* __restrict__ out16) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const jampr_dims_t& d = a.d;
  const jampr_instances_t& p = a.p;
  if (!a.per_instance) {
    if (!jampr_step_active(a.st.ctl, a.step_no)) return;
    if (jampr_rollout_idle(a.st, blockIdx.y, d.n_slots)) return;
  }
  const int N = d.n_nodes;
  const int tile = blockIdx.x, g = blockIdx.y, npad = gridDim.x * 128;
  const int ins = a.per_instance ? g : g / d.n_samples;
  const int row0 = tile * 128, nrows = min(128, N - row0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* __restrict__ hs = reinterpret_cast<const T*>(a.h_src) + (size_t)g * npad * 128;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;                    // two h tiles
  uint8_t* sW = base + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + 65568);
  uint64_t* full = bars;
  uint64_t* bar_root = bars + 2;
  uint64_t* bar_acc = bars + 3;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(bar_root, 1);
    mbar_init(bar_acc, 1);
    mbar_fence_init();
    lg_load_w<T>(sW, 0, a.w16, a.img + 0, full);
    lg_load_w<T>(sW, 1, a.w16, a.img + 1, full);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  for (int e = tid; e < 128 * 16; e += LG_THREADS) {
    const int r = e >> 4, cc = e & 15;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < nrows) v = __ldg(reinterpret_cast<const uint4*>(hs + (size_t)(row0 + r) * 128 + cc * 8));
    *reinterpret_cast<uint4*>(sA + (cc >> 3) * 16384 + sw128_off(r, cc & 7)) = v;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc(128, 128, Elem<T>::fmt);
    const uint32_t a_addr = smem_u32(sA), w_addr = smem_u32(sW);
    mbar_wait(&full[0], 0u);
    mbar_wait(&full[1], 0u);
    tc_fence_after();
    mma_kblock(tmem, a_addr, w_addr, idesc, true);
    mma_kblock(tmem, a_addr + 16384, w_addr + 16384, idesc, false);
    umma_commit(bar_root);
    mbar_wait(bar_root, 0u);
    lg_load_w<T>(sW, 0, a.w16, a.img + 2, full);
    lg_load_w<T>(sW, 1, a.w16, a.img + 3, full);
    mbar_wait(&full[0], 1u);
    mbar_wait(&full[1], 1u);
    tc_fence_after();
    mma_kblock(tmem + 128, a_addr, w_addr, idesc, true);
    mma_kblock(tmem + 128, a_addr + 16384, w_addr + 16384, idesc, false);
    umma_commit(bar_acc);
    mbar_wait(bar_acc, 0u);
  }
  __syncthreads();
  tc_fence_after();
  // Epilogue.  tcgen05.ld hands a thread 32 columns of ITS row; loading static_emb and storing the rows from there would
  // be 32 different 128-byte lines per warp instruction (LSU-bound).  Each warp turns its 32 x 32 block through 4 KB of
  // shared memory (the operand tiles are dead; 16-byte chunks XOR-ed with the row: conflict-free both ways), after
  // which 8 lanes cover 128 contiguous bytes of one row for the static_emb load and the fp32 / 16-bit stores.
  const int q = warp & 3, ch = warp >> 2;
  const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16);
  uint8_t* tt = sA + (size_t)warp * 4096;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int col0 = ch * 128 + c * 32;
    uint32_t acc[32];
    tmem_ld32_nowait(ta + col0, acc);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<uint4*>(tt + lane * 128 + ((i ^ (lane & 7)) << 4)) = make_uint4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    __syncwarp();
    const int cc = lane & 7, col = col0 + 4 * cc;
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.w + a.bias_off + col));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int rw = 4 * j + (lane >> 3), r = q * 32 + rw;
      if (r < nrows) {
        const uint4 x = *reinterpret_cast<const uint4*>(tt + rw * 128 + ((cc ^ (rw & 7)) << 4));
        float4 v = make_float4(__uint_as_float(x.x) + b4.x, __uint_as_float(x.y) + b4.y, __uint_as_float(x.z) + b4.z,
                               __uint_as_float(x.w) + b4.w);
        const size_t u = (size_t)row0 + r;
        if (a.mode == 0) {
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.static_emb + ((size_t)ins * N + u) * 256 + col));
          v.x += s4.x; v.y += s4.y; v.z += s4.z; v.w += s4.w;
        }
        *reinterpret_cast<float4*>(out + ((size_t)g * N + u) * 256 + col) = v;
        // 16-bit copy of the rows for the 128-rollout decoder (decoder_tc.cu gathers the candidate rows from it)
        if (out16)
          *reinterpret_cast<uint2*>(out16 + ((size_t)g * N + u) * 256 + col) =
              make_uint2(as_u32(Elem<T>::from2(v.x, v.y)), as_u32(Elem<T>::from2(v.z, v.w)));
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
  if (a.mode == 0 && stats_part) {
    // column sum / max over this tile's rows (read back through L2: the rows were written by this CTA)
    float s = 0.0f, m = -INFINITY;
    const float* col = out + ((size_t)g * N + row0) * 256 + tid;
    for (int rr = 0; rr < nrows; ++rr) {
      const float v = __ldcg(col + (size_t)rr * 256);
      s += v;
      m = fmaxf(m, v);
    }
    float* sp = stats_part + ((size_t)g * gridDim.x + tile) * 512;
    sp[tid] = s;
    sp[256 + tid] = m;
  }
}

// column mean / max of node_emb over all nodes (policy.py:179-181) from the per-tile partials
__global__ void lg_stats_kernel(jampr_dims_t d, jampr_state_t st, const float* __restrict__ stats_part, int nt, int step_no) {
  if (!jampr_step_active(st.ctl, step_no)) return;
  const int g = blockIdx.x, c = threadIdx.x;
  float s = 0.0f, m = -INFINITY;
  for (int t = 0; t < nt; ++t) {
    s += stats_part[((size_t)g * nt + t) * 512 + c];
    m = fmaxf(m, stats_part[((size_t)g * nt + t) * 512 + 256 + c]);
  }
  st.ne_stats[(size_t)g * 512 + c] = s / (float)d.n_nodes;
  st.ne_stats[(size_t)g * 512 + 256 + c] = m;
}

// h0 = static_emb W_in^T + b (K = 256): A = own static_emb rows as four 16-bit K-block tiles
template <typename T>
__global__ void __launch_bounds__(LG_THREADS, 2) lg_h0_kernel(const LgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const jampr_dims_t& d = a.d;
  const jampr_instances_t& p = a.p;
  const int N = d.n_nodes;
  const int tile = blockIdx.x, ins = blockIdx.y;
  const int row0 = tile * 128, nrows = min(128, N - row0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;                    // four tiles (K = 256)
  uint8_t* sW = base + 65536;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 98304);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + 98336);
  uint64_t* full = bars;
  uint64_t* bar_root = bars + 2;
  uint64_t* bar_acc = bars + 3;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(bar_root, 1);
    mbar_init(bar_acc, 1);
    mbar_fence_init();
    lg_load_w<T>(sW, 0, a.w16, a.img + 0, full);
    lg_load_w<T>(sW, 1, a.w16, a.img + 1, full);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  for (int e = tid; e < 128 * 32; e += LG_THREADS) {
    const int r = e >> 5, cc = e & 31;               // 32 chunks of 8 features
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < nrows) {
      const float4* src = reinterpret_cast<const float4*>(p.static_emb + ((size_t)ins * N + row0 + r) * 256 + cc * 8);
      const float4 x = __ldg(src), y = __ldg(src + 1);
      v = make_uint4(as_u32(Elem<T>::from2(x.x, x.y)), as_u32(Elem<T>::from2(x.z, x.w)), as_u32(Elem<T>::from2(y.x, y.y)),
                     as_u32(Elem<T>::from2(y.z, y.w)));
    }
    *reinterpret_cast<uint4*>(sA + (cc >> 3) * 16384 + sw128_off(r, cc & 7)) = v;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc(128, 128, Elem<T>::fmt);
    const uint32_t a_addr = smem_u32(sA), w_addr = smem_u32(sW);
    mbar_wait(&full[0], 0u);
    mbar_wait(&full[1], 0u);
    tc_fence_after();
    mma_kblock(tmem, a_addr, w_addr, idesc, true);
    mma_kblock(tmem, a_addr + 16384, w_addr + 16384, idesc, false);
    umma_commit(bar_root);
    mbar_wait(bar_root, 0u);
    lg_load_w<T>(sW, 0, a.w16, a.img + 2, full);
    lg_load_w<T>(sW, 1, a.w16, a.img + 3, full);
    mbar_wait(&full[0], 1u);
    mbar_wait(&full[1], 1u);
    tc_fence_after();
    mma_kblock(tmem, a_addr + 2 * 16384, w_addr, idesc, false);
    mma_kblock(tmem, a_addr + 3 * 16384, w_addr + 16384, idesc, false);
    umma_commit(bar_acc);
    mbar_wait(bar_acc, 0u);
  }
  __syncthreads();
  tc_fence_after();
  const int q = warp & 3, ch = warp >> 2, r = q * 32 + lane;
  const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    const int col0 = ch * 64 + c * 32;
    uint32_t acc[32];
    tmem_ld32_nowait(ta + col0, acc);
    tmem_wait_ld();
    if (r < nrows) {
      const float4* bo = reinterpret_cast<const float4*>(a.w + a.bias_off + col0);
      float4* dst = reinterpret_cast<float4*>(p.h0 + ((size_t)ins * N + row0 + r) * 128 + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = __ldg(bo + i);
        dst[i] = make_float4(__uint_as_float(acc[4 * i]) + b4.x, __uint_as_float(acc[4 * i + 1]) + b4.y,
                             __uint_as_float(acc[4 * i + 2]) + b4.z, __uint_as_float(acc[4 * i + 3]) + b4.w);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// ---- elementwise helpers -----------------------------------------------------------------------------------------
// float index of (node u, column c) inside one graph of the tile-chunk-row skip layout
__host__ __device__ inline size_t lg_skip_index(int u, int c) {
  return ((((size_t)(u >> 7) * 32 + (c >> 2)) * 128) + (u & 127)) * 4 + (c & 3);
}
// node-encoder input: h = Linear(7,128)([coords, demand, tw, service, time to depot]) (env.py:1047-1053) -> 16-bit + fp32
template <typename T>
__global__ void lg_input_kernel(jampr_dims_t d, jampr_instances_t p, const float* __restrict__ w, T* __restrict__ h16,
                                float* __restrict__ hskip, int npad) {
  const int N = d.n_nodes, ins = blockIdx.y, u = blockIdx.x * 2 + (threadIdx.x >> 7), c = threadIdx.x & 127;
  if (u >= N) return;
  const size_t iu = (size_t)ins * N + u;
  float x = __ldg(w + OFF_NE_IN_B + c);
  x = fmaf(p.coords[iu * 2], __ldg(w + OFF_NE_IN_WT + 0 * 128 + c), x);
  x = fmaf(p.coords[iu * 2 + 1], __ldg(w + OFF_NE_IN_WT + 1 * 128 + c), x);
  x = fmaf(p.demand[iu], __ldg(w + OFF_NE_IN_WT + 2 * 128 + c), x);
  x = fmaf(p.tw[iu * 2], __ldg(w + OFF_NE_IN_WT + 3 * 128 + c), x);
  x = fmaf(p.tw[iu * 2 + 1], __ldg(w + OFF_NE_IN_WT + 4 * 128 + c), x);
  x = fmaf(p.service[ins], __ldg(w + OFF_NE_IN_WT + 5 * 128 + c), x);
  x = fmaf(p.ttd[iu], __ldg(w + OFF_NE_IN_WT + 6 * 128 + c), x);
  const size_t o = ((size_t)ins * npad + u) * 128 + c;
  h16[o] = to_T<T>(x);
  hskip[(size_t)ins * npad * 128 + lg_skip_index(u, c)] = x;
}

// tour-encoder start: every rollout starts from its instance's h0 (16-bit rows from the row-major h0, fp32 skip operand
// from its tile-chunk-row copy h0_t: both copies are contiguous on both sides)
template <typename T>
__global__ void lg_start_kernel(jampr_dims_t d, jampr_instances_t p, jampr_state_t st, T* __restrict__ h16,
                                float* __restrict__ hskip, int npad, int step_no) {
  if (!jampr_step_active(st.ctl, step_no)) return;
  const int N = d.n_nodes, g = blockIdx.y, ins = g / d.n_samples;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;       // float4 index inside the graph
  if (e >= npad * 32) return;
  reinterpret_cast<float4*>(hskip + (size_t)g * npad * 128)[e] =
      __ldg(reinterpret_cast<const float4*>(p.h0_t + (size_t)ins * npad * 128) + e);
  if (e >= N * 32) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(p.h0 + (size_t)ins * N * 128) + e);
  uint2 pk;
  pk.x = as_u32(Elem<T>::from2(v.x, v.y));
  pk.y = as_u32(Elem<T>::from2(v.z, v.w));
  reinterpret_cast<uint2*>(h16 + (size_t)g * npad * 128)[e] = pk;
}

// tile-chunk-row copy of h0 for graphs of more than one tile (include/jampr_b200.h: h0_t), once per episode
__global__ void lg_h0t_kernel(int N, int npad, const float* __restrict__ h0, float* __restrict__ h0_t) {
  const size_t ins = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;       // float4 index inside the tile-chunk-row graph
  if (e >= npad * 32) return;
  const int r = e & 127, c4 = (e >> 7) & 31, tile = e >> 12, u = tile * 128 + r;
  float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (u < N) v = __ldg(reinterpret_cast<const float4*>(h0 + (ins * N + u) * 128) + c4);
  reinterpret_cast<float4*>(h0_t + ins * (size_t)npad * 128)[e] = v;
}

// packed kNN edges for the large path: low half = source row index, high half = 16-bit weight
template <typename T>
__global__ void lg_prepare_kernel(jampr_dims_t d, jampr_instances_t p) {
  const int N = d.n_nodes, k = d.k_nbh, k4 = (k + 3) & ~3, ins = blockIdx.x;
  const int16_t* gk = p.knn + (size_t)ins * N * k;
  const float* gw = p.knn_w + (size_t)ins * N * k;
  for (int e = threadIdx.x; e < N * k4; e += blockDim.x) {
    const int u = e / k4, j = min(e - u * k4, k - 1);
    p.tc_tab[(size_t)ins * N * k4 + e] = (uint32_t)(uint16_t)gk[u * k + j] | ((uint32_t)bits_T<T>(gw[u * k + j]) << 16);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
#define LG_LAYER_SMEM (1024 + 110144)
#define LG_PROJ_SMEM (1024 + 65600)
#define LG_H0_SMEM (1024 + 98368)

template <typename T>
static int lg_node_encoder_t(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w, cudaStream_t s) {
  const int N = d->n_nodes, I = d->n_inst, nt = (N + 127) / 128, npad = nt * 128;
  T* buf0 = reinterpret_cast<T*>(inst->lg_h16);
  T* buf1 = buf0 + (size_t)I * npad * 128;
  CUDA_RET(cudaFuncSetAttribute(lg_layer_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, LG_LAYER_SMEM));
  CUDA_RET(cudaFuncSetAttribute(lg_proj_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, LG_PROJ_SMEM));
  CUDA_RET(cudaFuncSetAttribute(lg_h0_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, LG_H0_SMEM));
  if (N > 128) lg_prepare_kernel<T><<<I, 256, 0, s>>>(*d, *inst);     // else: the step kernel's table (jampr_tc_prepare)
  lg_input_kernel<T><<<dim3((N + 1) / 2, I), 256, 0, s>>>(*d, *inst, w->blob, buf0, inst->lg_skip, npad);
  LgArgs a;
  a.d = *d; a.p = *inst; a.st = jampr_state_t{}; a.w = w->blob; a.w16 = w->blob_bf16;
  a.tab_shift = N > 128 ? 0 : 7;
  a.hskip = inst->lg_skip; a.hfin = nullptr; a.kind = 0; a.per_instance = 1; a.step_no = 0; a.mode = 1; a.bias_off = OFF_NE_OUT_B;
  T* src = buf0;
  T* dst = buf1;
  for (int l = 0; l < 3; ++l) {
    a.h_src = src; a.h_dst = dst; a.img = IMG_NE_BASE + l * 4; a.par_off = OFF_NE_BLK + l * BLK_FLOATS + BLK_B;
    lg_layer_kernel<T><<<dim3(nt, I), LG_THREADS, LG_LAYER_SMEM, s>>>(a);
    T* t = src; src = dst; dst = t;
  }
  a.h_src = src; a.img = IMG_NE_BASE + 12;
  lg_proj_kernel<T><<<dim3(nt, I), LG_THREADS, LG_PROJ_SMEM, s>>>(a, inst->static_emb, nullptr, nullptr);
  a.img = IMG_NE_BASE + 16; a.bias_off = OFF_TE_IN_B;
  lg_h0_kernel<T><<<dim3(nt, I), LG_THREADS, LG_H0_SMEM, s>>>(a);
  if (N > 128) {
    if (!inst->h0_t) return JAMPR_E_BADARG;
    lg_h0t_kernel<<<dim3((npad * 32 + 255) / 256, I), 256, 0, s>>>(N, npad, inst->h0, inst->h0_t);
  }
  return launch_status();
}

template <typename T>
static int lg_tour_encoder_t(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                             const jampr_state_t* st, int step_no, cudaStream_t s) {
  const int N = d->n_nodes, bs = d->n_inst * d->n_samples, nt = (N + 127) / 128, npad = nt * 128;
  T* buf0 = reinterpret_cast<T*>(st->hbuf);
  T* buf1 = buf0 + (size_t)bs * npad * 128;
  CUDA_RET(cudaFuncSetAttribute(lg_layer_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, LG_LAYER_SMEM));
  CUDA_RET(cudaFuncSetAttribute(lg_proj_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, LG_PROJ_SMEM));
  lg_start_kernel<T><<<dim3((npad * 32 + 255) / 256, bs), 256, 0, s>>>(*d, *inst, *st, buf0, st->hskip, npad, step_no);
  LgArgs a;
  a.d = *d; a.p = *inst; a.st = *st; a.w = w->blob; a.w16 = w->blob_bf16;
  a.tab_shift = N > 128 ? 0 : 7;
  a.hskip = st->hskip; a.per_instance = 0; a.step_no = step_no; a.mode = 0; a.bias_off = OFF_TE_OUT_B;
  T* src = buf0;
  T* dst = buf1;
  for (int l = 0; l < 6; ++l) {
    a.h_src = src; a.h_dst = dst; a.img = l * 4; a.par_off = OFF_TE_BLK + l * BLK_FLOATS + BLK_B;
    a.kind = (l == 2 || l == 5) ? 0 : (l < 2 ? 1 : 2);
    a.hfin = (l == 5) ? st->h_fin : nullptr;
    lg_layer_kernel<T><<<dim3(nt, bs), LG_THREADS, LG_LAYER_SMEM, s>>>(a);
    T* t = src; src = dst; dst = t;
  }
  a.h_src = src; a.img = 24; a.hfin = nullptr;
  lg_proj_kernel<T><<<dim3(nt, bs), LG_THREADS, LG_PROJ_SMEM, s>>>(a, st->ne, st->stats_part, reinterpret_cast<T*>(st->ne16));
  lg_stats_kernel<<<bs, 256, 0, s>>>(*d, *st, st->stats_part, nt, step_no);
  return launch_status();
}

// Hand-over of the layer-wise path to the 128-rollout decoder (decoder_tc.cu), every step: node statistics -> st.dq,
// per-tour means of h over buffer positions 0..idx (one depot entry included, tour_encoder.py:236-243) and the tour
// features (env.py:1061-1068) -> st.dqp; the 16-bit node-embedding rows were written by lg_proj_kernel on the last fresh step.
__global__ void __launch_bounds__(128) lg_handover_kernel(jampr_dims_t d, jampr_state_t st, int step_no) {
  if (!jampr_step_active(st.ctl, step_no)) return;
  const int b = blockIdx.x, K = d.n_slots, N = d.n_nodes, L = d.plan_len, c = threadIdx.x;
  if (jampr_rollout_idle(st, b, K)) return;
  for (int e = c; e < 512; e += 128) st.dq[(size_t)b * 512 + e] = st.ne_stats[(size_t)b * 512 + e];
  const float* hf = st.h_fin + (size_t)b * N * 128;
  for (int t = 0; t < K; ++t) {
    const size_t bk = (size_t)b * K + t;
    const int row = st.row[bk];
    const int16_t* pr = st.plan + ((size_t)b * d.k_max + row) * L;
    int idx = 0;
    while (idx < L && pr[idx] != 0) ++idx;
    if (idx >= L) idx = 0;                       // argmin over an all-nonzero row is position 0
    float sum = hf[(size_t)pr[0] * 128 + c];
    for (int q = 1; q <= idx; ++q) sum += hf[(size_t)pr[q] * 128 + c];
    float* o = st.dqp + ((size_t)b * 4 + t) * 256;
    o[c] = sum / (float)(idx + 1);
    if (c < 5) {
      const float f = c == 0 ? (float)st.cur[bk] : c == 1 ? st.cap[bk] : c == 2 ? st.time[bk] : c == 3 ? st.cttd[bk]
                                                                                 : (float)(d.k_max - row) / (float)d.k_max;
      o[128 + c] = f;
    }
  }
}

int jampr_lg_handover(const jampr_dims_t* d, const jampr_state_t* st, int step_no, cudaStream_t s) {
  if (!st->dq || !st->dqp || !st->h_fin || !st->ne_stats) return JAMPR_E_BADARG;
  lg_handover_kernel<<<d->n_inst * d->n_samples, 128, 0, s>>>(*d, *st, step_no);
  return launch_status();
}

static int lg_check(const jampr_dims_t* d, const jampr_weights_t* w) {
  if (d->n_nodes > 32767 || d->k_nbh > 1024) return JAMPR_E_TOOLARGE;
  if (!w->blob_bf16 || w->n_bf16 < (int64_t)(IMG_NE_BASE + 20) * W16_STAGE) return JAMPR_E_BADARG;
  return 0;
}

int jampr_lg_node_encoder(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w, cudaStream_t s) {
  const int rc = lg_check(d, w);
  if (rc) return rc;
  if (!inst->tc_tab || !inst->lg_h16 || !inst->lg_skip) return JAMPR_E_BADARG;
  return w->precision == JAMPR_PREC_F16 ? lg_node_encoder_t<__half>(d, inst, w, s)
                                        : lg_node_encoder_t<__nv_bfloat16>(d, inst, w, s);
}

int jampr_lg_tour_encoder(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                          const jampr_state_t* st, int step_no, cudaStream_t s) {
  const int rc = lg_check(d, w);
  if (rc) return rc;
  if (!st || !st->hbuf || !st->hskip || !st->stats_part || !st->h_fin || !st->ne || !st->ne_stats || !inst->h0_t) return JAMPR_E_BADARG;
  return w->precision == JAMPR_PREC_F16 ? lg_tour_encoder_t<__half>(d, inst, w, st, step_no, s)
                                        : lg_tour_encoder_t<__nv_bfloat16>(d, inst, w, st, step_no, s);
}
