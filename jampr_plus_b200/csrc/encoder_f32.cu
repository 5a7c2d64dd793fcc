// fp32 CUDA-core path of the two GNN encoders (the 1e-5 parity anchor; the bf16 tcgen05 path lives in
// encoder_tc.cu).  One CTA owns one graph: its node embedding h (N x 128) and the max-aggregated
// messages agg (N x 128) stay in shared memory for all layers, the (256 x 128) layer weights are
// streamed from L2 in 32-row chunks with cp.async double buffering, and every thread keeps a
// TM x 8 register tile of the (N x 128) layer output.
//
//   node_encoder_f32_kernel   node_encoder.py:128-155 once per instance, plus the hoisted
//                             node_emb_input_proj of tour_encoder.py:170-172
//   tour_encoder_f32_kernel   tour_encoder.py:159-220 once per rollout and refresh step:
//                             3 forward + 3 reversed GraphConv blocks (graph_conv.py:66-90), then
//                             node_emb = node_emb_output_proj(h) + static_emb and its column mean / max
// GraphConv(aggr="max") follows the published PyG semantics:
//   out_i = lin_rel(max_{j->i} w_ji x_j) + lin_root(x_i), nodes without in-edge aggregate to 0.
#include "common.cuh"

#define ENC_THREADS 256
#define HP 132    // row stride of h / agg in floats (128 + 4: adjacent rows land on different banks)
#define WCH 32    // weight rows per streamed chunk

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void load_w_chunk(float* dst, const float* __restrict__ src) {
  // 32 rows x 128 floats = 1024 float4, 4 per thread
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = (threadIdx.x + i * ENC_THREADS) * 4;
    cp_async16(dst + e, src + e);
  }
}

__device__ __forceinline__ float f4_get(const float4& v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// acc[i][0..3] -> output columns tx*4.., acc[i][4..7] -> 64 + tx*4.., row = ty + 16*i.
// A = [A0 | A1]: the first nch0 chunks of K come from A0, the rest from A1 (both smem, stride HP).
template <int TM>
__device__ __forceinline__ void gemm128(float (&acc)[TM][8], const float* A0, const float* A1, int nch0,
                                        int nch, const float* __restrict__ Wg, float* wst) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  load_w_chunk(wst, Wg);
  cp_async_commit();
  for (int c = 0; c < nch; ++c) {
    const float* wb = wst + (c & 1) * (WCH * 128);
    if (c + 1 < nch) {
      load_w_chunk(wst + ((c + 1) & 1) * (WCH * 128), Wg + (size_t)(c + 1) * (WCH * 128));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* A = (c < nch0) ? (A0 + c * WCH) : (A1 + (c - nch0) * WCH);
#pragma unroll 2
    for (int kk = 0; kk < WCH; kk += 4) {
      float4 a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty + 16 * i) * HP + kk);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const float4 w0 = *reinterpret_cast<const float4*>(wb + (kk + k4) * 128 + tx * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(wb + (kk + k4) * 128 + 64 + tx * 4);
        const float2 w00 = make_float2(w0.x, w0.y), w01 = make_float2(w0.z, w0.w);
        const float2 w10 = make_float2(w1.x, w1.y), w11 = make_float2(w1.z, w1.w);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          // packed fp32 FMA (FFMA2, sm_100): each component is the same IEEE fma as the scalar form, at twice the
          // issue rate of the three-register FFMA
          const float av = f4_get(a[i], k4);
          const float2 a2 = make_float2(av, av);
          float2 r0 = __ffma2_rn(a2, w00, make_float2(acc[i][0], acc[i][1]));
          float2 r1 = __ffma2_rn(a2, w01, make_float2(acc[i][2], acc[i][3]));
          float2 r2 = __ffma2_rn(a2, w10, make_float2(acc[i][4], acc[i][5]));
          float2 r3 = __ffma2_rn(a2, w11, make_float2(acc[i][6], acc[i][7]));
          acc[i][0] = r0.x; acc[i][1] = r0.y; acc[i][2] = r1.x; acc[i][3] = r1.y;
          acc[i][4] = r2.x; acc[i][5] = r2.y; acc[i][6] = r3.x; acc[i][7] = r3.y;
        }
      }
    }
    __syncthreads();
  }
}

template <int TM>
__device__ __forceinline__ void zero_acc(float (&acc)[TM][8]) {
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// GraphConv block: y = [agg | h] @ WT + b -> GELU -> + h -> LayerNorm, written back into h (rows < N).
template <int TM>
__device__ void conv_block(float* hs, float* as, float* wst, const float* __restrict__ blk, int N) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[TM][8];
  zero_acc<TM>(acc);
  gemm128<TM>(acc, as, hs, 4, 8, blk + BLK_WT, wst);
  float bias[8], gam[8], bet[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4);
    bias[j] = __ldg(blk + BLK_B + c);
    gam[j] = __ldg(blk + BLK_G + c);
    bet[j] = __ldg(blk + BLK_BETA + c);
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = ty + 16 * i;
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4);
      const float v = gelu_erf(acc[i][j] + bias[j]) + hs[r * HP + c];
      acc[i][j] = v;
      s += v;
    }
    const float mean = half_warp_sum(s) * (1.0f / 128.0f);
    float q = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dv = acc[i][j] - mean;
      acc[i][j] = dv;
      q += dv * dv;
    }
    const float rstd = rsqrtf(half_warp_sum(q) * (1.0f / 128.0f) + 1e-5f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = acc[i][j] * rstd * gam[j] + bet[j];
  }
  // gemm128 ends with a barrier: nobody reads the old h any more except this thread's own skip reads above
  __syncthreads();
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = ty + 16 * i;
    if (r < N) {
      *reinterpret_cast<float4*>(hs + r * HP + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(hs + r * HP + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
  __syncthreads();
}

// Max aggregation over the static neighbourhood graph (graph_utils.py:127-141): customer u receives
// from its k table entries (k-1 nearest customers incl. itself, then the depot), the depot from every node.
__device__ void agg_nbh(const float* hs, float* as, const int16_t* s_knn, const float* s_knw,
                        const float* s_ttd, float* s_red, int N, int k) {
  const int f = threadIdx.x & 127, half = threadIdx.x >> 7;
  for (int u = 1 + half; u < N; u += 2) {
    const int16_t* nb = s_knn + u * k;
    const float* wb = s_knw + u * k;
    float m = wb[0] * hs[nb[0] * HP + f];
    for (int j = 1; j < k; ++j) m = fmaxf(m, wb[j] * hs[nb[j] * HP + f]);
    as[u * HP + f] = m;
  }
  float m = -INFINITY;
  for (int u = half; u < N; u += 2) m = fmaxf(m, s_ttd[u] * hs[u * HP + f]);
  if (half) s_red[f] = m;
  __syncthreads();
  if (!half) as[f] = fmaxf(m, s_red[f]);
  __syncthreads();
}

// Max aggregation over the tour graph kept as predecessor / successor links (env.py:804-835):
// forward edges pred(u) -> u and last-of-tour -> depot; reversed edges succ(u) -> u and first-of-tour -> depot.
__device__ void agg_tour(const float* hs, float* as, const uint8_t* s_vis, const int16_t* s_link,
                         const float* s_wlink, const int16_t* s_other, const float* s_ttd, float* s_red, int N) {
  const int f = threadIdx.x & 127, half = threadIdx.x >> 7;
  float m = -INFINITY;
  for (int u = 1 + half; u < N; u += 2) {
    float v = 0.0f;
    if (s_vis[u]) {
      v = s_wlink[u] * hs[s_link[u] * HP + f];
      if (s_other[u] == 0) m = fmaxf(m, s_ttd[u] * hs[u * HP + f]);
    }
    as[u * HP + f] = v;
  }
  if (half) s_red[f] = m;
  __syncthreads();
  if (!half) {
    m = fmaxf(m, s_red[f]);
    as[f] = (m == -INFINITY) ? 0.0f : m;
  }
  __syncthreads();
}

struct EncSmem {
  float* hs;
  float* as;
  float* wst;
  float* ttd;
  float* knw;
  float* red;
  float* wpred;
  float* wsucc;
  int16_t* knn;
  int16_t* pred;
  int16_t* succ;
  uint8_t* vis;
};

__host__ __device__ inline size_t enc_smem_bytes(int rows, int N, int k) {
  size_t b = 0;
  b += (size_t)rows * HP * 4 * 2;       // hs, as
  b += 2 * WCH * 128 * 4;               // wst
  b += ((size_t)N + 3) / 4 * 4 * 4 * 3; // ttd, wpred, wsucc
  b += ((size_t)N * k + 3) / 4 * 4 * 4; // knw
  b += 128 * 4;                         // red
  b += ((size_t)N * k + 7) / 8 * 8 * 2; // knn
  b += ((size_t)N + 7) / 8 * 8 * 2 * 2; // pred, succ
  b += ((size_t)N + 15) / 16 * 16;      // vis
  return b;
}

__device__ inline EncSmem enc_carve(float* base, int rows, int N, int k) {
  EncSmem s;
  float* p = base;
  s.hs = p; p += (size_t)rows * HP;
  s.as = p; p += (size_t)rows * HP;
  s.wst = p; p += 2 * WCH * 128;
  const int n4 = (N + 3) / 4 * 4;
  s.ttd = p; p += n4;
  s.wpred = p; p += n4;
  s.wsucc = p; p += n4;
  s.knw = p; p += ((size_t)N * k + 3) / 4 * 4;
  s.red = p; p += 128;
  int16_t* q = reinterpret_cast<int16_t*>(p);
  s.knn = q; q += ((size_t)N * k + 7) / 8 * 8;
  s.pred = q; q += (N + 7) / 8 * 8;
  s.succ = q; q += (N + 7) / 8 * 8;
  s.vis = reinterpret_cast<uint8_t*>(q);
  return s;
}

__device__ inline void load_static_tables(const EncSmem& s, const jampr_instances_t& p, int ins, int N, int k) {
  for (int e = threadIdx.x; e < N * k; e += ENC_THREADS) {
    s.knn[e] = p.knn[(size_t)ins * N * k + e];
    s.knw[e] = p.knn_w[(size_t)ins * N * k + e];
  }
  for (int u = threadIdx.x; u < N; u += ENC_THREADS) s.ttd[u] = p.ttd[(size_t)ins * N + u];
}

// out (N x 256) = h @ WT2 (two 128-column halves) + bias [+ addend]; optionally reduces column mean / max.
template <int TM>
__device__ void project_256(const float* hs, float* wst, const float* __restrict__ wt2, const float* __restrict__ bias,
                            const float* __restrict__ addend, float* __restrict__ out, float* stats_smem,
                            float* __restrict__ stats_out, float* keep_half0, int N) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int hf = 0; hf < 2; ++hf) {
    float acc[TM][8];
    zero_acc<TM>(acc);
    gemm128<TM>(acc, hs, hs, 4, 4, wt2 + (size_t)hf * 128 * 128, wst);
    float cs[8], cm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { cs[j] = 0.0f; cm[j] = -INFINITY; }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int r = ty + 16 * i;
      if (r < N) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int c = hf * 128 + g * 64 + tx * 4;
          const float4 b4 = *reinterpret_cast<const float4*>(bias + c);
          float4 v = make_float4(acc[i][g * 4 + 0] + b4.x, acc[i][g * 4 + 1] + b4.y, acc[i][g * 4 + 2] + b4.z,
                                 acc[i][g * 4 + 3] + b4.w);
          if (addend) {
            const float4 a4 = *reinterpret_cast<const float4*>(addend + (size_t)r * 256 + c);
            v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
          }
          if (out) *reinterpret_cast<float4*>(out + (size_t)r * 256 + c) = v;
          if (keep_half0 && hf == 0)
            *reinterpret_cast<float4*>(keep_half0 + r * HP + g * 64 + tx * 4) = v;
          if (keep_half0 && hf == 1) { acc[i][g * 4 + 0] = v.x; acc[i][g * 4 + 1] = v.y; acc[i][g * 4 + 2] = v.z; acc[i][g * 4 + 3] = v.w; }
          cs[g * 4 + 0] += v.x; cs[g * 4 + 1] += v.y; cs[g * 4 + 2] += v.z; cs[g * 4 + 3] += v.w;
          cm[g * 4 + 0] = fmaxf(cm[g * 4 + 0], v.x); cm[g * 4 + 1] = fmaxf(cm[g * 4 + 1], v.y);
          cm[g * 4 + 2] = fmaxf(cm[g * 4 + 2], v.z); cm[g * 4 + 3] = fmaxf(cm[g * 4 + 3], v.w);
        }
      }
    }
    if (stats_out) {
      // reduce over the 16 row groups: xor 16 joins the two ty of a warp, then 8 warps through smem
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
        cm[j] = fmaxf(cm[j], __shfl_xor_sync(0xffffffffu, cm[j], 16));
      }
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      if (lane < 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4);
          stats_smem[(warp * 2 + 0) * 128 + c] = cs[j];
          stats_smem[(warp * 2 + 1) * 128 + c] = cm[j];
        }
      }
      __syncthreads();
      if (threadIdx.x < 128) {
        float s = 0.0f, m = -INFINITY;
        for (int w = 0; w < 8; ++w) {
          s += stats_smem[(w * 2 + 0) * 128 + threadIdx.x];
          m = fmaxf(m, stats_smem[(w * 2 + 1) * 128 + threadIdx.x]);
        }
        stats_out[hf * 128 + threadIdx.x] = s / (float)N;
        stats_out[256 + hf * 128 + threadIdx.x] = m;
      }
      __syncthreads();
    }
    if (keep_half0 && hf == 1) {
      // second half replaces h in place (all reads of h finished at the last barrier of gemm128)
      float* hw = const_cast<float*>(hs);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int r = ty + 16 * i;
        if (r < N) {
          *reinterpret_cast<float4*>(hw + r * HP + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          *reinterpret_cast<float4*>(hw + r * HP + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        }
      }
      __syncthreads();
    }
  }
}

template <int TM>
__global__ void __launch_bounds__(ENC_THREADS, 1)
node_encoder_f32_kernel(jampr_dims_t d, jampr_instances_t p, const float* __restrict__ w) {
  extern __shared__ __align__(16) float smem[];
  const int N = d.n_nodes, k = d.k_nbh, ins = blockIdx.x;
  const EncSmem s = enc_carve(smem, 16 * TM, N, k);
  for (int e = threadIdx.x; e < 16 * TM * HP; e += ENC_THREADS) { s.hs[e] = 0.0f; s.as[e] = 0.0f; }
  load_static_tables(s, p, ins, N, k);
  __syncthreads();
  // input_proj Linear(7,128) over [coords, demand, tw, service, time to depot] (env.py:1047-1053)
  {
    const int c = threadIdx.x & 127, half = threadIdx.x >> 7;
    float wc[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) wc[i] = __ldg(w + OFF_NE_IN_WT + i * 128 + c);
    const float bc = __ldg(w + OFF_NE_IN_B + c);
    const float sv = p.service[ins];
    for (int u = half; u < N; u += 2) {
      const size_t iu = (size_t)ins * N + u;
      float x = bc;
      x = fmaf(p.coords[iu * 2], wc[0], x);
      x = fmaf(p.coords[iu * 2 + 1], wc[1], x);
      x = fmaf(p.demand[iu], wc[2], x);
      x = fmaf(p.tw[iu * 2], wc[3], x);
      x = fmaf(p.tw[iu * 2 + 1], wc[4], x);
      x = fmaf(sv, wc[5], x);
      x = fmaf(s.ttd[u], wc[6], x);
      s.hs[u * HP + c] = x;
    }
  }
  __syncthreads();
  for (int l = 0; l < 3; ++l) {
    agg_nbh(s.hs, s.as, s.knn, s.knw, s.ttd, s.red, N, k);
    conv_block<TM>(s.hs, s.as, s.wst, w + OFF_NE_BLK + (size_t)l * BLK_FLOATS, N);
  }
  // static_emb = output_proj(h): columns 0..127 parked in `as`, 128..255 replace h
  project_256<TM>(s.hs, s.wst, w + OFF_NE_OUT_WT, w + OFF_NE_OUT_B, nullptr,
                  p.static_emb + (size_t)ins * N * 256, nullptr, nullptr, s.as, N);
  // h0 = node_emb_input_proj(static_emb), K = 256 = [as | hs]
  {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[TM][8];
    zero_acc<TM>(acc);
    gemm128<TM>(acc, s.as, s.hs, 4, 8, w + OFF_TE_IN_WT, s.wst);
    float* h0 = p.h0 + (size_t)ins * N * 128;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int r = ty + 16 * i;
      if (r < N) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int c = g * 64 + tx * 4;
          const float4 b4 = *reinterpret_cast<const float4*>(w + OFF_TE_IN_B + c);
          *reinterpret_cast<float4*>(h0 + (size_t)r * 128 + c) =
              make_float4(acc[i][g * 4 + 0] + b4.x, acc[i][g * 4 + 1] + b4.y, acc[i][g * 4 + 2] + b4.z,
                          acc[i][g * 4 + 3] + b4.w);
        }
      }
    }
  }
}

template <int TM>
__global__ void __launch_bounds__(ENC_THREADS, 1)
tour_encoder_f32_kernel(jampr_dims_t d, jampr_instances_t p, jampr_state_t st, const float* __restrict__ w,
                        int step_no) {
  extern __shared__ __align__(16) float smem[];
  if (!jampr_step_active(st.ctl, step_no)) return;
  const int N = d.n_nodes, k = d.k_nbh, b = blockIdx.x, ins = b / d.n_samples;
  const EncSmem s = enc_carve(smem, 16 * TM, N, k);
  for (int e = threadIdx.x; e < 16 * TM * HP; e += ENC_THREADS) { s.hs[e] = 0.0f; s.as[e] = 0.0f; }
  load_static_tables(s, p, ins, N, k);
  __syncthreads();
  {
    const float* h0 = p.h0 + (size_t)ins * N * 128;
    for (int e = threadIdx.x; e < N * 32; e += ENC_THREADS) {
      const int r = e >> 5, c4 = (e & 31) * 4;
      *reinterpret_cast<float4*>(s.hs + r * HP + c4) = *reinterpret_cast<const float4*>(h0 + (size_t)r * 128 + c4);
    }
    const float* dist = p.dist + (size_t)ins * N * N;
    for (int u = threadIdx.x; u < N; u += ENC_THREADS) {
      const int pr = st.pred[(size_t)b * N + u], su = st.succ[(size_t)b * N + u];
      s.pred[u] = (int16_t)pr;
      s.succ[u] = (int16_t)su;
      s.vis[u] = st.visited[(size_t)b * N + u];
      s.wpred[u] = dist[(size_t)u * N + pr];
      s.wsucc[u] = dist[(size_t)u * N + su];
    }
  }
  __syncthreads();
  for (int l = 0; l < 6; ++l) {
    if (l == 2 || l == 5) agg_nbh(s.hs, s.as, s.knn, s.knw, s.ttd, s.red, N, k);
    else if (l < 2) agg_tour(s.hs, s.as, s.vis, s.pred, s.wpred, s.succ, s.ttd, s.red, N);
    else agg_tour(s.hs, s.as, s.vis, s.succ, s.wsucc, s.pred, s.ttd, s.red, N);
    conv_block<TM>(s.hs, s.as, s.wst, w + OFF_TE_BLK + (size_t)l * BLK_FLOATS, N);
  }
  {
    float* hf = st.h_fin + (size_t)b * N * 128;
    for (int e = threadIdx.x; e < N * 32; e += ENC_THREADS) {
      const int r = e >> 5, c4 = (e & 31) * 4;
      *reinterpret_cast<float4*>(hf + (size_t)r * 128 + c4) = *reinterpret_cast<const float4*>(s.hs + r * HP + c4);
    }
  }
  // node_emb = node_emb_output_proj(h) + static_emb (tour_encoder.py:217-218); `as` is free now and
  // serves as scratch for the column statistics
  project_256<TM>(s.hs, s.wst, w + OFF_TE_OUT_WT, w + OFF_TE_OUT_B, p.static_emb + (size_t)ins * N * 256,
                  st.ne + (size_t)b * N * 256, s.as, st.ne_stats + (size_t)b * 512, nullptr, N);
}

static int tm_for(int N) { return N <= 64 ? 4 : (N <= 112 ? 7 : (N <= 128 ? 8 : 0)); }

template <int TM>
static int launch_node_enc(const jampr_dims_t* d, const jampr_instances_t* inst, const float* w, cudaStream_t s) {
  const size_t smem = enc_smem_bytes(16 * TM, d->n_nodes, d->k_nbh);
  if (smem > 227 * 1024) return JAMPR_E_TOOLARGE;
  CUDA_RET(cudaFuncSetAttribute(node_encoder_f32_kernel<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  node_encoder_f32_kernel<TM><<<d->n_inst, ENC_THREADS, smem, s>>>(*d, *inst, w);
  return launch_status();
}

template <int TM>
static int launch_tour_enc(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                           const float* w, int step_no, cudaStream_t s) {
  const size_t smem = enc_smem_bytes(16 * TM, d->n_nodes, d->k_nbh);
  if (smem > 227 * 1024) return JAMPR_E_TOOLARGE;
  CUDA_RET(cudaFuncSetAttribute(tour_encoder_f32_kernel<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tour_encoder_f32_kernel<TM><<<d->n_inst * d->n_samples, ENC_THREADS, smem, s>>>(*d, *inst, *st, w, step_no);
  return launch_status();
}

int jampr_node_encoder_f32(const jampr_dims_t* d, const jampr_instances_t* inst, const float* w, cudaStream_t s) {
  switch (tm_for(d->n_nodes)) {
    case 4: return launch_node_enc<4>(d, inst, w, s);
    case 7: return launch_node_enc<7>(d, inst, w, s);
    case 8: return launch_node_enc<8>(d, inst, w, s);
    default: return JAMPR_E_TOOLARGE;
  }
}

int jampr_tour_encoder_f32(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                           const float* w, int step_no, cudaStream_t s) {
  switch (tm_for(d->n_nodes)) {
    case 4: return launch_tour_enc<4>(d, inst, st, w, step_no, s);
    case 7: return launch_tour_enc<7>(d, inst, st, w, step_no, s);
    case 8: return launch_tour_enc<8>(d, inst, st, w, step_no, s);
    default: return JAMPR_E_TOOLARGE;
  }
}
