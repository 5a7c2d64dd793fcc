// Fused per-step decoder: per-tour embeddings, query, multi-head glimpse, tanh-clipped pointer logits,
// masked log-softmax and greedy / inverse-CDF action selection, one CTA per rollout.
// Reference: tour_encoder.py:202-213 + :226-243 (tour embeddings), policy.py:170-227
// (query / action embeddings, selection), attn_decoder.py:62-98.
//
// The reference projects every action embedding with set_proj (A x 256 x 768 MACs per rollout and
// step).  The three thirds of that projection are only ever used inside dot products, so the kernel
// moves the projection to the other operand instead (same algebra, 1/50 of the MACs):
//   glimpse score  ctxt_h . (Wgk_h act_a)      = (Wgk_h^T ctxt_h) . act_a
//   glimpse        sum_a p_a (Wgv_h act_a)      = Wgv_h (sum_a p_a act_a)
//   pointer logit  x . (Wlk act_a)              = (Wlk^T x) . act_a
// Everything is fp32; the reassociation stays inside the 1e-5 log-probability tolerance (tests/).
#include "common.cuh"

#define DEC_THREADS 256

struct DecArgs {
  jampr_dims_t d;
  jampr_instances_t p;
  jampr_state_t st;
  const float* w;
  const float* uniforms;
  uint64_t seed;
  int step_no;
  int decode_type;
};

__global__ void __launch_bounds__(DEC_THREADS, 2) decoder_kernel(const DecArgs a) {
  extern __shared__ __align__(16) float dsm[];
  const jampr_dims_t& d = a.d;
  const jampr_state_t& st = a.st;
  if (!jampr_step_active(st.ctl, a.step_no)) return;
  const int N = d.n_nodes, K = d.n_slots, k = d.k_nbh, A = K * k, L = d.plan_len;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (jampr_rollout_idle(st, b, K)) {
    if (tid == 0) {
      st.action[2 * (size_t)b] = 0;
      st.action[2 * (size_t)b + 1] = 0;
      st.ll[b] = 0.0f;
      st.entropy[b] = 0.0f;
    }
    return;
  }
  const float* __restrict__ w = a.w;

  float* s_act = dsm;                         // [A][256]
  float* s_te = s_act + (size_t)A * 256;      // [K][256] tour embeddings
  float* s_cat = s_te + K * 256;              // [K][256] cat(tour feature proj, per-tour mean)
  float* s_q = s_cat + K * 256;               // [512]
  float* s_ctxt = s_q + 512;                  // [256]
  float* s_qp = s_ctxt + 256;                 // [4][256]
  float* s_gbar = s_qp + 1024;                // [4][256]
  float* s_g = s_gbar + 1024;                 // [256]
  float* s_x = s_g + 256;                     // [256]
  float* s_lp = s_x + 256;                    // [256]
  float* s_sc = s_lp + 256;                   // [4][A] glimpse scores -> probabilities
  float* s_logit = s_sc + 4 * A;              // [A]
  float* s_feat = s_logit + A;                // [K][8]
  int* s_nbh = reinterpret_cast<int*>(s_feat + K * 8);   // [A]
  int* s_idx = s_nbh + A;                     // [K] tour length incl. the closing depot entry
  int* s_misc = s_idx + K;                    // [4]

  // ---- tour features (env.py:1061-1068) and tour buffer lengths
  if (tid < K) {
    const size_t bk = (size_t)b * K + tid;
    const int row = st.row[bk];
    s_feat[tid * 8 + 0] = (float)st.cur[bk];
    s_feat[tid * 8 + 1] = st.cap[bk];
    s_feat[tid * 8 + 2] = st.time[bk];
    s_feat[tid * 8 + 3] = st.cttd[bk];
    s_feat[tid * 8 + 4] = (float)(d.k_max - row) / (float)d.k_max;
    const int16_t* pr = st.plan + ((size_t)b * d.k_max + row) * L;
    int idx = 0;
    while (idx < L && pr[idx] != 0) ++idx;
    s_idx[tid] = (idx >= L) ? 0 : idx;        // argmin over an all-nonzero row is position 0
  }
  for (int e = tid; e < A; e += DEC_THREADS) s_nbh[e] = st.nbh[(size_t)b * A + e];
  __syncthreads();

  // ---- cat = [input_proj(tour_features) | mean of h over the tour incl. one depot entry]
  {
    const int c = tid & 127, part = tid >> 7;
    const float* hf = st.h_fin + (size_t)b * N * 128;
    for (int t = 0; t < K; ++t) {
      if (part == 0) {
        float v = __ldg(w + OFF_TF_B + c);
#pragma unroll
        for (int i = 0; i < 5; ++i) v = fmaf(s_feat[t * 8 + i], __ldg(w + OFF_TF_WT + i * 128 + c), v);
        s_cat[t * 256 + c] = v;
      } else {
        const int16_t* pr = st.plan + ((size_t)b * d.k_max + st.row[(size_t)b * K + t]) * L;
        const int idx = s_idx[t];
        float sum = hf[(size_t)pr[0] * 128 + c];
        for (int q = 1; q <= idx; ++q) sum += hf[(size_t)pr[q] * 128 + c];
        s_cat[t * 256 + 128 + c] = sum / (float)(idx + 1);
      }
    }
  }
  __syncthreads();

  // ---- tour_emb = output_proj(cat)   (K x 256 @ 256 x 256)
  {
    float acc[JAMPR_MAX_SLOTS];
    const float bo = __ldg(w + OFF_TO_B + tid);
#pragma unroll
    for (int t = 0; t < JAMPR_MAX_SLOTS; ++t) acc[t] = bo;
    for (int c = 0; c < 256; ++c) {
      const float wv = __ldg(w + OFF_TO_WT + (size_t)c * 256 + tid);
#pragma unroll
      for (int t = 0; t < JAMPR_MAX_SLOTS; ++t)
        if (t < K) acc[t] = fmaf(s_cat[t * 256 + c], wv, acc[t]);
    }
    float sm = 0.0f, mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < JAMPR_MAX_SLOTS; ++t)
      if (t < K) {
        s_te[t * 256 + tid] = acc[t];
        sm += acc[t];
        mx = fmaxf(mx, acc[t]);
      }
    // query = [mean_N(node) + mean_K(tour) | max_N(node) + max_K(tour)]  (policy.py:179-181)
    const float* nst = st.ne_stats + (size_t)b * 512;
    s_q[tid] = nst[tid] + sm / (float)K;
    s_q[256 + tid] = nst[256 + tid] + mx;
  }
  __syncthreads();

  // ---- action embeddings act[a] = tour_emb[a / k] + node_emb[nbh[a]]  (policy.py:184-201)
  {
    const float* ne = st.ne + (size_t)b * N * 256;
    for (int e = tid; e < A * 64; e += DEC_THREADS) {
      const int aa = e >> 6, c4 = (e & 63) * 4;
      const float4 nv = *reinterpret_cast<const float4*>(ne + (size_t)s_nbh[aa] * 256 + c4);
      const float4 tv = *reinterpret_cast<const float4*>(s_te + (aa / k) * 256 + c4);
      *reinterpret_cast<float4*>(s_act + (size_t)aa * 256 + c4) =
          make_float4(tv.x + nv.x, tv.y + nv.y, tv.z + nv.z, tv.w + nv.w);
    }
  }
  // ---- ctxt = ctxt_proj(query)
  {
    float acc = 0.0f;
    for (int c = 0; c < 512; ++c) acc = fmaf(s_q[c], __ldg(w + OFF_CTXT_WT + (size_t)c * 256 + tid), acc);
    s_ctxt[tid] = acc;
  }
  __syncthreads();
  // ---- qp[h] = Wgk_h^T ctxt_h
  {
#pragma unroll
    for (int h = 0; h < JAMPR_HEADS; ++h) {
      float acc = 0.0f;
      for (int e = 0; e < 64; ++e)
        acc = fmaf(s_ctxt[h * 64 + e], __ldg(w + OFF_GK_W + (size_t)(h * 64 + e) * 256 + tid), acc);
      s_qp[h * 256 + tid] = acc;
    }
  }
  __syncthreads();
  // ---- glimpse scores, one warp per action
  for (int aa = warp; aa < A; aa += DEC_THREADS / 32) {
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      const float v = s_act[(size_t)aa * 256 + c];
      p0 = fmaf(v, s_qp[c], p0);
      p1 = fmaf(v, s_qp[256 + c], p1);
      p2 = fmaf(v, s_qp[512 + c], p2);
      p3 = fmaf(v, s_qp[768 + c], p3);
    }
    p0 = warp_sum(p0); p1 = warp_sum(p1); p2 = warp_sum(p2); p3 = warp_sum(p3);
    if (lane == 0) {
      const bool m = st.mask[(size_t)b * A + aa] != 0;
      s_sc[0 * A + aa] = m ? -INFINITY : p0 * 0.125f;
      s_sc[1 * A + aa] = m ? -INFINITY : p1 * 0.125f;
      s_sc[2 * A + aa] = m ? -INFINITY : p2 * 0.125f;
      s_sc[3 * A + aa] = m ? -INFINITY : p3 * 0.125f;
    }
  }
  __syncthreads();
  // ---- softmax per head (warps 0..3)
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
  // ---- gbar[h] = sum_a p[h][a] act[a]
  {
    float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, g3 = 0.0f;
    for (int aa = 0; aa < A; ++aa) {
      const float v = s_act[(size_t)aa * 256 + tid];
      g0 = fmaf(s_sc[aa], v, g0);
      g1 = fmaf(s_sc[A + aa], v, g1);
      g2 = fmaf(s_sc[2 * A + aa], v, g2);
      g3 = fmaf(s_sc[3 * A + aa], v, g3);
    }
    s_gbar[tid] = g0; s_gbar[256 + tid] = g1; s_gbar[512 + tid] = g2; s_gbar[768 + tid] = g3;
  }
  __syncthreads();
  // ---- glimpse g[o] = Wgv[o] . gbar[head(o)]
  {
    const float* gb = s_gbar + (tid >> 6) * 256;
    float acc = 0.0f;
    for (int c = 0; c < 256; ++c) acc = fmaf(gb[c], __ldg(w + OFF_GV_WT + (size_t)c * 256 + tid), acc);
    s_g[tid] = acc;
  }
  __syncthreads();
  // ---- x = out_proj(g)
  {
    float acc = 0.0f;
    for (int c = 0; c < 256; ++c) acc = fmaf(s_g[c], __ldg(w + OFF_OUT_WT + (size_t)c * 256 + tid), acc);
    s_x[tid] = acc;
  }
  __syncthreads();
  // ---- lp = Wlk^T x
  {
    float acc = 0.0f;
    for (int o = 0; o < 256; ++o) acc = fmaf(s_x[o], __ldg(w + OFF_LK_W + (size_t)o * 256 + tid), acc);
    s_lp[tid] = acc;
  }
  __syncthreads();
  // ---- pointer logits: 10 tanh(lp . act / 16), mask -> -inf
  for (int aa = warp; aa < A; aa += DEC_THREADS / 32) {
    float pp = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      pp = fmaf(s_act[(size_t)aa * 256 + c], s_lp[c], pp);
    }
    pp = warp_sum(pp);
    if (lane == 0) {
      const bool m = st.mask[(size_t)b * A + aa] != 0;
      s_logit[aa] = m ? -INFINITY : 10.0f * tanhf(pp * 0.0625f);
    }
  }
  __syncthreads();
  // ---- log-softmax + selection (warp 0)
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
    float ent = 0.0f;
    float best = -INFINITY;
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
    // first-index argmax across lanes (policy.py:211; ties -> lowest index)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    int choice = besti;
    if (a.decode_type == JAMPR_DECODE_SAMPLING) {
      // inverse CDF over exp(logp): first index whose running sum exceeds u; if rounding leaves the total
      // below u, the last feasible index (restated in oracle/policy_oracle.py:select_inverse_cdf)
      const float u = a.uniforms ? a.uniforms[b] : jampr_uniform(a.seed, (uint32_t)b, (uint32_t)a.step_no);
      float run = 0.0f;
      int hit = -1, lastf = -1;
      for (int base = 0; base < A; base += 32) {
        const int e = base + lane;
        const float pe = (e < A) ? expf(s_logit[e]) : 0.0f;
        float inc = pe;   // inclusive scan in index order, sequential adds to mirror cumsum
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const float cum = run + inc;
        const unsigned hm = __ballot_sync(0xffffffffu, e < A && pe > 0.0f && cum > u);
        const unsigned fm = __ballot_sync(0xffffffffu, e < A && pe > 0.0f);
        if (hit < 0 && hm) hit = base + __ffs(hm) - 1;
        if (fm) lastf = base + 31 - __clz(fm);
        run = __shfl_sync(0xffffffffu, cum, 31);
      }
      choice = (hit >= 0) ? hit : lastf;
      if (choice < 0) choice = 0;
    }
    if (lane == 0) {
      if (choice >= A) choice = 0;   // every action masked: flagged below, keep indices in range
      st.action[2 * (size_t)b] = choice / k;
      st.action[2 * (size_t)b + 1] = s_nbh[choice];
      st.ll[b] = s_logit[choice];
      st.entropy[b] = ent;
      if (nan) atomicOr(st.status + b, JAMPR_ST_NAN_LOGITS);
    }
  }
}

size_t jampr_decoder_smem(const jampr_dims_t* d) {
  const size_t A = (size_t)d->n_slots * d->k_nbh, K = d->n_slots;
  return (A * 256 + K * 256 * 2 + 512 + 256 + 1024 + 1024 + 256 * 3 + 4 * A + A + K * 8) * 4 + (A + K + 4) * 4;
}

int jampr_decoder_launch(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                         const float* w, int step_no, int decode_type, uint64_t seed, const float* uniforms,
                         cudaStream_t s) {
  const size_t smem = jampr_decoder_smem(d);
  if (smem > 227 * 1024) return JAMPR_E_TOOLARGE;
  CUDA_RET(cudaFuncSetAttribute(decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DecArgs a;
  a.d = *d;
  a.p = *inst;
  a.st = *st;
  a.w = w;
  a.uniforms = uniforms;
  a.seed = seed;
  a.step_no = step_no;
  a.decode_type = decode_type;
  decoder_kernel<<<d->n_inst * d->n_samples, DEC_THREADS, smem, s>>>(a);
  return launch_status();
}
