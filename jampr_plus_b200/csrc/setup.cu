// Static per-instance tables: distance matrix, window fix-up, kNN neighbourhoods, depot order.
// Restates reference env.py:369-409 (load_data), challenge_utils.py:23-35, graph_utils.py:104-141 and
// env.py:855-861 as one kernel, one CTA per instance.  All float ops are explicit round-to-nearest
// intrinsics so that nvcc cannot contract them into FMAs: the distance has a floor() in it and the
// feasibility masks built on it must be bit-exact.
#include "common.cuh"

__global__ void __launch_bounds__(256) setup_instances_kernel(jampr_dims_t d, jampr_instances_t p) {
  extern __shared__ float sm[];
  const int N = d.n_nodes, k = d.k_nbh;
  float* cx = sm;            // [N]
  float* cy = sm + N;        // [N]
  float* nrm = sm + 2 * N;   // [N]
  const int i = blockIdx.x;
  const float* c = p.coords + (size_t)i * N * 2;
  for (int u = threadIdx.x; u < N; u += blockDim.x) {
    cx[u] = c[2 * u];
    cy[u] = c[2 * u + 1];
  }
  __syncthreads();
  const float hz = p.horizon[i];
  float* dist = p.dist + (size_t)i * N * N;
  const float* given = p.dist_given ? p.dist_given + (size_t)i * N * N : nullptr;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    if (given) {
      dist[e] = given[e];
      continue;
    }
    const int u = e / N, v = e - u * N;
    const float dx = __fmul_rn(100.0f, __fsub_rn(cx[u], cx[v]));
    const float dy = __fmul_rn(100.0f, __fsub_rn(cy[u], cy[v]));
    const float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float t = __fdiv_rn(floorf(__fmul_rn(10.0f, __fsqrt_rn(s))), 10.0f);
    dist[e] = __fdiv_rn(t, hz);
  }
  for (int u = threadIdx.x; u < N; u += blockDim.x) {
    const float dx = __fsub_rn(cx[u], cx[0]), dy = __fsub_rn(cy[u], cy[0]);
    nrm[u] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  }
  __syncthreads();   // dist rows written by this CTA are read back below (same CTA: visible after barrier)
  const float sv = p.service[i];
  float* tw = p.tw + (size_t)i * N * 2;
  float* ttd = p.ttd + (size_t)i * N;
  int bad = 0;
  for (int u = threadIdx.x; u < N; u += blockDim.x) {
    const float t0 = dist[(size_t)u * N];
    ttd[u] = t0;
    if (u > 0) {
      const float ret = __fadd_rn(__fadd_rn(tw[2 * u + 1], t0), sv);
      if (ret > 1.0f) {
        const float ne = __fsub_rn(tw[2 * u + 1], __fsub_rn(ret, 1.0f));
        tw[2 * u + 1] = ne;
        if (!(__fsub_rn(ne, tw[2 * u]) > 0.005f)) bad = 1;
      }
    }
    // depot order: rank of u under (norm, index)
    int rank = 0;
    const float nu = nrm[u];
    for (int v = 0; v < N; ++v) rank += (nrm[v] < nu) || (nrm[v] == nu && v < u);
    p.order[(size_t)i * N + rank] = (int16_t)u;
  }
  if (bad) atomicOr(p.inst_status + i, JAMPR_ST_TW_COLLAPSE);
  // kNN among customers, self included, ties -> lower index; depot appended
  int16_t* knn = p.knn + (size_t)i * N * k;
  float* knw = p.knn_w + (size_t)i * N * k;
  for (int u = threadIdx.x; u < N; u += blockDim.x) {
    if (u == 0) {
      for (int j = 0; j < k; ++j) { knn[j] = 0; knw[j] = 0.0f; }
      continue;
    }
    float last_d = -1.0f;
    int last_v = -1;
    for (int j = 0; j < k - 1; ++j) {
      float best_d = 3.0e38f;
      int best_v = 0x7fffffff;
      for (int v = 1; v < N; ++v) {
        const float dx = __fsub_rn(cx[u], cx[v]), dy = __fsub_rn(cy[u], cy[v]);
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const bool after = (d2 > last_d) || (d2 == last_d && v > last_v);
        const bool better = (d2 < best_d) || (d2 == best_d && v < best_v);
        if (after && better) { best_d = d2; best_v = v; }
      }
      knn[u * k + j] = (int16_t)best_v;
      knw[u * k + j] = dist[(size_t)u * N + best_v];
      last_d = best_d;
      last_v = best_v;
    }
    knn[u * k + k - 1] = 0;
    knw[u * k + k - 1] = dist[(size_t)u * N];
  }
}

extern "C" int jampr_setup_instances(const jampr_dims_t* d, const jampr_instances_t* inst, void* stream) {
  if (!d || !inst || d->n_inst <= 0 || d->n_nodes < 2 || d->k_nbh < 2 || d->k_nbh > d->n_nodes) return JAMPR_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_RET(cudaMemsetAsync(inst->inst_status, 0, sizeof(int32_t) * d->n_inst, s));
  const size_t smem = sizeof(float) * 3 * d->n_nodes;
  setup_instances_kernel<<<d->n_inst, 256, smem, s>>>(*d, *inst);
  return launch_status();
}

// ---- POMO start nodes (env.py:1221-1270): per instance the depot neighbourhood without the depot (the k - 1 nodes that
// follow the depot in `order`), reduced to the n_cand earliest window starts (stable); per rollout either the sample-th
// of them (pomo_single_start_node) or n_start distinct ones, uniformly: a partial Fisher-Yates shuffle with draws
// jampr_uniform(seed, rollout, 0x70000000 + t).  The reference draws with torch.multinomial on the global generator
// (env.py:924-930, 1258-1262): same distribution, a stream that cannot be matched across devices; oracle/rng.py replays
// this one on the CPU.
#define POMO_MAX_CAND 256
__global__ void pomo_start_kernel(jampr_dims_t d, jampr_instances_t p, int n_cand, int n_start, int single, uint64_t seed,
                                  int32_t* __restrict__ out) {
  __shared__ int16_t s_c[POMO_MAX_CAND];
  const int i = blockIdx.x, N = d.n_nodes, k = d.k_nbh, P = d.n_samples;
  if (threadIdx.x == 0) {
    // insertion sort of the k - 1 candidates by window start, stable in `order` position
    int m = 0;
    for (int j = 1; j < k; ++j) {
      const int16_t n = p.order[(size_t)i * N + j];
      const float key = p.tw[((size_t)i * N + n) * 2];
      int q = m++;
      while (q > 0 && p.tw[((size_t)i * N + s_c[q - 1]) * 2] > key) { s_c[q] = s_c[q - 1]; --q; }
      s_c[q] = n;
    }
  }
  __syncthreads();
  for (int smp = threadIdx.x; smp < P; smp += blockDim.x) {
    const int b = i * P + smp;
    if (single) {
      out[b] = s_c[smp];
      continue;
    }
    uint8_t idx[POMO_MAX_CAND];
    for (int j = 0; j < n_cand; ++j) idx[j] = (uint8_t)j;
    for (int t = 0; t < n_start; ++t) {
      const float u = jampr_uniform(seed, (uint32_t)b, 0x70000000u + (uint32_t)t);
      const int j = t + min((int)(u * (float)(n_cand - t)), n_cand - t - 1);
      const uint8_t tmp = idx[t];
      idx[t] = idx[j];
      idx[j] = tmp;
      out[(size_t)b * n_start + t] = s_c[idx[t]];
    }
  }
}

extern "C" int jampr_pomo_start_nodes(const jampr_dims_t* d, const jampr_instances_t* inst, int32_t n_cand, int32_t n_start,
                                      int32_t single_start, uint64_t seed, int32_t* start_nodes, void* stream) {
  if (!d || !inst || !start_nodes || d->n_inst <= 0 || d->k_nbh < 2 || d->k_nbh - 1 > POMO_MAX_CAND) return JAMPR_E_BADARG;
  if (single_start) {
    if (d->n_samples > d->k_nbh - 1) return JAMPR_E_BADARG;
  } else if (n_cand < n_start || n_cand > d->k_nbh - 1 || n_start <= 0 || n_start > d->n_slots) {
    return JAMPR_E_BADARG;
  }
  pomo_start_kernel<<<d->n_inst, 128, 0, (cudaStream_t)stream>>>(*d, *inst, n_cand, n_start, single_start, seed, start_nodes);
  return launch_status();
}
