// Tensor-native solution I/O for the LNS repair loop (sm_100a; thread per rollout, the batches are small).
//   * lns_destroy_kernel       the "tensor-based native destruction operator" the reference left as a stub
//                              (lib/routing/env.py:720-723): lib/lns/destructor.py:86-205 (random_from_route with
//                              random_nodes / random_cut / const_cut, random / smallest complete-route removal) applied
//                              directly to tour buffers, random numbers from the counter-based generator of common.cuh
//   * env_import_tours_kernel  RPEnv.import_sol (lib/routing/env.py:575-706, _recompute_cost :708-718) from dense tour
//                              arrays instead of nested Python lists: tour buffers, vehicle slots, capacity / clock of
//                              partial tours, recomputed totals — same operation order as the reference's loop
// The tail of import_sol (visited flags, tour links, first observation, new step counter) stays jampr_env_import_finish.
#include "common.cuh"

// jampr_uniform(seed, rollout, iteration * LNS_DRAWS + draw): one stream of draws per rollout and LNS iteration
#define LNS_DRAWS 8192u

struct Draws {
  uint64_t seed;
  uint32_t b, base, c;
  __device__ float next() { return jampr_uniform(seed, b, base + (c++)); }
};

// destructor.py:58-63 rnd_int_1: uniform integer in [1, n - 2] for n > 2, else 1
__device__ __forceinline__ int rnd_int_1(Draws& r, int n) {
  if (n <= 2) return 1;
  const float u = r.next();
  return 1 + min((int)(u * (float)(n - 2)), n - 3);
}

// marks the `n_rm` smallest keys among key[0..n) (ties -> lower index) in rm[]
__device__ void mark_smallest(const float* key, int n, int n_rm, uint8_t* rm) {
  for (int i = 0; i < n; ++i) rm[i] = 0;
  for (int q = 0; q < n_rm && q < n; ++q) {
    int best = -1;
    for (int i = 0; i < n; ++i)
      if (!rm[i] && (best < 0 || key[i] < key[best])) best = i;
    rm[best] = 1;
  }
}

#define LNS_MAX_ROWS 256
#define LNS_MAX_LEN 64

__global__ void lns_destroy_kernel(jampr_dims_t d, const int16_t* __restrict__ src_plan, const int32_t* __restrict__ src_index,
                                   jampr_destroy_t prm, int16_t* __restrict__ tours, uint8_t* __restrict__ kind) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int bs = d.n_inst * d.n_samples;
  if (b >= bs) return;
  const int R = d.k_max, L = d.plan_len;
  const int16_t* plan = src_plan + (size_t)src_index[b] * R * L;
  int16_t* out = tours + (size_t)b * R * L;
  uint8_t* okind = kind + (size_t)b * R;
  Draws rng{prm.seed, (uint32_t)b, (uint32_t)prm.iteration * LNS_DRAWS, 0u};
  // row lengths; rows of one customer are singleton tours: always destroyed (destructor.py:98-117)
  uint8_t len[LNS_MAX_ROWS], role[LNS_MAX_ROWS];      // role: 0 none / singleton, 1 complete, 2 picked partial, 3 removed
  float key[LNS_MAX_ROWS];
  uint8_t rm[LNS_MAX_ROWS];
  int n_cand = 0;
  for (int r = 0; r < R; ++r) {
    int l = 0;
    while (l < L && plan[r * L + l] != 0) ++l;
    len[r] = (uint8_t)l;
    role[r] = l > 1 ? 1 : 0;
    n_cand += l > 1;
  }
  // num_partial distinct candidate tours, uniformly (keys in row order, the smallest win)
  int16_t idx[LNS_MAX_ROWS];
  {
    int m = 0;
    for (int r = 0; r < R; ++r)
      if (role[r]) { key[m] = rng.next(); idx[m++] = (int16_t)r; }
    mark_smallest(key, m, min(prm.num_partial, m), rm);
    for (int i = 0; i < m; ++i)
      if (rm[i]) role[idx[i]] = 2;
  }
  for (int r = 0; r < R; ++r) okind[r] = 0;
  for (int e = 0; e < R * L; ++e) out[e] = 0;
  // partial tours first into a staging area at the END of the output rows (they are emitted after the complete ones)
  int n_part = 0;
  int16_t pbuf[JAMPR_MAX_SLOTS][LNS_MAX_LEN];
  uint8_t plen[JAMPR_MAX_SLOTS];
  for (int r = 0; r < R; ++r) {
    if (role[r] != 2) continue;
    const int n = len[r];
    int mode = prm.rm_partial_mode;
    if (mode == 3) mode = min((int)(rng.next() * 3.0f), 2);
    int m = 0;
    if (mode == 0) {               // random_nodes (destructor.py:141-149)
      int n_rm = prm.frac_partial > 0.0f ? max((int)rintf(prm.frac_partial * (float)n), 1) : rnd_int_1(rng, n);
      n_rm = min(n_rm, n);
      float nk[LNS_MAX_LEN];
      uint8_t nrm[LNS_MAX_LEN];
      for (int q = 0; q < n; ++q) nk[q] = rng.next();
      mark_smallest(nk, n, n_rm, nrm);
      for (int q = 0; q < n; ++q)
        if (!nrm[q]) pbuf[n_part][m++] = plan[r * L + q];
    } else if (mode == 1) {        // random_cut (:151-155)
      m = min(rnd_int_1(rng, n), n);
      for (int q = 0; q < m; ++q) pbuf[n_part][q] = plan[r * L + q];
    } else {                       // const_cut (:157-161): tour[:-cut]
      const int cut = max((int)rintf(prm.frac_partial * (float)n), 1);
      m = max(n - cut, 0);
      for (int q = 0; q < m; ++q) pbuf[n_part][q] = plan[r * L + q];
    }
    plen[n_part++] = (uint8_t)m;
  }
  // complete tours: drop some (destructor.py:186-205)
  int n_c = 0;
  for (int r = 0; r < R; ++r) n_c += role[r] == 1;
  if (prm.frac_complete > 0.0f && n_c > 0) {
    const int n_rm = max((int)rintf(prm.frac_complete * (float)n_c), 1);
    int m = 0;
    for (int r = 0; r < R; ++r)
      if (role[r] == 1) { key[m] = prm.rm_complete_mode == 0 ? rng.next() : (float)len[r]; idx[m++] = (int16_t)r; }
    mark_smallest(key, m, n_rm, rm);
    for (int i = 0; i < m; ++i)
      if (rm[i]) role[idx[i]] = 3;            // removed complete tour
  }
  int t = 0;
  for (int r = 0; r < R; ++r) {
    if (role[r] != 1) continue;
    for (int q = 0; q < len[r]; ++q) out[t * L + q] = plan[r * L + q];
    okind[t++] = 1;
  }
  for (int i = 0; i < n_part; ++i) {
    for (int q = 0; q < plen[i]; ++q) out[t * L + q] = pbuf[i][q];
    okind[t++] = 2;               // may be empty or a singleton: import drops those (env.py:628)
  }
}

// env.py:708-718 in fp32 with the reference's operation order: tm += dist; tm += (max(tw_start - tm, 0) + service)
__device__ __forceinline__ float recompute_leg(float tm, const float* dist, const float* tw, float srv, int N, int prev, int n) {
  tm = tm + dist[(size_t)prev * N + n];
  tm = tm + (fmaxf(tw[2 * n] - tm, 0.0f) + srv);
  return tm;
}

__global__ void env_import_tours_kernel(jampr_dims_t d, jampr_instances_t p, jampr_state_t s,
                                        const int16_t* __restrict__ tours, const uint8_t* __restrict__ kind, int T, int Lt,
                                        const int32_t* __restrict__ src_index, const float* __restrict__ total_in) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int bs = d.n_inst * d.n_samples;
  if (b >= bs) return;
  const int ins = b / d.n_samples, K = d.n_slots, N = d.n_nodes, L = d.plan_len, R = d.k_max;
  const int src = src_index ? src_index[b] : b;
  const float* dist = p.dist + (size_t)ins * N * N;
  const float* tw = p.tw + (size_t)ins * N * 2;
  const float* dem = p.demand + (size_t)ins * N;
  const float srv = p.service[ins];
  int16_t* plan = s.plan + (size_t)b * R * L;
  uint8_t* vfl = s.vflags + (size_t)b * R;
  for (int e = 0; e < R * L; ++e) plan[e] = 0;
  for (int r = 0; r < R; ++r) vfl[r] = 0;
  for (int n = 0; n < N; ++n) {
    s.visited[(size_t)b * N + n] = 0;
    s.pred[(size_t)b * N + n] = 0;
    s.succ[(size_t)b * N + n] = 0;
  }
  for (int t = 0; t < K; ++t) {
    const size_t bk = (size_t)b * K + t;
    s.row[bk] = 0; s.nxt[bk] = 0; s.cur[bk] = 0; s.row_last[bk] = 0;
    s.cap[bk] = 1.0f; s.time[bk] = 0.0f; s.cttd[bk] = 0.0f;
  }
  int t_idx = 0, n_part = 0, status = 0;
  double sum = 0.0;              // the reference adds Python floats (env.py:694) and narrows once
  for (int tt = 0; tt < T; ++tt) {
    const int kd = kind[(size_t)src * T + tt];
    if (kd == 0) continue;
    const int16_t* tr = tours + ((size_t)src * T + tt) * Lt;
    int len = 0;
    while (len < Lt && tr[len] != 0) ++len;
    if (kd == 2 && len <= 1) continue;                        // singleton / emptied partial tour (env.py:628, 667)
    if (t_idx >= R || len > L || (kd == 1 && len > L - 1) || (kd == 2 && n_part >= K)) { status |= JAMPR_ST_TOUR_OVERFLOW; break; }
    for (int q = 0; q < len; ++q) plan[t_idx * L + q] = tr[q];
    float tm = 0.0f;
    int prev = 0;
    if (kd == 1) {                                            // [0, c1 .. cn, 0]: the depot entries count (env.py:631-638)
      tm = recompute_leg(tm, dist, tw, srv, N, 0, 0);
      for (int q = 0; q < len; ++q) { tm = recompute_leg(tm, dist, tw, srv, N, prev, tr[q]); prev = tr[q]; }
      tm = recompute_leg(tm, dist, tw, srv, N, prev, 0);
      vfl[t_idx] |= 2;
    } else {
      float load = 0.0f;
      for (int q = 0; q < len; ++q) {
        tm = recompute_leg(tm, dist, tw, srv, N, prev, tr[q]);
        prev = tr[q];
        load += dem[tr[q]];
      }
      const size_t bk = (size_t)b * K + n_part;
      s.cur[bk] = prev;
      s.cap[bk] = 1.0f - load;
      s.cttd[bk] = p.ttd[(size_t)ins * N + prev];
      s.time[bk] = tm;
      s.nxt[bk] = len;
      s.row[bk] = t_idx;
      vfl[t_idx] |= 1;
      ++n_part;
    }
    sum += (double)tm;
    ++t_idx;
  }
  for (int t = n_part; t < K; ++t) {                          // env.py:672-679: a fresh vehicle for every unused slot
    int nf = 0;
    for (int r = 0; r < R; ++r)
      if (vfl[r] == 0) { nf = r; break; }
    vfl[nf] |= 1;
    s.row[(size_t)b * K + t] = nf;
  }
  s.total[b] = total_in ? total_in[b] : (float)sum;
  s.cost[b] = 0.0f;
  s.status[b] = status;
}

static int lns_check(const jampr_dims_t* d) {
  if (!d || d->n_inst <= 0 || d->n_samples <= 0 || d->k_max <= 0 || d->k_max > LNS_MAX_ROWS || d->plan_len <= 0 ||
      d->plan_len > LNS_MAX_LEN || d->n_slots <= 0 || d->n_slots > JAMPR_MAX_SLOTS)
    return JAMPR_E_BADARG;
  return 0;
}

extern "C" int jampr_lns_destroy(const jampr_dims_t* d, const int16_t* src_plan, const int32_t* src_index,
                                 const jampr_destroy_t* prm, int16_t* tours, uint8_t* kind, void* stream) {
  if (lns_check(d) || !src_plan || !src_index || !prm || !tours || !kind) return JAMPR_E_BADARG;
  if (prm->num_partial < 0 || prm->num_partial > JAMPR_MAX_SLOTS || prm->rm_partial_mode < 0 || prm->rm_partial_mode > 3 ||
      prm->rm_complete_mode < 0 || prm->rm_complete_mode > 1 || prm->frac_partial < 0.0f || prm->frac_partial >= 1.0f ||
      prm->frac_complete < 0.0f || prm->frac_complete >= 1.0f || (prm->rm_partial_mode == 2 && prm->frac_partial <= 0.0f))
    return JAMPR_E_BADARG;
  const int bs = d->n_inst * d->n_samples;
  lns_destroy_kernel<<<(bs + 31) / 32, 32, 0, (cudaStream_t)stream>>>(*d, src_plan, src_index, *prm, tours, kind);
  return launch_status();
}

extern "C" int jampr_env_import_tours(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                                      const int16_t* tours, const uint8_t* kind, int32_t n_tours, int32_t tour_len,
                                      const int32_t* src_index, const float* total_in, void* stream) {
  if (lns_check(d) || !inst || !st || !tours || !kind || n_tours <= 0 || tour_len <= 0) return JAMPR_E_BADARG;
  const int bs = d->n_inst * d->n_samples;
  env_import_tours_kernel<<<(bs + 31) / 32, 32, 0, (cudaStream_t)stream>>>(*d, *inst, *st, tours, kind, n_tours, tour_len,
                                                                          src_index, total_in);
  return launch_status();
}
