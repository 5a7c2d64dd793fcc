// Fused environment kernels: transition + tour buffers + vehicle hand-over (env_move, one thread per rollout),
// termination + candidate sets + feasibility masks (env_observe, one warp per rollout).  Per-rollout state is SoA
// in HBM (include/jampr_b200.h), static tables are per instance (L2 resident across the samples of an instance).
// Integer/compare work on ~1.5 KB of algorithmic bytes per rollout and step; each kernel is organised as a short
// chain of dependent round trips with all independent loads of a link issued together.
// Reference semantics: env.py:203-311 (step), :1075-1131 (_update), :848-907 (get_node_nbh),
// :932-1022 (_get_nbh_and_mask), :1133-1170 (_return_all), :1221-1297 (_reset_pomo).
#include "common.cuh"

#define ENV_WARPS 4
#define ENV_THREADS (ENV_WARPS * 32)

// ---- one transition, executed by one thread (A.4 steps 1-5). Returns the step cost.
// Every load is issued before the first store: the kernel is a chain of dependent HBM round trips, and the only
// true dependency is action -> slot state -> dist[prev][node].
__device__ float env_apply_move(const jampr_dims_t& d, const jampr_instances_t& p, const jampr_state_t& s,
                                int b, int ins, int slot, int node, int* nvis_after = nullptr) {
  const int N = d.n_nodes, K = d.n_slots, L = d.plan_len;
  const size_t bk = (size_t)b * K + slot, in = (size_t)ins * N + node, bn = (size_t)b * N + node;
  const int prev = s.cur[bk];
  const float cap0 = s.cap[bk], t0 = s.time[bk], cttd0 = s.cttd[bk];
  const int r = s.row[bk], pos = s.nxt[bk], last = s.row_last[bk];
  const float dem = p.demand[in], tw_open = p.tw[in * 2], ttd_new = p.ttd[in], service = p.service[ins];
  const uint8_t seen = s.visited[bn];
  const int nvis0 = s.nvis[b];
  float delta = p.dist[((size_t)ins * N + prev) * N + node];
  s.cur[bk] = node;
  s.cap[bk] = __fsub_rn(cap0, dem);
  const float arrival = __fadd_rn(t0, delta);
  if (node != 0) {
    const float wait = fmaxf(__fsub_rn(tw_open, arrival), 0.0f);
    delta = __fadd_rn(delta, __fadd_rn(wait, service));
  }
  s.time[bk] = __fadd_rn(t0, delta);
  const float cost = __fadd_rn(delta, __fsub_rn(ttd_new, cttd0));
  s.cttd[bk] = ttd_new;
  if (pos >= L) {
    s.status[b] |= JAMPR_ST_TOUR_OVERFLOW;
  } else {
    s.plan[((size_t)b * d.k_max + r) * L + pos] = (int16_t)node;
  }
  if (node != 0) {
    s.nxt[bk] = pos + 1;
    s.pred[bn] = (int16_t)last;
    if (last != 0) s.succ[(size_t)b * N + last] = (int16_t)node;
    s.row_last[bk] = node;
    if (!seen) {
      s.visited[bn] = 1;
      s.nvis[b] = nvis0 + 1;
    }
  }
  if (nvis_after) *nvis_after = nvis0 + ((node != 0 && !seen) ? 1 : 0);
  return cost;
}

__device__ __forceinline__ int env_next_free(const jampr_dims_t& d, const jampr_state_t& s, int b) {
  // first tour row that is neither finished nor active, 0 if none (env.py:1035-1040)
  const uint8_t* f = s.vflags + (size_t)b * d.k_max;
  for (int v = 0; v < d.k_max; ++v)
    if (f[v] == 0) return v;
  return 0;
}

// ---- observation: candidate sets + masks for the K slots of rollout b (whole warp).
// Written for memory-level parallelism: a warp's latency is a chain of three HBM/L2 round trips, each link issuing
// all of its independent loads together:
//   link 1 (env_obs_prefetch, issued by the kernels next to their control-word loads): counters, slot state,
//          vehicle flags, the first 128 entries of the depot order;
//   link 2: visited flags of those entries (depot scan) + kNN candidate ids of the vehicles en route;
//   link 3: the per-candidate gathers (visited, demand, distance, window close) of four slots at a time.
struct ObsPre {
  int nvis, cur, nfin, first_free;
  float cap, tm;
  int ord[4];
};

__device__ __forceinline__ ObsPre env_obs_prefetch(const jampr_dims_t& d, const jampr_instances_t& p,
                                                   const jampr_state_t& s, int b, int ins, int lane) {
  const int N = d.n_nodes, K = d.n_slots;
  ObsPre o;
  o.nvis = s.nvis[b];
  o.cur = 1;
  o.cap = 0.0f;
  o.tm = 0.0f;
  if (lane < K) {
    const size_t bk = (size_t)b * K + lane;
    o.cur = s.cur[bk];
    o.cap = s.cap[bk];
    o.tm = s.time[bk];
  }
  const int16_t* ord = p.order + (size_t)ins * N;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int q = 32 * c + lane;
    o.ord[c] = (q < N) ? (int)ord[q] : 0;
  }
  const uint8_t* f = s.vflags + (size_t)b * d.k_max;
  o.nfin = 0;
  o.first_free = 0x7fffffff;
  for (int v = lane; v < d.k_max; v += 32) {
    const uint8_t fl = f[v];
    o.nfin += (fl >> 1) & 1;
    if (fl == 0) o.first_free = min(o.first_free, v);
  }
  return o;
}

// G = slots handled together (compile-time width of the register arrays): min(K, 4)
template <int G>
__device__ __forceinline__ void env_obs_run_g(const jampr_dims_t& d, const jampr_instances_t& p, const jampr_state_t& s,
                                              int b, int ins, int lane, int16_t* dep /* smem [k] */, const ObsPre& o) {
  const int N = d.n_nodes, K = d.n_slots, k = d.k_nbh;
  // per-warp base pointers once, 32-bit offsets inside the rows
  const uint8_t* vis = s.visited + (size_t)b * N;
  const int16_t* knn = p.knn + (size_t)ins * N * k;
  const float* demand_i = p.demand + (size_t)ins * N;
  const float* dist_i = p.dist + (size_t)ins * N * N;
  const float* tw_i = p.tw + (size_t)ins * N * 2;
  int16_t* nbh_b = s.nbh + (size_t)b * K * k;
  uint8_t* mask_b = s.mask + (size_t)b * K * k;
  const bool allvis = o.nvis == N - 1;
  const bool any_depot = __any_sync(0xffffffffu, lane < K && o.cur == 0);
  // link 2
  int cur0[G], kn0[G];
#pragma unroll
  for (int i = 0; i < G; ++i) {
    cur0[i] = __shfl_sync(0xffffffffu, o.cur, min(i, K - 1));
    kn0[i] = 0;
    if (lane < k && i < K && cur0[i] != 0) kn0[i] = knn[cur0[i] * k + lane];
  }
  if (any_depot) {
    // first k unvisited nodes in depot order (the depot never counts as visited), padded with the
    // earliest visited ones, overall order kept (env.py:863-887); 128 positions per pass
    const int16_t* ord = p.order + (size_t)ins * N;
    const int n_unv = N - o.nvis;
    const int missing = max(k - n_unv, 0);
    const unsigned lt = (1u << lane) - 1u;
    int ru = 0, rv = 0, outp = 0;
    for (int base = 0; base < N && outp < k; base += 128) {
      int node[4];
      bool unv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int q = base + 32 * c + lane;
        node[c] = (base == 0) ? o.ord[c] : ((q < N) ? (int)ord[q] : 0);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) unv[c] = vis[node[c]] == 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool valid = base + 32 * c + lane < N;
        const unsigned mu = __ballot_sync(0xffffffffu, valid && unv[c]);
        const unsigned mv = __ballot_sync(0xffffffffu, valid && !unv[c]);
        const int my_ru = ru + __popc(mu & lt) + 1;   // inclusive rank among unvisited
        const int my_rv = rv + __popc(mv & lt) + 1;
        const bool pick = valid && ((unv[c] && my_ru <= k) || (!unv[c] && my_rv <= missing));
        const unsigned mp = __ballot_sync(0xffffffffu, pick);
        if (pick) {
          const int q = outp + __popc(mp & lt);
          if (q < k) dep[q] = (int16_t)node[c];
        }
        ru += __popc(mu);
        rv += __popc(mv);
        outp += __popc(mp);
      }
    }
    __syncwarp();
  }
  bool mask_depot = false;
  if (!allvis) {
    const int nfin = warp_sum_i(o.nfin);
    int first_free = o.first_free;
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) first_free = min(first_free, __shfl_xor_sync(0xffffffffu, first_free, sh));
    // first tour row that is neither finished nor active, if any and not row 0 (env.py:1035-1040)
    mask_depot = (nfin < K) || (first_free != 0x7fffffff && first_free != 0);
  }
  unsigned open_bits = 0u;       // bit t: this lane saw a feasible candidate of slot t
  for (int j0 = 0; j0 < k; j0 += 32) {
    const int j = j0 + lane;
    const bool valid = j < k;
    for (int tb = 0; tb < K; tb += G) {
      const bool first = (j0 | tb) == 0;
      int cur[G], n[G];
      float cap[G], tm[G];
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const int t = min(tb + i, K - 1);
        cur[i] = __shfl_sync(0xffffffffu, o.cur, t);
        cap[i] = __shfl_sync(0xffffffffu, o.cap, t);
        tm[i] = __shfl_sync(0xffffffffu, o.tm, t);
      }
#pragma unroll
      for (int i = 0; i < G; ++i) {
        n[i] = 0;
        if (valid && tb + i < K) {
          if (cur[i] == 0) n[i] = dep[j];
          else n[i] = first ? kn0[i] : (int)knn[cur[i] * k + j];
        }
      }
      // link 3
      uint8_t vs[G];
      float dm[G], dr[G], tc[G];
#pragma unroll
      for (int i = 0; i < G; ++i) {
        vs[i] = vis[n[i]];
        dm[i] = demand_i[n[i]];
        dr[i] = dist_i[cur[i] * N + n[i]];
        tc[i] = tw_i[n[i] * 2 + 1];
      }
#pragma unroll
      for (int i = 0; i < G; ++i) {
        if (valid && tb + i < K) {
          const int e = (tb + i) * k + j;
          bool m = vs[i] != 0;
          m |= cap[i] < dm[i];
          m |= __fadd_rn(tm[i], dr[i]) > tc[i];
          if (cur[i] == 0 && j == 0 && mask_depot) m = true;
          nbh_b[e] = (int16_t)n[i];
          mask_b[e] = m ? 1 : 0;
          if (!m) open_bits |= 1u << (tb + i);
        }
      }
    }
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) open_bits |= __shfl_xor_sync(0xffffffffu, open_bits, sh);
  if (lane == 0 && open_bits != ((1u << K) - 1u)) s.status[b] |= JAMPR_ST_NO_FEASIBLE;
}

__device__ void env_obs_run(const jampr_dims_t& d, const jampr_instances_t& p, const jampr_state_t& s,
                            int b, int ins, int lane, int16_t* dep /* smem [k] */, const ObsPre& o) {
  switch (d.n_slots) {
    case 1: env_obs_run_g<1>(d, p, s, b, ins, lane, dep, o); break;
    case 2: env_obs_run_g<2>(d, p, s, b, ins, lane, dep, o); break;
    case 3: env_obs_run_g<3>(d, p, s, b, ins, lane, dep, o); break;
    default: env_obs_run_g<4>(d, p, s, b, ins, lane, dep, o); break;
  }
}

__device__ __forceinline__ void env_observe(const jampr_dims_t& d, const jampr_instances_t& p, const jampr_state_t& s,
                                            int b, int ins, int lane, int16_t* dep) {
  const ObsPre o = env_obs_prefetch(d, p, s, b, ins, lane);
  env_obs_run(d, p, s, b, ins, lane, dep, o);
}

// ---- reset: zero slot state, K active vehicles, forced start moves, first observation
__global__ void __launch_bounds__(ENV_THREADS) env_reset_kernel(jampr_dims_t d, jampr_instances_t p,
                                                                jampr_state_t s,
                                                                const int32_t* __restrict__ start, int n_start) {
  extern __shared__ int16_t sdep[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * ENV_WARPS + warp;
  const int bs = d.n_inst * d.n_samples;
  if (b >= bs) return;
  const int ins = b / d.n_samples, K = d.n_slots;
  if (lane == 0) {
    for (int t = 0; t < K; ++t) {
      const size_t bk = (size_t)b * K + t;
      s.row[bk] = t;
      s.nxt[bk] = 0;
      s.cur[bk] = 0;
      s.row_last[bk] = 0;
      s.cap[bk] = 1.0f;
      s.time[bk] = 0.0f;
      s.cttd[bk] = 0.0f;
      s.vflags[(size_t)b * d.k_max + t] = 1;
    }
    s.nvis[b] = 0;
    s.status[b] = 0;
    s.allvis[b] = 0;
    float cost = 0.0f;
    for (int t = 0; t < n_start; ++t)
      cost = __fadd_rn(cost, env_apply_move(d, p, s, b, ins, t, start[(size_t)b * n_start + t]));
    s.total[b] = cost;
    s.cost[b] = cost;
    if (s.nvis[b] == d.n_nodes - 1) {   // degenerate tiny graphs
      s.allvis[b] = 1;
      atomicAdd(s.ctl + JAMPR_CTL_N_ALLVIS, 1);
    }
  }
  __syncwarp();
  env_observe(d, p, s, b, ins, lane, sdep + warp * d.k_nbh);
}

// ---- step part 1: transition + hand-over (lane 0), counts newly completed rollouts
__global__ void __launch_bounds__(ENV_THREADS) env_move_kernel(jampr_dims_t d, jampr_instances_t p,
                                                               jampr_state_t s,
                                                               const int32_t* __restrict__ action, int step_no) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int bs = d.n_inst * d.n_samples;
  if (b >= bs) return;
  // control words and the rollout's action in one round trip
  const int c_done = s.ctl[JAMPR_CTL_DONE_STEP], c_hold = s.ctl[JAMPR_CTL_HOLD], c_fin = s.ctl[JAMPR_CTL_FINAL_STEP],
            c_ld = s.ctl[JAMPR_CTL_LOCAL_DONE];
  const int2 act = *reinterpret_cast<const int2*>(action + 2 * (size_t)b);
  const bool before = s.allvis[b] != 0;
  if (!jampr_step_active_w(c_done, c_hold, c_fin, c_ld, step_no)) return;
  const int ins = b / d.n_samples, K = d.n_slots;
  const int slot = act.x, node = act.y;
  int nvis_now;
  s.cost[b] = env_apply_move(d, p, s, b, ins, slot, node, &nvis_now);
  if (!before && nvis_now == d.n_nodes - 1) {
    s.allvis[b] = 1;
    atomicAdd(s.ctl + JAMPR_CTL_N_ALLVIS, 1);
  }
  if (node == 0 && !before) {
    const int nf = env_next_free(d, s, b);
    if (nf == 0) {
      s.status[b] |= JAMPR_ST_FLEET_EXHAUSTED;
    } else {
      const size_t bk = (size_t)b * K + slot;
      const int old = s.row[bk];
      s.cap[bk] = 1.0f;
      s.time[bk] = 0.0f;
      s.vflags[(size_t)b * d.k_max + old] = 2;
      s.vflags[(size_t)b * d.k_max + nf] = 1;
      s.nxt[bk] = 0;
      s.row[bk] = nf;
      s.row_last[bk] = 0;
    }
  }
}

// ---- step part 2: termination (needs the batch-wide completed count), totals, next observation
__global__ void __launch_bounds__(ENV_THREADS, 8) env_observe_kernel(jampr_dims_t d, jampr_instances_t p,
                                                                  jampr_state_t s, int step_no) {
  extern __shared__ int16_t sdep[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * ENV_WARPS + warp;
  const int bs = d.n_inst * d.n_samples;
  if (b >= bs) return;
  const int ins = b / d.n_samples, K = d.n_slots, N = d.n_nodes;
  // link 1: control words, running totals and the rollout's slot state in one round trip
  const int c_done = s.ctl[JAMPR_CTL_DONE_STEP], c_hold = s.ctl[JAMPR_CTL_HOLD], c_fin = s.ctl[JAMPR_CTL_FINAL_STEP],
            c_ld = s.ctl[JAMPR_CTL_LOCAL_DONE], c_all = s.ctl[JAMPR_CTL_N_ALLVIS];
  const float total0 = s.total[b], cost0 = s.cost[b];
  const ObsPre o = env_obs_prefetch(d, p, s, b, ins, lane);
  if (!jampr_step_active_w(c_done, c_hold, c_fin, c_ld, step_no)) return;
  const bool local_done = (c_all == bs) || (step_no >= N + d.k_max + 1);
  bool done = local_done;
  if (c_hold != 0) {
    const int fin = c_fin;
    if (fin < 0) {
      // sharded run, union finishing step not known yet: park here without finalising or observing
      if (local_done) {
        if (b == 0 && lane == 0) s.ctl[JAMPR_CTL_LOCAL_DONE] = step_no;
        return;
      }
    } else {
      done = step_no >= fin;
    }
  }
  if (done) {
    // every vehicle still en route drives home (env.py:1133-1170); unvisited customers become
    // singleton tours at 2 * time_to_depot (env.py:287-295)
    float added = 0.0f;
    if (lane == 0) {
      for (int t = 0; t < K; ++t) {
        const size_t bk = (size_t)b * K + t;
        const int c = s.cur[bk];
        const float back = p.dist[((size_t)ins * N + c) * N];
        s.time[bk] = __fadd_rn(s.time[bk], back);
        s.cur[bk] = 0;
        s.cttd[bk] = 0.0f;
        added = (t == 0) ? back : __fadd_rn(added, back);
      }
    }
    float pen = 0.0f;
    if (s.nvis[b] != N - 1) {
      const uint8_t* vis = s.visited + (size_t)b * N;
      for (int u = 1 + lane; u < N; u += 32)
        if (!vis[u]) pen += 2.0f * p.ttd[(size_t)ins * N + u];
      pen = warp_sum(pen);
    }
    if (lane == 0) {
      float c = __fadd_rn(s.cost[b], added);
      if (s.nvis[b] != N - 1) c = __fadd_rn(c, pen);
      s.cost[b] = c;
      s.total[b] = __fadd_rn(s.total[b], c);
      if (b == 0) s.ctl[JAMPR_CTL_DONE_STEP] = step_no;
    }
    return;
  }
  if (lane == 0) s.total[b] = __fadd_rn(total0, cost0);
  env_obs_run(d, p, s, b, ins, lane, sdep + warp * d.k_nbh, o);
}

// ---- import_sol tail: derive visited / links / counters from the tour buffers
__global__ void __launch_bounds__(ENV_THREADS) env_import_kernel(jampr_dims_t d, jampr_instances_t p,
                                                                 jampr_state_t s, int32_t* max_step_out) {
  extern __shared__ int16_t sdep[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * ENV_WARPS + warp;
  const int bs = d.n_inst * d.n_samples;
  if (b >= bs) return;
  const int ins = b / d.n_samples, K = d.n_slots, N = d.n_nodes, L = d.plan_len;
  if (lane == 0) {
    int nvis = 0, nfin = 0;
    for (int r = 0; r < d.k_max; ++r) {
      const uint8_t fl = s.vflags[(size_t)b * d.k_max + r];
      nfin += (fl >> 1) & 1;
      const int16_t* row = s.plan + ((size_t)b * d.k_max + r) * L;
      int prev = 0;
      for (int q = 0; q < L; ++q) {
        const int n = row[q];
        if (n == 0) break;
        s.visited[(size_t)b * N + n] = 1;
        s.pred[(size_t)b * N + n] = (int16_t)prev;
        if (prev != 0) s.succ[(size_t)b * N + prev] = (int16_t)n;
        prev = n;
        ++nvis;
      }
      for (int t = 0; t < K; ++t)
        if (s.row[(size_t)b * K + t] == r && (fl & 1)) s.row_last[(size_t)b * K + t] = prev;
    }
    s.nvis[b] = nvis;            // s.status was set by whoever filled the buffers (jampr_env_import_tours or the caller)
    s.allvis[b] = (nvis == N - 1);
    if (nvis == N - 1) atomicAdd(s.ctl + JAMPR_CTL_N_ALLVIS, 1);
    atomicMax(max_step_out, nvis + nfin);
  }
  __syncwarp();
  env_observe(d, p, s, b, ins, lane, sdep + warp * d.k_nbh);
}

// ---- best-of-samples (training.py:80-94): first minimum over the P rollouts of an instance
__global__ void gather_best_kernel(jampr_dims_t d, jampr_state_t s, float* best_cost, int32_t* best_idx,
                                   int16_t* best_plan) {
  const int i = blockIdx.x;
  const int P = d.n_samples;
  __shared__ int sbest;
  if (threadIdx.x == 0) {
    float bc = s.total[(size_t)i * P];
    int bi = 0;
    for (int q = 1; q < P; ++q) {
      const float c = s.total[(size_t)i * P + q];
      if (c < bc) { bc = c; bi = q; }
    }
    best_cost[i] = bc;
    best_idx[i] = bi;
    sbest = bi;
  }
  __syncthreads();
  const size_t words = (size_t)d.k_max * d.plan_len;
  const int16_t* src = s.plan + ((size_t)i * P + sbest) * words;
  for (size_t e = threadIdx.x; e < words; e += blockDim.x) best_plan[(size_t)i * words + e] = src[e];
}

// ---- batch-independent tour cost, one thread per rollout
__global__ void true_cost_kernel(jampr_dims_t d, jampr_instances_t p, jampr_state_t s, float* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.n_inst * d.n_samples) return;
  const int ins = b / d.n_samples, N = d.n_nodes, L = d.plan_len;
  const float sv = p.service[ins];
  float tot = 0.0f;
  for (int r = 0; r < d.k_max; ++r) {
    const int16_t* row = s.plan + ((size_t)b * d.k_max + r) * L;
    float t = 0.0f;
    int prev = 0;
    for (int q = 0; q < L; ++q) {
      const int n = row[q];
      if (n == 0) break;
      t += p.dist[((size_t)ins * N + prev) * N + n];
      t += fmaxf(p.tw[((size_t)ins * N + n) * 2] - t, 0.0f) + sv;
      prev = n;
    }
    if (prev != 0) tot += t + p.dist[((size_t)ins * N + prev) * N];
  }
  for (int u = 1; u < N; ++u)
    if (!s.visited[(size_t)b * N + u]) tot += 2.0f * p.ttd[(size_t)ins * N + u];
  out[b] = tot;
}

static int check_dims(const jampr_dims_t* d) {
  if (!d || d->n_inst <= 0 || d->n_samples <= 0 || d->n_nodes < 2 || d->n_slots <= 0 ||
      d->n_slots > JAMPR_MAX_SLOTS || d->k_nbh < 2 || d->k_nbh > d->n_nodes || d->k_max < d->n_slots ||
      d->plan_len <= 0 || d->upd_step <= 0)
    return JAMPR_E_BADARG;
  return 0;
}

extern "C" int jampr_env_reset(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                               const int32_t* start_nodes, int32_t n_start, void* stream) {
  if (check_dims(d) || !inst || !st || n_start < 0 || n_start > d->n_slots || (n_start > 0 && !start_nodes))
    return JAMPR_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bs = (size_t)d->n_inst * d->n_samples, N = d->n_nodes;
  CUDA_RET(cudaMemsetAsync(st->visited, 0, bs * N, s));
  CUDA_RET(cudaMemsetAsync(st->pred, 0, bs * N * sizeof(int16_t), s));
  CUDA_RET(cudaMemsetAsync(st->succ, 0, bs * N * sizeof(int16_t), s));
  CUDA_RET(cudaMemsetAsync(st->plan, 0, bs * d->k_max * d->plan_len * sizeof(int16_t), s));
  CUDA_RET(cudaMemsetAsync(st->vflags, 0, bs * d->k_max, s));
  CUDA_RET(cudaMemsetAsync(st->ctl, 0, sizeof(int32_t) * JAMPR_CTL_WORDS, s));
  CUDA_RET(cudaMemsetAsync(st->ctl + JAMPR_CTL_DONE_STEP, 0xff, sizeof(int32_t), s));
  CUDA_RET(cudaMemsetAsync(st->ctl + JAMPR_CTL_LOCAL_DONE, 0xff, 2 * sizeof(int32_t), s));   // LOCAL_DONE, FINAL_STEP = -1
  const int grid = (int)((bs + ENV_WARPS - 1) / ENV_WARPS);
  env_reset_kernel<<<grid, ENV_THREADS, ENV_WARPS * d->k_nbh * sizeof(int16_t), s>>>(*d, *inst, *st,
                                                                                     start_nodes, n_start);
  return launch_status();
}

extern "C" int jampr_env_step(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                              const int32_t* action, int32_t step_no, void* stream) {
  if (check_dims(d) || !inst || !st) return JAMPR_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bs = (size_t)d->n_inst * d->n_samples;
  const int32_t* act = action ? action : st->action;
  env_move_kernel<<<(int)((bs + 127) / 128), 128, 0, s>>>(*d, *inst, *st, act, step_no);
  const int grid = (int)((bs + ENV_WARPS - 1) / ENV_WARPS);
  env_observe_kernel<<<grid, ENV_THREADS, ENV_WARPS * d->k_nbh * sizeof(int16_t), s>>>(*d, *inst, *st, step_no);
  return launch_status();
}

extern "C" int jampr_env_resume(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                                int32_t step_no, void* stream) {
  if (check_dims(d) || !inst || !st) return JAMPR_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bs = (size_t)d->n_inst * d->n_samples;
  const int grid = (int)((bs + ENV_WARPS - 1) / ENV_WARPS);
  env_observe_kernel<<<grid, ENV_THREADS, ENV_WARPS * d->k_nbh * sizeof(int16_t), s>>>(*d, *inst, *st, step_no);
  return launch_status();
}

extern "C" int jampr_env_import_finish(const jampr_dims_t* d, const jampr_instances_t* inst,
                                       const jampr_state_t* st, int32_t* max_step_out, void* stream) {
  if (check_dims(d) || !inst || !st || !max_step_out) return JAMPR_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bs = (size_t)d->n_inst * d->n_samples, N = d->n_nodes;
  CUDA_RET(cudaMemsetAsync(st->visited, 0, bs * N, s));
  CUDA_RET(cudaMemsetAsync(st->pred, 0, bs * N * sizeof(int16_t), s));
  CUDA_RET(cudaMemsetAsync(st->succ, 0, bs * N * sizeof(int16_t), s));
  CUDA_RET(cudaMemsetAsync(st->ctl, 0, sizeof(int32_t) * JAMPR_CTL_WORDS, s));
  CUDA_RET(cudaMemsetAsync(st->ctl + JAMPR_CTL_DONE_STEP, 0xff, sizeof(int32_t), s));
  CUDA_RET(cudaMemsetAsync(st->ctl + JAMPR_CTL_LOCAL_DONE, 0xff, 2 * sizeof(int32_t), s));
  CUDA_RET(cudaMemsetAsync(max_step_out, 0, sizeof(int32_t), s));
  const int grid = (int)((bs + ENV_WARPS - 1) / ENV_WARPS);
  env_import_kernel<<<grid, ENV_THREADS, ENV_WARPS * d->k_nbh * sizeof(int16_t), s>>>(*d, *inst, *st, max_step_out);
  return launch_status();
}

extern "C" int jampr_gather_best(const jampr_dims_t* d, const jampr_state_t* st, float* best_cost,
                                 int32_t* best_idx, int16_t* best_plan, void* stream) {
  if (check_dims(d) || !st || !best_cost || !best_idx || !best_plan) return JAMPR_E_BADARG;
  gather_best_kernel<<<d->n_inst, 128, 0, (cudaStream_t)stream>>>(*d, *st, best_cost, best_idx, best_plan);
  return launch_status();
}

extern "C" int jampr_true_cost(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                               float* out, void* stream) {
  if (check_dims(d) || !inst || !st || !out) return JAMPR_E_BADARG;
  const int bs = d->n_inst * d->n_samples;
  true_cost_kernel<<<(bs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*d, *inst, *st, out);
  return launch_status();
}

extern "C" int jampr_abi_version(void) { return JAMPR_ABI_VERSION; }
