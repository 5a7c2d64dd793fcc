// Device local search on tour buffers (sm_100a): the stand-in for the reference's OR-tools guided local search
// (lib/lns/gort.py:129-290, lib/utils/or_utils.py:10-171), which is neither installed nor pinned.  One CTA improves one
// solution by best-improvement descent over two classic VRPTW neighbourhoods:
//   relocate   move one customer to any position of any tour (or to an empty tour row)
//   2-opt*     exchange the tails of two tours
// Every candidate move is evaluated by re-simulating the one or two tours it touches with the ENVIRONMENT's own
// arithmetic (env.py:1081-1131 in fp32: capacity test cur_cap < demand, window test time + dist > tw_end, waiting and
// service), so a solution the search accepts is exactly a sequence of moves the environment would have allowed, and its
// cost is the batch-independent tour cost of jampr_true_cost (travel + waiting + service per tour incl. the return
// leg).  Unserved customers stay unserved (their 2 * time_to_depot penalty is constant under these moves).
// Deterministic: the best move is the smallest (delta, move id) pair; oracle/ls_oracle.py restates it on the CPU.
#include "common.cuh"

#define LS_THREADS 256
#define LS_MAX_ROWS 256
#define LS_EPS 1e-6f

struct LsTour {            // one tour row being simulated: up to two segments of stored rows + one inserted node
  const int16_t* a; int a0, a1;     // nodes a[a0 .. a1)
  int ins;                          // node inserted after the first segment (0 = none)
  const int16_t* b; int b0, b1;     // then b[b0 .. b1)
  int skip;                         // absolute index in segment a / b storage to skip (-1 = none), applies to `a` row
};

// simulate a tour given as (segment a with an optional skipped index) + optional inserted node + segment b.
// returns the tour cost (duration incl. return leg) or +inf when infeasible; empty tour -> 0
__device__ float ls_sim(const float* __restrict__ dist, const float* __restrict__ tw, const float* __restrict__ dem,
                        float srv, int N, const int16_t* a, int a0, int a1, int skip_a, int ins, const int16_t* b, int b0,
                        int b1, int skip_b) {
  float tm = 0.0f, cap = 1.0f;
  int prev = 0, cnt = 0;
  auto visit = [&](int n) -> bool {
    const float d = dist[(size_t)prev * N + n];
    if (cap < dem[n]) return false;                     // env.py:981
    if (tm + d > tw[2 * n + 1]) return false;           // env.py:985-990
    cap -= dem[n];
    const float arrive = tm + d;
    tm = tm + (d + (fmaxf(tw[2 * n] - arrive, 0.0f) + srv));     // env.py:1096-1103
    prev = n;
    ++cnt;
    return true;
  };
  for (int q = a0; q < a1; ++q)
    if (q != skip_a && !visit(a[q])) return INFINITY;
  if (ins > 0 && !visit(ins)) return INFINITY;
  for (int q = b0; q < b1; ++q)
    if (q != skip_b && !visit(b[q])) return INFINITY;
  if (cnt == 0) return 0.0f;
  return tm + dist[(size_t)prev * N];
}

__global__ void __launch_bounds__(LS_THREADS) local_search_kernel(jampr_dims_t d, jampr_instances_t p, int16_t* __restrict__ plans,
                                                                  const int32_t* __restrict__ inst_of, int max_iters,
                                                                  float* __restrict__ cost_out, int32_t* __restrict__ moves_out) {
  extern __shared__ uint8_t ls_smem[];
  const int R = d.k_max, L = d.plan_len, N = d.n_nodes;
  const int sol = blockIdx.x, tid = threadIdx.x;
  const int ins_i = inst_of ? inst_of[sol] : sol / d.n_samples;
  int16_t* s_plan = reinterpret_cast<int16_t*>(ls_smem);                       // [R][L]
  int16_t* s_len = s_plan + R * L;                                             // [R]
  float* s_cost = reinterpret_cast<float*>(s_len + ((R + 1) & ~1));            // [R]
  float* s_bd = s_cost + R;                                                    // [LS_THREADS] best delta per thread
  int* s_bm = reinterpret_cast<int*>(s_bd + LS_THREADS);                       // [LS_THREADS] best move id per thread
  int16_t* s_tmp = reinterpret_cast<int16_t*>(s_bm + LS_THREADS);              // [2][L] scratch rows for applying a move
  __shared__ int s_nmoves;
  const float* dist = p.dist + (size_t)ins_i * N * N;
  const float* tw = p.tw + (size_t)ins_i * N * 2;
  const float* dem = p.demand + (size_t)ins_i * N;
  const float srv = p.service[ins_i];
  int16_t* gplan = plans + (size_t)sol * R * L;
  for (int e = tid; e < R * L; e += LS_THREADS) s_plan[e] = gplan[e];
  if (tid == 0) s_nmoves = 0;
  __syncthreads();
  for (int r = tid; r < R; r += LS_THREADS) {
    int l = 0;
    while (l < L && s_plan[r * L + l] != 0) ++l;
    s_len[r] = (int16_t)l;
  }
  __syncthreads();
  for (int it = 0; it < max_iters; ++it) {
    for (int r = tid; r < R; r += LS_THREADS)
      s_cost[r] = ls_sim(dist, tw, dem, srv, N, s_plan + r * L, 0, s_len[r], -1, 0, nullptr, 0, 0, -1);
    __syncthreads();
    // ---- move space.  relocate: id = ((r1 * L + p1) * R + r2) * (L + 1) + p2  (customer at (r1, p1) to before position
    // p2 of row r2); 2-opt*: id = REL + ((r1 * (L + 1) + i) * R + r2) * (L + 1) + j, r1 < r2 (A[:i] + B[j:], B[:j] + A[i:])
    const long long REL = (long long)R * L * R * (L + 1);
    const long long TOT = REL + (long long)R * (L + 1) * R * (L + 1);
    float bd = -LS_EPS;
    long long bm = -1;
    // enumerate only existing (row, position) pairs: flatten over rows with their lengths
    // relocate
    for (int r1 = 0; r1 < R; ++r1) {
      const int l1 = s_len[r1];
      if (l1 == 0) continue;
      bool empty_used = false;
      for (int r2 = 0; r2 < R; ++r2) {
        const int l2 = s_len[r2];
        if (l2 == 0) {                      // one empty row stands for "open a new tour"
          if (empty_used) continue;
          empty_used = true;
        }
        if (l2 >= L - 1 && r2 != r1) continue;      // a complete tour needs its trailing zero (env.py:230-236)
        const int n_pairs = l1 * (l2 + 1);
        for (int e = tid; e < n_pairs; e += LS_THREADS) {
          const int p1 = e / (l2 + 1), p2 = e - p1 * (l2 + 1);
          const int node = s_plan[r1 * L + p1];
          float delta;
          if (r1 == r2) {
            if (p2 == p1 || p2 == p1 + 1) continue;            // same position
            // row without p1, node inserted before original position p2
            const int16_t* a = s_plan + r1 * L;
            const float nc = ls_sim(dist, tw, dem, srv, N, a, 0, p2, p1, node, a, p2, l1, p1);
            delta = nc - s_cost[r1];
          } else {
            const float c1 = ls_sim(dist, tw, dem, srv, N, s_plan + r1 * L, 0, l1, p1, 0, nullptr, 0, 0, -1);
            const float c2 = ls_sim(dist, tw, dem, srv, N, s_plan + r2 * L, 0, p2, -1, node, s_plan + r2 * L, p2, l2, -1);
            delta = (c1 + c2) - (s_cost[r1] + s_cost[r2]);
          }
          const long long id = (((long long)r1 * L + p1) * R + r2) * (L + 1) + p2;
          if (delta < bd || (delta == bd && id < bm)) { bd = delta; bm = id; }
        }
      }
    }
    // 2-opt*
    for (int r1 = 0; r1 < R; ++r1) {
      const int l1 = s_len[r1];
      if (l1 == 0) continue;
      for (int r2 = r1 + 1; r2 < R; ++r2) {
        const int l2 = s_len[r2];
        if (l2 == 0) continue;
        const int n_pairs = (l1 + 1) * (l2 + 1);
        for (int e = tid; e < n_pairs; e += LS_THREADS) {
          const int i = e / (l2 + 1), j = e - i * (l2 + 1);
          if ((i == l1 && j == l2) || (i == 0 && j == 0)) continue;        // no change / pure relabelling
          if (i + (l2 - j) > L - 1 || j + (l1 - i) > L - 1) continue;
          const float c1 = ls_sim(dist, tw, dem, srv, N, s_plan + r1 * L, 0, i, -1, 0, s_plan + r2 * L, j, l2, -1);
          const float c2 = ls_sim(dist, tw, dem, srv, N, s_plan + r2 * L, 0, j, -1, 0, s_plan + r1 * L, i, l1, -1);
          const float delta = (c1 + c2) - (s_cost[r1] + s_cost[r2]);
          const long long id = REL + (((long long)r1 * (L + 1) + i) * R + r2) * (L + 1) + j;
          if (delta < bd || (delta == bd && id < bm)) { bd = delta; bm = id; }
        }
      }
    }
    (void)TOT;
    // ---- block reduction of (delta, id); ids fit 32 bits for the supported sizes (checked on the host)
    s_bd[tid] = bd;
    s_bm[tid] = (int)bm;
    __syncthreads();
    for (int o = LS_THREADS / 2; o > 0; o >>= 1) {
      if (tid < o) {
        const float od = s_bd[tid + o];
        const int om = s_bm[tid + o];
        if (om >= 0 && (s_bm[tid] < 0 || od < s_bd[tid] || (od == s_bd[tid] && om < s_bm[tid]))) { s_bd[tid] = od; s_bm[tid] = om; }
      }
      __syncthreads();
    }
    const int best = s_bm[0];
    if (best < 0) break;                     // local optimum (uniform across the block)
    // ---- apply (thread 0; rows are short)
    if (tid == 0) {
      long long id = (unsigned int)best;
      if (id < REL) {
        const int p2 = (int)(id % (L + 1)); id /= (L + 1);
        const int r2 = (int)(id % R); id /= R;
        const int p1 = (int)(id % L);
        const int r1 = (int)(id / L);
        const int node = s_plan[r1 * L + p1];
        const int l1 = s_len[r1];
        if (r1 == r2) {
          int m = 0;
          for (int q = 0; q < l1; ++q) {
            if (q == p2) s_tmp[m++] = (int16_t)node;
            if (q != p1) s_tmp[m++] = s_plan[r1 * L + q];
          }
          if (p2 == l1) s_tmp[m++] = (int16_t)node;
          for (int q = 0; q < l1; ++q) s_plan[r1 * L + q] = s_tmp[q];
        } else {
          const int l2 = s_len[r2];
          for (int q = p1; q < l1 - 1; ++q) s_plan[r1 * L + q] = s_plan[r1 * L + q + 1];
          s_plan[r1 * L + l1 - 1] = 0;
          for (int q = l2; q > p2; --q) s_plan[r2 * L + q] = s_plan[r2 * L + q - 1];
          s_plan[r2 * L + p2] = (int16_t)node;
          s_len[r1] = (int16_t)(l1 - 1);
          s_len[r2] = (int16_t)(l2 + 1);
        }
      } else {
        id -= REL;
        const int j = (int)(id % (L + 1)); id /= (L + 1);
        const int r2 = (int)(id % R); id /= R;
        const int i = (int)(id % (L + 1));
        const int r1 = (int)(id / (L + 1));
        const int l1 = s_len[r1], l2 = s_len[r2];
        int m1 = 0, m2 = 0;
        for (int q = 0; q < i; ++q) s_tmp[m1++] = s_plan[r1 * L + q];
        for (int q = j; q < l2; ++q) s_tmp[m1++] = s_plan[r2 * L + q];
        for (int q = 0; q < j; ++q) s_tmp[L + m2++] = s_plan[r2 * L + q];
        for (int q = i; q < l1; ++q) s_tmp[L + m2++] = s_plan[r1 * L + q];
        for (int q = 0; q < L; ++q) {
          s_plan[r1 * L + q] = q < m1 ? s_tmp[q] : (int16_t)0;
          s_plan[r2 * L + q] = q < m2 ? s_tmp[L + q] : (int16_t)0;
        }
        s_len[r1] = (int16_t)m1;
        s_len[r2] = (int16_t)m2;
      }
      ++s_nmoves;
    }
    __syncthreads();
  }
  __syncthreads();
  for (int r = tid; r < R; r += LS_THREADS)
    s_cost[r] = ls_sim(dist, tw, dem, srv, N, s_plan + r * L, 0, s_len[r], -1, 0, nullptr, 0, 0, -1);
  __syncthreads();
  for (int e = tid; e < R * L; e += LS_THREADS) gplan[e] = s_plan[e];
  if (tid == 0) {
    float tot = 0.0f;
    for (int r = 0; r < R; ++r) tot += s_cost[r];
    if (cost_out) cost_out[sol] = tot;
    if (moves_out) moves_out[sol] = s_nmoves;
  }
}

extern "C" int jampr_local_search(const jampr_dims_t* d, const jampr_instances_t* inst, int16_t* plans, int32_t n_solutions,
                                  const int32_t* inst_of, int32_t max_iters, float* cost_out, int32_t* moves_out,
                                  void* stream) {
  if (!d || !inst || !plans || n_solutions <= 0 || max_iters < 0 || d->k_max <= 0 || d->k_max > LS_MAX_ROWS ||
      d->plan_len <= 1 || d->n_nodes < 2)
    return JAMPR_E_BADARG;
  const long long R = d->k_max, L = d->plan_len;
  if (R * L * R * (L + 1) + R * (L + 1) * R * (L + 1) > 0x7fffffffLL) return JAMPR_E_TOOLARGE;      // move ids are int32
  const int Rp = (d->k_max + 1) & ~1;
  const size_t smem = (size_t)d->k_max * d->plan_len * 2 + (size_t)Rp * 2 + (size_t)d->k_max * 4 + LS_THREADS * 8 +
                      2 * (size_t)d->plan_len * 2 + 16;
  if (smem > 200 * 1024) return JAMPR_E_TOOLARGE;
  CUDA_RET(cudaFuncSetAttribute(local_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  local_search_kernel<<<n_solutions, LS_THREADS, smem, (cudaStream_t)stream>>>(*d, *inst, plans, inst_of, max_iters, cost_out,
                                                                              moves_out);
  return launch_status();
}
