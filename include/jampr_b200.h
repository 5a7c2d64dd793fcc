/*
 * jampr_b200.h — C ABI of libjampr_b200.so: the B200 (sm_100a) construction-rollout hot path of JAMPR+.
 *
 * The reference (jokofa/JAMPR_plus) has no FFI layer; its boundary for this path is two Python
 * classes.  Each entry point below names the reference code it replaces, with
 * file:line under /root/reference/.  The Python mirror classes in jampr_plus_b200/ (RPEnv, Policy)
 * keep the reference signatures and call these functions through ctypes.
 *
 * Conventions
 *  - Every pointer inside the structs is a DEVICE pointer owned by the caller (torch-allocated); the
 *    library allocates nothing persistent and never synchronises the stream.  It holds no state between calls
 *    except one debugging switch that is read ONCE from the process environment on first use: JAMPR_NODE_ENCODER=fp32
 *    (CUDA-core node encoder on the 16-bit precisions; results within the documented tolerances either way).  Kernel
 *    variants of the policy step are selected per call through jampr_weights_t.tc_variant.
 *  - All calls enqueue work on `stream` (a cudaStream_t passed as void*) and return 0 on success,
 *    a negative JAMPR_E_* code for argument errors, or a positive cudaError_t for launch errors.
 *  - Data-dependent conditions the reference raises on are OR-ed into `status[rollout]`
 *    (JAMPR_ST_*), the caller turns them into the reference's RuntimeError / RuntimeWarning.
 *  - Rollout index b = instance * num_samples + sample (reference env.py:1189-1202 ordering).
 */
#ifndef JAMPR_B200_H
#define JAMPR_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define JAMPR_API __attribute__((visibility("default")))
#else
#define JAMPR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define JAMPR_ABI_VERSION 5

/* argument errors */
#define JAMPR_E_BADARG   (-1)
#define JAMPR_E_TOOLARGE (-2)   /* graph does not fit the on-chip tile of this build */
#define JAMPR_E_UNSUPPORTED (-3)

/* per-rollout status bits */
#define JAMPR_ST_NO_FEASIBLE     1   /* env.py:1020-1021 "no feasible nodes left" (RuntimeError)  */
#define JAMPR_ST_TOUR_OVERFLOW   2   /* env.py:230-236 tour buffer (64 slots) exceeded (RuntimeError) */
#define JAMPR_ST_FLEET_EXHAUSTED 4   /* env.py:254-258 "used all available vehicles!" (RuntimeWarning) */
#define JAMPR_ST_NAN_LOGITS      8   /* policy.py:208 "Logits contain NANs!" (AssertionError) */
#define JAMPR_ST_TW_COLLAPSE    16   /* env.py:408 assertion in the time-window fix-up (per instance) */

/* decode types (policy.py:229-240) */
#define JAMPR_DECODE_GREEDY   0
#define JAMPR_DECODE_SAMPLING 1

/* control block on the device: int32[JAMPR_CTL_WORDS].  The step counter env._step (env.py:139,309)
 * is host-known and travels as a launch argument; only data-dependent words live here. */
#define JAMPR_CTL_N_ALLVIS    0   /* rollouts whose customers are all visited */
#define JAMPR_CTL_DONE_STEP   1   /* value of env._step at which the episode terminated (env.py:281), -1 = running */
/* Sharded runs (one process per GPU): the reference terminates on the UNION batch (env.py:281) and its reported
 * total_cost depends on that step.  With HOLD set by the caller, a shard that reaches its own
 * finishing condition records the step in LOCAL_DONE and parks (all later launches are no-ops) instead of
 * finalising; the caller all-reduces MAX over the ranks, writes FINAL_STEP, calls jampr_env_resume for the parked
 * step and enqueues the remaining steps; the terminal accounting then happens at FINAL_STEP on every rank. */
#define JAMPR_CTL_HOLD        2   /* in: 1 = park at the local finishing step */
#define JAMPR_CTL_LOCAL_DONE  3   /* out: env._step at which this shard alone would have terminated, -1 = not yet */
#define JAMPR_CTL_FINAL_STEP  4   /* in: finishing step of the union batch, -1 = unknown */
#define JAMPR_CTL_WORDS       8

typedef struct {
  int32_t n_inst;        /* I: instances */
  int32_t n_samples;     /* P: rollouts per instance (env.num_samples) */
  int32_t n_nodes;       /* N: graph size incl. depot */
  int32_t n_slots;       /* K: max_concurrent_vehicles */
  int32_t k_nbh;         /* k: ceil(k_nbh_frac * N) (graph_utils.py:99) */
  int32_t k_max;         /* K_max: int(kv + floor(ln kv)) (env.py:362) */
  int32_t plan_len;      /* L: min(64, N) (env.py:165) */
  int32_t upd_step;      /* tour_graph_update_step (env.py:43) */
} jampr_dims_t;

/* Static data, once per INSTANCE (the reference replicates it per sample, env.py:1189-1202). */
typedef struct {
  const float* coords;     /* in  (I,N,2)  */
  const float* demand;     /* in  (I,N)    */
  float*       tw;         /* in/out (I,N,2): window ends shrunk where return is impossible (env.py:395-409) */
  const float* service;    /* in  (I)      */
  const float* horizon;    /* in  (I) org_service_horizon */
  float*       dist;       /* out (I,N,N) floor(10*||100 dx||)/10/horizon (challenge_utils.py:35, env.py:916-922) */
  float*       ttd;        /* out (I,N) time to depot = dist[:,:,0] */
  int16_t*     knn;        /* out (I,N,k) row u>=1: k-1 nearest customers incl. u, then depot (graph_utils.py:113-136) */
  float*       knn_w;      /* out (I,N,k) dist[u, knn[u,j]] */
  int16_t*     order;      /* out (I,N) nodes by Euclidean distance to depot (env.py:855-861) */
  int32_t*     inst_status;/* out (I) JAMPR_ST_TW_COLLAPSE */
  float*       static_emb; /* out (I,N,256) node-encoder output (node_encoder.py:128-155) */
  float*       h0;         /* out (I,N,128) node_emb_input_proj(static_emb) (tour_encoder.py:170-172), hoisted */
  uint32_t*    tc_tab;     /* out (I,N,k4) k4 = k rounded up to 4: packed edges of the kNN graph for the tensor-core step
                              (source-row offset | 16-bit weight, csrc/policy_tc_common.cuh agg_entry), written by
                              jampr_node_encoder for the precision in use; may be NULL on the fp32 path */
  uint32_t*    tc_dtab;    /* out (I,nd) nd = N rounded up to 8: packed edges node -> depot, same encoding */
  void*        lg_h16;     /* scratch (2,I,Npad,128) 16-bit, Npad = N rounded up to 128: node-encoder ping-pong embedding of
                              the large-graph path (N > 128, csrc/gnn_large.cu); NULL otherwise */
  float*       lg_skip;    /* scratch (I,Npad,128) fp32 skip operand of the same path; NULL otherwise */
  const float* dist_given; /* in  (I,N,N) or NULL: caller-supplied normalised transit matrix copied into `dist`
                              instead of evaluating the DIMACS formula (e.g. a matrix computed by torch on the
                              host: MKL's vsSqrt is not correctly rounded, sqrtf on the device is) */
  float*       h0_t;       /* out (I,32,N,4)  column-chunk-major copy of h0: element [i][c][n][e] = h0[i][n][4c+e].  The
                              single-tile tensor-core step reads these tables with thread = node; in this layout a warp's
                              128-bit loads cover 512 contiguous bytes instead of 32 different lines.  Written by
                              jampr_node_encoder on the 16-bit precisions; required there.  For N > 128 (layer-wise path) the
                              layout is (I,Npad/128,32,128,4), tile-chunk-row: [i][tile][c][r][e] = h0[i][128 tile + r][4c+e] */
  float*       static_emb_t; /* out (I,64,N,4) the same for static_emb (N <= 128 only, NULL otherwise) */
} jampr_instances_t;

/* Episode state, once per ROLLOUT (bs = I*P). SoA, all arrays contiguous. */
typedef struct {
  uint8_t*  visited;   /* (bs,N)        env._visited */
  int16_t*  pred;      /* (bs,N)        predecessor along the tour row (0 = depot) — tour graph, env.py:804-835 */
  int16_t*  succ;      /* (bs,N)        successor along the tour row (0 = depot / none yet) */
  int16_t*  plan;      /* (bs,K_max,L)  env.tour_plan */
  uint8_t*  vflags;    /* (bs,K_max)    bit0 active_vehicles, bit1 _finished */
  int32_t*  row;       /* (bs,K)        env.active_to_plan_idx */
  int32_t*  nxt;       /* (bs,K)        env.next_index_in_tour */
  int32_t*  cur;       /* (bs,K)        env.cur_node */
  int32_t*  row_last;  /* (bs,K)        last customer written to the slot's tour row */
  float*    cap;       /* (bs,K)        env.cur_cap */
  float*    time;      /* (bs,K)        env.cur_time */
  float*    cttd;      /* (bs,K)        env.cur_time_to_depot */
  float*    total;     /* (bs)          env._total */
  float*    cost;      /* (bs)          cost of the last step (return value of env.step) */
  int32_t*  status;    /* (bs)          JAMPR_ST_* */
  uint8_t*  allvis;    /* (bs)          all customers visited */
  int32_t*  nvis;      /* (bs)          number of visited customers */
  int16_t*  nbh;       /* (bs,K,k)      obs.nbh */
  uint8_t*  mask;      /* (bs,K,k)      obs.nbh_mask (1 = infeasible) */
  int32_t*  action;    /* (bs,2)        (slot, node) chosen by the policy */
  float*    ll;        /* (bs)          log-likelihood of the chosen action */
  float*    entropy;   /* (bs)          entropy of the action distribution (sampling) */
  float*    logp;      /* (bs,K*k) or NULL: full log-probabilities (parity / training) */
  float*    h_fin;     /* (bs,N,128)    tour-graph node embedding of the last refresh (tour_encoder.py:198 cache) */
  float*    ne;        /* (bs,N,256)    node_emb = node_emb_output_proj(h) + static_emb (tour_encoder.py:217-218) */
  float*    ne_stats;  /* (bs,2,256)    column mean and max of ne over the N nodes (policy.py:179-181) */
  int32_t*  ctl;       /* (JAMPR_CTL_WORDS) control block */
  void*     hbuf;      /* (2,bs,Npad,128) 16-bit ping-pong tour-graph embedding of the large-graph path (N > 128), or NULL */
  float*    hskip;     /* (bs,Npad,128) fp32 skip operand of the large-graph path, or NULL */
  float*    stats_part;/* (bs,Npad/128,2,256) per-tile column sums / maxima of ne (large-graph path), or NULL */
  /* split decoder of the tensor-core path (csrc/decoder_tc.cu: the decoder's dense stages as tcgen05 GEMMs over 128
   * rollouts per CTA).  All six NULL = the decoder stays fused into the per-rollout policy kernel. */
  void*     ne16;      /* (bs,N,256) 16-bit node_emb rows written by the GNN kernel, gathered by the decoder */
  float*    dq;        /* (bs,512)   decoder query [mean + mean, max + max] (policy.py:179-182) */
  float*    dte;       /* (bs,4,256) tour embeddings of the K <= 4 active slots (tour_encoder.py:202-213) */
  float*    dqp;       /* (bs,4,256) scratch: per-head glimpse query W_gk,h^T ctxt_h */
  float*    dgbar;     /* (bs,4,256) scratch: per-head glimpse-weighted sum of the action embeddings */
  float*    dlp;       /* (bs,256)   scratch: logit query W_lk^T out_proj(glimpse) */
} jampr_state_t;

/* Policy weights: one fp32 blob on the device, packed by jampr_plus_b200/model/packing.py.
 * Offsets (in floats) are given by jampr_weight_offsets(). */
#define JAMPR_PREC_F32  0   /* fp32 CUDA-core path: the 1e-5 parity anchor */
#define JAMPR_PREC_BF16 1   /* bf16 tcgen05 tensor-core GEMMs, fp32 accumulate and epilogues */
#define JAMPR_PREC_F16  2   /* same tensor-core path with fp16 operands (11-bit significand) */

typedef struct {
  const float* blob;       /* fp32 blob, layout = jampr_weight_offsets() */
  int64_t      n_floats;
  const void*  blob_bf16;  /* 16-bit (bf16 or fp16, see precision) operands in UMMA tile order (packing.pack_tc), or NULL */
  int64_t      n_bf16;     /* elements of blob_bf16 */
  int32_t      precision;  /* JAMPR_PREC_* */
  int32_t      tc_variant; /* kernel selection of the 16-bit path for graphs of up to 128 nodes: 0 = automatic (3 when the batch has
                              at least JAMPR_SPLIT_MIN_ROLLOUTS rollouts, else 2), 1 = one-CTA-per-SM tiling with the decoder fused
                              (csrc/policy_tc.cu), 2 = two-CTA tiling with the decoder fused per rollout (csrc/policy_tc2.cu),
                              3 = two-CTA tiling + the tcgen05 decoder over 128 rollouts per CTA (csrc/decoder_tc.cu).  All
                              variants meet the documented tolerances; 3 wins on large batches (weights read once per 128
                              rollouts), 2 on small ones (no latency-bound GEMM chain in a single CTA). */
} jampr_weights_t;
#define JAMPR_SPLIT_MIN_ROLLOUTS 8192

JAMPR_API int jampr_abi_version(void);

/* Number of floats of the packed weight blob and the offset table (names in packing.py).
 * Replaces nothing in the reference; it is the checkpoint -> device contract. */
JAMPR_API int64_t jampr_weight_blob_floats(void);
/* Number of 16-bit elements of the tensor-core operand blob (packing.pack_tc; layout in csrc/policy_tc.cu). */
JAMPR_API int64_t jampr_weight_blob16_elems(void);
JAMPR_API int jampr_weight_offsets(int64_t* out, int32_t max_entries);

/* RPEnv.load_data static part: distance matrix, window fix-up, kNN table, depot order.
 * Replaces env.py:369-409 and the per-instance Python loop env.py:782-794 (+ graph_utils.py:104-141). */
JAMPR_API int jampr_setup_instances(const jampr_dims_t* d, const jampr_instances_t* inst, void* stream);

/* Start nodes of the POMO pseudo-steps (env.py:1221-1270): candidates = depot neighbourhood without the depot, reduced to
 * the n_cand = floor((k - 1) * pomo_nbh_frac) earliest window starts; per rollout n_start distinct uniform picks
 * (counter-based generator keyed by seed; the reference uses torch.multinomial, env.py:1258-1262), or — single_start —
 * the sample-th earliest candidate for slot 0 only (env.py:1229-1242).  start_nodes (bs, n_start) int32 (n_start = 1
 * with single_start), to be handed to jampr_env_reset.  Needs jampr_setup_instances (order, fixed windows). */
JAMPR_API int jampr_pomo_start_nodes(const jampr_dims_t* d, const jampr_instances_t* inst, int32_t n_cand, int32_t n_start,
                           int32_t single_start, uint64_t seed, int32_t* start_nodes, void* stream);

/* NodeEncoder.forward, once per instance (node_encoder.py:128-155), plus the hoisted
 * node_emb_input_proj (tour_encoder.py:170-172). */
JAMPR_API int jampr_node_encoder(const jampr_dims_t* d, const jampr_instances_t* inst,
                       const jampr_weights_t* w, void* stream);

/* RPEnv.reset (env.py:146-201).  start_nodes: (bs, n_start) int32 forced first moves, one per
 * vehicle slot 0..n_start-1 (the POMO pseudo-steps, env.py:1271-1297), or NULL / n_start = 0.
 * Leaves the first observation (nbh, mask) in the state. */
JAMPR_API int jampr_env_reset(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                    const int32_t* start_nodes, int32_t n_start, void* stream);

/* RPEnv.step (env.py:203-311) on st->action (or `action` if not NULL, (bs,2) int32): transition,
 * tour buffers, vehicle hand-over, termination incl. _return_all and singleton penalties
 * (env.py:281-295, 1133-1170), then the next observation (env.py:848-1022).  step_no = env._step
 * before the call.  After termination (ctl[DONE_STEP] >= 0 from an earlier call) it is a no-op. */
JAMPR_API int jampr_env_step(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                   const int32_t* action, int32_t step_no, void* stream);

/* Second half of jampr_env_step (termination test, totals, next observation) for step_no alone: resumes a shard
 * parked by JAMPR_CTL_HOLD once JAMPR_CTL_FINAL_STEP is known. */
JAMPR_API int jampr_env_resume(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                     int32_t step_no, void* stream);

/* RPEnv.import_sol tail (env.py:681-706): state arrays were filled by the caller from solution
 * lists (plan, vflags, row, nxt, cur, cap, time, cttd, total); derives visited/pred/succ/row_last
 * from the tour buffers, writes max_b(visited + finished) (env.py:704) to *max_step_out (device int32)
 * and produces the observation. */
JAMPR_API int jampr_env_import_finish(const jampr_dims_t* d, const jampr_instances_t* inst,
                            const jampr_state_t* st, int32_t* max_step_out, void* stream);

/* RPEnv.import_sol (env.py:575-706, _recompute_cost :708-718) from DENSE tours instead of nested Python lists: tours
 * (n_sol, n_tours, tour_len) int16 customers in visiting order, zero padded; kind (n_sol, n_tours) uint8: 0 = no tour,
 * 1 = complete tour (the list form starts and ends at the depot), 2 = partial tour (open list; fewer than two customers
 * = dropped like the reference drops singletons).  Rollout b reads solution src_index[b] (NULL = b): bootstrap of the
 * selected solutions back to num_samples (lns.py:304-310) and _expand_sample_dimension (env.py:1204-1219) are index maps.
 * Fills plan / vflags / row / nxt / cur / cap / time / cttd / total (total_in (bs) overrides the recomputed totals, the
 * `cost` argument of import_sol); follow with jampr_env_import_finish for visited / links / observation / step. */
JAMPR_API int jampr_env_import_tours(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                           const int16_t* tours, const uint8_t* kind, int32_t n_tours, int32_t tour_len,
                           const int32_t* src_index, const float* total_in, void* stream);

/* The destruction operator on tour buffers — what the reference left as a NotImplemented stub (env.py:720-723, "tensor-based
 * native destruction operator circumventing expensive conversion to lists") — for lib/lns/destructor.py:86-205
 * `random_from_route`: per solution pick num_partial tours with more than one customer and mutilate them (random_nodes /
 * random_cut / const_cut), drop singleton tours, remove max(round(frac_complete * n), 1) of the remaining complete
 * tours (random or smallest).  Random numbers: the counter-based generator of csrc/common.cuh keyed by (seed, rollout,
 * iteration * 8192 + draw), restated in oracle/rng.py / oracle/destroy_oracle.py.  src_plan (n_src, K_max, L) tour buffers; rollout b mutilates solution src_index[b];
 * output in the dense format of jampr_env_import_tours with n_tours = K_max, tour_len = L (complete tours first). */
typedef struct {
  int32_t  num_partial;        /* tours to mutilate per solution */
  int32_t  rm_partial_mode;    /* 0 random_nodes, 1 random_cut, 2 const_cut, 3 one of the three at random per tour */
  int32_t  rm_complete_mode;   /* 0 random, 1 smallest */
  float    frac_partial;       /* frac_rm_nodes_from_partial, in [0, 1) */
  float    frac_complete;      /* frac_rm_complete_routes, in [0, 1) */
  int32_t  iteration;          /* LNS step: selects the random stream */
  uint64_t seed;
} jampr_destroy_t;
JAMPR_API int jampr_lns_destroy(const jampr_dims_t* d, const int16_t* src_plan, const int32_t* src_index,
                      const jampr_destroy_t* prm, int16_t* tours, uint8_t* kind, void* stream);

/* Local search on tour buffers: the stand-in for the reference's OR-tools guided local search (lib/lns/gort.py:129-290,
 * lib/utils/or_utils.py:10-171; OR-tools is neither installed nor pinned).  Best-improvement descent over relocate and
 * 2-opt* moves, every move evaluated with the environment's own feasibility arithmetic (env.py:981-990, 1081-1131); at most
 * max_iters moves per solution.  plans (n_solutions, K_max, L) is improved in place; inst_of (n_solutions) names the instance
 * of each solution (NULL = solution / n_samples); cost_out (n_solutions) = batch-independent cost of the served tours (travel +
 * waiting + service incl. return legs, normalised units; unserved customers are not touched and not charged here);
 * moves_out (n_solutions) = moves applied.  Deterministic; restated in oracle/ls_oracle.py. */
JAMPR_API int jampr_local_search(const jampr_dims_t* d, const jampr_instances_t* inst, int16_t* plans, int32_t n_solutions,
                       const int32_t* inst_of, int32_t max_iters, float* cost_out, int32_t* moves_out, void* stream);

/* Policy.forward for one step (policy.py:120-227; tour_encoder.py:159-243; attn_decoder.py:62-98):
 * tour-graph GNN (or cached embedding), tour embeddings, pointer attention, masked log-softmax,
 * greedy argmax or inverse-CDF sampling.  graph_fresh: 1 = rebuild the tour-graph embedding,
 * 0 = reuse st->h_fin / st->ne (the rule of env.py:801 evaluated by the caller).  uniforms: (bs) floats in [0,1) or NULL to draw them from
 * the counter-based generator keyed by (seed, rollout, step). */
JAMPR_API int jampr_policy_step(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                      const jampr_state_t* st, int32_t graph_fresh, int32_t step_no,
                      int32_t decode_type, uint64_t seed, const float* uniforms, void* stream);

/* The two launches of jampr_policy_step's split variant separately (part 1 = per-rollout GNN kernel, part 2 = decoder over
 * 128 rollouts per CTA; csrc/policy_tc2.cu, csrc/decoder_tc.cu), for callers that time them one by one (bench.py).
 * JAMPR_E_UNSUPPORTED when the configuration is not served by that variant.  Replaces nothing in the reference. */
JAMPR_API int jampr_policy_step_part(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                           const jampr_state_t* st, int32_t graph_fresh, int32_t step_no, int32_t decode_type,
                           uint64_t seed, const float* uniforms, int32_t part, void* stream);

/* The per-step tour-graph GNN alone (tour_encoder.py:159-198 + the node_emb projection :217-218): what
 * jampr_policy_step runs when graph_fresh = 1.  Exposed so that callers (tests, bench.py) can check and
 * time the dominant kernel in isolation. */
JAMPR_API int jampr_tour_encoder(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                       const jampr_state_t* st, int32_t step_no, void* stream);

/* The whole `while not done: action = policy(obs); obs = env.step(action)` loop
 * (training.py:72-77, lns.py:258-261) enqueued without host round trips: at most max_steps
 * iterations starting at env._step = first_step; kernels turn into no-ops once the device-side
 * done word is set.  Returns the number of iterations enqueued (>= 0) or an error (< 0). */
JAMPR_API int jampr_rollout(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                  const jampr_state_t* st, int32_t first_step, int32_t first_fresh,
                  int32_t decode_type, uint64_t seed, int32_t max_steps, void* stream);

/* Best-of-samples reduction (training.py:80-94): per instance argmin of total cost over its P
 * rollouts; copies that rollout's tour buffer.  best_cost (I), best_idx (I), best_plan (I,K_max,L). */
JAMPR_API int jampr_gather_best(const jampr_dims_t* d, const jampr_state_t* st, float* best_cost,
                      int32_t* best_idx, int16_t* best_plan, void* stream);

/* Batch-independent cost of the tours in the buffers: travel + waiting + service per tour incl. the
 * return leg, + 2*time_to_depot per unvisited customer (logic of env.py:708-718, 287-295). */
JAMPR_API int jampr_true_cost(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_state_t* st,
                    float* out, void* stream);

/* Diagnostic: one 128 x N x K tcgen05 tile product D = A * B^T built from the same pieces as the tensor-core
 * encoder (thread-written swizzled A tile, bulk-copied pre-swizzled B image, TMEM accumulator read-back).
 * A (128,K) row-major 16-bit, Bimg = packing.sw128_image(B (N,K)), D (128,N) fp32.  precision = JAMPR_PREC_BF16
 * or JAMPR_PREC_F16.  Replaces nothing in the reference; tests/test_gpu_tc.py pins the UMMA descriptors with it. */
JAMPR_API int jampr_selftest_umma(const void* A, const void* Bimg, float* D, int32_t K, int32_t N, int32_t precision,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JAMPR_B200_H */
