// Host-side orchestration of the policy step and of the whole construction rollout (C ABI).
// Nothing here synchronises: every call only enqueues kernels on the caller's stream, so a rollout can be
// captured into a CUDA graph by the caller.  Termination is data dependent (env.py:281) and lives in the
// device control block; kernels launched after the terminating step return immediately.
#include <stdlib.h>
#include "common.cuh"

int jampr_node_encoder_f32(const jampr_dims_t*, const jampr_instances_t*, const float*, cudaStream_t);
int jampr_tour_encoder_f32(const jampr_dims_t*, const jampr_instances_t*, const jampr_state_t*, const float*, int,
                           cudaStream_t);
int jampr_decoder_launch(const jampr_dims_t*, const jampr_instances_t*, const jampr_state_t*, const float*, int, int,
                         uint64_t, const float*, cudaStream_t);
int jampr_policy_step_tc(const jampr_dims_t*, const jampr_instances_t*, const jampr_weights_t*, const jampr_state_t*,
                         int, int, int, uint64_t, const float*, int, cudaStream_t);
int jampr_policy_step_tc2(const jampr_dims_t*, const jampr_instances_t*, const jampr_weights_t*, const jampr_state_t*,
                          int, int, int, uint64_t, const float*, int, int, cudaStream_t);
int jampr_decoder_tc(const jampr_dims_t*, const jampr_weights_t*, const jampr_state_t*, int, int, uint64_t, const float*,
                     cudaStream_t);
int jampr_decoder_tc_fits(const jampr_dims_t*);
int64_t jampr_w16_elems(void);
int jampr_tc_prepare(const jampr_dims_t*, const jampr_instances_t*, int, cudaStream_t);
int jampr_tc_finish(const jampr_dims_t*, const jampr_instances_t*, cudaStream_t);
int jampr_lg_node_encoder(const jampr_dims_t*, const jampr_instances_t*, const jampr_weights_t*, cudaStream_t);
int jampr_lg_tour_encoder(const jampr_dims_t*, const jampr_instances_t*, const jampr_weights_t*, const jampr_state_t*,
                          int, cudaStream_t);
int jampr_lg_handover(const jampr_dims_t*, const jampr_state_t*, int, cudaStream_t);
#define JAMPR_TILE_NODES 128     // graphs up to this size run the single-tile kernels, larger ones the layer-wise path

// Tensor-core policy step, variant per jampr_weights_t.tc_variant (include/jampr_b200.h): 3 = the two-CTAs-per-SM GNN
// kernel (policy_tc2.cu) in split mode followed by the 128-rollouts-per-CTA tcgen05 decoder (decoder_tc.cu), when the
// caller supplied the hand-over buffers and the shapes fit; 2 = the same GNN kernel with the decoder fused per rollout;
// 1 (and graphs that do not fit the two-CTA tiling) = the one-CTA tiling (policy_tc.cu); 0 = 3 for large batches, else 2.
static int policy_step_tc_any(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                              const jampr_state_t* st, int fresh, int step_no, int decode_type, uint64_t seed,
                              const float* uniforms, int store_cache, cudaStream_t s) {
  int variant = w->tc_variant;
  if (variant <= 0 || variant > 3)
    variant = ((int64_t)d->n_inst * d->n_samples >= JAMPR_SPLIT_MIN_ROLLOUTS) ? 3 : 2;
  if (variant >= 2) {
    const bool split = variant == 3 && st->ne16 && st->dq && st->dte && st->dqp && st->dgbar && st->dlp &&
                       jampr_decoder_tc_fits(d);
    const int rc = jampr_policy_step_tc2(d, inst, w, st, fresh, step_no, decode_type, seed, uniforms, store_cache,
                                         split ? 1 : 0, s);
    if (rc == 0 && split) return jampr_decoder_tc(d, w, st, step_no, decode_type, seed, uniforms, s);
    if (rc != JAMPR_E_TOOLARGE) return rc;
  }
  return jampr_policy_step_tc(d, inst, w, st, fresh, step_no, decode_type, seed, uniforms, store_cache, s);
}

static int check_policy_args(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w) {
  if (!d || !inst || !w || !w->blob || w->n_floats < (int64_t)OFF_END) return JAMPR_E_BADARG;
  if (d->n_inst <= 0 || d->n_samples <= 0 || d->n_nodes < 2 || d->n_slots <= 0 || d->n_slots > JAMPR_MAX_SLOTS ||
      d->k_nbh < 2 || d->k_nbh > d->n_nodes)
    return JAMPR_E_BADARG;
  if (w->precision != JAMPR_PREC_F32 && w->precision != JAMPR_PREC_BF16 && w->precision != JAMPR_PREC_F16)
    return JAMPR_E_BADARG;
  if (w->precision != JAMPR_PREC_F32 && (!w->blob_bf16 || w->n_bf16 < jampr_w16_elems())) return JAMPR_E_BADARG;
  return 0;
}

extern "C" int64_t jampr_weight_blob_floats(void) { return (int64_t)OFF_END; }
extern "C" int64_t jampr_weight_blob16_elems(void) { return jampr_w16_elems(); }

extern "C" int jampr_weight_offsets(int64_t* out, int32_t max_entries) {
  static const int64_t offs[] = {OFF_NE_IN_WT, OFF_NE_IN_B, OFF_NE_BLK, OFF_NE_OUT_WT, OFF_NE_OUT_B, OFF_TE_IN_WT,
                                 OFF_TE_IN_B, OFF_TE_BLK, OFF_TE_OUT_WT, OFF_TE_OUT_B, OFF_TF_WT, OFF_TF_B,
                                 OFF_TO_WT, OFF_TO_B, OFF_CTXT_WT, OFF_GK_W, OFF_GV_WT, OFF_LK_W, OFF_OUT_WT,
                                 OFF_TOF_WT, OFF_TOF_B, OFF_END, BLK_FLOATS};
  const int n = (int)(sizeof(offs) / sizeof(offs[0]));
  if (!out || max_entries < n) return JAMPR_E_BADARG;
  for (int i = 0; i < n; ++i) out[i] = offs[i];
  return n;
}

extern "C" int jampr_node_encoder(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                                  void* stream) {
  const int rc = check_policy_args(d, inst, w);
  if (rc) return rc;
  // once per episode and < 1 % of it: every precision runs the fp32 kernel, the tensor-core step kernel
  // converts h0 / static_emb on the fly
  // Tensor-core precisions: the node encoder runs as tcgen05 GraphConv tiles too (gnn_large.cu, any graph size;
  // JAMPR_NODE_ENCODER=fp32 keeps the CUDA-core kernel for single-tile graphs).  Once per episode, < 1 % of it.
  static const bool ne_fp32 = [] {
    const char* e = getenv("JAMPR_NODE_ENCODER");
    return e && e[0] == 'f';
  }();
  if (w->precision != JAMPR_PREC_F32 && d->n_nodes <= JAMPR_TILE_NODES) {
    const int rc2 = jampr_tc_prepare(d, inst, w->precision, (cudaStream_t)stream);
    if (rc2) return rc2;
  }
  int rc3;
  if (w->precision != JAMPR_PREC_F32 && (d->n_nodes > JAMPR_TILE_NODES || !ne_fp32))
    rc3 = jampr_lg_node_encoder(d, inst, w, (cudaStream_t)stream);
  else
    rc3 = jampr_node_encoder_f32(d, inst, w->blob, (cudaStream_t)stream);
  // single-tile tensor-core step: column-chunk-major copies of h0 / static_emb (include/jampr_b200.h)
  if (rc3 == 0 && w->precision != JAMPR_PREC_F32 && d->n_nodes <= JAMPR_TILE_NODES) rc3 = jampr_tc_finish(d, inst, (cudaStream_t)stream);
  return rc3;
}

extern "C" int jampr_tour_encoder(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                                  const jampr_state_t* st, int32_t step_no, void* stream) {
  const int rc = check_policy_args(d, inst, w);
  if (rc) return rc;
  if (!st || !st->h_fin || !st->ne || !st->ne_stats) return JAMPR_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  // on the tensor-core path the encoder is fused with the decoder: the call runs the whole policy step (greedy)
  // and additionally stores h_fin / ne / ne_stats
  if (w->precision != JAMPR_PREC_F32 && d->n_nodes > JAMPR_TILE_NODES)
    return jampr_lg_tour_encoder(d, inst, w, st, step_no, s);
  if (w->precision != JAMPR_PREC_F32)
    return policy_step_tc_any(d, inst, w, st, 1, step_no, JAMPR_DECODE_GREEDY, 0, nullptr, 1, s);
  return jampr_tour_encoder_f32(d, inst, st, w->blob, step_no, s);
}

extern "C" int jampr_policy_step(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                                 const jampr_state_t* st, int32_t graph_fresh, int32_t step_no, int32_t decode_type,
                                 uint64_t seed, const float* uniforms, void* stream) {
  int rc = check_policy_args(d, inst, w);
  if (rc) return rc;
  if (!st || (decode_type != JAMPR_DECODE_GREEDY && decode_type != JAMPR_DECODE_SAMPLING)) return JAMPR_E_BADARG;
  if (!st->h_fin || !st->ne || !st->ne_stats) return JAMPR_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (w->precision != JAMPR_PREC_F32 && d->n_nodes > JAMPR_TILE_NODES) {
    // large-graph path: layer-wise tensor-core GNN through HBM/L2; then the tcgen05 decoder over tiles of rollouts when
    // the caller supplied its hand-over buffers and the action set fits (decoder_tc.cu), else the fp32 decoder kernel
    if (graph_fresh) {
      rc = jampr_lg_tour_encoder(d, inst, w, st, step_no, s);
      if (rc) return rc;
    }
    if (st->ne16 && st->dq && st->dte && st->dqp && st->dgbar && st->dlp && jampr_decoder_tc_fits(d)) {
      rc = jampr_lg_handover(d, st, step_no, s);
      if (rc) return rc;
      return jampr_decoder_tc(d, w, st, step_no, decode_type, seed, uniforms, s);
    }
    return jampr_decoder_launch(d, inst, st, w->blob, step_no, decode_type, seed, uniforms, s);
  }
  if (w->precision != JAMPR_PREC_F32)
    return policy_step_tc_any(d, inst, w, st, graph_fresh, step_no, decode_type, seed, uniforms,
                              (d->upd_step != 1) || st->logp != nullptr, s);
  if (graph_fresh) {
    rc = jampr_tour_encoder_f32(d, inst, st, w->blob, step_no, s);
    if (rc) return rc;
  }
  return jampr_decoder_launch(d, inst, st, w->blob, step_no, decode_type, seed, uniforms, s);
}

// The two launches of the split policy step one by one (part 1 = per-rollout GNN kernel in split mode, part 2 = the
// 128-rollout decoder), so that callers can time them separately (bench.py roofline entries).  Only for configurations the
// split variant serves; jampr_policy_step is the product entry point.
extern "C" int jampr_policy_step_part(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                                      const jampr_state_t* st, int32_t graph_fresh, int32_t step_no, int32_t decode_type,
                                      uint64_t seed, const float* uniforms, int32_t part, void* stream) {
  const int rc = check_policy_args(d, inst, w);
  if (rc) return rc;
  if (!st || w->precision == JAMPR_PREC_F32 || d->n_nodes > JAMPR_TILE_NODES || (part != 1 && part != 2)) return JAMPR_E_BADARG;
  if (!(st->ne16 && st->dq && st->dte && st->dqp && st->dgbar && st->dlp) || !jampr_decoder_tc_fits(d)) return JAMPR_E_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (part == 1)
    return jampr_policy_step_tc2(d, inst, w, st, graph_fresh, step_no, decode_type, seed, uniforms,
                                 (d->upd_step != 1) || st->logp != nullptr, 1, s);
  return jampr_decoder_tc(d, w, st, step_no, decode_type, seed, uniforms, s);
}

extern "C" int jampr_rollout(const jampr_dims_t* d, const jampr_instances_t* inst, const jampr_weights_t* w,
                             const jampr_state_t* st, int32_t first_step, int32_t first_fresh, int32_t decode_type,
                             uint64_t seed, int32_t max_steps, void* stream) {
  int rc = check_policy_args(d, inst, w);
  if (rc) return rc;
  if (!st || first_step < 0 || d->upd_step <= 0) return JAMPR_E_BADARG;
  const int cap = d->n_nodes + d->k_max + 1;   // last admissible value of env._step (env.py:281)
  int fresh = first_fresh, n = 0;
  for (int step = first_step; step <= cap && n < max_steps; ++step, ++n) {
    rc = jampr_policy_step(d, inst, w, st, fresh, step, decode_type, seed, nullptr, stream);
    if (rc) return rc > 0 ? -1000 - rc : rc;
    rc = jampr_env_step(d, inst, st, nullptr, step, stream);
    if (rc) return rc > 0 ? -1000 - rc : rc;
    fresh = (step <= d->n_slots + 1) || (step % d->upd_step == 0);   // env.py:801 with the pre-increment _step
  }
  return n;
}
