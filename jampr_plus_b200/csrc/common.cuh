// Shared device helpers for libjampr_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/jampr_b200.h"

#define JAMPR_H 128   // hidden dim of the GNN encoders (gnn_attn.yaml:11,21)
#define JAMPR_D 256   // embedding / decoder dim (gnn_attn.yaml:6,34)
#define JAMPR_HEADS 4
#define JAMPR_MAX_SLOTS 8

#define CUDA_RET(call)                                   \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return (int)e__;             \
  } while (0)

static inline int launch_status() { return (int)cudaGetLastError(); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Is the kernel launched for env._step == step_no still live?  Not after the episode terminated (env.py:281), and
// not while a sharded run is parked at its local finishing step waiting for the global one (JAMPR_CTL_HOLD).
__device__ __forceinline__ bool jampr_step_active_w(int done_step, int hold, int fin, int ld, int step_no) {
  if (done_step >= 0 && done_step < step_no) return false;
  if (hold != 0 && fin < 0 && ld >= 0 && ld < step_no) return false;
  return true;
}
__device__ __forceinline__ bool jampr_step_active(const int32_t* ctl, int step_no) {
  return jampr_step_active_w(ctl[JAMPR_CTL_DONE_STEP], ctl[JAMPR_CTL_HOLD], ctl[JAMPR_CTL_FINAL_STEP],
                             ctl[JAMPR_CTL_LOCAL_DONE], step_no);
}

// A rollout whose customers are all served and whose vehicles are all back at the depot can only make zero-cost
// depot -> depot moves until the batch terminates (env.py:281: lockstep over the batch): no state changes whichever
// slot the policy picks.  The fused device loop skips the policy for such rollouts and emits action (slot 0, depot);
// callers that want the log-probabilities (st.logp set: step API of the parity tests) still get the full forward.
__device__ __forceinline__ bool jampr_rollout_idle(const jampr_state_t& st, int b, int K) {
  if (st.logp != nullptr || !st.allvis[b]) return false;
  for (int t = 0; t < K; ++t)
    if (st.cur[(size_t)b * K + t] != 0) return false;
  return true;
}
// same test with the all-visited flag already in hand (loaded next to the control word, off the critical path)
__device__ __forceinline__ bool jampr_rollout_idle2(const jampr_state_t& st, int b, int K, uint8_t allvis) {
  if (!allvis) return false;
  for (int t = 0; t < K; ++t)
    if (st.cur[(size_t)b * K + t] != 0) return false;
  return true;
}

// Counter-based uniform in [0,1): splitmix64 of (seed, rollout, step). Restated in
// oracle/rng.py so that sampling runs can be replayed on the CPU.
__host__ __device__ __forceinline__ float jampr_uniform(uint64_t seed, uint32_t rollout, uint32_t step) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((uint64_t)rollout * 0x100000001B3ull + step + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// ---- weight blob layout (floats). Mirrors jampr_plus_b200/model/packing.py -------------------
// A "conv block" is [WT (256 x 128): rows 0..127 lin_rel^T, rows 128..255 lin_root^T][bias 128]
// [ln gamma 128][ln beta 128].
#define BLK_WT 0
#define BLK_B (256 * 128)
#define BLK_G (BLK_B + 128)
#define BLK_BETA (BLK_G + 128)
#define BLK_FLOATS (BLK_BETA + 128)

enum {
  OFF_NE_IN_WT = 0,                                 // [8][128] (7 rows used)
  OFF_NE_IN_B = OFF_NE_IN_WT + 8 * 128,             // [128]
  OFF_NE_BLK = OFF_NE_IN_B + 128,                   // 3 blocks
  OFF_NE_OUT_WT = OFF_NE_BLK + 3 * BLK_FLOATS,      // [2][128][128] column halves of output_proj^T
  OFF_NE_OUT_B = OFF_NE_OUT_WT + 2 * 128 * 128,     // [256]
  OFF_TE_IN_WT = OFF_NE_OUT_B + 256,                // [256][128] node_emb_input_proj^T
  OFF_TE_IN_B = OFF_TE_IN_WT + 256 * 128,           // [128]
  OFF_TE_BLK = OFF_TE_IN_B + 128,                   // 6 blocks: tour_layers 0..2, rev_tour_layers 0..2
  OFF_TE_OUT_WT = OFF_TE_BLK + 6 * BLK_FLOATS,      // [2][128][128] node_emb_output_proj^T halves
  OFF_TE_OUT_B = OFF_TE_OUT_WT + 2 * 128 * 128,     // [256]
  OFF_TF_WT = OFF_TE_OUT_B + 256,                   // [8][128] tour feature input_proj^T (5 rows used)
  OFF_TF_B = OFF_TF_WT + 8 * 128,                   // [128]
  OFF_TO_WT = OFF_TF_B + 128,                       // [256][256] tour output_proj^T
  OFF_TO_B = OFF_TO_WT + 256 * 256,                 // [256]
  OFF_CTXT_WT = OFF_TO_B + 256,                     // [512][256] ctxt_proj^T
  OFF_GK_W = OFF_CTXT_WT + 512 * 256,               // [256][256] set_proj rows 0..255 (as stored)
  OFF_GV_WT = OFF_GK_W + 256 * 256,                 // [256][256] set_proj rows 256..511, transposed
  OFF_LK_W = OFF_GV_WT + 256 * 256,                 // [256][256] set_proj rows 512..767 (as stored)
  OFF_OUT_WT = OFF_LK_W + 256 * 256,                // [256][256] out_proj^T
  OFF_TOF_WT = OFF_OUT_WT + 256 * 256,              // [8][256] (5 rows used) output_proj[:, :128] @ tour input_proj: the
                                                    // tour-feature branch of tour_encoder.py:202-213 folded (tensor-core path)
  OFF_TOF_B = OFF_TOF_WT + 8 * 256,                 // [256] output_proj[:, :128] @ input_proj.bias + output_proj.bias
  OFF_END = OFF_TOF_B + 256
};
