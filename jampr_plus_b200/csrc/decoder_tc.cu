// Split decoder of the tensor-core policy step (sm_100a): tour_encoder.py:202-213, policy.py:179-182, attn_decoder.py:62-98
// and policy.py:203-227 for up to 128 rollouts per CTA.  policy_tc2.cu (split mode) leaves per rollout the 16-bit
// node-embedding rows (st.ne16), their column mean / max (st.dq), the per-tour means of h and the tour features (st.dqp)
// in HBM; this kernel runs
//
//   stage T   tour_emb[t] = W_to [input_proj(f_t) | mean_h[t]] + b as one product over K = 192 per vehicle slot (mean half,
//             then [features, 1]: feature matrix and bias folded into the third B tile); a second set of accumulator
//             columns collects sum_t, the running max_t stays in registers                       (tour_encoder.py:202-213)
//   stage A   q = [mean_n + mean_t, max_n + max_t] formed while the tiles are filled (policy.py:179-182);
//             ctxt = q W_c^T                       GEMM 128 x 256 x 512          (attn_decoder.py:73)
//   stage B   qp_h = W_gk,h^T ctxt_h  (4 heads)    4 GEMMs 128 x 256 x 64        (set_proj moved to the query side)
//   phase C   per rollout, one warp: glimpse scores qp_h . (node_emb[n_a] + tour_emb[t_a]) over the FEASIBLE actions,
//             masked softmax per head, gbar_h = sum_a p_h[a] act[a] — on warp-level tensor-core MMAs (see below)
//   stage D   g[64h..] = W_gv,h gbar_h  (4 heads)  4 GEMMs 128 x 64 x 256        (attn_decoder.py:86-88)
//   stage E   lp = (W_lk^T W_out) g                GEMM 128 x 256 x 256          (out_proj + logit keys, pre-multiplied)
//   phase F   per rollout, one warp: logits 10 tanh(lp . act[a] / 16), masked log-softmax, greedy argmax or
//             inverse-CDF sampling (policy.py:203-227)
//
// The dense stages are tcgen05.mma (M = 128 rollouts, fp32 accumulators in TMEM).  Both operands are SPLIT into a hi and
// a lo 16-bit part (x = hi + lo) and every K step issues hi*hi + lo*hi + hi*lo: ~21 significand bits, i.e. the stages
// are as accurate as fp32 mat-vecs.  That matters: the glimpse scores are O(10) before the softmax, and rounding the
// glimpse-key weights alone to fp16 costs 9e-3 in the log-probabilities (tools/emulate_tc_precision.py).  The tensor
// pipe has the room: 1.0 MFLOP per rollout x 3 is 6 % of the GNN's MMAs.  Weights are read once per tile
// (the fused per-rollout decoder streamed 704 KB of them per rollout).
// A operand tiles are written by the threads (from HBM vectors or straight from the previous accumulator in TMEM),
// B tiles are pre-swizzled images streamed by the bulk-copy engine; both are double-buffered and two items are in flight.
// A tile holds ceil(bs / (waves x SMs)) <= 128 rollouts, so that the one-CTA-per-SM grid runs whole waves.
//
// The per-rollout phases C and F are the HBM-facing part: each gathers the 16-bit node-embedding rows of the feasible
// actions (512 B per row, ~2.7 GB per launch at 65,536 rollouts).  A warp stages 16 rows at a time in shared memory
// with 16-byte asynchronous copies (no register staging: 8 KB in flight per warp, 128 KB per SM) and runs the dot
// products as mma.sync.m16n8k16 on fragments read with ldmatrix:
//   scores   S[h][a]  = sum_c Qp[h][c] R[a][c]     A = Qp as [hi(4 heads); lo(4 heads)] (fp32 = hi + lo), B = R
//   glimpse  G[c][h] += sum_a R[a][c] P[a][h]      A = R^T (ldmatrix.trans), B = [hi(p); lo(p)] straight from the score
//                                                  accumulators (flash-attention register reuse), online softmax
// so the rows are fetched once for scores and weighted sums, and a rollout costs ~5.6 k instructions instead of
// ~12 k on the CUDA cores.  Node-embedding rows are exact 16-bit values; queries and probabilities enter as hi + lo,
// products are exact and accumulate in fp32: the accuracy of the fp32 dot products it replaces.
#include <stdlib.h>
#include "policy_tc_common.cuh"

#define DT_THREADS 512
#define DT_WARPS (DT_THREADS / 32)
#define DT_M 128
#define DT_RROWS 16            // node-embedding rows staged per chunk (two mma n-tiles)
#define DT_RSTRIDE 528         // bytes per staged row: 512 + 16 (8 consecutive rows hit 32 distinct banks in ldmatrix)
#define DT_GSTRIDE 260         // floats per row of the [8][256] hand-over of the glimpse accumulators (conflict-free)
#define DT_ITEMS 32            // B tiles consumed after stage T (3 per vehicle slot): 8 (stage A) + 4 (B) + 16 (D) + 4 (E)

// per-phase time line of the CTAs (tools/timeline.py decoder): build with `make EXTRA=-DJAMPR_TIMELINE`
#ifdef JAMPR_TIMELINE
#define DTL_N 16
__device__ uint32_t g_dtimeline[1024][DTL_N];
#define DTL(k) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_dtimeline[blockIdx.x][k] = (uint32_t)clock64(); } while (0)
#else
#define DTL(k) do { } while (0)
#endif

struct DtArgs {
  jampr_dims_t d;
  jampr_state_t st;
  const float* w;        // fp32 blob: folded tour-feature branch (OFF_TOF_*)
  const void* w16;
  const float* uniforms;
  uint64_t seed;
  int step_no;
  int decode_type;
  int tile_rollouts;     // rollouts per CTA (<= DT_M), chosen by the launcher so that the tiles fill whole waves of the SMs
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// 16 fp32 values of row r, 16-column group g (0..3) of a K-block -> hi / lo operand tiles (two 16-byte chunks each)
template <typename T>
__device__ __forceinline__ void write_a16(uint8_t* sAhi, uint8_t* sAlo, int r, int g, const float (&v)[16]) {
  using T2 = typename Elem<T>::T2;
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const T2 hh = Elem<T>::from2(v[2 * i], v[2 * i + 1]);
    const float2 hf = Elem<T>::to2(hh);
    h[i] = as_u32(hh);
    l[i] = as_u32(Elem<T>::from2(v[2 * i] - hf.x, v[2 * i + 1] - hf.y));
  }
  const uint32_t o0 = sw128_off((uint32_t)r, (uint32_t)(2 * g)), o1 = sw128_off((uint32_t)r, (uint32_t)(2 * g + 1));
  *reinterpret_cast<uint4*>(sAhi + o0) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(sAhi + o1) = make_uint4(h[4], h[5], h[6], h[7]);
  *reinterpret_cast<uint4*>(sAlo + o0) = make_uint4(l[0], l[1], l[2], l[3]);
  *reinterpret_cast<uint4*>(sAlo + o1) = make_uint4(l[4], l[5], l[6], l[7]);
}

// B tile of item i inside one part of the decoder blob: element offset and element count
__device__ __forceinline__ void dt_item_src(int i, int n_t, uint32_t& off, uint32_t& elems) {
  if (i < n_t) { off = WD_TO + (uint32_t)(i % 3) * WD_TILE; elems = WD_TILE; return; }      // stage T: (slot, K-block) items
  i -= n_t;
  if (i < 8) { off = WD_C + (uint32_t)i * WD_TILE; elems = WD_TILE; }
  else if (i < 12) { off = WD_GK + (uint32_t)(i - 8) * WD_TILE; elems = WD_TILE; }
  else if (i < 28) { off = WD_GV + (uint32_t)(i - 12) * (WD_TILE / 4); elems = WD_TILE / 4; }
  else { off = WD_M1 + (uint32_t)(i - 28) * WD_TILE; elems = WD_TILE; }
}

__host__ __device__ inline uint32_t dt_a32(int A) { return (uint32_t)(A + 31) & ~31u; }
// per-warp scratch: staged rows [16][528 B] | query rows [8][528 B] (hi of 4 vectors, lo of 4 vectors) | logits [A32] f32 |
// feasible list [A32] i16 | nodes [A32] i16 | qt [16] f32.  The warps' regions overlay the GEMM operand tiles (dead
// during the per-rollout phases).
#define DT_R_BYTES (DT_RROWS * DT_RSTRIDE)
#define DT_Q_BYTES (8 * DT_RSTRIDE)
__host__ __device__ inline uint32_t dt_warp_bytes(int A) { return DT_R_BYTES + DT_Q_BYTES + dt_a32(A) * (4 + 2 + 2) + 64; }
#define DT_GEMM_BYTES (65536u + 131072u)
__host__ __device__ inline uint32_t dt_region_bytes(int A) {
  const uint32_t w = DT_WARPS * dt_warp_bytes(A);
  return w > DT_GEMM_BYTES ? w : DT_GEMM_BYTES;
}

template <typename T>
__global__ void __launch_bounds__(DT_THREADS, 1) decoder_tc_kernel(const DtArgs a) {
  using T2 = typename Elem<T>::T2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const jampr_dims_t& d = a.d;
  const jampr_state_t& st = a.st;
  if (!jampr_step_active(st.ctl, a.step_no)) return;
  const int N = d.n_nodes, K = d.n_slots, k = d.k_nbh, A = K * k;
  const int b0 = blockIdx.x * a.tile_rollouts;
  // one past the last rollout of this tile: accumulator rows beyond it are padding (computed, never stored or claimed)
  const int bs = min(d.n_inst * d.n_samples, b0 + a.tile_rollouts);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t A32 = dt_a32(A);

  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // pointer arithmetic keeps the shared space
  uint8_t* sA = base;                      // 2 buffers x ([128 rows][64 K] hi 16 KB | lo 16 KB)
  uint8_t* sB = base + 65536;              // 2 buffers x (hi 32 KB | lo 32 KB)
  uint8_t* scratch = base + dt_region_bytes(A);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch);          // full[0], full[1], mma[0], mma[1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(scratch + 32);
  int* s_any = reinterpret_cast<int*>(scratch + 48);
  int* s_next = reinterpret_cast<int*>(scratch + 52);            // [2]: next unclaimed rollout of the tile in phases C, F
  // per-warp region of the per-rollout phases (over the operand tiles)
  uint8_t* wbase = base + (size_t)warp * dt_warp_bytes(A);
  uint8_t* sR = wbase;                                                        // staged rows | glimpse hand-over
  uint8_t* sQ = wbase + DT_R_BYTES;                                           // query rows: hi 0..3, lo 4..7
  float* s_logit = reinterpret_cast<float*>(wbase + DT_R_BYTES + DT_Q_BYTES); // [A32]
  int16_t* s_list = reinterpret_cast<int16_t*>(wbase + DT_R_BYTES + DT_Q_BYTES + A32 * 4);      // feasible action indices
  int16_t* s_nb = reinterpret_cast<int16_t*>(wbase + DT_R_BYTES + DT_Q_BYTES + A32 * 6);        // node of action a
  float* s_qt = reinterpret_cast<float*>(wbase + DT_R_BYTES + DT_Q_BYTES + A32 * 8);            // [4 heads][4 slots]
  uint64_t* full = bars;
  uint64_t* bar_mma = bars + 2;            // [2]

  // ---- does any rollout of this tile need a decision?  (late in the episode whole tiles are idle)
  if (tid == 0) {
    *s_any = 0;
    s_next[0] = 0;
    s_next[1] = 0;
  }
  __syncthreads();
  {
    bool need = false;
    if (tid < DT_M && b0 + tid < bs) need = !jampr_rollout_idle(st, b0 + tid, K);
    if (need) *s_any = 1;
  }
  __syncthreads();
  if (*s_any == 0) {
    if (tid < DT_M && b0 + tid < bs) {
      const size_t b = (size_t)(b0 + tid);
      st.action[2 * b] = 0;
      st.action[2 * b + 1] = 0;
      st.ll[b] = 0.0f;
      st.entropy[b] = 0.0f;
    }
    return;
  }

  DTL(0);
  // The GNN kernel left this tile's per-tour means and features (st.dqp) in HBM: request them into L2 now, so that the
  // A-tile loads of stage T find them there
  for (int e = tid; e < DT_M * 4 * 5; e += DT_THREADS) {        // (rollout, slot, 128-byte line 0..4 of the 532 bytes read)
    const int r = e / 20, rem = e - r * 20, t = rem / 5, ln = rem - t * 5;
    if (b0 + r < bs && t < K)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(st.dqp + ((size_t)(b0 + r) * 4 + t) * 256 + ln * 32));
  }
  const int n_t = 3 * K, n_items = n_t + DT_ITEMS;      // stage T: three K-block items per vehicle slot
  const uint8_t* wdec = reinterpret_cast<const uint8_t*>(a.w16) + (size_t)W16_DEC * 2;
  auto issue_b = [&](int i) {          // one thread; the buffer's previous reader (item i - 2) must be complete
    uint32_t off, elems;
    dt_item_src(i, n_t, off, elems);
    const int buf = i & 1;
    mbar_arrive_expect_tx(&full[buf], elems * 4u);
    bulk_g2s(sB + buf * 65536, wdec + (size_t)off * 2, elems * 2u, &full[buf]);
    bulk_g2s(sB + buf * 65536 + 32768, wdec + ((size_t)WD_PART + off) * 2, elems * 2u, &full[buf]);
  };
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&bar_mma[0], 1);
    mbar_init(&bar_mma[1], 1);
    mbar_fence_init();
    issue_b(0);
    issue_b(1);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int q4 = warp & 3, cg = warp >> 2;            // TMEM lane quarter, column group of this warp
  const int r_t = q4 * 32 + lane;                     // accumulator row (= rollout of the tile) owned in TMEM accesses
  const uint32_t ta = tmem + ((uint32_t)(q4 * 32) << 16);
  const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
  const uint32_t idesc256 = umma_idesc(128, 256, Elem<T>::fmt), idesc64 = umma_idesc(128, 64, Elem<T>::fmt);
  DTL(9);

  // ---- the GEMM chain.  Items are numbered in program order; item i uses operand buffers i & 1 (A tile written by the
  // threads, B tile streamed by the copy engine) and completion barrier bar_mma[i & 1].  Two items are in flight: while
  // the tensor pipe works on item i the threads fill the A tile of item i + 1, so an item costs max(MMAs, tile fill)
  // instead of their sum.  Rules: the A tile of item i is written after every thread has seen item i - 2 complete; the
  // B tile of item i is requested after thread 0 has seen item i - 2 complete.
  // A tile from a row-major fp32 matrix in HBM: 2048 float4 per tile, thread owns float4 (j * 512 + tid): a warp reads two
  // whole 256-byte row segments per instruction
  auto load_global = [&](float4 (&v)[4], const float* src, size_t row_stride, int col0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = j * 32 + (tid >> 4);
      v[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (b0 + r < bs) v[j] = *reinterpret_cast<const float4*>(src + (size_t)(b0 + r) * row_stride + col0 + (tid & 15) * 4);
    }
  };
  auto store_tile = [&](int i, const float4 (&v)[4]) {
    uint8_t* hi = sA + (size_t)(i & 1) * 32768;
    uint8_t* lo = hi + 16384;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t r = (uint32_t)(j * 32 + (tid >> 4)), c4 = (uint32_t)(tid & 15);      // columns 4 c4 .. 4 c4 + 3
      const T2 h0 = Elem<T>::from2(v[j].x, v[j].y), h1 = Elem<T>::from2(v[j].z, v[j].w);
      const float2 f0 = Elem<T>::to2(h0), f1 = Elem<T>::to2(h1);
      const T2 l0 = Elem<T>::from2(v[j].x - f0.x, v[j].y - f0.y), l1 = Elem<T>::from2(v[j].z - f1.x, v[j].w - f1.y);
      const uint32_t o = sw128_off(r, c4 >> 1) + (c4 & 1u) * 8u;
      *reinterpret_cast<uint2*>(hi + o) = make_uint2(as_u32(h0), as_u32(h1));
      *reinterpret_cast<uint2*>(lo + o) = make_uint2(as_u32(l0), as_u32(l1));
    }
  };
  // A tile from 64 accumulator columns in TMEM: warp = (lane quarter, 16-column group)
  auto fill_from_tmem = [&](int i, uint32_t col0) {
    uint32_t raw[16];
    tmem_ld16_nowait(ta + col0 + (uint32_t)cg * 16u, raw);
    tmem_wait_ld();
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
    uint8_t* hi = sA + (size_t)(i & 1) * 32768;
    write_a16<T>(hi, hi + 16384, r_t, cg, v);
  };
  // block barrier, then thread 0 issues the MMAs of item i: nks K steps x (hi*hi, lo*hi, hi*lo) into accumulator columns
  // d_col (and, stage T, a second time into d_col2 >= 0), commits, and requests the B tile of item i + 1
  auto issue_item = [&](int i, uint32_t d_col, uint32_t idesc, bool first, int nks = 4, int d_col2 = -1, bool first2 = false,
                        bool prefetch = true) {
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(&full[i & 1], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t ah = a_addr + (uint32_t)(i & 1) * 32768u, al = ah + 16384u;
      const uint32_t bh = b_addr + (uint32_t)(i & 1) * 65536u, bl = bh + 32768u;
      for (int ks = 0; ks < nks; ++ks) {
        const uint64_t dah = umma_desc_sw128(ah + ks * 32), dal = umma_desc_sw128(al + ks * 32);
        const uint64_t dbh = umma_desc_sw128(bh + ks * 32), dbl = umma_desc_sw128(bl + ks * 32);
        umma_f16(tmem + d_col, dah, dbh, idesc, (first && ks == 0) ? 0u : 1u);
        umma_f16(tmem + d_col, dal, dbh, idesc, 1u);
        umma_f16(tmem + d_col, dah, dbl, idesc, 1u);
        if (d_col2 >= 0) {
          umma_f16(tmem + (uint32_t)d_col2, dah, dbh, idesc, (first2 && ks == 0) ? 0u : 1u);
          umma_f16(tmem + (uint32_t)d_col2, dal, dbh, idesc, 1u);
          umma_f16(tmem + (uint32_t)d_col2, dah, dbl, idesc, 1u);
        }
      }
      umma_commit(&bar_mma[i & 1]);
      if (i >= 1 && prefetch && i + 1 < n_items) {
        mbar_wait(&bar_mma[(i - 1) & 1], (uint32_t)(((i - 1) >> 1) & 1));      // item i - 1 read the buffer of item i + 1
        issue_b(i + 1);
      }
    }
  };
  auto wait_item = [&](int i) {
    mbar_wait(&bar_mma[i & 1], (uint32_t)((i >> 1) & 1));
    tc_fence_after();
  };
  // ---- accumulator rows -> HBM with whole-row-segment accesses.  tcgen05.ld hands a thread 32 columns of ITS row; stored
  // directly that is 32 different 128-byte lines per warp instruction (the launch was bound by exactly that).  Each warp
  // turns its 32 x 32 block through 4 KB of shared memory (16-byte chunks XOR-ed with the row: conflict-free both ways)
  // and calls f(j, row of the tile, column, float4), j = 0..7, with 8 lanes covering 128 contiguous bytes of one row
  // (row = 32 q4 + 4 j + lane / 8, column = col0 + 4 (lane % 8)).  The turn-table
  // lives in the B buffer of the item that just completed (`buf`), which the copy engine refills only later.
  auto coop_block = [&](int buf, const uint32_t (&vals)[32], int col0, auto&& f) {
    uint8_t* tt = sB + (size_t)buf * 65536 + (size_t)warp * 4096;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<uint4*>(tt + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(vals[4 * c], vals[4 * c + 1], vals[4 * c + 2], vals[4 * c + 3]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int rw = 4 * j + (lane >> 3), c = lane & 7;
      const uint4 x = *reinterpret_cast<const uint4*>(tt + rw * 128 + ((c ^ (rw & 7)) << 4));
      f(j, q4 * 32 + rw, col0 + 4 * c, make_float4(__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w)));
    }
    __syncwarp();
  };
  // accumulator columns [col0, col0 + 256) -> row-major fp32 rows in HBM (dst row stride in floats)
  auto drain_to_global = [&](int buf, uint32_t col0, float* dst, size_t row_stride) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t raw[32];
      tmem_ld32_nowait(ta + col0 + (uint32_t)(cg * 64 + c * 32), raw);
      tmem_wait_ld();
      coop_block(buf, raw, cg * 64 + c * 32, [&](int, int row, int col, float4 x) {
        if (b0 + row < bs) *reinterpret_cast<float4*>(dst + (size_t)(b0 + row) * row_stride + col) = x;
      });
    }
  };

  // ================================================================ stage T: tour embeddings and the decoder query
  // tour_encoder.py:202-213: tour_emb[t] = W_to [input_proj(features_t) | mean_h[t]] + b as one product over K = 192 per
  // vehicle slot: A = [per-tour mean of h (128) | features (5), 1, 0 ..] from st.dqp[:, t, 0..132] left by the GNN kernel
  // (the folded feature matrix and bias are K rows 0..5 of the third B tile).  Every item also accumulates into columns 256..511, which
  // end up holding sum_t tour_emb[t]; the running maximum over the slots stays in registers (thread = row, 64 columns).
  // Sum and maximum go to stage A through two scratch rows per rollout.
  {
    float mx[64];
    float4 v[4];
    load_global(v, st.dqp, 1024, 0);
#pragma unroll 1
    for (int e = 0; e < n_t; ++e) {
      const int t = e / 3, kb = e - 3 * t;
      if (e >= 2) wait_item(e - 2);
      store_tile(e, v);
      if (e == 0) DTL(10);
      issue_item(e, 0u, idesc256, kb == 0, kb == 2 ? 1 : 4, 256, e == 0);
      if (e == 0) DTL(11);
      if (e + 1 < n_t) {
        const int tn = (e + 1) / 3, kn = (e + 1) - 3 * tn;
        load_global(v, st.dqp + (size_t)tn * 256, 1024, kn * 64);
        if (kn == 2) {           // feature block [f0 .. f4, 1 (bias row), 0 ..]: the GNN kernel left the five features
          const int c4 = tid & 15;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c4 == 1) v[j] = make_float4(v[j].x, 1.0f, 0.0f, 0.0f);
            else if (c4 > 1) v[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          }
        }
      }
      if (kb == 2) {
        wait_item(e);
        if (e == 2) DTL(12);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t raw[32];
          tmem_ld32_nowait(ta + (uint32_t)(cg * 64 + c * 32), raw);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) mx[c * 32 + j] = (t == 0) ? __uint_as_float(raw[j]) : fmaxf(mx[c * 32 + j], __uint_as_float(raw[j]));
          coop_block(e & 1, raw, cg * 64 + c * 32, [&](int, int row, int col, float4 x) {
            if (b0 + row < bs) *reinterpret_cast<float4*>(st.dte + ((size_t)(b0 + row) * 4 + t) * 256 + col) = x;
          });
        }
        tc_fence_before();           // the next slot's first MMA overwrites these columns after the barrier in issue_item
        if (e == 2) DTL(13);
      }
    }
    DTL(14);
    // hand-over of sum_t tour_emb[t] (accumulator columns 256..511) and max_t tour_emb[t] (registers) through the rows
    // st.dgbar[:, 0] / [:, 1] (free until phase C); stage A adds the node statistics when it builds its operand tiles
    const int lbuf = (n_t - 1) & 1;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t raw[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(mx[c * 32 + j]);
      coop_block(lbuf, raw, cg * 64 + c * 32, [&](int, int row, int col, float4 x) {
        if (b0 + row < bs) *reinterpret_cast<float4*>(st.dgbar + (size_t)(b0 + row) * 1024 + 256 + col) = x;
      });
      tmem_ld32_nowait(ta + 256u + (uint32_t)(cg * 64 + c * 32), raw);
      tmem_wait_ld();
      coop_block(lbuf, raw, cg * 64 + c * 32, [&](int, int row, int col, float4 x) {
        if (b0 + row < bs) *reinterpret_cast<float4*>(st.dgbar + (size_t)(b0 + row) * 1024 + col) = x;
      });
    }
    tc_fence_before();
  }
  DTL(15);
  __threadfence_block();
  __syncthreads();               // the scratch rows and st.dte of this tile are visible to every warp of the CTA
  DTL(1);
  DTL(2);

  // ================================================================ stage A: ctxt (TMEM columns 0..255)
  // policy.py:179-182: q = [mean_n(node_emb) + mean_t(tour_emb), max_n(node_emb) + max_t(tour_emb)], built while the
  // operand tiles are filled: node statistics from st.dq (GNN kernel), tour part from stage T; same operation order as
  // the fused kernel (nmean + sum / K, nmax + max).  A data is requested two items ahead.
  {
    // (the sums are formed when the tile is written, not when the loads are issued: the loads of two items stay in flight)
    auto load_q = [&](float4 (&n)[4], float4 (&t)[4], int kb) {
      const int part = kb >> 2, col = (kb & 3) * 64 + (tid & 15) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = j * 32 + (tid >> 4);
        n[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        t[j] = n[j];
        if (b0 + r < bs) {
          n[j] = *reinterpret_cast<const float4*>(st.dq + (size_t)(b0 + r) * 512 + part * 256 + col);
          t[j] = *reinterpret_cast<const float4*>(st.dgbar + (size_t)(b0 + r) * 1024 + part * 256 + col);
        }
      }
    };
    auto store_q = [&](int i, int kb, const float4 (&n)[4], const float4 (&t)[4]) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        v[j] = kb < 4 ? make_float4(n[j].x + t[j].x / (float)K, n[j].y + t[j].y / (float)K, n[j].z + t[j].z / (float)K, n[j].w + t[j].w / (float)K)
                      : make_float4(n[j].x + t[j].x, n[j].y + t[j].y, n[j].z + t[j].z, n[j].w + t[j].w);
      store_tile(i, v);
    };
    float4 na[4], ta4[4], nb[4], tb4[4];
    load_q(na, ta4, 0);
    load_q(nb, tb4, 1);
#pragma unroll 1
    for (int kb = 0; kb < 8; kb += 2) {
      const int i = n_t + kb;
      wait_item(i - 2);
      store_q(i, kb, na, ta4);
      issue_item(i, 0u, idesc256, kb == 0);
      if (kb + 2 < 8) load_q(na, ta4, kb + 2);
      wait_item(i - 1);
      store_q(i + 1, kb + 1, nb, tb4);
      issue_item(i + 1, 0u, idesc256, false);
      if (kb + 3 < 8) load_q(nb, tb4, kb + 3);
    }
  }
  DTL(3);
  // ================================================================ stage B: qp_h (TMEM columns 256..511) -> st.dqp
  {
    const int i0 = n_t + 8;
    wait_item(i0 - 1);           // ctxt complete (and with it item i0 - 2)
    fill_from_tmem(i0, 0u);
    issue_item(i0, 256u, idesc256, true);
#pragma unroll 1
    for (int h = 0; h < 4; ++h) {
      const int i = i0 + h;
      if (h < 3) fill_from_tmem(i + 1, (uint32_t)(h + 1) * 64u);      // A buffer of item i - 1: complete since the last drain
      wait_item(i);
      drain_to_global(i & 1, 256u, st.dqp + (size_t)h * 256, 1024);
      tc_fence_before();           // the next item's MMAs overwrite these columns after the block barrier in issue_item
      if (h < 3) issue_item(i + 1, 256u, idesc256, true, 4, -1, false, h < 2);   // no B tile in flight during phase C
    }
  }
  __threadfence_block();
  __syncthreads();               // st.dqp rows of this tile are visible to every warp of the CTA (same SM, L1 coherent
                                 // for data written by the CTA itself; reads below use plain loads)

  DTL(4);
  // ================================================================ phase C: glimpse per rollout (one warp each)
  const T* ne16 = reinterpret_cast<const T*>(st.ne16);
  const int g8 = lane >> 2, t4 = lane & 3;            // mma fragment coordinates
  const uint32_t r_saddr = smem_u32(sR), q_saddr = smem_u32(sQ);
  // ldmatrix row addresses of this lane: staged rows as four 8x8 blocks (rows 0-7 | 8-15) x (columns 0-7 | 8-15) of a
  // 16-column step, query rows as two blocks (columns 0-7 | 8-15)
  const uint32_t r_lm = r_saddr + (uint32_t)(((lane >> 4) << 3) + (lane & 7)) * DT_RSTRIDE + (uint32_t)((lane >> 3) & 1) * 16u;
  const uint32_t q_lm = q_saddr + (uint32_t)(lane & 7) * DT_RSTRIDE + (uint32_t)((lane >> 3) & 1) * 16u;
  // feasible actions of rollout b in ascending order -> s_list (action | slot << 10), s_nb; returns their number
  auto build_list = [&](int b, bool init_logits) {
    int cnt = 0;
    for (int e00 = 0; e00 < A; e00 += 128) {        // 128 actions per round trip
      int nb[4], mk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = e00 + 32 * j + lane;
        nb[j] = 0;
        mk[j] = 1;
        if (e < A) {
          nb[j] = st.nbh[(size_t)b * A + e];
          mk[j] = st.mask[(size_t)b * A + e];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = e00 + 32 * j + lane;
        if (e < A) {
          s_nb[e] = (int16_t)nb[j];
          if (init_logits) s_logit[e] = -INFINITY;
        }
        const bool ok = mk[j] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) s_list[cnt + __popc(m & ((1u << lane) - 1u))] = (int16_t)(e | ((e / k) << 10));      // action | slot << 10
        cnt += __popc(m);
      }
    }
    __syncwarp();
    return cnt;
  };
  // rows of the feasible actions i0 .. i0+15 -> sR (rows past the end are zero-filled: they enter the glimpse MMA with
  // weight 0 and must be finite)
  auto stage_issue = [&](const uint8_t* rows_b, int i0, int cnt) {
    int goff = 0;
    if (lane < DT_RROWS) goff = (int)s_nb[s_list[min(i0 + lane, cnt - 1)] & 1023] * 512;
#pragma unroll
    for (int r = 0; r < DT_RROWS; ++r) {
      const int off = __shfl_sync(0xffffffffu, goff, r);
      const uint32_t nbytes = (i0 + r < cnt) ? 16u : 0u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(r_saddr + (uint32_t)r * DT_RSTRIDE + (uint32_t)lane * 16u),
                   "l"(rows_b + off), "r"(nbytes)
                   : "memory");
    }
  };
  auto stage_wait = [&]() {
    cp_async_wait_all();
    __syncwarp();
  };
  // S[v][a] = sum_c Q[v][c] R[a][c] for the 16 staged rows: this lane ends up with the full dot products (hi + lo) of
  // query vector g8 & 3 for the chunk positions 2 t4, 2 t4 + 1, 8 + 2 t4, 9 + 2 t4
  // (mma.sync has a dependent-issue latency of ~260 cycles on this part, tools/ubench/hmma_rate.cu: the 16 K steps run as
  // two independent accumulation chains per n-tile, summed at the end; four chains spill next to the glimpse accumulators)
  auto chunk_scores = [&](float (&sc)[4]) {
    float acc0[2][4], acc1[2][4];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc0[p][j] = 0.0f; acc1[p][j] = 0.0f; }
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      uint32_t b00, b01, b10, b11, a0, a2;
      ldsm_x4(r_lm + (uint32_t)ks * 32u, b00, b01, b10, b11);
      ldsm_x2(q_lm + (uint32_t)ks * 32u, a0, a2);
      mma16816<T>(acc0[ks & 1], a0, 0u, a2, 0u, b00, b01);
      mma16816<T>(acc1[ks & 1], a0, 0u, a2, 0u, b10, b11);
    }
    const float s0 = acc0[0][0] + acc0[1][0], s1 = acc0[0][1] + acc0[1][1];
    const float s2 = acc1[0][0] + acc1[1][0], s3 = acc1[0][1] + acc1[1][1];
    sc[0] = s0 + __shfl_xor_sync(0xffffffffu, s0, 16);
    sc[1] = s1 + __shfl_xor_sync(0xffffffffu, s1, 16);
    sc[2] = s2 + __shfl_xor_sync(0xffffffffu, s2, 16);
    sc[3] = s3 + __shfl_xor_sync(0xffffffffu, s3, 16);
  };
  // fp32 vector (this lane's 8 columns) -> query rows `row` (hi) and 4 + `row` (lo)
  auto put_query = [&](int row, const float2 (&v)[4]) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const T2 hh = Elem<T>::from2(v[i].x, v[i].y);
      const float2 hf = Elem<T>::to2(hh);
      hi[i] = as_u32(hh);
      lo[i] = as_u32(Elem<T>::from2(v[i].x - hf.x, v[i].y - hf.y));
    }
    *reinterpret_cast<uint4*>(sQ + (size_t)row * DT_RSTRIDE + lane * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(sQ + (size_t)(4 + row) * DT_RSTRIDE + lane * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  };
#pragma unroll 1
  for (;;) {
    // rollouts are claimed one by one: their cost varies with the number of feasible actions (1 .. A) and idle ones
    // cost nothing, a fixed assignment leaves the slowest warp ~25 % behind the average
    int rr = 0;
    if (lane == 0) rr = atomicAdd(&s_next[0], 1);
    rr = __shfl_sync(0xffffffffu, rr, 0);
    const int b = b0 + rr;
    if (rr >= DT_M || b >= bs) break;
    if (jampr_rollout_idle(st, b, K)) continue;
    // everything this rollout needs besides the rows is requested in one round trip: per-head glimpse queries and tour
    // embeddings (lane owns columns 8 lane .. 8 lane + 7), then the action list
    float2 qv[4][4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float4* p = reinterpret_cast<const float4*>(st.dqp + ((size_t)b * 4 + h) * 256 + lane * 8);
      const float4 x = p[0], y = p[1];
      qv[h][0] = make_float2(x.x, x.y); qv[h][1] = make_float2(x.z, x.w);
      qv[h][2] = make_float2(y.x, y.y); qv[h][3] = make_float2(y.z, y.w);
    }
    float4 tex[4], tey[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      tex[t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      tey[t] = tex[t];
      if (t < K) {
        const float4* p = reinterpret_cast<const float4*>(st.dte + ((size_t)b * 4 + t) * 256 + lane * 8);
        tex[t] = p[0];
        tey[t] = p[1];
      }
    }
    const int cnt = build_list(b, false);
    const uint8_t* rows_b = reinterpret_cast<const uint8_t*>(ne16 + (size_t)b * N * 256) + lane * 16;
    stage_issue(rows_b, 0, cnt);                 // first chunk in flight during the query preparation
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (t < K) {
        const float2 v0 = make_float2(tex[t].x, tex[t].y), v1 = make_float2(tex[t].z, tex[t].w);
        const float2 v2 = make_float2(tey[t].x, tey[t].y), v3 = make_float2(tey[t].z, tey[t].w);
        float dd[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float2 e = __ffma2_rn(v0, qv[h][0], make_float2(0.0f, 0.0f));
          e = __ffma2_rn(v1, qv[h][1], e);
          e = __ffma2_rn(v2, qv[h][2], e);
          e = __ffma2_rn(v3, qv[h][3], e);
          dd[h] = e.x + e.y;
        }
        const float z = warp_sum4(dd[0], dd[1], dd[2], dd[3], lane);
        if ((lane & 7) == 0) s_qt[(((lane >> 4) << 1) | ((lane >> 3) & 1)) * 4 + t] = z;
      }
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) put_query(h, qv[h]);
    __syncwarp();
    // this lane's head in the score fragments, its tour terms qp_h . tour_emb[t]
    const int hh = g8 & 3;
    float qt_r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) qt_r[q] = (q < K) ? s_qt[hh * 4 + q] : 0.0f;
    // online softmax over the chunks: running maximum m_run, this lane's partial denominator l_part and per-slot
    // partial weight sums ps[] (all relative to m_run), glimpse accumulators G[c][v] (v < 4: hi of p, v >= 4: lo of p)
    float m_run = -INFINITY, l_part = 0.0f, ps[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float G[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { G[i][0] = 0.0f; G[i][1] = 0.0f; G[i][2] = 0.0f; G[i][3] = 0.0f; }
#pragma unroll 1
    for (int i0 = 0; i0 < cnt; i0 += DT_RROWS) {
      if (i0 > 0) stage_issue(rows_b, i0, cnt);
      stage_wait();
      float sc[4];
      chunk_scores(sc);
      const uint32_t lwA = *reinterpret_cast<const uint32_t*>(s_list + i0 + 2 * t4);
      const uint32_t lwB = *reinterpret_cast<const uint32_t*>(s_list + i0 + 8 + 2 * t4);
      const int slot[4] = {(int)((lwA >> 10) & 63u), (int)(lwA >> 26), (int)((lwB >> 10) & 63u), (int)(lwB >> 26)};
      const int pos0 = i0 + 2 * t4;
      const bool ok[4] = {pos0 < cnt, pos0 + 1 < cnt, pos0 + 8 < cnt, pos0 + 9 < cnt};
      float mloc = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float qt = (slot[j] == 0) ? qt_r[0] : (slot[j] == 1) ? qt_r[1] : (slot[j] == 2) ? qt_r[2] : qt_r[3];
        sc[j] = ok[j] ? (sc[j] + qt) * 0.125f : -INFINITY;
        mloc = fmaxf(mloc, sc[j]);
      }
      mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 1));
      mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 2));
      const float m_new = fmaxf(m_run, mloc);             // finite: position i0 of the chunk is valid
      const float rs = expf(m_run - m_new);               // first chunk: exp(-inf) = 0
      m_run = m_new;
      float pj[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) pj[j] = expf(sc[j] - m_new);
      l_part = fmaf(l_part, rs, (pj[0] + pj[1]) + (pj[2] + pj[3]));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float add = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) add += (slot[j] == q) ? pj[j] : 0.0f;
        ps[q] = fmaf(ps[q], rs, add);
      }
      // probabilities as the B operand: lanes with g8 < 4 carry the hi parts, the others the lo parts
      const T2 h0 = Elem<T>::from2(pj[0], pj[1]), h1 = Elem<T>::from2(pj[2], pj[3]);
      const float2 f0 = Elem<T>::to2(h0), f1 = Elem<T>::to2(h1);
      const T2 l0 = Elem<T>::from2(pj[0] - f0.x, pj[1] - f0.y), l1 = Elem<T>::from2(pj[2] - f1.x, pj[3] - f1.y);
      const uint32_t pb0 = (g8 < 4) ? as_u32(h0) : as_u32(l0), pb1 = (g8 < 4) ? as_u32(h1) : as_u32(l1);
      // rescale the accumulators: this lane's columns belong to heads (2 t4) & 3 and (2 t4 + 1) & 3
      const int ha = (2 * t4) & 3;
      const float rsa = __shfl_sync(0xffffffffu, rs, 4 * ha), rsb = __shfl_sync(0xffffffffu, rs, 4 * ha + 4);
      if (!__all_sync(0xffffffffu, rs == 1.0f)) {
        const float2 r2 = make_float2(rsa, rsb);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 u = __fmul2_rn(make_float2(G[i][0], G[i][1]), r2), v = __fmul2_rn(make_float2(G[i][2], G[i][3]), r2);
          G[i][0] = u.x; G[i][1] = u.y; G[i][2] = v.x; G[i][3] = v.y;
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4_t(r_lm + (uint32_t)i * 32u, a0, a1, a2, a3);
        mma16816<T>(G[i], a0, a1, a2, a3, pb0, pb1);
      }
      __syncwarp();
    }
    // denominators and slot weights of this lane's head
    l_part += __shfl_xor_sync(0xffffffffu, l_part, 1);
    l_part += __shfl_xor_sync(0xffffffffu, l_part, 2);
    const float inv = 1.0f / l_part;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ps[q] += __shfl_xor_sync(0xffffffffu, ps[q], 1);
      ps[q] += __shfl_xor_sync(0xffffffffu, ps[q], 2);
      ps[q] *= inv;
    }
    // accumulators -> [8][256] fp32 in the (dead) row buffer, then per lane 8 columns x 4 heads:
    // gbar_h = (hi + lo) / denominator + sum_t P_h[t] tour_emb[t]   (tour embeddings re-requested first: L1 / L2 hits
    // whose latency overlaps the hand-over)
    {
      float* sG = reinterpret_cast<float*>(sR);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        sG[(2 * t4) * DT_GSTRIDE + 16 * i + g8] = G[i][0];
        sG[(2 * t4 + 1) * DT_GSTRIDE + 16 * i + g8] = G[i][1];
        sG[(2 * t4) * DT_GSTRIDE + 16 * i + g8 + 8] = G[i][2];
        sG[(2 * t4 + 1) * DT_GSTRIDE + 16 * i + g8 + 8] = G[i][3];
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (t < K) {
          const float4* p = reinterpret_cast<const float4*>(st.dte + ((size_t)b * 4 + t) * 256 + lane * 8);
          tex[t] = p[0];
          tey[t] = p[1];
        }
      }
      __syncwarp();
      float2 gb[4][4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float ih = __shfl_sync(0xffffffffu, inv, 4 * h);
        const float4 x = *reinterpret_cast<const float4*>(sG + h * DT_GSTRIDE + lane * 8);
        const float4 y = *reinterpret_cast<const float4*>(sG + h * DT_GSTRIDE + lane * 8 + 4);
        const float4 u = *reinterpret_cast<const float4*>(sG + (4 + h) * DT_GSTRIDE + lane * 8);
        const float4 v = *reinterpret_cast<const float4*>(sG + (4 + h) * DT_GSTRIDE + lane * 8 + 4);
        gb[h][0] = make_float2((x.x + u.x) * ih, (x.y + u.y) * ih);
        gb[h][1] = make_float2((x.z + u.z) * ih, (x.w + u.w) * ih);
        gb[h][2] = make_float2((y.x + v.x) * ih, (y.y + v.y) * ih);
        gb[h][3] = make_float2((y.z + v.z) * ih, (y.w + v.w) * ih);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (t < K) {
          const float2 t0 = make_float2(tex[t].x, tex[t].y), t1 = make_float2(tex[t].z, tex[t].w);
          const float2 t2 = make_float2(tey[t].x, tey[t].y), t3 = make_float2(tey[t].z, tey[t].w);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 p2 = f2(__shfl_sync(0xffffffffu, ps[t], 4 * h));
            gb[h][0] = __ffma2_rn(p2, t0, gb[h][0]);
            gb[h][1] = __ffma2_rn(p2, t1, gb[h][1]);
            gb[h][2] = __ffma2_rn(p2, t2, gb[h][2]);
            gb[h][3] = __ffma2_rn(p2, t3, gb[h][3]);
          }
        }
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        float4* o = reinterpret_cast<float4*>(st.dgbar + ((size_t)b * 4 + h) * 256 + lane * 8);
        o[0] = make_float4(gb[h][0].x, gb[h][0].y, gb[h][1].x, gb[h][1].y);
        o[1] = make_float4(gb[h][2].x, gb[h][2].y, gb[h][3].x, gb[h][3].y);
      }
    }
    __syncwarp();
  }
  fence_proxy_async();           // the warps' generic-proxy writes over the operand tiles precede the next bulk copies
  __threadfence_block();
  __syncthreads();
  if (tid == 0) issue_b(n_t + 12);
  DTL(5);

  // ================================================================ stage D: g (TMEM columns 0..255, 64 per head)
  {
    float4 va[4], vb[4];
    load_global(va, st.dgbar, 1024, 0);
    load_global(vb, st.dgbar, 1024, 64);
#pragma unroll 1
    for (int e = 0; e < 16; e += 2) {
      const int h = e >> 2, kb = e & 3, i = n_t + 12 + e;      // items e, e + 1 belong to the same head
      wait_item(i - 2);
      store_tile(i, va);
      issue_item(i, (uint32_t)h * 64u, idesc64, kb == 0);
      if (e + 2 < 16) load_global(va, st.dgbar + (size_t)((e + 2) >> 2) * 256, 1024, ((e + 2) & 3) * 64);
      wait_item(i - 1);
      store_tile(i + 1, vb);
      issue_item(i + 1, (uint32_t)h * 64u, idesc64, false);
      if (e + 3 < 16) load_global(vb, st.dgbar + (size_t)((e + 3) >> 2) * 256, 1024, ((e + 3) & 3) * 64);
    }
  }
  DTL(6);
  // ================================================================ stage E: lp (TMEM columns 256..511) -> st.dlp
  {
    const int i0 = n_t + 28;
    wait_item(i0 - 1);           // g complete
#pragma unroll 1
    for (int kb = 0; kb < 4; ++kb) {
      if (kb >= 2) wait_item(i0 + kb - 2);
      fill_from_tmem(i0 + kb, (uint32_t)kb * 64u);
      issue_item(i0 + kb, 256u, idesc256, kb == 0);
    }
    wait_item(i0 + 3);
    drain_to_global((i0 + 3) & 1, 256u, st.dlp, 256);
  }
  tc_fence_before();
  __threadfence_block();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
  DTL(7);

  // ================================================================ phase F: logits + selection per rollout
#pragma unroll 1
  for (;;) {
    int rr = 0;
    if (lane == 0) rr = atomicAdd(&s_next[1], 1);
    rr = __shfl_sync(0xffffffffu, rr, 0);
    const int b = b0 + rr;
    if (rr >= DT_M || b >= bs) break;
    if (jampr_rollout_idle(st, b, K)) {
      if (lane == 0) {
        st.action[2 * (size_t)b] = 0;
        st.action[2 * (size_t)b + 1] = 0;
        st.ll[b] = 0.0f;
        st.entropy[b] = 0.0f;
      }
      continue;
    }
    // logit vector and tour embeddings requested together with the action list (one round trip)
    float2 lv[4];
    {
      const float4* p = reinterpret_cast<const float4*>(st.dlp + (size_t)b * 256 + lane * 8);
      const float4 x = p[0], y = p[1];
      lv[0] = make_float2(x.x, x.y); lv[1] = make_float2(x.z, x.w); lv[2] = make_float2(y.x, y.y); lv[3] = make_float2(y.z, y.w);
    }
    float4 tex[4], tey[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      tex[t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      tey[t] = tex[t];
      if (t < K) {
        const float4* p = reinterpret_cast<const float4*>(st.dte + ((size_t)b * 4 + t) * 256 + lane * 8);
        tex[t] = p[0];
        tey[t] = p[1];
      }
    }
    const int cnt = build_list(b, true);
    const uint8_t* rows_b = reinterpret_cast<const uint8_t*>(ne16 + (size_t)b * N * 256) + lane * 16;
    stage_issue(rows_b, 0, cnt);
    float lt[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (t < K) {
        float2 e = __ffma2_rn(make_float2(tex[t].x, tex[t].y), lv[0], make_float2(0.0f, 0.0f));
        e = __ffma2_rn(make_float2(tex[t].z, tex[t].w), lv[1], e);
        e = __ffma2_rn(make_float2(tey[t].x, tey[t].y), lv[2], e);
        e = __ffma2_rn(make_float2(tey[t].z, tey[t].w), lv[3], e);
        lt[t] = warp_sum(e.x + e.y);
      }
    }
    // logit vector as query row 0 (hi) / 4 (lo); the other query rows keep finite leftovers of phase C or zeros, and
    // rows of the score MMA are independent
    put_query(0, lv);
    __syncwarp();
#pragma unroll 1
    for (int i0 = 0; i0 < cnt; i0 += DT_RROWS) {
      if (i0 > 0) stage_issue(rows_b, i0, cnt);
      stage_wait();
      float sc[4];
      chunk_scores(sc);
      if (g8 == 0) {          // raw dot products; the clip follows, vectorised
        const int pos0 = i0 + 2 * t4;
        if (pos0 < cnt) s_logit[s_list[pos0] & 1023] = sc[0];
        if (pos0 + 1 < cnt) s_logit[s_list[pos0 + 1] & 1023] = sc[1];
        if (pos0 + 8 < cnt) s_logit[s_list[pos0 + 8] & 1023] = sc[2];
        if (pos0 + 9 < cnt) s_logit[s_list[pos0 + 9] & 1023] = sc[3];
      }
      __syncwarp();
    }
    for (int i = lane; i < cnt; i += 32) {
      const int v = s_list[i], e = v & 1023, t = v >> 10;
      const float ltv = (t == 0) ? lt[0] : (t == 1) ? lt[1] : (t == 2) ? lt[2] : lt[3];
      s_logit[e] = 10.0f * tanhf((s_logit[e] + ltv) * 0.0625f);
    }
    __syncwarp();
    // ---- masked log-softmax + selection, same rules as decoder.cu / policy_tc2.cu
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
    float ent = 0.0f, best = -INFINITY;
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    __syncwarp();
    int choice = besti;
    if (a.decode_type == JAMPR_DECODE_SAMPLING) {
      const float u = a.uniforms ? a.uniforms[b] : jampr_uniform(a.seed, (uint32_t)b, (uint32_t)a.step_no);
      float run = 0.0f;
      int hit = -1, lastf = -1;
      for (int bs0 = 0; bs0 < A; bs0 += 32) {
        const int e = bs0 + lane;
        const float pe = (e < A) ? expf(s_logit[e]) : 0.0f;
        float inc = pe;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float tt = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += tt;
        }
        const float cum = run + inc;
        const unsigned hm = __ballot_sync(0xffffffffu, e < A && pe > 0.0f && cum > u);
        const unsigned fm = __ballot_sync(0xffffffffu, e < A && pe > 0.0f);
        if (hit < 0 && hm) hit = bs0 + __ffs(hm) - 1;
        if (fm) lastf = bs0 + 31 - __clz(fm);
        run = __shfl_sync(0xffffffffu, cum, 31);
      }
      choice = (hit >= 0) ? hit : lastf;
      if (choice < 0) choice = 0;
    }
    if (lane == 0) {
      if (choice >= A) choice = 0;
      st.action[2 * (size_t)b] = choice / k;
      st.action[2 * (size_t)b + 1] = s_nb[choice];
      st.ll[b] = s_logit[choice];
      st.entropy[b] = ent;
      if (nan) atomicOr(st.status + b, JAMPR_ST_NAN_LOGITS);
    }
    __syncwarp();
  }
  DTL(8);
}

#ifdef JAMPR_TIMELINE
extern "C" __attribute__((visibility("default"))) int jampr_debug_dtimeline(uint32_t* host_out) {
  CUDA_RET(cudaDeviceSynchronize());
  CUDA_RET(cudaMemcpyFromSymbol(host_out, g_dtimeline, sizeof(g_dtimeline)));
  return 0;
}
#endif

static size_t dt_smem_bytes(const jampr_dims_t* d) { return 1024 + (size_t)dt_region_bytes(d->n_slots * d->k_nbh) + 64; }

template <typename T>
static int launch_dt(DtArgs& a, size_t smem, cudaStream_t s) {
  CUDA_RET(cudaFuncSetAttribute(decoder_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int bs = a.d.n_inst * a.d.n_samples;
  // One CTA per SM: 128-rollout tiles would need ceil(bs / (128 n_sm)) waves with the last one partly empty (3.46 waves at
  // 65,536 rollouts on 148 SMs).  The same number of waves is filled evenly with slightly smaller tiles (111 rollouts):
  // the GEMM stages still run M = 128, but the per-rollout phases, two thirds of a tile's time, shrink with the tile.
  int dev = 0, n_sm = 0;
  CUDA_RET(cudaGetDevice(&dev));
  CUDA_RET(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const int waves = (bs + DT_M * n_sm - 1) / (DT_M * n_sm);
  int tr = (bs + waves * n_sm - 1) / (waves * n_sm);
  tr = tr < 1 ? 1 : (tr > DT_M ? DT_M : tr);
  a.tile_rollouts = tr;
  decoder_tc_kernel<T><<<(bs + tr - 1) / tr, DT_THREADS, smem, s>>>(a);
  return launch_status();
}

// Returns JAMPR_E_TOOLARGE when the configuration does not fit (caller keeps the fused decoder).
int jampr_decoder_tc(const jampr_dims_t* d, const jampr_weights_t* w, const jampr_state_t* st, int step_no,
                     int decode_type, uint64_t seed, const float* uniforms, cudaStream_t s) {
  if (d->n_slots > 4 || d->n_slots * d->k_nbh > 1023) return JAMPR_E_TOOLARGE;
  if (!st->ne16 || !st->dq || !st->dte || !st->dqp || !st->dgbar || !st->dlp) return JAMPR_E_BADARG;
  if (!w->blob_bf16 || w->n_bf16 < (int64_t)W16_END) return JAMPR_E_BADARG;
  const size_t smem = dt_smem_bytes(d);
  if (smem > 227 * 1024) return JAMPR_E_TOOLARGE;
  DtArgs a;
  a.d = *d;
  a.st = *st;
  a.w = w->blob;
  a.w16 = w->blob_bf16;
  a.uniforms = uniforms;
  a.seed = seed;
  a.step_no = step_no;
  a.decode_type = decode_type;
  if (w->precision == JAMPR_PREC_F16) return launch_dt<__half>(a, smem, s);
  return launch_dt<__nv_bfloat16>(a, smem, s);
}

int jampr_decoder_tc_fits(const jampr_dims_t* d) {
  return d->n_slots <= 4 && d->n_slots * d->k_nbh <= 1023 && dt_smem_bytes(d) <= 227 * 1024;      // action index: 10 bits
}
