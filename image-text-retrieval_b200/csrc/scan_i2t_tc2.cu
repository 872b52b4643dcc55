// SCAN image-to-text cross-attention scores, fused on the CTA-pair tcgen05 main loop (sm_100a only).
//
// Replaces xattn_score_i2t + func_attention + cosine_similarity (itr/modalmodule/Objectives.py:376-417, 421-476, 10-15)
// for raw_feature_norm in {clipped_l2norm, l2norm}, every agg_func, captions of up to 32 words (longer captions are left
// to the two-phase path: their columns are not written).  For image i (regions v_k) and caption c (words w_j, j < n):
//     A[k][j]  = v_k . w_j                                   tensor cores, exactly the t2i main loop (scan_t2i_tc2.cu)
//     ahat[j][k] = leaky(A) / (sqrt(sum_k leaky(A[k][j])^2) + 1e-8)      per WORD, over the regions: in-thread
//     e[j][k]  = exp2(lambda log2e (ahat - 1))               softmax numerators over the caption's WORDS, per region
//     Z_k = sum_j e,  P_k = sum_j e A,  Q_k = sum_j e[j][k] U[j][k],  U[j][k] = sum_j' G_c[j][j'] e[j'][k]
//     r_k      = P_k / max(|v_k| sqrt(Q_k), 1e-8 Z_k)        cos(v_k, sum_j alpha_kj w_j) with |ctx|^2 = alpha^T G_c alpha
//     S[i][c]  = aggregate over the 36 regions               LSE / Mean / Max / Sum
// G_c = W_c W_c^T is the caption's word Gram (fp32, from the bf16-rounded words), handed in "caption-relative" form
// gq_rel[tile][d][row] = w_row . w_(d-th word of row's caption) (0 beyond the caption) so that lanes read it coalesced.
//
// Where t2i reduces over regions (in-thread, TMEM lane = word) and needs one cross-lane sum per region, i2t needs three
// (Z, P, Q) plus the contraction with G_c.  Per (warp = 32 word rows, image), with two 4.6 KB shared-memory scratches:
//   1. e -> scratch ES (fp32, [row][36]); t = e A is parked in spare tensor-memory columns;
//   2. U by ONE loop over the caption's words (a coalesced Gram value requested a step ahead, nine 16-byte broadcast reads
//      of the partner row's e, 36 FMAs per step); y = e U; then for each third of the regions (12): rows (t, y | e) ->
//      scratch WK and 18 lanes walk the 32 rows restarting at caption ends (the row walk of the t2i kernel, three
//      quantities at once) and leave the caption's totals P | Q | Z in its last row; the lane of that last word finishes
//      r_k for the twelve regions and folds them into its aggregate -- one scalar per caption, in a register, is all that
//      survives between the thirds;
//   3. that lane turns the aggregate into the score and stores it.
// Main loop, barriers and cluster protocol: scan_t2i_tc2.cu (no Gram MMA, no parking: the accumulator is free again as
// soon as the 32 epilogue warps of the pair have loaded it).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc2_common.cuh"

namespace itr {
namespace tc2i {

using namespace itr::tc;
using namespace itr::tc2;

constexpr int STAGES = 3;
// word-tile pairs per work unit: items differ a lot in cost here (the loop over a caption's words is as long as the
// longest caption of the warp) and an evaluation fold is small, so units are kept short for the static schedule's balance
constexpr int I2T_BAND = 32;
constexpr int AUX_BYTES = BLOCK_M * 16;            // row metadata of the CTA's word tile
constexpr int ROW_BYTES = 36 * 4;                  // one scratch row: 36 floats
constexpr int WARP_SCRATCH = 32 * ROW_BYTES;       // 4608
constexpr int SMEM_STAGES = 0;
constexpr int SMEM_AUX = SMEM_STAGES + STAGES * STAGE_BYTES;
constexpr int SMEM_ES = SMEM_AUX + 2 * AUX_BYTES;
constexpr int SMEM_WK = SMEM_ES + NUM_EPI_WARPS * WARP_SCRATCH;
constexpr int SMEM_BARS = SMEM_WK + NUM_EPI_WARPS * WARP_SCRATCH;
constexpr int NUM_BARS = 2 * STAGES + 8;
constexpr int SMEM_TMEMPTR = SMEM_BARS + NUM_BARS * 8;
constexpr int SMEM_BYTES = SMEM_TMEMPTR + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;
static_assert(SMEM_ALLOC <= 232448, "shared memory budget");
static_assert(SMEM_AUX % 16 == 0 && SMEM_ES % 16 == 0 && SMEM_WK % 16 == 0 && SMEM_BARS % 8 == 0, "alignment");

struct Params {
  const int4* row_meta;        // [n_wt*128]
  const float* gq_rel;         // [n_wt][32][128] caption-relative word Gram
  const float* vnorm;          // [n_img][36] region norms
  int n_img, n_wt, n_wp, n_it;
  int agg;
  float c_sm, c_lse, inv_lse;
  float* scores; long long ld;
};

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f4(uint32_t addr, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ int4 lds_i4(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

template <bool CLIPPED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
scan_i2t_tc2_kernel(const __grid_constant__ CUtensorMap map_words, const __grid_constant__ CUtensorMap map_imgs, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = (int)((threadIdx.x >> 5) + EPI_WARP0) % (NUM_THREADS / 32);      // control warpgroup physically last
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const uint32_t bar0 = sbase + SMEM_BARS;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };                          // leader: its producer's arrival + both CTAs' bytes
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };              // both (multicast commit)
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 0 + b); };      // both (multicast commit)
  auto loaded_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 + b); };     // leader: 32 epilogue warps
  auto afull_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 4 + b); };      // local
  auto aempty_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 6 + b); };     // local

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_words) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_imgs) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1); mbar_init(loaded_bar(b), 2 * NUM_EPI_WARPS);
      mbar_init(afull_bar(b), 1); mbar_init(aempty_bar(b), NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SMEM_TMEMPTR), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  constexpr uint32_t tmem_base = 0u;       // the pair owns both SMs and asks for all 512 columns

  using Schedule = ScheduleT<false>;
  using ItemIter = ItemIterT<false>;
  const Schedule sched(p.n_wp, p.n_it, nullptr, 0, I2T_BAND);
  const int first = (int)(blockIdx.x >> 1);
  const int step = (int)(gridDim.x >> 1);

  if (warp < EPI_WARP0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ==============================
    int stage = 0; uint32_t phase = 0;
    for (ItemIter item(sched, first, step); item.valid(); item.next()) {
      const int row_w = item.tile((int)rank) * BLOCK_M;
      const int row_i = item.n * BLOCK_N + (int)rank * HALF_N;
#pragma unroll 1
      for (int kb = 0; kb < K_BLOCKS; ++kb) {
        mbar_wait_sleep(empty_bar(stage), phase ^ 1);
        const uint32_t sa = sbase + SMEM_STAGES + stage * STAGE_BYTES, fb = full_bar(stage);
        if (elect_one()) {
          if (leader) mbar_expect_tx(fb, 2 * STAGE_BYTES);
          tma_load_2d_pair(sa, &map_words, fb, kb * BLOCK_K, row_w);
          tma_load_2d_pair(sa + A_BYTES, &map_imgs, fb, kb * BLOCK_K, row_i);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader only) ==============================
    if (leader) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      const uint64_t adesc0 = umma_desc_sw128(sbase + SMEM_STAGES);
      const uint64_t bdesc0 = umma_desc_sw128(sbase + SMEM_STAGES + A_BYTES);
      for (ItemIter item(sched, first, step); item.valid(); item.next(), ++it) {
        const int ab = it & 1;
        const uint32_t tacc = tmem_base + ab * ACC_PITCH;
        mbar_wait_sleep(loaded_bar(ab), ((it >> 1) & 1) ^ 1);      // every epilogue warp of the pair holds item it-2 in registers
        tc_fence_after();
#pragma unroll 1
        for (int kb = 0; kb < K_BLOCKS; ++kb) {
          mbar_wait_sleep(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t soff = (uint64_t)((uint32_t)stage * (uint32_t)(STAGE_BYTES >> 4));
          const uint64_t adesc = adesc0 + soff, bdesc = bdesc0 + soff;
          if (elect_one()) {
            umma2_bf16(tacc, adesc, bdesc, IDESC, (uint32_t)kb);
            umma2_bf16(tacc, adesc + 2, bdesc + 2, IDESC, 1u);
            umma2_bf16(tacc, adesc + 4, bdesc + 4, IDESC, 1u);
            umma2_bf16(tacc, adesc + 6, bdesc + 6, IDESC, 1u);
            umma2_commit_both(empty_bar(stage));
            if (kb == K_BLOCKS - 1) umma2_commit_both(tfull_bar(ab));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // =============================== aux loader (both CTAs): row metadata ==================
    int it = 0;
    for (ItemIter item(sched, first, step); item.valid(); item.next(), ++it) {
      const int m = item.tile((int)rank);
      const int b = it & 1;
      mbar_wait_sleep(aempty_bar(b), ((it >> 1) & 1) ^ 1);
      if (elect_one()) {
        if (m < p.n_wt) {
          mbar_expect_tx(afull_bar(b), AUX_BYTES);
          bulk_load(sbase + SMEM_AUX + b * AUX_BYTES, p.row_meta + (size_t)m * BLOCK_M, AUX_BYTES, afull_bar(b));
        } else {
          mbar_arrive(afull_bar(b));
        }
      }
      __syncwarp();
    }
  }
  } else {
    // =============================== epilogue (both CTAs) ==================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int q = warp & 3;
    const int g = (warp - EPI_WARP0) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t es = sbase + SMEM_ES + (uint32_t)(warp - EPI_WARP0) * WARP_SCRATCH;
    const uint32_t wk = sbase + SMEM_WK + (uint32_t)(warp - EPI_WARP0) * WARP_SCRATCH;
    const uint32_t tpark = tmem_base + 2 * ACC_PITCH + 56 * g + lane_sel;      // 36 of the group's 56 spare TMEM columns
    const float shift = -fabsf(p.c_sm);
    const float agg_identity = (p.agg == ITR_AGG_MAX) ? -INFINITY : 0.f;

    int it = 0;
    for (ItemIter item(sched, first, step); item.valid(); item.next(), ++it) {
      const int n = item.n;
      const int m = item.tile((int)rank);
      const bool word_ok = m < p.n_wt;
      const int b = it & 1;
      mbar_wait_sleep(afull_bar(b), (it >> 1) & 1);
      int4 meta = make_int4(-1, 0, lane | (lane << 8), 0);
      if (word_ok) meta = lds_i4(sbase + SMEM_AUX + b * AUX_BYTES + 16u * (uint32_t)row);
      const int seg_lo = meta.z & 0xff, seg_hi = (meta.z >> 8) & 0xff;
      // The buffer may be refilled only once every lane HOLDS its metadata: an arrive issued while the loads are still in
      // flight lets the refill overtake them (seen as rare stale rows).  The vote consumes every lane's value, and the
      // arrive depends on its result (lane 31 always ends a caption or is padding, so the mask is never 0).
      const uint32_t endmask = __ballot_sync(0xffffffffu, lane == seg_hi);
      if (lane == 0 && endmask != 0u) mbar_arrive(aempty_bar(b));
      const bool long_tile = (meta.z >> 16) & 1;
      const int img = n * IMGS + g;
      const bool valid = img < p.n_img && word_ok && !long_tile;      // warp-uniform

      const uint32_t tacc = tmem_base + b * ACC_PITCH + lane_sel;
      mbar_wait_sleep(tfull_bar(b), (it >> 1) & 1);
      tc_fence_after();
      float A[R];
      TMEM_LD_X32(tacc + g * R, A, 0);
      TMEM_LD_X4(tacc + g * R + 32, A, 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(loaded_bar(b), leader);
      if (!valid) continue;

      // ---- 1. per-word l2norm over the regions, softmax numerators e (-> ES), t = e A (-> parked in tensor memory) -----
      float n2 = 0.f;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const float a = CLIPPED ? fmaxf(A[k], 0.1f * A[k]) : A[k];
        n2 = fmaf(a, a, n2);
      }
      const float cw = __fdividef(p.c_sm, sqrtf(n2) + 1e-8f);
      const uint32_t myrow = es + (uint32_t)lane * ROW_BYTES;
#pragma unroll
      for (int k = 0; k < R; k += 4) {
        float e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float raw = A[k + i];
          const float a = CLIPPED ? fmaxf(raw, 0.1f * raw) : raw;
          e[i] = ex2f(fmaf(a, cw, shift));
          A[k + i] = e[i] * raw;                           // t
        }
        sts_f4(myrow + 4 * k, e[0], e[1], e[2], e[3]);
      }
      {
        // t leaves the registers: 36 of the group's 56 spare tensor-memory columns (the accumulators end at column 288;
        // the accumulator itself may be overwritten by the MMA of item t+2 as soon as every warp has loaded it)
        uint32_t tv[R];
#pragma unroll
        for (int k = 0; k < R; ++k) tv[k] = __float_as_uint(A[k]);
        TMEM_ST_X16(tpark, tv, 0);
        TMEM_ST_X16(tpark + 16, tv, 16);
        TMEM_ST_X4(tpark + 32, tv, 32);
      }
      const int len = meta.x >= 0 ? meta.w : 0;
      const int maxlen = __reduce_max_sync(0xffffffffu, len);
      tmem_st_wait();
      __syncwarp();

      // ---- 2a. U[j][k] = sum over the caption's words j' of G_c[j][j'] e[j'][k], all 36 regions: one coalesced Gram value
      // and nine 16-byte broadcast reads of the partner row per step; the next step's Gram value is requested a step ahead
      float U[R];
#pragma unroll
      for (int k = 0; k < R; ++k) U[k] = 0.f;
      {
        const float* gq = p.gq_rel + (size_t)m * (32 * BLOCK_M) + row;
        float gv = maxlen > 0 ? gq[0] : 0.f;
#pragma unroll 1
        for (int dlt = 0; dlt < maxlen; ++dlt) {
          const float gnext = dlt + 1 < maxlen ? gq[(dlt + 1) * BLOCK_M] : 0.f;
          const uint32_t src = es + (uint32_t)min(seg_lo + dlt, 31) * ROW_BYTES;
          // two reads in flight: the next 16 bytes are requested before the FMAs of the current ones
          float4 cur = lds_f4(src);
#pragma unroll
          for (int c4 = 0; c4 < R / 4; ++c4) {
            float4 nxt = cur;
            if (c4 + 1 < R / 4) nxt = lds_f4(src + 16 * (c4 + 1));
            U[4 * c4 + 0] = fmaf(gv, cur.x, U[4 * c4 + 0]); U[4 * c4 + 1] = fmaf(gv, cur.y, U[4 * c4 + 1]);
            U[4 * c4 + 2] = fmaf(gv, cur.z, U[4 * c4 + 2]); U[4 * c4 + 3] = fmaf(gv, cur.w, U[4 * c4 + 3]);
            cur = nxt;
          }
          gv = gnext;
        }
        // y = e U (this word's own e from its ES row)
#pragma unroll
        for (int c4 = 0; c4 < R / 4; ++c4) {
          const float4 e4 = lds_f4(myrow + 16 * c4);
          U[4 * c4 + 0] *= e4.x; U[4 * c4 + 1] *= e4.y; U[4 * c4 + 2] *= e4.z; U[4 * c4 + 3] *= e4.w;
        }
      }

      // per-caption aggregate over the regions, held by the lane of the caption's last word
      const bool end_lane = lane == seg_hi && meta.x >= 0;
      float aggacc = agg_identity;
#pragma unroll
      for (int T = 0; T < 3; ++T) {
        // ---- 2b. this word's row of the walk scratch, twelve regions: [t | y | e] -----------------------------------
        {
          float t[12];
          TMEM_LD_X4(tpark + 12 * T, t, 0);
          TMEM_LD_X4(tpark + 12 * T + 4, t, 4);
          TMEM_LD_X4(tpark + 12 * T + 8, t, 8);
          const uint32_t own = myrow + 48u * T;
          const float4 e0 = lds_f4(own), e1 = lds_f4(own + 16), e2 = lds_f4(own + 32);
          const uint32_t dst = wk + (uint32_t)lane * ROW_BYTES;
          sts_f4(dst + 48, U[12 * T + 0], U[12 * T + 1], U[12 * T + 2], U[12 * T + 3]);
          sts_f4(dst + 64, U[12 * T + 4], U[12 * T + 5], U[12 * T + 6], U[12 * T + 7]);
          sts_f4(dst + 80, U[12 * T + 8], U[12 * T + 9], U[12 * T + 10], U[12 * T + 11]);
          sts_f4(dst + 96, e0.x, e0.y, e0.z, e0.w);
          sts_f4(dst + 112, e1.x, e1.y, e1.z, e1.w);
          sts_f4(dst + 128, e2.x, e2.y, e2.z, e2.w);
          tmem_ld_wait();
          sts_f4(dst, t[0], t[1], t[2], t[3]);
          sts_f4(dst + 16, t[4], t[5], t[6], t[7]);
          sts_f4(dst + 32, t[8], t[9], t[10], t[11]);
        }
        __syncwarp();
        // ---- 2c. row walk (as in the t2i kernel): 18 lanes, two columns each, restart at every caption end and leave the
        // caption's totals (P | Q | Z for the twelve regions) in its last row ------------------------------------------
        if (lane < 18) {
          const uint32_t col = wk + 8u * (uint32_t)lane;
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll 1
          for (int j0 = 0; j0 < 32; j0 += 8) {
            float2 v[8];
            const uint32_t blk = col + (uint32_t)(j0 * ROW_BYTES);
            const uint32_t ends = endmask >> j0;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = lds_f2(blk + (uint32_t)(j * ROW_BYTES));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc.x += v[j].x; acc.y += v[j].y;
              if ((ends >> j) & 1u) {
                sts_f2(blk + (uint32_t)(j * ROW_BYTES), acc.x, acc.y);
                acc = make_float2(0.f, 0.f);
              }
            }
          }
        }
        __syncwarp();
        // ---- 2d. the lane of each caption's last word finishes r_k for the twelve regions and folds them into its
        // aggregate: r_k = P_k / max(|v_k| sqrt(Q_k), 1e-8 Z_k) -------------------------------------------------------
        if (end_lane) {
          const uint32_t tot = wk + (uint32_t)lane * ROW_BYTES;
          const float4* vn4 = reinterpret_cast<const float4*>(p.vnorm + (size_t)img * R + 12 * T);
#pragma unroll
          for (int c4 = 0; c4 < 3; ++c4) {
            const float4 P4 = lds_f4(tot + 16 * c4), Q4 = lds_f4(tot + 48 + 16 * c4), Z4 = lds_f4(tot + 96 + 16 * c4);
            const float4 v4 = vn4[c4];
            const float Pv[4] = {P4.x, P4.y, P4.z, P4.w}, Qv[4] = {Q4.x, Q4.y, Q4.z, Q4.w}, Zv[4] = {Z4.x, Z4.y, Z4.z, Z4.w};
            const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float r = __fdividef(Pv[i], fmaxf(vv[i] * sqrtf(fmaxf(Qv[i], 0.f)), 1e-8f * Zv[i]));
              if (p.agg == ITR_AGG_MAX) aggacc = fmaxf(aggacc, r);
              else aggacc += (p.agg == ITR_AGG_LSE) ? ex2f(r * p.c_lse) : r;
            }
          }
        }
        __syncwarp();
      }
      // ---- 3. ... and stores the score --------------------------------------------------------------------------------
      if (end_lane) {
        float tot = aggacc;
        if (p.agg == ITR_AGG_LSE) tot = lg2f(tot) * p.inv_lse;
        if (p.agg == ITR_AGG_MEAN) tot = tot * (1.0f / (float)R);
        p.scores[(size_t)img * p.ld + meta.x] = tot;
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// Caption-relative word Gram of every packed word tile: out[tile][d][row] = w_row . w_(first row of row's caption + d)
// for d < caption length, else 0 (padding rows: 0).  One warp per word row; the row stays in registers (bf16 pairs).
__global__ void __launch_bounds__(256)
caption_gram_rel_kernel(const uint16_t* __restrict__ words, const int4* __restrict__ row_meta, int n_tiles, float* __restrict__ out) {
  const int tile = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < BLOCK_M; r += 8) {
    const size_t row = (size_t)tile * BLOCK_M + r;
    const int4 meta = row_meta[row];
    const int len = (meta.x >= 0 && !((meta.z >> 16) & 1)) ? meta.w : 0;       // long tiles are not scored by the fused kernel
    float* dst = out + ((size_t)tile * 32) * BLOCK_M + r;
    uint4 mine[4];                                                             // 1024 bf16 = 32 lanes x 4 x 8
    if (len > 0) {
      const uint4* src = reinterpret_cast<const uint4*>(words + row * D);
#pragma unroll
      for (int v = 0; v < 4; ++v) mine[v] = src[lane + 32 * v];
    }
    for (int dlt = 0; dlt < 32; ++dlt) {
      float s = 0.f;
      if (dlt < len) {
        const uint4* oth = reinterpret_cast<const uint4*>(words + (row - meta.y + dlt) * D);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const uint4 o = oth[lane + 32 * v];
          const uint32_t a[4] = {mine[v].x, mine[v].y, mine[v].z, mine[v].w}, bb[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            s = fmaf(__uint_as_float(a[w] << 16), __uint_as_float(bb[w] << 16), s);
            s = fmaf(__uint_as_float(a[w] & 0xffff0000u), __uint_as_float(bb[w] & 0xffff0000u), s);
          }
        }
        s = warp_sum(s);
      }
      if (lane == 0) dst[(size_t)dlt * BLOCK_M] = s;
    }
  }
}

template <bool CLIPPED>
static int launch(const CUtensorMap& map_w, const CUtensorMap& map_i, const Params& p, cudaStream_t stream) {
  auto kern = scan_i2t_tc2_kernel<CLIPPED>;
  ITR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
  int dev = 0, sms = 0;
  ITR_CHECK_CUDA(cudaGetDevice(&dev));
  ITR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static int max_pairs = -1;
  if (max_pairs < 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(sms & ~1)); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM_ALLOC;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension; attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = sms / 2; }
    max_pairs = n < sms / 2 ? n : sms / 2;
  }
  const long long units = (long long)((p.n_wp + I2T_BAND - 1) / I2T_BAND) * p.n_it;
  if (units <= 0) return ITR_OK;
  const int pairs = (int)(units < max_pairs ? units : max_pairs);
  kern<<<2 * pairs, NUM_THREADS, SMEM_ALLOC, stream>>>(map_w, map_i, p);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

}  // namespace tc2i
}  // namespace itr

using namespace itr;

extern "C" int itr_scan_caption_gram_rel_bf16(const uint16_t* words_bf16, const int32_t* row_meta, int n_tiles, float* gq_rel,
                                              void* stream) {
  ITR_REQUIRE(words_bf16 && row_meta && gq_rel && n_tiles >= 0, "itr_scan_caption_gram_rel_bf16: bad arguments");
  ITR_REQUIRE(((uintptr_t)words_bf16 & 15) == 0 && ((uintptr_t)row_meta & 15) == 0, "itr_scan_caption_gram_rel_bf16: buffers must be 16-byte aligned");
  if (n_tiles == 0) return ITR_OK;
  tc2i::caption_gram_rel_kernel<<<n_tiles, 256, 0, as_stream(stream)>>>(words_bf16, reinterpret_cast<const int4*>(row_meta), n_tiles, gq_rel);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_scan_i2t_scores_bf16(const uint16_t* images_bf16, const float* region_norm, int n_img,
                                        const uint16_t* words_bf16, const int32_t* row_meta, const float* gq_rel, int n_tiles,
                                        int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                                        float* scores, int64_t ld_scores, void* stream) {
  ITR_REQUIRE(images_bf16 && region_norm && words_bf16 && row_meta && gq_rel && scores, "itr_scan_i2t_scores_bf16: null pointer");
  ITR_REQUIRE(feature_norm == ITR_NORM_CLIPPED_L2 || feature_norm == ITR_NORM_L2,
              "itr_scan_i2t_scores_bf16: raw_feature_norm %d is only available in the two-phase / float32 paths", feature_norm);
  ITR_REQUIRE(agg >= 0 && agg <= ITR_AGG_SUM, "unknown aggfunc: %d", agg);
  ITR_REQUIRE(lambda_lse != 0.f || agg != ITR_AGG_LSE, "itr_scan_i2t_scores_bf16: lambda_lse must be non-zero");
  ITR_REQUIRE(lambda_softmax > -80.f && lambda_softmax < 80.f, "itr_scan_i2t_scores_bf16: |lambda_softmax| must be < 80");
  ITR_REQUIRE(((uintptr_t)images_bf16 & 15) == 0 && ((uintptr_t)words_bf16 & 15) == 0 && ((uintptr_t)row_meta & 15) == 0,
              "itr_scan_i2t_scores_bf16: buffers must be 16-byte aligned");
  if (n_img <= 0 || n_tiles <= 0) return ITR_OK;
  int dev = 0, major = 0;
  ITR_CHECK_CUDA(cudaGetDevice(&dev));
  ITR_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(ITR_ERR_UNSUPPORTED, "the tensor-core SCAN path needs an sm_100 device (found sm_%d0)", major);
  CUtensorMap map_w, map_i;
  int rc = tc2::make_map(&map_w, words_bf16, (uint64_t)n_tiles * tc2::BLOCK_M, tc2::BLOCK_M);
  if (rc) return rc;
  rc = tc2::make_map(&map_i, images_bf16, (uint64_t)n_img * tc2::R, tc2::HALF_N);
  if (rc) return rc;
  tc2i::Params p{};
  p.row_meta = reinterpret_cast<const int4*>(row_meta);
  p.gq_rel = gq_rel; p.vnorm = region_norm;
  p.n_img = n_img; p.n_wt = n_tiles; p.n_wp = (n_tiles + 1) / 2; p.n_it = (n_img + tc2::IMGS - 1) / tc2::IMGS;
  p.agg = agg;
  p.c_sm = lambda_softmax * 1.4426950408889634f;
  p.c_lse = lambda_lse * 1.4426950408889634f;
  p.inv_lse = 0.6931471805599453f / lambda_lse;
  p.scores = scores; p.ld = ld_scores;
  if ((long long)p.n_wp * p.n_it >= (1ll << 31))
    return fail(ITR_ERR_INVALID, "itr_scan_i2t_scores_bf16: %lld tile pairs exceed the 2^31 scheduler range; split the call", (long long)p.n_wp * p.n_it);
  cudaStream_t st = as_stream(stream);
  return feature_norm == ITR_NORM_CLIPPED_L2 ? tc2i::launch<true>(map_w, map_i, p, st) : tc2i::launch<false>(map_w, map_i, p, st);
}
