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
// G_c = W_c W_c^T is the caption's word Gram (fp32 accumulation over the bf16-rounded words, rounded to tf32), handed in
// the fragment order of the contraction below (caption_gram_frag_kernel).
//
// Where t2i reduces over regions (in-thread, TMEM lane = word) and needs one cross-lane sum per region, i2t needs three
// (Z, P, Q) plus the contraction with G_c.  Per (warp = 32 word rows, image), with two 4.6 KB shared-memory scratches:
//   1. e (rounded to tf32, the same value everywhere) -> scratch ES (fp32, [row][36]); t = e A is parked in spare
//      tensor-memory columns;
//   2. U = G e on the warp's own tensor cores: mma.sync m16n8k8 tf32, A = the quarter's block-diagonal Gram (eight
//      coalesced 16-byte loads per lane), B = e from ES, 40 MMAs; y = e U at the accumulator positions; then for each half
//      of the regions (18): rows (t | y) -> scratch WK and 27 lanes walk the 32 rows of WK and, for Z, of ES in place,
//      restarting at caption ends (the row walk of the t2i kernel) and leave the caption's totals P | Q | Z in its last
//      row; the lane of that last word finishes r_k for the eighteen regions and folds them into its aggregate -- one
//      scalar per caption, in a register, is all that survives between the halves;
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
constexpr int SMEM_LIST = SMEM_WK + NUM_EPI_WARPS * WARP_SCRATCH;       // per warp: rows of its caption ends, 32 bytes
constexpr int SMEM_BARS = SMEM_LIST + NUM_EPI_WARPS * 32;
constexpr int NUM_BARS = 2 * STAGES + 8;
constexpr int SMEM_TMEMPTR = SMEM_BARS + NUM_BARS * 8;
constexpr int SMEM_BYTES = SMEM_TMEMPTR + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;
static_assert(SMEM_ALLOC <= 232448, "shared memory budget");
static_assert(SMEM_AUX % 16 == 0 && SMEM_ES % 16 == 0 && SMEM_WK % 16 == 0 && SMEM_BARS % 8 == 0, "alignment");

struct Params {
  const int4* row_meta;        // [n_wt*128]
  const float* gq_frag;        // [n_wt][4][2][4][32] float4: word Gram in mma.sync fragment order
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
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// D (16x8, f32) += A (16x8, row) * B (8x8, col), tf32 operands held as f32 bit patterns
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float4& a, float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
                 "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
#define TMEM_LD_X16(taddr, v, o)                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=f"(v[o + 0]), "=f"(v[o + 1]), "=f"(v[o + 2]), "=f"(v[o + 3]), "=f"(v[o + 4]), "=f"(v[o + 5]),      \
                 "=f"(v[o + 6]), "=f"(v[o + 7]), "=f"(v[o + 8]), "=f"(v[o + 9]), "=f"(v[o + 10]), "=f"(v[o + 11]),    \
                 "=f"(v[o + 12]), "=f"(v[o + 13]), "=f"(v[o + 14]), "=f"(v[o + 15])                                   \
               : "r"(taddr))
#define TMEM_LD_X2(taddr, v, o) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=f"(v[o + 0]), "=f"(v[o + 1]) : "r"(taddr))
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
    const uint32_t endlist = sbase + SMEM_LIST + (uint32_t)(warp - EPI_WARP0) * 32;
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
          e[i] = round_tf32(ex2f(fmaf(a, cw, shift)));      // the tensor-core contraction below reads e as tf32: use the same e everywhere
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
      // the quarter's Gram fragments (step 2a) are requested as soon as t has left the registers
      float4 ga[8];
      {
        const float4* gf = reinterpret_cast<const float4*>(p.gq_frag) + ((size_t)(m * 4 + q) * 8) * 32 + lane;
#pragma unroll
        for (int i = 0; i < 8; ++i) ga[i] = __ldg(gf + i * 32);
      }
      // rows of the (real) caption ends, in order: the pairs (caption end, region) of step 2d are dealt to the lanes
      const bool end_lane = lane == seg_hi && meta.x >= 0;
      const uint32_t endv = __ballot_sync(0xffffffffu, end_lane);
      const int npairs = __popc(endv) * 18;
      if (end_lane) sts_u8(endlist + (uint32_t)__popc(endv & ((1u << lane) - 1u)), (uint32_t)lane);
      tmem_st_wait();
      __syncwarp();

      // ---- 2a. U[j][k] = sum over the caption's words j' of G_c[j][j'] e[j'][k] on the warp's own tensor cores: the
      // quarter's 32x32 block-diagonal Gram (A, pre-arranged in fragment order and tf32-rounded by the Gram kernel) times
      // e (B, read straight from ES), mma.sync m16n8k8 tf32 -> 2 x 5 accumulator fragments (columns 36..39 are padding).
      // Within a k-step the contraction index is permuted (logical k = a -> word 8s+2a, k = a+4 -> word 8s+2a+1, the same
      // in A and B) so that the B reads are free of bank conflicts at the 36-float row pitch.
      const int fa = lane & 3, fg = lane >> 2;
      float y[2][5][4];
      {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 5; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) y[mt][nt][i] = 0.f;
        const uint32_t bbase = es + (uint32_t)(2 * fa) * ROW_BYTES + 4u * (uint32_t)fg;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          float b0[5], b1[5];
#pragma unroll
          for (int nt = 0; nt < 5; ++nt) {
            b0[nt] = lds_f1(bbase + (uint32_t)(8 * s * ROW_BYTES + 32 * nt));
            b1[nt] = lds_f1(bbase + (uint32_t)((8 * s + 1) * ROW_BYTES + 32 * nt));
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) mma_tf32(y[mt][nt], ga[mt * 4 + s], b0[nt], b1[nt]);
        }
        // y = e U at the accumulator positions: rows 16 mt + 8 h + fg, columns 8 nt + 2 fa (+1)
        const uint32_t ebase = es + (uint32_t)fg * ROW_BYTES + 8u * (uint32_t)fa;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) {
              const float2 e2 = lds_f2(ebase + (uint32_t)((16 * mt + 8 * h) * ROW_BYTES + 32 * nt));
              y[mt][nt][2 * h] *= e2.x; y[mt][nt][2 * h + 1] *= e2.y;
            }
      }

      // per-caption aggregate over the regions, held by the lane of the caption's last word
      float aggacc = agg_identity;
#pragma unroll
      for (int H = 0; H < 2; ++H) {
        // ---- 2b. walk scratch rows for eighteen regions: [t | y]; e is walked in place in ES -----------------------------
        {
          float t[18];
          TMEM_LD_X16(tpark + 18 * H, t, 0);
          TMEM_LD_X2(tpark + 18 * H + 16, t, 16);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int nt = 2 * H; nt < 2 * H + 3; ++nt) {
                const int rel = 8 * nt + 2 * fa - 18 * H;           // column pair within the half (even)
                if (rel >= 0 && rel < 18)
                  sts_f2(wk + (uint32_t)((16 * mt + 8 * h + fg) * ROW_BYTES + 72 + 4 * rel), y[mt][nt][2 * h], y[mt][nt][2 * h + 1]);
              }
          tmem_ld_wait();
          const uint32_t dst = wk + (uint32_t)lane * ROW_BYTES;
          sts_f4(dst, t[0], t[1], t[2], t[3]);
          sts_f4(dst + 16, t[4], t[5], t[6], t[7]);
          sts_f4(dst + 32, t[8], t[9], t[10], t[11]);
          sts_f4(dst + 48, t[12], t[13], t[14], t[15]);
          sts_f2(dst + 64, t[16], t[17]);
        }
        __syncwarp();
        // ---- 2c. row walk (as in the t2i kernel): 27 lanes, two columns each (18 on the scratch: P and Q, 9 on ES: Z),
        // restart at every caption end and leave the caption's totals in its last row -------------------------------------
        if (lane < 27) {
          const uint32_t col = lane < 18 ? wk + 8u * (uint32_t)lane : es + (uint32_t)(72 * H) + 8u * (uint32_t)(lane - 18);
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
        // ---- 2d. r_k = P_k / max(|v_k| sqrt(Q_k), 1e-8 Z_k) for every (caption end, region) pair, one pair per lane and
        // round; the value (already exponentiated for LSE) replaces P_k in the end row, whose lane then folds the eighteen
        // of them into its aggregate -----------------------------------------------------------------------------------
        for (int pi = lane; pi < npairs; pi += 32) {
          const int eo = (pi * 3641) >> 16;                        // pi / 18 for pi < 32 * 18
          const int k = pi - 18 * eo;
          const uint32_t rowoff = lds_u8(endlist + (uint32_t)eo) * ROW_BYTES + 4u * (uint32_t)k;
          const float P = lds_f1(wk + rowoff), Q = lds_f1(wk + rowoff + 72), Z = lds_f1(es + rowoff + (uint32_t)(72 * H));
          const float vn = __ldg(p.vnorm + (size_t)img * R + 18 * H + k);
          const float r = __fdividef(P, fmaxf(vn * sqrt_approx(fmaxf(Q, 0.f)), 1e-8f * Z));
          sts_f1(wk + rowoff, p.agg == ITR_AGG_LSE ? ex2f(r * p.c_lse) : r);
        }
        __syncwarp();
        if (end_lane) {
          const uint32_t tw = wk + (uint32_t)lane * ROW_BYTES;
          const float4 v0 = lds_f4(tw), v1 = lds_f4(tw + 16), v2 = lds_f4(tw + 32), v3 = lds_f4(tw + 48);
          const float2 v4 = lds_f2(tw + 64);
          if (p.agg == ITR_AGG_MAX) {
            aggacc = fmaxf(aggacc, fmaxf(fmaxf(fmaxf(fmaxf(v0.x, v0.y), fmaxf(v0.z, v0.w)), fmaxf(fmaxf(v1.x, v1.y), fmaxf(v1.z, v1.w))),
                                         fmaxf(fmaxf(fmaxf(v2.x, v2.y), fmaxf(v2.z, v2.w)), fmaxf(fmaxf(fmaxf(v3.x, v3.y), fmaxf(v3.z, v3.w)), fmaxf(v4.x, v4.y)))));
          } else {
            aggacc += (((v0.x + v0.y) + (v0.z + v0.w)) + ((v1.x + v1.y) + (v1.z + v1.w))) +
                      (((v2.x + v2.y) + (v2.z + v2.w)) + (((v3.x + v3.y) + (v3.z + v3.w)) + (v4.x + v4.y)));
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

// Word Gram of every packed word tile in the order the fused kernel's mma.sync A operand wants it.  Per 32-row quarter the
// Gram is block-diagonal (G[j][j'] = w_j . w_j' for two words of the same caption, else 0); out[tile][quarter][mt][s][lane]
// is the float4 (a0..a3) of the m16n8k8 fragment of rows 16 mt .. +15 and k-step s, with the kernel's permutation of the
// contraction index: a0 = G[16mt+g][8s+2a], a1 = G[16mt+g+8][8s+2a], a2 = G[16mt+g][8s+2a+1], a3 = G[16mt+g+8][8s+2a+1]
// (g = lane / 4, a = lane % 4), rounded to tf32.  One warp per word row; the row stays in registers (bf16 pairs).
__global__ void __launch_bounds__(256)
caption_gram_frag_kernel(const uint16_t* __restrict__ words, const int4* __restrict__ row_meta, int n_tiles, float* __restrict__ out) {
  const int tile = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile_out = out + (size_t)tile * (32 * BLOCK_M);
  for (int i = threadIdx.x; i < 32 * BLOCK_M / 4; i += 256) reinterpret_cast<float4*>(tile_out)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int r = warp; r < BLOCK_M; r += 8) {
    const size_t row = (size_t)tile * BLOCK_M + r;
    const int4 meta = row_meta[row];
    const int len = (meta.x >= 0 && !((meta.z >> 16) & 1)) ? meta.w : 0;       // long tiles are not scored by the fused kernel
    if (len == 0) continue;
    const int q = r >> 5, jl = r & 31, seg_lo = meta.z & 0xff;
    uint4 mine[4];                                                             // 1024 bf16 = 32 lanes x 4 x 8
    {
      const uint4* src = reinterpret_cast<const uint4*>(words + row * D);
#pragma unroll
      for (int v = 0; v < 4; ++v) mine[v] = src[lane + 32 * v];
    }
    for (int dlt = 0; dlt < len; ++dlt) {
      float s = 0.f;
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
      if (lane == 0) {
        const int jp = seg_lo + dlt;                                           // quarter-local index of the partner word
        const int mt = jl >> 4, g = jl & 7, hrow = (jl >> 3) & 1, ks = jp >> 3, a = (jp & 7) >> 1, odd = jp & 1;
        tile_out[((((q * 2 + mt) * 4 + ks) * 32) + 4 * g + a) * 4 + hrow + 2 * odd] = round_tf32(s);
      }
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

extern "C" int itr_scan_caption_gram_frag_bf16(const uint16_t* words_bf16, const int32_t* row_meta, int n_tiles, float* gq_frag,
                                              void* stream) {
  ITR_REQUIRE(words_bf16 && row_meta && gq_frag && n_tiles >= 0, "itr_scan_caption_gram_frag_bf16: bad arguments");
  ITR_REQUIRE(((uintptr_t)words_bf16 & 15) == 0 && ((uintptr_t)row_meta & 15) == 0, "itr_scan_caption_gram_frag_bf16: buffers must be 16-byte aligned");
  if (n_tiles == 0) return ITR_OK;
  tc2i::caption_gram_frag_kernel<<<n_tiles, 256, 0, as_stream(stream)>>>(words_bf16, reinterpret_cast<const int4*>(row_meta), n_tiles, gq_frag);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_scan_i2t_scores_bf16(const uint16_t* images_bf16, const float* region_norm, int n_img,
                                        const uint16_t* words_bf16, const int32_t* row_meta, const float* gq_frag, int n_tiles,
                                        int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                                        float* scores, int64_t ld_scores, void* stream) {
  ITR_REQUIRE(images_bf16 && region_norm && words_bf16 && row_meta && gq_frag && scores, "itr_scan_i2t_scores_bf16: null pointer");
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
  p.gq_frag = gq_frag; p.vnorm = region_norm;
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
