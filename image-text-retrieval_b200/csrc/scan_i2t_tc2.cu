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
// Where t2i reduces over regions (in-thread, TMEM lane = word), i2t needs three sums over a caption's WORDS per region
// (Z, P, Q) plus the contraction with G_c.  A first version walked shared-memory rows for them and was bound by the
// shared-memory pipe (1085 wavefronts per warp and item, 93 % busy together with the MMA operand reads); this one moves
// every contraction onto the warp's own tensor cores (mma.sync m16n8k8 tf32, fp32 accumulation) on TRANSPOSED tiles (rows =
// regions, columns = words), so that a sum over words is a contraction over the accumulator's column index and chains
// from registers.  Per (warp = 32 word rows, image), with two 4.6 KB shared-memory scratches:
//   1. lane = word: e (rounded to tf32, the same value everywhere) -> scratch ES, t = e A -> scratch TS (fp32, [word][36]);
//   2. U^T = e^T G (A = e^T read from ES in fragment order, B = the quarter's block-diagonal Gram, eight coalesced 16-byte
//      loads per lane), y^T = e^T o U^T in registers;
//   3. Z^T = e^T S, Q^T = y^T S, P^T = t^T S with S[word][caption] = [the word belongs to the caption], built from the
//      end mask in registers; y and t enter split hi + lo, so the sums keep fp32 accuracy;
//   4. r_k at the accumulator positions (region, caption), aggregated over the regions in-thread and with three shuffles;
//      the lane of each caption's last word fetches and stores the score.
// Shared-memory traffic: 72 wavefronts written + 80 read per warp and item.
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
// word-tile pairs per work unit: an evaluation fold is small (206 pairs x 250 image tiles over 74 CTA pairs), so units are
// kept short for the static schedule's balance (measured: 8 -> 7.83 ms, 32 -> 8.08 ms, 64 -> 8.18 ms on one fold)
constexpr int I2T_BAND = 8;
constexpr int AUX_BYTES = BLOCK_M * 16;            // row metadata of the CTA's word tile
constexpr int ROW_BYTES = 36 * 4;                  // one scratch row: 36 floats
constexpr int WARP_SCRATCH = 32 * ROW_BYTES;       // 4608
constexpr int SMEM_STAGES = 0;
constexpr int SMEM_AUX = SMEM_STAGES + STAGES * STAGE_BYTES;
constexpr int SMEM_ES = SMEM_AUX + 2 * AUX_BYTES;
constexpr int SMEM_TS = SMEM_ES + NUM_EPI_WARPS * WARP_SCRATCH;
constexpr int SMEM_BARS = SMEM_TS + NUM_EPI_WARPS * WARP_SCRATCH;
constexpr int NUM_BARS = 2 * STAGES + 8;
constexpr int SMEM_TMEMPTR = SMEM_BARS + NUM_BARS * 8;
constexpr int SMEM_BYTES = SMEM_TMEMPTR + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;
static_assert(SMEM_ALLOC <= 232448, "shared memory budget");
static_assert(SMEM_AUX % 16 == 0 && SMEM_ES % 16 == 0 && SMEM_TS % 16 == 0 && SMEM_BARS % 8 == 0, "alignment");

struct Params {
  const int4* row_meta;        // [n_wt*128]
  const float* gq_frag;        // [n_wt][4][2][4][32] float4: word Gram in mma.sync fragment order
  const float* vnorm;          // [n_img][36] region norms
  int n_img, n_wt, n_wp, n_it;
  int agg;
  float c_sm, c_lse, inv_lse;
  float* scores; long long ld;
  int band;
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
// Round-to-nearest (ties away in magnitude) to tf32's 10-bit mantissa with two integer instructions; cvt.rna.tf32.f32
// expands to a NaN/Inf-aware sequence of ~6.  The values seen here are finite.
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// hi part of a hi + lo split: the tensor core ignores the low 13 bits of a tf32 operand, so plain truncation is the hi
// the hardware would see anyway; lo = x - hi is exact in fp32
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
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
  const Schedule sched(p.n_wp, p.n_it, nullptr, 0, p.band);
  const int first = (int)(blockIdx.x >> 1);
  const int step = (int)(gridDim.x >> 1);

  if (warp < EPI_WARP0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
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
    const uint32_t ts = sbase + SMEM_TS + (uint32_t)(warp - EPI_WARP0) * WARP_SCRATCH;
    const int fa = lane & 3, fg = lane >> 2;                                   // mma.sync fragment coordinates
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
      const int seg_hi = (meta.z >> 8) & 0xff;
      // The buffer may be refilled only once every lane HOLDS its metadata: an arrive issued while the loads are still in
      // flight lets the refill overtake them (seen as rare stale rows).  The vote consumes every lane's value, and the
      // arrive depends on its result (lane 31 always ends a caption or is padding, so the mask is never 0).
      const uint32_t endmask = __ballot_sync(0xffffffffu, lane == seg_hi);
      __syncwarp();                                            // orders every lane's read before lane 0's release
      if (lane == 0 && endmask != 0u) mbar_arrive(aempty_bar(b));
      const bool long_tile = (meta.z >> 16) & 1;
      const int img = n * IMGS + g;
      const bool end_lane = lane == seg_hi && meta.x >= 0;
      const uint32_t endv = __ballot_sync(0xffffffffu, end_lane);                   // last words of the real captions
      const uint32_t realv = __ballot_sync(0xffffffffu, meta.x >= 0);              // real (non-padding) word rows
      const bool valid = img < p.n_img && word_ok && !long_tile && endv != 0u;      // warp-uniform

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

      // ---- 1. per-word l2norm over the regions, softmax numerators e (rounded to tf32) and t = e A, lane = word ---------
      float n2 = 0.f;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const float a = CLIPPED ? fmaxf(A[k], 0.1f * A[k]) : A[k];
        n2 = fmaf(a, a, n2);
      }
      const float cw = __fdividef(p.c_sm, sqrtf(n2) + 1e-8f);
      const uint32_t erow = es + (uint32_t)lane * ROW_BYTES, trow = ts + (uint32_t)lane * ROW_BYTES;
#pragma unroll
      for (int k = 0; k < R; k += 4) {
        float e[4], t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float raw = A[k + i];
          const float a = CLIPPED ? fmaxf(raw, 0.1f * raw) : raw;
          e[i] = round_tf32(ex2f(fmaf(a, cw, shift)));      // the tensor-core contractions below read e as tf32: the same e everywhere
          t[i] = e[i] * raw;
        }
        sts_f4(erow + 4 * k, e[0], e[1], e[2], e[3]);
        sts_f4(trow + 4 * k, t[0], t[1], t[2], t[3]);
      }
      // the quarter's Gram in B-fragment order (step 2): b0 = G[8s+2a][8nt+g], b1 = G[8s+2a+1][8nt+g], two n-tiles per float4;
      // streamed from L1/L2 once per row tile (holding it would cost 32 registers)
      const float4* gf = reinterpret_cast<const float4*>(p.gq_frag) + ((size_t)(m * 4 + q) * 8) * 32 + lane;
      const float* vp = p.vnorm + (size_t)img * R + fg;          // |v_k| of this lane's accumulator rows: regions 16 mt + 8 h + g
      const int my_ord = __popc(endv & ((1u << lane) - 1u));      // meaningful on the lanes that end a caption
      const int n_ct = (__popc(endv) + 7) >> 3;                  // caption n-tiles of eight: one unless the captions are tiny
      __syncwarp();

      // Everything below works on TRANSPOSED tiles (rows = regions, padded 36 -> 48 = three m16 tiles; columns = the 32
      // words), mma.sync m16n8k8 tf32 with fp32 accumulation, so that sums over a caption's words are contractions over
      // the accumulator's COLUMN index and chain from registers.  Contraction index of k-step s: logical k = a -> word
      // 8s+2a, k = a+4 -> word 8s+2a+1 (the same permutation in every operand): with it the A fragment of k-step s sits at
      // exactly the accumulator positions of word n-tile s, and the fragment reads are free of bank conflicts at the
      // 36-float row pitch.
      //   2. U^T = e^T G            A = e^T from ES, B = the quarter's block-diagonal Gram (registers)
      //      y^T = e^T o U^T        element-wise, in registers
      //   3. Z^T = e^T S, Q^T = y^T S, P^T = t^T S     S[word][caption] = 1 if the word belongs to the caption (built from
      //      the end mask in registers); y and t are split hi + lo so that the sums keep fp32 accuracy
      //   4. r = P / max(|v| sqrt(Q), 1e-8 Z) at the accumulator positions (region, caption); aggregate over the regions:
      //      in-thread over the row tiles, three shuffles over g; the lane that ends the caption fetches and stores it.
#pragma unroll 1
      for (int ct = 0; ct < n_ct; ++ct) {
        float sb[4][2];
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            // caption ordinal (within the quarter) of word 8 s + 2 a + o; padding rows belong to no caption
            const int w = 8 * s + 2 * fa + o;
            const bool mine = ((realv >> w) & 1u) && __popc(endv & ((1u << w) - 1u)) == 8 * ct + fg;
            sb[s][o] = mine ? 1.f : 0.f;
          }
        float agg[2] = {agg_identity, agg_identity};             // captions 8 ct + 2 a (+1), this lane's regions
#pragma unroll 1
        for (int mt = 0; mt < 3; ++mt) {
          // A fragments of e^T: a0 = e[8s+2a][16mt+g], a1 = e[8s+2a][16mt+g+8], a2, a3 = the same of word 8s+2a+1; rows 36..47 are
          // padding (a1 = a3 = 0 in the last tile; its a0/a2 of g >= 4 re-read region 35 and feed rows nobody uses)
          const uint32_t fo = (uint32_t)(2 * fa) * ROW_BYTES + 4u * (uint32_t)min(16 * mt + fg, R - 1);
          float4 ea[4];
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const uint32_t ad = es + fo + (uint32_t)(8 * s * ROW_BYTES);
            ea[s].x = lds_f1(ad); ea[s].z = lds_f1(ad + ROW_BYTES);
            ea[s].y = mt < 2 ? lds_f1(ad + 32) : 0.f; ea[s].w = mt < 2 ? lds_f1(ad + ROW_BYTES + 32) : 0.f;
          }
          float u[4][4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) u[nt][i] = 0.f;
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const float4 g01 = __ldg(gf + (2 * s) * 32), g23 = __ldg(gf + (2 * s + 1) * 32);
            // (the all-zero 8x8 blocks of the block-diagonal Gram are not skipped: a branch per MMA cost more than it saved)
            mma_tf32(u[0], ea[s], g01.x, g01.y);
            mma_tf32(u[1], ea[s], g01.z, g01.w);
            mma_tf32(u[2], ea[s], g23.x, g23.y);
            mma_tf32(u[3], ea[s], g23.z, g23.w);
          }
          // hi and lo parts accumulate separately: independent chains of four MMAs instead of one of eight
          float zq[4] = {0.f, 0.f, 0.f, 0.f}, qq[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};
          float ql[4] = {0.f, 0.f, 0.f, 0.f}, pl[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            mma_tf32(zq, ea[s], sb[s][0], sb[s][1]);
            // y^T at the accumulator positions of word tile s, re-ordered into an A fragment: (c0, c2, c1, c3)
            float4 y = make_float4(u[s][0] * ea[s].x, u[s][2] * ea[s].y, u[s][1] * ea[s].z, u[s][3] * ea[s].w);
            float4 yh = make_float4(trunc_tf32(y.x), trunc_tf32(y.y), trunc_tf32(y.z), trunc_tf32(y.w));
            mma_tf32(qq, yh, sb[s][0], sb[s][1]);
            mma_tf32(ql, make_float4(y.x - yh.x, y.y - yh.y, y.z - yh.z, y.w - yh.w), sb[s][0], sb[s][1]);
          }
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const uint32_t ad = ts + fo + (uint32_t)(8 * s * ROW_BYTES);
            float4 t4;
            t4.x = lds_f1(ad); t4.z = lds_f1(ad + ROW_BYTES);
            t4.y = mt < 2 ? lds_f1(ad + 32) : 0.f; t4.w = mt < 2 ? lds_f1(ad + ROW_BYTES + 32) : 0.f;
            float4 th = make_float4(trunc_tf32(t4.x), trunc_tf32(t4.y), trunc_tf32(t4.z), trunc_tf32(t4.w));
            mma_tf32(pq, th, sb[s][0], sb[s][1]);
            mma_tf32(pl, make_float4(t4.x - th.x, t4.y - th.y, t4.z - th.z, t4.w - th.w), sb[s][0], sb[s][1]);
          }
          // accumulator positions: [2h + cc] = (region 16 mt + 8 h + g, caption 8 ct + 2 a + cc)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const bool region_ok = 16 * mt + 8 * h + fg < R;                  // regions 36..47: padding
            const float vk = region_ok ? __ldg(vp + 16 * mt + 8 * h) : 1.f;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const float P = pq[2 * h + cc] + pl[2 * h + cc], Q = qq[2 * h + cc] + ql[2 * h + cc], Z = zq[2 * h + cc];
              const float r = __fdividef(P, fmaxf(vk * sqrt_approx(fmaxf(Q, 0.f)), 1e-8f * Z));
              const float val = p.agg == ITR_AGG_LSE ? ex2f(r * p.c_lse) : r;
              if (p.agg == ITR_AGG_MAX) agg[cc] = region_ok ? fmaxf(agg[cc], val) : agg[cc];
              else agg[cc] += region_ok ? val : 0.f;
            }
          }
        }
        // over the regions held by the other seven g of the same a
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
          for (int sh = 4; sh < 32; sh <<= 1) {
            const float o = __shfl_xor_sync(0xffffffffu, agg[cc], sh);
            agg[cc] = p.agg == ITR_AGG_MAX ? fmaxf(agg[cc], o) : agg[cc] + o;
          }
        // ---- 5. the lane of each caption's last word fetches its caption's aggregate (lane a = (ordinal % 8) / 2) and stores it
        const int src = (my_ord & 7) >> 1;
        const float v0 = __shfl_sync(0xffffffffu, agg[0], src), v1 = __shfl_sync(0xffffffffu, agg[1], src);
        if (end_lane && (my_ord >> 3) == ct) {
          float tot = (my_ord & 1) ? v1 : v0;
          if (p.agg == ITR_AGG_LSE) tot = lg2f(tot) * p.inv_lse;
          if (p.agg == ITR_AGG_MEAN) tot = tot * (1.0f / (float)R);
          p.scores[(size_t)img * p.ld + meta.x] = tot;
        }
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

// Word Gram of every packed word tile in the order the fused kernel's mma.sync B operand wants it.  Per 32-row quarter the
// Gram is block-diagonal (G[j][j'] = w_j . w_j' for two words of the same caption, else 0) and symmetric;
// out[tile][quarter][s][p][lane] is the float4 (b0, b1 of word n-tile 2p, b0, b1 of n-tile 2p+1) of k-step s with the kernel's
// permutation of the contraction index: b0 = G[8s+2a][8nt+g], b1 = G[8s+2a+1][8nt+g] (g = lane / 4, a = lane % 4), rounded
// to tf32.  One warp per word row; the row stays in registers (bf16 pairs).
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
        const int ks = jp >> 3, a = (jp & 7) >> 1, odd = jp & 1, nt = jl >> 3, g = jl & 7;
        tile_out[((((q * 4 + ks) * 2 + (nt >> 1)) * 32) + 4 * g + a) * 4 + 2 * (nt & 1) + odd] = round_tf32(s);
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
  const long long units = (long long)((p.n_wp + p.band - 1) / p.band) * p.n_it;
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
  p.band = tc2i::I2T_BAND;
  if ((long long)p.n_wp * p.n_it >= (1ll << 31))
    return fail(ITR_ERR_INVALID, "itr_scan_i2t_scores_bf16: %lld tile pairs exceed the 2^31 scheduler range; split the call", (long long)p.n_wp * p.n_it);
  cudaStream_t st = as_stream(stream);
  return feature_norm == ITR_NORM_CLIPPED_L2 ? tc2i::launch<true>(map_w, map_i, p, st) : tc2i::launch<false>(map_w, map_i, p, st);
}
