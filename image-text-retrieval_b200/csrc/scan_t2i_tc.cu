// SCAN text-to-image cross-attention scores on Blackwell tensor cores (sm_100a only).
//
// Replaces xattn_score_t2i + func_attention + cosine_similarity
// (itr/modalmodule/Objectives.py:329-372, 421-476, 10-15) for
// raw_feature_norm in {clipped_l2norm, l2norm} and every agg_func.
//
// Formulation: for caption c (words w_j, j < n_c) and image i
// (unit regions v_k, k < 36)
//     A[k][j] = v_k . w_j                    <- the only D-wide contraction (tensor cores)
//     a       = leaky_0.1(A)                 (clipped_l2norm; identity for l2norm)
//     ahat    = a / (sqrt(sum_j a[k][j]^2) + 1e-8)          per region, over the caption's words
//     alpha_j = softmax_k(lambda_sm * ahat[k][j])            per word, over regions
//     r_j     = (sum_k alpha_jk A[k][j]) / max(|w_j| sqrt(alpha_j^T G_i alpha_j), 1e-8)
//     S[i][c] = agg_j r_j                                     LSE / Mean / Max / Sum
// The reference's second batched matmul (attended context, K = 36 -> D = 1024) is removed by
// the identities  w_j . ctx_j = sum_k alpha_jk A[k][j]  and  |ctx_j|^2 = alpha_j^T G_i alpha_j
// with G_i = V_i V_i^T the 36x36 region Gram, precomputed once per image.
//
// Kernel structure: persistent, warp-specialised, one CTA per SM, 20 warps.
//   tile       = 128 packed words (UMMA M) x 4 images = 144 region columns (UMMA N), K = 1024
//   warp 0     TMA producer: 5-stage ring of {words 128x64, regions 144x64} bf16 tiles, SWIZZLE_128B
//   warp 1     main tcgen05.mma issuer (cta_group::1, kind::f16, bf16 x bf16 -> fp32 in TMEM)
//   warp 2     TMEM allocator + issuer of the small "Gram" MMAs (see below)
//   warp 3     aux loader: per-tile Gram packs, row metadata, word norms (bulk copies)
//   warps 4-19 epilogue, four groups of four warps; group g owns image g of the tile and all 128 word
//              rows (thread = one word row = one TMEM lane).  Two accumulators alternate between items.  Per item:
//                load    tcgen05.ld of the 36 raw affinities of item t is issued first, and while it is in flight
//                B(t-1)  finishes the PREVIOUS item: tcgen05.ld U = e (G - I) (its Gram MMA completed long ago),
//                        |ctx|^2 Z^2 = D + sum_k e_k U_k with e(t-1) still in registers, cosine, aggregation over
//                        the caption's words, store;
//                A(t)    leaky, l2norm over the caption's words (segmented warp scan),
//                        e_k = exp2(lambda*ahat_k - lambda), P = sum e A, D = sum e^2;
//                park(t) tcgen05.st e as fp16 into 24 of the group's OWN 36 columns of the accumulator item t came
//                        from (its values are in registers now) and wake the Gram issuer, which runs U = e (G - I)
//                        as a 128x48x48 tcgen05.mma with A FROM TMEM and the image's fp16 Gram pack as the SMEM
//                        B operand.  The MMA issuer reuses that accumulator for item t+2 once the Gram MMAs have
//                        completed, so neither the accumulator hand-off nor the Gram latency is on the item cycle.
//              The two cross-row reductions (l2norm over the caption's words, aggregation over words) are
//              segmented warp scans -- itr_scan_plan_words guarantees a caption never straddles a warp,
//              except in `long` tiles which exchange through shared memory.
// History and measurements behind these choices: DESIGN.md section 5, profiles/r01/.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace itr {
namespace tc {

constexpr int R = ITR_REGIONS;                 // 36
constexpr int IMGS = ITR_TILE_IMAGES;          // 4
constexpr int BLOCK_M = ITR_TILE_WORDS;        // 128
constexpr int BLOCK_N = IMGS * R;              // 144
constexpr int BLOCK_K = 64;                    // bf16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int D = ITR_EMBED;                   // 1024
constexpr int K_BLOCKS = D / BLOCK_K;          // 16
constexpr int STAGES = 5;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2; // 16384
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2; // 18432
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
// per-image Gram pack: fp16 48x48 off-diagonal Gram in the canonical no-swizzle K-major UMMA layout
// (8x8 core matrices of 128 bytes; element (n,k) at (n/8)*768 + (k/8)*128 + (n%8)*16 + (k%8)*2),
// followed by the 36 fp32 diagonal entries.
constexpr int GRAM_N = 48;
constexpr int G16_BYTES = GRAM_N * GRAM_N * 2;     // 4608
constexpr int GRAM_BYTES = ITR_GRAM_BYTES;         // 4752 = 4608 + 36*4
constexpr int G_LBO = 128, G_SBO = 768;
constexpr int AUX_GRAM = IMGS * GRAM_BYTES;        // 19008
constexpr int AUX_META = BLOCK_M * 16;             // 2048
constexpr int AUX_WNORM = BLOCK_M * 4;             // 512
constexpr int AUX_BYTES = AUX_GRAM + AUX_META + AUX_WNORM;   // 21568
constexpr int XCH_FLOATS = 4 /*group*/ * 4 /*warp*/ * 40;
// Tensor memory (512 columns): two accumulators [0,144) and [144,288), four Gram products at 288 + 48 g.
// The fp16 numerators e(t) of group g are parked INSIDE the accumulator the item came from, in 24 of the group's own
// 36 columns (the group has its raw affinities in registers by then), so a second accumulator fits.
constexpr int ACC_PITCH = BLOCK_N;               // 144
constexpr int U_BASE = 2 * ACC_PITCH;            // 288
__host__ __device__ constexpr int park_col(int g) { return g == 0 ? 0 : 16 * ((36 * g + 15) / 16); }   // 0, 48, 80, 112
static_assert(park_col(1) >= 36 && park_col(1) + 24 <= 72 && park_col(2) >= 72 && park_col(2) + 24 <= 108 &&
              park_col(3) >= 108 && park_col(3) + 24 <= 144, "parking area must stay inside the group's own columns");
constexpr int TMEM_COLS = 512;
constexpr int BAND = 64;                       // word tiles kept L2-resident while images stream
constexpr int NUM_THREADS = 640;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 16;

constexpr int SMEM_STAGES = 0;
constexpr int SMEM_AUX = SMEM_STAGES + STAGES * STAGE_BYTES;
constexpr int SMEM_XCH = SMEM_AUX + 2 * AUX_BYTES;
constexpr int SMEM_BARS = SMEM_XCH + XCH_FLOATS * 4;
constexpr int NUM_BARS = 2 * STAGES + 18;
constexpr int SMEM_TMEMPTR = SMEM_BARS + NUM_BARS * 8;
constexpr int SMEM_BYTES = SMEM_TMEMPTR + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;   // slack for manual 1024-byte alignment

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=144
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
// Gram MMA: D=f32, A=B=fp16, K-major, M=128, N=48
constexpr uint32_t IDESC_GRAM = (1u << 4) | ((uint32_t)(GRAM_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);

// ---------------------------------------------------------------------------- tile schedule
// Work unit = (band of BAND consecutive word tiles, image tile n); a CTA takes units u = cta, cta + grid, ...
// and walks the band's word tiles against the SAME image tile.  So an image tile comes from HBM once per band
// and is then re-read from L2 by the one CTA that owns it, while the band of word tiles (8 MB) stays L2-resident
// and is shared by all CTAs.  (The first schedule let 32 CTAs request each image tile at the same instant: ncu
// showed 159 GB of DRAM reads per COCO-5K launch against ~30 GB expected -- concurrent misses are not merged.)
struct Schedule {
  int n_wt, n_it, n_bands, last_band;
  __device__ Schedule(int n_wt_, int n_it_) : n_wt(n_wt_), n_it(n_it_) {
    n_bands = (n_wt + BAND - 1) / BAND;
    last_band = n_wt - (n_bands - 1) * BAND;
  }
  __device__ int units() const { return n_bands * n_it; }   // host guarantees n_wt * n_it < 2^31
};

// Walks the items of one CTA: units u = first, first + step, ...; inside a unit the word tile advances.
struct ItemIter {
  const Schedule& s;
  int u, step, m, n, left;
  __device__ ItemIter(const Schedule& s_, int first, int step_) : s(s_), u(first), step(step_) { open(); }
  __device__ void open() {
    if (u < s.units()) {
      const int band = u / s.n_it;
      n = u - band * s.n_it;
      m = band * BAND;
      left = (band == s.n_bands - 1) ? s.last_band : BAND;
    }
  }
  __device__ bool valid() const { return u < s.units(); }
  __device__ void next() {
    ++m;
    if (--left == 0) { u += step; open(); }
  }
};

struct Params {
  const uint8_t* gram_pack;    // [n_img][GRAM_BYTES]
  const int4* row_meta;        // [n_wt*128]
  const float* row_wnorm;      // [n_wt*128]
  int n_img, n_wt, n_it;
  int clipped, agg;
  float c_sm;                  // lambda_softmax * log2(e)
  float c_lse;                 // lambda_lse * log2(e)
  float inv_lse;               // ln(2) / lambda_lse
  float* scores; long long ld;
  float* dump;                 // debug: raw affinities of item (dbg_m, dbg_n)
  int dbg_m, dbg_n;
  long long* prof;             // optional [grid][16] cycle counters (see itr_scan_t2i_profile)
  int ctrl_last;               // 1: the control warpgroup is the LAST one (highest warp ids), 0: the first
  int skip_math;               // tuning only: epilogue loads and releases the accumulator, no arithmetic
  int wait_mode;               // tuning only (PROF builds): 1 producer spins on `empty`, 2 MMA issuer spins on `full`
};

// inclusive segmented scan over the lanes [seg_lo, lane], then broadcast of the segment total
template <bool MAXOP>
__device__ __forceinline__ float seg_total(float x, const bool (&p)[5], int seg_hi) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    float y = __shfl_up_sync(0xffffffffu, x, 1 << s);
    if (p[s]) x = MAXOP ? fmaxf(x, y) : x + y;
  }
  return __shfl_sync(0xffffffffu, x, seg_hi);
}

// what phase B of an item needs from its phase A (phase B runs one item later, see below)
struct Carry {
  float P, D, wnorm;
  int cap, seg, n_words, img, b;
  bool valid, live;
};

// DEBUG: one (word tile, image tile) pair, raw accumulator to p.dump[128][144] (bring-up).
// DUMP : the whole problem, raw affinities to p.dump[word tile][image][row][36] -- phase 1 of the generic
//        two-phase path (itr_scan_affinity_bf16 + itr_scan_epilogue_f32) used for i2t and the norm modes the
//        fused epilogue does not implement.  No Gram MMA, no aux loads, the epilogue is a TMEM -> HBM copy.
template <bool DEBUG, bool PROF, bool DUMP = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
scan_t2i_tc_kernel(const __grid_constant__ CUtensorMap map_words, const __grid_constant__ CUtensorMap map_imgs, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  // logical warp index: 0-3 control, 4-19 epilogue.  The shift is a multiple of 4, so the TMEM lane
  // quarter (physical warp id % 4) and warpgroup alignment (setmaxnreg) are preserved.
  const int warp = (int)((threadIdx.x >> 5) + (p.ctrl_last ? EPI_WARP0 : 0)) % (NUM_THREADS / 32);
  const int lane = threadIdx.x & 31;

  const uint32_t bar0 = sbase + SMEM_BARS;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 0 + b); };     // accumulator b complete
  auto loaded_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 + b); };    // all 16 epilogue warps hold accumulator b in registers
  auto gfree_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 16 + b); };    // the Gram MMAs that read the numerators parked in b are done
  auto afull_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 4 + b); };
  auto aempty_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 6 + b); };
  auto eready_bar = [&](int g) { return bar0 + 8u * (2 * STAGES + 8 + g); };
  auto uready_bar = [&](int g) { return bar0 + 8u * (2 * STAGES + 12 + g); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + SMEM_TMEMPTR);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_words) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_imgs) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1); mbar_init(loaded_bar(b), NUM_EPI_WARPS); mbar_init(gfree_bar(b), 1);
      mbar_init(afull_bar(b), 1); mbar_init(aempty_bar(b), NUM_EPI_WARPS);
    }
    for (int g = 0; g < IMGS; ++g) { mbar_init(eready_bar(g), 4); mbar_init(uready_bar(g), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SMEM_TMEMPTR), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The CTA owns the SM (1 CTA/SM, all 512 columns allocated), so the allocation starts at column 0, lane 0.
  // Treating the base as a compile-time constant keeps it out of the register file of the issuer warps.
  if (*tmem_ptr_smem != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;

  constexpr bool prof_on = PROF;         // wait-cycle counters are compiled out of the production kernel
  const Schedule sched(p.n_wt, p.n_it);
  const int first = DEBUG ? 0 : (int)blockIdx.x;
  const int step = DEBUG ? 1 : (int)gridDim.x;

  // Register budget: the control warpgroup (warps 0-3) gives registers back, the four epilogue
  // warpgroups take them (128*56 + 512*104 <= 64K).
  if (warp < EPI_WARP0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // =============================== TMA producer: operand ring ============================
    int stage = 0; uint32_t phase = 0;
    long long w_empty = 0; const long long t_begin = prof_on ? clock64() : 0;
    int dbg_it = 0;
    for (ItemIter item(sched, first, step); DEBUG ? dbg_it == 0 : item.valid(); item.next(), ++dbg_it) {
      const int row_w = (DEBUG ? p.dbg_m : item.m) * BLOCK_M, row_i = (DEBUG ? p.dbg_n : item.n) * BLOCK_N;
#pragma unroll 1
      for (int kb = 0; kb < K_BLOCKS; ++kb) {
        if (PROF && (p.wait_mode & 1)) mbar_wait_t(empty_bar(stage), phase ^ 1, w_empty, prof_on);
        else mbar_wait_sleep_t(empty_bar(stage), phase ^ 1, w_empty, prof_on);
        const uint32_t sa = sbase + SMEM_STAGES + stage * STAGE_BYTES, fb = full_bar(stage);
        if (elect_one()) {
          if (PROF && (p.skip_math & 2)) {       // tuning only: no operand traffic at all (stale SMEM)
            mbar_arrive(fb);
          } else {
            mbar_expect_tx(fb, STAGE_BYTES);
            tma_load_2d(sa, &map_words, fb, kb * BLOCK_K, row_w);
            tma_load_2d(sa + A_BYTES, &map_imgs, fb, kb * BLOCK_K, row_i);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (prof_on && lane == 0) { p.prof[blockIdx.x * 16 + 0] = clock64() - t_begin; p.prof[blockIdx.x * 16 + 1] = w_empty; }
  } else if (warp == 1) {
    // =============================== main MMA issuer ======================================
    // This warp shares its scheduler with four busy epilogue warps (measured: it is starved of issue
    // slots, not the tensor pipe of work), so its per-k-block instruction count is kept minimal:
    // descriptors are one 64-bit add away from a precomputed base.
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    long long w_tempty = 0, w_full = 0; const long long t_begin = prof_on ? clock64() : 0;
    const uint64_t adesc0 = umma_desc_sw128(sbase + SMEM_STAGES);
    const uint64_t bdesc0 = umma_desc_sw128(sbase + SMEM_STAGES + A_BYTES);
    for (ItemIter item(sched, first, step); DEBUG ? it == 0 : item.valid(); item.next(), ++it) {
      // accumulator b = it & 1 is reusable once (a) every epilogue warp has item it-2 in registers and (b) the Gram MMAs
      // of item it-2, which read the numerators parked in it, have completed.
      const int ab = it & 1;
      const uint32_t tacc = tmem_base + ab * ACC_PITCH;
      mbar_wait_sleep_t(loaded_bar(ab), ((it >> 1) & 1) ^ 1, w_tempty, prof_on);
      if (!DEBUG && !DUMP) mbar_wait_sleep_t(gfree_bar(ab), ((it >> 1) & 1) ^ 1, w_tempty, prof_on);
      tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < K_BLOCKS; ++kb) {
        if (PROF && (p.wait_mode & 2)) mbar_wait_t(full_bar(stage), phase, w_full, prof_on);
        else mbar_wait_sleep_t(full_bar(stage), phase, w_full, prof_on);
        tc_fence_after();
        const uint64_t soff = (uint64_t)((uint32_t)stage * (uint32_t)(STAGE_BYTES >> 4));
        const uint64_t adesc = adesc0 + soff, bdesc = bdesc0 + soff;
        if (elect_one()) {
          umma_bf16(tacc, adesc, bdesc, IDESC, (uint32_t)kb);                    // first MMA of a tile overwrites
          umma_bf16(tacc, adesc + 2, bdesc + 2, IDESC, 1u);                      // +32 bytes per K step
          umma_bf16(tacc, adesc + 4, bdesc + 4, IDESC, 1u);
          umma_bf16(tacc, adesc + 6, bdesc + 6, IDESC, 1u);
          umma_commit(empty_bar(stage));
          if (kb == K_BLOCKS - 1) umma_commit(tfull_bar(ab));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (prof_on && lane == 0) {
      p.prof[blockIdx.x * 16 + 2] = clock64() - t_begin; p.prof[blockIdx.x * 16 + 3] = w_tempty; p.prof[blockIdx.x * 16 + 4] = w_full;
      p.prof[blockIdx.x * 16 + 5] = it;
    }
  } else if (warp == 2) {
    // =============================== Gram MMA issuer =======================================
    // U_g = e_g (G_g - I): A = the group's parked fp16 numerators (TMEM), B = the image's Gram pack (SMEM)
    if (!DEBUG && !DUMP) {
      int it = 0;
      uint32_t used[IMGS] = {0u, 0u, 0u, 0u};    // completed phases of eready[g] (tail images are skipped)
      for (ItemIter item(sched, first, step); item.valid(); item.next(), ++it) {
        const int n = item.n;
        const int b = it & 1;
        mbar_wait_sleep(afull_bar(b), (it >> 1) & 1);
        const uint32_t aux = sbase + SMEM_AUX + b * AUX_BYTES;
#pragma unroll
        for (int g = 0; g < IMGS; ++g) {
          if (n * IMGS + g >= p.n_img || (PROF && (p.skip_math & 1))) continue;
          mbar_wait_sleep(eready_bar(g), used[g]++ & 1);
          tc_fence_after();
          const uint32_t te = tmem_base + b * ACC_PITCH + park_col(g);
          const uint32_t tu = tmem_base + U_BASE + g * GRAM_N;
          const uint64_t gdesc = umma_desc_nosw(aux + g * GRAM_BYTES, G_LBO, G_SBO);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GRAM_N / UMMA_K; ++k)          // 16 fp16 = 8 TMEM columns, 2 core matrices
              umma_f16_ts(tu, te + 8 * k, gdesc + (uint64_t)((2 * G_LBO * k) >> 4), IDESC_GRAM, k != 0);
            umma_commit(uready_bar(g));
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(gfree_bar(b));      // every Gram MMA of this item done -> accumulator b may be overwritten
        __syncwarp();
      }
    }
  } else {
    // =============================== aux loader: Gram packs, row metadata, word norms ======
    int it = 0;
    if (!DUMP)
    for (ItemIter item(sched, first, step); DEBUG ? it == 0 : item.valid(); item.next(), ++it) {
      const int m = DEBUG ? p.dbg_m : item.m, n = DEBUG ? p.dbg_n : item.n;
      const int b = it & 1;
      mbar_wait_sleep(aempty_bar(b), ((it >> 1) & 1) ^ 1);
      if (elect_one()) {
        const int n_valid = min(IMGS, p.n_img - n * IMGS);
        const uint32_t aux = sbase + SMEM_AUX + b * AUX_BYTES;
        mbar_expect_tx(afull_bar(b), n_valid * GRAM_BYTES + AUX_META + AUX_WNORM);
        bulk_load(aux, p.gram_pack + (size_t)n * IMGS * GRAM_BYTES, n_valid * GRAM_BYTES, afull_bar(b));
        bulk_load(aux + AUX_GRAM, p.row_meta + (size_t)m * BLOCK_M, AUX_META, afull_bar(b));
        bulk_load(aux + AUX_GRAM + AUX_META, p.row_wnorm + (size_t)m * BLOCK_M, AUX_WNORM, afull_bar(b));
      }
      __syncwarp();
    }
  }
  } else {
    // =============================== epilogue =============================================
    // Group g (4 warps = all 128 word rows) owns image g of every tile.  Per item:
    //   load   issue the TMEM loads of the 36 raw affinities of item t
    //   B(t-1) finish the PREVIOUS item while they are in flight (its Gram product landed long ago)
    //   A(t)   l2norm scan, exp
    //   park(t) write e(t) as fp16 into the group's own columns of the accumulator item t came from, signal the Gram issuer
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int g = (warp - EPI_WARP0) >> 2;      // epilogue group = image of the tile
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t tu = tmem_base + U_BASE + g * GRAM_N + lane_sel;
    const uint32_t tpark0 = tmem_base + park_col(g) + lane_sel;      // + ACC_PITCH for odd items
    float* xch = reinterpret_cast<float*>(smem + SMEM_XCH) + g * 4 * 40;
    uint32_t used = 0u;                         // completed phases of uready[g]
    long long w_tfull = 0, w_afull = 0, w_uready = 0; const long long t_begin = prof_on ? clock64() : 0;
    Carry c;
    uint32_t hvp[18];                           // the carried item's numerators (fp16 pairs), as the tensor core saw them
#pragma unroll
    for (int k = 0; k < 18; ++k) hvp[k] = 0u;
    c.live = false; c.valid = false; c.P = c.D = c.wnorm = 0.f; c.cap = -1; c.seg = 0; c.n_words = 0; c.img = 0; c.b = 0;

    auto phase_b = [&]() {
      // ---------------- phase B of the carried item: cosine + aggregation + store -------------
      if (c.valid) {
        const int seg_lo = c.seg & 0xff, seg_hi = (c.seg >> 8) & 0xff;
        const bool long_tile = (c.seg >> 16) & 1;
        mbar_wait_sleep_t(uready_bar(g), used++ & 1, w_uready, prof_on);
        tc_fence_after();
        // off-diagonal part of e^T G e with the fp16-rounded e the tensor core saw (symmetric form); U in two
        // halves of 18 columns: the raw affinities of the NEXT item are already on their way into registers
        float q0 = 0.f, q1 = 0.f, Zsum;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t Uh[18];
          TMEM_LD_X16U(tu + 18 * h, Uh, 0);
          TMEM_LD_X2U(tu + 18 * h + 16, Uh, 16);
          if (h == 1) TMEM_LD_X1F(tu + 36, Zsum);            // ones column of the Gram pack: sum_k e_k
          tmem_ld_wait();
          if (!(PROF && (p.skip_math & 16)))
#pragma unroll
          for (int cidx = 0; cidx < 9; ++cidx) {
            float2 ef = unpack_f16x2(hvp[9 * h + cidx]);
            q0 = fmaf(ef.x, __uint_as_float(Uh[2 * cidx]), q0); q1 = fmaf(ef.y, __uint_as_float(Uh[2 * cidx + 1]), q1);
          }
        }
        // e^T G e = sum e_k^2 (unit diagonal, fp32) + e^T (G - I) e (tensor core, fp16 operands)
        const float Qf = c.D + (q0 + q1);
        // r_j = (P/Z) / max(|w| sqrt(Q)/Z, 1e-8)
        const float rj = c.P / fmaxf(c.wnorm * sqrtf(fmaxf(Qf, 0.f)), 1e-8f * Zsum);
        bool pr[5];
#pragma unroll
        for (int s = 0; s < 5; ++s) pr[s] = (lane - (1 << s)) >= seg_lo;
        float v = (p.agg == ITR_AGG_LSE) ? ex2f(rj * p.c_lse) : rj;
        if (c.cap < 0) v = (p.agg == ITR_AGG_MAX) ? -INFINITY : 0.f;
        float tot;
        if (!long_tile) {
          tot = (p.agg == ITR_AGG_MAX) ? seg_total<true>(v, pr, seg_hi) : seg_total<false>(v, pr, seg_hi);
        } else {
          float* x = xch + 36;
          tot = (p.agg == ITR_AGG_MAX) ? warp_max(v) : warp_sum(v);
          named_bar_sync(1 + g, 128);
          if (lane == 0) x[q * 40] = tot;
          named_bar_sync(1 + g, 128);
          float t0 = x[0], t1 = x[40], t2 = x[80], t3 = x[120];
          tot = (p.agg == ITR_AGG_MAX) ? fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)) : (t0 + t1) + (t2 + t3);
        }
        if (p.agg == ITR_AGG_LSE) tot = lg2f(tot) * p.inv_lse;
        if (p.agg == ITR_AGG_MEAN) tot = tot / (float)c.n_words;
        const bool writer = long_tile ? (row == 0) : (lane == seg_lo);
        if (writer && c.cap >= 0) p.scores[(size_t)c.img * p.ld + c.cap] = tot;
      }
      if (c.live) {
        // the Gram MMA of that item has completed (or the image was a tail image): its aux buffer is free
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(aempty_bar(c.b));
      }
    };

    int it = 0;
    for (ItemIter item(sched, first, step); DEBUG ? it == 0 : item.valid(); item.next(), ++it) {
      const int n = DEBUG ? p.dbg_n : item.n;
      const int b = it & 1;
      int4 meta = make_int4(-1, 0, 0, 0);
      float wnorm = 0.f;
      if (!DUMP) {
        mbar_wait_sleep_t(afull_bar(b), (it >> 1) & 1, w_afull, prof_on);
        const uint8_t* aux = smem + SMEM_AUX + b * AUX_BYTES;
        meta = reinterpret_cast<const int4*>(aux + AUX_GRAM)[row];
        wnorm = reinterpret_cast<const float*>(aux + AUX_GRAM + AUX_META)[row];
      }
      const int seg_lo = meta.z & 0xff, seg_hi = (meta.z >> 8) & 0xff;
      const bool long_tile = (meta.z >> 16) & 1;
      const int img = n * IMGS + g;
      const bool valid = !DEBUG && !DUMP && img < p.n_img && !(PROF && (p.skip_math & 1));

      // ---------------- raw affinities of item t on their way to registers ... ---------------------
      const uint32_t tacc = tmem_base + b * ACC_PITCH + lane_sel;
      mbar_wait_sleep_t(tfull_bar(b), (it >> 1) & 1, w_tfull, prof_on);
      tc_fence_after();
      float A[R];
      TMEM_LD_X32(tacc + g * R, A, 0);
      TMEM_LD_X4(tacc + g * R + 32, A, 32);

      // ---------------- ... while phase B(t-1) finishes the previous item (its Gram product has landed) ----
      phase_b();

      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(loaded_bar(b));
      if (DEBUG) {
#pragma unroll
        for (int k = 0; k < R; ++k) p.dump[(size_t)row * BLOCK_N + g * R + k] = A[k];
      }
      if (DUMP && img < p.n_img) {
        // [word tile][image][row][36]: a warp writes 32 rows x 144 bytes = 4.6 KB contiguous
        float4* dst = reinterpret_cast<float4*>(p.dump + (((size_t)item.m * p.n_img + img) * BLOCK_M + row) * R);
#pragma unroll
        for (int k = 0; k < R; k += 4) dst[k / 4] = make_float4(A[k], A[k + 1], A[k + 2], A[k + 3]);
      }

      uint32_t hv[18];
      float P = 0.f, Dd = 0.f;
      if (PROF && (p.skip_math & 32)) {
        // tuning only: pure register ALU work (no SMEM / TMEM / MUFU / SHFL), ~1150 dependent-free FMAs
        float x0 = A[0], x1 = A[1], x2 = A[2], x3 = A[3];
#pragma unroll 1
        for (int r = 0; r < 36; ++r) {
#pragma unroll
          for (int u = 0; u < 8; ++u) { x0 = fmaf(x0, 1.0001f, 0.5f); x1 = fmaf(x1, 1.0001f, 0.25f); x2 = fmaf(x2, 0.9999f, 0.125f); x3 = fmaf(x3, 0.9999f, 1.f); }
        }
        if (x0 + x1 + x2 + x3 == 123.456f) p.scores[0] = x0;
      }
      if (valid) {
        bool pr[5];
#pragma unroll
        for (int s = 0; s < 5; ++s) pr[s] = (lane - (1 << s)) >= seg_lo;
        const float shift = -fabsf(p.c_sm);
#pragma unroll
        for (int h = 0; h < 2; ++h) {          // two halves of 18 regions: bounds the live registers
          float E[R / 2];
#pragma unroll
          for (int k = 0; k < R / 2; ++k) {
            float a = p.clipped ? fmaxf(A[18 * h + k], 0.1f * A[18 * h + k]) : A[18 * h + k];
            E[k] = a * a;
          }
          // l2norm denominators: sum over the caption's words of a^2, per region
          if (PROF && (p.skip_math & 4)) {
            // tuning only: no cross-lane reduction
          } else if (!long_tile) {
#pragma unroll
            for (int k = 0; k < R / 2; ++k) E[k] = seg_total<false>(E[k], pr, seg_hi);
          } else {
#pragma unroll
            for (int k = 0; k < R / 2; ++k) E[k] = warp_sum(E[k]);
            named_bar_sync(1 + g, 128);        // previous readers of the exchange buffer are done
            if (lane == 0) {
#pragma unroll
              for (int k = 0; k < R / 2; ++k) xch[q * 40 + k] = E[k];
            }
            named_bar_sync(1 + g, 128);
#pragma unroll
            for (int k = 0; k < R / 2; ++k) E[k] = (xch[k] + xch[40 + k]) + (xch[80 + k] + xch[120 + k]);
          }
          // e_k = exp2(lambda ahat_k - lambda) <= 1 (|ahat| <= 1): no overflow in fp16 for any lambda.
          // 1/(sqrt(S)+1e-8) = rsqrt(S) (1 - O(1e-8 rsqrt(S))): the shortcut is exact to 3e-4 relative
          // for S >= 1e-9; below that (numerically orthogonal caption/region) take the exact form.
          float smin = E[0];
#pragma unroll
          for (int k = 1; k < R / 2; ++k) smin = fminf(smin, E[k]);
          const bool exact = __any_sync(0xffffffffu, smin < 1e-9f && meta.x >= 0);
          if (PROF && (p.skip_math & 8)) {
#pragma unroll
            for (int k = 0; k < R / 2; ++k) { P += E[k]; }      // tuning only: no exp / rsqrt
          } else if (!exact) {
#pragma unroll
            for (int k = 0; k < R / 2; ++k) {
              const float raw = A[18 * h + k];
              const float a = p.clipped ? fmaxf(raw, 0.1f * raw) : raw;
              const float e = ex2f(fmaf(a, p.c_sm * rsqf(E[k]), shift));
              E[k] = e; P = fmaf(e, raw, P); Dd = fmaf(e, e, Dd);
            }
          } else {
#pragma unroll
            for (int k = 0; k < R / 2; ++k) {
              const float raw = A[18 * h + k];
              const float a = p.clipped ? fmaxf(raw, 0.1f * raw) : raw;
              const float e = ex2f(fmaf(a, __fdividef(p.c_sm, sqrtf(E[k]) + 1e-8f), shift));
              E[k] = e; P = fmaf(e, raw, P); Dd = fmaf(e, e, Dd);
            }
          }
#pragma unroll
          for (int cidx = 0; cidx < 9; ++cidx) hv[9 * h + cidx] = pack_f16x2(E[2 * cidx], E[2 * cidx + 1]);
        }
      }

      // ---------------- park(t): e(t) as fp16 (K padded 36 -> 48 with zeros) in the group's own columns of the
      // accumulator it came from, wake the Gram issuer; the same values stay in registers for phase B(t)
      if (valid) {
        const uint32_t tpark = tpark0 + b * ACC_PITCH;
        uint32_t z[6] = {0u, 0u, 0u, 0u, 0u, 0u};
        TMEM_ST_X16(tpark, hv, 0);
        TMEM_ST_X2(tpark + 16, hv, 16);
        TMEM_ST_X4(tpark + 18, z, 0);
        TMEM_ST_X2(tpark + 22, z, 4);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(eready_bar(g));
      }
#pragma unroll
      for (int k = 0; k < 18; ++k) hvp[k] = hv[k];
      c.P = P; c.D = Dd; c.wnorm = wnorm; c.cap = meta.x; c.seg = meta.z; c.n_words = meta.w;
      c.img = img; c.b = b; c.valid = valid; c.live = !DUMP;
    }
    phase_b();                                    // drain
    if (prof_on && lane == 0 && q == 0) {
      long long* o = p.prof + blockIdx.x * 16 + 6 + g * 2;      // groups 0..3 -> slots 6..13
      o[0] = w_tfull + w_afull; o[1] = w_uready;
      if (g == 0) { p.prof[blockIdx.x * 16 + 14] = clock64() - t_begin; p.prof[blockIdx.x * 16 + 15] = w_afull; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------- prep kernels
// One warp per packed row: gather the word (or zeros), round to bf16, norm of the rounded row.
// 128 threads x <= 32 registers per block: exactly the register-file slice the persistent score kernel leaves free
// (640 threads x 96 registers), so the PCIe gather of the next caption chunk co-resides with a running score kernel
// instead of queueing behind it (ops.scan_t2i_scores_from_host).
__global__ void __launch_bounds__(128, 16)
pack_words_kernel(const float* __restrict__ captions, int lmax, int d, const int4* __restrict__ row_meta, int n_rows,
                  uint16_t* __restrict__ out, float* __restrict__ wnorm) {
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 4 + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * 4) {
    const int4 meta = row_meta[row];
    uint2* dst = reinterpret_cast<uint2*>(out + (size_t)row * d);
    float ss = 0.f;
    if (meta.x < 0) {
      for (int v = lane; v < d / 4; v += 32) dst[v] = make_uint2(0u, 0u);
    } else {
      const float4* src = reinterpret_cast<const float4*>(captions + ((size_t)meta.x * lmax + meta.y) * d);
#pragma unroll 4
      for (int v = lane; v < d / 4; v += 32) {
        float4 x = src[v];
        uint16_t b0 = f32_to_bf16_rn(x.x), b1 = f32_to_bf16_rn(x.y), b2 = f32_to_bf16_rn(x.z), b3 = f32_to_bf16_rn(x.w);
        float r0 = bf16_to_f32(b0), r1 = bf16_to_f32(b1), r2 = bf16_to_f32(b2), r3 = bf16_to_f32(b3);
        ss = fmaf(r0, r0, ss); ss = fmaf(r1, r1, ss); ss = fmaf(r2, r2, ss); ss = fmaf(r3, r3, ss);
        dst[v] = make_uint2((uint32_t)b0 | ((uint32_t)b1 << 16), (uint32_t)b2 | ((uint32_t)b3 << 16));
      }
    }
    ss = warp_sum(ss);
    if (lane == 0) wnorm[row] = sqrtf(ss);
  }
}

// One block (192 threads) per image: round the 36 regions to bf16 and build the image's Gram pack (fp16
// G - I in UMMA core-matrix order + fp32 diagonal + the ones row) from the ROUNDED regions.
// The 36x36x1024 Gram is register-tiled: 144 threads each own a 3x3 block of G and sweep K from shared memory
// (K-major rows with a 4-float skew -> conflict-free float4 reads); HBM-bound (6 B per element moved).
constexpr int PREP_THREADS = 192;
__global__ void __launch_bounds__(PREP_THREADS)
prep_images_kernel(const float* __restrict__ images, uint16_t* __restrict__ out, uint8_t* __restrict__ gram_pack) {
  extern __shared__ float sv[];   // [R][D+4] rounded values
  constexpr int LD = D + 4;
  const float* src = images + (size_t)blockIdx.x * R * D;
  uint16_t* dst = out + (size_t)blockIdx.x * R * D;
  uint8_t* gp = gram_pack + (size_t)blockIdx.x * GRAM_BYTES;
  for (int e = threadIdx.x; e < R * D / 4; e += PREP_THREADS) {
    float4 x = reinterpret_cast<const float4*>(src)[e];
    uint16_t b0 = f32_to_bf16_rn(x.x), b1 = f32_to_bf16_rn(x.y), b2 = f32_to_bf16_rn(x.z), b3 = f32_to_bf16_rn(x.w);
    reinterpret_cast<uint2*>(dst)[e] = make_uint2((uint32_t)b0 | ((uint32_t)b1 << 16), (uint32_t)b2 | ((uint32_t)b3 << 16));
    int r = (e * 4) / D, c = (e * 4) % D;
    *reinterpret_cast<float4*>(&sv[r * LD + c]) = make_float4(bf16_to_f32(b0), bf16_to_f32(b1), bf16_to_f32(b2), bf16_to_f32(b3));
  }
  // zero the fp16 block (padding rows / columns 36..47), then the ones row n = 36 (U[:, 36] = sum_k e_k)
  for (int e = threadIdx.x; e < G16_BYTES / 4; e += PREP_THREADS) reinterpret_cast<uint32_t*>(gp)[e] = 0u;
  __syncthreads();
  __half* g16 = reinterpret_cast<__half*>(gp);
  float* gd = reinterpret_cast<float*>(gp + G16_BYTES);
  auto g16_at = [&](int n_, int k_) -> __half& { return g16[(n_ / 8) * (G_SBO / 2) + (k_ / 8) * (G_LBO / 2) + (n_ % 8) * 8 + (k_ % 8)]; };
  if (threadIdx.x < R) g16_at(36, threadIdx.x) = __float2half_rn(1.0f);
  if (threadIdx.x < 144) {
    const int bi = threadIdx.x / 12, bj = threadIdx.x % 12;   // 12 x 12 blocks of 3 x 3
    if (bj <= bi) {                                            // lower triangle of blocks (G is symmetric)
      float4 acc4[3][3];                                       // four partial sums per output: accuracy + ILP
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc4[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* ra = sv + (3 * bi) * LD;
      const float* rb = sv + (3 * bj) * LD;
      for (int c = 0; c < D; c += 4) {
        float4 a[3], bq[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { a[i] = *reinterpret_cast<const float4*>(ra + i * LD + c); bq[i] = *reinterpret_cast<const float4*>(rb + i * LD + c); }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            acc4[i][j].x = fmaf(a[i].x, bq[j].x, acc4[i][j].x); acc4[i][j].y = fmaf(a[i].y, bq[j].y, acc4[i][j].y);
            acc4[i][j].z = fmaf(a[i].z, bq[j].z, acc4[i][j].z); acc4[i][j].w = fmaf(a[i].w, bq[j].w, acc4[i][j].w);
          }
      }
      float acc[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[i][j] = (acc4[i][j].x + acc4[i][j].y) + (acc4[i][j].z + acc4[i][j].w);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int k = 3 * bi + i, k2 = 3 * bj + j;
          if (k == k2) {
            gd[k] = acc[i][j];
            g16_at(k, k) = __float2half_rn(acc[i][j] - 1.0f);
          } else if (bi != bj || k2 < k) {
            const __half hs = __float2half_rn(acc[i][j]);
            g16_at(k, k2) = hs;
            g16_at(k2, k) = hs;
          }
        }
    }
  }
}

// ---------------------------------------------------------------------------- tcgen05 microbenchmark
// One CTA per SM issues `iters` kind::f16 MMAs (M = 128, K = 16) on whatever SMEM holds and reports
// cycles per MMA.  Used to establish the cost model the tile shape was chosen on.
//   n_cols : UMMA N (multiple of 16, <= 256)        n_acc : accumulators cycled through (1..3 with N <= 160)
//   a_tmem : 1 = A operand from tensor memory (TS), 0 = from shared memory (SS)
//   kadv   : 1 = walk the four K sub-steps of a 128-byte swizzle row like the real kernel, 0 = same address
__global__ void __launch_bounds__(192, 1)
mma_microbench_kernel(int n_cols, int n_acc, int iters, int a_tmem, int kadv, int n_issuers, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ long long t_issuer[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_ptr;
  if (warp >= 1 && warp <= n_issuers) {
    const int w = warp - 1;                     // issuer w owns accumulators at columns w*... (n_acc * n_issuers * n_cols <= 480)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc = umma_desc_sw128(sbase), bdesc = umma_desc_sw128(sbase + 16384);
    const long long t0 = clock64();
    int acc = 0;
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tbase + (uint32_t)((w * n_acc + acc) * n_cols);
      const int k = kadv ? (i & 3) : 0;
      if (elect_one()) {
        if (a_tmem) umma_f16_ts(d, tbase + 480, bdesc + 2 * k, idesc, 1u);
        else umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
      }
      __syncwarp();
      if (++acc == n_acc) acc = 0;
    }
    if (elect_one()) umma_commit(smem_u32(&bars[w]));
    __syncwarp();
    mbar_wait(smem_u32(&bars[w]), 0);
    const long long t1 = clock64();
    if (lane == 0) t_issuer[w] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    long long m = 0;
    for (int i = 0; i < n_issuers; ++i) m = t_issuer[i] > m ? t_issuer[i] : m;
    out[blockIdx.x] = m;
  }
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
  }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    ITR_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !ptr) return fail(ITR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cached = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  *fn = cached;
  return ITR_OK;
}

// 2-D bf16 tensor [rows][1024], box = [box_rows][64], 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ITR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return ITR_OK;
}

static int require_sm100() {
  int dev = 0;
  ITR_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0;
  ITR_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(ITR_ERR_UNSUPPORTED, "the tensor-core SCAN path needs an sm_100 device (found sm_%d0)", major);
  return ITR_OK;
}

}  // namespace tc
}  // namespace itr

namespace itr {
namespace tc2 {
// CTA-pair form of the fused kernel (scan_t2i_tc2.cu)
int launch_tc2(const uint16_t* images_bf16, const void* gram_pack, int n_img, const uint16_t* words_bf16,
               const int32_t* row_meta, const float* row_wnorm, int n_tiles, int feature_norm, int agg,
               float lambda_softmax, float lambda_lse, float* scores, int64_t ld_scores, void* stream, long long* prof);
}  // namespace tc2
}  // namespace itr

using namespace itr;
using namespace itr::tc;

// ITR_B200_SCORE_KERNEL=single selects the one-CTA kernel of round 1 (A/B measurements); default: the CTA-pair kernel
static bool use_pair_kernel() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("ITR_B200_SCORE_KERNEL");
    cached = (e && e[0] == 's') ? 0 : 1;
  }
  return cached == 1;
}

extern "C" int itr_scan_pack_words_bf16(const float* captions, int n_cap, int lmax, int d, const int32_t* row_meta,
                                        int n_tiles, uint16_t* words_bf16, float* row_wnorm, void* stream) {
  ITR_REQUIRE(captions && row_meta && words_bf16 && row_wnorm, "itr_scan_pack_words_bf16: null pointer");
  ITR_REQUIRE(d == D, "itr_scan_pack_words_bf16: built for embed size %d, got %d", D, d);
  ITR_REQUIRE(n_cap >= 0 && lmax >= 1 && n_tiles >= 0, "itr_scan_pack_words_bf16: bad shape");
  if (n_tiles == 0) return ITR_OK;
  const int n_rows = n_tiles * BLOCK_M;
  // Host (pinned) source: the gather is PCIe-bound, so one resident block per SM is plenty -- and one block per SM is
  // what fits next to a running score CTA whatever the launch order.  Device source: one block per 4 rows (HBM-bound).
  int grid = (n_rows + 3) / 4;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, captions) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (grid > sms) grid = sms;
  } else {
    cudaGetLastError();   // a plain cudaMalloc'ed / unregistered pointer is not an error here
  }
  pack_words_kernel<<<grid, 128, 0, as_stream(stream)>>>(captions, lmax, d, reinterpret_cast<const int4*>(row_meta),
                                                         n_rows, words_bf16, row_wnorm);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_scan_prep_images_bf16(const float* images, int n_img, int n_regions, int d, uint16_t* images_bf16,
                                         void* gram_pack, void* stream) {
  ITR_REQUIRE(images && images_bf16 && gram_pack, "itr_scan_prep_images_bf16: null pointer");
  ITR_REQUIRE(n_regions == R && d == D, "itr_scan_prep_images_bf16: built for %d regions x %d dims, got %d x %d", R, D, n_regions, d);
  if (n_img <= 0) return ITR_OK;
  const int smem = R * (D + 4) * 4;
  ITR_CHECK_CUDA(cudaFuncSetAttribute(prep_images_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  prep_images_kernel<<<n_img, PREP_THREADS, smem, as_stream(stream)>>>(images, images_bf16, reinterpret_cast<uint8_t*>(gram_pack));
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

static int launch_tc(const uint16_t* images_bf16, const void* gram_pack, int n_img, const uint16_t* words_bf16,
                     const int32_t* row_meta, const float* row_wnorm, int n_tiles, int feature_norm, int agg,
                     float lambda_softmax, float lambda_lse, float* scores, int64_t ld_scores, float* dump, int dbg_m,
                     int dbg_n, void* stream, long long* prof = nullptr, int mode = -1, bool full_dump = false) {
  int rc = require_sm100();
  if (rc) return rc;
  CUtensorMap map_w, map_i;
  rc = make_map(&map_w, words_bf16, (uint64_t)n_tiles * BLOCK_M, BLOCK_M);
  if (rc) return rc;
  rc = make_map(&map_i, images_bf16, (uint64_t)n_img * R, BLOCK_N);
  if (rc) return rc;
  Params p{};
  p.gram_pack = reinterpret_cast<const uint8_t*>(gram_pack);
  p.row_meta = reinterpret_cast<const int4*>(row_meta);
  p.row_wnorm = row_wnorm;
  p.n_img = n_img; p.n_wt = n_tiles; p.n_it = (n_img + IMGS - 1) / IMGS;
  p.clipped = (feature_norm == ITR_NORM_CLIPPED_L2); p.agg = agg;
  p.c_sm = lambda_softmax * 1.4426950408889634f;
  p.c_lse = lambda_lse * 1.4426950408889634f;
  p.inv_lse = 0.6931471805599453f / lambda_lse;
  p.scores = scores; p.ld = ld_scores; p.dump = dump; p.dbg_m = dbg_m; p.dbg_n = dbg_n; p.prof = prof;
  {
    static int env_ctrl_last = -1;
    if (env_ctrl_last < 0) { const char* e = getenv("ITR_B200_CTRL_LAST"); env_ctrl_last = e ? atoi(e) : 1; }
    p.ctrl_last = mode >= 0 ? (mode & 1) : env_ctrl_last;
    p.skip_math = mode >= 0 ? ((mode >> 1) & 63) : 0;
    p.wait_mode = mode >= 0 ? ((mode >> 7) & 3) : 0;    // tuning bits: 1 no epilogue arithmetic, 2 no TMA loads, 4 no scan, 8 no exp, 16 no phase-B dot
  }
  int dev = 0, sms = 0;
  ITR_CHECK_CUDA(cudaGetDevice(&dev));
  ITR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (dump && !full_dump) {
    ITR_CHECK_CUDA(cudaFuncSetAttribute(scan_t2i_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    scan_t2i_tc_kernel<true, false><<<1, NUM_THREADS, SMEM_ALLOC, as_stream(stream)>>>(map_w, map_i, p);
  } else if (full_dump) {
    long long total = (long long)p.n_wt * p.n_it;
    if (total >= (1ll << 31)) return fail(ITR_ERR_INVALID, "itr_scan_affinity_bf16: %lld tiles exceed the 2^31 scheduler range; split the call", total);
    long long units = (long long)((p.n_wt + BAND - 1) / BAND) * p.n_it;
    int grid = (int)(units < sms ? units : sms);
    ITR_CHECK_CUDA(cudaFuncSetAttribute(scan_t2i_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    scan_t2i_tc_kernel<false, false, true><<<grid, NUM_THREADS, SMEM_ALLOC, as_stream(stream)>>>(map_w, map_i, p);
  } else {
    long long total = (long long)p.n_wt * p.n_it;
    if (total >= (1ll << 31)) return fail(ITR_ERR_INVALID, "itr_scan_t2i_scores_bf16: %lld tiles exceed the 2^31 scheduler range; split the call", total);
    long long units = (long long)((p.n_wt + BAND - 1) / BAND) * p.n_it;
    int grid = (int)(units < sms ? units : sms);
    if (prof) {
      ITR_CHECK_CUDA(cudaFuncSetAttribute(scan_t2i_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
      scan_t2i_tc_kernel<false, true><<<grid, NUM_THREADS, SMEM_ALLOC, as_stream(stream)>>>(map_w, map_i, p);
    } else {
      ITR_CHECK_CUDA(cudaFuncSetAttribute(scan_t2i_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
      scan_t2i_tc_kernel<false, false><<<grid, NUM_THREADS, SMEM_ALLOC, as_stream(stream)>>>(map_w, map_i, p);
    }
  }
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_scan_t2i_scores_bf16(const uint16_t* images_bf16, const void* gram_pack, int n_img,
                                        const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm,
                                        int n_tiles, int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                                        float* scores, int64_t ld_scores, void* stream) {
  ITR_REQUIRE(images_bf16 && gram_pack && words_bf16 && row_meta && row_wnorm && scores, "itr_scan_t2i_scores_bf16: null pointer");
  ITR_REQUIRE(feature_norm == ITR_NORM_CLIPPED_L2 || feature_norm == ITR_NORM_L2,
              "itr_scan_t2i_scores_bf16: raw_feature_norm %d is only available in the float32 path", feature_norm);
  ITR_REQUIRE(agg >= 0 && agg <= ITR_AGG_SUM, "unknown aggfunc: %d", agg);
  ITR_REQUIRE(lambda_lse != 0.f || agg != ITR_AGG_LSE, "itr_scan_t2i_scores_bf16: lambda_lse must be non-zero");
  ITR_REQUIRE(lambda_softmax > -80.f && lambda_softmax < 80.f, "itr_scan_t2i_scores_bf16: |lambda_softmax| must be < 80");
  ITR_REQUIRE(((uintptr_t)images_bf16 & 15) == 0 && ((uintptr_t)words_bf16 & 15) == 0 && ((uintptr_t)gram_pack & 15) == 0 &&
              ((uintptr_t)row_meta & 15) == 0 && ((uintptr_t)row_wnorm & 15) == 0, "itr_scan_t2i_scores_bf16: buffers must be 16-byte aligned");
  if (n_img <= 0 || n_tiles <= 0) return ITR_OK;
  if (use_pair_kernel()) {
    int rc = require_sm100();
    if (rc) return rc;
    return tc2::launch_tc2(images_bf16, gram_pack, n_img, words_bf16, row_meta, row_wnorm, n_tiles, feature_norm, agg,
                           lambda_softmax, lambda_lse, scores, ld_scores, stream, nullptr);
  }
  return launch_tc(images_bf16, gram_pack, n_img, words_bf16, row_meta, row_wnorm, n_tiles, feature_norm, agg,
                   lambda_softmax, lambda_lse, scores, ld_scores, nullptr, 0, 0, stream);
}

extern "C" int itr_scan_t2i_pair_profile(const uint16_t* images_bf16, const void* gram_pack, int n_img,
                                         const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm,
                                         int n_tiles, float* scores, int64_t ld_scores, int64_t* counters, void* stream) {
  ITR_REQUIRE(images_bf16 && gram_pack && words_bf16 && row_meta && row_wnorm && scores && counters, "itr_scan_t2i_pair_profile: null pointer");
  if (n_img <= 0 || n_tiles <= 0) return ITR_OK;
  int rc = require_sm100();
  if (rc) return rc;
  return tc2::launch_tc2(images_bf16, gram_pack, n_img, words_bf16, row_meta, row_wnorm, n_tiles, ITR_NORM_CLIPPED_L2, ITR_AGG_LSE,
                         9.f, 6.f, scores, ld_scores, stream, reinterpret_cast<long long*>(counters));
}

extern "C" int itr_scan_t2i_affinity_debug(const uint16_t* images_bf16, int n_img, const uint16_t* words_bf16, int n_tiles,
                                           int word_tile, int image_tile, float* out, void* stream) {
  ITR_REQUIRE(images_bf16 && words_bf16 && out, "itr_scan_t2i_affinity_debug: null pointer");
  ITR_REQUIRE(word_tile >= 0 && word_tile < n_tiles && image_tile >= 0 && image_tile * IMGS < n_img,
              "itr_scan_t2i_affinity_debug: tile index out of range");
  // aux loads still run; point them at the operand buffers (contents unused in debug mode)
  return launch_tc(images_bf16, images_bf16, n_img, words_bf16,
                   reinterpret_cast<const int32_t*>(words_bf16), reinterpret_cast<const float*>(words_bf16), n_tiles,
                   ITR_NORM_CLIPPED_L2, ITR_AGG_SUM, 1.f, 1.f, out, 0, out, word_tile, image_tile, stream);
}

extern "C" int itr_scan_t2i_profile(const uint16_t* images_bf16, const void* gram_pack, int n_img,
                                    const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm,
                                    int n_tiles, float* scores, int64_t ld_scores, int64_t* counters, int mode, void* stream) {
  ITR_REQUIRE(images_bf16 && gram_pack && words_bf16 && row_meta && row_wnorm && scores && counters, "itr_scan_t2i_profile: null pointer");
  if (n_img <= 0 || n_tiles <= 0) return ITR_OK;
  return launch_tc(images_bf16, gram_pack, n_img, words_bf16, row_meta, row_wnorm, n_tiles, ITR_NORM_CLIPPED_L2, ITR_AGG_LSE,
                   9.f, 6.f, scores, ld_scores, nullptr, 0, 0, stream, reinterpret_cast<long long*>(counters), mode);
}

extern "C" int itr_tc_mma_microbench(int n_cols, int n_acc, int iters, int a_tmem, int kadv, int n_issuers, int n_ctas, int64_t* cycles, void* stream) {
  ITR_REQUIRE(cycles && n_cols >= 16 && n_cols <= 256 && n_cols % 16 == 0 && n_acc >= 1 && n_issuers >= 1 && n_issuers <= 4 &&
              n_acc * n_issuers * n_cols <= 480 && iters > 0 && n_ctas > 0, "itr_tc_mma_microbench: bad arguments");
  int rc = require_sm100();
  if (rc) return rc;
  const int smem = 64 * 1024 + 1024;
  ITR_CHECK_CUDA(cudaFuncSetAttribute(mma_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_microbench_kernel<<<n_ctas, 192, smem, as_stream(stream)>>>(n_cols, n_acc, iters, a_tmem, kadv, n_issuers, reinterpret_cast<long long*>(cycles));
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_scan_affinity_bf16(const uint16_t* images_bf16, int n_img, const uint16_t* words_bf16, int n_tiles,
                                      float* affinity, void* stream) {
  ITR_REQUIRE(images_bf16 && words_bf16 && affinity, "itr_scan_affinity_bf16: null pointer");
  ITR_REQUIRE(((uintptr_t)images_bf16 & 15) == 0 && ((uintptr_t)words_bf16 & 15) == 0 && ((uintptr_t)affinity & 15) == 0,
              "itr_scan_affinity_bf16: buffers must be 16-byte aligned");
  if (n_img <= 0 || n_tiles <= 0) return ITR_OK;
  return launch_tc(images_bf16, images_bf16, n_img, words_bf16, reinterpret_cast<const int32_t*>(words_bf16),
                   reinterpret_cast<const float*>(words_bf16), n_tiles, ITR_NORM_CLIPPED_L2, ITR_AGG_SUM, 1.f, 1.f, affinity, 0,
                   affinity, 0, 0, stream, nullptr, -1, true);
}
