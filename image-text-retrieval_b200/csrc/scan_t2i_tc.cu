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
// Kernel structure: persistent, warp-specialised, one CTA per SM.
//   tile       = 128 packed words (UMMA M) x 4 images = 144 region columns (UMMA N), K = 1024
//   warp 0     TMA producer: 5-stage ring of {words 128x64, regions 144x64} bf16 tiles, SWIZZLE_128B,
//              plus per-tile aux data (4 packed Grams, row metadata, word norms) by bulk copy
//   warp 1     tcgen05.mma issuer (cta_group::1, kind::f16, bf16 x bf16 -> fp32 in TMEM)
//   warp 2     TMEM allocator (2 accumulator buffers of 144 columns)
//   warps 4-11 epilogue: thread = one word row; tcgen05.ld its 36 columns per image, then the
//              whole softmax / cosine / aggregation chain in registers; the two cross-row
//              reductions (l2norm over the caption's words, aggregation over words) are
//              segmented warp scans -- itr_scan_plan_words guarantees a caption never
//              straddles a warp, except in `long` tiles which exchange through shared memory.
#include <cuda.h>

#include "common.cuh"

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
constexpr int GRAM_FLOATS = ITR_GRAM_TRI;      // 720 per image
constexpr int AUX_GRAM = IMGS * GRAM_FLOATS * 4;   // 11520
constexpr int AUX_META = BLOCK_M * 16;             // 2048
constexpr int AUX_WNORM = BLOCK_M * 4;             // 512
constexpr int AUX_BYTES = AUX_GRAM + AUX_META + AUX_WNORM;   // 14080
constexpr int XCH_FLOATS = 2 /*half*/ * 2 /*image*/ * 4 /*warp*/ * 40;
constexpr int ACC_COLS = 256;                  // column stride between the two accumulators
constexpr int TMEM_COLS = 512;
constexpr int BAND = 32;                       // word tiles kept L2-resident while images stream
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 8;

constexpr int SMEM_STAGES = 0;
constexpr int SMEM_AUX = SMEM_STAGES + STAGES * STAGE_BYTES;
constexpr int SMEM_XCH = SMEM_AUX + 2 * AUX_BYTES;
constexpr int SMEM_BARS = SMEM_XCH + XCH_FLOATS * 4;
constexpr int NUM_BARS = 2 * STAGES + 8;
constexpr int SMEM_TMEMPTR = SMEM_BARS + NUM_BARS * 8;
constexpr int SMEM_BYTES = SMEM_TMEMPTR + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;   // slack for manual 1024-byte alignment

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=144
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (reported as a CUDA error) instead of a hang.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, SWIZZLE_128B operand tile whose rows are 128 bytes: 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
#define TMEM_LD_X32(taddr, v, o)                                                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"        \
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                              \
               : "=f"(v[o + 0]), "=f"(v[o + 1]), "=f"(v[o + 2]), "=f"(v[o + 3]), "=f"(v[o + 4]), "=f"(v[o + 5]),       \
                 "=f"(v[o + 6]), "=f"(v[o + 7]), "=f"(v[o + 8]), "=f"(v[o + 9]), "=f"(v[o + 10]), "=f"(v[o + 11]),     \
                 "=f"(v[o + 12]), "=f"(v[o + 13]), "=f"(v[o + 14]), "=f"(v[o + 15]), "=f"(v[o + 16]), "=f"(v[o + 17]), \
                 "=f"(v[o + 18]), "=f"(v[o + 19]), "=f"(v[o + 20]), "=f"(v[o + 21]), "=f"(v[o + 22]), "=f"(v[o + 23]), \
                 "=f"(v[o + 24]), "=f"(v[o + 25]), "=f"(v[o + 26]), "=f"(v[o + 27]), "=f"(v[o + 28]), "=f"(v[o + 29]), \
                 "=f"(v[o + 30]), "=f"(v[o + 31])                                                                      \
               : "r"(taddr))
#define TMEM_LD_X4(taddr, v, o)                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"           \
               : "=f"(v[o + 0]), "=f"(v[o + 1]), "=f"(v[o + 2]), "=f"(v[o + 3])       \
               : "r"(taddr))
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqf(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---------------------------------------------------------------------------- tile schedule
// Item t -> (word tile m, image tile n).  Word tiles are taken in bands of BAND; inside a band
// the word tile varies fastest so the CTAs of one wave share a handful of image tiles and one
// band of word tiles (both L2 resident) while the image set streams from HBM once per band.
struct Schedule {
  int n_wt, n_it, full_items, last_band;
  __device__ Schedule(int n_wt_, int n_it_) : n_wt(n_wt_), n_it(n_it_) {
    int full_bands = n_wt / BAND;
    last_band = n_wt - full_bands * BAND;
    full_items = full_bands * BAND * n_it;
  }
  __device__ long long total() const { return (long long)n_wt * n_it; }
  __device__ void map(long long t, int& m, int& n) const {
    if (t < full_items) {
      int band = (int)(t / ((long long)BAND * n_it));
      int local = (int)(t - (long long)band * BAND * n_it);
      n = local / BAND;
      m = band * BAND + local % BAND;
    } else {
      int local = (int)(t - full_items);
      n = local / last_band;
      m = (n_wt - last_band) + local % last_band;
    }
  }
};

struct Params {
  const float* gram_tri;       // [n_img][720]
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

template <bool DEBUG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
scan_t2i_tc_kernel(const __grid_constant__ CUtensorMap map_words, const __grid_constant__ CUtensorMap map_imgs, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bar0 = sbase + SMEM_BARS;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 + b); };
  auto afull_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 4 + b); };
  auto aempty_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 6 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + SMEM_TMEMPTR);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_words) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_imgs) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), NUM_EPI_WARPS);
      mbar_init(afull_bar(b), 1); mbar_init(aempty_bar(b), NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SMEM_TMEMPTR), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const Schedule sched(p.n_wt, p.n_it);
  const long long total = DEBUG ? 1 : sched.total();
  const long long first = DEBUG ? 0 : blockIdx.x;
  const long long step = DEBUG ? 1 : gridDim.x;

  if (warp == 0) {
    // =============================== TMA producer =========================================
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (long long t = first; t < total; t += step, ++it) {
      int m, n;
      if (DEBUG) { m = p.dbg_m; n = p.dbg_n; } else sched.map(t, m, n);
      const int b = it & 1;
      mbar_wait(aempty_bar(b), ((it >> 1) & 1) ^ 1);
      if (lane == 0) {
        const int n_valid = min(IMGS, p.n_img - n * IMGS);
        const uint32_t aux = sbase + SMEM_AUX + b * AUX_BYTES;
        mbar_expect_tx(afull_bar(b), n_valid * GRAM_FLOATS * 4 + AUX_META + AUX_WNORM);
        bulk_load(aux, p.gram_tri + (size_t)n * IMGS * GRAM_FLOATS, n_valid * GRAM_FLOATS * 4, afull_bar(b));
        bulk_load(aux + AUX_GRAM, p.row_meta + (size_t)m * BLOCK_M, AUX_META, afull_bar(b));
        bulk_load(aux + AUX_GRAM + AUX_META, p.row_wnorm + (size_t)m * BLOCK_M, AUX_WNORM, afull_bar(b));
      }
      for (int kb = 0; kb < K_BLOCKS; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (lane == 0) {
          const uint32_t sa = sbase + SMEM_STAGES + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_2d(sa, &map_words, full_bar(stage), kb * BLOCK_K, m * BLOCK_M);
          tma_load_2d(sa + A_BYTES, &map_imgs, full_bar(stage), kb * BLOCK_K, n * BLOCK_N);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===========================================
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (long long t = first; t < total; t += step, ++it) {
      const int b = it & 1;
      mbar_wait(tempty_bar(b), ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + b * ACC_COLS;
      for (int kb = 0; kb < K_BLOCKS; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = sbase + SMEM_STAGES + stage * STAGE_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);   // +32 bytes per K step
          umma_commit(empty_bar(stage));
          if (kb == K_BLOCKS - 1) umma_commit(tfull_bar(b));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // =============================== epilogue =============================================
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int h = (warp - EPI_WARP0) >> 2;      // which pair of images of the tile
    const int row = q * 32 + lane;
    float* xch = reinterpret_cast<float*>(smem + SMEM_XCH);
    int it = 0;
    for (long long t = first; t < total; t += step, ++it) {
      int m, n;
      if (DEBUG) { m = p.dbg_m; n = p.dbg_n; } else sched.map(t, m, n);
      const int b = it & 1;
      const uint32_t par = (it >> 1) & 1;
      mbar_wait(afull_bar(b), par);
      const uint8_t* aux = smem + SMEM_AUX + b * AUX_BYTES;
      const int4 meta = reinterpret_cast<const int4*>(aux + AUX_GRAM)[row];
      const float wnorm = reinterpret_cast<const float*>(aux + AUX_GRAM + AUX_META)[row];
      const int cap = meta.x, seg_lo = meta.z & 0xff, seg_hi = (meta.z >> 8) & 0xff, n_words = meta.w;
      const bool long_tile = (meta.z >> 16) & 1;
      bool pr[5];
#pragma unroll
      for (int s = 0; s < 5; ++s) pr[s] = (lane - (1 << s)) >= seg_lo;

      mbar_wait(tfull_bar(b), par);
      tc_fence_after();
      const uint32_t tacc = tmem_base + b * ACC_COLS + ((uint32_t)(q * 32) << 16);

#pragma unroll 1
      for (int ii = 0; ii < 2; ++ii) {
        const int img_in_tile = h * 2 + ii;
        const int img = n * IMGS + img_in_tile;
        float A[R];
        TMEM_LD_X32(tacc + img_in_tile * R, A, 0);
        TMEM_LD_X4(tacc + img_in_tile * R + 32, A, 32);
        tmem_ld_wait();
        if (ii == 1) {
          // both images of this warp are in registers: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(b));
        }
        if (DEBUG) {
#pragma unroll
          for (int k = 0; k < R; ++k) p.dump[(size_t)row * BLOCK_N + img_in_tile * R + k] = A[k];
          continue;
        }
        if (img >= p.n_img) continue;           // warp-uniform: zero-filled tail of the image set

        // ---- l2norm denominators: sum over the caption's words of a^2, per region ----------
        float E[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
          float a = p.clipped ? fmaxf(A[k], 0.1f * A[k]) : A[k];
          E[k] = a * a;
        }
        if (!long_tile) {
#pragma unroll
          for (int k = 0; k < R; ++k) E[k] = seg_total<false>(E[k], pr, seg_hi);
        } else {
          float* x = xch + ((h * 2 + ii) * 4) * 40;
#pragma unroll
          for (int k = 0; k < R; ++k) E[k] = warp_sum(E[k]);
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < R; k += 4) *reinterpret_cast<float4*>(x + q * 40 + k) = make_float4(E[k], E[k + 1], E[k + 2], E[k + 3]);
          }
          named_bar_sync(1 + h, 128);
#pragma unroll
          for (int k = 0; k < R; k += 4) {
            float4 s0 = *reinterpret_cast<const float4*>(x + 0 * 40 + k), s1 = *reinterpret_cast<const float4*>(x + 1 * 40 + k);
            float4 s2 = *reinterpret_cast<const float4*>(x + 2 * 40 + k), s3 = *reinterpret_cast<const float4*>(x + 3 * 40 + k);
            E[k] = (s0.x + s1.x) + (s2.x + s3.x); E[k + 1] = (s0.y + s1.y) + (s2.y + s3.y);
            E[k + 2] = (s0.z + s1.z) + (s2.z + s3.z); E[k + 3] = (s0.w + s1.w) + (s2.w + s3.w);
          }
        }
        // ---- softmax numerators e_k = exp(lambda * ahat_k), Z, P -----------------------------
        float smin = E[0];
#pragma unroll
        for (int k = 1; k < R; ++k) smin = fminf(smin, E[k]);
        float Z = 0.f, P = 0.f;
        if (smin >= 1e-6f || cap < 0) {
          // 1/(sqrt(S)+1e-8) = r (1 - 1e-8 r) + O((1e-8 r)^2),  r = rsqrt(S) <= 1e3
#pragma unroll
          for (int k = 0; k < R; ++k) {
            float a = p.clipped ? fmaxf(A[k], 0.1f * A[k]) : A[k];
            float r = rsqf(E[k]);
            float cr = p.c_sm * r;
            float inv = fmaf(cr, -1e-8f * r, cr);
            float e = ex2f(a * inv);
            E[k] = e; Z += e; P = fmaf(e, A[k], P);
          }
        } else {
#pragma unroll
          for (int k = 0; k < R; ++k) {
            float a = p.clipped ? fmaxf(A[k], 0.1f * A[k]) : A[k];
            float inv = p.c_sm / (sqrtf(E[k]) + 1e-8f);
            float e = ex2f(a * inv);
            E[k] = e; Z += e; P = fmaf(e, A[k], P);
          }
        }
        // ---- |ctx|^2 * Z^2 = e^T G e with the packed lower-triangular Gram (diagonal halved) ----
        const float4* G = reinterpret_cast<const float4*>(aux + img_in_tile * GRAM_FLOATS * 4);
        float Qh0 = 0.f, Qh1 = 0.f;
        {
          int off = 0;
#pragma unroll
          for (int k = 0; k < R; ++k) {
            float u0 = 0.f, u1 = 0.f;
#pragma unroll
            for (int g = 0; g <= k / 4; ++g) {
              float4 gv = G[off + g];
              u0 = fmaf(gv.x, E[4 * g + 0], u0); u1 = fmaf(gv.y, E[4 * g + 1], u1);
              u0 = fmaf(gv.z, E[4 * g + 2], u0); u1 = fmaf(gv.w, E[4 * g + 3], u1);
            }
            off += k / 4 + 1;
            if (k & 1) Qh1 = fmaf(E[k], u0 + u1, Qh1); else Qh0 = fmaf(E[k], u0 + u1, Qh0);
          }
        }
        const float Qf = 2.f * (Qh0 + Qh1);
        // r_j = (P/Z) / max(|w| sqrt(Q)/Z, 1e-8)
        const float rj = P / fmaxf(wnorm * sqrtf(fmaxf(Qf, 0.f)), 1e-8f * Z);

        // ---- aggregate over the caption's words ------------------------------------------------
        float v = (p.agg == ITR_AGG_LSE) ? ex2f(rj * p.c_lse) : rj;
        if (cap < 0) v = (p.agg == ITR_AGG_MAX) ? -INFINITY : 0.f;
        float tot;
        if (!long_tile) {
          tot = (p.agg == ITR_AGG_MAX) ? seg_total<true>(v, pr, seg_hi) : seg_total<false>(v, pr, seg_hi);
        } else {
          float* x = xch + ((h * 2 + ii) * 4) * 40 + 36;
          tot = (p.agg == ITR_AGG_MAX) ? warp_max(v) : warp_sum(v);
          if (lane == 0) x[q * 40] = tot;
          named_bar_sync(1 + h, 128);
          float t0 = x[0], t1 = x[40], t2 = x[80], t3 = x[120];
          tot = (p.agg == ITR_AGG_MAX) ? fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)) : (t0 + t1) + (t2 + t3);
        }
        if (p.agg == ITR_AGG_LSE) tot = lg2f(tot) * p.inv_lse;
        if (p.agg == ITR_AGG_MEAN) tot = tot / (float)n_words;
        const bool writer = long_tile ? (row == 0) : (lane == seg_lo);
        if (writer && cap >= 0) p.scores[(size_t)img * p.ld + cap] = tot;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(aempty_bar(b));
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
__global__ void __launch_bounds__(256)
pack_words_kernel(const float* __restrict__ captions, int lmax, int d, const int4* __restrict__ row_meta, int n_rows,
                  uint16_t* __restrict__ out, float* __restrict__ wnorm) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int4 meta = row_meta[row];
  uint2* dst = reinterpret_cast<uint2*>(out + (size_t)row * d);
  float ss = 0.f;
  if (meta.x < 0) {
    for (int v = lane; v < d / 4; v += 32) dst[v] = make_uint2(0u, 0u);
  } else {
    const float4* src = reinterpret_cast<const float4*>(captions + ((size_t)meta.x * lmax + meta.y) * d);
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

// One block per image: round the 36 regions to bf16 and build the packed lower-triangular Gram.
__global__ void __launch_bounds__(256)
prep_images_kernel(const float* __restrict__ images, uint16_t* __restrict__ out, float* __restrict__ gram_tri) {
  extern __shared__ float sv[];   // [R][D+4] rounded values
  constexpr int LD = D + 4;
  const float* src = images + (size_t)blockIdx.x * R * D;
  uint16_t* dst = out + (size_t)blockIdx.x * R * D;
  for (int e = threadIdx.x; e < R * D / 4; e += 256) {
    float4 x = reinterpret_cast<const float4*>(src)[e];
    uint16_t b0 = f32_to_bf16_rn(x.x), b1 = f32_to_bf16_rn(x.y), b2 = f32_to_bf16_rn(x.z), b3 = f32_to_bf16_rn(x.w);
    reinterpret_cast<uint2*>(dst)[e] = make_uint2((uint32_t)b0 | ((uint32_t)b1 << 16), (uint32_t)b2 | ((uint32_t)b3 << 16));
    int r = (e * 4) / D, c = (e * 4) % D;
    *reinterpret_cast<float4*>(&sv[r * LD + c]) = make_float4(bf16_to_f32(b0), bf16_to_f32(b1), bf16_to_f32(b2), bf16_to_f32(b3));
  }
  __syncthreads();
  // 720 packed outputs: row k holds k2 = 0 .. 4*(k/4)+3, zero beyond k, diagonal halved
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* g = gram_tri + (size_t)blockIdx.x * GRAM_FLOATS;
  for (int o = warp; o < GRAM_FLOATS; o += 8) {
    // locate (k, k2) of packed offset o: rows come in groups of four with equal length 4*(grp+1)
    int grp = 0, base = 0;
    while (o >= base + 16 * (grp + 1)) { base += 16 * (grp + 1); ++grp; }
    int len = 4 * (grp + 1);
    int k = 4 * grp + (o - base) / len, k2 = (o - base) % len;
    float s = 0.f;
    if (k2 <= k) {
      const float* a = sv + k * LD;
      const float* bq = sv + k2 * LD;
      for (int c = lane; c < D; c += 32) s = fmaf(a[c], bq[c], s);
      s = warp_sum(s);
      if (k2 == k) s *= 0.5f;
    }
    if (lane == 0) g[o] = s;
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

using namespace itr;
using namespace itr::tc;

extern "C" int itr_scan_pack_words_bf16(const float* captions, int n_cap, int lmax, int d, const int32_t* row_meta,
                                        int n_tiles, uint16_t* words_bf16, float* row_wnorm, void* stream) {
  ITR_REQUIRE(captions && row_meta && words_bf16 && row_wnorm, "itr_scan_pack_words_bf16: null pointer");
  ITR_REQUIRE(d == D, "itr_scan_pack_words_bf16: built for embed size %d, got %d", D, d);
  ITR_REQUIRE(n_cap >= 0 && lmax >= 1 && n_tiles >= 0, "itr_scan_pack_words_bf16: bad shape");
  if (n_tiles == 0) return ITR_OK;
  const int n_rows = n_tiles * BLOCK_M;
  pack_words_kernel<<<(n_rows + 7) / 8, 256, 0, as_stream(stream)>>>(captions, lmax, d, reinterpret_cast<const int4*>(row_meta),
                                                                     n_rows, words_bf16, row_wnorm);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_scan_prep_images_bf16(const float* images, int n_img, int n_regions, int d, uint16_t* images_bf16,
                                         float* gram_tri, void* stream) {
  ITR_REQUIRE(images && images_bf16 && gram_tri, "itr_scan_prep_images_bf16: null pointer");
  ITR_REQUIRE(n_regions == R && d == D, "itr_scan_prep_images_bf16: built for %d regions x %d dims, got %d x %d", R, D, n_regions, d);
  if (n_img <= 0) return ITR_OK;
  const int smem = R * (D + 4) * 4;
  ITR_CHECK_CUDA(cudaFuncSetAttribute(prep_images_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  prep_images_kernel<<<n_img, 256, smem, as_stream(stream)>>>(images, images_bf16, gram_tri);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

static int launch_tc(const uint16_t* images_bf16, const float* gram_tri, int n_img, const uint16_t* words_bf16,
                     const int32_t* row_meta, const float* row_wnorm, int n_tiles, int feature_norm, int agg,
                     float lambda_softmax, float lambda_lse, float* scores, int64_t ld_scores, float* dump, int dbg_m,
                     int dbg_n, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  CUtensorMap map_w, map_i;
  rc = make_map(&map_w, words_bf16, (uint64_t)n_tiles * BLOCK_M, BLOCK_M);
  if (rc) return rc;
  rc = make_map(&map_i, images_bf16, (uint64_t)n_img * R, BLOCK_N);
  if (rc) return rc;
  Params p{};
  p.gram_tri = gram_tri;
  p.row_meta = reinterpret_cast<const int4*>(row_meta);
  p.row_wnorm = row_wnorm;
  p.n_img = n_img; p.n_wt = n_tiles; p.n_it = (n_img + IMGS - 1) / IMGS;
  p.clipped = (feature_norm == ITR_NORM_CLIPPED_L2); p.agg = agg;
  p.c_sm = lambda_softmax * 1.4426950408889634f;
  p.c_lse = lambda_lse * 1.4426950408889634f;
  p.inv_lse = 0.6931471805599453f / lambda_lse;
  p.scores = scores; p.ld = ld_scores; p.dump = dump; p.dbg_m = dbg_m; p.dbg_n = dbg_n;
  int dev = 0, sms = 0;
  ITR_CHECK_CUDA(cudaGetDevice(&dev));
  ITR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (dump) {
    ITR_CHECK_CUDA(cudaFuncSetAttribute(scan_t2i_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    scan_t2i_tc_kernel<true><<<1, NUM_THREADS, SMEM_ALLOC, as_stream(stream)>>>(map_w, map_i, p);
  } else {
    long long total = (long long)p.n_wt * p.n_it;
    int grid = (int)(total < sms ? total : sms);
    ITR_CHECK_CUDA(cudaFuncSetAttribute(scan_t2i_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    scan_t2i_tc_kernel<false><<<grid, NUM_THREADS, SMEM_ALLOC, as_stream(stream)>>>(map_w, map_i, p);
  }
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_scan_t2i_scores_bf16(const uint16_t* images_bf16, const float* gram_tri, int n_img,
                                        const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm,
                                        int n_tiles, int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                                        float* scores, int64_t ld_scores, void* stream) {
  ITR_REQUIRE(images_bf16 && gram_tri && words_bf16 && row_meta && row_wnorm && scores, "itr_scan_t2i_scores_bf16: null pointer");
  ITR_REQUIRE(feature_norm == ITR_NORM_CLIPPED_L2 || feature_norm == ITR_NORM_L2,
              "itr_scan_t2i_scores_bf16: raw_feature_norm %d is only available in the float32 path", feature_norm);
  ITR_REQUIRE(agg >= 0 && agg <= ITR_AGG_SUM, "unknown aggfunc: %d", agg);
  ITR_REQUIRE(lambda_lse != 0.f || agg != ITR_AGG_LSE, "itr_scan_t2i_scores_bf16: lambda_lse must be non-zero");
  ITR_REQUIRE(lambda_softmax > -80.f && lambda_softmax < 80.f, "itr_scan_t2i_scores_bf16: |lambda_softmax| must be < 80");
  ITR_REQUIRE(((uintptr_t)images_bf16 & 15) == 0 && ((uintptr_t)words_bf16 & 15) == 0 && ((uintptr_t)gram_tri & 15) == 0 &&
              ((uintptr_t)row_meta & 15) == 0 && ((uintptr_t)row_wnorm & 15) == 0, "itr_scan_t2i_scores_bf16: buffers must be 16-byte aligned");
  if (n_img <= 0 || n_tiles <= 0) return ITR_OK;
  return launch_tc(images_bf16, gram_tri, n_img, words_bf16, row_meta, row_wnorm, n_tiles, feature_norm, agg,
                   lambda_softmax, lambda_lse, scores, ld_scores, nullptr, 0, 0, stream);
}

extern "C" int itr_scan_t2i_affinity_debug(const uint16_t* images_bf16, int n_img, const uint16_t* words_bf16, int n_tiles,
                                           int word_tile, int image_tile, float* out, void* stream) {
  ITR_REQUIRE(images_bf16 && words_bf16 && out, "itr_scan_t2i_affinity_debug: null pointer");
  ITR_REQUIRE(word_tile >= 0 && word_tile < n_tiles && image_tile >= 0 && image_tile * IMGS < n_img,
              "itr_scan_t2i_affinity_debug: tile index out of range");
  // aux loads still run; point them at the operand buffers (contents unused in debug mode)
  return launch_tc(images_bf16, reinterpret_cast<const float*>(images_bf16), n_img, words_bf16,
                   reinterpret_cast<const int32_t*>(words_bf16), reinterpret_cast<const float*>(words_bf16), n_tiles,
                   ITR_NORM_CLIPPED_L2, ITR_AGG_SUM, 1.f, 1.f, out, 0, out, word_tile, image_tile, stream);
}
