// SCAN text-to-image cross-attention scores on Blackwell tensor cores, CTA-PAIR form (sm_100a only).
//
// Same mathematics and the same epilogue as scan_t2i_tc.cu (read its header first); what changes is the MMA:
// two CTAs on the two SMs of a TPC form a cluster and run ONE `tcgen05.mma.cta_group::2` per K step with
// M = 256: CTA r of the pair supplies word tile 2*mp + r (128 rows of A) and HALF of the image tile (72 of the
// 144 region rows of B, i.e. two of its four images); the hardware exchanges the B halves, and each CTA ends up
// with its own 128 x 144 accumulator in its own tensor memory.  Why (profiles/r02/mma2_microbench.txt):
//   * a tcgen05.mma costs its issuer ~115 clk whatever its size (max(115, N/2 + 45) in SS mode) and the pair form
//     costs the same 117 clk for twice the work, so the per-SM issue work halves (64 main + 12 Gram instructions
//     per TWO items) and the MMA side drops from ~8.7K to ~4K clk per item, below the epilogue;
//   * each SM stages 25 KB per K block instead of 34 KB (its B half is 9 KB): the L2 -> SMEM operand stream of the
//     image tiles halves and the ring holds 7 stages instead of 5 in the same shared memory.
// Replaces xattn_score_t2i + func_attention + cosine_similarity (itr/modalmodule/Objectives.py:329-372, 421-476,
// 10-15) for raw_feature_norm in {clipped_l2norm, l2norm} and every agg_func.
//
// Roles per CTA (20 warps, as before): warp 0 TMA producer (both CTAs; the peer's loads complete on the LEADER's
// `full` barrier), warp 1 main MMA issuer (leader only), warp 2 TMEM allocator (both) + Gram MMA issuer (leader
// only, also cta_group::2: M = 256, N = 48, each CTA holds 24 of the 48 rows of every Gram pack), warp 3 aux
// loader (both), warps 4-19 epilogue (both).  Barriers that gate the leader's issuers collect arrivals from both
// CTAs (the peer arrives through the cluster-mapped address); barriers the tensor pipe signals are multicast to
// both CTAs with one tcgen05.commit.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "tc2_common.cuh"

namespace itr {
namespace tc2 {

constexpr int STAGES = 5;
constexpr int GRAM_N = 48;
constexpr int GRAM_BYTES = ITR_GRAM_BYTES;     // 4752 per image in global memory (fp16 48x48 + 36 fp32)
constexpr int GH_BYTES = GRAM_N * GRAM_N;      // 2304 = 24 rows x 48 k x 2 bytes: this CTA's half of one Gram pack
constexpr int G_LBO = 128, G_SBO = 768;
constexpr int AUX_GRAM = IMGS * GH_BYTES;      // 9216
constexpr int AUX_META = BLOCK_M * 16;         // 2048
constexpr int AUX_WNORM = BLOCK_M * 4;         // 512
constexpr int AUX_BYTES = AUX_GRAM + AUX_META + AUX_WNORM;   // 11776
constexpr int XCH_FLOATS = 4 * 4 * 40;
// per-epilogue-warp scratch for the l2norm denominators: a^2 of the warp's 32 word rows x 36 regions.  Pitch 36 floats:
// the 8-byte row stores are 2-way bank conflicted (a conflict-free pitch of 38 does not leave room for 5 stages).
constexpr int SQ_PITCH = 36;
constexpr int SQ_WARP_FLOATS = 32 * SQ_PITCH;
constexpr int U_BASE = 2 * ACC_PITCH;          // four Gram products at 288 + 48 g
__host__ __device__ constexpr int park_col(int g) { return g == 0 ? 0 : 16 * ((36 * g + 15) / 16); }   // 0, 48, 80, 112

constexpr int SMEM_STAGES = 0;
constexpr int SMEM_AUX = SMEM_STAGES + STAGES * STAGE_BYTES;
constexpr int SMEM_XCH = SMEM_AUX + 2 * AUX_BYTES;
constexpr int SMEM_SQ = SMEM_XCH + XCH_FLOATS * 4;
constexpr int SMEM_BARS = SMEM_SQ + NUM_EPI_WARPS * SQ_WARP_FLOATS * 4;
constexpr int NUM_BARS = 2 * STAGES + 18;
constexpr int SMEM_TMEMPTR = SMEM_BARS + NUM_BARS * 8;
constexpr int SMEM_BYTES = SMEM_TMEMPTR + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;
static_assert(SMEM_ALLOC <= 232448, "shared memory budget");
static_assert(STAGE_BYTES % 1024 == 0 && SMEM_AUX % 16 == 0 && SMEM_SQ % 16 == 0 && SMEM_BARS % 8 == 0, "alignment");

// kind::f16 instruction descriptors, M = 256 across the CTA pair: D=f32, A=B=bf16 (main) / fp16 (Gram), K-major
constexpr uint32_t IDESC_GRAM = (1u << 4) | ((uint32_t)(GRAM_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

struct Params {
  const uint8_t* gram_pack;    // [n_img][GRAM_BYTES]
  const int4* row_meta;        // [n_wt*128]
  const float* row_wnorm;      // [n_wt*128]
  int n_img, n_wt, n_wp, n_it; // n_wp = ceil(n_wt / 2): the peer's tile of the last pair may not exist
  int clipped, agg;
  float c_sm, c_lse, inv_lse;
  float* scores; long long ld; // MODE_SCORES: required; MODE_COUNT: optional (NULL = the matrix is never written)
  long long* prof;             // optional [cluster][2][16] cycle counters (PROF instantiation)
  // fused evaluation (i2t / t2i ranking, evaluation.py:156-222): caption c of this launch is global caption cap_offset + c,
  // whose ground-truth image is (cap_offset + c) / cpi
  const int4* items; int n_items;          // MODE_GT: (leader's word tile, peer's word tile, image tile, 0) items that hold ground-truth pairs
  int cap_offset, cpi;
  int img_offset;                          // MODE_COUNT: global index of image 0 of this launch (arg-max keys of the columns)
  float* thr_col;                          // [n_cap]  MODE_GT: written; MODE_COUNT: read (NaN = no ground-truth image here)
  unsigned int* thr_row_key;               // [n_img]  MODE_GT: atomicMax of the orderable key of the ground-truth scores
  const float* thr_row;                    // [n_img]  MODE_COUNT
  int* cnt_col; int* cnt_row;              // MODE_COUNT: += #scores strictly above the threshold
  unsigned long long* best_col; unsigned long long* best_row;   // MODE_COUNT: atomicMax of (orderable score << 32 | ~index)
};
constexpr int MODE_SCORES = 0, MODE_GT = 1, MODE_COUNT = 2;

template <bool MAXOP>
__device__ __forceinline__ float seg_total(float x, const bool (&p)[5], int seg_hi) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    float y = __shfl_up_sync(0xffffffffu, x, 1 << s);
    if (p[s]) x = MAXOP ? fmaxf(x, y) : x + y;
  }
  return __shfl_sync(0xffffffffu, x, seg_hi);
}

struct Carry {
  float P, D, wnorm;
  float thr_c, thr_r;          // MODE_COUNT: thresholds of the row's caption / the group's image, fetched an item ahead
  int cap, seg, n_words, img, b;
  bool valid, img_ok, live;
};

template <bool PROF, bool CLIPPED, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
scan_t2i_tc2_kernel(const __grid_constant__ CUtensorMap map_words, const __grid_constant__ CUtensorMap map_imgs, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  // logical warp index: 0-3 control, 4-19 epilogue; the control warpgroup is the LAST one physically (measured best,
  // scan_t2i_tc.cu).  The shift is a multiple of 4: TMEM lane quarter and warpgroup alignment are preserved.
  const int warp = (int)((threadIdx.x >> 5) + EPI_WARP0) % (NUM_THREADS / 32);
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
  auto eready_bar = [&](int g) { return bar0 + 8u * (2 * STAGES + 8 + g); };     // leader: 8 epilogue warps
  auto uready_bar = [&](int g) { return bar0 + 8u * (2 * STAGES + 12 + g); };    // both (multicast commit)
  auto gfree_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 16 + b); };     // leader (local commit)
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + SMEM_TMEMPTR);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_words) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_imgs) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1); mbar_init(loaded_bar(b), 2 * NUM_EPI_WARPS); mbar_init(gfree_bar(b), 1);
      mbar_init(afull_bar(b), 1); mbar_init(aempty_bar(b), NUM_EPI_WARPS);
    }
    for (int g = 0; g < IMGS; ++g) { mbar_init(eready_bar(g), 8); mbar_init(uready_bar(g), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SMEM_TMEMPTR), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the peer's barriers are initialised before anybody arrives on them remotely
  tc_fence_after();
  // The pair owns both SMs and asks for all 512 columns, so the allocation starts at column 0 (the allocator blocks until
  // they are free).  Only the profiling build reads the address back: compute-sanitizer's racecheck pairs that read with
  // the allocator's hardware write into the peer's copy of the slot (same value, ordered by the barriers above).
  if (PROF && *tmem_ptr_smem != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;

  constexpr bool prof_on = PROF;
  using Schedule = ScheduleT<MODE == MODE_GT>;
  using ItemIter = ItemIterT<MODE == MODE_GT>;
  const Schedule sched(p.n_wp, p.n_it, MODE == MODE_GT ? p.items : nullptr, MODE == MODE_GT ? p.n_items : 0);
  const int first = (int)(blockIdx.x >> 1);
  const int step = (int)(gridDim.x >> 1);
  long long* prof = PROF ? p.prof + ((size_t)(blockIdx.x >> 1) * 2 + rank) * 16 : nullptr;

  if (warp < EPI_WARP0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ==============================
    int stage = 0; uint32_t phase = 0;
    long long w_empty = 0; const long long t_begin = prof_on ? clock64() : 0;
    for (ItemIter item(sched, first, step); item.valid(); item.next()) {
      const int row_w = item.tile((int)rank) * BLOCK_M;                // rows past the tensor are zero-filled
      const int row_i = item.n * BLOCK_N + (int)rank * HALF_N;
#pragma unroll 1
      for (int kb = 0; kb < K_BLOCKS; ++kb) {
        mbar_wait_sleep_t(empty_bar(stage), phase ^ 1, w_empty, prof_on);
        const uint32_t sa = sbase + SMEM_STAGES + stage * STAGE_BYTES, fb = full_bar(stage);
        if (elect_one()) {
          // one arrival per phase: the leader's, which expects the bytes of BOTH CTAs.  The peer cannot run a phase
          // ahead: it refills stage s only after the MMAs that consumed it completed (its own `empty`, multicast by
          // the leader), i.e. after the leader's `full` phase for that stage is over.
          if (leader) mbar_expect_tx(fb, 2 * STAGE_BYTES);
          tma_load_2d_pair(sa, &map_words, fb, kb * BLOCK_K, row_w);
          tma_load_2d_pair(sa + A_BYTES, &map_imgs, fb, kb * BLOCK_K, row_i);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (prof_on && lane == 0) { prof[0] = clock64() - t_begin; prof[1] = w_empty; }
  } else if (warp == 1) {
    // =============================== main MMA issuer (leader only) =========================
    if (leader) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      long long w_tempty = 0, w_full = 0; const long long t_begin = prof_on ? clock64() : 0;
      const uint64_t adesc0 = umma_desc_sw128(sbase + SMEM_STAGES);
      const uint64_t bdesc0 = umma_desc_sw128(sbase + SMEM_STAGES + A_BYTES);
      for (ItemIter item(sched, first, step); item.valid(); item.next(), ++it) {
        // accumulator b = it & 1 (in both CTAs) is reusable once every epilogue warp of BOTH CTAs has item it-2 in
        // registers and the Gram MMAs of item it-2, which read the numerators parked in it, have completed
        const int ab = it & 1;
        const uint32_t tacc = tmem_base + ab * ACC_PITCH;
        mbar_wait_sleep_t(loaded_bar(ab), ((it >> 1) & 1) ^ 1, w_tempty, prof_on);
        mbar_wait_sleep_t(gfree_bar(ab), ((it >> 1) & 1) ^ 1, w_tempty, prof_on);
        tc_fence_after();
#pragma unroll 1
        for (int kb = 0; kb < K_BLOCKS; ++kb) {
          mbar_wait_sleep_t(full_bar(stage), phase, w_full, prof_on);
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
      if (prof_on && lane == 0) { prof[2] = clock64() - t_begin; prof[3] = w_tempty; prof[4] = w_full; prof[5] = it; }
    }
  } else if (warp == 2) {
    // =============================== Gram MMA issuer (leader only) =========================
    // U_g = e_g (G_g - I) for the image's column group in BOTH CTAs: A = the parked fp16 numerators (each CTA's own
    // tensor memory), B = the image's Gram pack, rows [0,24) from the leader's copy and [24,48) from the peer's
    if (leader) {
      int it = 0;
      uint32_t used[IMGS] = {0u, 0u, 0u, 0u};
      for (ItemIter item(sched, first, step); item.valid(); item.next(), ++it) {
        const int n = item.n;
        const int b = it & 1;
        mbar_wait_sleep(afull_bar(b), (it >> 1) & 1);      // the leader's half; the peer's is ordered by its eready arrivals
        const uint32_t aux = sbase + SMEM_AUX + b * AUX_BYTES;
#pragma unroll
        for (int g = 0; g < IMGS; ++g) {
          if (n * IMGS + g >= p.n_img) continue;
          mbar_wait_sleep(eready_bar(g), used[g]++ & 1);
          tc_fence_after();
          const uint32_t te = tmem_base + b * ACC_PITCH + park_col(g);
          const uint32_t tu = tmem_base + U_BASE + g * GRAM_N;
          const uint64_t gdesc = umma_desc_nosw(aux + g * GH_BYTES, G_LBO, G_SBO);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GRAM_N / UMMA_K; ++k)
              umma2_f16_ts(tu, te + 8 * k, gdesc + (uint64_t)((2 * G_LBO * k) >> 4), IDESC_GRAM, k != 0);
            umma2_commit_both(uready_bar(g));
          }
          __syncwarp();
        }
        if (elect_one()) umma2_commit_local(gfree_bar(b));
        __syncwarp();
      }
    }
  } else {
    // =============================== aux loader (both CTAs) ================================
    int it = 0;
    for (ItemIter item(sched, first, step); item.valid(); item.next(), ++it) {
      const int m = item.tile((int)rank), n = item.n;
      const int b = it & 1;
      mbar_wait_sleep(aempty_bar(b), ((it >> 1) & 1) ^ 1);
      if (elect_one()) {
        const int n_valid = min(IMGS, p.n_img - n * IMGS);
        const bool word_ok = m < p.n_wt;
        const uint32_t aux = sbase + SMEM_AUX + b * AUX_BYTES;
        mbar_expect_tx(afull_bar(b), n_valid * GH_BYTES + (word_ok ? AUX_META + AUX_WNORM : 0));
        for (int g = 0; g < n_valid; ++g)
          bulk_load(aux + g * GH_BYTES, p.gram_pack + (size_t)(n * IMGS + g) * GRAM_BYTES + rank * GH_BYTES, GH_BYTES, afull_bar(b));
        if (word_ok) {
          bulk_load(aux + AUX_GRAM, p.row_meta + (size_t)m * BLOCK_M, AUX_META, afull_bar(b));
          bulk_load(aux + AUX_GRAM + AUX_META, p.row_wnorm + (size_t)m * BLOCK_M, AUX_WNORM, afull_bar(b));
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
    const uint32_t tu = tmem_base + U_BASE + g * GRAM_N + lane_sel;
    const uint32_t tpark0 = tmem_base + park_col(g) + lane_sel;
    float* xch = reinterpret_cast<float*>(smem + SMEM_XCH) + g * 4 * 40;
    const uint32_t sq = sbase + SMEM_SQ + (uint32_t)(warp - EPI_WARP0) * (SQ_WARP_FLOATS * 4);      // shared-space address
    uint32_t used = 0u;
    long long w_tfull = 0, w_afull = 0, w_uready = 0; const long long t_begin = prof_on ? clock64() : 0;
    Carry c;
    uint32_t hvp[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) hvp[k] = 0u;
    c.live = false; c.valid = false; c.img_ok = false; c.P = c.D = c.wnorm = 0.f; c.thr_c = c.thr_r = 0.f; c.cap = -1; c.seg = 0; c.n_words = 0; c.img = 0; c.b = 0;

    auto phase_b = [&]() {
      if (c.img_ok) {
        // the Gram issuer commits uready[g] for every existing image, whether or not this CTA's word tile exists
        mbar_wait_sleep_t(uready_bar(g), used++ & 1, w_uready, prof_on);
        tc_fence_after();
      }
      if (c.valid) {
        const int seg_lo = c.seg & 0xff, seg_hi = (c.seg >> 8) & 0xff;
        const bool long_tile = (c.seg >> 16) & 1;
        float q0 = 0.f, q1 = 0.f, Zsum;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t Uh[18];
          TMEM_LD_X16U(tu + 18 * h, Uh, 0);
          TMEM_LD_X2U(tu + 18 * h + 16, Uh, 16);
          if (h == 1) TMEM_LD_X1F(tu + 36, Zsum);
          tmem_ld_wait();
#pragma unroll
          for (int cidx = 0; cidx < 9; ++cidx) {
            float2 ef = unpack_f16x2(hvp[9 * h + cidx]);
            q0 = fmaf(ef.x, __uint_as_float(Uh[2 * cidx]), q0); q1 = fmaf(ef.y, __uint_as_float(Uh[2 * cidx + 1]), q1);
          }
        }
        const float Qf = c.D + (q0 + q1);
        float sq;                                   // approximate sqrt / division: 2^-22 relative, far inside the bf16 contract
        asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(fmaxf(Qf, 0.f)));
        const float rj = __fdividef(c.P, fmaxf(c.wnorm * sq, 1e-8f * Zsum));
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
        if (p.agg == ITR_AGG_MEAN) tot = __fdividef(tot, (float)c.n_words);
        const bool writer = long_tile ? (row == 0) : (lane == seg_lo);
        if (writer && c.cap >= 0) {
          if (MODE == MODE_SCORES) {
            p.scores[(size_t)c.img * p.ld + c.cap] = tot;
          } else if (MODE == MODE_GT) {
            // ground-truth pre-pass: the same arithmetic on the same packed rows as the counting pass, so the
            // thresholds are bit-identical to the scores it will compare them with
            if ((p.cap_offset + c.cap) / p.cpi == c.img) {
              p.thr_col[c.cap] = tot;
              atomicMax(p.thr_row_key + c.img, orderable(tot));
            }
          } else {
            if (p.scores) p.scores[(size_t)c.img * p.ld + c.cap] = tot;
            // rank = #scores strictly above the ground-truth score; a score can be the arg-max of its row / column only
            // if it is not below that score, so the packed keys are built on the rare path only
            const float tc = c.thr_c, tr = c.thr_r;
            if (!(tot < tc)) {
              if (tot > tc) atomicAdd(p.cnt_col + c.cap, 1);
              atomicMax(p.best_col + c.cap, ((unsigned long long)orderable(tot) << 32) | (unsigned)(~(unsigned)(p.img_offset + c.img)));
            }
            if (!(tot < tr)) {
              if (tot > tr) atomicAdd(p.cnt_row + c.img, 1);
              atomicMax(p.best_row + c.img, ((unsigned long long)orderable(tot) << 32) | (unsigned)(~(unsigned)(p.cap_offset + c.cap)));
            }
          }
        }
      }
      if (c.live) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(aempty_bar(c.b));
      }
    };

    int it = 0;
    for (ItemIter item(sched, first, step); item.valid(); item.next(), ++it) {
      const int n = item.n;
      const int m = item.tile((int)rank);
      const bool word_ok = m < p.n_wt;
      const int b = it & 1;
      mbar_wait_sleep_t(afull_bar(b), (it >> 1) & 1, w_afull, prof_on);
      const uint8_t* aux = smem + SMEM_AUX + b * AUX_BYTES;
      int4 meta = make_int4(-1, 0, lane | (lane << 8), 0);
      float wnorm = 0.f;
      if (word_ok) {
        meta = reinterpret_cast<const int4*>(aux + AUX_GRAM)[row];
        wnorm = reinterpret_cast<const float*>(aux + AUX_GRAM + AUX_META)[row];
      }
      const int seg_lo = meta.z & 0xff, seg_hi = (meta.z >> 8) & 0xff;
      const bool long_tile = (meta.z >> 16) & 1;
      const int img = n * IMGS + g;
      const bool img_ok = img < p.n_img;
      const bool valid = img_ok && word_ok;
      // MODE_COUNT: the two thresholds this row's score will be compared with, requested now and consumed in phase B one
      // item later (a dependent global load inside phase B costs every warp an L2 round trip per item)
      float thr_c = 0.f, thr_r = 0.f;
      if (MODE == MODE_COUNT && valid && meta.x >= 0) { thr_c = p.thr_col[meta.x]; thr_r = p.thr_row[img]; }

      const uint32_t tacc = tmem_base + b * ACC_PITCH + lane_sel;
      mbar_wait_sleep_t(tfull_bar(b), (it >> 1) & 1, w_tfull, prof_on);
      tc_fence_after();
      float A[R];
      TMEM_LD_X32(tacc + g * R, A, 0);
      TMEM_LD_X4(tacc + g * R + 32, A, 32);

      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(loaded_bar(b), leader);

      // ---- A(t), first half: l2norm denominators.  S[c][k] = sum over the caption's words of a^2: every lane writes the
      // a^2 of its word row to the warp's scratch, then lanes 0..17 walk the 32 rows, two regions each, restarting at
      // every caption end (a caption never straddles a warp) and leaving lambda*log2e / sqrt(S) -- the softmax scale the
      // whole caption shares -- in the caption's LAST row: ~175 instructions and ~70 shared-memory operations instead of
      // 36 x (6 shuffles + 5 adds) of segmented warp scans, and one rsqrt per (caption, region) instead of per (word, region).
      // lambda / (sqrt(S) + 1e-8), the reference's l2norm + softmax temperature (utils.py:11-15), as t / (1 + 1e-8 r) with
      // r = rsqrt(S), t = lambda r: computed per (caption, region), so it costs nothing to take the exact form always --
      // and the result of a caption does not depend on which captions share its warp.
      const float eps_c = 1e-8f;
      if (valid && !long_tile) {
        const uint32_t myrow = sq + (uint32_t)lane * (SQ_PITCH * 4);
#pragma unroll
        for (int k = 0; k < R; k += 2) {
          const float a0 = CLIPPED ? fmaxf(A[k], 0.1f * A[k]) : A[k];
          const float a1 = CLIPPED ? fmaxf(A[k + 1], 0.1f * A[k + 1]) : A[k + 1];
          sts_f2(myrow + 4 * k, a0 * a0, a1 * a1);
        }
        __syncwarp();
        const uint32_t endmask = __ballot_sync(0xffffffffu, lane == seg_hi);
        if (lane < R / 2) {
          const uint32_t col = sq + 8u * (uint32_t)lane;
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll 1
          for (int j0 = 0; j0 < 32; j0 += 8) {        // eight independent loads in flight, then the dependent adds
            float2 v[8];                              // (rolled over the four row blocks: the epilogue is i-cache bound)
            const uint32_t blk = col + (uint32_t)(j0 * SQ_PITCH * 4);
            const uint32_t ends = endmask >> j0;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = lds_f2(blk + (uint32_t)(j * SQ_PITCH * 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc.x += v[j].x; acc.y += v[j].y;
              if ((ends >> j) & 1u) {
                const float r0 = rsqf(fmaxf(acc.x, 1e-36f)), r1 = rsqf(fmaxf(acc.y, 1e-36f));
                sts_f2(blk + (uint32_t)(j * SQ_PITCH * 4), __fdividef(p.c_sm * r0, fmaf(eps_c, r0, 1.0f)),
                       __fdividef(p.c_sm * r1, fmaf(eps_c, r1, 1.0f)));
                acc = make_float2(0.f, 0.f);
              }
            }
          }
        }
        __syncwarp();                                // the scale stores are visible to the whole warp
      }

      // ---- B(t-1): finish the previous item.  Its Gram product was issued when the last of the eight warps parked,
      // ~2K clk ago by now, behind whatever main MMAs were queued: nobody waits for it here.
      phase_b();

      uint32_t hv[18];
#pragma unroll
      for (int k = 0; k < 18; ++k) hv[k] = 0u;
      float P = 0.f, Dd = 0.f;
      if (valid) {
        const float shift = -fabsf(p.c_sm);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float E[R / 2];                       // softmax scales lambda*log2e/sqrt(S[c][k]) of this half's regions
          if (!long_tile) {
            const uint32_t tot = sq + (uint32_t)(seg_hi * SQ_PITCH + 18 * h) * 4u;
#pragma unroll
            for (int k = 0; k < R / 2; k += 2) {
              const float2 t = lds_f2(tot + 4 * k);
              E[k] = t.x; E[k + 1] = t.y;
            }
          } else {
#pragma unroll
            for (int k = 0; k < R / 2; ++k) {
              const float a = CLIPPED ? fmaxf(A[18 * h + k], 0.1f * A[18 * h + k]) : A[18 * h + k];
              E[k] = warp_sum(a * a);
            }
            named_bar_sync(1 + g, 128);
            if (lane == 0) {
#pragma unroll
              for (int k = 0; k < R / 2; ++k) xch[q * 40 + k] = E[k];
            }
            named_bar_sync(1 + g, 128);
#pragma unroll
            for (int k = 0; k < R / 2; ++k) {
              const float r = rsqf(fmaxf((xch[k] + xch[40 + k]) + (xch[80 + k] + xch[120 + k]), 1e-36f));
              E[k] = __fdividef(p.c_sm * r, fmaf(eps_c, r, 1.0f));
            }
          }
#pragma unroll
          for (int k = 0; k < R / 2; ++k) {
            const float raw = A[18 * h + k];
            const float a = CLIPPED ? fmaxf(raw, 0.1f * raw) : raw;
            const float e = ex2f(fmaf(a, E[k], shift));
            E[k] = e; P = fmaf(e, raw, P); Dd = fmaf(e, e, Dd);
          }
#pragma unroll
          for (int cidx = 0; cidx < 9; ++cidx) hv[9 * h + cidx] = pack_f16x2(E[2 * cidx], E[2 * cidx + 1]);
        }
        if (!long_tile) __syncwarp();           // every lane has read its caption's scales before the next item's stores
      }

      // park(t): e(t) as fp16 (K padded 36 -> 48 with zeros) in the group's own columns of the accumulator it came from,
      // wake the (leader's) Gram issuer.  An image that exists is signalled by BOTH CTAs, word tile or not.
      if (img_ok) {
        const uint32_t tpark = tpark0 + b * ACC_PITCH;
        uint32_t z[6] = {0u, 0u, 0u, 0u, 0u, 0u};
        TMEM_ST_X16(tpark, hv, 0);
        TMEM_ST_X2(tpark + 16, hv, 16);
        TMEM_ST_X4(tpark + 18, z, 0);
        TMEM_ST_X2(tpark + 22, z, 4);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(eready_bar(g), leader);
      }
#pragma unroll
      for (int k = 0; k < 18; ++k) hvp[k] = hv[k];
      c.P = P; c.D = Dd; c.wnorm = wnorm; c.cap = meta.x; c.seg = meta.z; c.n_words = meta.w;
      if (MODE == MODE_COUNT) { c.thr_c = thr_c; c.thr_r = thr_r; }
      c.img = img; c.b = b; c.valid = valid; c.img_ok = img_ok; c.live = true;
    }
    phase_b();
    if (prof_on && lane == 0 && q == 0) {
      long long* o = prof + 6 + g * 2;
      o[0] = w_tfull + w_afull; o[1] = w_uready;
      if (g == 0) { prof[14] = clock64() - t_begin; prof[15] = w_afull; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // nobody frees tensor memory or exits while the peer can still signal or read it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

template <bool PROF, bool CLIPPED, int MODE>
static int launch(const CUtensorMap& map_w, const CUtensorMap& map_i, const Params& p, cudaStream_t stream) {
  auto kern = scan_t2i_tc2_kernel<PROF, CLIPPED, MODE>;
  ITR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
  int dev = 0, sms = 0;
  ITR_CHECK_CUDA(cudaGetDevice(&dev));
  ITR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // how many CTA pairs the device keeps resident at once (a static round-robin schedule must not be split in waves)
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
  const long long units = MODE == MODE_GT ? (long long)p.n_items : (long long)((p.n_wp + BAND - 1) / BAND) * p.n_it;
  if (units <= 0) return ITR_OK;
  const int pairs = (int)(units < max_pairs ? units : max_pairs);
  kern<<<2 * pairs, NUM_THREADS, SMEM_ALLOC, stream>>>(map_w, map_i, p);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

static int fill_params(Params& p, CUtensorMap& map_w, CUtensorMap& map_i, const uint16_t* images_bf16, const void* gram_pack,
                       int n_img, const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm, int n_tiles,
                       int feature_norm, int agg, float lambda_softmax, float lambda_lse) {
  int rc = make_map(&map_w, words_bf16, (uint64_t)n_tiles * BLOCK_M, BLOCK_M);
  if (rc) return rc;
  rc = make_map(&map_i, images_bf16, (uint64_t)n_img * R, HALF_N);
  if (rc) return rc;
  p.gram_pack = reinterpret_cast<const uint8_t*>(gram_pack);
  p.row_meta = reinterpret_cast<const int4*>(row_meta);
  p.row_wnorm = row_wnorm;
  p.n_img = n_img; p.n_wt = n_tiles; p.n_wp = (n_tiles + 1) / 2; p.n_it = (n_img + IMGS - 1) / IMGS;
  p.clipped = (feature_norm == ITR_NORM_CLIPPED_L2); p.agg = agg;
  p.c_sm = lambda_softmax * 1.4426950408889634f;
  p.c_lse = lambda_lse * 1.4426950408889634f;
  p.inv_lse = 0.6931471805599453f / lambda_lse;
  if ((long long)p.n_wp * p.n_it >= (1ll << 31))
    return fail(ITR_ERR_INVALID, "itr_scan_t2i_*_bf16: %lld tile pairs exceed the 2^31 scheduler range; split the call", (long long)p.n_wp * p.n_it);
  return ITR_OK;
}

int launch_tc2(const uint16_t* images_bf16, const void* gram_pack, int n_img, const uint16_t* words_bf16,
               const int32_t* row_meta, const float* row_wnorm, int n_tiles, int feature_norm, int agg,
               float lambda_softmax, float lambda_lse, float* scores, int64_t ld_scores, void* stream, long long* prof) {
  CUtensorMap map_w, map_i;
  Params p{};
  int rc = fill_params(p, map_w, map_i, images_bf16, gram_pack, n_img, words_bf16, row_meta, row_wnorm, n_tiles, feature_norm, agg,
                       lambda_softmax, lambda_lse);
  if (rc) return rc;
  p.scores = scores; p.ld = ld_scores; p.prof = prof;
  cudaStream_t st = as_stream(stream);
  if (prof) return launch<true, true, MODE_SCORES>(map_w, map_i, p, st);       // the profile entry point always runs clipped_l2norm
  return p.clipped ? launch<false, true, MODE_SCORES>(map_w, map_i, p, st) : launch<false, false, MODE_SCORES>(map_w, map_i, p, st);
}

// thr_row keys (atomicMax of orderable(score); 0 = no ground-truth caption seen) -> float thresholds (-inf)
__global__ void unkey_thresholds_kernel(const unsigned int* __restrict__ key, float* __restrict__ thr_row, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const unsigned int u = key[i];
    thr_row[i] = u == 0u ? -INFINITY : __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
  }
}

// Ground-truth pre-pass of the fused evaluation: scores of the (image, caption) pairs with caption / cpi == image, on
// the listed items only.  thr_col[c] = that score (NaN where the caption's image is not in [0, n_img)), thr_row[i] = the
// best of image i's ground-truth captions among THIS launch's captions (-inf if none).
int launch_tc2_gt(const uint16_t* images_bf16, const void* gram_pack, int n_img, const uint16_t* words_bf16,
                  const int32_t* row_meta, const float* row_wnorm, int n_tiles, int n_cap, const int32_t* items, int n_items,
                  int feature_norm, int agg, float lambda_softmax, float lambda_lse, int cap_offset, int cpi,
                  float* thr_col, float* thr_row, void* stream) {
  CUtensorMap map_w, map_i;
  Params p{};
  int rc = fill_params(p, map_w, map_i, images_bf16, gram_pack, n_img, words_bf16, row_meta, row_wnorm, n_tiles, feature_norm, agg,
                       lambda_softmax, lambda_lse);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  p.items = reinterpret_cast<const int4*>(items); p.n_items = n_items;
  p.cap_offset = cap_offset; p.cpi = cpi;
  p.thr_col = thr_col;
  p.thr_row_key = reinterpret_cast<unsigned int*>(thr_row);        // keys first, converted in place below
  ITR_CHECK_CUDA(cudaMemsetAsync(thr_col, 0xFF, sizeof(float) * (size_t)n_cap, st));      // NaN
  ITR_CHECK_CUDA(cudaMemsetAsync(thr_row, 0, sizeof(float) * (size_t)n_img, st));
  rc = p.clipped ? launch<false, true, MODE_GT>(map_w, map_i, p, st) : launch<false, false, MODE_GT>(map_w, map_i, p, st);
  if (rc) return rc;
  unkey_thresholds_kernel<<<(n_img + 255) / 256, 256, 0, st>>>(p.thr_row_key, thr_row, n_img);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

// Counting pass of the fused evaluation: every score is compared with its row / column threshold as it is produced;
// the matrix itself is written only if `scores` is given.
int launch_tc2_count(const uint16_t* images_bf16, const void* gram_pack, int n_img, const uint16_t* words_bf16,
                     const int32_t* row_meta, const float* row_wnorm, int n_tiles, int n_cap, int feature_norm, int agg,
                     float lambda_softmax, float lambda_lse, int cap_offset, const float* thr_col, const float* thr_row,
                     float* scores, int64_t ld_scores, int32_t* cnt_row, int32_t* cnt_col, uint64_t* best_row, uint64_t* best_col,
                     int img_offset, int accumulate, void* stream) {
  CUtensorMap map_w, map_i;
  Params p{};
  int rc = fill_params(p, map_w, map_i, images_bf16, gram_pack, n_img, words_bf16, row_meta, row_wnorm, n_tiles, feature_norm, agg,
                       lambda_softmax, lambda_lse);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  p.scores = scores; p.ld = ld_scores;
  p.cap_offset = cap_offset; p.cpi = 1; p.img_offset = img_offset;
  p.thr_col = const_cast<float*>(thr_col); p.thr_row = thr_row;
  p.cnt_row = cnt_row; p.cnt_col = cnt_col;
  p.best_row = reinterpret_cast<unsigned long long*>(best_row); p.best_col = reinterpret_cast<unsigned long long*>(best_col);
  ITR_CHECK_CUDA(cudaMemsetAsync(cnt_row, 0, sizeof(int32_t) * (size_t)n_img, st));        // rows belong to this launch alone
  ITR_CHECK_CUDA(cudaMemsetAsync(best_row, 0, sizeof(uint64_t) * (size_t)n_img, st));
  if (!accumulate) {                                                                         // columns may collect several image ranges
    ITR_CHECK_CUDA(cudaMemsetAsync(cnt_col, 0, sizeof(int32_t) * (size_t)n_cap, st));
    ITR_CHECK_CUDA(cudaMemsetAsync(best_col, 0, sizeof(uint64_t) * (size_t)n_cap, st));
  }
  return p.clipped ? launch<false, true, MODE_COUNT>(map_w, map_i, p, st) : launch<false, false, MODE_COUNT>(map_w, map_i, p, st);
}

}  // namespace tc2
}  // namespace itr

using namespace itr;

static int require_sm100_tc2() {
  int dev = 0, major = 0;
  ITR_CHECK_CUDA(cudaGetDevice(&dev));
  ITR_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(ITR_ERR_UNSUPPORTED, "the tensor-core SCAN path needs an sm_100 device (found sm_%d0)", major);
  return ITR_OK;
}

#define ITR_TC2_COMMON_CHECKS(name)                                                                                              \
  ITR_REQUIRE(images_bf16 && gram_pack && words_bf16 && row_meta && row_wnorm, name ": null pointer");                            \
  ITR_REQUIRE(feature_norm == ITR_NORM_CLIPPED_L2 || feature_norm == ITR_NORM_L2,                                                \
              name ": raw_feature_norm %d is only available in the float32 path", feature_norm);                                 \
  ITR_REQUIRE(agg >= 0 && agg <= ITR_AGG_SUM, "unknown aggfunc: %d", agg);                                                       \
  ITR_REQUIRE(lambda_lse != 0.f || agg != ITR_AGG_LSE, name ": lambda_lse must be non-zero");                                    \
  ITR_REQUIRE(lambda_softmax > -80.f && lambda_softmax < 80.f, name ": |lambda_softmax| must be < 80");                          \
  ITR_REQUIRE(((uintptr_t)images_bf16 & 15) == 0 && ((uintptr_t)words_bf16 & 15) == 0 && ((uintptr_t)gram_pack & 15) == 0 &&     \
              ((uintptr_t)row_meta & 15) == 0 && ((uintptr_t)row_wnorm & 15) == 0, name ": buffers must be 16-byte aligned");    \
  ITR_REQUIRE(n_img >= 0 && n_tiles >= 0 && n_cap >= 0 && cap_offset >= 0, name ": bad shape")

extern "C" int itr_scan_t2i_gt_thresholds_bf16(const uint16_t* images_bf16, const void* gram_pack, int n_img,
                                               const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm,
                                               int n_tiles, int n_cap, const int32_t* items, int n_items, int feature_norm, int agg,
                                               float lambda_softmax, float lambda_lse, int cap_offset, int caps_per_img,
                                               float* thr_col, float* thr_row, void* stream) {
  ITR_TC2_COMMON_CHECKS("itr_scan_t2i_gt_thresholds_bf16");
  ITR_REQUIRE(thr_col && thr_row && (items || n_items == 0) && n_items >= 0 && caps_per_img >= 1, "itr_scan_t2i_gt_thresholds_bf16: bad arguments");
  if (n_img == 0 && n_cap == 0) return ITR_OK;
  int rc = require_sm100_tc2();
  if (rc) return rc;
  if (n_img == 0 || n_tiles == 0 || n_cap == 0) {          // nothing to score: thresholds keep their "none" values
    cudaStream_t st = as_stream(stream);
    if (n_cap) ITR_CHECK_CUDA(cudaMemsetAsync(thr_col, 0xFF, sizeof(float) * (size_t)n_cap, st));
    if (n_img) {
      ITR_CHECK_CUDA(cudaMemsetAsync(thr_row, 0, sizeof(float) * (size_t)n_img, st));
      tc2::unkey_thresholds_kernel<<<(n_img + 255) / 256, 256, 0, st>>>(reinterpret_cast<unsigned int*>(thr_row), thr_row, n_img);
      ITR_CHECK_LAUNCH();
    }
    return ITR_OK;
  }
  return tc2::launch_tc2_gt(images_bf16, gram_pack, n_img, words_bf16, row_meta, row_wnorm, n_tiles, n_cap, items, n_items,
                            feature_norm, agg, lambda_softmax, lambda_lse, cap_offset, caps_per_img, thr_col, thr_row, stream);
}

extern "C" int itr_scan_t2i_count_bf16(const uint16_t* images_bf16, const void* gram_pack, int n_img,
                                       const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm,
                                       int n_tiles, int n_cap, int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                                       int cap_offset, const float* thr_col, const float* thr_row, float* scores, int64_t ld_scores,
                                       int32_t* cnt_row, int32_t* cnt_col, uint64_t* best_row, uint64_t* best_col,
                                       int img_offset, int accumulate_columns, void* stream) {
  ITR_TC2_COMMON_CHECKS("itr_scan_t2i_count_bf16");
  ITR_REQUIRE(img_offset >= 0, "itr_scan_t2i_count_bf16: img_offset < 0");
  ITR_REQUIRE(thr_col && thr_row && cnt_row && cnt_col && best_row && best_col, "itr_scan_t2i_count_bf16: null pointer");
  ITR_REQUIRE(!scores || ld_scores >= n_cap, "itr_scan_t2i_count_bf16: ld_scores < n_cap");
  if (n_img == 0 && n_cap == 0) return ITR_OK;
  int rc = require_sm100_tc2();
  if (rc) return rc;
  if (n_img == 0 || n_tiles == 0 || n_cap == 0) {
    cudaStream_t st = as_stream(stream);
    if (n_img) { ITR_CHECK_CUDA(cudaMemsetAsync(cnt_row, 0, 4 * (size_t)n_img, st)); ITR_CHECK_CUDA(cudaMemsetAsync(best_row, 0, 8 * (size_t)n_img, st)); }
    if (n_cap && !accumulate_columns) { ITR_CHECK_CUDA(cudaMemsetAsync(cnt_col, 0, 4 * (size_t)n_cap, st)); ITR_CHECK_CUDA(cudaMemsetAsync(best_col, 0, 8 * (size_t)n_cap, st)); }
    return ITR_OK;
  }
  return tc2::launch_tc2_count(images_bf16, gram_pack, n_img, words_bf16, row_meta, row_wnorm, n_tiles, n_cap, feature_norm, agg,
                               lambda_softmax, lambda_lse, cap_offset, thr_col, thr_row, scores, ld_scores, cnt_row, cnt_col,
                               best_row, best_col, img_offset, accumulate_columns, stream);
}
