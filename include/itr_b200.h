/*
 * itr_b200 -- C ABI of the B200-native similarity / hinge-loss / Recall@K hot path.
 *
 * The reference (WangFei-2019/Image-text-Retrieval) is pure Python/PyTorch and has
 * no FFI or plugin registry; its hot path is a handful of
 * module-level Python functions.  This header is the native boundary underneath the
 * Python drop-ins for those functions: each entry point below names the reference
 * function (file:line, relative to the reference root) whose arithmetic it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every `const T*` / `T*` is a DEVICE pointer unless the
 *     parameter name ends in `_host`; `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream); all launches are asynchronous on `stream`.
 *   - matrices are row-major; `ld_*` are leading dimensions in ELEMENTS.
 *   - return value: ITR_OK, or an ITR_ERR_* code with a message in itr_last_error()
 *     (thread-local).  No entry point falls back to the CPU.
 *   - bf16 buffers are passed as `uint16_t*` (raw bit patterns).
 */
#ifndef ITR_B200_H_
#define ITR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ITR_OK               0
#define ITR_ERR_INVALID      1   /* bad argument / unsupported mode  -> ValueError   */
#define ITR_ERR_CUDA         2   /* CUDA runtime or driver error     -> RuntimeError */
#define ITR_ERR_UNSUPPORTED  3   /* device is not sm_100             -> RuntimeError */

/* config['cross_attn'], Objectives.py:64-71 */
#define ITR_T2I 0
#define ITR_I2T 1
/* config['raw_feature_norm'], Objectives.py:436-457 (the l1 modes raise NameError upstream, defect D4) */
#define ITR_NORM_CLIPPED_L2 0
#define ITR_NORM_L2         1
#define ITR_NORM_SOFTMAX    2
#define ITR_NORM_CLIPPED    3
#define ITR_NORM_NONE       4
/* config['agg_func'], Objectives.py:355-366 */
#define ITR_AGG_LSE  0
#define ITR_AGG_MEAN 1
#define ITR_AGG_MAX  2
#define ITR_AGG_SUM  3

#define ITR_REGIONS        36    /* precomp regions per image the tensor-core path is built for */
#define ITR_EMBED          1024  /* embed_size, itr/config.py:73 */
#define ITR_TILE_WORDS     128   /* rows of one packed word tile (= UMMA M) */
#define ITR_TILE_IMAGES    4     /* images per accumulator tile (UMMA N = 4*36 = 144) */
#define ITR_GRAM_BYTES     4752  /* bytes per image of the Gram pack: fp16 48x48 off-diagonal Gram in UMMA core-matrix order + 36 fp32 diagonal entries */
#define ITR_MAX_WORDS_F32  96    /* longest caption the fp32 kernels (validation mode, training backward) accept */

/* ---- library ------------------------------------------------------------------------- */
const char* itr_last_error(void);
int         itr_version(void);
/* 1 if `device` is compute capability 10.x, 0 if not, negative ITR_ERR_* on error */
int         itr_device_supported(int device);

/* ---- VSE++ scores: cosine_sim(im, s), Objectives.py:18-21 ------------------------------
 * scores[i, c] = sum_d im[i, d] * s[c, d]   (inputs already unit-norm; fp32 FMA, 1e-5 mode) */
int itr_cosine_scores_f32(const float* im, const float* s, int n_img, int n_cap, int d,
                          float* scores, int64_t ld_scores, void* stream);

/* ---- order-embedding scores: order_sim(im, s), Objectives.py:24-30 ------------------------
 * scores[i, c] = -sqrt(sum_d max(s[c, d] - im[i, d], 0)^2).  The backward returns the gradients autograd produces for
 * sum(scores * d_scores) (zero where the distance is exactly 0); either output may be NULL. */
int itr_order_scores_f32(const float* im, const float* s, int n_img, int n_cap, int d, float* scores, int64_t ld_scores,
                         void* stream);
int itr_order_backward_f32(const float* im, const float* s, const float* scores, int64_t ld_scores, const float* d_scores,
                           int64_t ld_dscores, int n_img, int n_cap, int d, float* d_im, float* d_s, void* stream);

/* ---- CAMERA multi-view scores: MultiViewMatching.forward, Fusionmodule.py:670-692 ----------
 * scores[i, c] = max_v imgs[i, v, :] . caps[c, :]; argmax (optional, (n_img, n_cap) int32) records the winning view
 * (first on ties), which is all the backward needs.  workspace: n_img * n_views * n_cap floats of device scratch. */
int itr_multiview_scores_f32(const float* imgs, const float* caps, int n_img, int n_views, int n_cap, int d,
                             float* workspace, float* scores, int64_t ld_scores, int32_t* argmax, void* stream);
int itr_multiview_backward_f32(const float* imgs, const float* caps, int n_img, int n_views, int n_cap, int d,
                               const float* d_scores, int64_t ld_dscores, const int32_t* argmax, float* workspace,
                               float* d_imgs, float* d_caps, void* stream);

/* ---- SCAN scores, float32 validation mode ---------------------------------------------
 * xattn_score_t2i / xattn_score_i2t + func_attention + cosine_similarity,
 * Objectives.py:329-372, 376-417, 421-476, 10-15; l2norm utils.py:11-15.
 * All five working raw_feature_norm modes x four agg_func x both directions; 1 to 36 regions per image, any embed size.
 * images (n_img, n_regions, d); captions (n_cap, lmax, d) zero padded; cap_lens (n_cap).
 * gram: (n_img, n_regions, n_regions) from itr_region_gram_f32 (needed for ITR_T2I; may be NULL for ITR_I2T). */
int itr_region_gram_f32(const float* images, int n_img, int n_regions, int d, float* gram, void* stream);
int itr_scan_scores_f32(const float* images, const float* gram, const float* captions, const int32_t* cap_lens,
                        int n_img, int n_regions, int n_cap, int lmax, int d,
                        int cross_attn, int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                        float* scores, int64_t ld_scores, void* stream);

/* ---- SCAN training backward, float32 --------------------------------------------------
 * Replaces autograd through xattn_score_t2i / xattn_score_i2t when ContrastiveLoss is back-propagated
 * (Models.py:219-222, Objectives.py:76-115, 329-476): given d_scores = dLoss/dScores (n_img, n_cap) it returns
 * dLoss/dImages and dLoss/dCaptions for the same modes as itr_scan_scores_f32.  The affinity tile is recomputed;
 * nothing is saved by the forward call.
 *   n_words      sum(cap_lens);  sum_len_sq = sum(cap_lens^2)  (host-known; they size the workspace)
 *   d_images     (n_img, n_regions, d)  overwritten
 *   d_captions   (n_cap, lmax, d)       ACCUMULATED into (zero it before the first call; lets the caller split
 *                                       the images into chunks); rows of padding words are left untouched
 *   workspace    device scratch of itr_scan_backward_workspace_f32(...) bytes, 16-byte aligned
 * All pointers are device pointers; launches are asynchronous on `stream`. */
int64_t itr_scan_backward_workspace_f32(int n_img, int n_regions, int n_cap, int64_t n_words, int64_t sum_len_sq,
                                        int cross_attn);
int itr_scan_backward_f32(const float* images, const float* gram, const float* captions, const int32_t* cap_lens,
                          int n_img, int n_regions, int n_cap, int lmax, int d, int64_t n_words, int64_t sum_len_sq,
                          int cross_attn, int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                          const float* d_scores, int64_t ld_dscores, float* d_images, float* d_captions,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* ---- SCAN t2i scores, tcgen05 tensor-core path (bf16 inputs, fp32 accumulate) ----------
 * Same function as above for cross_attn = t2i, raw_feature_norm in {clipped_l2norm, l2norm}.
 *
 * 1. itr_scan_plan_words (HOST, no CUDA): bin-packs captions into 128-row word tiles so that
 *    no caption of <= 32 words straddles a 32-row quarter; longer captions get a tile of
 *    their own.  Output, per packed row: row_meta[4*row + {0,1,2,3}] =
 *      { caption id (-1 = padding), word index, seg_lo | seg_hi<<8 | long_tile<<16, caption length }.
 *    Returns the number of tiles through *n_tiles; arrays must hold itr_scan_plan_max_tiles() tiles.
 * 2. itr_scan_pack_words_bf16: gathers the words (captions may live in device memory or in
 *    pinned host memory mapped into the device address space), rounds to bf16 and writes each
 *    row's L2 norm (of the rounded values).
 * 3. itr_scan_prep_images_bf16: rounds regions to bf16 and writes, per image, the Gram pack of the
 *    rounded regions (ITR_GRAM_BYTES): G = V V^T with the off-diagonal part in fp16, laid out as the
 *    SMEM B operand of a 128x48x48 tcgen05.mma, and the diagonal in fp32.
 * 4. itr_scan_t2i_scores_bf16: persistent TMA -> tcgen05.mma -> TMEM epilogue kernel; two CTAs per cluster run one
 *    tcgen05.mma.cta_group::2 (M = 256) per K step (csrc/scan_t2i_tc2.cu).  ITR_B200_SCORE_KERNEL=single in the
 *    environment selects the one-CTA kernel of csrc/scan_t2i_tc.cu instead (A/B measurements).
 */
int itr_scan_plan_max_tiles(const int32_t* cap_lens_host, int n_cap);
int itr_scan_plan_words(const int32_t* cap_lens_host, int n_cap, int32_t* row_meta_host, int* n_tiles);
int itr_scan_pack_words_bf16(const float* captions, int n_cap, int lmax, int d,
                             const int32_t* row_meta, int n_tiles,
                             uint16_t* words_bf16, float* row_wnorm, void* stream);
int itr_scan_prep_images_bf16(const float* images, int n_img, int n_regions, int d,
                              uint16_t* images_bf16, void* gram_pack, void* stream);
int itr_scan_t2i_scores_bf16(const uint16_t* images_bf16, const void* gram_pack, int n_img,
                             const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm,
                             int n_tiles, int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                             float* scores, int64_t ld_scores, void* stream);
/* ---- SCAN scores, generic two-phase tensor-core path (both directions, all raw_feature_norm modes) --------
 * Phase 1, itr_scan_affinity_bf16: the same TMA + tcgen05 main loop; its epilogue only copies the raw region-word
 *   affinities A = V W^T (bf16 inputs, fp32 accumulate) to `affinity[n_tiles][n_img][128][36]` (fp32).
 * Phase 2, itr_scan_epilogue_f32: the reference's func_attention / cosine / aggregation (Objectives.py:421-476, 10-15,
 *   355-366) in fp32 on those affinities, one block per (caption, 4 images).  Inputs: cap_row0[c] = packed row of the
 *   caption's first word (words of a caption are consecutive rows), region_norm[n_img][36] = |v_k|,
 *   region_gram[n_img][36][36] (t2i) or the packed word Grams from itr_scan_caption_gram_f32 (i2t; gram_off[c] =
 *   sum_{c' < c} len_c'^2).  cap_ids[n_ids] selects the captions of this launch (all of length <= max_len); the
 *   shared-memory tile of a block is sized by max_len, so the caller launches once per length class and chunks over
 *   images to bound the affinity buffer. */
int itr_scan_affinity_bf16(const uint16_t* images_bf16, int n_img, const uint16_t* words_bf16, int n_tiles,
                           float* affinity, void* stream);
int itr_scan_caption_gram_f32(const uint16_t* words_bf16, const int32_t* cap_row0, const int32_t* cap_lens,
                              const int64_t* gram_off, int n_cap, int d, float* gram, void* stream);
int itr_scan_epilogue_f32(const float* affinity, int n_img, const int32_t* cap_row0, const int32_t* cap_lens,
                          const int32_t* cap_ids, int n_ids, int max_len, const float* row_wnorm, const float* region_norm,
                          const float* region_gram, const float* cap_gram, const int64_t* gram_off,
                          int cross_attn, int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                          float* scores, int64_t ld_scores, void* stream);

/* ---- SCAN i2t scores, fused on the CTA-pair tcgen05 main loop (csrc/scan_i2t_tc2.cu) --------------------------------
 * xattn_score_i2t (Objectives.py:376-417) for raw_feature_norm in {clipped_l2norm, l2norm}, every agg_func.  Operands as
 * for itr_scan_t2i_scores_bf16 (same plan / pack / image prep) plus
 *   region_norm [n_img][36] f32: |v_k| of the (rounded) regions (the square roots of the Gram pack's diagonal);
 *   gq_frag [n_tiles][4096] f32 from itr_scan_caption_gram_frag_bf16: per 32-row quarter the block-diagonal word Gram
 *          (w_j . w_j' for two words of the same caption, else 0; fp32 accumulation, rounded to tf32) in the fragment
 *          order of the kernel's mma.sync A operand -- an opaque companion of the packed words, valid for the same plan.
 * Captions of more than 32 words (the planner's `long` tiles) are NOT scored: their columns of `scores` are left
 * untouched for the two-phase path (itr_scan_affinity_bf16 + itr_scan_epilogue_f32) to fill. */
int itr_scan_caption_gram_frag_bf16(const uint16_t* words_bf16, const int32_t* row_meta, int n_tiles, float* gq_frag, void* stream);
int itr_scan_i2t_scores_bf16(const uint16_t* images_bf16, const float* region_norm, int n_img,
                             const uint16_t* words_bf16, const int32_t* row_meta, const float* gq_frag, int n_tiles,
                             int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                             float* scores, int64_t ld_scores, void* stream);

/* ---- fused evaluation: scores + i2t / t2i ranking without the score matrix (evaluation.py:124-153 + 156-222) -------
 * rank of a query = number of scores strictly above its (best) ground-truth score, so the ranks need the ground-truth
 * scores first:
 *   1. itr_scan_plan_gt_items (HOST): int32 quadruples (word tile A, word tile B or n_tiles, image tile, 0): two word
 *      tiles in which a packed caption meets a ground-truth image of that image tile -- global caption cap_offset + c
 *      belongs to image (cap_offset + c) / caps_per_img.
 *   2. itr_scan_t2i_gt_thresholds_bf16: the fused kernel on those items only (< 1 % of the matrix), same packed rows and
 *      same arithmetic as the full pass, so the thresholds are bit-identical to the scores they are compared with.
 *      thr_col[c] = score of caption c with its image (NaN if that image is not in [0, n_img)); thr_row[i] = best score
 *      of image i with its captions AMONG THIS LAUNCH'S (-inf if none).  Multi-GPU: all-reduce(MAX) thr_row across the
 *      caption shards before step 3.
 *   3. itr_scan_t2i_count_bf16: the full pass; every score is compared with its thresholds as it is produced:
 *      cnt_col[c] / cnt_row[i] = #scores strictly above (the t2i rank of caption c / this shard's part of the i2t rank
 *      of image i), best_col / best_row = max of (orderable(score) << 32 | ~index) (index = image / GLOBAL caption).
 *      `scores` may be NULL: the matrix is then never written.  The images may be fed in several launches (a range of
 *      rows each: pointers advanced to the range, img_offset = its first image, accumulate_columns = 1 from the second
 *      launch on) -- how the multi-GPU path scores its own image shard while the others are still arriving.
 * Replaces cal_sims + i2t + t2i for SCAN t2i (same feature norms as itr_scan_t2i_scores_bf16). */
int itr_scan_plan_gt_items(const int32_t* row_meta_host, int n_tiles, int cap_offset, int caps_per_img, int n_img,
                           int32_t* items_host, int max_items, int* n_items);
int itr_scan_t2i_gt_thresholds_bf16(const uint16_t* images_bf16, const void* gram_pack, int n_img,
                                    const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm,
                                    int n_tiles, int n_cap, const int32_t* items, int n_items, int feature_norm, int agg,
                                    float lambda_softmax, float lambda_lse, int cap_offset, int caps_per_img,
                                    float* thr_col, float* thr_row, void* stream);
int itr_scan_t2i_count_bf16(const uint16_t* images_bf16, const void* gram_pack, int n_img,
                            const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm,
                            int n_tiles, int n_cap, int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                            int cap_offset, const float* thr_col, const float* thr_row, float* scores, int64_t ld_scores,
                            int32_t* cnt_row, int32_t* cnt_col, uint64_t* best_row, uint64_t* best_col,
                            int img_offset, int accumulate_columns, void* stream);

/* Debug / bring-up: raw region-word affinities of ONE (word tile, image tile) pair as the
 * tensor cores produced them: out[128 rows][144 cols] fp32. */
int itr_scan_t2i_affinity_debug(const uint16_t* images_bf16, int n_img, const uint16_t* words_bf16, int n_tiles,
                                int word_tile, int image_tile, float* out, void* stream);

/* Debug / tuning: the score kernel (clipped_l2norm, LSE, lambda 9 / 6) with per-role wait-cycle counters,
 * counters[cta][16] int64 (cta < #SMs): 0 producer total, 1 producer wait(empty), 2 MMA total, 3 MMA wait(tempty),
 * 4 MMA wait(full), 5 items, 6+2g epilogue group g wait(tfull+afull), 7+2g wait(uready), 14 epilogue total, 15 wait(afull).
 * mode bit 0: control warpgroup placed last (1) or first (0); bit 1: skip the epilogue arithmetic (scores are garbage). */
int itr_scan_t2i_profile(const uint16_t* images_bf16, const void* gram_pack, int n_img,
                         const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm,
                         int n_tiles, float* scores, int64_t ld_scores, int64_t* counters, int mode, void* stream);

/* Same counters for the CTA-pair kernel (csrc/scan_t2i_tc2.cu): counters[pair][rank][16] int64, pair < #SMs / 2. */
int itr_scan_t2i_pair_profile(const uint16_t* images_bf16, const void* gram_pack, int n_img,
                              const uint16_t* words_bf16, const int32_t* row_meta, const float* row_wnorm,
                              int n_tiles, float* scores, int64_t ld_scores, int64_t* counters, void* stream);

/* Tuning: cycles until `n_issuers` warps have each pushed `iters` tcgen05.mma (M=128, K=16, kind::f16) through the
 * tensor pipe of one CTA, on `n_ctas` CTAs; see csrc/scan_t2i_tc.cu.  cycles[n_ctas]. */
int itr_tc_mma_microbench(int n_cols, int n_acc, int iters, int a_tmem, int kadv, int n_issuers, int n_ctas, int64_t* cycles, void* stream);
/* Tuning: the same for the CTA-pair form (tcgen05.mma.cta_group::2, M=256 across the two SMs of a cluster of 2; each CTA
 * supplies N/2 rows of B), on `n_pairs` clusters; see csrc/tc_microbench2.cu.  cycles[n_pairs]. */
int itr_tc_mma2_microbench(int n_cols, int n_acc, int iters, int a_tmem, int n_issuers, int n_pairs, int64_t* cycles, void* stream);

/* ---- hinge loss: ContrastiveLoss.forward / TripletLoss.forward, Objectives.py:93-115, 492-517
 * loss (1 float, device) = sum of both directions; dscores (n x n, may be NULL) = dloss/dscores. */
int itr_hinge_fwd_bwd_f32(const float* scores, int64_t ld_scores, int n, float margin, int max_violation,
                          float* loss, float* dscores, int64_t ld_dscores, void* stream);
/* VSE++ training step of the path: scores = im @ s.T, hinge, and the gradients w.r.t. both
 * embedding matrices (d_im = dS @ s, d_s = dS.T @ im).  ws: itr_cosine_hinge_workspace_f32(n, d) floats, 16-byte aligned.
 * Batches up to 264 with d % 4 == 0 run as ONE cooperative launch (csrc/vse_step.cu: split-K scores, fixed-order
 * reduction + hinge statistics, gradients, two grid barriers); larger ones as five launches. */
int64_t itr_cosine_hinge_workspace_f32(int n, int d);
int itr_cosine_hinge_fwd_bwd_f32(const float* im, const float* s, int n, int d, float margin, int max_violation,
                                 float* ws, float* loss, float* d_im, float* d_s, void* stream);

/* ---- ranking: i2t / t2i, evaluation.py:156-189, 192-222 ---------------------------------
 * Works on a column block [n_img x n_cap_local] of the full matrix whose first column is global
 * caption `cap_offset` (a multiple of caps_per_img), so caption shards rank locally.
 *   thr_col[c]  = score of caption c's ground-truth image  (global image (cap_offset+c)/caps_per_img)
 *   thr_row[i]  = best score among image i's ground-truth captions inside this block (-inf if none)
 *   cnt_row[i]  = #{c : scores[i,c] > thr_row[i]}        cnt_col[c] = #{i : scores[i,c] > thr_col[c]}
 *   best_row[i] / best_col[c] = arg-max packed as (orderable score bits << 32) | ~index
 * With one block, rank_i2t = cnt_row and rank_t2i = cnt_col (position among strictly greater scores).
 * Multi-GPU: all-reduce MAX thr_row, run the count pass, all-reduce SUM cnt_row / MAX best_row. */
int itr_rank_thresholds_f32(const float* scores, int64_t ld_scores, int n_img, int n_cap_local,
                            int cap_offset, int caps_per_img, float* thr_row, float* thr_col, void* stream);
int itr_rank_count_f32(const float* scores, int64_t ld_scores, int n_img, int n_cap_local, int cap_offset,
                       const float* thr_row, const float* thr_col,
                       int32_t* cnt_row, int32_t* cnt_col, uint64_t* best_row, uint64_t* best_col, void* stream);

/* cal_sims hands back a host float64 matrix (evaluation.py:140,152): converts a device float32 score block and writes it
 * straight into MAPPED page-locked host memory (cudaHostAlloc / pinned torch memory; row pitch ld_host in elements).  A
 * 128-thread block per SM: runs next to the persistent score kernel, so finished caption blocks leave the device while
 * the next ones are scored. */
int itr_scores_to_host_f64(const float* scores, int64_t ld_scores, int n_rows, int n_cols, double* host_mapped,
                           int64_t ld_host, void* stream);

/* float64 variant for a whole host-provided matrix (i2t(sims) / t2i(sims) take the float64
 * array cal_sims returns, evaluation.py:140,156,192): ranks (strictly-greater counts) and the
 * arg-max (lowest index on ties) of every row and every column. */
int itr_rank_f64(const double* scores, int64_t ld_scores, int n_img, int n_cap, int caps_per_img,
                 int32_t* rank_row, int32_t* rank_col, int32_t* top1_row, int32_t* top1_col, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ITR_B200_H_ */
