// Phase 1 of the float32 SCAN kernels (forward scores and backward coefficients): the raw affinity tile of
// one caption x 4 images in shared memory, plus word / region norms and (i2t) the caption's word Gram.
#pragma once

#include "common.cuh"

namespace itr {

constexpr int SF_IMGS = 4;
constexpr int SF_BK = 32;
constexpr int SF_LMAX = ITR_MAX_WORDS_F32;   // 96
constexpr int SF_WP = SF_LMAX + 4;           // row pitch of the word tile Ws
constexpr int SF_LP = SF_LMAX + 1;           // padded row length of the A arrays

struct ScanF32Params {
  const float* images; const float* gram; const float* captions; const int32_t* cap_lens;
  int n_img, R, n_cap, lmax, d;
  int cross_attn, feature_norm, agg;
  float lambda_softmax, lambda_lse;
  float* scores; int64_t ld_scores;
};

template <int CPT>
__device__ __forceinline__ void scan_f32_gemm(const ScanF32Params& p, int img0, int n_rows, const float* W, int n,
                                              float* Vs, float* Ws, float* Araw, float* qn_w, float* vn2, float* Gcap) {
  // Vs[SF_BK][148], Ws[SF_BK][SF_WP]; thread (ty, tx): rows ty*9..ty*9+8, cols tx + 16*c
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int RT = SF_IMGS * p.R;    // 144 rows when R = 36
  const int rpt = (RT + 15) / 16;  // rows per thread (9)
  float acc[9][CPT];
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[i][c] = 0.f;
  float wn_acc = 0.f, vn_acc = 0.f;
  for (int k0 = 0; k0 < p.d; k0 += SF_BK) {
    for (int e = tid; e < RT * SF_BK; e += 256) {
      int r = e / SF_BK, k = e % SF_BK;
      float v = 0.f;
      if (r < n_rows && k0 + k < p.d) v = p.images[((int64_t)img0 * p.R + r) * p.d + k0 + k];
      Vs[k * 148 + r] = v;
    }
    for (int e = tid; e < CPT * 16 * SF_BK; e += 256) {      // only the columns this caption's tile uses
      int j = e / SF_BK, k = e % SF_BK;
      float v = 0.f;
      if (j < n && k0 + k < p.d) v = W[(int64_t)j * p.d + k0 + k];
      Ws[k * SF_WP + j] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < SF_BK; ++k) {
      float wv[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) wv[c] = Ws[k * SF_WP + tx + 16 * c];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        float vv = (i < rpt) ? Vs[k * 148 + ty * rpt + i] : 0.f;
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(vv, wv[c], acc[i][c]);
      }
    }
    // squared norms of words / regions, and the caption's word Gram (i2t only)
    if (tid < n) {
      float s = 0.f;
      for (int k = 0; k < SF_BK; ++k) s = fmaf(Ws[k * SF_WP + tid], Ws[k * SF_WP + tid], s);
      wn_acc += s;
    }
    if (tid < RT) {
      float s = 0.f;
      for (int k = 0; k < SF_BK; ++k) s = fmaf(Vs[k * 148 + tid], Vs[k * 148 + tid], s);
      vn_acc += s;
    }
    if (p.cross_attn == ITR_I2T) {
      for (int o = tid; o < n * n; o += 256) {
        int a = o / n, b = o % n;
        float s = 0.f;
        for (int k = 0; k < SF_BK; ++k) s = fmaf(Ws[k * SF_WP + a], Ws[k * SF_WP + b], s);
        Gcap[a * SF_LP + b] += s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    int r = ty * rpt + i;
    if (i < rpt && r < RT) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        int j = tx + 16 * c;
        if (j < SF_LMAX) Araw[r * SF_LP + j] = acc[i][c];
      }
    }
  }
  if (tid < n) qn_w[tid] = sqrtf(wn_acc);
  if (tid < RT) vn2[tid] = sqrtf(vn_acc);
}

__device__ __forceinline__ float leaky01(float x) { return x > 0.f ? x : 0.1f * x; }

}  // namespace itr
