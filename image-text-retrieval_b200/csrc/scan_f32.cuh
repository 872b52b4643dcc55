// Phase 1 of the float32 SCAN kernels (forward scores and backward coefficients): the raw affinity tile of
// one caption x 4 images in shared memory, plus word / region norms and (i2t) the caption's word Gram.
#pragma once

#include "common.cuh"

namespace itr {

constexpr int SF_IMGS = 4;
constexpr int SF_BK = 32;
constexpr int SF_LMAX = ITR_MAX_WORDS_F32;   // 96
constexpr int SF_LP = SF_LMAX + 1;           // largest padded row length of the A arrays
// odd row pitch for a batch whose longest caption has lmax words (the i2t accesses stride by LP across lanes)
__host__ __device__ inline int sf_pitch(int lmax) { return (lmax + 1) | 1; }

struct ScanF32Params {
  const float* images; const float* gram; const float* captions; const int32_t* cap_lens;
  int n_img, R, n_cap, lmax, d;
  int cross_attn, feature_norm, agg;
  float lambda_softmax, lambda_lse;
  float* scores; int64_t ld_scores;
};

// Shared-memory operand tiles of one k-slab, both k-fastest with a 4-float skew (16-byte aligned rows, so the inner
// product reads float4 = 4 k-values per LDS): Vs[144 rows][SF_KP], Ws[96 rows][SF_KP].
constexpr int SF_KP = SF_BK + 4;
constexpr int SF_VS_FLOATS = 144 * SF_KP;
constexpr int SF_WS_FLOATS = SF_LMAX * SF_KP;

// one slab row (32 consecutive k) from global memory into a tile row; zero beyond the row / embedding bounds
__device__ __forceinline__ void sf_load_rows(const float* __restrict__ src, int n_valid, int n_rows, int d, int k0, bool vec,
                                             float* __restrict__ tile) {
  for (int e = threadIdx.x; e < n_rows * (SF_BK / 4); e += 256) {
    const int r = e >> 3, k = (e & 7) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < n_valid) {
      const float* g = src + (int64_t)r * d + k0 + k;
      if (vec && k0 + k + 3 < d) {
        v = *reinterpret_cast<const float4*>(g);
      } else {
        if (k0 + k + 0 < d) v.x = g[0];
        if (k0 + k + 1 < d) v.y = g[1];
        if (k0 + k + 2 < d) v.z = g[2];
        if (k0 + k + 3 < d) v.w = g[3];
      }
    }
    *reinterpret_cast<float4*>(tile + r * SF_KP + k) = v;
  }
}

__device__ __forceinline__ float sf_dot_slab(const float* a, const float* b) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < SF_BK; k += 4) {
    const float4 x = *reinterpret_cast<const float4*>(a + k), y = *reinterpret_cast<const float4*>(b + k);
    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
  }
  return s;
}

template <int CPT>
__device__ __forceinline__ void scan_f32_gemm(const ScanF32Params& p, int img0, int n_rows, const float* W, int n, int LP,
                                              float* Vs, float* Ws, float* Araw, float* qn_w, float* vn2, float* Gcap) {
  // thread (ty, tx): region rows ty*rpt .. ty*rpt + rpt - 1, word columns tx + 16*c
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int RT = SF_IMGS * p.R;    // 144 rows when R = 36
  const int rpt = (RT + 15) / 16;  // rows per thread (9)
  const float* V = p.images + (int64_t)img0 * p.R * p.d;
  const bool vec = (p.d % 4 == 0) && ((reinterpret_cast<uintptr_t>(V) & 15) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  float acc[9][CPT];
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[i][c] = 0.f;
  float wn_acc = 0.f, vn_acc = 0.f;
  for (int k0 = 0; k0 < p.d; k0 += SF_BK) {
    sf_load_rows(V, n_rows, 16 * rpt, p.d, k0, vec, Vs);
    sf_load_rows(W, n, CPT * 16, p.d, k0, vec, Ws);          // only the columns this caption's tile uses
    __syncthreads();
#pragma unroll 2
    for (int k = 0; k < SF_BK; k += 4) {
      float4 wv[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) wv[c] = *reinterpret_cast<const float4*>(Ws + (tx + 16 * c) * SF_KP + k);
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        if (i < rpt) {
          const float4 vv = *reinterpret_cast<const float4*>(Vs + (ty * rpt + i) * SF_KP + k);
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            float a = acc[i][c];
            a = fmaf(vv.x, wv[c].x, a); a = fmaf(vv.y, wv[c].y, a); a = fmaf(vv.z, wv[c].z, a); a = fmaf(vv.w, wv[c].w, a);
            acc[i][c] = a;
          }
        }
      }
    }
    // squared norms of words / regions, and the caption's word Gram (i2t only)
    if (tid < n) wn_acc += sf_dot_slab(Ws + tid * SF_KP, Ws + tid * SF_KP);
    if (tid < RT) vn_acc += sf_dot_slab(Vs + tid * SF_KP, Vs + tid * SF_KP);
    if (p.cross_attn == ITR_I2T) {
      for (int o = tid; o < n * n; o += 256) {
        const int a = o / n, b = o % n;
        Gcap[a * LP + b] += sf_dot_slab(Ws + a * SF_KP, Ws + b * SF_KP);
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
        if (j < LP) Araw[r * LP + j] = acc[i][c];
      }
    }
  }
  if (tid < n) qn_w[tid] = sqrtf(wn_acc);
  if (tid < RT) vn2[tid] = sqrtf(vn_acc);
}

__device__ __forceinline__ float leaky01(float x) { return x > 0.f ? x : 0.1f * x; }

}  // namespace itr
