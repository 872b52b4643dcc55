// float32 CUDA-core kernels of libitr_b200: the 1e-5 "fp32 mode" of the hot path.
//   - cosine scores / small dense products      (Objectives.py:18-21)
//   - SCAN cross-attention scores, every mode   (Objectives.py:329-476)
//   - max-violation / sum hinge, fwd + bwd      (Objectives.py:93-115, 492-517)
//   - i2t / t2i rank counting                   (evaluation.py:156-222)
// The throughput path for SCAN t2i lives in scan_t2i_tc.cu (tcgen05).
#include <cfloat>
#include <cstdlib>
#include <cstdarg>

#include "common.cuh"
#include "scan_f32.cuh"

namespace itr {

std::string& last_error() {
  static thread_local std::string s;
  return s;
}
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

// =========================================================================================
// Generic strided SGEMM:  C[m][n] = sum_k A(m,k) * B(n,k)
//   A(m,k) = A[m*a_rs + k*a_cs],  B(n,k) = B[n*b_rs + k*b_cs]
// TB x TB x 16 block tile, (TB/16) x (TB/16) register tile, 256 threads.  TB = 64 for large outputs,
// TB = 32 when the output is so small (the 128 x 128 training batch) that 64-wide tiles would leave
// the GPU with a handful of CTAs.
// =========================================================================================
constexpr int GK = 16;

template <int TB>
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, int64_t a_rs, int64_t a_cs,
                     const float* __restrict__ B, int64_t b_rs, int64_t b_cs,
                     float* __restrict__ C, int64_t ldc, int M, int N, int K) {
  constexpr int T = TB / 16;                       // register tile edge: 4 or 2
  __shared__ __align__(16) float As[GK][TB + 4];
  __shared__ __align__(16) float Bs[GK][TB + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TB, n0 = blockIdx.x * TB;
  const bool a_kfast = (a_cs == 1), b_kfast = (b_cs == 1);
  float acc[T][T] = {};
  for (int k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
    for (int i = 0; i < TB * GK / 256; ++i) {
      int e = tid + 256 * i;
      int mm = a_kfast ? e / GK : e % TB, kk = a_kfast ? e % GK : e / TB;
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? A[gm * a_rs + gk * a_cs] : 0.f;
      int nn = b_kfast ? e / GK : e % TB;
      kk = b_kfast ? e % GK : e / TB;
      int gn = n0 + nn;
      gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < K) ? B[gn * b_rs + gk * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float av[T], bv[T];
#pragma unroll
      for (int i = 0; i < T; ++i) { av[i] = As[kk][ty * T + i]; bv[i] = Bs[kk][tx * T + i]; }
#pragma unroll
      for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < T; ++i) {
    int gm = m0 + ty * T + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < T; ++j) {
      int gn = n0 + tx * T + j;
      if (gn < N) C[gm * ldc + gn] = acc[i][j];
    }
  }
}

static int launch_sgemm(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                        float* C, int64_t ldc, int M, int N, int K, cudaStream_t st) {
  static int force_tb = -1;
  if (force_tb < 0) { const char* e = getenv("ITR_B200_SGEMM_TB"); force_tb = e ? atoi(e) : 0; }
  if (force_tb == 32 || (force_tb == 0 && (int64_t)M * N <= 256 * 1024)) {
    dim3 grid((N + 31) / 32, (M + 31) / 32);
    sgemm_strided_kernel<32><<<grid, 256, 0, st>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K);
  } else {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    sgemm_strided_kernel<64><<<grid, 256, 0, st>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K);
  }
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

// =========================================================================================
// Region Gram  G_i = V_i V_i^T  (n_regions x n_regions per image), fp32.
// =========================================================================================
__global__ void __launch_bounds__(256)
region_gram_kernel(const float* __restrict__ images, int R, int d, float* __restrict__ gram) {
  extern __shared__ float sm[];   // [R][KC+1]
  constexpr int KC = 128;
  const float* V = images + (int64_t)blockIdx.x * R * d;
  const int n_out = R * R;
  float acc[8] = {};               // supports R*R <= 2048
  for (int k0 = 0; k0 < d; k0 += KC) {
    for (int e = threadIdx.x; e < R * KC; e += blockDim.x) {
      int r = e / KC, k = e % KC;
      sm[r * (KC + 1) + k] = (k0 + k < d) ? V[(int64_t)r * d + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      int idx = threadIdx.x + o * 256;
      if (idx < n_out) {
        const float* a = sm + (idx / R) * (KC + 1);
        const float* b = sm + (idx % R) * (KC + 1);
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < KC; ++k) s = fmaf(a[k], b[k], s);
        acc[o] += s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    int idx = threadIdx.x + o * 256;
    if (idx < n_out) gram[(int64_t)blockIdx.x * n_out + idx] = acc[o];
  }
}

// =========================================================================================
// SCAN scores, fp32, all modes.  One block = one caption x 4 images.
//   phase 1: raw affinities A[region][word] for the 4 images (register-tiled FMA), word and
//            region norms, and (i2t) the caption's word Gram.
//   phase 2: the reference's epilogue, written once for both directions in terms of
//            source index s / query index q (func_attention, Objectives.py:421-476):
//              t2i: s = region, q = word      i2t: s = word, q = region
//            X[s][q]  -> raw_feature_norm over q -> softmax over s (x lambda_softmax)
//            r[q] = cos(query_q, sum_s alpha[q][s] context_s)
//                 = (sum_s alpha A[s][q]) / max(|query_q| * sqrt(alpha^T Gctx alpha), 1e-8)
//            with Gctx the context Gram (regions: precomputed; words: built in phase 1).
//   phase 3: aggregate r over q (LSE / Mean / Max / Sum).
// =========================================================================================
// (ScanF32Params, scan_f32_gemm and the SF_* tile constants live in scan_f32.cuh)

struct ScanEpiParams {
  int R, cross_attn, feature_norm, agg;
  float lambda_softmax, lambda_lse;
  float* scores; int64_t ld_scores;
};

// The reference's epilogue on a block-resident affinity tile (shared by the fp32 kernel and by phase 2 of the
// tensor-core generic path).  Araw / X: [n_im * R rows][LP] (row = image-major region, column = word), X is scratch.
// Gctx: t2i -> n_im region Grams (R x R each); i2t -> the caption's word Gram with row stride LP.
template <bool T2I>
__device__ __forceinline__ void scan_epilogue_dir(const ScanEpiParams& p, int n_im, int n, int img0, int c, int LP,
                                                  float* Araw, float* X, const float* Gctx, const float* wnorm,
                                                  const float* vnorm, float* rsim, int RS) {
  const int R = p.R, tid = threadIdx.x, nthr = blockDim.x;
  constexpr bool t2i = T2I;
  // index helpers: element (image m, source s, query q) of the [row][word] arrays
  const int S = t2i ? R : n, Q = t2i ? n : R;
  auto at = [&](float* base, int m, int s, int q) -> float& {
    return t2i ? base[(m * R + s) * LP + q] : base[(m * R + q) * LP + s];
  };

  // ---- step 1: raw_feature_norm over q, for every (image, s) ---------------------------
  for (int it = tid; it < n_im * S; it += nthr) {
    int m = it / S, s = it % S;
    const int mode = p.feature_norm;
    if (mode == ITR_NORM_CLIPPED_L2 || mode == ITR_NORM_L2) {
      float ss = 0.f;
      for (int q = 0; q < Q; ++q) {
        float a = at(Araw, m, s, q);
        if (mode == ITR_NORM_CLIPPED_L2) a = leaky01(a);
        ss = fmaf(a, a, ss);
      }
      float inv = 1.f / (sqrtf(ss) + 1e-8f);
      for (int q = 0; q < Q; ++q) {
        float a = at(Araw, m, s, q);
        if (mode == ITR_NORM_CLIPPED_L2) a = leaky01(a);
        at(X, m, s, q) = a * inv;
      }
    } else if (mode == ITR_NORM_SOFTMAX) {
      float mx = -FLT_MAX;
      for (int q = 0; q < Q; ++q) mx = fmaxf(mx, at(Araw, m, s, q));
      float z = 0.f;
      for (int q = 0; q < Q; ++q) z += expf(at(Araw, m, s, q) - mx);
      float inv = 1.f / z;
      for (int q = 0; q < Q; ++q) at(X, m, s, q) = expf(at(Araw, m, s, q) - mx) * inv;
    } else {
      for (int q = 0; q < Q; ++q) {
        float a = at(Araw, m, s, q);
        at(X, m, s, q) = (mode == ITR_NORM_CLIPPED) ? leaky01(a) : a;
      }
    }
  }
  __syncthreads();

  // ---- step 2+3: softmax over s, attended cosine, for every (image, q) -------------------
  for (int it = tid; it < n_im * Q; it += nthr) {
    int m = it / Q, q = it % Q;
    float mx = -FLT_MAX;
    for (int s = 0; s < S; ++s) mx = fmaxf(mx, at(X, m, s, q) * p.lambda_softmax);
    float Z = 0.f, P = 0.f;
    for (int s = 0; s < S; ++s) {
      float e = expf(at(X, m, s, q) * p.lambda_softmax - mx);
      at(X, m, s, q) = e;
      Z += e;
      P = fmaf(e, at(Araw, m, s, q), P);
    }
    const float* G = t2i ? Gctx + m * R * R : Gctx;
    const int gs = t2i ? R : LP;
    // e^T G e with G symmetric: diagonal once, strict upper triangle twice
    float Qf = 0.f;
    for (int s = 0; s < S; ++s) {
      const float es = at(X, m, s, q);
      float u = 0.f;
      for (int s2 = s + 1; s2 < S; ++s2) u = fmaf(G[s * gs + s2], at(X, m, s2, q), u);
      Qf = fmaf(es, fmaf(G[s * gs + s], es, 2.f * u), Qf);
    }
    float qn = t2i ? wnorm[q] : vnorm[m * R + q];
    float invZ = 1.f / Z;
    float w12 = P * invZ;
    float w2 = sqrtf(fmaxf(Qf, 0.f)) * invZ;
    rsim[m * RS + q] = w12 / fmaxf(qn * w2, 1e-8f);
  }
  __syncthreads();

  // ---- step 4: aggregate over q, one warp per image --------------------------------------
  const int lane = tid & 31;
  for (int warp = tid >> 5; warp < n_im; warp += nthr >> 5) {
    const float* r = rsim + warp * RS;
    float v;
    if (p.agg == ITR_AGG_MAX) {
      v = -FLT_MAX;
      for (int q = lane; q < Q; q += 32) v = fmaxf(v, r[q]);
      v = warp_max(v);
    } else {
      v = 0.f;
      for (int q = lane; q < Q; q += 32) v += (p.agg == ITR_AGG_LSE) ? expf(r[q] * p.lambda_lse) : r[q];
      v = warp_sum(v);
      if (p.agg == ITR_AGG_LSE) v = logf(v) / p.lambda_lse;
      if (p.agg == ITR_AGG_MEAN) v = v / (float)Q;
    }
    if (lane == 0) p.scores[(int64_t)(img0 + warp) * p.ld_scores + c] = v;
  }
}

__device__ __forceinline__ void scan_epilogue_smem(const ScanEpiParams& p, int n_im, int n, int img0, int c, int LP,
                                                   float* Araw, float* X, const float* Gctx, const float* wnorm,
                                                   const float* vnorm, float* rsim, int RS) {
  if (p.cross_attn == ITR_T2I) scan_epilogue_dir<true>(p, n_im, n, img0, c, LP, Araw, X, Gctx, wnorm, vnorm, rsim, RS);
  else scan_epilogue_dir<false>(p, n_im, n, img0, c, LP, Araw, X, Gctx, wnorm, vnorm, rsim, RS);
}

__global__ void __launch_bounds__(256)
scan_f32_kernel(ScanF32Params p) {
  extern __shared__ __align__(16) float smem[];
  const int R = p.R, RT = SF_IMGS * R, LP = sf_pitch(p.lmax);
  float* Vs = smem;                           // SF_VS_FLOATS
  float* Ws = Vs + SF_VS_FLOATS;              // SF_WS_FLOATS
  float* Araw = Ws + SF_WS_FLOATS;            // RT*LP   raw affinities [row = img*R + region][word]
  float* X = Araw + RT * LP;                  // RT*LP   normalised / exponentiated copy
  float* Gctx = X + RT * LP;                  // t2i: SF_IMGS*R*R region Grams; i2t: LP*LP word Gram
  const int g_floats = (p.cross_attn == ITR_T2I) ? SF_IMGS * R * R : LP * LP;
  float* wnorm = Gctx + g_floats;             // SF_LMAX  |w_j|
  float* vnorm = wnorm + SF_LMAX;             // RT       |v_k|
  float* rsim = vnorm + RT;                   // SF_IMGS * max(R, SF_LMAX)
  const int RS = max(R, SF_LMAX);

  const int c = blockIdx.x;
  const int img0 = blockIdx.y * SF_IMGS;
  const int n_im = min(SF_IMGS, p.n_img - img0);
  const int n = p.cap_lens[c];
  const float* W = p.captions + (int64_t)c * p.lmax * p.d;
  const int tid = threadIdx.x;
  const bool t2i = (p.cross_attn == ITR_T2I);

  if (t2i) {
    for (int e = tid; e < n_im * R * R; e += 256) Gctx[e] = p.gram[(int64_t)img0 * R * R + e];
  } else {
    for (int e = tid; e < LP * LP; e += 256) Gctx[e] = 0.f;
  }
  __syncthreads();

  const int cpt = (n + 15) / 16;
  switch (cpt) {
    case 1: scan_f32_gemm<1>(p, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
    case 2: scan_f32_gemm<2>(p, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
    case 3: scan_f32_gemm<3>(p, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
    case 4: scan_f32_gemm<4>(p, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
    case 5: scan_f32_gemm<5>(p, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
    default: scan_f32_gemm<6>(p, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
  }
  __syncthreads();

  ScanEpiParams ep{R, p.cross_attn, p.feature_norm, p.agg, p.lambda_softmax, p.lambda_lse, p.scores, p.ld_scores};
  scan_epilogue_smem(ep, n_im, n, img0, c, LP, Araw, X, Gctx, wnorm, vnorm, rsim, RS);
}

// =========================================================================================
// Phase 2 of the tensor-core generic path: the same epilogue, fed with the raw affinities the tcgen05 kernel
// dumped (itr_scan_affinity_bf16), layout [word tile][image][row][R].  One block = one caption x 4 images.
// =========================================================================================
struct ScanEpiKernelParams {
  const float* affinity; int n_img;            // images in this chunk (= second dimension of the dump)
  const int32_t* cap_row0; const int32_t* cap_lens; const int32_t* cap_ids; int n_ids; int LP;
  int imgs;                                    // images per block
  const float* row_wnorm;                      // [packed rows]
  const float* region_norm;                    // [n_img][R]
  const float* region_gram;                    // t2i: [n_img][R][R]
  const float* cap_gram; const int64_t* gram_off;   // i2t: packed n_c x n_c word Grams
  ScanEpiParams e;
};

// i2t: 256 threads and 7 images per block (7 x 36 regions = 252 work items in the softmax step) while the tile fits
// two blocks per SM, else 4 images; t2i: 128 threads, 4 images (4 x n words).
__global__ void __launch_bounds__(256)
scan_epilogue_kernel(ScanEpiKernelParams p) {
  const int EPI_THREADS = blockDim.x;
  extern __shared__ __align__(16) float smem[];
  const int R = p.e.R, RT = p.imgs * R, LP = p.LP;
  float* Araw = smem;
  float* X = Araw + RT * LP;
  float* Gctx = X + RT * LP;
  const int g_floats = (p.e.cross_attn == ITR_T2I) ? p.imgs * R * R : LP * LP;
  float* wnorm = Gctx + g_floats;              // LP
  float* vnorm = wnorm + LP;                   // RT
  float* rsim = vnorm + RT;                    // imgs * RS
  const int RS = max(R, LP);
  const int c = p.cap_ids ? p.cap_ids[blockIdx.x] : (int)blockIdx.x;
  const int img0 = blockIdx.y * p.imgs, tid = threadIdx.x;
  const int n_im = min(p.imgs, p.n_img - img0);
  const int n = p.cap_lens[c];
  const int row0 = p.cap_row0[c];
  const int tile = row0 / ITR_TILE_WORDS, r_in = row0 % ITR_TILE_WORDS;
  const bool t2i = (p.e.cross_attn == ITR_T2I);
  // affinities: global [tile][img][row][k] (k fastest) -> smem Araw[(m*R + k)*LP + j]
  for (int m = 0; m < n_im; ++m) {                      // n contiguous rows of R floats per image
    const float* src = p.affinity + (((size_t)tile * p.n_img + img0 + m) * ITR_TILE_WORDS + r_in) * R;
    float* dst = Araw + m * R * LP;
    for (int e = tid; e < n * R; e += EPI_THREADS) {
      const int j = e / R, k = e - j * R;
      dst[k * LP + j] = src[e];
    }
  }
  if (t2i) {
    for (int e = tid; e < n_im * R * R; e += EPI_THREADS) Gctx[e] = p.region_gram[(size_t)img0 * R * R + e];
  } else {
    const float* g = p.cap_gram + p.gram_off[c];
    for (int e = tid; e < n * n; e += EPI_THREADS) Gctx[(e / n) * LP + (e % n)] = g[e];
  }
  for (int j = tid; j < n; j += EPI_THREADS) wnorm[j] = p.row_wnorm[row0 + j];
  for (int e = tid; e < n_im * R; e += EPI_THREADS) vnorm[e] = p.region_norm[(size_t)img0 * R + e];
  __syncthreads();
  scan_epilogue_smem(p.e, n_im, n, img0, c, LP, Araw, X, Gctx, wnorm, vnorm, rsim, RS);
}

// i2t, captions of at most NMAX <= 32 words (all but a handful): the same phase 2 with the softmax-over-words step
// specialised.  In the i2t layout a thread's (image, region) column is a contiguous shared-memory row, so its NMAX
// attention weights live in registers, the word Gram is read as broadcast float4 and alpha^T G alpha is fully
// unrolled over the strict upper triangle.  One block = one caption x 7 images (252 of 256 threads busy).
constexpr int I2T_IMGS = 7;

template <int NMAX>
__global__ void __launch_bounds__(256)
scan_i2t_epilogue_kernel(ScanEpiKernelParams p) {
  constexpr int R = ITR_REGIONS, RT = I2T_IMGS * R, LP = NMAX + 1;     // odd pitch
  extern __shared__ __align__(16) float smem[];
  float* Araw = smem;                          // RT*LP  [img*R + region][word]
  float* X = Araw + RT * LP;                   // RT*LP
  float* G = X + RT * LP;                      // NMAX*NMAX word Gram, zero padded
  float* vnorm = G + NMAX * NMAX;              // RT
  float* rsim = vnorm + RT;                    // RT
  const int c = p.cap_ids ? p.cap_ids[blockIdx.x] : (int)blockIdx.x;
  const int img0 = blockIdx.y * I2T_IMGS, tid = threadIdx.x;
  const int n_im = min(I2T_IMGS, p.n_img - img0);
  const int n = p.cap_lens[c];
  const int row0 = p.cap_row0[c];
  const int tile = row0 / ITR_TILE_WORDS, r_in = row0 % ITR_TILE_WORDS;
  const int mode = p.e.feature_norm;

  {
    const float* g = p.cap_gram + p.gram_off[c];
    for (int e = tid; e < NMAX * NMAX; e += 256) {
      const int a = e / NMAX, b = e - a * NMAX;
      G[e] = (a < n && b < n) ? g[a * n + b] : 0.f;
    }
  }
  for (int e = tid; e < n_im * R; e += 256) vnorm[e] = p.region_norm[(size_t)img0 * R + e];

  // ---- load + step 1: one thread per (image, word).  The word's 36 affinities are one contiguous, 16-byte aligned
  // 144-byte row of the dump: nine float4 loads, raw_feature_norm over the regions in registers, then the
  // transposed scatter into the [region][word] tiles (lanes = consecutive words: conflict-free).
  for (int it = tid; it < n_im * n; it += 256) {
    const int m = it / n, s = it - m * n;
    const float4* src = reinterpret_cast<const float4*>(
        p.affinity + (((size_t)tile * p.n_img + img0 + m) * ITR_TILE_WORDS + r_in + s) * R);
    float v[R];
#pragma unroll
    for (int q4 = 0; q4 < R / 4; ++q4) {
      const float4 t = src[q4];
      v[q4 * 4 + 0] = t.x; v[q4 * 4 + 1] = t.y; v[q4 * 4 + 2] = t.z; v[q4 * 4 + 3] = t.w;
    }
    float* a = Araw + m * R * LP + s;
    float* x = X + m * R * LP + s;
#pragma unroll
    for (int q = 0; q < R; ++q) a[q * LP] = v[q];
    if (mode == ITR_NORM_CLIPPED_L2 || mode == ITR_NORM_L2) {
      float ss = 0.f;
#pragma unroll
      for (int q = 0; q < R; ++q) {
        if (mode == ITR_NORM_CLIPPED_L2) v[q] = leaky01(v[q]);
        ss = fmaf(v[q], v[q], ss);
      }
      const float inv = 1.f / (sqrtf(ss) + 1e-8f);
#pragma unroll
      for (int q = 0; q < R; ++q) x[q * LP] = v[q] * inv;
    } else if (mode == ITR_NORM_SOFTMAX) {
      float mx = -FLT_MAX;
#pragma unroll
      for (int q = 0; q < R; ++q) mx = fmaxf(mx, v[q]);
      float z = 0.f;
#pragma unroll
      for (int q = 0; q < R; ++q) { v[q] = expf(v[q] - mx); z += v[q]; }
      const float inv = 1.f / z;
#pragma unroll
      for (int q = 0; q < R; ++q) x[q * LP] = v[q] * inv;
    } else {
#pragma unroll
      for (int q = 0; q < R; ++q) x[q * LP] = (mode == ITR_NORM_CLIPPED) ? leaky01(v[q]) : v[q];
    }
  }
  __syncthreads();

  // ---- step 2: softmax over the caption's words, attended cosine, one thread per (image, region) ----
  if (tid < n_im * R) {
    const float* a = Araw + tid * LP;
    const float* x = X + tid * LP;
    const float c2 = p.e.lambda_softmax * 1.4426950408889634f;      // exp(l x) = exp2(l log2e x)
    float al[NMAX];
    float mx = -FLT_MAX;
#pragma unroll
    for (int s = 0; s < NMAX; ++s) {
      al[s] = (s < n) ? x[s] * c2 : -FLT_MAX;
      mx = fmaxf(mx, al[s]);
    }
    float Z = 0.f;
#pragma unroll
    for (int s = 0; s < NMAX; ++s) {
      al[s] = (s < n) ? exp2f(al[s] - mx) : 0.f;
      Z += al[s];
    }
    const float invZ = 1.f / Z;
    float P = 0.f;
#pragma unroll
    for (int s = 0; s < NMAX; ++s) {
      al[s] *= invZ;
      if (s < n) P = fmaf(al[s], a[s], P);
    }
    float Qf = 0.f;
#pragma unroll
    for (int s = 0; s < NMAX; ++s) {
      float u = 0.f;
#pragma unroll
      for (int s2 = s + 1; s2 < NMAX; ++s2) u = fmaf(G[s * NMAX + s2], al[s2], u);
      Qf = fmaf(al[s], fmaf(G[s * NMAX + s], al[s], 2.f * u), Qf);
    }
    const float w2 = sqrtf(fmaxf(Qf, 0.f));
    rsim[tid] = P / fmaxf(vnorm[tid] * w2, 1e-8f);
  }
  __syncthreads();

  // ---- step 3: aggregate over the 36 regions, one warp per image -----------------------------------
  const int lane = tid & 31;
  for (int m = tid >> 5; m < n_im; m += 8) {
    const float* r = rsim + m * R;
    float v;
    if (p.e.agg == ITR_AGG_MAX) {
      v = -FLT_MAX;
      for (int q = lane; q < R; q += 32) v = fmaxf(v, r[q]);
      v = warp_max(v);
    } else {
      v = 0.f;
      for (int q = lane; q < R; q += 32) v += (p.e.agg == ITR_AGG_LSE) ? expf(r[q] * p.e.lambda_lse) : r[q];
      v = warp_sum(v);
      if (p.e.agg == ITR_AGG_LSE) v = logf(v) / p.e.lambda_lse;
      if (p.e.agg == ITR_AGG_MEAN) v = v / (float)R;
    }
    if (lane == 0) p.e.scores[(int64_t)(img0 + m) * p.e.ld_scores + c] = v;
  }
}

template <int NMAX>
static int launch_i2t_epilogue(const ScanEpiKernelParams& p, int n_ids, int n_img, cudaStream_t st) {
  constexpr int RT = I2T_IMGS * ITR_REGIONS, LP = NMAX + 1;
  const size_t smem = sizeof(float) * (2 * (size_t)RT * LP + NMAX * NMAX + 2 * RT);
  ITR_CHECK_CUDA(cudaFuncSetAttribute(scan_i2t_epilogue_kernel<NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_ids, (n_img + I2T_IMGS - 1) / I2T_IMGS);
  ITR_REQUIRE(grid.y <= 65535, "itr_scan_epilogue_f32: more than %d images per call", 65535 * I2T_IMGS);
  scan_i2t_epilogue_kernel<NMAX><<<grid, 256, smem, st>>>(p);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

// Word Gram of every caption from the packed bf16 rows: gram[gram_off[c] + j*n + j2] = w_j . w_j2 (fp32 accumulate).
__global__ void __launch_bounds__(128)
caption_gram_kernel(const uint16_t* __restrict__ words, const int32_t* __restrict__ cap_row0, const int32_t* __restrict__ cap_lens,
                    const int64_t* __restrict__ gram_off, int d, float* __restrict__ gram) {
  const int c = blockIdx.x, n = cap_lens[c], warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint16_t* W = words + (size_t)cap_row0[c] * d;
  float* g = gram + gram_off[c];
  // blockIdx.y splits the caption's n (n + 1) / 2 word pairs: a handful of long captions would otherwise keep a handful
  // of blocks busy for a millisecond (measured: 6 captions of up to 72 words, 0.78 ms with one block per caption)
  for (int o = warp + 4 * blockIdx.y; o < n * (n + 1) / 2; o += 4 * gridDim.y) {
    // pair o -> (j, j2 <= j): j = floor((sqrt(8 o + 1) - 1) / 2), corrected for rounding
    int j = (int)((sqrtf(8.f * (float)o + 1.f) - 1.f) * 0.5f);
    while (j * (j + 1) / 2 > o) --j;
    while ((j + 1) * (j + 2) / 2 <= o) ++j;
    const int j2 = o - j * (j + 1) / 2;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(W + (size_t)j * d);
    const uint32_t* b = reinterpret_cast<const uint32_t*>(W + (size_t)j2 * d);
    float s = 0.f;
    for (int v = lane; v < d / 2; v += 32) {
      uint32_t x = a[v], y = b[v];
      s = fmaf(__uint_as_float(x << 16), __uint_as_float(y << 16), s);
      s = fmaf(__uint_as_float(x & 0xffff0000u), __uint_as_float(y & 0xffff0000u), s);
    }
    s = warp_sum(s);
    if (lane == 0) { g[j * n + j2] = s; g[j2 * n + j] = s; }
  }
}

// =========================================================================================
// Hinge loss (fwd + bwd).  stats[i] = {row value, row arg / count, col value, col arg / count}
// =========================================================================================
__global__ void __launch_bounds__(128)
hinge_stats_kernel(const float* __restrict__ S, int64_t ld, int n, float margin, int max_violation,
                   float4* __restrict__ stats) {
  __shared__ float sv[2][4];
  __shared__ int si[2][4];
  const int i = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float d = S[(int64_t)i * ld + i];
  // row i: cost_s[i][j] = relu(margin + S[i][j] - d_i); column i: cost_im[j][i] = relu(margin + S[j][i] - d_i)
  float rv = max_violation ? -1.f : 0.f, cv = rv;
  int ra = 0x7fffffff, ca = 0x7fffffff, rc = 0, cc = 0;
  for (int j = tid; j < n; j += 128) {
    if (j == i) continue;
    float a = fmaxf(margin + S[(int64_t)i * ld + j] - d, 0.f);
    float b = fmaxf(margin + S[(int64_t)j * ld + i] - d, 0.f);
    if (max_violation) {
      if (a > rv) { rv = a; ra = j; }
      if (b > cv) { cv = b; ca = j; }
    } else {
      rv += a; cv += b; rc += (a > 0.f); cc += (b > 0.f);
    }
  }
  // block reduce: (max, lowest index) or (sum, count)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float orv = __shfl_xor_sync(0xffffffffu, rv, o), ocv = __shfl_xor_sync(0xffffffffu, cv, o);
    int ora = __shfl_xor_sync(0xffffffffu, ra, o), oca = __shfl_xor_sync(0xffffffffu, ca, o);
    int orc = __shfl_xor_sync(0xffffffffu, rc, o), occ = __shfl_xor_sync(0xffffffffu, cc, o);
    if (max_violation) {
      if (orv > rv || (orv == rv && ora < ra)) { rv = orv; ra = ora; }
      if (ocv > cv || (ocv == cv && oca < ca)) { cv = ocv; ca = oca; }
    } else {
      rv += orv; cv += ocv; rc += orc; cc += occ;
    }
  }
  if (lane == 0) { sv[0][warp] = rv; sv[1][warp] = cv; si[0][warp] = max_violation ? ra : rc; si[1][warp] = max_violation ? ca : cc; }
  __syncthreads();
  if (tid == 0) {
    float R = sv[0][0], Cc = sv[1][0];
    int RI = si[0][0], CI = si[1][0];
    for (int w = 1; w < 4; ++w) {
      if (max_violation) {
        if (sv[0][w] > R || (sv[0][w] == R && si[0][w] < RI)) { R = sv[0][w]; RI = si[0][w]; }
        if (sv[1][w] > Cc || (sv[1][w] == Cc && si[1][w] < CI)) { Cc = sv[1][w]; CI = si[1][w]; }
      } else {
        R += sv[0][w]; Cc += sv[1][w]; RI += si[0][w]; CI += si[1][w];
      }
    }
    if (max_violation) { R = fmaxf(R, 0.f); Cc = fmaxf(Cc, 0.f); }   // n == 1: no negatives
    stats[i] = make_float4(R, __int_as_float(RI), Cc, __int_as_float(CI));
  }
}

// grid-stride over the n x n matrix; block 0 additionally reduces the loss in a fixed order.
__global__ void __launch_bounds__(256)
hinge_finish_kernel(const float* __restrict__ S, int64_t ld, int n, float margin, int max_violation,
                    const float4* __restrict__ stats, float* __restrict__ loss,
                    float* __restrict__ dS, int64_t ldd) {
  if (blockIdx.x == 0) {
    __shared__ float part[8];
    float v = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) v += stats[i].x + stats[i].z;
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += part[w];
      *loss = t;
    }
  }
  if (dS == nullptr) return;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < (int64_t)n * n; e += (int64_t)gridDim.x * 256) {
    int i = (int)(e / n), j = (int)(e % n);
    float g;
    float4 si = stats[i];
    if (i == j) {
      g = max_violation ? -((si.x > 0.f) + (si.z > 0.f)) : -(float)(__float_as_int(si.y) + __float_as_int(si.w));
    } else {
      float4 sj = stats[j];
      if (max_violation) {
        g = (float)((__float_as_int(si.y) == j && si.x > 0.f) + (__float_as_int(sj.w) == i && sj.z > 0.f));
      } else {
        float s = S[(int64_t)i * ld + j];
        float di = S[(int64_t)i * ld + i], dj = S[(int64_t)j * ld + j];
        g = (float)((margin + s - di > 0.f) + (margin + s - dj > 0.f));
      }
    }
    dS[(int64_t)i * ldd + j] = g;
  }
}

// =========================================================================================
// Ranking
// =========================================================================================
__global__ void rank_thresholds_kernel(const float* __restrict__ S, int64_t ld, int n_img, int n_cap, int cap_offset,
                                       int cpi, float* __restrict__ thr_row, float* __restrict__ thr_col) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_cap) {
    int g = (cap_offset + t) / cpi;
    thr_col[t] = (g < n_img) ? S[(int64_t)g * ld + t] : INFINITY;
  }
  if (t < n_img) {
    float best = -INFINITY;
    for (int k = 0; k < cpi; ++k) {
      int c = t * cpi + k - cap_offset;
      if (c >= 0 && c < n_cap) best = fmaxf(best, S[(int64_t)t * ld + c]);
    }
    thr_row[t] = best;
  }
}

// One block scans a [RK_ROWS x 256] tile: column stats stay in registers, row stats go
// through a warp reduction; both are merged with atomics (integers / packed keys only, so
// the result does not depend on the order of arrival).
constexpr int RK_ROWS = 64;
__global__ void __launch_bounds__(256)
rank_count_kernel(const float* __restrict__ S, int64_t ld, int n_img, int n_cap, int cap_offset,
                  const float* __restrict__ thr_row, const float* __restrict__ thr_col,
                  int* __restrict__ cnt_row, int* __restrict__ cnt_col,
                  unsigned long long* __restrict__ best_row, unsigned long long* __restrict__ best_col) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * RK_ROWS;
  const int r1 = min(r0 + RK_ROWS, n_img);
  const bool live = c < n_cap;
  const float tc = live ? thr_col[c] : INFINITY;
  const int lane = threadIdx.x & 31;
  int ccnt = 0;
  unsigned long long cbest = 0ull;
  for (int r = r0; r < r1; ++r) {
    float v = live ? S[(int64_t)r * ld + c] : -INFINITY;
    ccnt += (v > tc);
    unsigned long long key = live ? (((unsigned long long)orderable(v)) << 32) : 0ull;
    unsigned long long ck = key | (unsigned)(~(unsigned)r);
    cbest = ck > cbest ? ck : cbest;
    // row side
    int rc = warp_sum_i(live && v > thr_row[r]);
    unsigned long long rk = live ? (key | (unsigned)(~(unsigned)(cap_offset + c))) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(0xffffffffu, rk, o);
      rk = other > rk ? other : rk;
    }
    if (lane == 0) {
      if (rc) atomicAdd(&cnt_row[r], rc);
      atomicMax(&best_row[r], rk);
    }
  }
  if (live) {
    if (ccnt) atomicAdd(&cnt_col[c], ccnt);
    atomicMax(&best_col[c], cbest);
  }
}


// float64 variant for host matrices handed to i2t()/t2i() (evaluation.py ranks a float64
// matrix, e.g. the average of two models' scores, which float32 cannot represent exactly).
// Single block of the matrix, three passes: thresholds, counts + best value, arg of best.
__device__ __forceinline__ unsigned long long orderable64(double f) {
  unsigned long long u = (unsigned long long)__double_as_longlong(f);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__global__ void rank64_thresholds_kernel(const double* __restrict__ S, int64_t ld, int n_img, int n_cap, int cpi,
                                         double* __restrict__ thr_row, double* __restrict__ thr_col) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_cap) {
    int g = t / cpi;
    thr_col[t] = (g < n_img) ? S[(int64_t)g * ld + t] : INFINITY;
  }
  if (t < n_img) {
    double best = -INFINITY;
    for (int k = 0; k < cpi; ++k) {
      int c = t * cpi + k;
      if (c < n_cap) best = fmax(best, S[(int64_t)t * ld + c]);
    }
    thr_row[t] = best;
  }
}
template <bool ARG>
__global__ void __launch_bounds__(256)
rank64_pass_kernel(const double* __restrict__ S, int64_t ld, int n_img, int n_cap,
                   const double* __restrict__ thr_row, const double* __restrict__ thr_col,
                   int* __restrict__ cnt_row, int* __restrict__ cnt_col,
                   unsigned long long* __restrict__ best_row, unsigned long long* __restrict__ best_col,
                   int* __restrict__ arg_row, int* __restrict__ arg_col) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * RK_ROWS, r1 = min(r0 + RK_ROWS, n_img);
  if (c >= n_cap) return;
  if (!ARG) {
    const double tc = thr_col[c];
    int ccnt = 0;
    unsigned long long cb = 0ull;
    for (int r = r0; r < r1; ++r) {
      double v = S[(int64_t)r * ld + c];
      ccnt += (v > tc);
      unsigned long long key = orderable64(v);
      cb = key > cb ? key : cb;
      if (v > thr_row[r]) atomicAdd(&cnt_row[r], 1);
      if (key > best_row[r]) atomicMax(&best_row[r], key);
    }
    if (ccnt) atomicAdd(&cnt_col[c], ccnt);
    atomicMax(&best_col[c], cb);
  } else {
    const unsigned long long bc = best_col[c];
    int ca = 0x7fffffff;
    for (int r = r0; r < r1; ++r) {
      unsigned long long key = orderable64(S[(int64_t)r * ld + c]);
      if (key == bc && r < ca) ca = r;
      if (key == best_row[r]) atomicMin(&arg_row[r], c);
    }
    if (ca != 0x7fffffff) atomicMin(&arg_col[c], ca);
  }
}

}  // namespace itr

// =========================================================================================
// C ABI
// =========================================================================================
using namespace itr;

// The hinge / rank-f64 entry points take their few-KB scratch from the device's default stream-ordered pool.
// With the default release threshold (0) the pool hands its memory back to the driver at every synchronisation
// and the next cudaMallocAsync pays for a real allocation; keep up to 64 MB cached instead.
static void keep_scratch_pool_warm() {
  static thread_local int done_for = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = 64ull << 20;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done_for = dev;
}

extern "C" const char* itr_last_error(void) { return last_error().c_str(); }
extern "C" int itr_version(void) { return 100; }

extern "C" int itr_device_supported(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { fail(ITR_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return -ITR_ERR_CUDA; }
  return prop.major == 10 ? 1 : 0;
}

extern "C" int itr_cosine_scores_f32(const float* im, const float* s, int n_img, int n_cap, int d,
                                     float* scores, int64_t ld_scores, void* stream) {
  ITR_REQUIRE(im && s && scores, "itr_cosine_scores_f32: null pointer");
  ITR_REQUIRE(n_img >= 0 && n_cap >= 0 && d > 0 && ld_scores >= n_cap, "itr_cosine_scores_f32: bad shape");
  if (n_img == 0 || n_cap == 0) return ITR_OK;
  return launch_sgemm(im, d, 1, s, d, 1, scores, ld_scores, n_img, n_cap, d, as_stream(stream));
}

extern "C" int itr_region_gram_f32(const float* images, int n_img, int n_regions, int d, float* gram, void* stream) {
  ITR_REQUIRE(images && gram, "itr_region_gram_f32: null pointer");
  ITR_REQUIRE(n_regions > 0 && n_regions * n_regions <= 2048 && d > 0, "itr_region_gram_f32: n_regions must be in [1, 45]");
  if (n_img <= 0) return ITR_OK;
  size_t smem = (size_t)n_regions * 129 * sizeof(float);
  region_gram_kernel<<<n_img, 256, smem, as_stream(stream)>>>(images, n_regions, d, gram);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_scan_scores_f32(const float* images, const float* gram, const float* captions,
                                   const int32_t* cap_lens, int n_img, int n_regions, int n_cap, int lmax, int d,
                                   int cross_attn, int feature_norm, int agg, float lambda_softmax, float lambda_lse,
                                   float* scores, int64_t ld_scores, void* stream) {
  ITR_REQUIRE(images && captions && cap_lens && scores, "itr_scan_scores_f32: null pointer");
  ITR_REQUIRE(cross_attn == ITR_T2I || cross_attn == ITR_I2T, "unknown cross_attn: %d", cross_attn);
  ITR_REQUIRE(feature_norm >= 0 && feature_norm <= ITR_NORM_NONE, "unknown first norm type: %d", feature_norm);
  ITR_REQUIRE(agg >= 0 && agg <= ITR_AGG_SUM, "unknown aggfunc: %d", agg);
  ITR_REQUIRE(n_regions >= 1 && n_regions <= ITR_REGIONS, "itr_scan_scores_f32: 1 to %d regions per image, got %d", ITR_REGIONS, n_regions);
  ITR_REQUIRE(lmax >= 1 && lmax <= ITR_MAX_WORDS_F32, "itr_scan_scores_f32: padded caption width %d outside [1, %d]", lmax, ITR_MAX_WORDS_F32);
  ITR_REQUIRE(cross_attn == ITR_I2T || gram != nullptr, "itr_scan_scores_f32: t2i needs the region Gram");
  ITR_REQUIRE(d > 0 && ld_scores >= n_cap, "itr_scan_scores_f32: bad shape");
  if (n_img <= 0 || n_cap <= 0) return ITR_OK;
  ScanF32Params p{images, gram, captions, cap_lens, n_img, n_regions, n_cap, lmax, d,
                  cross_attn, feature_norm, agg, lambda_softmax, lambda_lse, scores, ld_scores};
  const int RT = SF_IMGS * n_regions, LP = sf_pitch(lmax);
  const int g_floats = (cross_attn == ITR_T2I) ? SF_IMGS * n_regions * n_regions : LP * LP;
  int rs = n_regions > SF_LMAX ? n_regions : SF_LMAX;
  size_t smem = sizeof(float) * ((size_t)SF_VS_FLOATS + SF_WS_FLOATS + 2 * (size_t)RT * LP + g_floats + SF_LMAX + RT + SF_IMGS * rs);
  ITR_CHECK_CUDA(cudaFuncSetAttribute(scan_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_cap, (n_img + SF_IMGS - 1) / SF_IMGS);
  ITR_REQUIRE(grid.y <= 65535, "itr_scan_scores_f32: more than %d images per call", 65535 * SF_IMGS);
  scan_f32_kernel<<<grid, 256, smem, as_stream(stream)>>>(p);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_hinge_fwd_bwd_f32(const float* scores, int64_t ld_scores, int n, float margin, int max_violation,
                                     float* loss, float* dscores, int64_t ld_dscores, void* stream) {
  ITR_REQUIRE(scores && loss, "itr_hinge_fwd_bwd_f32: null pointer");
  ITR_REQUIRE(n >= 1 && ld_scores >= n && (dscores == nullptr || ld_dscores >= n), "itr_hinge_fwd_bwd_f32: bad shape");
  cudaStream_t st = as_stream(stream);
  float4* stats = nullptr;
  keep_scratch_pool_warm();
  ITR_CHECK_CUDA(cudaMallocAsync(&stats, sizeof(float4) * n, st));
  hinge_stats_kernel<<<n, 128, 0, st>>>(scores, ld_scores, n, margin, max_violation, stats);
  int blocks = dscores ? (int)(((int64_t)n * n + 255) / 256) : 1;
  if (blocks > 1184) blocks = 1184;
  hinge_finish_kernel<<<blocks, 256, 0, st>>>(scores, ld_scores, n, margin, max_violation, stats, loss, dscores, ld_dscores);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(stats, st);
  ITR_CHECK_CUDA(e);
  return ITR_OK;
}

// one cooperative launch for batches up to 264 (vse_step.cu); 1 = done, 0 = not applicable, < 0 = error in *status
int vse_step_fused(const float* im, const float* s, int n, int d, float margin, int max_violation, float* ws, float* loss,
                   float* d_im, float* d_s, cudaStream_t st, int* status);

extern "C" int itr_cosine_hinge_fwd_bwd_f32(const float* im, const float* s, int n, int d, float margin,
                                            int max_violation, float* ws, float* loss, float* d_im, float* d_s,
                                            void* stream) {
  ITR_REQUIRE(im && s && ws && loss, "itr_cosine_hinge_fwd_bwd_f32: null pointer");
  ITR_REQUIRE(n >= 1 && d >= 1, "itr_cosine_hinge_fwd_bwd_f32: bad shape");
  cudaStream_t st = as_stream(stream);
  {
    static int use_fused = -1;
    if (use_fused < 0) { const char* e = getenv("ITR_B200_VSE_STEP"); use_fused = (e && e[0] == 'm') ? 0 : 1; }    // "multi" = round-1 path
    if (use_fused) {
      int status = ITR_OK;
      const int done = vse_step_fused(im, s, n, d, margin, max_violation, ws, loss, d_im, d_s, st, &status);
      if (done < 0) return status;
      if (done > 0) return ITR_OK;
    }
  }
  float* S = ws;
  float* dS = ws + (int64_t)n * n;
  const bool need_grad = d_im != nullptr || d_s != nullptr;
  int rc = launch_sgemm(im, d, 1, s, d, 1, S, n, n, n, d, st);
  if (rc) return rc;
  rc = itr_hinge_fwd_bwd_f32(S, n, n, margin, max_violation, loss, need_grad ? dS : nullptr, n, stream);
  if (rc) return rc;
  // d_im[i][k] = sum_j dS[i][j] s[j][k]   ;   d_s[j][k] = sum_i dS[i][j] im[i][k]
  if (d_im) { rc = launch_sgemm(dS, n, 1, s, 1, d, d_im, d, n, d, n, st); if (rc) return rc; }
  if (d_s) { rc = launch_sgemm(dS, 1, n, im, 1, d, d_s, d, n, d, n, st); if (rc) return rc; }
  return ITR_OK;
}

extern "C" int itr_rank_thresholds_f32(const float* scores, int64_t ld_scores, int n_img, int n_cap_local,
                                       int cap_offset, int caps_per_img, float* thr_row, float* thr_col, void* stream) {
  ITR_REQUIRE(scores && thr_row && thr_col, "itr_rank_thresholds_f32: null pointer");
  ITR_REQUIRE(n_img >= 1 && n_cap_local >= 1 && caps_per_img >= 1 && cap_offset >= 0 && ld_scores >= n_cap_local,
              "itr_rank_thresholds_f32: bad shape");
  int n = n_img > n_cap_local ? n_img : n_cap_local;
  rank_thresholds_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(scores, ld_scores, n_img, n_cap_local, cap_offset,
                                                                        caps_per_img, thr_row, thr_col);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_rank_count_f32(const float* scores, int64_t ld_scores, int n_img, int n_cap_local, int cap_offset,
                                  const float* thr_row, const float* thr_col, int32_t* cnt_row, int32_t* cnt_col,
                                  uint64_t* best_row, uint64_t* best_col, void* stream) {
  ITR_REQUIRE(scores && thr_row && thr_col && cnt_row && cnt_col && best_row && best_col, "itr_rank_count_f32: null pointer");
  ITR_REQUIRE(n_img >= 1 && n_cap_local >= 1 && ld_scores >= n_cap_local, "itr_rank_count_f32: bad shape");
  cudaStream_t st = as_stream(stream);
  ITR_CHECK_CUDA(cudaMemsetAsync(cnt_row, 0, sizeof(int32_t) * n_img, st));
  ITR_CHECK_CUDA(cudaMemsetAsync(cnt_col, 0, sizeof(int32_t) * n_cap_local, st));
  ITR_CHECK_CUDA(cudaMemsetAsync(best_row, 0, sizeof(uint64_t) * n_img, st));
  ITR_CHECK_CUDA(cudaMemsetAsync(best_col, 0, sizeof(uint64_t) * n_cap_local, st));
  dim3 grid((n_cap_local + 255) / 256, (n_img + RK_ROWS - 1) / RK_ROWS);
  rank_count_kernel<<<grid, 256, 0, st>>>(scores, ld_scores, n_img, n_cap_local, cap_offset, thr_row, thr_col, cnt_row,
                                          cnt_col, reinterpret_cast<unsigned long long*>(best_row),
                                          reinterpret_cast<unsigned long long*>(best_col));
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_rank_f64(const double* scores, int64_t ld_scores, int n_img, int n_cap, int caps_per_img,
                            int32_t* rank_row, int32_t* rank_col, int32_t* top1_row, int32_t* top1_col, void* stream) {
  ITR_REQUIRE(scores && rank_row && rank_col && top1_row && top1_col, "itr_rank_f64: null pointer");
  ITR_REQUIRE(n_img >= 1 && n_cap >= 1 && caps_per_img >= 1 && ld_scores >= n_cap, "itr_rank_f64: bad shape");
  cudaStream_t st = as_stream(stream);
  double *thr_row = nullptr, *thr_col = nullptr;
  unsigned long long *best_row = nullptr, *best_col = nullptr;
  keep_scratch_pool_warm();
  ITR_CHECK_CUDA(cudaMallocAsync(&thr_row, sizeof(double) * (n_img + n_cap), st));
  thr_col = thr_row + n_img;
  ITR_CHECK_CUDA(cudaMallocAsync(&best_row, sizeof(unsigned long long) * (n_img + n_cap), st));
  best_col = best_row + n_img;
  cudaMemsetAsync(best_row, 0, sizeof(unsigned long long) * (n_img + n_cap), st);
  cudaMemsetAsync(rank_row, 0, sizeof(int32_t) * n_img, st);
  cudaMemsetAsync(rank_col, 0, sizeof(int32_t) * n_cap, st);
  cudaMemsetAsync(top1_row, 0x7f, sizeof(int32_t) * n_img, st);
  cudaMemsetAsync(top1_col, 0x7f, sizeof(int32_t) * n_cap, st);
  int n = n_img > n_cap ? n_img : n_cap;
  rank64_thresholds_kernel<<<(n + 255) / 256, 256, 0, st>>>(scores, ld_scores, n_img, n_cap, caps_per_img, thr_row, thr_col);
  dim3 grid((n_cap + 255) / 256, (n_img + RK_ROWS - 1) / RK_ROWS);
  rank64_pass_kernel<false><<<grid, 256, 0, st>>>(scores, ld_scores, n_img, n_cap, thr_row, thr_col, rank_row, rank_col,
                                                  best_row, best_col, top1_row, top1_col);
  rank64_pass_kernel<true><<<grid, 256, 0, st>>>(scores, ld_scores, n_img, n_cap, thr_row, thr_col, rank_row, rank_col,
                                                 best_row, best_col, top1_row, top1_col);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(thr_row, st);
  cudaFreeAsync(best_row, st);
  ITR_CHECK_CUDA(e);
  return ITR_OK;
}

extern "C" int itr_scan_caption_gram_f32(const uint16_t* words_bf16, const int32_t* cap_row0, const int32_t* cap_lens,
                                         const int64_t* gram_off, int n_cap, int d, float* gram, void* stream) {
  ITR_REQUIRE(words_bf16 && cap_row0 && cap_lens && gram_off && gram, "itr_scan_caption_gram_f32: null pointer");
  ITR_REQUIRE(d > 0 && d % 2 == 0, "itr_scan_caption_gram_f32: embed size must be even");
  if (n_cap <= 0) return ITR_OK;
  // about four blocks per SM in total: many captions -> one block each, few captions -> their pairs spread over blocks
  int split = (4 * 148 + n_cap - 1) / n_cap;
  split = split < 1 ? 1 : (split > 64 ? 64 : split);
  caption_gram_kernel<<<dim3((unsigned)n_cap, (unsigned)split), 128, 0, as_stream(stream)>>>(words_bf16, cap_row0, cap_lens, gram_off, d, gram);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_scan_epilogue_f32(const float* affinity, int n_img, const int32_t* cap_row0, const int32_t* cap_lens,
                                     const int32_t* cap_ids, int n_ids, int max_len, const float* row_wnorm,
                                     const float* region_norm, const float* region_gram, const float* cap_gram,
                                     const int64_t* gram_off, int cross_attn, int feature_norm, int agg,
                                     float lambda_softmax, float lambda_lse, float* scores, int64_t ld_scores, void* stream) {
  ITR_REQUIRE(affinity && cap_row0 && cap_lens && row_wnorm && region_norm && scores, "itr_scan_epilogue_f32: null pointer");
  ITR_REQUIRE(cross_attn == ITR_T2I || cross_attn == ITR_I2T, "unknown cross_attn: %d", cross_attn);
  ITR_REQUIRE(feature_norm >= 0 && feature_norm <= ITR_NORM_NONE, "unknown first norm type: %d", feature_norm);
  ITR_REQUIRE(agg >= 0 && agg <= ITR_AGG_SUM, "unknown aggfunc: %d", agg);
  ITR_REQUIRE(cross_attn == ITR_I2T ? (cap_gram && gram_off) : (region_gram != nullptr), "itr_scan_epilogue_f32: missing Gram input");
  ITR_REQUIRE(max_len >= 1 && max_len <= ITR_TILE_WORDS && n_ids >= 0, "itr_scan_epilogue_f32: bad shape");
  if (n_img <= 0 || n_ids <= 0) return ITR_OK;
  const int R = ITR_REGIONS, LP = (max_len + 1) | 1;      // odd row pitch: the i2t accesses stride by LP across lanes
  auto smem_for = [&](int imgs) {
    const int RT = imgs * R;
    const int g_floats = (cross_attn == ITR_T2I) ? imgs * R * R : LP * LP;
    const int rs = R > LP ? R : LP;
    return sizeof(float) * (2 * (size_t)RT * LP + g_floats + LP + RT + (size_t)imgs * rs);
  };
  if (cross_attn == ITR_I2T && max_len <= 32) {        // the specialised kernel: words in registers
    ScanEpiKernelParams q{affinity, n_img, cap_row0, cap_lens, cap_ids, n_ids, 0, I2T_IMGS, row_wnorm, region_norm, region_gram, cap_gram,
                          gram_off, ScanEpiParams{R, cross_attn, feature_norm, agg, lambda_softmax, lambda_lse, scores, ld_scores}};
    cudaStream_t st = as_stream(stream);
    if (max_len <= 8) return launch_i2t_epilogue<8>(q, n_ids, n_img, st);
    if (max_len <= 12) return launch_i2t_epilogue<12>(q, n_ids, n_img, st);
    if (max_len <= 16) return launch_i2t_epilogue<16>(q, n_ids, n_img, st);
    if (max_len <= 24) return launch_i2t_epilogue<24>(q, n_ids, n_img, st);
    return launch_i2t_epilogue<32>(q, n_ids, n_img, st);
  }
  int imgs = SF_IMGS;
  if (cross_attn == ITR_I2T && smem_for(7) <= 110 * 1024) imgs = 7;
  const size_t smem = smem_for(imgs);
  ScanEpiKernelParams p{affinity, n_img, cap_row0, cap_lens, cap_ids, n_ids, LP, imgs, row_wnorm, region_norm, region_gram, cap_gram, gram_off,
                        ScanEpiParams{R, cross_attn, feature_norm, agg, lambda_softmax, lambda_lse, scores, ld_scores}};
  ITR_REQUIRE(smem <= 227 * 1024, "itr_scan_epilogue_f32: caption of %d words needs %zu bytes of shared memory", max_len, smem);
  ITR_CHECK_CUDA(cudaFuncSetAttribute(scan_epilogue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_ids, (n_img + imgs - 1) / imgs);
  ITR_REQUIRE(grid.y <= 65535, "itr_scan_epilogue_f32: more than %d images per call", 65535 * imgs);
  scan_epilogue_kernel<<<grid, cross_attn == ITR_I2T ? 256 : 128, smem, as_stream(stream)>>>(p);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

// =========================================================================================
// the two remaining dense similarity measures.
//   CAMERA MultiViewMatching (Fusionmodule.py:670-692): score[i][c] = max_v  imgs[i][v] . caps[c]
//   order_sim (Objectives.py:24-30):                    score[i][c] = -|max(s_c - im_i, 0)|_2
// =========================================================================================
namespace itr {

// C[(i, v)][c] -> score[i][c] = max_v, arg[i][c] = first v attaining it
__global__ void __launch_bounds__(256)
view_max_kernel(const float* __restrict__ C, int n_img, int n_views, int n_cap, float* __restrict__ scores, int64_t ld,
                int32_t* __restrict__ arg) {
  const int c = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
  if (c >= n_cap) return;
  const float* src = C + (int64_t)i * n_views * n_cap + c;
  float best = src[0];
  int bv = 0;
  for (int v = 1; v < n_views; ++v) {
    const float x = src[(int64_t)v * n_cap];
    if (x > best) { best = x; bv = v; }
  }
  scores[(int64_t)i * ld + c] = best;
  if (arg) arg[(int64_t)i * n_cap + c] = bv;
}

// E[(i, v)][c] = dS[i][c] if v == arg[i][c] else 0
__global__ void __launch_bounds__(256)
view_scatter_kernel(const float* __restrict__ dS, int64_t ld, const int32_t* __restrict__ arg, int n_img, int n_views,
                    int n_cap, float* __restrict__ E) {
  const int c = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
  if (c >= n_cap) return;
  const float g = dS[(int64_t)i * ld + c];
  const int a = arg[(int64_t)i * n_cap + c];
  for (int v = 0; v < n_views; ++v) E[((int64_t)i * n_views + v) * n_cap + c] = (v == a) ? g : 0.f;
}

// order_sim forward: 32 x 32 scores per block, 2 x 2 per thread, K swept through shared memory in slabs of 32.
__global__ void __launch_bounds__(256)
order_scores_kernel(const float* __restrict__ im, const float* __restrict__ s, int n_img, int n_cap, int d,
                    float* __restrict__ scores, int64_t ld) {
  __shared__ float Is[32][33], Ss[32][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  float acc[2][2] = {};
  for (int k0 = 0; k0 < d; k0 += 32) {
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      const int r = e >> 5, k = e & 31;
      Is[r][k] = (i0 + r < n_img && k0 + k < d) ? im[(int64_t)(i0 + r) * d + k0 + k] : 0.f;
      Ss[r][k] = (c0 + r < n_cap && k0 + k < d) ? s[(int64_t)(c0 + r) * d + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float a0 = Is[ty * 2][k], a1 = Is[ty * 2 + 1][k], b0 = Ss[tx * 2][k], b1 = Ss[tx * 2 + 1][k];
      float y;
      y = fmaxf(b0 - a0, 0.f); acc[0][0] = fmaf(y, y, acc[0][0]);
      y = fmaxf(b1 - a0, 0.f); acc[0][1] = fmaf(y, y, acc[0][1]);
      y = fmaxf(b0 - a1, 0.f); acc[1][0] = fmaf(y, y, acc[1][0]);
      y = fmaxf(b1 - a1, 0.f); acc[1][1] = fmaf(y, y, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = i0 + ty * 2 + a, c = c0 + tx * 2 + b;
      if (i < n_img && c < n_cap) scores[(int64_t)i * ld + c] = -sqrtf(acc[a][b]);
    }
}

// order_sim backward.  With y = max(s_c - im_i, 0), N = |y| = -score and w = dS / N (0 where N = 0):
//   d_im[i] = +sum_c w[i][c] y        d_s[c] = -sum_i w[i][c] y
// ROWS_ARE_IMAGES: block (row i, 256-wide slab of d) loops over all captions; otherwise block (row c) loops over images.
template <bool ROWS_ARE_IMAGES>
__global__ void __launch_bounds__(256)
order_backward_kernel(const float* __restrict__ im, const float* __restrict__ s, const float* __restrict__ scores, int64_t ld_s,
                      const float* __restrict__ dS, int64_t ld_ds, int n_img, int n_cap, int d, float* __restrict__ out) {
  const int row = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
  const int n_other = ROWS_ARE_IMAGES ? n_cap : n_img;
  __shared__ float w[256];
  const float mine = (k < d) ? (ROWS_ARE_IMAGES ? im : s)[(int64_t)row * d + k] : 0.f;
  float acc = 0.f;
  for (int o0 = 0; o0 < n_other; o0 += 256) {
    const int o = o0 + threadIdx.x;
    float wv = 0.f;
    if (o < n_other) {
      const int i = ROWS_ARE_IMAGES ? row : o, c = ROWS_ARE_IMAGES ? o : row;
      const float nrm = -scores[(int64_t)i * ld_s + c];
      wv = nrm > 0.f ? dS[(int64_t)i * ld_ds + c] / nrm : 0.f;
    }
    __syncthreads();
    w[threadIdx.x] = wv;
    __syncthreads();
    if (k < d) {
      const int lim = min(256, n_other - o0);
      const float* other = (ROWS_ARE_IMAGES ? s : im) + (int64_t)o0 * d + k;
      for (int j = 0; j < lim; ++j) {
        const float x = other[(int64_t)j * d];
        const float y = ROWS_ARE_IMAGES ? fmaxf(x - mine, 0.f) : fmaxf(mine - x, 0.f);
        acc = fmaf(w[j], y, acc);
      }
    }
  }
  if (k < d) out[(int64_t)row * d + k] = ROWS_ARE_IMAGES ? acc : -acc;
}

}  // namespace itr

extern "C" int itr_multiview_scores_f32(const float* imgs, const float* caps, int n_img, int n_views, int n_cap, int d,
                                        float* workspace, float* scores, int64_t ld_scores, int32_t* argmax, void* stream) {
  ITR_REQUIRE(imgs && caps && workspace && scores, "itr_multiview_scores_f32: null pointer");
  ITR_REQUIRE(n_img >= 0 && n_cap >= 0 && n_views >= 1 && d > 0 && ld_scores >= n_cap, "itr_multiview_scores_f32: bad shape");
  ITR_REQUIRE((int64_t)n_img * n_views < (1ll << 31) && n_img <= 65535, "itr_multiview_scores_f32: too many image views per call");
  if (n_img == 0 || n_cap == 0) return ITR_OK;
  cudaStream_t st = as_stream(stream);
  int rc = launch_sgemm(imgs, d, 1, caps, d, 1, workspace, n_cap, n_img * n_views, n_cap, d, st);
  if (rc) return rc;
  view_max_kernel<<<dim3((n_cap + 255) / 256, n_img), 256, 0, st>>>(workspace, n_img, n_views, n_cap, scores, ld_scores, argmax);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_multiview_backward_f32(const float* imgs, const float* caps, int n_img, int n_views, int n_cap, int d,
                                          const float* d_scores, int64_t ld_dscores, const int32_t* argmax, float* workspace,
                                          float* d_imgs, float* d_caps, void* stream) {
  ITR_REQUIRE(imgs && caps && d_scores && argmax && workspace, "itr_multiview_backward_f32: null pointer");
  ITR_REQUIRE(n_img >= 0 && n_cap >= 0 && n_views >= 1 && d > 0 && ld_dscores >= n_cap, "itr_multiview_backward_f32: bad shape");
  ITR_REQUIRE((int64_t)n_img * n_views < (1ll << 31) && n_img <= 65535, "itr_multiview_backward_f32: too many image views per call");
  if (n_img == 0 || n_cap == 0) return ITR_OK;
  cudaStream_t st = as_stream(stream);
  const int rows = n_img * n_views;
  view_scatter_kernel<<<dim3((n_cap + 255) / 256, n_img), 256, 0, st>>>(d_scores, ld_dscores, argmax, n_img, n_views, n_cap, workspace);
  ITR_CHECK_LAUNCH();
  int rc = ITR_OK;
  // d_imgs[(i,v)][k] = sum_c E[(i,v)][c] caps[c][k];   d_caps[c][k] = sum_(i,v) E[(i,v)][c] imgs[(i,v)][k]
  if (d_imgs) { rc = launch_sgemm(workspace, n_cap, 1, caps, 1, d, d_imgs, d, rows, d, n_cap, st); if (rc) return rc; }
  if (d_caps) { rc = launch_sgemm(workspace, 1, n_cap, imgs, 1, d, d_caps, d, n_cap, d, rows, st); if (rc) return rc; }
  return ITR_OK;
}

// Score block -> the host's float64 matrix (evaluation.py:140: cal_sims returns float64), written straight into mapped
// page-locked host memory over PCIe.  One 128-thread block per SM with a handful of registers: it fits next to the
// persistent score CTA, so the block of caption chunk k leaves the device while chunk k+1 is being scored.
__global__ void __launch_bounds__(128, 16)
scores_to_host_f64_kernel(const float* __restrict__ src, int64_t ld_src, int n_rows, int n_cols, double* __restrict__ dst, int64_t ld_dst) {
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const float* s = src + (int64_t)row * ld_src;
    double* d = dst + (int64_t)row * ld_dst;
    int c = threadIdx.x;
    for (; c + 384 < n_cols; c += 512) {                 // four independent loads in flight per thread
      const float v0 = s[c], v1 = s[c + 128], v2 = s[c + 256], v3 = s[c + 384];
      d[c] = (double)v0; d[c + 128] = (double)v1; d[c + 256] = (double)v2; d[c + 384] = (double)v3;
    }
    for (; c < n_cols; c += 128) d[c] = (double)s[c];
  }
}

extern "C" int itr_scores_to_host_f64(const float* scores, int64_t ld_scores, int n_rows, int n_cols, double* host_mapped,
                                      int64_t ld_host, void* stream) {
  ITR_REQUIRE(scores && host_mapped, "itr_scores_to_host_f64: null pointer");
  ITR_REQUIRE(n_rows >= 0 && n_cols >= 0 && ld_scores >= n_cols && ld_host >= n_cols, "itr_scores_to_host_f64: bad shape");
  if (n_rows == 0 || n_cols == 0) return ITR_OK;
  int dev = 0, sms = 0;
  ITR_CHECK_CUDA(cudaGetDevice(&dev));
  ITR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  scores_to_host_f64_kernel<<<sms < n_rows ? sms : n_rows, 128, 0, as_stream(stream)>>>(scores, ld_scores, n_rows, n_cols, host_mapped, ld_host);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_order_scores_f32(const float* im, const float* s, int n_img, int n_cap, int d, float* scores,
                                    int64_t ld_scores, void* stream) {
  ITR_REQUIRE(im && s && scores, "itr_order_scores_f32: null pointer");
  ITR_REQUIRE(n_img >= 0 && n_cap >= 0 && d > 0 && ld_scores >= n_cap, "itr_order_scores_f32: bad shape");
  if (n_img == 0 || n_cap == 0) return ITR_OK;
  dim3 grid((n_cap + 31) / 32, (n_img + 31) / 32);
  ITR_REQUIRE(grid.y <= 65535, "itr_order_scores_f32: too many images per call");
  order_scores_kernel<<<grid, 256, 0, as_stream(stream)>>>(im, s, n_img, n_cap, d, scores, ld_scores);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}

extern "C" int itr_order_backward_f32(const float* im, const float* s, const float* scores, int64_t ld_scores,
                                      const float* d_scores, int64_t ld_dscores, int n_img, int n_cap, int d,
                                      float* d_im, float* d_s, void* stream) {
  ITR_REQUIRE(im && s && scores && d_scores, "itr_order_backward_f32: null pointer");
  ITR_REQUIRE(n_img >= 0 && n_cap >= 0 && d > 0 && ld_scores >= n_cap && ld_dscores >= n_cap, "itr_order_backward_f32: bad shape");
  ITR_REQUIRE(n_img <= 65535 && n_cap <= 65535, "itr_order_backward_f32: more than 65535 rows per call");
  if (n_img == 0 || n_cap == 0) return ITR_OK;
  cudaStream_t st = as_stream(stream);
  if (d_im) order_backward_kernel<true><<<dim3((d + 255) / 256, n_img), 256, 0, st>>>(im, s, scores, ld_scores, d_scores, ld_dscores, n_img, n_cap, d, d_im);
  if (d_s) order_backward_kernel<false><<<dim3((d + 255) / 256, n_cap), 256, 0, st>>>(im, s, scores, ld_scores, d_scores, ld_dscores, n_img, n_cap, d, d_s);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}
