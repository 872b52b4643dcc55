// SCAN training backward, float32.
//
// The reference back-propagates ContrastiveLoss through xattn_score_t2i / xattn_score_i2t with autograd
// (Models.py:219-222, Objectives.py:76-115, 329-476): per caption it re-materialises the (n_img, L, D) context
// tensors.  Here the gradient is brought into a closed form whose per-pair part needs only the raw affinity
// tile, the context Gram and dS (nothing D-wide), and whose D-wide part is two plain GEMMs over the whole batch:
//
//     dImages   = M   . Words   + [t2i: blockdiag(MC_i) . V_i   | i2t: diag(T) . V]
//     dCaptions = M^T . Regions + [t2i: diag(T) . W             | i2t: blockdiag(MC_c) . W_c]
//
// M[(i, region k)][(c, word j)] couples v_ik and w_cj (oracle/scan_backward.py derives and checks it against
// autograd).  Kernels:
//   scan_bwd_coeff_kernel   one block = one caption x 4 images: recompute the affinity tile (same phase 1 as the
//                           forward kernel), run the epilogue forward and backward in shared memory, write the
//                           block's slice of M and M^T and its partial T / MC.
//   leading_sum_kernel      deterministic reduction of the partial T / MC over captions or image groups.
//   sgemm_nn_kernel         C (+)= A . B (+ diag(t) . X), rows of B / C / X optionally indirected (packed words
//                           <-> padded caption rows).
//   blockdiag_apply_kernel  C_b += G_b . X_b for the small per-image / per-caption blocks.
#include <cfloat>

#include "common.cuh"
#include "scan_f32.cuh"

namespace itr {

struct ScanBwdParams {
  ScanF32Params f;                 // forward inputs (f.scores unused)
  const float* d_scores; int64_t ld_ds;
  const int32_t* cap_off;          // [n_cap] first packed word of caption c
  const int64_t* gram_off;         // [n_cap] offset of caption c in the packed n_c x n_c array (i2t)
  float* M; int64_t ldm;           // [n_img*R][ldm]   columns = packed words
  float* MT; int64_t ldmt;         // [n_words][ldmt]  columns = image-major regions
  float* Tpart;                    // t2i: [n_groups][n_words]     i2t: [n_cap][n_img*R]
  float* MCpart;                   // t2i: [n_cap][n_img*R*R]      i2t: [n_groups][sum n_c^2]
  int64_t n_words, sum_n2;
};

constexpr int SB_RS = SF_LMAX;     // per-image stride of the small per-query / per-source arrays (>= R, >= lmax)

__global__ void __launch_bounds__(256)
scan_bwd_coeff_kernel(ScanBwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int R = p.f.R, RT = SF_IMGS * R, LP = sf_pitch(p.f.lmax);
  const bool t2i_ = (p.f.cross_attn == ITR_T2I);
  // Vs / Ws live only during phase 1 and Y only after it: they share the first region.
  const int head = max(SF_VS_FLOATS + SF_WS_FLOATS, RT * LP);
  float* Vs = smem;                            // SF_VS_FLOATS (phase 1 only)
  float* Ws = Vs + SF_VS_FLOATS;               // SF_WS_FLOATS (phase 1 only)
  float* Y = smem;                             // RT*LP  G alpha, then d xh, then d a
  float* Araw = smem + head;                   // RT*LP  raw affinities [img*R + region][word]
  float* X = Araw + RT * LP;                   // RT*LP  xh, then alpha
  float* Gctx = X + RT * LP;                   // t2i: SF_IMGS*R*R; i2t: LP*LP
  const int g_floats = t2i_ ? SF_IMGS * R * R : LP * LP;
  float* wnorm = Gctx + g_floats;              // SF_LMAX
  float* vnorm = wnorm + SF_LMAX;              // RT
  float* rsim = vnorm + RT;                    // SF_IMGS*SB_RS   r_q
  float* qfv = rsim + SF_IMGS * SB_RS;         // |ctx_q|^2
  float* pq = qfv + SF_IMGS * SB_RS;           // d r / d(query . ctx)   (times g)
  float* uq = pq + SF_IMGS * SB_RS;            // coefficient of ctx in d r / d ctx
  float* tq = uq + SF_IMGS * SB_RS;            // coefficient of query in d r / d query
  float* st0 = tq + SF_IMGS * SB_RS;           // per (image, source): norm / row max
  float* st1 = st0 + SF_IMGS * SB_RS;          // per (image, source): sqrt(sum sq) / softmax denominator

  const int c = blockIdx.x;
  const int img0 = blockIdx.y * SF_IMGS;
  const int n_im = min(SF_IMGS, p.f.n_img - img0);
  const int n = p.f.cap_lens[c];
  const float* W = p.f.captions + (int64_t)c * p.f.lmax * p.f.d;
  const int tid = threadIdx.x;
  const bool t2i = (p.f.cross_attn == ITR_T2I);
  const int mode = p.f.feature_norm;
  const float lam = p.f.lambda_softmax;

  if (t2i) {
    for (int e = tid; e < n_im * R * R; e += 256) Gctx[e] = p.f.gram[(int64_t)img0 * R * R + e];
  } else {
    for (int e = tid; e < LP * LP; e += 256) Gctx[e] = 0.f;
  }
  __syncthreads();
  const int cpt = (n + 15) / 16;
  switch (cpt) {
    case 1: scan_f32_gemm<1>(p.f, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
    case 2: scan_f32_gemm<2>(p.f, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
    case 3: scan_f32_gemm<3>(p.f, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
    case 4: scan_f32_gemm<4>(p.f, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
    case 5: scan_f32_gemm<5>(p.f, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
    default: scan_f32_gemm<6>(p.f, img0, n_im * R, W, n, LP, Vs, Ws, Araw, wnorm, vnorm, Gctx); break;
  }
  __syncthreads();

  const int S = t2i ? R : n, Q = t2i ? n : R;
  auto at = [&](float* base, int m, int s, int q) -> float& {
    return t2i ? base[(m * R + s) * LP + q] : base[(m * R + q) * LP + s];
  };
  const bool clip = (mode == ITR_NORM_CLIPPED_L2 || mode == ITR_NORM_CLIPPED);
  const bool l2 = (mode == ITR_NORM_CLIPPED_L2 || mode == ITR_NORM_L2);

  // ---- forward 1: raw_feature_norm over q for every (image, source) ------------------------
  for (int it = tid; it < n_im * S; it += 256) {
    const int m = it / S, s = it % S;
    if (l2) {
      float ss = 0.f;
      for (int q = 0; q < Q; ++q) {
        float a = at(Araw, m, s, q);
        if (clip) a = leaky01(a);
        ss = fmaf(a, a, ss);
      }
      const float rs = sqrtf(ss), nrm = rs + 1e-8f, inv = 1.f / nrm;
      st0[m * SB_RS + s] = nrm;
      st1[m * SB_RS + s] = rs;
      for (int q = 0; q < Q; ++q) {
        float a = at(Araw, m, s, q);
        if (clip) a = leaky01(a);
        at(X, m, s, q) = a * inv;
      }
    } else if (mode == ITR_NORM_SOFTMAX) {
      float mx = -FLT_MAX;
      for (int q = 0; q < Q; ++q) mx = fmaxf(mx, at(Araw, m, s, q));
      float z = 0.f;
      for (int q = 0; q < Q; ++q) z += expf(at(Araw, m, s, q) - mx);
      const float inv = 1.f / z;
      for (int q = 0; q < Q; ++q) at(X, m, s, q) = expf(at(Araw, m, s, q) - mx) * inv;
    } else {
      for (int q = 0; q < Q; ++q) {
        float a = at(Araw, m, s, q);
        at(X, m, s, q) = clip ? leaky01(a) : a;
      }
    }
  }
  __syncthreads();

  // ---- forward 2: alpha = softmax over s, G alpha, attended cosine -------------------------
  for (int it = tid; it < n_im * Q; it += 256) {
    const int m = it / Q, q = it % Q;
    float mx = -FLT_MAX;
    for (int s = 0; s < S; ++s) mx = fmaxf(mx, at(X, m, s, q) * lam);
    float Z = 0.f;
    for (int s = 0; s < S; ++s) {
      float e = expf(at(X, m, s, q) * lam - mx);
      at(X, m, s, q) = e;
      Z += e;
    }
    const float invZ = 1.f / Z;
    float P = 0.f;
    for (int s = 0; s < S; ++s) {
      float al = at(X, m, s, q) * invZ;
      at(X, m, s, q) = al;
      P = fmaf(al, at(Araw, m, s, q), P);
    }
    const float* G = t2i ? Gctx + m * R * R : Gctx;
    const int gs = t2i ? R : LP;
    float Qf = 0.f;
    for (int s = 0; s < S; ++s) {
      float u = 0.f;
      for (int s2 = 0; s2 < S; ++s2) u = fmaf(G[s * gs + s2], at(X, m, s2, q), u);
      at(Y, m, s, q) = u;
      Qf = fmaf(at(X, m, s, q), u, Qf);
    }
    const float qn = t2i ? wnorm[q] : vnorm[m * R + q];
    const float den = qn * sqrtf(fmaxf(Qf, 0.f));
    rsim[m * SB_RS + q] = P / fmaxf(den, 1e-8f);
    qfv[m * SB_RS + q] = Qf;
  }
  __syncthreads();

  // ---- backward 1: d score / d r_q times dS, one warp per image ------------------------------
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < n_im) {
    const float* r = rsim + warp * SB_RS;
    float* g = pq + warp * SB_RS;               // staged in pq, consumed by the next step
    const float ds = p.d_scores[(int64_t)(img0 + warp) * p.ld_ds + c];
    if (p.f.agg == ITR_AGG_LSE) {
      float mx = -FLT_MAX;
      for (int q = lane; q < Q; q += 32) mx = fmaxf(mx, r[q]);
      mx = warp_max(mx);
      float z = 0.f;
      for (int q = lane; q < Q; q += 32) z += expf((r[q] - mx) * p.f.lambda_lse);
      z = warp_sum(z);
      for (int q = lane; q < Q; q += 32) g[q] = ds * expf((r[q] - mx) * p.f.lambda_lse) / z;
    } else if (p.f.agg == ITR_AGG_MAX) {
      float mx = -FLT_MAX;
      for (int q = lane; q < Q; q += 32) mx = fmaxf(mx, r[q]);
      mx = warp_max(mx);
      int first = 1 << 30;
      for (int q = lane; q < Q; q += 32) if (r[q] == mx) first = min(first, q);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      for (int q = lane; q < Q; q += 32) g[q] = (q == first) ? ds : 0.f;
    } else {
      const float w = (p.f.agg == ITR_AGG_MEAN) ? ds / (float)Q : ds;
      for (int q = lane; q < Q; q += 32) g[q] = w;
    }
  }
  __syncthreads();

  // ---- backward 2: cosine and softmax-over-s, for every (image, query) -----------------------
  for (int it = tid; it < n_im * Q; it += 256) {
    const int m = it / Q, q = it % Q;
    const float g = pq[m * SB_RS + q], r = rsim[m * SB_RS + q], Qf = qfv[m * SB_RS + q];
    const float qn = t2i ? wnorm[q] : vnorm[m * R + q];
    const float den = qn * sqrtf(fmaxf(Qf, 0.f));
    float pc, uc, tc;
    if (den < 1e-8f) { pc = g * 1e8f; uc = 0.f; tc = 0.f; }
    else { pc = g / den; uc = -g * r / Qf; tc = -g * r / (qn * qn); }
    pq[m * SB_RS + q] = pc;
    uq[m * SB_RS + q] = uc;
    tq[m * SB_RS + q] = tc;
    float dot = 0.f;
    for (int s = 0; s < S; ++s) {
      float dal = fmaf(pc, at(Araw, m, s, q), uc * at(Y, m, s, q));
      at(Y, m, s, q) = dal;
      dot = fmaf(at(X, m, s, q), dal, dot);
    }
    for (int s = 0; s < S; ++s) at(Y, m, s, q) = lam * at(X, m, s, q) * (at(Y, m, s, q) - dot);
  }
  __syncthreads();

  // ---- backward 3: raw_feature_norm over q, for every (image, source): Y <- d a --------------
  for (int it = tid; it < n_im * S; it += 256) {
    const int m = it / S, s = it % S;
    if (l2) {
      const float nrm = st0[m * SB_RS + s], rs = st1[m * SB_RS + s];
      float proj = 0.f;
      for (int q = 0; q < Q; ++q) {
        float a = at(Araw, m, s, q);
        if (clip) a = leaky01(a);
        proj = fmaf(at(Y, m, s, q), a, proj);
      }
      const float inv = 1.f / nrm, k2 = proj / (nrm * nrm * fmaxf(rs, 1e-30f));
      for (int q = 0; q < Q; ++q) {
        const float a = at(Araw, m, s, q);
        const float l = clip ? leaky01(a) : a;
        float dl = at(Y, m, s, q) * inv - l * k2;
        if (clip && !(a > 0.f)) dl *= 0.1f;
        at(Y, m, s, q) = dl;
      }
    } else if (mode == ITR_NORM_SOFTMAX) {
      // xh is gone (X holds alpha): rebuild it from the raw affinities
      float mx = -FLT_MAX;
      for (int q = 0; q < Q; ++q) mx = fmaxf(mx, at(Araw, m, s, q));
      float z = 0.f;
      for (int q = 0; q < Q; ++q) z += expf(at(Araw, m, s, q) - mx);
      const float inv = 1.f / z;
      float dot = 0.f;
      for (int q = 0; q < Q; ++q) dot = fmaf(expf(at(Araw, m, s, q) - mx) * inv, at(Y, m, s, q), dot);
      for (int q = 0; q < Q; ++q) {
        const float xh = expf(at(Araw, m, s, q) - mx) * inv;
        at(Y, m, s, q) = xh * (at(Y, m, s, q) - dot);
      }
    } else if (clip) {
      for (int q = 0; q < Q; ++q)
        if (!(at(Araw, m, s, q) > 0.f)) at(Y, m, s, q) *= 0.1f;
    }
  }
  __syncthreads();

  // ---- outputs ----------------------------------------------------------------------------------
  // MQ(region k, word j) = p_q alpha + d a, with q = word (t2i) or region (i2t); X/Y are [img*R + region][word].
  const int64_t col0 = p.cap_off[c];
  for (int e = tid; e < n_im * R * n; e += 256) {                 // M: word fastest
    const int j = e % n, row = e / n, m = row / R, k = row % R;
    const float pc = pq[m * SB_RS + (t2i ? j : k)];
    p.M[((int64_t)img0 * R + row) * p.ldm + col0 + j] = fmaf(pc, X[row * LP + j], Y[row * LP + j]);
  }
  const int rows_blk = n_im * R;
  for (int e = tid; e < n * rows_blk; e += 256) {                 // M^T: region fastest
    const int row = e % rows_blk, j = e / rows_blk, m = row / R, k = row % R;
    const float pc = pq[m * SB_RS + (t2i ? j : k)];
    p.MT[(col0 + j) * p.ldmt + (int64_t)img0 * R + row] = fmaf(pc, X[row * LP + j], Y[row * LP + j]);
  }
  if (t2i) {
    for (int j = tid; j < n; j += 256) {
      float s = 0.f;
      for (int m = 0; m < n_im; ++m) s += tq[m * SB_RS + j];
      p.Tpart[(int64_t)blockIdx.y * p.n_words + col0 + j] = s;
    }
    float* mc = p.MCpart + ((int64_t)c * p.f.n_img + img0) * R * R;
    for (int e = tid; e < n_im * R * R; e += 256) {
      const int m = e / (R * R), s = (e / R) % R, s2 = e % R;
      float acc = 0.f;
      for (int q = 0; q < Q; ++q) acc = fmaf(at(X, m, s, q) * uq[m * SB_RS + q], at(X, m, s2, q), acc);
      mc[e] = acc;
    }
  } else {
    for (int e = tid; e < n_im * R; e += 256)
      p.Tpart[(int64_t)c * p.f.n_img * R + (int64_t)img0 * R + e] = tq[(e / R) * SB_RS + (e % R)];
    float* mc = p.MCpart + (int64_t)blockIdx.y * p.sum_n2 + p.gram_off[c];
    for (int e = tid; e < n * n; e += 256) {
      const int s = e / n, s2 = e % n;
      float acc = 0.f;
      for (int m = 0; m < n_im; ++m)
        for (int q = 0; q < Q; ++q) acc = fmaf(at(X, m, s, q) * uq[m * SB_RS + q], at(X, m, s2, q), acc);
      mc[e] = acc;
    }
  }
}

// cap_off / gram_off (exclusive prefix sums of len and len^2) and word_row (packed word -> padded row) from the
// caption lengths, one block; n_cap is a training batch, so the serial prefix is a few microseconds.
__global__ void __launch_bounds__(256)
caption_index_kernel(const int32_t* __restrict__ cap_lens, int n_cap, int lmax, int32_t* __restrict__ cap_off,
                     int64_t* __restrict__ gram_off, int32_t* __restrict__ word_row) {
  if (threadIdx.x == 0) {
    int32_t o = 0;
    int64_t g = 0;
    for (int c = 0; c < n_cap; ++c) {
      cap_off[c] = o;
      gram_off[c] = g;
      const int n = cap_lens[c];
      o += n;
      g += (int64_t)n * n;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n_cap; c += 256) {
    const int n = cap_lens[c], o = cap_off[c];
    for (int j = 0; j < n; ++j) word_row[o + j] = c * lmax + j;
  }
}

// out[e] = sum_g in[g * n + e], fixed order.
__global__ void __launch_bounds__(256)
leading_sum_kernel(const float* __restrict__ in, int64_t n, int groups, float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n) return;
  float s = 0.f;
  for (int g = 0; g < groups; ++g) s += in[(int64_t)g * n + e];
  out[e] = s;
}

// C[row(m)][n] (+)= sum_k A[m][k] * B[brow(k)][n]  (+ tvec[m] * Xm[row(m)][n]);  64 x 64 x 16 tile, 4 x 4 per thread.
// brow / crow: optional row indirections (nullptr = identity).  N and ldb / ldc / ldx must be multiples of 4.
struct SgemmNN {
  const float* A; int64_t lda;
  const float* B; int64_t ldb; const int32_t* brow;
  float* C; int64_t ldc; const int32_t* crow;
  const float* tvec; const float* Xm; int64_t ldx;
  int M, N, K, accumulate;
};

__global__ void __launch_bounds__(256)
sgemm_nn_kernel(SgemmNN g) {
  constexpr int TB = 64, KB = 16;
  __shared__ __align__(16) float As[KB][TB + 4];
  __shared__ __align__(16) float Bs[KB][TB + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TB, n0 = blockIdx.x * TB;
  float acc[4][4] = {};
  const int am = tid >> 2, ak = (tid & 3) * 4;        // A tile: 64 rows x 16 k, 4 consecutive k per thread
  const int bk = tid >> 4, bn = (tid & 15) * 4;       // B tile: 16 k x 64 n, float4 per thread
  const bool a_vec = (g.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  for (int k0 = 0; k0 < g.K; k0 += KB) {
    {
      const int gm = m0 + am, gk = k0 + ak;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gm < g.M) {
        const float* src = g.A + (int64_t)gm * g.lda + gk;
        if (a_vec && gk + 3 < g.K) {
          const float4 t = *reinterpret_cast<const float4*>(src);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (gk + i < g.K) v[i] = src[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[ak + i][am] = v[i];
    }
    {
      const int gk = k0 + bk, gn = n0 + bn;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gk < g.K && gn < g.N) {
        const int64_t row = g.brow ? g.brow[gk] : gk;
        t = *reinterpret_cast<const float4*>(g.B + row * g.ldb + gn);
      }
      *reinterpret_cast<float4*>(&Bs[bk][bn]) = t;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int gn = n0 + tx * 4;
  if (gn >= g.N) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
    const int64_t row = g.crow ? g.crow[gm] : gm;
    float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (g.tvec) {
      const float t = g.tvec[gm];
      const float4 x = *reinterpret_cast<const float4*>(g.Xm + row * g.ldx + gn);
      o.x = fmaf(t, x.x, o.x); o.y = fmaf(t, x.y, o.y); o.z = fmaf(t, x.z, o.z); o.w = fmaf(t, x.w, o.w);
    }
    float4* dst = reinterpret_cast<float4*>(g.C + row * g.ldc + gn);
    if (g.accumulate) { const float4 c = *dst; o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
    *dst = o;
  }
}

// C_b[a][:] += sum_b G_b[a][b] X_b[b][:]  for block b = blockIdx.x: rows row0 .. row0 + n_b of X / C (same row
// numbering, leading dimension ld), G_b = G + goff, n_b x n_b.  blockIdx.y picks a 128-column slab.
struct BlockDiag {
  const float* G; const float* X; float* C; int64_t ld; int d;
  const int32_t* lens; const int64_t* goff;     // per block; nullptr -> fixed size / stride
  int fixed_n; int64_t row_stride;              // rows of block b start at b * row_stride
};

__global__ void __launch_bounds__(128)
blockdiag_apply_kernel(BlockDiag p) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x, nb = p.lens ? p.lens[b] : p.fixed_n;
  const float* G = p.G + (p.goff ? p.goff[b] : (int64_t)b * p.fixed_n * p.fixed_n);
  float* Gs = smem;                  // nb*nb
  float* Xs = smem + nb * nb;        // nb*128
  const int col = blockIdx.y * 128 + threadIdx.x;
  const int64_t row0 = (int64_t)b * p.row_stride;
  for (int e = threadIdx.x; e < nb * nb; e += 128) Gs[e] = G[e];
  for (int r = 0; r < nb; ++r) Xs[r * 128 + threadIdx.x] = (col < p.d) ? p.X[(row0 + r) * p.ld + col] : 0.f;
  __syncthreads();
  if (col >= p.d) return;
  for (int a = 0; a < nb; ++a) {
    float s = 0.f;
    for (int r = 0; r < nb; ++r) s = fmaf(Gs[a * nb + r], Xs[r * 128 + threadIdx.x], s);
    p.C[(row0 + a) * p.ld + col] += s;
  }
}

static int64_t align64(int64_t floats) { return (floats + 63) / 64 * 64; }

struct BwdWorkspace {
  int64_t m, mt, tpart, tsum, mcpart, mcsum, cap_off, gram_off, word_row, total;   // offsets in floats (4-byte units)
  int64_t ldm;
  int n_groups;
};

static BwdWorkspace bwd_layout(int n_img, int R, int n_cap, int64_t n_words, int64_t sum_n2, int cross_attn) {
  BwdWorkspace w;
  w.n_groups = (n_img + SF_IMGS - 1) / SF_IMGS;
  w.ldm = (n_words + 3) / 4 * 4;
  const int64_t rows = (int64_t)n_img * R;
  int64_t o = 0;
  w.m = o; o += align64(rows * w.ldm);
  w.mt = o; o += align64(n_words * rows);
  if (cross_attn == ITR_T2I) {
    w.tpart = o; o += align64((int64_t)w.n_groups * n_words);
    w.tsum = o; o += align64(n_words);
    w.mcpart = o; o += align64((int64_t)n_cap * rows * R);
    w.mcsum = o; o += align64(rows * R);
  } else {
    w.tpart = o; o += align64((int64_t)n_cap * rows);
    w.tsum = o; o += align64(rows);
    w.mcpart = o; o += align64((int64_t)w.n_groups * sum_n2);
    w.mcsum = o; o += align64(sum_n2);
  }
  w.cap_off = o; o += align64(n_cap);
  w.gram_off = o; o += align64(2 * (int64_t)n_cap);
  w.word_row = o; o += align64(n_words);
  w.total = o;
  return w;
}

}  // namespace itr

using namespace itr;

extern "C" int64_t itr_scan_backward_workspace_f32(int n_img, int n_regions, int n_cap, int64_t n_words,
                                                   int64_t sum_len_sq, int cross_attn) {
  if (n_img < 0 || n_regions <= 0 || n_cap < 0 || n_words < 0 || sum_len_sq < 0) return -1;
  return bwd_layout(n_img, n_regions, n_cap, n_words, sum_len_sq, cross_attn).total * (int64_t)sizeof(float);
}

extern "C" int itr_scan_backward_f32(const float* images, const float* gram, const float* captions,
                                     const int32_t* cap_lens, int n_img, int n_regions, int n_cap, int lmax, int d,
                                     int64_t n_words, int64_t sum_len_sq, int cross_attn, int feature_norm, int agg,
                                     float lambda_softmax, float lambda_lse, const float* d_scores, int64_t ld_dscores,
                                     float* d_images, float* d_captions, void* workspace, int64_t workspace_bytes,
                                     void* stream) {
  ITR_REQUIRE(images && captions && cap_lens && d_scores && d_images && d_captions && workspace,
              "itr_scan_backward_f32: null pointer");
  ITR_REQUIRE(cross_attn == ITR_T2I || cross_attn == ITR_I2T, "unknown cross_attn: %d", cross_attn);
  ITR_REQUIRE(feature_norm >= 0 && feature_norm <= ITR_NORM_NONE, "unknown first norm type: %d", feature_norm);
  ITR_REQUIRE(agg >= 0 && agg <= ITR_AGG_SUM, "unknown aggfunc: %d", agg);
  ITR_REQUIRE(n_regions >= 1 && n_regions <= ITR_REGIONS, "itr_scan_backward_f32: 1 to %d regions per image, got %d", ITR_REGIONS, n_regions);
  ITR_REQUIRE(lmax >= 1 && lmax <= ITR_MAX_WORDS_F32, "itr_scan_backward_f32: padded caption width %d outside [1, %d]", lmax, ITR_MAX_WORDS_F32);
  ITR_REQUIRE(cross_attn == ITR_I2T || gram != nullptr, "itr_scan_backward_f32: t2i needs the region Gram");
  ITR_REQUIRE(d > 0 && d % 4 == 0 && ld_dscores >= n_cap, "itr_scan_backward_f32: bad shape (d must be a multiple of 4)");
  ITR_REQUIRE(n_words >= n_cap && n_words <= (int64_t)n_cap * lmax, "itr_scan_backward_f32: n_words inconsistent with n_cap / lmax");
  if (n_img <= 0 || n_cap <= 0) return ITR_OK;
  const int R = n_regions;
  const BwdWorkspace w = bwd_layout(n_img, R, n_cap, n_words, sum_len_sq, cross_attn);
  ITR_REQUIRE(workspace_bytes >= w.total * (int64_t)sizeof(float), "itr_scan_backward_f32: workspace too small (%lld < %lld bytes)",
              (long long)workspace_bytes, (long long)(w.total * sizeof(float)));
  ITR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(captions) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(images) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_images) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(d_captions) & 15) == 0, "itr_scan_backward_f32: pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  float* ws = static_cast<float*>(workspace);
  const int64_t rows = (int64_t)n_img * R;
  const bool t2i = cross_attn == ITR_T2I;
  int32_t* cap_off = reinterpret_cast<int32_t*>(ws + w.cap_off);
  int64_t* gram_off = reinterpret_cast<int64_t*>(ws + w.gram_off);
  int32_t* word_row = reinterpret_cast<int32_t*>(ws + w.word_row);
  caption_index_kernel<<<1, 256, 0, st>>>(cap_lens, n_cap, lmax, cap_off, gram_off, word_row);
  ITR_CHECK_LAUNCH();

  ScanBwdParams p;
  p.f = ScanF32Params{images, gram, captions, cap_lens, n_img, R, n_cap, lmax, d, cross_attn, feature_norm, agg,
                      lambda_softmax, lambda_lse, nullptr, 0};
  p.d_scores = d_scores; p.ld_ds = ld_dscores; p.cap_off = cap_off; p.gram_off = gram_off;
  p.M = ws + w.m; p.ldm = w.ldm; p.MT = ws + w.mt; p.ldmt = rows;
  p.Tpart = ws + w.tpart; p.MCpart = ws + w.mcpart; p.n_words = n_words; p.sum_n2 = sum_len_sq;
  const int RT = SF_IMGS * R, LP = sf_pitch(lmax);
  const int g_floats = t2i ? SF_IMGS * R * R : LP * LP;
  size_t head = (size_t)SF_VS_FLOATS + SF_WS_FLOATS;
  if ((size_t)RT * LP > head) head = (size_t)RT * LP;
  const size_t smem = sizeof(float) * (head + 2 * (size_t)RT * LP + g_floats + SF_LMAX + RT + 7 * (size_t)SF_IMGS * SB_RS);
  ITR_CHECK_CUDA(cudaFuncSetAttribute(scan_bwd_coeff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_cap, w.n_groups);
  ITR_REQUIRE(grid.y <= 65535, "itr_scan_backward_f32: more than %d images per call", 65535 * SF_IMGS);
  scan_bwd_coeff_kernel<<<grid, 256, smem, st>>>(p);
  ITR_CHECK_LAUNCH();

  // reduce the partial T / MC
  const int64_t t_n = t2i ? n_words : rows, mc_n = t2i ? rows * R : sum_len_sq;
  const int t_groups = t2i ? w.n_groups : n_cap, mc_groups = t2i ? n_cap : w.n_groups;
  leading_sum_kernel<<<(unsigned)((t_n + 255) / 256), 256, 0, st>>>(ws + w.tpart, t_n, t_groups, ws + w.tsum);
  ITR_CHECK_LAUNCH();
  leading_sum_kernel<<<(unsigned)((mc_n + 255) / 256), 256, 0, st>>>(ws + w.mcpart, mc_n, mc_groups, ws + w.mcsum);
  ITR_CHECK_LAUNCH();

  // dImages = M . Words (+ i2t: diag(T) . V)          overwritten
  SgemmNN gv{ws + w.m, w.ldm, captions, d, word_row, d_images, d, nullptr,
             t2i ? nullptr : ws + w.tsum, images, d, (int)rows, d, (int)n_words, 0};
  sgemm_nn_kernel<<<dim3((d + 63) / 64, (unsigned)((rows + 63) / 64)), 256, 0, st>>>(gv);
  ITR_CHECK_LAUNCH();
  // dCaptions += M^T . Regions (+ t2i: diag(T) . W)    accumulated into the padded rows word_row[p]
  SgemmNN gw{ws + w.mt, rows, images, d, nullptr, d_captions, d, word_row,
             t2i ? ws + w.tsum : nullptr, captions, d, (int)n_words, d, (int)rows, 1};
  sgemm_nn_kernel<<<dim3((d + 63) / 64, (unsigned)((n_words + 63) / 64)), 256, 0, st>>>(gw);
  ITR_CHECK_LAUNCH();

  // block-diagonal context terms
  BlockDiag bd;
  if (t2i) bd = BlockDiag{ws + w.mcsum, images, d_images, d, d, nullptr, nullptr, R, R};
  else bd = BlockDiag{ws + w.mcsum, captions, d_captions, d, d, cap_lens, gram_off, 0, lmax};
  const int nb_max = t2i ? R : lmax;
  const size_t bsm = sizeof(float) * ((size_t)nb_max * nb_max + (size_t)nb_max * 128);
  ITR_CHECK_CUDA(cudaFuncSetAttribute(blockdiag_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
  blockdiag_apply_kernel<<<dim3(t2i ? n_img : n_cap, (d + 127) / 128), 128, bsm, st>>>(bd);
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}
