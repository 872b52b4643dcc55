// VSE++ training step of the path in ONE cooperative launch (sm_100a):
//   scores = im @ s.T (Objectives.py:18-21), the hinge of ContrastiveLoss.forward (Objectives.py:93-115: max_violation or
//   sum), and the gradients w.r.t. both embedding matrices, d_im = dS @ s, d_s = dS.T @ im.
// At batch 128 x embed 1024 the step is 100 MFLOP over 1 MB: launch- and dependency-latency bound, not arithmetic bound.
// Five back-to-back launches plus a stream-ordered allocation (round 1) cost ~150 us; here the three dependent phases
// run inside one grid with two grid-wide barriers:
//   phase 1  split-K partial products of the n x n scores, (32 x 32 output tile, K chunk) per CTA -- every SM has work
//   phase 2  CTA i reduces row i and column i of the partials in a fixed order (bit-reproducible), writes the row of
//            S and the statistics the hinge needs: hardest negative + index (max_violation) or sum + count
//   phase 3  loss (block 0, fixed order) and both gradient products with dS rebuilt on the fly from S and the stats
// Exact float32 throughout (the 1e-5 contract of the fp32 mode).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace itr {
namespace vse {

constexpr int THREADS = 256;
constexpr int TILE = 32;          // phase-1 output tile
constexpr int KC = 128;           // phase-1 K chunk (floats)
constexpr int GT_ROWS = 16;       // phase-3 output tile: 16 rows x 128 columns
constexpr int GT_COLS = 128;

struct Params {
  const float* im; const float* s;
  int n, d, k_chunks, k_len;      // K is cut into k_chunks ranges of k_len floats (a multiple of KC)
  float margin; int max_violation;
  float* part;                    // [k_chunks][n][n]
  float* S;                       // [n][n]
  float4* stats;                  // [n]
  float* loss; float* d_im; float* d_s;
};

__device__ __forceinline__ float ds_entry(const Params& p, int i, int j, const float4& si, const float4& sj, float s_ij,
                                          float d_i, float d_j) {
  if (i == j)
    return p.max_violation ? -(float)((si.x > 0.f) + (si.z > 0.f)) : -(float)(__float_as_int(si.y) + __float_as_int(si.w));
  if (p.max_violation)
    return (float)((__float_as_int(si.y) == j && si.x > 0.f) + (__float_as_int(sj.w) == i && sj.z > 0.f));
  return (float)((p.margin + s_ij - d_i > 0.f) + (p.margin + s_ij - d_j > 0.f));
}

__global__ void __launch_bounds__(THREADS)
vse_step_kernel(Params p) {
  cg::grid_group grid = cg::this_grid();
  __shared__ __align__(16) float sa[TILE][KC + 4];
  __shared__ __align__(16) float sb[TILE][KC + 4];
  const int tid = threadIdx.x, n = p.n, d = p.d;

  // ---------------------------------------------------------------- phase 1: split-K partial scores
  {
    const int tiles = (n + TILE - 1) / TILE;
    const int items = tiles * tiles * p.k_chunks;
    const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 2 x 2 outputs each
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int kc = it % p.k_chunks, t = it / p.k_chunks;
      const int i0 = (t / tiles) * TILE, j0 = (t % tiles) * TILE;
      float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
      for (int k0 = kc * p.k_len; k0 < min(d, (kc + 1) * p.k_len); k0 += KC) {
        const int klen = min(KC, d - k0);
        __syncthreads();
        for (int e = tid; e < TILE * (KC / 4); e += THREADS) {
          const int r = e / (KC / 4), c4 = (e % (KC / 4)) * 4;
          float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
          if (c4 < klen) {           // d is a multiple of 4 (checked on the host), so a float4 is all in or all out
            if (i0 + r < n) va = *reinterpret_cast<const float4*>(p.im + (size_t)(i0 + r) * d + k0 + c4);
            if (j0 + r < n) vb = *reinterpret_cast<const float4*>(p.s + (size_t)(j0 + r) * d + k0 + c4);
          }
          *reinterpret_cast<float4*>(&sa[r][c4]) = va;
          *reinterpret_cast<float4*>(&sb[r][c4]) = vb;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < KC; k += 4) {
          const float4 a0 = *reinterpret_cast<const float4*>(&sa[ty][k]), a1 = *reinterpret_cast<const float4*>(&sa[ty + 16][k]);
          const float4 b0 = *reinterpret_cast<const float4*>(&sb[tx][k]), b1 = *reinterpret_cast<const float4*>(&sb[tx + 16][k]);
          acc[0][0] = fmaf(a0.x, b0.x, fmaf(a0.y, b0.y, fmaf(a0.z, b0.z, fmaf(a0.w, b0.w, acc[0][0]))));
          acc[0][1] = fmaf(a0.x, b1.x, fmaf(a0.y, b1.y, fmaf(a0.z, b1.z, fmaf(a0.w, b1.w, acc[0][1]))));
          acc[1][0] = fmaf(a1.x, b0.x, fmaf(a1.y, b0.y, fmaf(a1.z, b0.z, fmaf(a1.w, b0.w, acc[1][0]))));
          acc[1][1] = fmaf(a1.x, b1.x, fmaf(a1.y, b1.y, fmaf(a1.z, b1.z, fmaf(a1.w, b1.w, acc[1][1]))));
        }
      }
      float* out = p.part + (size_t)kc * n * n;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
          if (i < n && j < n) out[(size_t)i * n + j] = acc[a][b];
        }
    }
  }
  grid.sync();

  // ---------------------------------------------------------------- phase 2: reduce, row of S, hinge statistics
  {
    __shared__ float sv[2][8];
    __shared__ int si[2][8];
    __shared__ float sdiag;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
      __syncthreads();
      if (tid == 0) {
        float v = 0.f;
        for (int kc = 0; kc < p.k_chunks; ++kc) v += p.part[(size_t)kc * n * n + (size_t)i * n + i];
        sdiag = v;
      }
      __syncthreads();
      const float dg = sdiag;
      float rv = p.max_violation ? -1.f : 0.f, cv = rv;
      int ra = 0x7fffffff, ca = 0x7fffffff, rc = 0, cc = 0;
      for (int j = tid; j < n; j += THREADS) {
        float row = 0.f, col = 0.f;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          const float* pp = p.part + (size_t)kc * n * n;
          row += pp[(size_t)i * n + j];
          col += pp[(size_t)j * n + i];
        }
        p.S[(size_t)i * n + j] = row;
        if (j == i) continue;
        const float a = fmaxf(p.margin + row - dg, 0.f), b = fmaxf(p.margin + col - dg, 0.f);
        if (p.max_violation) {
          if (a > rv) { rv = a; ra = j; }
          if (b > cv) { cv = b; ca = j; }
        } else {
          rv += a; cv += b; rc += (a > 0.f); cc += (b > 0.f);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float orv = __shfl_xor_sync(0xffffffffu, rv, o), ocv = __shfl_xor_sync(0xffffffffu, cv, o);
        const int ora = __shfl_xor_sync(0xffffffffu, ra, o), oca = __shfl_xor_sync(0xffffffffu, ca, o);
        const int orc = __shfl_xor_sync(0xffffffffu, rc, o), occ = __shfl_xor_sync(0xffffffffu, cc, o);
        if (p.max_violation) {
          if (orv > rv || (orv == rv && ora < ra)) { rv = orv; ra = ora; }
          if (ocv > cv || (ocv == cv && oca < ca)) { cv = ocv; ca = oca; }
        } else {
          rv += orv; cv += ocv; rc += orc; cc += occ;
        }
      }
      if (lane == 0) { sv[0][warp] = rv; sv[1][warp] = cv; si[0][warp] = p.max_violation ? ra : rc; si[1][warp] = p.max_violation ? ca : cc; }
      __syncthreads();
      if (tid == 0) {
        float R = sv[0][0], C = sv[1][0];
        int RI = si[0][0], CI = si[1][0];
        for (int w = 1; w < THREADS / 32; ++w) {
          if (p.max_violation) {
            if (sv[0][w] > R || (sv[0][w] == R && si[0][w] < RI)) { R = sv[0][w]; RI = si[0][w]; }
            if (sv[1][w] > C || (sv[1][w] == C && si[1][w] < CI)) { C = sv[1][w]; CI = si[1][w]; }
          } else {
            R += sv[0][w]; C += sv[1][w]; RI += si[0][w]; CI += si[1][w];
          }
        }
        if (p.max_violation) { R = fmaxf(R, 0.f); C = fmaxf(C, 0.f); }
        p.stats[i] = make_float4(R, __int_as_float(RI), C, __int_as_float(CI));
      }
    }
  }
  grid.sync();

  // ---------------------------------------------------------------- phase 3: loss and gradients
  if (blockIdx.x == 0) {
    __shared__ float part[THREADS / 32];
    float v = 0.f;
    for (int i = tid; i < n; i += THREADS) v += p.stats[i].x + p.stats[i].z;
    v = warp_sum(v);
    if ((tid & 31) == 0) part[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < THREADS / 32; ++w) t += part[w];
      *p.loss = t;
    }
  }
  if (p.d_im == nullptr && p.d_s == nullptr) return;
  {
    // d_im[i][k] = sum_j dS[i][j] s[j][k]   and   d_s[j][k] = sum_i dS[i][j] im[i][k]:
    // tile = GT_ROWS output rows x GT_COLS columns; the dS block (GT_ROWS x n) is rebuilt in shared memory per tile
    float* dsb = &sa[0][0];                           // GT_ROWS x n floats (n <= 264 fits: 32*132 floats available)
    const int row_tiles = (n + GT_ROWS - 1) / GT_ROWS, col_tiles = (d + GT_COLS - 1) / GT_COLS;
    const int per = row_tiles * col_tiles;
    const int which0 = p.d_im ? 0 : 1, n_which = (p.d_im ? 1 : 0) + (p.d_s ? 1 : 0);
    const int items = per * n_which;
    const int c = tid & 31, rgrp = tid >> 5;          // thread: columns c, c+32, c+64, c+96 of rows rgrp, rgrp+8
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int which = which0 + it / per;            // 0: d_im, 1: d_s
      const int t = it % per;
      const int r0 = (t / col_tiles) * GT_ROWS, k0 = (t % col_tiles) * GT_COLS;
      __syncthreads();
      for (int e = tid; e < GT_ROWS * n; e += THREADS) {
        const int r = e / n, j = e % n;
        const int row = r0 + r;
        float g = 0.f;
        if (row < n) {
          // d_im: output row = image i = row, contraction over captions j;  d_s: output row = caption j' = row,
          // contraction over images i = j (the roles swap: dS[i][j'] with i running)
          const int i = which == 0 ? row : j, jj = which == 0 ? j : row;
          const float4 si_ = p.stats[i], sj_ = p.stats[jj];
          g = ds_entry(p, i, jj, si_, sj_, p.S[(size_t)i * n + jj], p.S[(size_t)i * n + i], p.S[(size_t)jj * n + jj]);
        }
        dsb[r * n + j] = g;
      }
      __syncthreads();
      const float* src = which == 0 ? p.s : p.im;
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      for (int j = 0; j < n; ++j) {
        const float g0 = dsb[rgrp * n + j], g1 = dsb[(rgrp + 8) * n + j];
        if (g0 == 0.f && g1 == 0.f) continue;        // max_violation: two non-zeros per row
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = k0 + c + 32 * q;
          const float v = k < d ? src[(size_t)j * d + k] : 0.f;
          acc[0][q] = fmaf(g0, v, acc[0][q]);
          acc[1][q] = fmaf(g1, v, acc[1][q]);
        }
      }
      float* dst = which == 0 ? p.d_im : p.d_s;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int row = r0 + rgrp + 8 * a;
        if (row >= n) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = k0 + c + 32 * q;
          if (k < d) dst[(size_t)row * d + k] = acc[a][q];
        }
      }
    }
  }
}

}  // namespace vse
}  // namespace itr

using namespace itr;

// largest batch the one-launch step handles (the dS block of a phase-3 tile must fit the static shared memory)
static constexpr int VSE_FUSED_MAX_N = 264;

// K ranges per output tile: as many as it takes to give every SM an item, each a multiple of KC floats long
static void vse_k_split(int n, int d, int sms, int* k_chunks, int* k_len) {
  const int tiles = ((n + vse::TILE - 1) / vse::TILE) * ((n + vse::TILE - 1) / vse::TILE);
  const int blocks = (d + vse::KC - 1) / vse::KC;            // KC-sized blocks of K
  int want = (sms + tiles - 1) / tiles;                      // ranges per tile
  if (want > blocks) want = blocks;
  if (want < 1) want = 1;
  const int per = (blocks + want - 1) / want;                // blocks per range
  *k_len = per * vse::KC;
  *k_chunks = (blocks + per - 1) / per;
}

extern "C" int64_t itr_cosine_hinge_workspace_f32(int n, int d) {
  if (n < 1 || d < 1) return -1;
  const int64_t nn = (int64_t)n * n;
  if (n > VSE_FUSED_MAX_N || d % 4 != 0) return 2 * nn;                   // multi-launch path (simt_kernels.cu)
  return ((d + vse::KC - 1) / vse::KC + 1) * nn + 4 * (int64_t)n + 8;      // partials (worst case) + S + stats (+ alignment)
}

// 1 = handled, 0 = not applicable (caller takes the multi-launch path), < 0 = error
int vse_step_fused(const float* im, const float* s, int n, int d, float margin, int max_violation, float* ws, float* loss,
                   float* d_im, float* d_s, cudaStream_t st, int* status) {
  *status = ITR_OK;
  if (n > VSE_FUSED_MAX_N || d % 4 != 0 || ((uintptr_t)im & 15) || ((uintptr_t)s & 15) || ((uintptr_t)ws & 15)) return 0;
  static int sms = 0, coop = -1, per_sm = 0;
  if (coop < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vse::vse_step_kernel, vse::THREADS, 0) != cudaSuccess) per_sm = 0;
  }
  if (!coop || per_sm < 1 || sms < 1) return 0;
  vse::Params p;
  p.im = im; p.s = s; p.n = n; p.d = d;
  vse_k_split(n, d, sms, &p.k_chunks, &p.k_len);
  p.margin = margin; p.max_violation = max_violation;
  const int64_t nn = (int64_t)n * n;
  const int64_t off_s = (int64_t)p.k_chunks * nn, off_stats = (off_s + nn + 3) & ~(int64_t)3;      // stats: 16-byte aligned
  p.part = ws; p.S = ws + off_s;
  p.stats = reinterpret_cast<float4*>(ws + off_stats);
  p.loss = loss; p.d_im = d_im; p.d_s = d_s;
  void* args[] = {&p};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(vse::vse_step_kernel), dim3(sms), dim3(vse::THREADS), args, 0, st);
  if (e != cudaSuccess) { *status = fail(ITR_ERR_CUDA, "vse_step_kernel launch failed: %s", cudaGetErrorString(e)); return -1; }
  return 1;
}
