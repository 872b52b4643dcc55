// Pieces shared by the CTA-pair tcgen05 kernels (scan_t2i_tc2.cu, scan_i2t_tc2.cu): tile geometry, cluster / pair PTX
// wrappers, explicit shared-space accesses, the tile schedule and the TMA descriptor helper.  sm_100a only.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace itr {
namespace tc2 {

using namespace itr::tc;

constexpr int R = ITR_REGIONS;                 // 36
constexpr int IMGS = ITR_TILE_IMAGES;          // 4
constexpr int BLOCK_M = ITR_TILE_WORDS;        // 128 word rows per CTA (UMMA M = 256 across the pair)
constexpr int BLOCK_N = IMGS * R;              // 144
constexpr int HALF_N = BLOCK_N / 2;            // 72 region rows of B staged by each CTA
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int D = ITR_EMBED;                   // 1024
constexpr int K_BLOCKS = D / BLOCK_K;          // 16
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2; // 16384
constexpr int B_BYTES = HALF_N * BLOCK_K * 2;  // 9216
constexpr int STAGE_BYTES = A_BYTES + B_BYTES; // 25600 (1024-aligned: SWIZZLE_128B tiles)
constexpr int ACC_PITCH = BLOCK_N;             // two accumulators [0,144) and [144,288)
constexpr int TMEM_COLS = 512;
constexpr int BAND = 32;                       // word-tile PAIRS (64 tiles, 16 MB) kept L2-resident while images stream
constexpr int NUM_THREADS = 640;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 16;
// kind::f16 instruction descriptor, M = 256 across the CTA pair: D=f32, A=B=bf16, K-major, N = 144
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

// ---------------------------------------------------------------------------- cluster / pair PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// In a cluster of two, the shared-memory windows of the CTAs differ in one address bit (bit 24 = rank): clearing it
// turns the address of an object of this CTA into the shared::cluster address of the same object in the leader.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
// arrive on the LEADER's copy of the barrier at `bar` (an address in this CTA's layout).  The leader takes the plain
// shared::cta form; the peer the shared::cluster form WITHOUT .release.cluster: measured (profiles/r02/
// role_profile_pair_v1_slow.txt), a cluster-scope release costs the arriving warp ~850 clk per arrive, and nothing needs
// it here -- the tcgen05.ld / tcgen05.st the arrival publishes have completed (wait::ld / wait::st) before it is issued.
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar, bool leader) {
  if (leader) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
  else asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_BIT_MASK) : "memory");
}
// TMA tile load of one CTA of a pair: the bytes complete on the LEADER's barrier (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once every tcgen05 operation issued so far has completed
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// ... on the issuing (leader) CTA's barrier only
__device__ __forceinline__ void umma2_commit_local(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// explicit shared-space accesses (a pointer derived from the manually aligned dynamic-SMEM base is a GENERIC pointer to
// the compiler: it would emit LD.E / ST.E with address translation instead of LDS / STS)
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f2(uint32_t addr, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}

// ---------------------------------------------------------------------------- tile schedule
// Work unit = (band of BAND consecutive word-tile PAIRS, image tile n); a CTA pair takes units u = pair, pair + #pairs, ...
// and walks the band against the SAME image tile (see scan_t2i_tc.cu: an image tile comes from HBM once per band,
// the band stays L2-resident).
// With an explicit item list (the ground-truth pre-pass of the fused evaluation) a unit is one listed item
// (word tile of the leader, word tile of the peer, image tile): the two CTAs take ANY two word tiles that need that
// image tile (a tile index >= n_wt means "none").
template <bool LIST>
struct ScheduleT {
  int n_wp, n_it, n_bands, last_band, band;
  const int4* items; int n_items;
  __device__ ScheduleT(int n_wp_, int n_it_, const int4* items_, int n_items_, int band_ = BAND)
      : n_wp(n_wp_), n_it(n_it_), band(band_), items(items_), n_items(n_items_) {
    n_bands = (n_wp + band - 1) / band;
    last_band = n_wp - (n_bands - 1) * band;
  }
  __device__ int units() const { return LIST ? n_items : n_bands * n_it; }
};
template <bool LIST>
struct ItemIterT {
  const ScheduleT<LIST>& s;
  int u, step, m, n, left;      // m = word-tile PAIR index (LIST: the leader's word tile)
  int m_peer;                   // LIST: the peer's word tile
  // word tile of CTA `rank` of the pair
  __device__ int tile(int rank) const { return LIST ? (rank ? m_peer : m) : 2 * m + rank; }
  __device__ ItemIterT(const ScheduleT<LIST>& s_, int first, int step_) : s(s_), u(first), step(step_) { open(); }
  __device__ void open() {
    if (LIST) {
      if (u < s.n_items) { const int4 e = s.items[u]; m = e.x; m_peer = e.y; n = e.z; left = 1; }
      return;
    }
    if (u < s.units()) {
      const int band = u / s.n_it;
      n = u - band * s.n_it;
      m = band * s.band;
      left = (band == s.n_bands - 1) ? s.last_band : s.band;
    }
  }
  __device__ bool valid() const { return u < s.units(); }
  __device__ void next() {
    ++m;
    if (--left == 0) { u += step; open(); }
  }
};


// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int get_encode_fn(EncodeTiledFn* fn) {
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
inline int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint32_t box_rows) {
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


}  // namespace tc2
}  // namespace itr
