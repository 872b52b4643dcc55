// tcgen05.mma cost model, CTA-pair form (cta_group::2, M = 256 across two SMs of a TPC).
// One cluster of two CTAs per SM pair; the leader CTA issues `iters` kind::f16 MMAs (K = 16) on whatever shared
// memory holds and reports cycles per MMA.  Companion of mma_microbench_kernel (scan_t2i_tc.cu); the numbers
// decide the tile shape of the 2-CTA score kernel.
//   n_cols : UMMA N (multiple of 16, <= 256); each CTA supplies N/2 rows of the B operand
//   n_acc  : accumulators cycled through
//   a_tmem : 1 = A operand from tensor memory (TS), 0 = from shared memory (SS)
//   n_issuers : warps of the leader CTA issuing concurrently (each on its own accumulators)
#include "tc_ptx.cuh"

namespace itr {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
mma2_microbench_kernel(int n_cols, int n_acc, int iters, int a_tmem, int n_issuers, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ long long t_issuer[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tbase = tmem_ptr;
  if (warp >= 1 && warp <= n_issuers) {
    const int w = warp - 1;
    long long t0 = 0;
    if (rank == 0) {
      // M = 256 across the pair: idesc M field = 256 >> 4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint64_t adesc = umma_desc_sw128(sbase), bdesc = umma_desc_sw128(sbase + 16384);
      t0 = clock64();
      int acc = 0;
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tbase + (uint32_t)((w * n_acc + acc) * n_cols);
        const int k = i & 3;
        if (elect_one()) {
          if (a_tmem) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(d), "r"(tbase + 480), "l"(bdesc + 2 * k), "r"(idesc), "r"(1u) : "memory");
          } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(1u) : "memory");
          }
        }
        __syncwarp();
        if (++acc == n_acc) acc = 0;
      }
      if (elect_one()) {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bars[w])), "h"((uint16_t)3) : "memory");
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&bars[w]), 0);      // both CTAs: the commit is multicast to the same barrier offset in each
    const long long t1 = clock64();
    if (lane == 0) t_issuer[w] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && rank == 0) {
    long long m = 0;
    for (int i = 0; i < n_issuers; ++i) m = t_issuer[i] > m ? t_issuer[i] : m;
    out[blockIdx.x >> 1] = m;
  }
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
  }
}

}  // namespace tc
}  // namespace itr

using namespace itr;
using namespace itr::tc;

extern "C" int itr_tc_mma2_microbench(int n_cols, int n_acc, int iters, int a_tmem, int n_issuers, int n_pairs, int64_t* cycles, void* stream) {
  ITR_REQUIRE(cycles && n_cols >= 32 && n_cols <= 256 && n_cols % 16 == 0 && n_acc >= 1 && n_issuers >= 1 && n_issuers <= 4 &&
              n_acc * n_issuers * n_cols <= 480 && iters > 0 && n_pairs > 0, "itr_tc_mma2_microbench: bad arguments");
  int dev = 0, major = 0;
  ITR_CHECK_CUDA(cudaGetDevice(&dev));
  ITR_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(ITR_ERR_UNSUPPORTED, "itr_tc_mma2_microbench needs an sm_100 device");
  const int smem = 64 * 1024 + 1024;
  ITR_CHECK_CUDA(cudaFuncSetAttribute(mma2_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma2_microbench_kernel<<<2 * n_pairs, 192, smem, as_stream(stream)>>>(n_cols, n_acc, iters, a_tmem, n_issuers, reinterpret_cast<long long*>(cycles));
  ITR_CHECK_LAUNCH();
  return ITR_OK;
}
