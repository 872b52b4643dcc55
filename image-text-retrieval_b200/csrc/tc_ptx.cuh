// PTX wrappers shared by the tcgen05 kernels (sm_100a only): mbarriers, TMA, tcgen05.mma / ld / st / commit,
// shared-memory operand descriptors.  Included inside namespace-less translation units; everything lives in itr::tc.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace itr {
namespace tc {

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One elected lane of a converged warp.  Unlike `lane == 0`, the compiler knows the branch is entered by
// exactly one thread with warp-uniform operands, so uniform-datapath instructions (UTCHMMA, UTMALDG, UTCBAR)
// are issued directly instead of through an ELECT/BRA.U.ANY uniformizing loop (measured: 187 -> N/2 clk per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait blocks in hardware for a short, implementation-defined time before it returns false.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// same with a suspend-time hint: the thread sleeps until the phase completes or ~hint_ns elapse.  Cheap in
// issue slots, but the wake-up is slow (measured ~1-2K clk round trips on the operand ring), so only the
// epilogue warps use it; the control warps spin.
__device__ __forceinline__ bool mbar_try_wait_sleep(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(hint_ns) : "memory");
  return ok != 0;
}
// Bounded waits: a protocol bug becomes a trap (reported as a CUDA error) instead of a hang.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {        // latency-critical: spin
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  int n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++n & 1023) == 0 && clock64() - t0 > 4000000000ll) __trap();
  }
}
// The watchdog of the sleeping waits reads the clock once every 1024 failed polls, inline: an out-of-line call here made
// ptxas spill 27 registers of the counting kernel around the call site (COCO-5K step 134.1 -> 132.8 ms without it).
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {  // throughput warps: sleep
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  int n = 0;
  while (!mbar_try_wait_sleep(bar, parity, 20000u)) {
    if ((++n & 1023) == 0 && clock64() - t0 > 6000000000ll) __trap();
  }
}
// wait and add the cycles spent waiting to `acc` (profiling builds of the role loops)
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, long long& acc, const bool on) {
  if (!on) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}
__device__ __forceinline__ void mbar_wait_sleep_t(uint32_t bar, uint32_t parity, long long& acc, const bool on) {
  if (!on) { mbar_wait_sleep(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait_sleep(bar, parity);
  acc += clock64() - t0;
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
// A operand read from tensor memory (lanes = M rows, two fp16 K elements per 32-bit column)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, SWIZZLE_128B operand tile whose rows are 128 bytes: 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// K-major, no swizzle: 8x8 core matrices, LBO = stride between core matrices along K, SBO = along N
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
#define TMEM_ST_X16(taddr, v, o)                                                                                     \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               ::"r"(taddr), "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]),  \
                 "r"(v[o + 6]), "r"(v[o + 7]), "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]),            \
                 "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), "r"(v[o + 15]) : "memory")
#define TMEM_ST_X8(taddr, v, o)                                                                     \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"             \
               ::"r"(taddr), "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), \
                 "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7]) : "memory")
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
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
#define TMEM_LD_X16U(taddr, v, o)                                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),      \
                 "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),    \
                 "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])                                   \
               : "r"(taddr))
#define TMEM_LD_X8U(taddr, v, o)                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"             \
               : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]),    \
                 "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])                                     \
               : "r"(taddr))
#define TMEM_LD_X2U(taddr, v, o) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[o + 0]), "=r"(v[o + 1]) : "r"(taddr))
#define TMEM_LD_X1F(taddr, f) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(f) : "r"(taddr))
#define TMEM_ST_X4(taddr, v, o) \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]) : "memory")
#define TMEM_ST_X2(taddr, v, o) \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(v[o + 0]), "r"(v[o + 1]) : "memory")
#define TMEM_ST_X1(taddr, r0) \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r0) : "memory")
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqf(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

}  // namespace tc
}  // namespace itr
