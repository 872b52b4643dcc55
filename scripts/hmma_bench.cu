// Throughput of the legacy warp-level MMA shapes on sm_100a: instructions per clock per SM with 4 / 8 / 16 warps per SM,
// 4 independent accumulator chains per warp.  nvcc -arch=sm_100a -O3 -o /tmp/hb scripts/hmma_bench.cu && /tmp/hb
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(float* out, long long* clk, int iters) {
  float d[4][4] = {};
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 2)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(b0));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int c = 0; c < 4; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int KIND> void run(const char* name) {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  for (int warps : {4, 8, 16}) {
    const int iters = 2000;
    k<KIND><<<148, warps * 32>>>(out, clk, iters); cudaDeviceSynchronize();
    k<KIND><<<148, warps * 32>>>(out, clk, iters); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("%-22s warps/SM %2d: %.2f clk per MMA per SM sub-partition  (%.3f MMA/clk/SM)\n", name, warps,
           c / (iters * 4.0 * warps / 4.0), iters * 4.0 * warps / c);
  }
}
int main() {
  run<0>("m16n8k8 tf32"); run<3>("m16n8k4 tf32"); run<1>("m16n8k16 f16"); run<2>("m16n8k16 bf16");
  cudaError_t e = cudaGetLastError(); printf("%s\n", cudaGetErrorString(e));
  return 0;
}
