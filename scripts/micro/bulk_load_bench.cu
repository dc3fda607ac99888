// Microbenchmark: cp.async.bulk global -> shared throughput per SM from an L2-resident buffer, as a function of the
// copy size, the number of copies in flight and whether all SMs stream the SAME bytes (weights) or private ones.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/micro/bulk_load_bench.bin scripts/micro/bulk_load_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// `issuers` warps each run their own ring of `depth` stages (lane 0 issues): are copies of different warps concurrent?
__global__ void __launch_bounds__(128, 1) kern(const uint8_t* buf, size_t region, int shared_region, uint32_t unit, int depth,
                                               int n_units, long long* cycles, int issuers) {
  extern __shared__ __align__(128) uint8_t smem_all[];
  __shared__ uint64_t bars_all[4][16];
  const int w = threadIdx.x >> 5;
  uint8_t* smem = smem_all + (size_t)w * unit * depth;
  uint64_t* bars = bars_all[w];
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* base = buf + (shared_region ? 0 : (size_t)blockIdx.x * region) + (size_t)w * (region / 4);
  region /= 4;
  if ((threadIdx.x & 31) == 0 && w < issuers) {
    const long long t0 = clock64();
    const int per_region = (int)(region / unit);
    for (int i = 0; i < n_units + depth; ++i) {
      const int st = i % depth;
      if (i >= depth) {                       // wait for the copy that used this stage
        const uint32_t parity = (uint32_t)(((i / depth) - 1) & 1);
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(s32(&bars[st])), "r"(parity) : "memory");
      }
      if (i < n_units) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bars[st])), "r"(unit) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(s32(smem + (size_t)st * unit)), "l"(base + (size_t)(i % per_region) * unit), "r"(unit), "r"(s32(&bars[st])) : "memory");
      }
    }
    if (w == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const size_t region = 512 << 10;            // 512 KB per SM (or shared by all): 74 MB in all, L2 resident
  uint8_t* buf; long long* cyc;
  cudaMalloc(&buf, region * 148); cudaMemset(buf, 1, region * 148); cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10);
  const int total_kb = 16 << 10;              // 16 MB per SM
  for (int issuers : {1, 2, 4})
  for (int shared_region : {1})
    for (uint32_t unit : {4096u, 16384u, 32768u, 65536u})
      for (int depth : {1, 2, 4}) {
        if ((size_t)unit * depth * issuers > (192u << 10)) continue;
        const int n_units = (int)((size_t)total_kb * 1024 / unit);
        float ms = 0; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          kern<<<148, 128, (size_t)unit * depth * issuers>>>(buf, region, shared_region, unit, depth, n_units, cyc, issuers);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaEventElapsedTime(&ms, e0, e1);
        }
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%d issuing warps, unit %2u KB x depth %d: %6.1f B/clk/SM (all warps), %6.0f clk per copy per warp\n",
               issuers, unit >> 10, depth, (double)total_kb * 1024 * issuers / h[0], (double)h[0] / n_units);
      }
  return 0;
}
