// Microbenchmark: tcgen05.ld throughput per SM on B200 as a function of the number of reading warps, the load width
// and the number of loads in flight.  One CTA per SM, 512 TMEM columns.  Prints bytes/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tmem_ld_bench scripts/micro/tmem_ld_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld16x256(uint32_t taddr, uint32_t (&v)[32]) {   // 16 lanes x 256 bit, x8 -> 32 regs
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: one x32 load, wait, consume.  MODE 1: two x32 loads in flight.  MODE 2: four in flight.  MODE 3: 16x256b.x8
template <int MODE>
__global__ void __launch_bounds__(512, 1) kern(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&holder)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = holder + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  uint32_t a[32], b[32], c[32], d[32];
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)((i * 128 + (warp >> 2) * 32) & 511);
    if (MODE == 0) {
      ld32(base + col, a); wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= a[j];
    } else if (MODE == 1) {
      ld32(base + col, a); ld32(base + ((col + 32) & 511), b); wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= a[j] + b[j];
    } else if (MODE == 2) {
      ld32(base + col, a); ld32(base + ((col + 32) & 511), b); ld32(base + ((col + 64) & 511), c); ld32(base + ((col + 96) & 511), d); wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= a[j] + b[j] + c[j] + d[j];
    } else {
      ld16x256(base + col, a); ld16x256(base + ((col + 64) & 511), b); wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= a[j] + b[j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(holder), "r"(512) : "memory");
}

int main() {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4096);
  const int iters = 20000;
  const int loads_per_iter[4] = {1, 2, 4, 2};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {1, 2, 4, 8, 12, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) kern<0><<<148, warps * 32>>>(iters, cyc, sink);
        if (mode == 1) kern<1><<<148, warps * 32>>>(iters, cyc, sink);
        if (mode == 2) kern<2><<<148, warps * 32>>>(iters, cyc, sink);
        if (mode == 3) kern<3><<<148, warps * 32>>>(iters, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double bytes = (double)iters * loads_per_iter[mode] * 4096.0 * warps;
      printf("mode %d warps %2d: %9lld cycles  -> %7.1f B/clk/SM, %6.1f clk per 4KB load per warp\n", mode, warps, h[0], bytes / h[0],
             (double)h[0] / (iters * loads_per_iter[mode]));
    }
  return 0;
}
