// Microbenchmark: tcgen05.ld throughput of 8 epilogue warps WHILE the tensor core accumulates into other TMEM columns.
// One CTA per SM: warp 8 issues back-to-back M128 x N256 x K16 bf16 MMAs (operands: zero-filled shared memory, K-major
// no-swizzle core matrices) into columns [256, 512); warps 0-7 read columns [0, 256) with 32x32b.x32 loads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/micro/tmem_ld_mma_bench.bin scripts/micro/tmem_ld_mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo & 0x3FFFF) >> 4) << 16) | ((uint64_t)((sbo & 0x3FFFF) >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mma_mode 0: no MMAs; 1: MMAs N=256 back to back; 2: MMAs N=128
__global__ void __launch_bounds__(288, 1) kern(int iters, int mma_mode, int ld_warps, long long* cycles, long long* mma_cycles, uint32_t* sink) {
  extern __shared__ __align__(128) uint8_t smem[];     // A: 128 x 16 bf16 = 4 KB, B: 256 x 16 bf16 = 8 KB (zeros)
  __shared__ uint32_t holder;
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 12288 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { stop = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(&holder)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = holder;
  if (warp == 8) {
    if (lane == 0 && mma_mode) {
      const int N = mma_mode == 1 ? 256 : 128;
      const uint32_t idesc = make_idesc(128, N);
      const uint64_t a = make_desc(s32(smem), 2048, 128), b = make_desc(s32(smem) + 4096, (uint32_t)N * 16, 128);
      const long long t0 = clock64();
      long long n = 0;
      uint32_t phase = 0;
      while (!stop) {
        for (int i = 0; i < 16; ++i)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       :: "r"(tbase + 256u), "l"(a), "l"(b), "r"(idesc), "r"(1u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(s32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(s32(&bar)), "r"(phase) : "memory");
        phase ^= 1u;
        n += 16;
      }
      mma_cycles[blockIdx.x * 2] = clock64() - t0;
      mma_cycles[blockIdx.x * 2 + 1] = n;
    }
  } else if (warp < ld_warps) {
    const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    uint32_t a[32], b[32];
    for (int i = 0; i < iters; ++i) {
      const uint32_t col = (uint32_t)(((i * 2) * 64 + (warp >> 2) * 32) & 255);
      ld32(base + col, a); ld32(base + ((col + 64) & 255), b);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= a[j] + b[j];
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    asm volatile("bar.sync 1, %0;" :: "r"(ld_warps * 32) : "memory");
    if (threadIdx.x == 0) stop = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512) : "memory");
}

int main() {
  long long *cyc, *mc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&mc, 148 * 16); cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  const int iters = 20000;
  for (int ld_warps : {4, 8})
    for (int mode = 0; mode < 3; ++mode) {
      cudaMemset(mc, 0, 148 * 16);
      for (int rep = 0; rep < 2; ++rep) {
        kern<<<148, 288, 16384>>>(iters, mode, ld_warps, cyc, mc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[148], m[296]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(m, mc, sizeof(m), cudaMemcpyDeviceToHost);
      printf("%d reading warps, MMA %s: tcgen05.ld %6.1f B/clk/SM (%5.1f clk per 4 KB load per warp)", ld_warps,
             mode == 0 ? "off     " : (mode == 1 ? "N=256 on" : "N=128 on"), (double)iters * 2 * 4096.0 * ld_warps / h[0], (double)h[0] / (iters * 2));
      if (mode) printf(";  MMA %6.1f clk each", (double)m[0] / (double)m[1]);
      printf("\n");
    }
  return 0;
}
