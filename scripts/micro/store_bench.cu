// Microbenchmark: global-store throughput per SM and for the whole GPU, for the store shapes the decoder-chain kernel
// uses, as a function of how many SMs write at the same time.  Prints B/clk/SM and aggregate GB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/micro/store_bench.bin scripts/micro/store_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// pattern 0: a warp writes 512 contiguous bytes per instruction (16 B per lane), consecutive instructions 3584 B apart
//            (the q/k/v image pieces); pattern 1: the same but consecutive instructions contiguous (pure streaming);
// pattern 2: 32 B per lane at a 512 B lane stride (x' rows); pattern 3: 16 KB cp.async.bulk shared -> global
template <int PAT>
__global__ void __launch_bounds__(256, 1) kern(uint8_t* buf, size_t per_cta, int active, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  if ((int)blockIdx.x >= active) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* base = buf + (size_t)blockIdx.x * per_cta;
  for (int i = threadIdx.x; i < 16384 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = i;
  __syncthreads();
  const long long t0 = clock64();
  const uint4 v = make_uint4(lane, warp, 3, 4);
  if (PAT == 0) {
    // per warp a region of per_cta/8; inside it 7 rows of 512 B are interleaved at stride 3584 = 7*512
    uint8_t* wb = base + (size_t)warp * (per_cta / 8);
    const size_t n = per_cta / 8 / 512;
    for (size_t i = 0; i < n; ++i) {
      const size_t blk = i / 7, r = i % 7;   // block of 7 instructions covering 7*3584... keep it simple: permuted order
      const size_t off = (blk * 7 + (r * 3) % 7) * 512;
      asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" :: "l"(wb + off + lane * 16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  } else if (PAT == 1) {
    uint8_t* wb = base + (size_t)warp * (per_cta / 8);
    const size_t n = per_cta / 8 / 512;
    for (size_t i = 0; i < n; ++i)
      asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" :: "l"(wb + i * 512 + lane * 16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  } else if (PAT == 2) {
    // 32 rows of 512 B per warp-instruction group: lane -> row, 16 instructions cover the rows' 512 B
    uint8_t* wb = base + (size_t)warp * (per_cta / 8);
    const size_t n = per_cta / 8 / (32 * 512);
    for (size_t i = 0; i < n; ++i)
      for (int q = 0; q < 16; ++q)
        asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(wb + i * 16384 + lane * 512 + q * 32), "f"(1.0f) : "memory");
  } else {
    if (threadIdx.x == 0) {
      const size_t n = per_cta / 16384;
      for (size_t i = 0; i < n; ++i) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(base + i * 16384), "r"((uint32_t)__cvta_generic_to_shared(smem)), "r"(16384u) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  const size_t per_cta = 8u << 20;
  uint8_t* buf; long long* cyc;
  cudaMalloc(&buf, per_cta * 148); cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(kern<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int pat = 0; pat < 4; ++pat)
    for (int active : {1, 8, 37, 74, 148}) {
      float ms = 0;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (pat == 0) kern<0><<<148, 256, 16384>>>(buf, per_cta, active, cyc);
        if (pat == 1) kern<1><<<148, 256, 16384>>>(buf, per_cta, active, cyc);
        if (pat == 2) kern<2><<<148, 256, 16384>>>(buf, per_cta, active, cyc);
        if (pat == 3) kern<3><<<148, 256, 16384>>>(buf, per_cta, active, cyc);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("pattern %d, %3d SMs writing 8 MB each: %8.3f ms, %7.1f GB/s aggregate, %6.1f B/clk/SM (CTA 0)\n", pat, active, ms,
             (double)per_cta * active / ms / 1e6, (double)per_cta / h[0]);
    }
  return 0;
}
