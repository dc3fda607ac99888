// The row-local part of a post-norm decoder layer as ONE persistent tcgen05 kernel (d = 128, ffn = 256):
//
//   t  = attn . Wo^T + bo + x            y  = LN2( LN1(t) + c )            (out_proj, norm1, zero-memory
//                                                                           cross-attention constant, norm2)
//   f  = relu(y . W1^T + b1)             x' = LN3( y + f . W2^T + b2 )     (linear1, relu, linear2, norm3)
//   qkv' = x' . Win^T + bin                                                 (in_proj of the NEXT layer)
//
// Between two attention kernels everything is local to an activation row, so a 128-row tile (UMMA
// M = 128, TMEM lane = row) goes through all four GEMMs without leaving the SM: per tile the kernel
// reads attn and x (2 x 64 KB) and writes x' and qkv' (64 + 192 KB) -- the five separate launches it
// replaces move 2.5x that.  Precision is the same fp32-faithful scheme as the other tensor-core
// kernels: operands split x = hi + lo (bf16), three accumulating MMAs (lo*hi + hi*lo + hi*hi), fp32
// accumulators in TMEM, all epilogue arithmetic in fp32.
//
// Warp roles (384 threads = 168 registers per thread, 1 CTA / SM, persistent over tiles):
//   warps 0-7   epilogue: TMEM lane quadrant = warp % 4 (row = 32*(warp%4) + lane), column half = warp / 4.
//               Every phase reads its accumulator columns ONCE into registers; the LayerNorm statistics of a
//               row are combined between the two threads that share it through shared memory + a 64-thread
//               named barrier.  E1 (LN1/LN2 -> y: fp32 copy parked in TMEM, hi/lo image to smem), E2
//               (bias+relu -> hi/lo image, 64 hidden units at a time), E3 (LN3 -> x' to HBM + image), E4
//               (bias -> qkv' to HBM).  256-bit global loads/stores: one full 32 B sector per lane.
//   warps 8-9   loaders: attn tile fp32 -> bf16 hi/lo K-major core-matrix image (region Q)
//   warp  10    weight streamer: the four weight matrices are re-tiled ONCE into the exact order the MMA
//               consumes them (32 units of 16 KB per tile, L2 resident), one cp.async.bulk per unit, 4-deep ring
//   warp  11    MMA issuer (one thread) + TMEM allocator
// Shared memory: P 64 KB (y image; relu(f) slots 2-3 during linear2; then x' image) | Q 64 KB (attn image; relu(f)
// slots 0-1 during linear2) | weight ring 64 KB | epilogue vectors 7 KB | statistics exchange 3 KB.
// TMEM (512 columns): [0,256) FFN hidden accumulator, later qkv' columns 0-255 (both N = 256 MMAs: the wide
// shape keeps operand fetch under the shared-memory bandwidth); [256,384) out_proj accumulator, then the fp32
// y (residual of norm3); [384,512) linear2 accumulator, later qkv' columns 256-383.  The next tile's out_proj
// GEMM is issued right after this tile's in_proj GEMM, so it overlaps the qkv' epilogue.
//   reference: nn.TransformerDecoderLayer (post-norm) as configured at model/influentialRS.py:67-74,
//   invoked :189-193 with an all-zero memory (:172-173); model/uRS.py:42-44,62-66.
#include "tc_common.cuh"
#include "qkv_image.cuh"

namespace irs {
namespace tcl {

using namespace irs::tc;

constexpr int BM = 128, D = 128, F = 256;
constexpr int RING = 4;
constexpr uint32_t UNIT_BYTES = 16384, UNIT_HALF = 8192;
constexpr uint32_t A_LBO = BM * 16;                // 2048: K-direction stride between 8-wide slabs of an A image
constexpr uint32_t SBO = 128;
constexpr uint32_t OFF_P = 0;                      // [hi 32 KB | lo 32 KB], K = 128
constexpr uint32_t OFF_Q = 65536;                  // attn image (same layout) / 2 x [hi 16 KB | lo 16 KB], K = 64
constexpr uint32_t OFF_RING = 131072;
constexpr uint32_t OFF_VEC = OFF_RING + RING * UNIT_BYTES;      // 196608
// epilogue vectors (floats)
constexpr int V_BO = 0, V_G1 = 128, V_B1 = 256, V_C2 = 384, V_G2 = 512, V_B2 = 640, V_BF1 = 768, V_BF2 = 1024,
              V_G3 = 1152, V_B3 = 1280, V_BIN = 1408, V_TOTAL = 1792;
constexpr uint32_t OFF_XCH = OFF_VEC + V_TOTAL * 4;             // float2 [3 exchanges][2 halves][128 rows]
constexpr uint32_t OFF_BARS = OFF_XCH + 3 * 2 * 128 * 8;
enum Bars { B_FULL = 0, B_EMPTY = 4, B_A0 = 8, B_D1 = 9, B_A1 = 10, B_D2 = 11, B_A2 = 12 /* 4: one per relu slot */, B_D3 = 16,
            B_A3 = 17, B_D4 = 18, B_COUNT = 20 };
constexpr uint32_t OFF_TMEM = OFF_BARS + B_COUNT * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16;
constexpr int EPI_WARPS = 8, EPI_THREADS = EPI_WARPS * 32;
constexpr int LOAD_WARPS = 2;                      // 12 warps in all: 384 threads leave 168 registers per thread
constexpr int WARP_LOAD0 = 8, WARP_PROD = 10, WARP_MMA = 11;
constexpr int THREADS = 12 * 32;
constexpr int UNITS_BODY = 20;                     // out_proj 4 + linear1 8 + linear2 8
constexpr int UNITS_QKV = 12;
constexpr uint32_t T_H = 0, T_Y = 256, T_D3 = 384; // TMEM column layout

struct Params {
  const float* attn;      // [R, 128]
  const float* x;         // [R, 128] residual stream in
  float* x_out;           // [R, 128] (may alias x)
  float* qkv_out;         // [R, 384] fp32, or null
  uint8_t* qkv_images;    // operand images of the attention kernel (qkv_image.cuh), or null; with both null Win is not applied
  int L, mask_mode;       // window length and mask mode (image addressing: token -> q slot / key column)
  int qkv_only;           // 1: `attn` holds x itself; only qkv' = x Win^T + bin is computed (first layer's in_proj)
  const uint4* wstream;   // prepared weight stream
  const float* vec[11];   // bo g1 b1 c2 g2 b2 bf1 bf2 g3 b3 bin
  float eps1, eps2, eps3;
  int64_t R; int64_t n_tiles; int n_units;
  int* error_flag;
  long long* timeline;    // debug: [8 tiles][3 roles][32 events] clock64 stamps of CTA 0 (null in production)
};

// ---- weight stream ------------------------------------------------------------------------------
// Unit kinds: 0 = [part hi|lo][4 slabs][128 rows][8 bf16]  (N = 128, K = 32)
//             2 = [part hi|lo][2 slabs][256 rows][8 bf16]  (N = 256, K = 16)
struct UnitDesc { int mat, n0, k0, kind; };         // mat: 0 Wo[128,128] 1 W1[256,128] 2 W2[128,256] 3 Win[384,128]
__host__ __device__ inline UnitDesc unit_desc(int u) {
  UnitDesc d;
  if (u < 4) { d.mat = 0; d.n0 = 0; d.k0 = 32 * u; d.kind = 0; }
  else if (u < 12) { d.mat = 1; d.n0 = 0; d.k0 = 16 * (u - 4); d.kind = 2; }
  else if (u < 20) { d.mat = 2; d.n0 = 0; d.k0 = 32 * (u - 12); d.kind = 0; }
  else if (u < 28) { d.mat = 3; d.n0 = 0; d.k0 = 16 * (u - 20); d.kind = 2; }
  else { d.mat = 3; d.n0 = 256; d.k0 = 32 * (u - 28); d.kind = 0; }
  return d;
}

__global__ void __launch_bounds__(256)
prepare_chain_kernel(const float* __restrict__ Wo, const float* __restrict__ W1, const float* __restrict__ W2,
                     const float* __restrict__ Win, int n_units, uint4* __restrict__ out) {
  const int total = n_units * 1024;                 // 16-byte elements
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int u = idx >> 10, w = idx & 1023;
    const int part = w >> 9, rem = w & 511;
    const UnitDesc ud = unit_desc(u);
    int slab, row;
    if (ud.kind == 0) { slab = rem >> 7; row = rem & 127; } else { slab = rem >> 8; row = rem & 255; }
    const float* W; int ldw;
    switch (ud.mat) { case 0: W = Wo; ldw = D; break; case 1: W = W1; ldw = D; break; case 2: W = W2; ldw = F; break;
                      default: W = Win; ldw = D; break; }
    const float* src = W + (int64_t)(ud.n0 + row) * ldw + ud.k0 + slab * 8;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = src[e];
    uint4 hi, lo;
    split8(x, hi, lo);
    out[idx] = part ? lo : hi;
  }
}

// ---- 256-bit global accesses ----------------------------------------------------------------------
__device__ __forceinline__ void ldg256_nc(const float* p, float (&a)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]) : "l"(p));
}
__device__ __forceinline__ void ldg256(const float* p, float (&a)[8]) {
  asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(float* p, const float* a) {
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]) : "memory");
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}

// One streamed weight unit against an A image: KSTEPS k16-steps of three MMAs.
template <int KSTEPS>
__device__ __forceinline__ void mma_unit(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_stage, uint32_t b_lbo,
                                         uint32_t idesc, bool first) {
#pragma unroll
  for (int kk = 0; kk < KSTEPS; ++kk) {
    const uint64_t ah = make_desc(a_hi + (uint32_t)(2 * kk) * A_LBO, A_LBO, SBO);
    const uint64_t al = make_desc(a_lo + (uint32_t)(2 * kk) * A_LBO, A_LBO, SBO);
    const uint64_t bh = make_desc(b_stage + (uint32_t)(2 * kk) * b_lbo, b_lbo, SBO);
    const uint64_t bl = make_desc(b_stage + UNIT_HALF + (uint32_t)(2 * kk) * b_lbo, b_lbo, SBO);
    tc_mma_bf16(d_tmem, al, bh, idesc, (first && kk == 0) ? 0u : 1u);
    tc_mma_bf16(d_tmem, ah, bl, idesc, 1u);
    tc_mma_bf16(d_tmem, ah, bh, idesc, 1u);
  }
}

#define IRS_TL(role, idx)                                                                         \
  do {                                                                                            \
    if (p.timeline && blockIdx.x == 0 && it < 8 && lane == 0)                                     \
      p.timeline[(it * 3 + (role)) * 32 + (idx)] = clock64();                                     \
  } while (0)

__device__ __forceinline__ void pair_barrier(int quad) {          // the two epilogue warps that share a lane quadrant
  asm volatile("bar.sync %0, 64;" :: "r"(1 + quad) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
decoder_chain_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto bar = [&](int i) { return sbase + OFF_BARS + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);
  float* vecs = reinterpret_cast<float*>(smem + OFF_VEC);
  const bool with_qkv = (p.qkv_out != nullptr) || (p.qkv_images != nullptr);
  const bool qkv_only = p.qkv_only != 0;

  if (tid == 0) {
    for (int s = 0; s < RING; ++s) { mbar_init(bar(B_FULL + s), 1); mbar_init(bar(B_EMPTY + s), 1); }
    mbar_init(bar(B_A0), LOAD_WARPS);
    mbar_init(bar(B_D1), 1);
    mbar_init(bar(B_A1), EPI_THREADS);
    mbar_init(bar(B_D2), 1);
    for (int i = 0; i < 4; ++i) mbar_init(bar(B_A2 + i), EPI_THREADS);
    mbar_init(bar(B_D3), 1);
    mbar_init(bar(B_A3), EPI_THREADS);
    for (int i = 0; i < 2; ++i) mbar_init(bar(B_D4 + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(sbase + OFF_TMEM), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const int vlen[11] = {128, 128, 128, 128, 128, 128, 256, 128, 128, 128, 384};
    int off = 0;
    const float qscale0 = 1.4426950408889634f / sqrtf((float)img::DH);
    for (int v = 0; v < 11; ++v) {
      for (int i = tid; i < vlen[v]; i += THREADS) {
        float x = p.vec[v] ? p.vec[v][i] : 0.f;
        if (v == 2 && p.vec[3]) x += p.vec[3][i];                       // norm1 bias + cross-attention constant, added once
        if (v == 10 && p.qkv_images && i < D) x *= qscale0;             // q columns of the images carry log2(e)/sqrt(dh)
        vecs[off + i] = x;
      }
      off += vlen[v];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int64_t first_tile = blockIdx.x, tile_step = gridDim.x;

  if (warp >= WARP_LOAD0 && warp < WARP_LOAD0 + LOAD_WARPS) {
    // ===== loaders: attn tile -> hi/lo image in region Q =====
    const int lw = warp - WARP_LOAD0;
    const int r8 = lane & 7, sg = lane >> 3;
    const uint32_t a0_off = qkv_only ? OFF_P : OFF_Q;
    int it = 0;
    for (int64_t tile = first_tile; tile < p.n_tiles; tile += tile_step, ++it) {
      const int64_t r0 = tile * BM;
#pragma unroll 1
      for (int batch = 0; batch < 64 / LOAD_WARPS / 8; ++batch) {      // 16 row groups x 4 slab groups = 64 units, 8 per batch
        float v[8][8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int qq = batch * 8 + q;
          const int row = (lw * (16 / LOAD_WARPS) + (qq >> 2)) * 8 + r8;
          const int slab = (qq & 3) * 4 + sg;
          if (r0 + row < p.R) ldg256_nc(p.attn + (r0 + row) * D + slab * 8, v[q]);
          else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[q][e] = 0.f;
          }
        }
        if (batch == 0 && lw == 0) IRS_TL(2, 0);
        // region free: Q after this CTA's previous linear2 GEMM; in qkv_only mode P after the previous in_proj GEMM
        if (batch == 0 && it > 0) mbar_wait(bar(qkv_only ? B_D4 + 1 : B_D3), (uint32_t)((it - 1) & 1), p.error_flag, 31);
        if (batch == 0 && lw == 0) IRS_TL(2, 1);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int qq = batch * 8 + q;
          const int row = (lw * (16 / LOAD_WARPS) + (qq >> 2)) * 8 + r8;
          const int slab = (qq & 3) * 4 + sg;
          uint4 hi, lo;
          split8(v[q], hi, lo);
          *reinterpret_cast<uint4*>(smem + a0_off + slab * A_LBO + row * 16) = hi;
          *reinterpret_cast<uint4*>(smem + a0_off + 32768 + slab * A_LBO + row * 16) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_A0));
      if (lw == 0) IRS_TL(2, 2);
    }
  } else if (warp == WARP_PROD) {
    // ===== weight stream: n_units x 16 KB per tile through a 4-deep ring; L2 prefetch of the next tile's rows =====
    if (lane == 0) {
      int64_t g = 0;
      for (int64_t tile = first_tile; tile < p.n_tiles; tile += tile_step) {
        const int64_t nt = tile + tile_step;
        if (nt < p.n_tiles) {
          const int64_t rows = (p.R - nt * BM) < BM ? (p.R - nt * BM) : BM;
          if (p.x) prefetch_l2_bulk(p.x + nt * BM * D, (uint32_t)(rows * D * 4));
          prefetch_l2_bulk(p.attn + nt * BM * D, (uint32_t)(rows * D * 4));
        }
        for (int u = 0; u < p.n_units; ++u, ++g) {
          const int stage = (int)(g % RING);
          const uint32_t phase = (uint32_t)((g / RING) & 1);
          mbar_wait(bar(B_EMPTY + stage), phase ^ 1u, p.error_flag, 32);
          mbar_arrive_expect_tx(bar(B_FULL + stage), UNIT_BYTES);
          bulk_g2s(sbase + OFF_RING + stage * UNIT_BYTES, p.wstream + (int64_t)u * (UNIT_BYTES / 16), UNIT_BYTES, bar(B_FULL + stage));
        }
      }
    }
  } else if (warp == WARP_MMA) {
    if (lane == 0) {
      const uint32_t idesc128 = make_idesc_bf16(BM, 128), idesc256 = make_idesc_bf16(BM, 256);
      const uint32_t P_HI = sbase + OFF_P, P_LO = sbase + OFF_P + 32768, Q_HI = sbase + OFF_Q, Q_LO = sbase + OFF_Q + 32768;
      int64_t g = 0;
      auto next_unit = [&]() -> uint32_t {          // waits for the next streamed unit; returns its smem address
        const int stage = (int)(g % RING);
        mbar_wait(bar(B_FULL + stage), (uint32_t)((g / RING) & 1), p.error_flag, 33);
        tc_fence_after();
        return sbase + OFF_RING + stage * UNIT_BYTES;
      };
      auto release_unit = [&]() { tc_commit(bar(B_EMPTY + (int)(g % RING))); ++g; };
      int it = 0;
      for (int64_t tile = first_tile; tile < p.n_tiles; tile += tile_step, ++it) {
        const uint32_t ph = (uint32_t)(it & 1);
        if (qkv_only) {
          mbar_wait(bar(B_A0), ph, p.error_flag, 34);                          // x image in P
          if (it > 0) mbar_wait(bar(B_A1), (uint32_t)((it - 1) & 1), p.error_flag, 38);   // previous tile's accumulators drained
          tc_fence_after();
          for (int u = 0; u < 8; ++u) {
            const uint32_t bs = next_unit();
            mma_unit<1>(tmem_base + T_H, P_HI + (uint32_t)(2 * u) * A_LBO, P_LO + (uint32_t)(2 * u) * A_LBO, bs, 4096u, idesc256, u == 0);
            release_unit();
          }
          tc_commit(bar(B_D4 + 0));
          for (int c = 0; c < 4; ++c) {
            const uint32_t bs = next_unit();
            mma_unit<2>(tmem_base + T_D3, P_HI + (uint32_t)(4 * c) * A_LBO, P_LO + (uint32_t)(4 * c) * A_LBO, bs, 2048u, idesc128, c == 0);
            release_unit();
          }
          tc_commit(bar(B_D4 + 1));
          continue;
        }
        // ---- G1: out_proj -> T_Y
        IRS_TL(0, 0);
        mbar_wait(bar(B_A0), ph, p.error_flag, 34);
        tc_fence_after();
        IRS_TL(0, 1);
        for (int c = 0; c < 4; ++c) {
          const uint32_t bs = next_unit();
          mma_unit<2>(tmem_base + T_Y, Q_HI + (uint32_t)(4 * c) * A_LBO, Q_LO + (uint32_t)(4 * c) * A_LBO, bs, 2048u, idesc128, c == 0);
          release_unit();
        }
        tc_commit(bar(B_D1));
        IRS_TL(0, 2);
        // ---- G2: linear1, all 256 hidden units at once -> T_H
        mbar_wait(bar(B_A1), ph, p.error_flag, 35);
        tc_fence_after();
        IRS_TL(0, 3);
        for (int u = 0; u < 8; ++u) {
          const uint32_t bs = next_unit();
          mma_unit<1>(tmem_base + T_H, P_HI + (uint32_t)(2 * u) * A_LBO, P_LO + (uint32_t)(2 * u) * A_LBO, bs, 4096u, idesc256, u == 0);
          release_unit();
        }
        tc_commit(bar(B_D2));
        // ---- G3: linear2 over the relu images, 64 hidden units (one 32 KB slot) at a time -> T_D3.  Four slots: region Q
        // (the attn image is consumed) and region P (the y image is consumed once linear1 has completed), so the relu
        // conversion of all four pieces runs ahead of the MMAs instead of alternating with them.
        for (int j = 0; j < 4; ++j) {
          IRS_TL(0, 4 + 2 * j);
          mbar_wait(bar(B_A2 + j), ph, p.error_flag, 36);
          tc_fence_after();
          IRS_TL(0, 5 + 2 * j);
          const uint32_t slot = (j < 2 ? Q_HI : P_HI) + (uint32_t)(j & 1) * 32768u;
          for (int i = 0; i < 2; ++i) {
            const uint32_t bs = next_unit();
            mma_unit<2>(tmem_base + T_D3, slot + (uint32_t)(4 * i) * A_LBO, slot + 16384u + (uint32_t)(4 * i) * A_LBO, bs,
                        2048u, idesc128, j == 0 && i == 0);
            release_unit();
          }
        }
        tc_commit(bar(B_D3));
        IRS_TL(0, 12);
        // ---- G4: in_proj of the next layer: columns 0-255 -> T_H (N = 256), 256-383 -> T_D3
        mbar_wait(bar(B_A3), ph, p.error_flag, 37);     // E3 has read y / D3 (and written the x' image)
        tc_fence_after();
        IRS_TL(0, 13);
        if (with_qkv) {
          for (int u = 0; u < 8; ++u) {
            const uint32_t bs = next_unit();
            mma_unit<1>(tmem_base + T_H, P_HI + (uint32_t)(2 * u) * A_LBO, P_LO + (uint32_t)(2 * u) * A_LBO, bs, 4096u, idesc256, u == 0);
            release_unit();
          }
          tc_commit(bar(B_D4 + 0));
          IRS_TL(0, 14);
          for (int c = 0; c < 4; ++c) {
            const uint32_t bs = next_unit();
            mma_unit<2>(tmem_base + T_D3, P_HI + (uint32_t)(4 * c) * A_LBO, P_LO + (uint32_t)(4 * c) * A_LBO, bs, 2048u, idesc128, c == 0);
            release_unit();
          }
          tc_commit(bar(B_D4 + 1));
          IRS_TL(0, 15);
        }
      }
    }
  } else if (warp < EPI_WARPS) {
    // ===== epilogue: (lane quadrant, column half) =====
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t tlane = tmem_base + (((uint32_t)(quad * 32)) << 16);
    float2* xch = reinterpret_cast<float2*>(smem + OFF_XCH);
    const float inv_n = 1.0f / (float)D;
    const int c0 = half * 64;                        // this thread's columns of a 128-wide row
    // ---------------- E4: qkv' = acc + bin -> HBM, one piece (0: columns [0,256) in T_H, 1: [256,384) in T_D3) ----------------
    // fp32 rows [R, 384], or operand images: q (pre-scaled) / k / v of every head as bf16 hi/lo 16-byte pieces.
    // An SM retires 32 B of global stores per clock (scripts/micro/store_bench.cu), so the 192 KB of a tile are >= 6k cycles
    // during which the storing warps stall.  Piece 0 therefore runs while the tensor core works through in_proj piece 1 and
    // the next tile's out_proj (the next E1 cannot start before that anyway), and piece 1 is deferred behind the next
    // tile's E1, into the wait for its linear1 -- T_D3 is not overwritten before that tile's linear2.
    auto e4_piece = [&](int64_t tile, int it, int piece) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int64_t r = tile * BM + row;
      const bool row_ok = r < p.R;
      float* qo = p.qkv_out ? p.qkv_out + (row_ok ? r : 0) * (3 * D) : nullptr;
      const int bb = (int)(r / p.L), ll = (int)(r - (int64_t)bb * p.L);
      const int n_chunks = (p.L + 31) / 32;
      const int qs = img::q_slot(n_chunks, ll);
      const uint32_t q_off = img::OFF_Q + (uint32_t)(qs / BM) * img::Q_TILE + (uint32_t)(qs % BM) * 16u;
      const uint32_t kv_off = (uint32_t)img::kv_col(p.mask_mode == IRS_MASK_PIM, p.L, ll) * 16u;
      uint8_t* item0 = p.qkv_images ? p.qkv_images + (int64_t)bb * (D / img::DH) * img::ITEM_BYTES : nullptr;
      const float qscale = 1.4426950408889634f / sqrtf((float)img::DH);
      {
      mbar_wait(bar(B_D4 + piece), ph, p.error_flag, 45);
      tc_fence_after();
      if (warp == 0) IRS_TL(1, 13 + 2 * piece);
      // piece 0: qkv' columns [0,256) in T_H, 128 per thread; piece 1: columns [256,384) in T_D3, 64 per thread
      const int ncol = piece == 0 ? 128 : 64;
      const int col0 = piece == 0 ? half * 128 : 256 + half * 64;
      const uint32_t tcol = piece == 0 ? (T_H + (uint32_t)half * 128u) : (T_D3 + (uint32_t)half * 64u);
      auto emit = [&](const uint32_t (&v)[32], int ch) {
        float o32[32];
        const float sc = (p.qkv_images && col0 + ch * 32 < D) ? qscale : 1.0f;        // q columns: scale folded into one FFMA
#pragma unroll
        for (int j = 0; j < 32; ++j) o32[j] = fmaf(__uint_as_float(v[j]), sc, vecs[V_BIN + col0 + ch * 32 + j]);
        if (qo) {
          if (row_ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) stg256(qo + col0 + ch * 32 + q * 8, &o32[q * 8]);
          }
        } else if (row_ok) {
          // 32 columns = one head of q, k or v: four 8-wide slabs, consecutive tokens 16 bytes apart
          const int c = col0 + ch * 32;
          const int which = c >> 7, head = (c & 127) >> 5;
          uint8_t* dst = item0 + (int64_t)head * img::ITEM_BYTES + (which == 0 ? q_off : (which == 1 ? img::OFF_K : img::OFF_V) + kv_off);
          const uint32_t lbo = which == 0 ? img::Q_LBO : img::K_LBO, part = which == 0 ? img::Q_PART : img::K_PART;
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8) {
            uint4 hi, lo;
            split8(*reinterpret_cast<float(*)[8]>(&o32[s8 * 8]), hi, lo);
            stg128(dst + s8 * lbo, hi);
            stg128(dst + part + s8 * lbo, lo);
          }
        }
      };
      // one TMEM load in flight behind the conversion + stores of the previous chunk
      {
        const int nch = ncol / 32;
        uint32_t va[32], vb[32];
        tc_ld32(tlane + tcol, va);
#pragma unroll 1
        for (int ch = 0; ch < nch; ch += 2) {
          tc_wait_ld();
          if (ch + 1 < nch) tc_ld32(tlane + tcol + (ch + 1) * 32, vb);
          emit(va, ch);
          if (ch + 1 < nch) {
            tc_wait_ld();
            if (ch + 2 < nch) tc_ld32(tlane + tcol + (ch + 2) * 32, va);
            emit(vb, ch + 1);
          }
        }
      }
      if (warp == 0) IRS_TL(1, 14 + 2 * piece);
      }
    };
    int it = 0;
    for (int64_t tile = first_tile; tile < p.n_tiles; tile += tile_step, ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int64_t r = tile * BM + row;
      const bool row_ok = r < p.R;
      if (!qkv_only) {
      // ---------------- E1: t = acc + bo + x ; y = LN2(LN1(t) + c) ----------------
      float t[64];
      {
        const float* xrow = p.x + (row_ok ? r : 0) * D + c0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (row_ok) ldg256(xrow + q * 8, *reinterpret_cast<float(*)[8]>(&t[q * 8]));
          else {
#pragma unroll
            for (int e = 0; e < 8; ++e) t[q * 8 + e] = 0.f;
          }
        }
      }
      if (warp == 0) IRS_TL(1, 0);
      mbar_wait(bar(B_D1), ph, p.error_flag, 41);
      tc_fence_after();
      if (warp == 0) IRS_TL(1, 1);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[32];
        tc_ld32(tlane + T_Y + c0 + ch * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float tt = __uint_as_float(v[j]) + vecs[V_BO + c0 + ch * 32 + j] + t[ch * 32 + j];
          s1 += tt; s2 = fmaf(tt, tt, s2);
          t[ch * 32 + j] = tt;
        }
      }
      xch[(0 * 2 + half) * 128 + row] = make_float2(s1, s2);
      pair_barrier(quad);
      {
        const float2 o = xch[(0 * 2 + (half ^ 1)) * 128 + row];
        s1 += o.x; s2 += o.y;
      }
      const float mean1 = s1 * inv_n;
      const float rstd1 = rsqrtf(fmaxf(s2 * inv_n - mean1 * mean1, 0.f) + p.eps1);
      const float nmr1 = -mean1 * rstd1;
      float u1 = 0.f, u2 = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const int n = c0 + j;
        const float y = fmaf(fmaf(t[j], rstd1, nmr1), vecs[V_G1 + n], vecs[V_B1 + n]);      // V_B1 holds b1 + c2
        u1 += y; u2 = fmaf(y, y, u2);
        t[j] = y;
      }
      xch[(1 * 2 + half) * 128 + row] = make_float2(u1, u2);
      pair_barrier(quad);
      {
        const float2 o = xch[(1 * 2 + (half ^ 1)) * 128 + row];
        u1 += o.x; u2 += o.y;
      }
      const float mean2 = u1 * inv_n;
      const float rstd2 = rsqrtf(fmaxf(u2 * inv_n - mean2 * mean2, 0.f) + p.eps2);
      const float nmr2 = -mean2 * rstd2;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = c0 + ch * 32 + j;
          const float y = fmaf(fmaf(t[ch * 32 + j], rstd2, nmr2), vecs[V_G2 + n], vecs[V_B2 + n]);
          t[ch * 32 + j] = y;
          v[j] = __float_as_uint(y);
        }
        tc_st32(tlane + T_Y + c0 + ch * 32, v);      // fp32 y stays in TMEM: residual of norm3
#pragma unroll
        for (int s8 = 0; s8 < 4; ++s8) {
          uint4 hi, lo;
          split8(*reinterpret_cast<float(*)[8]>(&t[ch * 32 + s8 * 8]), hi, lo);
          const uint32_t o = (uint32_t)(half * 8 + ch * 4 + s8) * A_LBO + (uint32_t)row * 16u;
          *reinterpret_cast<uint4*>(smem + OFF_P + o) = hi;
          *reinterpret_cast<uint4*>(smem + OFF_P + 32768 + o) = lo;
        }
      }
      tc_wait_st();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      mbar_arrive(bar(B_A1));
      if (warp == 0) IRS_TL(1, 2);
      if (with_qkv && it > 0) e4_piece(tile - tile_step, it - 1, 1);         // previous tile's v columns, behind this E1
      // ---------------- E2: relu(f + b1) -> image, 64 hidden units at a time (32 per thread) ----------------
      mbar_wait(bar(B_D2), ph, p.error_flag, 42);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        uint32_t v[32];
        tc_ld32(tlane + T_H + 64u * (uint32_t)j + 32u * (uint32_t)half, v);
        if (warp == 0) IRS_TL(1, 3 + 2 * j);
        uint8_t* slot = smem + (j < 2 ? OFF_Q : OFF_P) + (j & 1) * 32768;      // P: linear1 (D2) has consumed the y image
        tc_wait_ld();
#pragma unroll
        for (int s8 = 0; s8 < 4; ++s8) {
          float x8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            x8[e] = fmaxf(__uint_as_float(v[s8 * 8 + e]) + vecs[V_BF1 + j * 64 + half * 32 + s8 * 8 + e], 0.f);
          uint4 hi, lo;
          split8(x8, hi, lo);
          const uint32_t o = (uint32_t)(half * 4 + s8) * A_LBO + (uint32_t)row * 16u;
          *reinterpret_cast<uint4*>(slot + o) = hi;
          *reinterpret_cast<uint4*>(slot + 16384 + o) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        mbar_arrive(bar(B_A2 + j));
        if (warp == 0) IRS_TL(1, 4 + 2 * j);
      }
      // ---------------- E3: x' = LN3(y + acc + b2) ----------------
      mbar_wait(bar(B_D3), ph, p.error_flag, 44);
      tc_fence_after();
      if (warp == 0) IRS_TL(1, 11);
      float w1 = 0.f, w2 = 0.f;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[32], y[32];
        tc_ld32(tlane + T_D3 + c0 + ch * 32, v);
        tc_ld32(tlane + T_Y + c0 + ch * 32, y);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float tt = __uint_as_float(v[j]) + vecs[V_BF2 + c0 + ch * 32 + j] + __uint_as_float(y[j]);
          w1 += tt; w2 = fmaf(tt, tt, w2);
          t[ch * 32 + j] = tt;
        }
      }
      xch[(2 * 2 + half) * 128 + row] = make_float2(w1, w2);
      pair_barrier(quad);
      {
        const float2 o = xch[(2 * 2 + (half ^ 1)) * 128 + row];
        w1 += o.x; w2 += o.y;
      }
      const float mean3 = w1 * inv_n;
      const float rstd3 = rsqrtf(fmaxf(w2 * inv_n - mean3 * mean3, 0.f) + p.eps3);
      const float nmr3 = -mean3 * rstd3;
      float* xo = p.x_out + (row_ok ? r : 0) * D + c0;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = c0 + ch * 32 + j;
          t[ch * 32 + j] = fmaf(fmaf(t[ch * 32 + j], rstd3, nmr3), vecs[V_G3 + n], vecs[V_B3 + n]);
        }
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) stg256(xo + ch * 32 + q * 8, &t[ch * 32 + q * 8]);
        }
        if (with_qkv) {
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8) {
            uint4 hi, lo;
            split8(*reinterpret_cast<float(*)[8]>(&t[ch * 32 + s8 * 8]), hi, lo);
            const uint32_t o = (uint32_t)(half * 8 + ch * 4 + s8) * A_LBO + (uint32_t)row * 16u;
            *reinterpret_cast<uint4*>(smem + OFF_P + o) = hi;
            *reinterpret_cast<uint4*>(smem + OFF_P + 32768 + o) = lo;
          }
        }
      }
      if (with_qkv) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      mbar_arrive(bar(B_A3));
      if (warp == 0) IRS_TL(1, 12);
      }  // !qkv_only
      if (qkv_only) {
        e4_piece(tile, it, 0);
        e4_piece(tile, it, 1);
        tc_fence_before();
        mbar_arrive(bar(B_A1));                                              // accumulators drained
      } else if (with_qkv) {
        e4_piece(tile, it, 0);                                               // piece 1: after the next tile's E1
      }
      tc_fence_before();
    }
    if (!qkv_only && with_qkv && it > 0) e4_piece(first_tile + (int64_t)(it - 1) * tile_step, it - 1, 1);
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace tcl
}  // namespace irs

using namespace irs;

static long long* g_chain_timeline = nullptr;

/* debug hook (not in the public header): device buffer of 8*3*32 int64 receiving CTA 0's phase time stamps */
extern "C" void irs_decoder_chain_debug_timeline(long long* buf) { g_chain_timeline = buf; }

extern "C" int irs_decoder_chain_supported(int d, int ffn) { return (d == tcl::D && ffn == tcl::F) ? 1 : 0; }

extern "C" size_t irs_decoder_chain_prepared_bytes(int d, int ffn, int with_in_proj) {
  if (!irs_decoder_chain_supported(d, ffn)) return 0;
  return (size_t)(tcl::UNITS_BODY + (with_in_proj ? tcl::UNITS_QKV : 0)) * tcl::UNIT_BYTES;
}

extern "C" int irs_decoder_chain_prepare_weights(const float* Wo, const float* W1, const float* W2, const float* Win,
                                                 int d, int ffn, void* prepared, void* stream) {
  if (!Wo || !W1 || !W2 || !prepared) return IRS_E_BADARG;
  if (!irs_decoder_chain_supported(d, ffn)) return IRS_E_SHAPE;
  const int n_units = tcl::UNITS_BODY + (Win ? tcl::UNITS_QKV : 0);
  tcl::prepare_chain_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(Wo, W1, W2, Win, n_units, (uint4*)prepared);
  IRS_LAUNCHED();
  return 0;
}

static int chain_launch(tcl::Params& p, int64_t R, int* error_flag, void* stream) {
  p.R = R; p.n_tiles = ceil_div(R, tcl::BM);
  p.error_flag = error_flag;
  p.timeline = g_chain_timeline;
  static bool configured = false;
  if (!configured) {
    IRS_CUDA(cudaFuncSetAttribute(tcl::decoder_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcl::SMEM_BYTES));
    configured = true;
  }
  const unsigned grid = (unsigned)(p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs);
  tcl::decoder_chain_kernel<<<grid, tcl::THREADS, tcl::SMEM_BYTES, (cudaStream_t)stream>>>(p);
  IRS_LAUNCHED();
  return 0;
}

extern "C" int irs_decoder_chain_tc(const float* attn, const float* x, const void* prepared,
                                    const float* bo, const float* g1, const float* b1, const float* c2,
                                    const float* g2, const float* b2, const float* bf1, const float* bf2,
                                    const float* g3, const float* b3, const float* bin,
                                    float eps1, float eps2, float eps3,
                                    float* x_out, float* qkv_out, void* qkv_images, int L, int mask_mode,
                                    int64_t R, int d, int ffn, int* error_flag, void* stream) {
  if (!attn || !x || !prepared || !x_out || !g1 || !b1 || !g2 || !b2 || !g3 || !b3) return IRS_E_BADARG;
  if (R < 0 || (qkv_out && qkv_images)) return IRS_E_BADARG;
  if (R == 0) return 0;
  if (!irs_decoder_chain_supported(d, ffn)) return IRS_E_SHAPE;
  if (qkv_images && (L <= 128 || L > irs::img::KEYS - 1 || R % L != 0 || mask_mode < 0 || mask_mode > 2)) return IRS_E_SHAPE;
  if (((uintptr_t)attn & 31) || ((uintptr_t)x & 31) || ((uintptr_t)x_out & 31) || ((uintptr_t)qkv_out & 31) ||
      ((uintptr_t)qkv_images & 15) || ((uintptr_t)prepared & 15))
    return IRS_E_SHAPE;
  tcl::Params p = {};
  p.attn = attn; p.x = x; p.x_out = x_out; p.qkv_out = qkv_out; p.qkv_images = (uint8_t*)qkv_images;
  p.L = qkv_images ? L : 1; p.mask_mode = mask_mode; p.qkv_only = 0;
  p.wstream = (const uint4*)prepared;
  const float* vec[11] = {bo, g1, b1, c2, g2, b2, bf1, bf2, g3, b3, bin};
  for (int i = 0; i < 11; ++i) p.vec[i] = vec[i];
  p.eps1 = eps1; p.eps2 = eps2; p.eps3 = eps3;
  p.n_units = tcl::UNITS_BODY + ((qkv_out || qkv_images) ? tcl::UNITS_QKV : 0);
  return chain_launch(p, R, error_flag, stream);
}

extern "C" int irs_in_proj_images_tc(const float* x, const void* prepared_with_in_proj, const float* bin,
                                     void* qkv_images, int L, int mask_mode, int64_t R, int d,
                                     int* error_flag, void* stream) {
  if (!x || !prepared_with_in_proj || !qkv_images) return IRS_E_BADARG;
  if (R < 0) return IRS_E_BADARG;
  if (R == 0) return 0;
  if (d != tcl::D) return IRS_E_SHAPE;
  if (L <= 128 || L > irs::img::KEYS - 1 || R % L != 0 || mask_mode < 0 || mask_mode > 2) return IRS_E_SHAPE;
  if (((uintptr_t)x & 31) || ((uintptr_t)qkv_images & 15) || ((uintptr_t)prepared_with_in_proj & 15)) return IRS_E_SHAPE;
  tcl::Params p = {};
  p.attn = x; p.x = nullptr; p.x_out = nullptr; p.qkv_out = nullptr; p.qkv_images = (uint8_t*)qkv_images;
  p.L = L; p.mask_mode = mask_mode; p.qkv_only = 1;
  // the in_proj units are the tail of a full chain stream
  p.wstream = (const uint4*)prepared_with_in_proj + (size_t)tcl::UNITS_BODY * (tcl::UNIT_BYTES / 16);
  p.vec[10] = bin;
  p.n_units = tcl::UNITS_QKV;
  return chain_launch(p, R, error_flag, stream);
}
