// K3, persistent edition: PIM / causal self-attention for full windows of 129..223 positions with 32-wide
// heads (the BASELINE cfg3/cfg5 shape: L = 201, dh = 32) on the 5th-generation tensor cores.
//
// One persistent CTA per SM walks over (batch, head) items.  Compared with the one-CTA-per-head kernel in
// attention_tc.cu (still used for L <= 128, other head sizes and row subsets) it
//   * overlaps global traffic with math: four loader warps stage the NEXT item's K, V^T and Q as bf16 hi/lo
//     K-major core-matrix images into the second half of a double buffer while the softmax warps work;
//   * keeps the probabilities out of shared memory: P = exp2(s - max) is written back IN PLACE over the
//     scores in TMEM as packed bf16 pairs (hi in 16 columns, lo in the next 16 of every 32-key block) and
//     the P.V product reads its A operand from TMEM (tcgen05.mma with a tensor-memory A operand);
//   * balances the causal triangle: the window is cut into 32-row chunks and the two softmax warpgroups
//     (two M = 128 tiles in flight, 256 TMEM columns each) take chunks such that every SM sub-partition
//     sees about the same number of visible key blocks (long rows paired with short rows).
// Arithmetic is that of attention_tc.cu: scores in the log2 domain, three bf16 MMAs per product
// (lo*hi + hi*lo + hi*hi) with fp32 accumulation, softmax in fp32.
// Warps: 0-3 softmax group 0, 4-7 softmax group 1 (TMEM lane quadrant = warp % 4), 8-11 loaders, 12 MMA issuer.
//   reference: model/influentialRS.py:139-151,171,189-193; torch multi_head_attention_forward; model/uRS.py:47-61.
#include "tc_common.cuh"

namespace irs {
namespace tcp {

using namespace irs::tc;

constexpr int BM = 128, KEYS = 224, PB = 32, DH = 32, SLABS = DH / 8;
constexpr int THREADS = 13 * 32;
constexpr int WARP_LOAD0 = 8, WARP_MMA = 12;
constexpr uint32_t SBO = 128;
constexpr uint32_t K_LBO = KEYS * 16, V_LBO = DH * 16, Q_LBO = BM * 16;
constexpr uint32_t K_PART = SLABS * K_LBO;          // 14336
constexpr uint32_t V_PART = (KEYS / 8) * V_LBO;     // 14336
constexpr uint32_t Q_TILE = SLABS * Q_LBO;          // 8192
constexpr uint32_t Q_PART = 2 * Q_TILE;             // 16384
constexpr uint32_t OFF_K_HI = 0, OFF_K_LO = K_PART, OFF_V_HI = 2 * K_PART, OFF_V_LO = 2 * K_PART + V_PART,
                   OFF_Q_HI = 2 * K_PART + 2 * V_PART, OFF_Q_LO = OFF_Q_HI + Q_PART, OFF_KB = OFF_Q_LO + Q_PART;
constexpr uint32_t BUF_BYTES = ((OFF_KB + KEYS * 4 + 127) / 128) * 128;    // 91136
enum Bars { B_KV_FULL = 0, B_KV_FREE = 2, B_S = 4, B_P = 6, B_O = 8, B_COUNT = 10 };
constexpr uint32_t OFF_BARS = 2 * BUF_BYTES;
constexpr uint32_t OFF_TMEM = OFF_BARS + B_COUNT * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16;
constexpr uint32_t O_COL = 224;

struct Params {
  const float* q; const float* k; const float* v; int64_t ld_q, ld_k, ld_v;
  const int64_t* ids; const float* r_u; float w_h, w_obj; int mode;
  float* out; int B, L, H; int n_items;
  int* error_flag;
};

// 32-row chunk handled by softmax warp `quad` of group `g` (-1: none).  Group 0 takes the four longest
// chunks ordered [second longest, longest, third, fourth]; group 1 the remaining short ones on quadrants
// 0, 2, 3 -- so that per quadrant the visible key blocks add up to about the same number.
__device__ __forceinline__ int chunk_of(int n_chunks, int g, int quad) {
  if (g == 0) return quad == 0 ? n_chunks - 2 : (quad == 1 ? n_chunks - 1 : (quad == 2 ? n_chunks - 3 : n_chunks - 4));
  const int c = quad == 0 ? 0 : (quad == 1 ? -1 : quad - 1);
  return (c >= 0 && c < n_chunks - 4) ? c : -1;
}

__device__ __forceinline__ void ldg256_nc(const float* p, float (&a)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* a) {
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T : A operand = packed bf16 pairs, lane = row, column = k / 2.
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t"
               "setp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
pim_attn_persistent_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto bar = [&](int i) { return sbase + OFF_BARS + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);
  const int L = p.L, H = p.H;
  const bool pim = (p.mode == IRS_MASK_PIM);
  const int koff = pim ? 1 : 0;                      // column c <-> key c - koff; column 0 <-> key L-1 in PIM mode
  const int n_chunks = (L + 31) / 32;                // 5..7
  const float l2e = 1.4426950408889634f;
  const float qscale = l2e / sqrtf((float)DH);
  auto key_of = [&](int c) { return (pim && c == 0) ? L - 1 : c - koff; };
  // columns each group's tile can see, padded to the MMA N granularity
  const int rem = n_chunks - 4;
  const int tcols_g0 = L, tcols_g1 = min(32 * rem, L - koff) + koff;
  auto tcols_of = [&](int g) { return g == 0 ? tcols_g0 : tcols_g1; };
  auto tpad_of = [&](int g) { return (tcols_of(g) + 15) & ~15; };

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_KV_FULL + i), 4); mbar_init(bar(B_KV_FREE + i), 1);
      mbar_init(bar(B_S + i), 1); mbar_init(bar(B_P + i), 128); mbar_init(bar(B_O + i), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(sbase + OFF_TMEM), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int first = blockIdx.x, step = gridDim.x;

  if (warp >= WARP_LOAD0 && warp < WARP_LOAD0 + 4) {
    // ===== loaders: stage item `it` into buffer it & 1 =====
    const int lw = warp - WARP_LOAD0, t = tid - WARP_LOAD0 * 32;
    int it = 0;
    for (int item = first; item < p.n_items; item += step, ++it) {
      const int b = item / H, h = item % H;
      uint8_t* buf = smem + (it & 1) * BUF_BYTES;
      const float* qbase = p.q + (int64_t)b * L * p.ld_q + h * DH;
      const float* kbase = p.k + (int64_t)b * L * p.ld_k + h * DH;
      const float* vbase = p.v + (int64_t)b * L * p.ld_v + h * DH;
      bool waited = false;
      auto wait_free = [&]() {
        if (!waited) { mbar_wait(bar(B_KV_FREE + (it & 1)), (uint32_t)(((it >> 1) & 1) ^ 1), p.error_flag, 51); waited = true; }
      };
      // ---- K image: thread <-> key column
      {
        float x[2][4][8];
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const int c = t + rep * 128;
          const bool ok = c < L;
          const float* src = kbase + (int64_t)(ok ? key_of(c) : 0) * p.ld_k;
#pragma unroll
          for (int s = 0; s < SLABS; ++s) {
            if (ok) ldg256_nc(src + s * 8, x[rep][s]);
            else {
#pragma unroll
              for (int e = 0; e < 8; ++e) x[rep][s][e] = 0.f;
            }
          }
        }
        wait_free();
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const int c = t + rep * 128;
          if (c < KEYS) {
#pragma unroll
            for (int s = 0; s < SLABS; ++s) {
              uint4 hi, lo;
              split8(x[rep][s], hi, lo);
              *reinterpret_cast<uint4*>(buf + OFF_K_HI + s * K_LBO + c * 16) = hi;
              *reinterpret_cast<uint4*>(buf + OFF_K_LO + s * K_LBO + c * 16) = lo;
            }
            // per-column additive term of every row that can see the column: l2e * (mask weight + key padding)
            float bias = -INFINITY;
            if (c < L) {
              const bool pad = (p.mode != IRS_MASK_CAUSAL && p.ids[(int64_t)b * L + key_of(c)] == 0);
              bias = pad ? -INFINITY : l2e * ((pim && c == 0) ? p.w_obj * p.r_u[b] : (pim ? p.w_h : 0.f));
            }
            reinterpret_cast<float*>(buf + OFF_KB)[c] = bias;
          }
        }
      }
      // ---- Q images: thread <-> (group, quadrant, lane) row slot, pre-scaled by log2(e)/sqrt(dh)
      {
        float x[2][4][8];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int ch = chunk_of(n_chunks, g, t >> 5);
          const int i = ch * 32 + (t & 31);
          const bool ok = ch >= 0 && i < L;
          const float* src = qbase + (int64_t)(ok ? i : 0) * p.ld_q;
#pragma unroll
          for (int s = 0; s < SLABS; ++s) {
            if (ok) ldg256_nc(src + s * 8, x[g][s]);
            else {
#pragma unroll
              for (int e = 0; e < 8; ++e) x[g][s][e] = 0.f;
            }
          }
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
#pragma unroll
          for (int s = 0; s < SLABS; ++s) {
#pragma unroll
            for (int e = 0; e < 8; ++e) x[g][s][e] *= qscale;
            uint4 hi, lo;
            split8(x[g][s], hi, lo);
            *reinterpret_cast<uint4*>(buf + OFF_Q_HI + g * Q_TILE + s * Q_LBO + t * 16) = hi;
            *reinterpret_cast<uint4*>(buf + OFF_Q_LO + g * Q_TILE + s * Q_LBO + t * 16) = lo;
          }
        }
      }
      // ---- V^T image: warp <-> group of 8 keys, lane <-> head dim (register transpose, coalesced 128-byte rows)
      {
        float x[7][8];
#pragma unroll
        for (int r = 0; r < 7; ++r) {
          const int ks = lw + 4 * r;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int col = ks * 8 + e;
            x[r][e] = (col < L) ? __ldg(vbase + (int64_t)key_of(col) * p.ld_v + lane) : 0.f;
          }
        }
#pragma unroll
        for (int r = 0; r < 7; ++r) {
          const int ks = lw + 4 * r;
          uint4 hi, lo;
          split8(x[r], hi, lo);
          *reinterpret_cast<uint4*>(buf + OFF_V_HI + ks * V_LBO + lane * 16) = hi;
          *reinterpret_cast<uint4*>(buf + OFF_V_LO + ks * V_LBO + lane * 16) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_KV_FULL + (it & 1)));
    }
  } else if (warp == WARP_MMA) {
    if (lane == 0) {
      const uint32_t idesc_o = make_idesc_bf16(BM, DH);
      auto issue_qk = [&](int g, int bufi) {
        const uint32_t bb = sbase + (uint32_t)bufi * BUF_BYTES;
        const uint32_t idesc_s = make_idesc_bf16(BM, tpad_of(g));
        const uint32_t d = tmem_base + (uint32_t)g * 256u;
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          const uint64_t q_hi = make_desc(bb + OFF_Q_HI + (uint32_t)g * Q_TILE + (uint32_t)(kk * 2) * Q_LBO, Q_LBO, SBO);
          const uint64_t q_lo = make_desc(bb + OFF_Q_LO + (uint32_t)g * Q_TILE + (uint32_t)(kk * 2) * Q_LBO, Q_LBO, SBO);
          const uint64_t k_hi = make_desc(bb + OFF_K_HI + (uint32_t)(kk * 2) * K_LBO, K_LBO, SBO);
          const uint64_t k_lo = make_desc(bb + OFF_K_LO + (uint32_t)(kk * 2) * K_LBO, K_LBO, SBO);
          tc_mma_bf16(d, q_lo, k_hi, idesc_s, kk != 0 ? 1u : 0u);
          tc_mma_bf16(d, q_hi, k_lo, idesc_s, 1u);
          tc_mma_bf16(d, q_hi, k_hi, idesc_s, 1u);
        }
        tc_commit(bar(B_S + g));
      };
      auto issue_pv = [&](int g, int bufi) {
        const uint32_t bb = sbase + (uint32_t)bufi * BUF_BYTES;
        const uint32_t ts = tmem_base + (uint32_t)g * 256u;
        const int nblk = (tpad_of(g) + PB - 1) / PB;
        for (int blk = 0; blk < nblk; ++blk) {
#pragma unroll
          for (int kk = 0; kk < PB / 16; ++kk) {
            const uint32_t a_hi = ts + (uint32_t)(blk * PB + kk * 8), a_lo = a_hi + 16u;
            const uint32_t vo = (uint32_t)(blk * (PB / 8) + kk * 2) * V_LBO;
            const uint64_t v_hi = make_desc(bb + OFF_V_HI + vo, V_LBO, SBO);
            const uint64_t v_lo = make_desc(bb + OFF_V_LO + vo, V_LBO, SBO);
            tc_mma_bf16_ts(ts + O_COL, a_lo, v_hi, idesc_o, (blk | kk) != 0 ? 1u : 0u);
            tc_mma_bf16_ts(ts + O_COL, a_hi, v_lo, idesc_o, 1u);
            tc_mma_bf16_ts(ts + O_COL, a_hi, v_hi, idesc_o, 1u);
          }
        }
        tc_commit(bar(B_O + g));
      };
      int it = 0;
      if (first < p.n_items) {
        mbar_wait(bar(B_KV_FULL + 0), 0u, p.error_flag, 52);
        tc_fence_after();
        issue_qk(0, 0);
        issue_qk(1, 0);
      }
      for (int item = first; item < p.n_items; item += step, ++it) {
        const int bufi = it & 1;
        const bool has_next = item + step < p.n_items;
        for (int g = 0; g < 2; ++g) {
          mbar_wait(bar(B_P + g), (uint32_t)(it & 1), p.error_flag, 53);      // P written (and O of the previous item read)
          tc_fence_after();
          issue_pv(g, bufi);
          if (has_next) {
            if (g == 0) {
              mbar_wait(bar(B_KV_FULL + (bufi ^ 1)), (uint32_t)(((it + 1) >> 1) & 1), p.error_flag, 54);
              tc_fence_after();
            }
            issue_qk(g, bufi ^ 1);                                           // S of the next item: runs behind this P.V in the pipe
          }
        }
        tc_commit(bar(B_KV_FREE + bufi));
      }
    }
  } else {
    // ===== softmax warps: thread <-> query row =====
    const int g = warp >> 2, quad = warp & 3;
    const int chunk = chunk_of(n_chunks, g, quad);
    const bool active = chunk >= 0;
    const int r_lo = chunk * 32;
    const int i = r_lo + lane;
    const int tcols = tcols_of(g), tcols_pad = tpad_of(g);
    const int nblk = (tcols_pad + PB - 1) / PB;
    const int vis_last = min(r_lo + 31 + koff, tcols - 1);         // last column any row of the warp sees
    const int nb_warp = vis_last / PB + 1;                         // blocks this warp must evaluate
    const int n_full = (r_lo + koff + 1) / PB;                     // blocks [0, n_full) are visible to every row
    const int my_last = i + koff;                                  // last visible column of this row
    const uint32_t trow = tmem_base + (((uint32_t)(quad * 32)) << 16) + (uint32_t)g * 256u;
    int it = 0;
    for (int item = first; item < p.n_items; item += step, ++it) {
      const int b = item / H, h = item % H;
      const float* kb = reinterpret_cast<const float*>(smem + (it & 1) * BUF_BYTES + OFF_KB);
      mbar_wait(bar(B_S + g), (uint32_t)(it & 1), p.error_flag, 55);
      tc_fence_after();
      float sum = 0.f;
      if (active) {
        float mx = -INFINITY;
        for (int blk = 0; blk < nb_warp; ++blk) {
          uint32_t v[32];
          tc_ld32(trow + blk * PB, v);
          const float4* cw4 = reinterpret_cast<const float4*>(kb + blk * PB);
          tc_wait_ld();
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
          if (blk < n_full) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 w = cw4[j4];
              m0 = fmaxf(m0, __uint_as_float(v[4 * j4 + 0]) + w.x); m1 = fmaxf(m1, __uint_as_float(v[4 * j4 + 1]) + w.y);
              m2 = fmaxf(m2, __uint_as_float(v[4 * j4 + 2]) + w.z); m3 = fmaxf(m3, __uint_as_float(v[4 * j4 + 3]) + w.w);
            }
          } else {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 w = cw4[j4];
              const int c = blk * PB + 4 * j4;
              const bool firstc = pim && c == 0;                  // the objective column is visible to every row
              if (c + 0 <= my_last || firstc) m0 = fmaxf(m0, __uint_as_float(v[4 * j4 + 0]) + w.x);
              if (c + 1 <= my_last) m1 = fmaxf(m1, __uint_as_float(v[4 * j4 + 1]) + w.y);
              if (c + 2 <= my_last) m2 = fmaxf(m2, __uint_as_float(v[4 * j4 + 2]) + w.z);
              if (c + 3 <= my_last) m3 = fmaxf(m3, __uint_as_float(v[4 * j4 + 3]) + w.w);
            }
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        }
        for (int blk = 0; blk < nblk; ++blk) {
          uint32_t pk[32];
          if (blk < nb_warp) {
            uint32_t v[32];
            tc_ld32(trow + blk * PB, v);
            const float4* cw4 = reinterpret_cast<const float4*>(kb + blk * PB);
            tc_wait_ld();
            const bool full = blk < n_full;
#pragma unroll
            for (int s8 = 0; s8 < PB / 8; ++s8) {
              float x[8];
#pragma unroll
              for (int hq = 0; hq < 2; ++hq) {
                const float4 w = cw4[s8 * 2 + hq];
                const int c = blk * PB + s8 * 8 + hq * 4;
                const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  // fully masked row: -inf - -inf = NaN, as torch's softmax
                  float pv = exp2f(__uint_as_float(v[s8 * 8 + hq * 4 + e]) + ww[e] - mx);
                  if (!full && !(c + e <= my_last || (pim && c + e == 0))) pv = 0.f;
                  x[hq * 4 + e] = pv;
                  sum += pv;
                }
              }
              uint4 hi, lo;
              split8(x, hi, lo);
              pk[s8 * 4 + 0] = hi.x; pk[s8 * 4 + 1] = hi.y; pk[s8 * 4 + 2] = hi.z; pk[s8 * 4 + 3] = hi.w;
              pk[16 + s8 * 4 + 0] = lo.x; pk[16 + s8 * 4 + 1] = lo.y; pk[16 + s8 * 4 + 2] = lo.z; pk[16 + s8 * 4 + 3] = lo.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) pk[j] = 0u;
          }
          tc_st32(trow + blk * PB, pk);              // P over S, in place: [hi: 16 columns | lo: 16 columns]
        }
        tc_wait_st();
      }
      tc_fence_before();
      mbar_arrive(bar(B_P + g));
      mbar_wait(bar(B_O + g), (uint32_t)(it & 1), p.error_flag, 56);
      tc_fence_after();
      if (active) {
        uint32_t v[32];
        tc_ld32(trow + O_COL, v);
        tc_wait_ld();
        if (i < L) {
          const float inv = 1.0f / sum;
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]) * inv;
          float* dst = p.out + ((int64_t)b * L + i) * (H * DH) + h * DH;
#pragma unroll
          for (int q = 0; q < 4; ++q) stg256(dst + q * 8, &o[q * 8]);
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace tcp
}  // namespace irs

using namespace irs;

// Full-window fast path of irs_pim_attn_fwd_tc (dispatch in attention_tc.cu).
extern "C" int irs_pim_attn_persistent_supported(int L, int dh, int q_row0, int n_q) {
  return (dh == tcp::DH && L > 128 && L <= tcp::KEYS - 1 && q_row0 == 0 && n_q == L) ? 1 : 0;
}

int irs_pim_attn_persistent_launch(const float* q, const float* k, const float* v, int64_t ld_q, int64_t ld_k, int64_t ld_v,
                                   const int64_t* ids, const float* r_u, float w_h, float w_obj, int mode,
                                   float* out, int B, int L, int H, int* error_flag, cudaStream_t stream) {
  if ((ld_q & 7) || (ld_k & 7) || (ld_v & 7) || ((uintptr_t)q & 31) || ((uintptr_t)k & 31) || ((uintptr_t)out & 31)) return IRS_E_SHAPE;
  static bool configured = false;
  if (!configured) {
    IRS_CUDA(cudaFuncSetAttribute(tcp::pim_attn_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcp::SMEM_BYTES));
    configured = true;
  }
  tcp::Params p = {};
  p.q = q; p.k = k; p.v = v; p.ld_q = ld_q; p.ld_k = ld_k; p.ld_v = ld_v; p.ids = ids; p.r_u = r_u;
  p.w_h = w_h; p.w_obj = w_obj; p.mode = mode; p.out = out; p.B = B; p.L = L; p.H = H; p.n_items = B * H;
  p.error_flag = error_flag;
  const unsigned grid = (unsigned)(p.n_items < kNumSMs ? p.n_items : kNumSMs);
  tcp::pim_attn_persistent_kernel<<<grid, tcp::THREADS, tcp::SMEM_BYTES, stream>>>(p);
  IRS_LAUNCHED();
  return 0;
}
