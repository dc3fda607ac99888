// K1 / K2: item-embedding gather (+ sqrt(d) scale + positional encoding) and its scatter-add
// backward.  HBM-bound, random 4*d-byte rows: 128-bit loads, one warp per d=128 row.
//   reference: model/influentialRS.py:174-175 (forward), :111,307 (backward via autograd),
//              model/layers.py:31-32 (pe slice).
#include "common.cuh"

namespace irs {

// ---------------------------------------------------------------------------------------------
// forward: out[r,:] = table[ids[r],:] * scale + pe[r % L,:]
// Mul and add are rounded separately (__fmul_rn/__fadd_rn forbid FMA contraction) so the result
// is bit-identical to torch's  emb * sqrt(d) + pe.
// ---------------------------------------------------------------------------------------------
template <int UNROLL>
__global__ void __launch_bounds__(256)
embed_gather_v4_kernel(const int64_t* __restrict__ ids, const float4* __restrict__ table,
                       const float4* __restrict__ pe, float scale, float4* __restrict__ out,
                       int64_t rows, int L, int dv /* d/4 */, int64_t table_rows) {
  const int64_t total = rows * dv;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; idx < total; idx += stride * UNROLL) {
    float4 e[UNROLL];
    int64_t r[UNROLL];
    int c[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {           // issue all gathers first (memory-level parallelism)
      const int64_t i = idx + (int64_t)u * stride;
      ok[u] = i < total;
      r[u] = ok[u] ? i / dv : 0;
      c[u] = ok[u] ? (int)(i - r[u] * dv) : 0;
      int64_t id = ok[u] ? ids[r[u]] : 0;
      if (id < 0 || id >= table_rows) id = 0;    // out-of-range ids read the PAD row (torch would raise)
      e[u] = ld_nc_f4(table + id * dv + c[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (!ok[u]) continue;
      float4 o;
      o.x = __fmul_rn(e[u].x, scale); o.y = __fmul_rn(e[u].y, scale);
      o.z = __fmul_rn(e[u].z, scale); o.w = __fmul_rn(e[u].w, scale);
      if (pe != nullptr) {
        const float4 p = __ldg(pe + (int64_t)(r[u] % L) * dv + c[u]);
        o.x = __fadd_rn(o.x, p.x); o.y = __fadd_rn(o.y, p.y);
        o.z = __fadd_rn(o.z, p.z); o.w = __fadd_rn(o.w, p.w);
      }
      st_na_f4(out + r[u] * dv + c[u], o);
    }
  }
}

__global__ void __launch_bounds__(256)
embed_gather_scalar_kernel(const int64_t* __restrict__ ids, const float* __restrict__ table,
                           const float* __restrict__ pe, float scale, float* __restrict__ out,
                           int64_t rows, int L, int d, int64_t table_rows) {
  const int64_t total = rows * d;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / d;
    const int c = (int)(i - r * d);
    int64_t id = ids[r];
    if (id < 0 || id >= table_rows) id = 0;
    float o = __fmul_rn(__ldg(table + id * d + c), scale);
    if (pe != nullptr) o = __fadd_rn(o, __ldg(pe + (int64_t)(r % L) * d + c));
    out[i] = o;
  }
}

// ---------------------------------------------------------------------------------------------
// backward: d_table[ids[r],:] += d_out[r,:] * scale, PAD rows skipped.
// Warp-aggregated: a warp owns 32 consecutive (b,l) rows; rows of the chunk that hit the same item
// are summed in registers by the lowest lane holding that id (match.any), so each distinct id of
// the chunk costs one vector atomic per 16 bytes instead of one per occurrence.  fp32 atomics
// reorder the sum across chunks: gradients agree to ~1e-6 relative, inside the 1e-3 contract.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_scatter_add_kernel(const int64_t* __restrict__ ids, const float* __restrict__ d_out, float scale,
                         float* __restrict__ d_table, int64_t rows, int d, int64_t table_rows, int64_t pad_id) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool vec = (d & 3) == 0;
  for (int64_t base = warp * 32; base < rows; base += n_warps * 32) {
    const int64_t my_row = base + lane;
    int64_t my_id = (my_row < rows) ? ids[my_row] : pad_id;
    if (my_id < 0 || my_id >= table_rows) my_id = pad_id;
    const unsigned same = __match_any_sync(0xffffffffu, my_id);
    const bool leader = (my_id != pad_id) && ((int)(__ffs(same) - 1) == lane);
    unsigned leaders = __ballot_sync(0xffffffffu, leader);
    while (leaders) {
      const int src = __ffs(leaders) - 1;
      leaders &= leaders - 1;
      const int64_t id = __shfl_sync(0xffffffffu, my_id, src);
      const unsigned members = __shfl_sync(0xffffffffu, same, src);
      if (vec) {
        for (int c = lane * 4; c < d; c += 128) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          unsigned m = members;
          while (m) {
            const int r = __ffs(m) - 1;
            m &= m - 1;
            const float4 g = ld_nc_f4(reinterpret_cast<const float4*>(d_out + (base + r) * d + c));
            acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
          }
          acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
          atomicAdd(reinterpret_cast<float4*>(d_table + id * d + c), acc);   // red.global.add.v4.f32
        }
      } else {
        for (int c = lane; c < d; c += 32) {
          float acc = 0.f;
          unsigned m = members;
          while (m) {
            const int r = __ffs(m) - 1;
            m &= m - 1;
            acc += d_out[(base + r) * d + c];
          }
          atomicAdd(d_table + id * d + c, acc * scale);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// a2: personalised impressionability factor r_u[b] = w . U[user[b],:] + c   (B dot products of length du <= 32:
// one warp per 32 users would waste lanes on du = 10, so a thread owns a user and reads its row once).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pif_kernel(const int64_t* __restrict__ users, const float* __restrict__ U, const float* __restrict__ w,
           const float* __restrict__ c, float* __restrict__ out, int B, int du, int64_t n_user) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int64_t u = users[b];
  if (u < 0 || u >= n_user) u = 0;
  const float* row = U + u * du;
  float acc = 0.f;
  for (int k = 0; k < du; ++k) acc = fmaf(__ldg(row + k), __ldg(w + k), acc);
  out[b] = acc + (c ? __ldg(c) : 0.f);
}

}  // namespace irs

extern "C" int irs_pif_fwd(const int64_t* users, const float* user_table, const float* w, const float* c, float* r_u,
                           int B, int du, int64_t n_user, void* stream) {
  if (B == 0) return 0;
  if (!users || !user_table || !w || !r_u) return IRS_E_BADARG;
  if (B < 0 || du <= 0 || n_user <= 0) return IRS_E_BADARG;
  irs::pif_kernel<<<(unsigned)irs::ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(users, user_table, w, c, r_u, B, du, n_user);
  IRS_LAUNCHED();
  return 0;
}

extern "C" int irs_embed_gather_fwd(const int64_t* ids, const float* table, const float* pe, float scale,
                                    float* out, int64_t rows, int L, int d, int64_t table_rows, void* stream) {
  if (rows == 0) return 0;               // empty batch: nothing to do (pointers may be null)
  if (!ids || !table || !out) return IRS_E_BADARG;
  if (rows < 0 || L <= 0 || d <= 0 || table_rows <= 0) return IRS_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  const bool aligned = ((uintptr_t)table % 16 == 0) && ((uintptr_t)out % 16 == 0) && (!pe || (uintptr_t)pe % 16 == 0);
  if ((d & 3) == 0 && aligned) {
    const int dv = d / 4;
    constexpr int U = 4;
    const int64_t total = rows * dv;
    int64_t blocks = irs::ceil_div(total, 256 * U);
    const int64_t cap = (int64_t)irs::kNumSMs * 8 * 4;        // 8 resident CTAs/SM, <= 4 waves
    if (blocks > cap) blocks = cap;
    irs::embed_gather_v4_kernel<U><<<(unsigned)blocks, 256, 0, s>>>(
        ids, (const float4*)table, (const float4*)pe, scale, (float4*)out, rows, L, dv, table_rows);
  } else {
    int64_t blocks = irs::ceil_div(rows * d, 256);
    const int64_t cap = (int64_t)irs::kNumSMs * 32;
    if (blocks > cap) blocks = cap;
    irs::embed_gather_scalar_kernel<<<(unsigned)blocks, 256, 0, s>>>(ids, table, pe, scale, out, rows, L, d, table_rows);
  }
  IRS_LAUNCHED();
  return 0;
}

extern "C" int irs_embed_scatter_add_bwd(const int64_t* ids, const float* d_out, float scale, float* d_table,
                                         int64_t rows, int d, int64_t table_rows, int64_t pad_id, void* stream) {
  if (rows == 0) return 0;
  if (!ids || !d_out || !d_table) return IRS_E_BADARG;
  if (rows < 0 || d <= 0 || table_rows <= 0) return IRS_E_BADARG;
  if ((d & 3) == 0 && (((uintptr_t)d_out % 16) || ((uintptr_t)d_table % 16))) return IRS_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  int64_t warps = irs::ceil_div(rows, 32);
  int64_t blocks = irs::ceil_div(warps, 8);
  const int64_t cap = (int64_t)irs::kNumSMs * 8 * 4;
  if (blocks > cap) blocks = cap;
  irs::embed_scatter_add_kernel<<<(unsigned)blocks, 256, 0, s>>>(ids, d_out, scale, d_table, rows, d, table_rows, pad_id);
  IRS_LAUNCHED();
  return 0;
}
