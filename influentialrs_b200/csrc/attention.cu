// K3 / K4: self-attention with the Personalized Impressionability Mask computed in-kernel
// (no [B*H,L,L] mask tensor ever exists).  fp32 CUDA-core version: K and V of one (batch, head)
// live in shared memory, a warp owns 4 query rows at a time, lanes split the keys for QK^T and the
// head dimension for PV.  Causality is exploited: row i only visits keys j <= i plus, in PIM mode,
// the objective column L-1 that every row sees.
//   reference: model/influentialRS.py:139-151 (PIM, keyword branch), :171 (key padding), :189-193
//   (decoder call), torch.nn.functional.multi_head_attention_forward (scores/softmax/PV);
//   model/uRS.py:47-61 (causal + padding), model/sas.py:168-177 (causal only).
#include "common.cuh"

namespace irs {

constexpr int kAttWarps = 8;
constexpr int kAttRows = 4;     // query rows a warp processes together

struct AttnSmem {
  int kstride;      // floats per K row (odd => conflict-free lane-per-key reads)
  int off_v, off_kb, off_warp, warp_floats, total_floats;
};

__host__ __device__ inline AttnSmem attn_smem_layout(int L, int dh) {
  AttnSmem s;
  s.kstride = dh | 1;
  s.off_v = (L * s.kstride + 3) & ~3;
  s.off_kb = s.off_v + ((L * dh + 3) & ~3);
  s.off_warp = s.off_kb + ((L + 3) & ~3);
  s.warp_floats = dh * kAttRows + L * kAttRows;      // q^T [dh][4] + p [L][4], both float4-aligned
  s.total_floats = s.off_warp + kAttWarps * s.warp_floats;
  return s;
}

// Attention-probability dropout (nn.MultiheadAttention(dropout=p) in train mode; model/influentialRS.py:67-74 builds the
// decoder layers with dropout=self.dropout): a counter-based keep mask, a pure function of (seed, batch*head, query, key),
// so the backward kernel regenerates exactly the mask of the forward.  Returns 0 or 1/(1-p).
__device__ __forceinline__ float drop_scale(unsigned long long seed, uint32_t thresh, float inv_keep, int bh, int i, int j, int L) {
  if (thresh == 0u) return 1.0f;
  unsigned long long x = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)(((long long)bh * L + i) * L + j) + 1ull);
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;   // murmur3 finaliser
  return ((uint32_t)(x >> 32) >= thresh) ? inv_keep : 0.0f;
}

__device__ __forceinline__ float mask_value(int mode, int i, int j, int L, float w_h, float obj) {
  if (mode == IRS_MASK_PIM) {
    if (j == L - 1) return obj;
    return (j <= i) ? w_h : -INFINITY;
  }
  return (j <= i) ? 0.f : -INFINITY;
}

__global__ void __launch_bounds__(kAttWarps * 32)
pim_attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                    int64_t ld_q, int64_t ld_k, int64_t ld_v, const int64_t* __restrict__ ids,
                    const float* __restrict__ r_u, float w_h, float w_obj, int mode,
                    float* __restrict__ out, float* __restrict__ lse,
                    int L, int H, int dh, int q_row0, int n_q, int rows_per_cta,
                    unsigned long long seed, uint32_t drop_thresh, float inv_keep) {
  extern __shared__ __align__(16) float smem[];
  const AttnSmem lay = attn_smem_layout(L, dh);
  float* Ks = smem;
  float* Vs = smem + lay.off_v;
  float* kb = smem + lay.off_kb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = smem + lay.off_warp + warp * lay.warp_floats;      // [dh][4]
  float* ps = qs + dh * kAttRows;                                // [L][4]

  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int q_begin = q_row0 + blockIdx.y * rows_per_cta;
  const int q_end = min(q_row0 + n_q, q_begin + rows_per_cta);
  if (q_begin >= q_end) return;
  const int d = H * dh;
  const float scale = 1.0f / sqrtf((float)dh);
  const float obj = (mode == IRS_MASK_PIM) ? w_obj * r_u[b] : 0.f;
  const bool pim = (mode == IRS_MASK_PIM);

  // Keys this CTA can touch: j <= q_end-1, plus the objective column.
  const int kmax = min(q_end - 1, L - 1);
  const float* kbase = k + (int64_t)b * L * ld_k + h * dh;
  const float* vbase = v + (int64_t)b * L * ld_v + h * dh;
  for (int idx = threadIdx.x; idx < L * dh; idx += blockDim.x) {
    const int j = idx / dh, c = idx - j * dh;
    if (j <= kmax || (pim && j == L - 1)) {
      Ks[j * lay.kstride + c] = kbase[(int64_t)j * ld_k + c];
      Vs[j * dh + c] = vbase[(int64_t)j * ld_v + c];
    }
  }
  for (int j = threadIdx.x; j < L; j += blockDim.x)
    kb[j] = (mode != IRS_MASK_CAUSAL && ids[(int64_t)b * L + j] == 0) ? -INFINITY : 0.f;
  __syncthreads();

  for (int i0 = q_begin + warp * kAttRows; i0 < q_end; i0 += kAttWarps * kAttRows) {
    const int nrows = min(kAttRows, q_end - i0);
    for (int idx = lane; idx < dh * kAttRows; idx += 32) {
      const int r = idx / dh, c = idx - r * dh;
      qs[c * kAttRows + r] = (r < nrows) ? q[((int64_t)b * L + i0 + r) * ld_q + h * dh + c] : 0.f;
    }
    __syncwarp();
    const int jmax = min(i0 + nrows - 1, L - 1);
    const int nslots = jmax + 1 + ((pim && jmax < L - 1) ? 1 : 0);
    float m[kAttRows];
#pragma unroll
    for (int r = 0; r < kAttRows; ++r) m[r] = -INFINITY;
    for (int t = lane; t < nslots; t += 32) {
      const int j = (t <= jmax) ? t : L - 1;
      const float* kr = Ks + j * lay.kstride;
      float acc[kAttRows] = {0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < dh; ++c) {
        const float kv = kr[c];
        const float4 q4 = *reinterpret_cast<const float4*>(qs + c * kAttRows);
        acc[0] = fmaf(q4.x, kv, acc[0]); acc[1] = fmaf(q4.y, kv, acc[1]);
        acc[2] = fmaf(q4.z, kv, acc[2]); acc[3] = fmaf(q4.w, kv, acc[3]);
      }
      float4 s4;
      float* sp = reinterpret_cast<float*>(&s4);
#pragma unroll
      for (int r = 0; r < kAttRows; ++r) {
        const float s = acc[r] * scale + (mask_value(mode, i0 + r, j, L, w_h, obj) + kb[j]);
        sp[r] = s;
        m[r] = fmaxf(m[r], s);
      }
      *reinterpret_cast<float4*>(ps + t * kAttRows) = s4;
    }
#pragma unroll
    for (int r = 0; r < kAttRows; ++r) m[r] = warp_max(m[r]);
    __syncwarp();
    float sum[kAttRows] = {0.f, 0.f, 0.f, 0.f};
    for (int t = lane; t < nslots; t += 32) {
      float4 s4 = *reinterpret_cast<const float4*>(ps + t * kAttRows);
      float* sp = reinterpret_cast<float*>(&s4);
#pragma unroll
      const int jd = (t <= jmax) ? t : L - 1;
      for (int r = 0; r < kAttRows; ++r) {
        const float p = expf(sp[r] - m[r]);    // fully masked row: -inf - -inf = NaN, as torch
        sum[r] += p;                           // the softmax normaliser is taken BEFORE dropout (F.dropout(softmax(s)))
        sp[r] = p * drop_scale(seed, drop_thresh, inv_keep, blockIdx.x, i0 + r, jd, L);
      }
      *reinterpret_cast<float4*>(ps + t * kAttRows) = s4;
    }
#pragma unroll
    for (int r = 0; r < kAttRows; ++r) sum[r] = warp_sum(sum[r]);
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float o[kAttRows] = {0.f, 0.f, 0.f, 0.f};
      for (int t = 0; t < nslots; ++t) {
        const int j = (t <= jmax) ? t : L - 1;
        const float vv = Vs[j * dh + c];
        const float4 p4 = *reinterpret_cast<const float4*>(ps + t * kAttRows);
        o[0] = fmaf(p4.x, vv, o[0]); o[1] = fmaf(p4.y, vv, o[1]);
        o[2] = fmaf(p4.z, vv, o[2]); o[3] = fmaf(p4.w, vv, o[3]);
      }
#pragma unroll
      for (int r = 0; r < kAttRows; ++r)
        if (r < nrows) out[((int64_t)b * n_q + (i0 + r - q_row0)) * d + h * dh + c] = o[r] / sum[r];
    }
    if (lse != nullptr && lane < nrows)
      lse[((int64_t)b * H + h) * n_q + (i0 + lane - q_row0)] = m[lane] + logf(sum[lane]);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// backward.  Phase A (row-wise, warp per 4 query rows): D_i = dO_i . O_i and
//   dQ_i = scale * sum_j dS_ij K_j with dS_ij = P_ij (dO_i . V_j - D_i), P recomputed from lse.
// Phase B (column-wise, warp per 4 keys): dV_j = sum_i P_ij dO_i,  dK_j = scale * sum_i dS_ij Q_i,
//   and the PIM column's  d r_u[b] += w_obj * sum_{h,i} dS[i, L-1]  (one atomic per (b,h)).
// Recomputing P in both phases keeps everything in shared memory / registers (no L x L buffer, no
// atomics on dK/dV).
// ---------------------------------------------------------------------------------------------
struct AttnBwdSmem {
  int stride;   // odd row stride for Q,K,V,dO
  int off_k, off_v, off_do, off_kb, off_lse, off_D, off_warp, warp_floats, total_floats;
};
__host__ __device__ inline AttnBwdSmem attn_bwd_smem_layout(int L, int dh) {
  AttnBwdSmem s;
  s.stride = dh | 1;
  const int mat = (L * s.stride + 3) & ~3;
  s.off_k = mat; s.off_v = 2 * mat; s.off_do = 3 * mat;
  s.off_kb = 4 * mat;
  s.off_lse = s.off_kb + ((L + 3) & ~3);
  s.off_D = s.off_lse + ((L + 3) & ~3);
  s.off_warp = s.off_D + ((L + 3) & ~3);
  s.warp_floats = dh * kAttRows + 2 * L * kAttRows;    // x^T [dh][4], p [L][4], ds [L][4]
  s.total_floats = s.off_warp + kAttWarps * s.warp_floats;
  return s;
}

__global__ void __launch_bounds__(kAttWarps * 32)
pim_attn_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                    int64_t ld_q, int64_t ld_k, int64_t ld_v, const int64_t* __restrict__ ids,
                    const float* __restrict__ r_u, float w_h, float w_obj, int mode,
                    const float* __restrict__ o, const float* __restrict__ lse_g, const float* __restrict__ d_o,
                    float* __restrict__ d_q, float* __restrict__ d_k, float* __restrict__ d_v,
                    float* __restrict__ d_r_u, int L, int H, int dh,
                    unsigned long long seed, uint32_t drop_thresh, float inv_keep) {
  extern __shared__ __align__(16) float smem[];
  const AttnBwdSmem lay = attn_bwd_smem_layout(L, dh);
  float* Qs = smem;
  float* Ks = smem + lay.off_k;
  float* Vs = smem + lay.off_v;
  float* dOs = smem + lay.off_do;
  float* kb = smem + lay.off_kb;
  float* lses = smem + lay.off_lse;
  float* Ds = smem + lay.off_D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xs = smem + lay.off_warp + warp * lay.warp_floats;   // [dh][4]
  float* ps = xs + dh * kAttRows;                             // [L][4]
  float* dss = ps + L * kAttRows;                             // [L][4]
  const int st = lay.stride;

  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int d = H * dh;
  const float scale = 1.0f / sqrtf((float)dh);
  const bool pim = (mode == IRS_MASK_PIM);
  const float obj = pim ? w_obj * r_u[b] : 0.f;

  for (int idx = threadIdx.x; idx < L * dh; idx += blockDim.x) {
    const int j = idx / dh, c = idx - j * dh;
    const int64_t row = (int64_t)b * L + j;
    Qs[j * st + c] = q[row * ld_q + h * dh + c];
    Ks[j * st + c] = k[row * ld_k + h * dh + c];
    Vs[j * st + c] = v[row * ld_v + h * dh + c];
    dOs[j * st + c] = d_o[row * d + h * dh + c];
  }
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    kb[j] = (mode != IRS_MASK_CAUSAL && ids[(int64_t)b * L + j] == 0) ? -INFINITY : 0.f;
    lses[j] = lse_g[((int64_t)b * H + h) * L + j];
  }
  __syncthreads();
  // D_i = dO_i . O_i
  for (int i = warp; i < L; i += kAttWarps) {
    float acc = 0.f;
    for (int c = lane; c < dh; c += 32) acc += dOs[i * st + c] * o[((int64_t)b * L + i) * d + h * dh + c];
    acc = warp_sum(acc);
    if (lane == 0) Ds[i] = acc;
  }
  __syncthreads();

  // ---- Phase A: dQ
  for (int i0 = warp * kAttRows; i0 < L; i0 += kAttWarps * kAttRows) {
    const int nrows = min(kAttRows, L - i0);
    const int jmax = min(i0 + nrows - 1, L - 1);
    const int nslots = jmax + 1 + ((pim && jmax < L - 1) ? 1 : 0);
    for (int t = lane; t < nslots; t += 32) {
      const int j = (t <= jmax) ? t : L - 1;
      float4 ds4;
      float* dsp = reinterpret_cast<float*>(&ds4);
#pragma unroll
      for (int r = 0; r < kAttRows; ++r) {
        const int i = min(i0 + r, L - 1);
        float s = 0.f, dp = 0.f;
        for (int c = 0; c < dh; ++c) {
          const float kv = Ks[j * st + c];
          s = fmaf(Qs[i * st + c], kv, s);
          dp = fmaf(dOs[i * st + c], Vs[j * st + c], dp);
        }
        s = s * scale + (mask_value(mode, i0 + r, j, L, w_h, obj) + kb[j]);
        const float p = (r < nrows) ? expf(s - lses[i]) : 0.f;
        // with dropout O = (P o M) V: dP = (dO V^T) o M and D_i = dO_i . O_i still equals sum_j P_ij dP_ij
        dsp[r] = p * (dp * drop_scale(seed, drop_thresh, inv_keep, blockIdx.x, i, j, L) - Ds[i]);
      }
      *reinterpret_cast<float4*>(dss + t * kAttRows) = ds4;
    }
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float a[kAttRows] = {0.f, 0.f, 0.f, 0.f};
      for (int t = 0; t < nslots; ++t) {
        const int j = (t <= jmax) ? t : L - 1;
        const float kv = Ks[j * st + c];
        const float4 d4 = *reinterpret_cast<const float4*>(dss + t * kAttRows);
        a[0] = fmaf(d4.x, kv, a[0]); a[1] = fmaf(d4.y, kv, a[1]);
        a[2] = fmaf(d4.z, kv, a[2]); a[3] = fmaf(d4.w, kv, a[3]);
      }
#pragma unroll
      for (int r = 0; r < kAttRows; ++r)
        if (r < nrows) d_q[((int64_t)b * L + i0 + r) * ld_q + h * dh + c] = a[r] * scale;
    }
    __syncwarp();
  }

  // ---- Phase B: dK, dV (warp per 4 keys; lanes over query rows i >= j, or all rows for the
  //      objective column in PIM mode)
  float dru = 0.f;
  for (int j0 = warp * kAttRows; j0 < L; j0 += kAttWarps * kAttRows) {
    const int ncols = min(kAttRows, L - j0);
    const bool has_obj = pim && (j0 + ncols - 1 == L - 1);
    const int ibeg = has_obj ? 0 : j0;
    for (int i = ibeg + lane; i < L; i += 32) {
      float4 p4, ds4;
      float* pp = reinterpret_cast<float*>(&p4);
      float* dsp = reinterpret_cast<float*>(&ds4);
#pragma unroll
      for (int r = 0; r < kAttRows; ++r) {
        const int j = min(j0 + r, L - 1);
        float s = 0.f, dp = 0.f;
        for (int c = 0; c < dh; ++c) {
          s = fmaf(Qs[i * st + c], Ks[j * st + c], s);
          dp = fmaf(dOs[i * st + c], Vs[j * st + c], dp);
        }
        s = s * scale + (mask_value(mode, i, j, L, w_h, obj) + kb[j]);
        const float p = (r < ncols) ? expf(s - lses[i]) : 0.f;
        const float dm = drop_scale(seed, drop_thresh, inv_keep, blockIdx.x, i, j, L);
        pp[r] = p * dm;                         // dV_j = sum_i (P o M)_ij dO_i
        dsp[r] = p * (dp * dm - Ds[i]);
        if (pim && j0 + r == L - 1 && r < ncols) dru += dsp[r];
      }
      *reinterpret_cast<float4*>(ps + (i - ibeg) * kAttRows) = p4;
      *reinterpret_cast<float4*>(dss + (i - ibeg) * kAttRows) = ds4;
    }
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float ak[kAttRows] = {0.f, 0.f, 0.f, 0.f}, av[kAttRows] = {0.f, 0.f, 0.f, 0.f};
      for (int i = ibeg; i < L; ++i) {
        const float qv = Qs[i * st + c], dov = dOs[i * st + c];
        const float4 p4 = *reinterpret_cast<const float4*>(ps + (i - ibeg) * kAttRows);
        const float4 d4 = *reinterpret_cast<const float4*>(dss + (i - ibeg) * kAttRows);
        ak[0] = fmaf(d4.x, qv, ak[0]); ak[1] = fmaf(d4.y, qv, ak[1]);
        ak[2] = fmaf(d4.z, qv, ak[2]); ak[3] = fmaf(d4.w, qv, ak[3]);
        av[0] = fmaf(p4.x, dov, av[0]); av[1] = fmaf(p4.y, dov, av[1]);
        av[2] = fmaf(p4.z, dov, av[2]); av[3] = fmaf(p4.w, dov, av[3]);
      }
#pragma unroll
      for (int r = 0; r < kAttRows; ++r)
        if (r < ncols) {
          const int64_t row = (int64_t)b * L + j0 + r;
          d_k[row * ld_k + h * dh + c] = ak[r] * scale;
          d_v[row * ld_v + h * dh + c] = av[r];
        }
    }
    __syncwarp();
  }
  if (pim && d_r_u != nullptr) {
    dru = warp_sum(dru);
    if (lane == 0 && dru != 0.f) atomicAdd(d_r_u + b, w_obj * dru);
  }
}

}  // namespace irs

static int attn_check(const float* q, const float* k, const float* v, const int64_t* ids, const float* r_u,
                      int mode, int B, int L, int H, int dh) {
  if (!q || !k || !v) return IRS_E_BADARG;
  if (B <= 0 || L <= 0 || H <= 0 || dh <= 0) return IRS_E_BADARG;
  if (mode < 0 || mode > 2) return IRS_E_BADARG;
  if (mode != IRS_MASK_CAUSAL && !ids) return IRS_E_BADARG;
  if (mode == IRS_MASK_PIM && !r_u) return IRS_E_BADARG;
  return 0;
}

extern "C" int irs_pim_attn_fwd(const float* q, const float* k, const float* v, int64_t ld_q, int64_t ld_k, int64_t ld_v,
                                const int64_t* ids, const float* r_u, float w_h, float w_obj, int mode,
                                float* out, float* lse, int B, int L, int H, int dh, int q_row0, int n_q,
                                float p_drop, unsigned long long seed, void* stream) {
  int rc = attn_check(q, k, v, ids, r_u, mode, B, L, H, dh);
  if (rc) return rc;
  if (!out || q_row0 < 0 || n_q <= 0 || q_row0 + n_q > L) return IRS_E_BADARG;
  if (!(p_drop >= 0.f) || p_drop >= 1.f) return IRS_E_BADARG;
  const uint32_t drop_thresh = p_drop > 0.f ? (uint32_t)((double)p_drop * 4294967296.0) : 0u;
  const float inv_keep = 1.0f / (1.0f - p_drop);
  const irs::AttnSmem lay = irs::attn_smem_layout(L, dh);
  const size_t bytes = (size_t)lay.total_floats * sizeof(float);
  if (bytes > 227 * 1024) return IRS_E_SHAPE;
  static size_t configured = 0;
  if (bytes > configured) {
    IRS_CUDA(cudaFuncSetAttribute(irs::pim_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    configured = bytes;
  }
  // Enough CTAs to fill the machine: split the query rows when B*H alone is too few.
  int chunks = 1;
  const int per_pass = irs::kAttWarps * irs::kAttRows;
  while ((int64_t)B * H * chunks < 2 * irs::kNumSMs && (n_q + chunks - 1) / chunks > per_pass) chunks *= 2;
  int rows_per_cta = (n_q + chunks - 1) / chunks;
  rows_per_cta = ((rows_per_cta + irs::kAttRows - 1) / irs::kAttRows) * irs::kAttRows;
  chunks = (n_q + rows_per_cta - 1) / rows_per_cta;
  dim3 grid((unsigned)(B * H), (unsigned)chunks);
  irs::pim_attn_fwd_kernel<<<grid, irs::kAttWarps * 32, bytes, (cudaStream_t)stream>>>(
      q, k, v, ld_q, ld_k, ld_v, ids, r_u, w_h, w_obj, mode, out, lse, L, H, dh, q_row0, n_q, rows_per_cta,
      seed, drop_thresh, inv_keep);
  IRS_LAUNCHED();
  return 0;
}

extern "C" int irs_pim_attn_bwd(const float* q, const float* k, const float* v, int64_t ld_q, int64_t ld_k, int64_t ld_v,
                                const int64_t* ids, const float* r_u, float w_h, float w_obj, int mode,
                                const float* out, const float* lse, const float* d_out,
                                float* d_q, float* d_k, float* d_v, float* d_r_u,
                                int B, int L, int H, int dh, float p_drop, unsigned long long seed, void* stream) {
  int rc = attn_check(q, k, v, ids, r_u, mode, B, L, H, dh);
  if (rc) return rc;
  if (!out || !lse || !d_out || !d_q || !d_k || !d_v) return IRS_E_BADARG;
  if (!(p_drop >= 0.f) || p_drop >= 1.f) return IRS_E_BADARG;
  const uint32_t drop_thresh = p_drop > 0.f ? (uint32_t)((double)p_drop * 4294967296.0) : 0u;
  const float inv_keep = 1.0f / (1.0f - p_drop);
  const irs::AttnBwdSmem lay = irs::attn_bwd_smem_layout(L, dh);
  const size_t bytes = (size_t)lay.total_floats * sizeof(float);
  if (bytes > 227 * 1024) return IRS_E_SHAPE;
  static size_t configured = 0;
  if (bytes > configured) {
    IRS_CUDA(cudaFuncSetAttribute(irs::pim_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    configured = bytes;
  }
  irs::pim_attn_bwd_kernel<<<(unsigned)(B * H), irs::kAttWarps * 32, bytes, (cudaStream_t)stream>>>(
      q, k, v, ld_q, ld_k, ld_v, ids, r_u, w_h, w_obj, mode, out, lse, d_out, d_q, d_k, d_v, d_r_u, L, H, dh,
      seed, drop_thresh, inv_keep);
  IRS_LAUNCHED();
  return 0;
}
