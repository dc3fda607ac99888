// Fused (bias +) residual + LayerNorm, optionally chained with "+ const, LayerNorm".
// HBM-bound: one warp per row, the row lives in registers between the two normalisations so the
// norm1 -> (+cross-attention constant) -> norm2 pair of nn.TransformerDecoderLayer costs one read
// of x, one read of y and one write.
//   reference: model/influentialRS.py:67-74 (post-norm decoder layer), :172-173 (zero memory =>
//   the cross-attention block is the constant W_o b_v + b_o).
#include "common.cuh"

namespace irs {

template <int EPL>   // elements per lane; d <= 32*EPL
__device__ __forceinline__ void ln_inplace(float (&v)[EPL], int d, int lane, const float* __restrict__ g,
                                           const float* __restrict__ b, float eps) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < EPL; ++e) { const int c = lane + 32 * e; if (c < d) s += v[e]; }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int e = 0; e < EPL; ++e) { const int c = lane + 32 * e; if (c < d) { const float t = v[e] - mean; q += t * t; } }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int c = lane + 32 * e;
    if (c < d) v[e] = (v[e] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
  }
}

template <int EPL>
__global__ void __launch_bounds__(256)
residual_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ y_bias,
                          const float* __restrict__ g1, const float* __restrict__ b1,
                          const float* __restrict__ c2, const float* __restrict__ g2, const float* __restrict__ b2,
                          float eps, float* __restrict__ out, int64_t rows, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += n_warps) {
    float v[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int c = lane + 32 * e;
      float t = 0.f;
      if (c < d) {
        t = x[r * d + c];
        if (y) t += y[r * d + c];
        if (y_bias) t += __ldg(y_bias + c);
      }
      v[e] = t;
    }
    ln_inplace<EPL>(v, d, lane, g1, b1, eps);
    if (g2) {
#pragma unroll
      for (int e = 0; e < EPL; ++e) { const int c = lane + 32 * e; if (c < d && c2) v[e] += __ldg(c2 + c); }
      ln_inplace<EPL>(v, d, lane, g2, b2, eps);
    }
#pragma unroll
    for (int e = 0; e < EPL; ++e) { const int c = lane + 32 * e; if (c < d) out[r * d + c] = v[e]; }
  }
}

}  // namespace irs

extern "C" int irs_residual_layernorm(const float* x, const float* y, const float* y_bias,
                                      const float* g1, const float* b1, const float* c2, const float* g2, const float* b2,
                                      float eps, float* out, int64_t rows, int d, void* stream) {
  if (!x || !g1 || !b1 || !out) return IRS_E_BADARG;
  if (g2 && !b2) return IRS_E_BADARG;
  if (rows < 0 || d <= 0) return IRS_E_BADARG;
  if (d > 1024) return IRS_E_SHAPE;
  if (rows == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int64_t blocks = irs::ceil_div(rows, 8);
  const int64_t cap = (int64_t)irs::kNumSMs * 8 * 4;
  if (blocks > cap) blocks = cap;
#define IRS_LN_LAUNCH(E) irs::residual_layernorm_kernel<E><<<(unsigned)blocks, 256, 0, s>>>(x, y, y_bias, g1, b1, c2, g2, b2, eps, out, rows, d)
  if (d <= 32) IRS_LN_LAUNCH(1);
  else if (d <= 64) IRS_LN_LAUNCH(2);
  else if (d <= 128) IRS_LN_LAUNCH(4);
  else if (d <= 256) IRS_LN_LAUNCH(8);
  else if (d <= 512) IRS_LN_LAUNCH(16);
  else IRS_LN_LAUNCH(32);
#undef IRS_LN_LAUNCH
  IRS_LAUNCHED();
  return 0;
}
