// "Operand image" format of the packed q/k/v projection for the persistent attention kernel
// (full windows of 129..223 positions, 32-wide heads).
//
// For every (batch, head) item one contiguous block of IMG_ITEM_BYTES holds q, k and v already split
// into bf16 hi/lo halves and laid out exactly as the tcgen05 shared-memory operands the attention
// kernel needs, so that staging an item is ONE cp.async.bulk (no conversion, no register traffic):
//   Q  [part hi|lo][tile 2][slab 4][row 128][8 bf16]   K-major A operand of S = Q K^T; pre-scaled by
//                                                      log2(e)/sqrt(dh); row = slot of the token in the
//                                                      balanced chunk -> (tile, quadrant) map below
//   K  [part hi|lo][slab 4][col 224][8 bf16]           K-major B operand; col = key column (PIM: the
//                                                      objective key L-1 sits in column 0)
//   V  [part hi|lo][slab 4][col 224][8 bf16]           the same layout, consumed as an MN-major B
//                                                      operand of O = P V (no transpose needed)
// The producer (decoder_chain_tc.cu epilogue, or irs_qkv_to_images) writes 16-byte pieces; consecutive
// tokens are adjacent, so a warp stores 512 contiguous bytes.  Rows / columns that no token maps to
// (padding up to 128 / 224) are never written: the buffer must be zero-initialised once.
#pragma once
#include <stdint.h>

namespace irs {
namespace img {

constexpr int DH = 32, SLABS = 4, KEYS = 224, BM = 128;
constexpr uint32_t Q_LBO = BM * 16, K_LBO = KEYS * 16;
constexpr uint32_t Q_TILE = SLABS * Q_LBO;            // 8192
constexpr uint32_t Q_PART = 2 * Q_TILE;               // 16384
constexpr uint32_t K_PART = SLABS * K_LBO;            // 14336
constexpr uint32_t OFF_Q = 0, OFF_K = 2 * Q_PART, OFF_V = OFF_K + 2 * K_PART;
constexpr uint32_t ITEM_BYTES = OFF_V + 2 * K_PART;   // 90112

// 32-row chunk handled by softmax warp `quad` of group `g` (-1: none).  Group 0 takes the four longest
// chunks ordered [second longest, longest, third, fourth]; group 1 the remaining short ones on
// quadrants 0, 2, 3 -- per quadrant the visible key blocks then add up to about the same number.
__host__ __device__ __forceinline__ int chunk_of(int n_chunks, int g, int quad) {
  if (g == 0) return quad == 0 ? n_chunks - 2 : (quad == 1 ? n_chunks - 1 : (quad == 2 ? n_chunks - 3 : n_chunks - 4));
  const int c = quad == 0 ? 0 : (quad == 1 ? -1 : quad - 1);
  return (c >= 0 && c < n_chunks - 4) ? c : -1;
}
// inverse: token l -> tile * 128 + row
__host__ __device__ __forceinline__ int q_slot(int n_chunks, int l) {
  const int chunk = l >> 5;
  int g, quad;
  if (chunk >= n_chunks - 4) { g = 0; quad = chunk == n_chunks - 2 ? 0 : (chunk == n_chunks - 1 ? 1 : (chunk == n_chunks - 3 ? 2 : 3)); }
  else { g = 1; quad = chunk == 0 ? 0 : chunk + 1; }
  return g * BM + quad * 32 + (l & 31);
}
__host__ __device__ __forceinline__ int kv_col(bool pim, int L, int l) { return pim ? (l == L - 1 ? 0 : l + 1) : l; }

}  // namespace img
}  // namespace irs
