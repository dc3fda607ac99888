// Parameters shared by the fused-scorer tile engines (CUDA-core and tcgen05) and their epilogues.
#pragma once
#include "common.cuh"

namespace irs {

enum ScoreMode { MODE_MAX = 0, MODE_COLLECT = 1, MODE_LSE = 2, MODE_RANK = 3 };

struct ScoreParams {
  // problem
  const float* h; int64_t ld_h;       // [M, d], rows ld_h apart
  const float* W;                     // [N, d]
  const float* bias;                  // [N] or null
  int M; int64_t N; int d;
  int64_t item_base;
  // exclusion lists (sorted columns), or null
  const int32_t* excl_sorted; const int32_t* excl_count; int Lx;
  // tiling (filled by the launcher)
  int m_tiles; int64_t n_tiles; int64_t tiles_per_split; int n_splits; int max_splits; int vec_ok;
  // MODE_MAX
  unsigned long long* slice_keys; int n_slices; int slices_per_split;
  // MODE_COLLECT
  const unsigned long long* thr_keys; unsigned long long* cand_keys; int* cand_count; int cand_cap;
  // MODE_LSE
  float* part_max; float* part_sum; const int64_t* sel; int n_sel; float* sel_logit;
  // MODE_RANK
  const int64_t* label; const float* label_score; int* rank_count; int* rank_excluded;
};

int launch_score_simt(int mode, ScoreParams& p, cudaStream_t s);
int score_simt_max_splits(int M, int64_t N);

}  // namespace irs
