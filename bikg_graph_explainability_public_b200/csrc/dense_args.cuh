// Arguments of the dense row transform, shared by the SIMT (engine.cu) and tensor-core (dense_tc.cu) paths.
#pragma once
#include "common.cuh"

namespace xpgnn {

// out[m][n] (+)= act(sum_k in[m][k] w[n][k] + b[n]); m enumerates (coalition slot s, row) pairs:
// optional row list, per-slot strides.
struct DenseArgs {
  const float* in;
  int64_t in_s_stride;
  int ld_in, k;
  const float* w;  // [n_out][k]
  const float* b;
  int n_out;
  float* out;
  int64_t out_s_stride;
  int ld_out;
  const int32_t* rows;
  int rows_per_s, row_lo;  // rows == NULL: row = row_lo + (m % rows_per_s)
  int64_t M;               // n_slots * rows_per_s
  int accumulate, act_fn;
  int dst_lo, dst_hi;      // rows outside [dst_lo, dst_hi) are skipped (row lists of hetero layers)
};

__device__ __forceinline__ float apply_act(float x, int a) {
  if (a == XPGNN_ACT_RELU) return fmaxf(x, 0.0f);
  if (a == XPGNN_ACT_SIGMOID) return 1.0f / (1.0f + expf(-x));
  return x;
}

enum { DENSE_SIMT = 0, DENSE_TC_BF16 = 1, DENSE_TC_TF32X3 = 2 };
bool dense_tc_eligible(const DenseArgs& d, int mode /*0 tf32x3, 1 bf16*/);
int launch_dense_tc(const DenseArgs& d, int mode, cudaStream_t st);

}  // namespace xpgnn
