// Arguments of the fused masked SpMM + tcgen05 dense kernel (fused.cu).
#pragma once
#include "common.cuh"

namespace xpgnn {

struct FusedArgs {
  const int32_t* rowptr;
  const int32_t* col;
  const uint32_t* ebits;     // per-edge activity word of the current coalition word
  const float* scale;        // [N][32]
  int kind;                  // XPGNN_CONV_GCN | XPGNN_CONV_SAGE_MEAN (without root weight)
  const float* in;           // H[s][u][K]
  int64_t in_s_stride;
  int ld_in, K;
  const float* w_image;      // pre-formatted TF32 hi/lo image of W (fused_build_w_image)
  const float* b;
  int n_out, n_pad;
  float* out;
  int64_t out_s_stride;
  int ld_out, act_fn;
  int row_lo, n_rows;        // destination rows [row_lo, row_lo + n_rows)
  int b0, n_bits;            // first bit of the tile inside the word, coalitions in the tile
  int SB, slot_major;        // slots per A tile (power of two <= 32; 0 = auto), tile order
  uint32_t w_off;            // byte offset of the W stages in dynamic shared memory
};

bool fused_eligible(int K, int n_out, int ld_in, int ld_out, int64_t in_s_stride, int64_t out_s_stride, const void* in,
                    const void* out);
int64_t fused_w_image_bytes(int K, int n_out);
int fused_build_w_image(const float* w, int n_out, int K, float* img, cudaStream_t st);
int launch_fused(FusedArgs a, cudaStream_t st);

}  // namespace xpgnn
