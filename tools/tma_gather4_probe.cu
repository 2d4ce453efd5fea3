// Probe: rate of 128-byte row gathers through the TMA unit, Blackwell's tile::gather4 form (one instruction = 4 rows of a
// 2D tensor map, row indices in registers) against plain bulk copies (cp.async.bulk, one instruction per row).
// Background (profiles/r02_summary.md): the warp-specialised SpMM measured ~11-13 cycles per 128-byte UBLKCP per SM and
// 16 bytes per cycle per SM for LDGSTS, against > 40 bytes per cycle per SM for LDG.128 -- if gather4 costs one TMA
// instruction per FOUR rows it could feed a shared-memory ring at the fabric rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4_probe tma_gather4_probe.cu
//   ./tma_gather4_probe <box_rows: 1 | 4> <mode: 0 gather4 | 1 bulk copies> [producer warps]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

constexpr int kStages = 8;           // ring stages of 32 lanes x 512 bytes
constexpr int kStageBytes = 32 * 512;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
    if (spin > (1u << 24)) __trap();
  }
}

struct Smem {
  alignas(128) char ring[kStages][kStageBytes];
  uint64_t full[kStages], empty[kStages];
};

// every producer warp owns the stages g = pw (mod P); each lane moves 4 rows (512 bytes) per stage
template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe_kernel(const __grid_constant__ CUtensorMap tmap, const float* tab, uint32_t n_rows, int stages_per_cta, int P,
                                                        int C, float* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  Smem& S = *reinterpret_cast<Smem*>(raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], C); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp < P) {
    for (int g = warp; g < stages_per_cta; g += P) {
      const int slot = g % kStages, use = g / kStages;
      if (use > 0) mbar_wait(&S.empty[slot], (use - 1) & 1);
      if (lane == 0) mbar_expect(&S.full[slot], kStageBytes);
      __syncwarp();
      uint32_t h = (uint32_t)(blockIdx.x * 7919 + g) * 2654435761u + (uint32_t)lane * 40503u;
      int r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { h = h * 1664525u + 1013904223u; r[j] = (int)((h >> 4) % n_rows); }
      const uint32_t dst = s32(&S.ring[slot][lane * 512]);
      if (MODE == 0) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
                     "l"(&tmap), "r"(s32(&S.full[slot])), "r"(0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                     : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + j * 128),
                       "l"(tab + (size_t)r[j] * 32), "r"(128), "r"(s32(&S.full[slot]))
                       : "memory");
      }
    }
  } else if (warp < P + C) {
    // consumers: every consumer warp sees every stage (checks the data of its share, then releases the stage)
    const int cw = warp - P;
    float acc = 0.f;
    for (int g = 0; g < stages_per_cta; ++g) {
      const int slot = g % kStages, use = g / kStages;
      mbar_wait(&S.full[slot], use & 1);
      if ((g % C) == cw) {
        // re-derive the row ids of this lane's 4 rows and check column 0 / 31 of each (table row i holds i in every column)
        uint32_t h = (uint32_t)(blockIdx.x * 7919 + g) * 2654435761u + (uint32_t)lane * 40503u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          h = h * 1664525u + 1013904223u;
          const float want = (float)((h >> 4) % n_rows);
          const float* p = reinterpret_cast<const float*>(&S.ring[slot][lane * 512 + j * 128]);
          acc += fabsf(p[0] - want) + fabsf(p[31] - want);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[slot]);
    }
    if (acc != 0.f) atomicAdd(out, acc);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  const int mode = argc > 2 ? atoi(argv[2]) : 0;
  const int P = argc > 3 ? atoi(argv[3]) : 4;
  const int C = 4;
  const size_t table_mb = argc > 4 ? atoi(argv[4]) : 64;
  const uint32_t n_rows = (uint32_t)(table_mb * 1024 * 1024 / 128);
  float *tab, *out;
  cudaMalloc(&tab, (size_t)n_rows * 128);
  cudaMalloc(&out, 4);
  cudaMemset(out, 0, 4);
  std::vector<float> h((size_t)n_rows * 32);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i / 32);
  cudaMemcpy(tab, h.data(), h.size() * 4, cudaMemcpyHostToDevice);

  CUtensorMap tmap;
  memset(&tmap, 0, sizeof tmap);
  if (mode == 0) {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    const cuuint64_t dims[2] = {32, n_rows};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, tab, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed: %d (box rows %d)\n", (int)r, box_rows); return 3; }
  }
  const int stages_per_cta = 4096;  // 4096 stages x 16 KB = 64 MB gathered per CTA
  void (*k)(const CUtensorMap, const float*, uint32_t, int, int, int, float*) = mode == 0 ? probe_kernel<0> : probe_kernel<1>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int it = 0; it < 4; ++it) {
    cudaEventRecord(a);
    k<<<148, 32 * (P + C), sizeof(Smem)>>>(tmap, tab, n_rows, stages_per_cta, P, C, out);
    cudaEventRecord(b);
    const cudaError_t e = cudaEventSynchronize(b);
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 4; }
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it > 0 && ms < best) best = ms;
  }
  float err = -1.f;
  cudaMemcpy(&err, out, 4, cudaMemcpyDeviceToHost);
  const double bytes = 148.0 * stages_per_cta * kStageBytes;
  printf("mode=%s box_rows=%d producers=%d table=%zu MB: %.3f ms  %.0f GB/s  = %.1f B/clk/SM @1.9GHz, %.1f cycles per 128-byte row per SM; data error sum %.1f  (%s)\n",
         mode == 0 ? "gather4" : "bulk", box_rows, P, table_mb, best, bytes / best / 1e6, bytes / best / 1e6 / 148 / 1.9, 128.0 / (bytes / best / 1e6 / 148 / 1.9), err,
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
