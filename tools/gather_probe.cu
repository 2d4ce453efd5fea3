// Probe: bandwidth of random row gathers (the masked SpMM access pattern) as a function of the table size.
// Rows of CW floats are gathered by groups of CW/4 lanes (float4 per lane), DEG gathers summed per output row.
// Prints L2->SM GB/s (gathered bytes / time) per table size; decides the feature-chunk width of the engine.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

template <int CW>
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ tab, const int32_t* __restrict__ col, int deg, int n_rows,
                                                     float* __restrict__ out) {
  constexpr int G = CW / 4;          // lanes per row
  constexpr int RPW = 32 / G;        // rows per warp
  const int lane = threadIdx.x & 31, sub = lane % G, grp = lane / G;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w * RPW < n_rows; w += warps) {
    const int i = w * RPW + grp;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n_rows) {
      const int32_t* c = col + (int64_t)i * deg;
      for (int base = 0; base < deg; base += G) {
        const int my = base + sub < deg ? __ldg(c + base + sub) : -1;
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const int u = __shfl_sync(0xffffffffu, my, grp * G + j);
          if (u >= 0) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(tab + (int64_t)u * CW + sub * 4));
            acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
          }
        }
      }
      *reinterpret_cast<float4*>(out + (int64_t)i * CW + sub * 4) = acc;
    }
  }
}

template <int CW>
static void run(size_t table_mb, int n_rows, int deg, int grid_mult) {
  const int64_t tab_rows = (int64_t)table_mb * 1024 * 1024 / (CW * 4);
  float *tab, *out;
  int32_t* col;
  cudaMalloc(&tab, tab_rows * CW * 4);
  cudaMalloc(&out, (int64_t)n_rows * CW * 4);
  cudaMalloc(&col, (int64_t)n_rows * deg * 4);
  cudaMemset(tab, 0, tab_rows * CW * 4);
  std::vector<int32_t> h((size_t)n_rows * deg);
  uint64_t s = 88172645463325252ull;
  for (auto& x : h) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    x = (int32_t)(s % (uint64_t)tab_rows);
  }
  cudaMemcpy(col, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int it = 0; it < 6; ++it) {
    cudaEventRecord(a);
    gather_kernel<CW><<<148 * grid_mult, 256>>>(tab, col, deg, n_rows, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it > 0 && ms < best) best = ms;
  }
  const double gathered = (double)n_rows * deg * CW * 4;
  const double dram = (double)n_rows * CW * 4 + (double)n_rows * deg * 4;
  printf("CW=%d table=%4zu MB rows=%d deg=%d grid=148x%d : %.3f ms  gather %.0f GB/s  (+%.0f GB/s streaming)  err=%s\n", CW, table_mb, n_rows, deg,
         grid_mult, best, gathered / best / 1e6, dram / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(tab); cudaFree(out); cudaFree(col);
}

int main() {
  const int n_rows = 500000, deg = 10;  // one coalition of C3: 0.5 M active rows, 5 M active edges
  for (size_t mb : {8, 16, 32, 48, 64, 96, 128, 256, 1024}) run<32>(mb, n_rows, deg, 8);
  for (size_t mb : {8, 16, 32, 48, 64, 96, 128, 256, 1024}) run<16>(mb, n_rows, deg, 8);
  for (size_t mb : {16, 32, 64, 128}) run<64>(mb, n_rows, deg, 8);
  for (int gm : {2, 4, 16}) run<32>(64, n_rows, deg, gm);
  for (int gm : {2, 4, 16}) run<16>(32, n_rows, deg, gm);
  // bigger launch (several coalition passes back to back in one grid) to amortise ramp/tail
  run<32>(64, 4000000, deg, 8);
  run<16>(32, 4000000, deg, 8);
  return 0;
}
