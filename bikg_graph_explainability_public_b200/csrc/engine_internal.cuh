// Host-side helpers shared by the two forward drivers (engine.cu: tile path, compact.cu: compact path).
#pragma once
#include <nvtx3/nvToolsExt.h>

#include <vector>

#include "common.cuh"
#include "dense_args.cuh"

namespace xpgnn {

// Optional per-category kernel timing with CUDA events on the launching stream (bench.py roofline).
enum { PROF_SCALE = 0, PROF_SPMM_INVARIANT, PROF_SPMM_TILE, PROF_DENSE, PROF_HEAD, PROF_COMPACT, PROF_N };
struct Profile {
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[PROF_N];
  void reset() {
    for (auto& v : ev) {
      for (auto& e : v) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
      v.clear();
    }
  }
};
extern Profile g_prof;
// Every engine stage is also an NVTX range (nvtx3 is header-only: no extra library), so a timeline tool shows
// masked degree / compaction / layer-0 SpMM / layer >= 1 SpMM / dense transform / head per tile.
inline const char* prof_name(int cat) {
  static const char* const names[PROF_N] = {"xpgnn:masked_degree", "xpgnn:spmm_invariant_l0", "xpgnn:spmm_tile_l1", "xpgnn:dense",
                                            "xpgnn:head", "xpgnn:compaction"};
  return cat >= 0 && cat < PROF_N ? names[cat] : "xpgnn";
}
struct ProfScope {
  cudaStream_t st;
  cudaEvent_t stop = nullptr;
  ProfScope(int cat, cudaStream_t s) : st(s) {
    nvtxRangePushA(prof_name(cat));
    if (!g_prof.on) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
    g_prof.ev[cat].push_back({a, b});
    stop = b;
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, st);
    nvtxRangePop();
  }
};

// bump allocator over the caller's workspace (base == nullptr: size query)
struct Bump {
  char* base;
  int64_t off = 0, cap;
  Bump(void* p, int64_t c) : base((char*)p), cap(c) {}
  template <class T>
  T* take(int64_t n) {
    off = (off + 255) & ~255ll;
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * (int64_t)sizeof(T);
    return r;
  }
};

// precision: DENSE_SIMT exact fp32 FMA | DENSE_TC_TF32X3 fp32 via 3 TF32 MMAs | DENSE_TC_BF16
int launch_dense(const DenseArgs& d, cudaStream_t st, int precision);

// ---- compact path (compact.cu) ----
bool compact_eligible(const xpgnn_plan_t* p);
int64_t compact_workspace_bytes(const xpgnn_plan_t* p, int tile);
bool compact_hetero_eligible(const xpgnn_plan_t* p);
int64_t compact_hetero_workspace_bytes(const xpgnn_plan_t* p, int tile);
int forward_compact_hetero(const xpgnn_plan_t* p, const uint32_t* act, int32_t W, int32_t s0, int32_t n_s, float* y, void* workspace,
                           int64_t workspace_bytes, int64_t* stats, cudaStream_t st, int dense_prec);
int forward_compact(const xpgnn_plan_t* p, const uint32_t* act, int32_t W, int32_t s0, int32_t n_s, float* y, void* workspace,
                    int64_t workspace_bytes, int64_t* stats, cudaStream_t st, int dense_prec);

}  // namespace xpgnn
