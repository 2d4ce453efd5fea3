// k-hop computational graph (frontier BFS over the COO edge list) and CSR construction.
//
// Replaces Data.comp_graph (data.py:281-361) and the PyG k_hop_subgraph it calls: per hop one
// coalesced pass over the E edges marks the sources of edges whose target is in the frontier
// (flow = source_to_target); the final pass keeps every edge with both ends in the node set, in
// original order, relabelled by the rank of the node id (subset is ascending, like torch.unique).
// Stream compaction / sorting use CUB device primitives (plumbing, not the hot path).
#include <cub/cub.cuh>

#include "common.cuh"

namespace xpgnn {

__global__ void khop_init_kernel(int8_t* hop, uint8_t* frontier, int64_t N, int64_t query) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < N) {
    hop[i] = (i == query) ? 0 : -1;
    frontier[i] = (i == query) ? 1 : 0;
  }
}

// one pass: next[src] = 1 for every edge whose dst is in cur.  PyG re-expands the whole previous
// layer (visited nodes included), so `next` is not filtered by `hop`.
__global__ void __launch_bounds__(256) khop_expand_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                                          int64_t N, const uint8_t* __restrict__ cur, uint8_t* __restrict__ next) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += stride) {
    const int64_t s = src[e], d = dst[e];
    if ((uint64_t)s >= (uint64_t)N || (uint64_t)d >= (uint64_t)N) continue;  // counted by khop_validate_kernel
    if (cur[d]) next[s] = 1;
  }
}

// endpoints outside [0, N): PyG raises an index error; here they are counted (counts[2]) and the edge is ignored
__global__ void __launch_bounds__(256) khop_validate_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                                            int64_t N, int64_t* __restrict__ n_bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int bad = 0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += stride)
    bad += ((uint64_t)src[e] >= (uint64_t)N) || ((uint64_t)dst[e] >= (uint64_t)N);
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(reinterpret_cast<unsigned long long*>(n_bad), (unsigned long long)bad);
}

__global__ void khop_mark_kernel(int8_t* hop, const uint8_t* next, int64_t N, int level) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < N && next[i] && hop[i] < 0) hop[i] = (int8_t)level;
}

__global__ void khop_nodeflag_kernel(const int8_t* hop, int32_t* flag, int64_t N) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < N) flag[i] = hop[i] >= 0;
}

__global__ void khop_relabel_kernel(const int8_t* hop, const int32_t* rank, int64_t N, int32_t* relabel, int64_t* subset) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < N) {
    if (hop[i] >= 0) {
      relabel[i] = rank[i];
      subset[rank[i]] = i;
    } else {
      relabel[i] = -1;
    }
  }
}

__global__ void __launch_bounds__(256) khop_edgeflag_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                                            int64_t N, const int8_t* __restrict__ hop, uint8_t* __restrict__ edge_mask,
                                                            int32_t* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += stride) {
    const int64_t s = src[e], d = dst[e];
    const bool ok = (uint64_t)s < (uint64_t)N && (uint64_t)d < (uint64_t)N;
    const int k = ok && (hop[s] >= 0) && (hop[d] >= 0);
    edge_mask[e] = (uint8_t)k;
    flag[e] = k;
  }
}

__global__ void __launch_bounds__(256) khop_edgewrite_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                                             const int32_t* __restrict__ flag, const int32_t* __restrict__ rank,
                                                             const int32_t* __restrict__ relabel, int64_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += stride)
    if (flag[e]) {
      out[rank[e]] = relabel[src[e]];
      out[E + rank[e]] = relabel[dst[e]];
    }
}

__global__ void khop_counts_kernel(const int32_t* nflag, const int32_t* nrank, int64_t N, const int32_t* eflag, const int32_t* erank,
                                   int64_t E, const int32_t* relabel, int64_t query, int64_t* sub_edge_index, int64_t* counts) {
  const int64_t ns = N > 0 ? nrank[N - 1] + nflag[N - 1] : 0;
  int64_t es = E > 0 ? erank[E - 1] + eflag[E - 1] : 0;
  if (es == 0) {  // data.py:337-339: leave at least the node self connected
    sub_edge_index[0] = relabel[query];
    sub_edge_index[E > 0 ? E : 1] = relabel[query];
    es = 1;
  }
  counts[0] = ns;
  counts[1] = es;
}

// ---------------------------------------------------------------- CSR
__global__ void __launch_bounds__(256) csr_key_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E, int drop,
                                                      int32_t sentinel, int32_t* __restrict__ key, int32_t* __restrict__ val) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += stride) {
    const int64_t s = src[e], d = dst[e];
    key[e] = (drop && s == d) ? sentinel : (int32_t)d;  // dropped edges sort behind every row
    val[e] = (int32_t)s;
  }
}

__global__ void __launch_bounds__(256) csr_rowptr_kernel(const int32_t* __restrict__ sorted_key, int64_t E, int32_t N, int32_t* __restrict__ rowptr,
                                                         int64_t* __restrict__ n_kept) {
  // rowptr[r] = first position whose key >= r  (keys sorted ascending; sentinel N sorts last)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= N; r += stride) {
    int64_t lo = 0, hi = E;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (sorted_key[mid] < r) lo = mid + 1; else hi = mid;
    }
    rowptr[r] = (int32_t)lo;
    if (r == N) n_kept[0] = lo;
  }
}

}  // namespace xpgnn

using namespace xpgnn;

extern "C" {

int xpgnn_khop_subgraph(const int64_t* edge_index, int64_t E, int64_t N, int64_t query, int32_t hops, int64_t* subset,
                        int32_t* relabel, int8_t* hop, uint8_t* edge_mask, int64_t* sub_edge_index, int64_t* counts,
                        void* stream) {
  XP_REQUIRE(N > 0 && query >= 0 && query < N, "query outside [0, N)");
  XP_REQUIRE(E >= 0 && hops >= 0 && hops < 127, "bad E / hops");
  XP_REQUIRE(N < (1ll << 31) - 1 && E < (1ll << 31) - 1, "N/E out of int32 range (the scans and ranks are 32-bit)");
  XP_REQUIRE(subset && relabel && hop && counts && sub_edge_index && (edge_mask || E == 0), "null output");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t* src = edge_index;
  const int64_t* dst = edge_index + E;
  Scratch fr(st), nflag(st), nrank(st), eflag(st), erank(st), tmp(st);
  XP_CHECK(fr.alloc(2 * N));
  XP_CHECK(nflag.alloc(sizeof(int32_t) * N));
  XP_CHECK(nrank.alloc(sizeof(int32_t) * N));
  XP_CHECK(eflag.alloc(sizeof(int32_t) * (E + 1)));
  XP_CHECK(erank.alloc(sizeof(int32_t) * (E + 1)));
  uint8_t* cur = fr.as<uint8_t>();
  uint8_t* nxt = cur + N;
  const int nb = (int)ceil_div(N, 256);
  const int eb = (int)std::min<int64_t>(std::max<int64_t>(ceil_div(E, 256), 1), kNumSMs * 16);
  XP_CHECK(cudaMemsetAsync(counts + 2, 0, sizeof(int64_t), st));
  if (E > 0) XP_LAUNCH(khop_validate_kernel, eb, 256, 0, st, src, dst, E, N, counts + 2);
  XP_LAUNCH(khop_init_kernel, nb, 256, 0, st, hop, cur, N, query);
  for (int h = 1; h <= hops; ++h) {
    XP_CHECK(cudaMemsetAsync(nxt, 0, N, st));
    if (E > 0) XP_LAUNCH(khop_expand_kernel, eb, 256, 0, st, src, dst, E, N, cur, nxt);
    XP_LAUNCH(khop_mark_kernel, nb, 256, 0, st, hop, nxt, N, h);
    std::swap(cur, nxt);
  }
  size_t tb = 0, tb2 = 0;
  XP_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, nflag.as<int32_t>(), nrank.as<int32_t>(), (int)N, st));
  if (E > 0) XP_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb2, eflag.as<int32_t>(), erank.as<int32_t>(), (int)E, st));
  XP_CHECK(tmp.alloc(std::max(tb, tb2)));
  tb = tb2 = std::max(tb, tb2);
  XP_LAUNCH(khop_nodeflag_kernel, nb, 256, 0, st, hop, nflag.as<int32_t>(), N);
  XP_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, nflag.as<int32_t>(), nrank.as<int32_t>(), (int)N, st));
  XP_LAUNCH(khop_relabel_kernel, nb, 256, 0, st, hop, nrank.as<int32_t>(), N, relabel, subset);
  if (E > 0) {
    XP_LAUNCH(khop_edgeflag_kernel, eb, 256, 0, st, src, dst, E, N, hop, edge_mask, eflag.as<int32_t>());
    XP_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, eflag.as<int32_t>(), erank.as<int32_t>(), (int)E, st));
    XP_LAUNCH(khop_edgewrite_kernel, eb, 256, 0, st, src, dst, E, eflag.as<int32_t>(), erank.as<int32_t>(), relabel,
              sub_edge_index);
  }
  XP_LAUNCH(khop_counts_kernel, 1, 1, 0, st, nflag.as<int32_t>(), nrank.as<int32_t>(), N, eflag.as<int32_t>(),
            erank.as<int32_t>(), E, relabel, query, sub_edge_index, counts);
  return 0;
}

int xpgnn_build_csr(const int64_t* src, const int64_t* dst, int64_t E, int64_t N, int32_t drop_self_loops, int32_t* rowptr,
                    int32_t* col, int64_t* n_kept, void* stream) {
  XP_REQUIRE(N > 0 && N < (1ll << 31) - 1 && E >= 0 && E < (1ll << 31) - 1, "N/E out of int32 range");
  XP_REQUIRE(rowptr && n_kept && (col || E == 0), "null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (E == 0) {
    XP_CHECK(cudaMemsetAsync(rowptr, 0, sizeof(int32_t) * (N + 1), st));
    XP_CHECK(cudaMemsetAsync(n_kept, 0, sizeof(int64_t), st));
    return 0;
  }
  Scratch key(st), key2(st), val(st), tmp(st);
  XP_CHECK(key.alloc(sizeof(int32_t) * E));
  XP_CHECK(key2.alloc(sizeof(int32_t) * E));
  XP_CHECK(val.alloc(sizeof(int32_t) * E));
  const int eb = (int)std::min<int64_t>(ceil_div(E, 256), kNumSMs * 16);
  XP_LAUNCH(csr_key_kernel, eb, 256, 0, st, src, dst, E, drop_self_loops, (int32_t)N, key.as<int32_t>(), val.as<int32_t>());
  int bits = 1;
  while ((1ll << bits) <= N) ++bits;  // keys in [0, N]
  size_t tb = 0;
  XP_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.as<int32_t>(), key2.as<int32_t>(), val.as<int32_t>(), col, (int)E, 0,
                                           bits, st));
  XP_CHECK(tmp.alloc(tb));
  XP_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.as<int32_t>(), key2.as<int32_t>(), val.as<int32_t>(), col, (int)E, 0,
                                           bits, st));  // LSD radix sort is stable: edge order kept inside a row
  const int rb = (int)std::min<int64_t>(ceil_div(N + 1, 256), kNumSMs * 16);
  XP_LAUNCH(csr_rowptr_kernel, rb, 256, 0, st, key2.as<int32_t>(), E, (int32_t)N, rowptr, n_kept);
  return 0;
}

}  // extern "C"
