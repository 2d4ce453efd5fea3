// Compact path of the coalition-batched masked message passing (homogeneous GCN / SAGE stacks, full mode).
//
// Same arithmetic as engine.cu (which replaces data.py:390-648 + model.py:62-328 of the reference), laid
// out for the memory system of a B200:
//
//  1. A node that is inactive in coalition s has no active in-edge (edge (u->v) needs both endpoint bits),
//     so its activation in every layer is the coalition-invariant "isolated node" value and no active
//     edge ever gathers it.  Per coalition only the rows of ACTIVE nodes are computed and stored; the
//     head evaluates the isolated chain itself when the query node is inactive.
//  2. Per tile of <= 32 coalitions the active in-edges of every coalition are compacted once into a
//     per-coalition CSR over its active rows (degree pass -> one CUB scan of packed (node, edge) counts ->
//     list pass), shared by all conv layers.
//  3. Activations are chunk-major, [chunk][node][CW] with CW = 32 floats (128 B): one (coalition, chunk)
//     pass of the SpMM touches 64 MB at C3 (0.5 M active rows x 128 B) instead of 16 GB for a whole
//     tile, so the ~10x re-gathers of every source row are served by the 126 MB L2 (measured with
//     tools/gather_probe.cu: 11.3 TB/s from a 64 MB set vs 4.6 TB/s from 1 GB).
//  4. GCN operands of layers >= 1 are stored pre-multiplied by deg^-1/2, so those gathers are
//     unweighted sums; layer 0 gathers the coalition-invariant Z = X W^T with a per-source weight.
#include <cub/device/device_scan.cuh>
#include <cuda_bf16.h>

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <string>
#include <vector>

#include "compact_internal.cuh"

namespace xpgnn {

constexpr int kKeyShift = 36;  // packed scan key: active-node count << 36 | active in-edge count
constexpr unsigned long long kEdgeMask = (1ull << kKeyShift) - 1ull;
constexpr int kMaxConvIso = 8;
constexpr int kMaxHeadC = 8;

// ------------------------------------------------------------------------------------------
// pass 1: edge-activity words of the coalition word + packed per-(coalition, node) counts
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) compact_degree_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                             const uint32_t* __restrict__ act, int W, int w, int b0, int nb, int n_rows,
                                                             uint32_t* __restrict__ ebits, unsigned long long* __restrict__ keys,
                                                             float* __restrict__ scale, int kind, int32_t* __restrict__ counter,
                                                             int long_threshold, int row_lo) {
  // rows = the destination range [row_lo, row_lo + n_rows) of the relation; keys are indexed by the row's position in it
  const int lane = threadIdx.x & 31;
  for (int ib = grab_rows(counter, lane); ib < n_rows; ib = grab_rows(counter, lane))
  for (int idx = ib; idx < min(n_rows, ib + kRowGrab); ++idx) {
    const int v = row_lo + idx;
    const int e0 = rowptr[v], e1 = rowptr[v + 1];
    if (long_threshold > 0 && e1 - e0 > long_threshold) continue;  // hub row: compact_degree_long_kernel
    const uint32_t av = act[(int64_t)v * W + w];
    int cnt = 0;
    for (int base = e0; base < e1; base += 32) {
      const int e = base + lane;
      uint32_t bits = 0;
      if (e < e1) {
        bits = act[(int64_t)col[e] * W + w] & av;
        ebits[e] = bits;
      }
      const int n = min(32, e1 - base);
      for (int j = 0; j < n; ++j) cnt += (__shfl_sync(0xffffffffu, bits, j) >> lane) & 1u;
    }
    // [N][32] by bit of the word: GCN (1 + masked in-degree)^-1/2, SAGE 1 / max(1, masked in-degree)
    if (scale) scale[(int64_t)v * 32 + lane] = kind == XPGNN_CONV_GCN ? gcn_dinv((uint32_t)cnt) : 1.0f / (float)max(cnt, 1);
    const int t = lane - b0;  // lane = bit of the word, t = slot of the tile
    if (t >= 0 && t < nb) keys[(int64_t)t * n_rows + idx] = ((av >> lane) & 1u) ? ((1ull << kKeyShift) | (unsigned long long)cnt) : 0ull;
  }
}

// hub rows of pass 1: one CTA per row, the warps take the 32-edge batches round-robin, counts summed through shared memory
__global__ void __launch_bounds__(256) compact_degree_long_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                  const uint32_t* __restrict__ act, int W, int w, int b0, int nb, int n_rows,
                                                                  uint32_t* __restrict__ ebits, unsigned long long* __restrict__ keys,
                                                                  float* __restrict__ scale, int kind, const int32_t* __restrict__ long_rows,
                                                                  int row_lo) {
  __shared__ int s_cnt[8][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int v = long_rows[blockIdx.x];
  const int e0 = rowptr[v], e1 = rowptr[v + 1];
  const uint32_t av = act[(int64_t)v * W + w];
  int cnt = 0;
  for (int base = e0 + 32 * wib; base < e1; base += 256) {
    const int e = base + lane;
    uint32_t bits = 0;
    if (e < e1) {
      bits = act[(int64_t)col[e] * W + w] & av;
      ebits[e] = bits;
    }
    const int n = min(32, e1 - base);
    for (int j = 0; j < n; ++j) cnt += (__shfl_sync(0xffffffffu, bits, j) >> lane) & 1u;
  }
  s_cnt[wib][lane] = cnt;
  __syncthreads();
  if (wib != 0) return;
  for (int w8 = 1; w8 < 8; ++w8) cnt += s_cnt[w8][lane];
  if (scale) scale[(int64_t)v * 32 + lane] = kind == XPGNN_CONV_GCN ? gcn_dinv((uint32_t)cnt) : 1.0f / (float)max(cnt, 1);
  const int t = lane - b0;
  if (t >= 0 && t < nb) keys[(int64_t)t * n_rows + (v - row_lo)] = ((av >> lane) & 1u) ? ((1ull << kKeyShift) | (unsigned long long)cnt) : 0ull;
}

// word `w` of every node's coalition bits as a dense array: the kernels of a tile gather act[u] once per edge, and with W = 128 words
// per node (the 4096-coalition job) every such gather touched its own 32-byte sector of a 512 MB matrix; the column is 4 MB
// (r02: masked degree 1.02 -> per-tile cost of the W = 4 case)
__global__ void __launch_bounds__(256) extract_word_kernel(const uint32_t* __restrict__ act, int W, int w, int N, uint32_t* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) out[v] = act[(int64_t)v * W + w];
}

// destination rows with more than `threshold` in-edges (a property of the graph: found once per forward call)
__global__ void __launch_bounds__(256) find_long_rows_kernel(const int32_t* __restrict__ rowptr, int N, int threshold,
                                                             int32_t* __restrict__ list, int32_t* __restrict__ count) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N && rowptr[v + 1] - rowptr[v] > threshold) list[atomicAdd(count, 1)] = v;
}

// pass 2 (after the exclusive scan of the keys): per-coalition list of active rows, compact in-edge
// offsets, per-source GCN weight of layer 0, per-slot totals
__global__ void __launch_bounds__(256) compact_finalize_kernel(const unsigned long long* __restrict__ keys,
                                                               const unsigned long long* __restrict__ scanned, int N, int row_lo,
                                                               int32_t* __restrict__ act_list, uint32_t* __restrict__ rowptr_c,
                                                               float* __restrict__ wgt, int2* __restrict__ slot_info,
                                                               long long* __restrict__ slot_base, int32_t* __restrict__ rows_packed,
                                                               float* __restrict__ rs_packed, int long_cnt, int32_t* __restrict__ long_list,
                                                               int32_t* __restrict__ n_long_list, int long_cap) {
  const int t = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  // first 128-row tile of this slot in the concatenated, per-slot padded tile table
  __shared__ int s_tile0;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int q = 0; q < t; ++q)
      acc += ((int)((scanned[(int64_t)(q + 1) * N] >> kKeyShift) - (scanned[(int64_t)q * N] >> kKeyShift)) + 127) / 128;
    s_tile0 = acc;
  }
  __syncthreads();
  if (v >= N) return;
  const int64_t g = (int64_t)t * N + v;
  const unsigned long long K = keys[g], S = scanned[g], B = scanned[(int64_t)t * N];
  if (K >> kKeyShift) {
    const int64_t i = (int64_t)((S >> kKeyShift) - (B >> kKeyShift));
    act_list[(int64_t)t * N + i] = row_lo + v;  // N = rows of the destination range, v = position in it
    rowptr_c[(int64_t)t * (N + 1) + i] = (uint32_t)((S & kEdgeMask) - (B & kEdgeMask));
    rows_packed[(int64_t)s_tile0 * 128 + i] = (int32_t)(((uint32_t)t << kPackShift) | (uint32_t)(row_lo + v));
    rs_packed[(int64_t)s_tile0 * 128 + i] = gcn_dinv((uint32_t)(K & kEdgeMask));
    if (long_cnt > 0 && (K & kEdgeMask) > (unsigned long long)long_cnt)  // hub row of this coalition: cspmm_long_kernel
      long_list[(int64_t)t * long_cap + atomicAdd(n_long_list + t, 1)] = i;  // per-slot segment: the long-row kernel works pass by pass
  }
  if (wgt) wgt[g] = gcn_dinv((uint32_t)(K & kEdgeMask));
  if (v == N - 1) {
    const unsigned long long T = S + K;
    const int n_act = (int)((T >> kKeyShift) - (B >> kKeyShift));
    const unsigned long long tot = (T & kEdgeMask) - (B & kEdgeMask);
    rowptr_c[(int64_t)t * (N + 1) + n_act] = (uint32_t)tot;
    slot_info[t] = make_int2(n_act, (int)min(tot, (unsigned long long)INT_MAX));
    slot_base[t] = (long long)(B & kEdgeMask);
    for (int i = n_act; i < (n_act + 127) / 128 * 128; ++i) rows_packed[(int64_t)s_tile0 * 128 + i] = -1;  // pad the last tile
  }
}

// first 128-row tile of every slot's active-row list, tile total, work counters, stats
__global__ void __launch_bounds__(256) compact_tilemap_kernel(const int2* __restrict__ slot_info, int nb, int32_t* __restrict__ slot_tile_start,
                                                              int32_t* __restrict__ n_tiles, int n_layers,
                                                              int64_t* stats, int32_t* __restrict__ counters, int count_tile) {
  __shared__ int start[33];
  if (threadIdx.x < 32) counters[threadIdx.x] = 0;  // work counters of the SpMM launches of this tile (16 + l: long-row kernel of layer l)
  if (threadIdx.x == 0) {
    int acc = 0;
    long long active = 0;
    for (int t = 0; t < nb; ++t) {
      start[t] = acc;
      acc += (slot_info[t].x + 127) / 128;
      active += slot_info[t].y;
    }
    start[nb] = acc;
    *n_tiles = acc;
    if (stats) {
      stats[1] += active * n_layers;  // active edge visits of this CSR (one per conv layer that uses it)
      stats[3] += count_tile;
    }
  }
  __syncthreads();
  if (threadIdx.x <= nb) slot_tile_start[threadIdx.x] = start[threadIdx.x];
}

// pass 3: per-coalition compacted source lists (original node ids, CSR order kept)
__global__ void __launch_bounds__(256) compact_edges_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                            const uint32_t* __restrict__ ebits, const uint32_t* __restrict__ act, int W,
                                                            int w, int b0, int nb, int N, const unsigned long long* __restrict__ scanned,
                                                            int32_t* __restrict__ ccol, int32_t* __restrict__ counter, int long_threshold,
                                                            int row_lo) {
  // N = rows of the destination range [row_lo, row_lo + N)
  const int lane = threadIdx.x & 31;
  const uint32_t tile_mask = nb == 32 ? 0xffffffffu : ((1u << nb) - 1u);
  const uint32_t lt = (1u << lane) - 1u;
  for (int ib = grab_rows(counter, lane); ib < N; ib = grab_rows(counter, lane))
  for (int idx = ib; idx < min(N, ib + kRowGrab); ++idx) {
    const int v = row_lo + idx;
    const uint32_t av = (act[(int64_t)v * W + w] >> b0) & tile_mask;
    if (!av) continue;
    const int e0 = rowptr[v], e1 = rowptr[v + 1];
    if (e0 == e1) continue;
    if (long_threshold > 0 && e1 - e0 > long_threshold) continue;  // hub row: compact_edges_long_kernel
    unsigned long long base = 0;  // lane t: write cursor of slot t (offset into the concatenated lists)
    if ((av >> lane) & 1u) base = scanned[(int64_t)lane * N + idx] & kEdgeMask;
    for (int b = e0; b < e1; b += 32) {
      const int e = b + lane;
      int u = -1;
      uint32_t bits = 0;
      if (e < e1) {
        u = __ldg(col + e);
        bits = (__ldg(ebits + e) >> b0) & tile_mask;
      }
      uint32_t rem = __reduce_or_sync(0xffffffffu, bits);  // slots with an active edge in this batch
      while (rem) {
        const int t = __ffs(rem) - 1;
        rem &= rem - 1;
        const uint32_t m = __ballot_sync(0xffffffffu, (bits >> t) & 1u);
        const unsigned long long bt = __shfl_sync(0xffffffffu, base, t);
        if ((bits >> t) & 1u) ccol[bt + __popc(m & lt)] = u;
        if (lane == t) base += __popc(m);
      }
    }
  }
}



// hub rows of pass 3: one CTA per row.  The warps take contiguous slices of the row's in-edges; a first sweep counts the
// active edges of every slot per slice, the exclusive sum over the slices gives each warp its write cursors, a second
// sweep writes the sources (CSR order kept).
__global__ void __launch_bounds__(256) compact_edges_long_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                 const uint32_t* __restrict__ ebits, const uint32_t* __restrict__ act, int W,
                                                                 int w, int b0, int nb, int N, const unsigned long long* __restrict__ scanned,
                                                                 int32_t* __restrict__ ccol, const int32_t* __restrict__ long_rows, int row_lo) {
  __shared__ int s_cnt[8][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t tile_mask = nb == 32 ? 0xffffffffu : ((1u << nb) - 1u);
  const uint32_t lt = (1u << lane) - 1u;
  const int v = long_rows[blockIdx.x];
  const uint32_t av = (act[(int64_t)v * W + w] >> b0) & tile_mask;
  if (!av) return;
  const int e0 = rowptr[v], e1 = rowptr[v + 1];
  const int nbatch = (e1 - e0 + 31) >> 5;
  const int bs = e0 + 32 * ((wib * nbatch) >> 3), be = min(e1, e0 + 32 * (((wib + 1) * nbatch) >> 3));
  int cnt = 0;  // lane t: active edges of slot t in this warp's slice
  for (int b = bs; b < be; b += 32) {
    const int e = b + lane;
    const uint32_t bits = e < be ? (__ldg(ebits + e) >> b0) & tile_mask : 0u;
    const int n = min(32, be - b);
    for (int j = 0; j < n; ++j) cnt += (__shfl_sync(0xffffffffu, bits, j) >> lane) & 1u;
  }
  s_cnt[wib][lane] = cnt;
  __syncthreads();
  unsigned long long base = 0;
  if ((av >> lane) & 1u) {
    base = scanned[(int64_t)lane * N + (v - row_lo)] & kEdgeMask;
    for (int w8 = 0; w8 < wib; ++w8) base += s_cnt[w8][lane];
  }
  for (int b = bs; b < be; b += 32) {
    const int e = b + lane;
    int u = -1;
    uint32_t bits = 0;
    if (e < be) {
      u = __ldg(col + e);
      bits = (__ldg(ebits + e) >> b0) & tile_mask;
    }
    uint32_t rem = __reduce_or_sync(0xffffffffu, bits);
    while (rem) {
      const int t = __ffs(rem) - 1;
      rem &= rem - 1;
      const uint32_t m = __ballot_sync(0xffffffffu, (bits >> t) & 1u);
      const unsigned long long bt = __shfl_sync(0xffffffffu, base, t);
      if ((bits >> t) & 1u) ccol[bt + __popc(m & lt)] = u;
      if (lane == t) base += __popc(m);
    }
  }
}

// ---- L2 eviction-priority hints (createpolicy + .L2::cache_hint): the (slot, chunk) pass wants its gathered
// operand to own the L2; index streams, the self / addend rows and the output are touched once ----
__device__ __forceinline__ uint64_t l2_policy(int kind /*0 normal, 1 evict_first, 2 evict_last*/) {
  uint64_t p;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ int ld_hint(const int32_t* ptr, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ld_hint(const uint32_t* ptr, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ float4 ld_hint4(const float* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_hint4(float* ptr, const float4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// ------------------------------------------------------------------------------------------
// masked SpMM over the compact lists: work item = (coalition slot, chunk, 128 active rows), items
// ordered slot-major / chunk / tile and dealt round-robin to a persistent grid, so that at any time
// the whole GPU works on one (slot, chunk) pass (+- skew) whose gathered operand fits in L2.
// A row is owned by CW/4 lanes (float4 each); a warp advances 32/(CW/4) rows together.
// ------------------------------------------------------------------------------------------
// (CspmmArgs lives in compact_internal.cuh: compact_bulk.cu launches over the same lists)

// finishes one row piece (CW / 4 lanes x float4): normalisation, self term, addend, activation, pre-scale, store
template <int CW>
__device__ __forceinline__ void cspmm_epilogue(const CspmmArgs& a, uint32_t cnt, const float4& acc, const float* in_c, float* out_c,
                                               const float* add_c, const float4& bias, int v, uint64_t pol_s, uint64_t pol_g) {
  float4 o;
  float dinv = 1.0f;
  if (a.kind == XPGNN_CONV_GCN) {
    dinv = gcn_dinv(cnt);
    const float4 self = ld_hint4(in_c + (int64_t)v * CW, pol_g);
    const float sw = a.layer0 ? dinv : 1.0f;  // layers >= 1 gather operands already scaled by deg^-1/2
    o.x = dinv * fmaf(sw, self.x, acc.x); o.y = dinv * fmaf(sw, self.y, acc.y);
    o.z = dinv * fmaf(sw, self.z, acc.z); o.w = dinv * fmaf(sw, self.w, acc.w);
  } else {
    const float inv = 1.0f / (float)max(cnt, 1u);
    o.x = acc.x * inv; o.y = acc.y * inv; o.z = acc.z * inv; o.w = acc.w * inv;
  }
  o.x += bias.x; o.y += bias.y; o.z += bias.z; o.w += bias.w;
  if (add_c) {
    const float4 ad = ld_hint4(add_c + (int64_t)v * CW, pol_s);
    o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
  }
  o.x = apply_act(o.x, a.act_fn); o.y = apply_act(o.y, a.act_fn);
  o.z = apply_act(o.z, a.act_fn); o.w = apply_act(o.w, a.act_fn);
  if (a.prescale) { o.x *= dinv; o.y *= dinv; o.z *= dinv; o.w *= dinv; }
  st_hint4(out_c + (int64_t)v * CW, o, pol_s);
}

template <int CW, bool WEIGHTED, int OCC>
__global__ void __launch_bounds__(256, OCC) cspmm_kernel(const CspmmArgs a) {
  constexpr int G = CW / 4;        // lanes per row
  constexpr int RPW = 32 / G;      // rows per warp
  constexpr int RPI = 8 * RPW;     // rows per CTA iteration
  constexpr int ITERS = 128 / RPI;
  __shared__ int s_start[33];
  __shared__ int s_item;
  if (threadIdx.x <= a.nb) s_start[threadIdx.x] = a.slot_tile_start[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % G, grp = lane / G;
  const int total = s_start[a.nb] * a.n_chunks;
  const uint64_t pol_s = l2_policy(a.l2_stream), pol_g = l2_policy(a.l2_gather);
  int t = 0;
  int idx = blockIdx.x;
  while (true) {
    if (a.counter) {  // items are handed out in global order: the in-flight window is one grid wide
      if (threadIdx.x == 0) s_item = atomicAdd(a.counter, 1);
      __syncthreads();
      idx = s_item;
      __syncthreads();
    }
    if (idx >= total) break;
    while (idx >= s_start[t + 1] * a.n_chunks) ++t;
    const int ntb = s_start[t + 1] - s_start[t];
    const int rem = idx - s_start[t] * a.n_chunks;
    const int c = rem / ntb, tb = rem - c * ntb;
    const int n_act = a.slot_info[t].x;
    const int32_t* al = a.act_list + (int64_t)t * a.N;
    const uint32_t* rp = a.rowptr_c + (int64_t)t * (a.N + 1);
    const int32_t* cc = a.ccol + a.slot_base[t];
    const float* in_c = a.in + (int64_t)t * a.in_s_stride + (int64_t)c * a.in_chunk_stride + sub * 4;
    const float* wg = WEIGHTED ? a.wgt + (int64_t)t * a.N : nullptr;
    float* out_c = a.out + (int64_t)t * a.out_s_stride + (int64_t)c * a.out_chunk_stride + sub * 4;
    const float* add_c = a.addend ? a.addend + (int64_t)c * a.add_chunk_stride + sub * 4 : nullptr;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) bias = __ldg(reinterpret_cast<const float4*>(a.bias + c * CW + sub * 4));
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
      const int i = tb * 128 + it * RPI + warp * RPW + grp;
      const bool valid = i < n_act;
      int v = 0;
      uint32_t e0 = 0, cnt = 0;
      if (valid) {  // streaming reads: keep the L2 for the gathered operand
        v = ld_hint(al + i, pol_s);
        e0 = ld_hint(rp + i, pol_s);
        cnt = ld_hint(rp + i + 1, pol_s) - e0;
      }
      const bool is_long = a.long_cnt > 0 && cnt > (uint32_t)a.long_cnt;  // left to cspmm_long_kernel
      const uint32_t cnt_all = cnt;
      if (is_long) cnt = 0;
      const uint32_t maxcnt = __reduce_max_sync(0xffffffffu, cnt);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (uint32_t base = 0; base < maxcnt; base += G) {
        int my = -1;
        float myw = 0.f;
        if (base + sub < cnt) {
          my = ld_hint(cc + e0 + base + sub, pol_s);
          if (WEIGHTED) myw = __ldg(wg + my);
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const int u = __shfl_sync(0xffffffffu, my, grp * G + j);
          const float wj = WEIGHTED ? __shfl_sync(0xffffffffu, myw, grp * G + j) : 1.0f;
          if (u >= 0) {
            const float4 x = ld_hint4(in_c + (int64_t)u * CW, pol_g);
            acc.x = fmaf(wj, x.x, acc.x); acc.y = fmaf(wj, x.y, acc.y);
            acc.z = fmaf(wj, x.z, acc.z); acc.w = fmaf(wj, x.w, acc.w);
          }
        }
      }
      if (valid && !is_long) cspmm_epilogue<CW>(a, cnt_all, acc, in_c, out_c, add_c, bias, v, pol_s, pol_g);
    }
    if (!a.counter) idx += gridDim.x;
  }
}

// Long compact rows (hub rows of power-law graphs): one CTA per (row, chunk).  The 32 row groups of the CTA take
// contiguous slices of the row's active in-edges; the partial sums are added in a fixed order through shared memory.
// Items are ordered slot / chunk / row (the long rows of a slot sit in their own segment of the list) and handed out in
// that order by a work counter, so the whole grid gathers from ONE (slot, chunk) operand at a time, like the main kernel
// (r02 launch list of the R-MAT C3 step: with the items in list order -- 4 chunks of a row next to each other, slots mixed --
// this kernel took 5.8 ms per tile for a third of the active edges).  The source ids of the next batch load while the current
// batch gathers.
template <int CW, bool WEIGHTED>
__global__ void __launch_bounds__(256) cspmm_long_kernel(const CspmmArgs a) {
  constexpr int G = CW / 4, NG = 256 / G;
  __shared__ float4 s_red[256];
  __shared__ int s_off[33];  // first item of every slot: items = rows x chunks
  __shared__ int s_item;
  const int lane = threadIdx.x & 31, sub = threadIdx.x % G, g = threadIdx.x / G, grp = lane / G;
  const uint64_t pol_s = l2_policy(a.l2_stream), pol_g = l2_policy(a.l2_gather);
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int t = 0; t < a.nb; ++t) { s_off[t] = acc; acc += a.n_long_list[t] * a.n_chunks; }
    s_off[a.nb] = acc;
  }
  __syncthreads();
  const int total = s_off[a.nb];
  int t = 0;
  int idx = blockIdx.x;
  while (true) {
    if (a.counter_long) {
      __syncthreads();  // everybody has read the previous item (and, the first time, s_off)
      if (threadIdx.x == 0) s_item = atomicAdd(a.counter_long, 1);
      __syncthreads();
      idx = s_item;
    }
    if (idx >= total) break;
    while (idx >= s_off[t + 1]) ++t;
    const int n_t = a.n_long_list[t];
    const int rem = idx - s_off[t];
    const int c = rem / n_t, i = a.long_list[(int64_t)t * a.long_cap + (rem - c * n_t)];
    const uint32_t* rp = a.rowptr_c + (int64_t)t * (a.N + 1);
    const int v = a.act_list[(int64_t)t * a.N + i];
    const uint32_t e0 = rp[i], cnt = rp[i + 1] - e0;
    const int32_t* cc = a.ccol + a.slot_base[t] + e0;
    const float* in_c = a.in + (int64_t)t * a.in_s_stride + (int64_t)c * a.in_chunk_stride + sub * 4;
    const float* wg = WEIGHTED ? a.wgt + (int64_t)t * a.N : nullptr;
    float* out_c = a.out + (int64_t)t * a.out_s_stride + (int64_t)c * a.out_chunk_stride + sub * 4;
    const float* add_c = a.addend ? a.addend + (int64_t)c * a.add_chunk_stride + sub * 4 : nullptr;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) bias = __ldg(reinterpret_cast<const float4*>(a.bias + c * CW + sub * 4));
    const uint32_t nbt = (cnt + G - 1) / G;  // batches of G edges, split evenly over the NG groups
    const uint32_t bs = (uint32_t)(((uint64_t)g * nbt) / NG) * G, be = min(cnt, (uint32_t)(((uint64_t)(g + 1) * nbt) / NG) * G);
    const uint32_t span = __reduce_max_sync(0xffffffffu, be > bs ? be - bs : 0u);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int my = -1;
    float myw = 0.f;
    if (bs + sub < be) {
      my = ld_hint(cc + bs + sub, pol_s);
      if (WEIGHTED) myw = __ldg(wg + my);
    }
    for (uint32_t off = 0; off < span; off += G) {
      const int cur = my;
      const float curw = myw;
      my = -1;
      myw = 0.f;
      if (off + G < span && bs + off + G + sub < be) {  // next batch's ids fly while this batch gathers
        my = ld_hint(cc + bs + off + G + sub, pol_s);
        if (WEIGHTED) myw = __ldg(wg + my);
      }
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const int u = __shfl_sync(0xffffffffu, cur, grp * G + j);
        const float wj = WEIGHTED ? __shfl_sync(0xffffffffu, curw, grp * G + j) : 1.0f;
        if (u >= 0) {
          const float4 x = ld_hint4(in_c + (int64_t)u * CW, pol_g);
          acc.x = fmaf(wj, x.x, acc.x); acc.y = fmaf(wj, x.y, acc.y);
          acc.z = fmaf(wj, x.z, acc.z); acc.w = fmaf(wj, x.w, acc.w);
        }
      }
    }
    __syncthreads();  // the previous item's reduction has been read
    s_red[threadIdx.x] = acc;
    __syncthreads();
    if (g == 0) {
      for (int k = 1; k < NG; ++k) {
        const float4 p = s_red[k * G + sub];
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
      }
      cspmm_epilogue<CW>(a, cnt, acc, in_c, out_c, add_c, bias, v, pol_s, pol_g);
    }
    if (!a.counter_long) idx += gridDim.x;
  }
}

// ------------------------------------------------------------------------------------------
// Segmented masked SpMM (layers >= 1, aggregate-first: unweighted sums of pre-scaled rows, out = scale * sum).
//
// The row-lockstep kernel above keeps ~1 gather instruction in flight per warp: every 4-row step is a chain of
// dependent round trips (row ids -> list offsets -> source ids -> gathers -> self row), rows of unequal length
// idle their lanes, and at 32 registers (8 CTAs / SM) it spills (r01 ncu: as many local as global load requests,
// 111.6 GB through the L2 -> SM fabric for 97 GB of useful sectors).  Here a warp owns a block of 32 compact rows:
//   * one coalesced load brings the block's row ids and list offsets, issued one block ahead of its use
//     (the block index comes from the global in-order counter two blocks ahead);
//   * the block's gather stream -- for GCN the row itself first (the operands of layers >= 1 are pre-scaled by
//     deg^-1/2, so the unit self loop is one more unweighted term), then its active sources -- is staged in shared
//     memory as words  id | row << 26 | last << 31 ;
//   * the stream is cut at ROW boundaries into four pieces of about equal length, one per group of 8 lanes (float4
//     each = one 128-byte row piece per group and gather): a warp-level segmented sum whose lanes stay busy whatever
//     the row lengths.  Every lane keeps D gathers in flight through a rotating register queue; a word with the `last`
//     bit closes its row: the sum goes to the warp's shared-memory tile (one predicated store, no divergent code);
//   * the 32 sums are scaled and written by a converged epilogue, 4 rows per instruction.
// Sums run in list order (self first), so results do not depend on the schedule.  Rows with more than long_cnt active
// in-edges stay with cspmm_long_kernel.
// ------------------------------------------------------------------------------------------
constexpr int kSegCap = 448;     // stream positions staged per round; a longer row is summed alone, chunk by chunk.  (320 until the r02 profile: a C3 block of 32 rows has 352 +- 18 positions, so nearly every block took a second round for its last 3 rows.)
constexpr int kSegTileLd = 32;   // floats per tile row (128 B; a warp-wide 16-byte access is 4 wavefronts with or without padding)
// Shared memory per CTA: 25.9 KB (448 staged positions, unpadded tile rows), so that SEVEN CTAs fit the 196 KB carve-out and the L1
// keeps 60 KB.  Measured (r02 calls 48 - 51, ms per C3 tile): with 28.9 KB per CTA (512 positions, padded rows) 7 CTAs need the
// 228 KB carve-out and the kernel takes 14.0 with 28 KB of L1 (6 CTAs, forced 228 KB: 14.3; 6 CTAs at 196 KB: 9.4); with 25.9 KB:
// 8 in flight x 6 CTAs 9.42, 7 x 7 CTAs 9.08 (shipped, option seg = 7), 8 x 7 CTAs 9.11, 6 x 7 CTAs 9.20.  (Every earlier
// occupancy sweep of this kernel above 6 CTAs per SM had silently run with 28 KB of L1.)
constexpr int kSegRowwise = 24;  // blocks whose longest row has at most this many entries are staged row by row
static_assert(kSegCap % 32 == 0 && kSegCap >= 128, "staging buffer: whole 32-position steps, room for the 4 x 128-byte partial sums");

__device__ __forceinline__ float4 seg_gather(const char* base, uint32_t word) {  // word = source id | last << 31
  const char* p;
  asm("mad.wide.u32 %0, %1, 128, %2;" : "=l"(p) : "r"(word & 0x7fffffffu), "l"(base));  // one IMAD.WIDE: base + id * 128 bytes
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void seg_add(float4& acc, const float4& x) {
  // two packed adds (FFMA2 with a unit multiplier) instead of four FADD: the loop is issue-bound before it is fabric-bound
  asm("{\n\t.reg .b64 a0, a1, x0, x1, one;\n\t"
      "mov.b64 a0, {%0, %1};\n\tmov.b64 a1, {%2, %3};\n\tmov.b64 x0, {%4, %5};\n\tmov.b64 x1, {%6, %7};\n\t"
      "mov.b64 one, {0f3F800000, 0f3F800000};\n\t"
      "fma.rn.f32x2 a0, x0, one, a0;\n\tfma.rn.f32x2 a1, x1, one, a1;\n\t"
      "mov.b64 {%0, %1}, a0;\n\tmov.b64 {%2, %3}, a1;\n\t}"
      : "+f"(acc.x), "+f"(acc.y), "+f"(acc.z), "+f"(acc.w)
      : "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w));
}

template <int D, int OCC, int PF>  // PF: list entries per row prefetched during the previous block's epilogue (0: none)
__global__ void __launch_bounds__(128, OCC) cspmm_seg_kernel(const CspmmArgs a) {
  constexpr int WARPS = 4;
  __shared__ int s_start[33];
  __shared__ int s_nact[32];
  __shared__ __align__(16) uint32_t s_ids[WARPS][kSegCap];
  __shared__ __align__(16) float s_tile[WARPS][32 * kSegTileLd];
  __shared__ int s_rowv[WARPS][32];        // original node id | -1: no row / hub row (not written here)
  __shared__ int s_aend[WARPS][32];        // stream position one past the row's last entry
  __shared__ uint32_t s_e[WARPS][32];      // first list entry of the row
  __shared__ float s_sc[WARPS][32];        // output scale of the row: deg^-1/2 (GCN) | 1 / active in-edges (SAGE mean; 0: none).
                                           // Computed ONCE per row by the lane that owns its metadata: the sqrt + division cost ~40
                                           // instructions, and the epilogue used to run them in every lane of every 4-row step
  __shared__ int s_next[WARPS][2];         // decoded item n + 1 (slot, chunk)
  if (threadIdx.x <= a.nb) s_start[threadIdx.x] = a.slot_tile_start[threadIdx.x];
  if (threadIdx.x < a.nb) s_nact[threadIdx.x] = a.slot_info[threadIdx.x].x;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7, grp = lane >> 3;
  const int total = s_start[a.nb] * a.n_chunks * 4;  // items = 32-row blocks, ordered slot / chunk / 128-row tile / quarter
  const bool gcn = a.kind == XPGNN_CONV_GCN;
  uint32_t* ids = s_ids[warp];
  float* tile = s_tile[warp];

  // ---- two-deep item pipeline: item n is processed while the metadata loads of n + 1 and the counter fetch of n + 2 fly ----
  int t_cur = 0;
  // the counter fetch of block n + 2 is ISSUED a block before its value is needed and broadcast only then (r02 source-level
  // profile: 4.6 % of the warp stalls sat in a fetch-and-broadcast at this point -- one atomic round trip per block)
  auto grab_raw = [&]() {
    int it = 0;
    if (lane == 0) it = atomicAdd(a.counter, 1);
    return it;  // valid in lane 0
  };
  // decodes the item into s_next and issues the loads of its row ids / list offsets
  auto load_meta = [&](int item, int& v, uint32_t& e, uint32_t& f) {
    v = -1; e = 0; f = 0;
    if (item >= total) return;
    const int idx = item >> 2;
    while (idx >= s_start[t_cur + 1] * a.n_chunks) ++t_cur;
    const int ntb = s_start[t_cur + 1] - s_start[t_cur];
    const int rem = idx - s_start[t_cur] * a.n_chunks;
    const int c = rem / ntb;
    const int row0 = (rem - c * ntb) * 128 + (item & 3) * 32;
    if (lane == 0) { s_next[warp][0] = t_cur; s_next[warp][1] = c; }
    const int i = row0 + lane;
    if (i < s_nact[t_cur]) {
      v = __ldcs(a.act_list + (int64_t)t_cur * a.N + i);
      const uint32_t* rp = a.rowptr_c + (int64_t)t_cur * (a.N + 1) + i;
      e = __ldcs(rp);
      f = __ldcs(rp + 1);
    }
  };
  int item1 = __shfl_sync(0xffffffffu, grab_raw(), 0);
  int v1;
  uint32_t e1, f1;
  load_meta(item1, v1, e1, f1);
  int raw2 = grab_raw();
  // The first kPf list entries of every row of the NEXT block are loaded into registers while the current block runs its
  // epilogue (the gather queue's registers are free then; the next block's row offsets arrived long ago): the staging of
  // a block then stores them without waiting -- its two exposed list round trips were 7 % of the warp stalls (r02 profile).
  constexpr int kPf = PF;
  uint32_t pf[kPf > 0 ? kPf : 1];
  auto prefetch_ids = [&]() {  // (v1, e1, f1) and s_next describe the block whose staging comes next
    if (kPf == 0) return;
    __syncwarp();
    const uint32_t cnt_n = f1 - e1;
    const bool mine_n = item1 < total && v1 >= 0 && !(a.long_cnt > 0 && cnt_n > (uint32_t)a.long_cnt);
    const int32_t* cc_n = a.ccol + (item1 < total ? a.slot_base[s_next[warp][0]] : 0ll) + e1;
#pragma unroll
    for (int j = 0; j < kPf; ++j) pf[j] = (mine_n && (uint32_t)j < cnt_n) ? (uint32_t)__ldg(cc_n + j) : 0u;
  };
  prefetch_ids();

  while (item1 < total) {
    __syncwarp();
    const int t = s_next[warp][0], c = s_next[warp][1];
    __syncwarp();
    int nA, n_l, v_l;
    uint32_t e_l, nonempty;
    {
      const uint32_t cnt_l = f1 - e1;
      const bool is_long = a.long_cnt > 0 && cnt_l > (uint32_t)a.long_cnt;
      const bool mine = v1 >= 0 && !is_long;
      n_l = mine ? (int)cnt_l + (gcn ? 1 : 0) : 0;
      v_l = v1; e_l = e1;
      int a_end = n_l;  // inclusive prefix of the stream lengths
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, a_end, o);
        if (lane >= o) a_end += y;
      }
      s_rowv[warp][lane] = mine ? v1 : -1;
      s_aend[warp][lane] = a_end;
      s_e[warp][lane] = e1;
      s_sc[warp][lane] = gcn ? gcn_dinv(cnt_l) : (cnt_l ? 1.0f / (float)cnt_l : 0.0f);
      nA = __shfl_sync(0xffffffffu, a_end, 31);
      nonempty = __ballot_sync(0xffffffffu, n_l > 0);  // the tile holds the sums of these rows, in row order
      item1 = __shfl_sync(0xffffffffu, raw2, 0);
      load_meta(item1, v1, e1, f1);   // consumed at the top of the next iteration
      raw2 = grab_raw();
    }
    const int n_max = __reduce_max_sync(0xffffffffu, n_l);
    __syncwarp();
    const int32_t* cc = a.ccol + a.slot_base[t];
    const char* in_c = reinterpret_cast<const char*>(a.in + (int64_t)t * a.in_s_stride + (int64_t)c * a.in_chunk_stride + sub * 4);

    int r_lo = 0;  // rows [r_lo, r_hi) of this round: as many whole rows as fit kSegCap positions
    while (r_lo < 32) {
      const int base = r_lo ? s_aend[warp][r_lo - 1] : 0;
      if (base >= nA) break;  // the remaining rows have no entries
      const uint32_t over = __ballot_sync(0xffffffffu, s_aend[warp][lane] - base > kSegCap) & ~((1u << r_lo) - 1u);
      const int r_hi = over ? __ffs(over) - 1 : 32;
      if (r_hi == r_lo) {
        // one row longer than a round (graphs without hub rows have no long-row list, so any length can arrive here): the
        // four groups sum contiguous quarters of every kSegCap-position chunk, the four partial sums are added in group order
        const int n_row = s_aend[warp][r_lo] - base;
        const uint32_t e_row = s_e[warp][r_lo];
        const int v_row = s_rowv[warp][r_lo];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int S = 0; S < n_row; S += kSegCap) {
          const int len = min(kSegCap, n_row - S);
          for (int i = lane; i < len; i += 32) {
            const int pos = S + i;
            ids[i] = (gcn && pos == 0) ? (uint32_t)v_row : (uint32_t)__ldcs(cc + e_row + (uint32_t)(pos - (gcn ? 1 : 0)));
          }
          __syncwarp();
          const int per = (len + 3) >> 2;
          const int q1 = min((grp + 1) * per, len);
          for (int q = min(grp * per, len); q < q1; ++q) seg_add(acc, seg_gather(in_c, ids[q]));
          __syncwarp();
        }
        float4* part = reinterpret_cast<float4*>(ids);  // 4 groups x 8 lanes x float4 = 512 bytes of the staging buffer
        part[grp * 8 + sub] = acc;
        __syncwarp();
        if (grp == 0) {
          float4 tot = part[sub];
          for (int g = 1; g < 4; ++g) { const float4 pg = part[g * 8 + sub]; tot.x += pg.x; tot.y += pg.y; tot.z += pg.z; tot.w += pg.w; }
          *reinterpret_cast<float4*>(tile + __popc(nonempty & ((1u << r_lo) - 1u)) * kSegTileLd + sub * 4) = tot;
        }
        __syncwarp();
        r_lo = r_lo + 1;
        continue;
      }
      const int len = s_aend[warp][r_hi - 1] - base;
      const int a_l = (lane ? s_aend[warp][lane - 1] : 0) - base;  // start of row `lane` relative to the round
      const bool in_round = lane >= r_lo && lane < r_hi;
      // ---- stage the stream words of this round: source id | last << 31 ----
      if (n_max <= kSegRowwise) {  // short rows: lane = row
        // 8 list entries per step, loads first: the r02 source-level profile showed 7 % of the warp stalls in a one-entry-per-
        // iteration form of this loop (load -> use -> store, one exposed round trip per entry)
        const bool have = in_round && n_l > 0;
        const int ne = have ? n_l - (gcn ? 1 : 0) : 0;
        int o = a_l;
        if (have && gcn) ids[o++] = (uint32_t)v_l | (n_l == 1 ? 0x80000000u : 0u);
#pragma unroll
        for (int j = 0; j < kPf; ++j)  // entries prefetched during the previous block's epilogue
          if (j < ne) ids[o + j] = pf[j] | (j == ne - 1 ? 0x80000000u : 0u);
        for (int k0 = kPf; k0 < n_max; k0 += 8) {  // the rest (rows longer than kPf entries); n_max: warp-uniform bound
          uint32_t t8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) t8[j] = k0 + j < ne ? (uint32_t)__ldg(cc + e_l + k0 + j) : 0u;  // (ld.global.cs here: 9.72 vs 9.50 ms -- the entries of a row share sectors in L1)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (k0 + j < ne) ids[o + k0 + j] = t8[j] | (k0 + j == ne - 1 ? 0x80000000u : 0u);
        }
      } else {                     // lane = stream position, its row by binary search over the row ends
        // four positions per lane and step, searches and list loads first, stores last: one exposed list round trip per 128
        // positions instead of one per 32 (r02 source-level profile of R-MAT C3: 11 % of the warp stalls in this loop)
        for (int i0 = lane; i0 < len; i0 += 128) {
          uint32_t word[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
            word[u] = 0;
            if (i < len) {
              const int pos = base + i;
              int r = 0;  // rows whose stream ends at or before pos
#pragma unroll
              for (int step = 16; step >= 1; step >>= 1)
                if (s_aend[warp][r + step - 1] <= pos) r += step;
              const int ar = r ? s_aend[warp][r - 1] : 0;
              int id;
              if (gcn) id = pos == ar ? s_rowv[warp][r] : __ldcs(cc + s_e[warp][r] + (uint32_t)(pos - ar - 1));
              else id = __ldcs(cc + s_e[warp][r] + (uint32_t)(pos - ar));
              word[u] = (uint32_t)id | (pos == s_aend[warp][r] - 1 ? 0x80000000u : 0u);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i0 + 32 * u < len) ids[i0 + 32 * u] = word[u];
        }
      }
      // ---- cut at row boundaries into four pieces of ~len / 4 positions: piece g starts at the first non-empty row that
      // begins at or after position g * per; its sums go to consecutive tile rows ----
      // (the boundary is the row start NEAREST to g * per: taking the first start at or after it made piece 0 about one row
      // longer and piece 3 one row shorter than the mean, and the warp runs as long as its longest piece)
      const int per = (len + 3) >> 2;
      const int half = (len / max(r_hi - r_lo, 1)) >> 1;  // half an average row
      int q = 0, q1 = len, trow = __popc(nonempty & ((1u << r_lo) - 1u));
#pragma unroll
      for (int g = 1; g < 4; ++g) {
        const uint32_t m = __ballot_sync(0xffffffffu, in_round && n_l > 0 && a_l + half >= g * per);
        const int first = m ? __ffs(m) - 1 : 32;
        const int start = m ? __shfl_sync(0xffffffffu, a_l, first & 31) : len;
        if (grp == g) { q = start; trow = __popc(nonempty & ((1u << (first & 31)) - 1u)); }
        if (grp == g - 1) q1 = start;
      }
      uint32_t tp = (uint32_t)__cvta_generic_to_shared(tile + trow * kSegTileLd + sub * 4);
      __syncwarp();
      // ---- stream: D gathers in flight per lane; a `last` word closes its row into the tile ----
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 x[D];
      uint32_t w[D];
      auto close_row = [&](uint32_t wk) {
        if ((int)wk < 0) {
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tp), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
          tp += kSegTileLd * 4;
          acc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      // (folding the reset into the add -- acc = acc * keep + x with keep = 0 after a flush, no zeroing moves -- measured the
      // same: 10.79 vs 10.77 ms per C3 tile)
      // (a single predicated loop over the longest piece -- no divergence between groups whose pieces differ in length --
      // measured slower: 13.8 vs 11.1 ms per C3 tile; the unpredicated steady state matters more)
#pragma unroll
      for (int k = 0; k < D; ++k) {
        w[k] = 0; x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q + k < q1) {
          w[k] = ids[q + k];
          x[k] = seg_gather(in_c, w[k]);
        }
      }
      for (; q + 2 * D <= q1; q += D) {  // steady state: every queue slot is consumed and refilled, no predicates
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const uint32_t wk = w[k];
          seg_add(acc, x[k]);
          w[k] = ids[q + k + D];
          x[k] = seg_gather(in_c, w[k]);
          close_row(wk);
        }
      }
      for (; q < q1; q += D) {           // drain
#pragma unroll
        for (int k = 0; k < D; ++k) {
          if (q + k < q1) {
            const uint32_t wk = w[k];
            seg_add(acc, x[k]);
            if (q + k + D < q1) {
              w[k] = ids[q + k + D];
              x[k] = seg_gather(in_c, w[k]);
            }
            close_row(wk);
          }
        }
      }
      __syncwarp();
      r_lo = r_hi;
    }
    prefetch_ids();  // list entries of the next block: in flight during the epilogue and the next block's scan
    // ---- epilogue: scale and write the 32 rows, 4 rows per instruction ----
    float* out_c = a.out + (int64_t)t * a.out_s_stride + (int64_t)c * a.out_chunk_stride + sub * 4;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + grp;
      const int v = s_rowv[warp][r];
      if (v >= 0) {
        const float sc = s_sc[warp][r];
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);  // SAGE row without an active in-edge: empty mean (no tile row, scale 0)
        if ((nonempty >> r) & 1u) o = *reinterpret_cast<const float4*>(tile + __popc(nonempty & ((1u << r) - 1u)) * kSegTileLd + sub * 4);
        __stcs(reinterpret_cast<float4*>(out_c + (int64_t)v * 32), make_float4(o.x * sc, o.y * sc, o.z * sc, o.w * sc));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Segmented masked SpMM, shared-memory ring variant.  The register-queue kernel above is latency bound (r02 ncu: 7.5 of
// 9.4 warp-cycles per issue stalled on the long scoreboard, 8.3 TB/s through the fabric): what it can keep in flight is
// what its registers hold (24 warps x 4 groups x 8 row pieces = 98 KB per SM).  Here the gathers are cp.async copies
// (LDGSTS, 16 bytes per lane = one 128-byte row piece per group) into a per-group ring of R slots in shared memory, so
// the data in flight is bounded by shared memory instead: 16 warps x 4 groups x 24 slots = 196 KB per SM.  A lane reads
// back exactly the 16 bytes it copied (cp.async.wait_group orders a thread's own copies), accumulates, and refills the
// slot R positions ahead.  Finished rows are scaled and stored straight from registers.  Same stream, same cut at row
// boundaries, same summation order as cspmm_seg_kernel: bit-identical results.
// ------------------------------------------------------------------------------------------
constexpr int kRingWarps = 4;
template <int R>
struct RingSmem {
  float4 ring[kRingWarps][4][R][8];   // [warp][group][slot][lane of the group]: 16 bytes each
  uint32_t ids[kRingWarps][kSegCap];
  int rowv[kRingWarps][32];           // per block row: node id | -1
  int aend[kRingWarps][32];
  uint32_t e[kRingWarps][32];
  uint32_t cnt[kRingWarps][32];
  int outv[kRingWarps][32];           // k-th non-empty row of the block: node id and scale (the order rows close in)
  float outs[kRingWarps][32];
  int next[kRingWarps][2];
  int start[33];
  int nact[32];
};

template <int R, int OCC>
__global__ void __launch_bounds__(32 * kRingWarps, OCC) cspmm_ring_kernel(const CspmmArgs a) {
  extern __shared__ __align__(16) uint8_t ring_smem_raw[];
  RingSmem<R>& S = *reinterpret_cast<RingSmem<R>*>(ring_smem_raw);
  if (threadIdx.x <= a.nb) S.start[threadIdx.x] = a.slot_tile_start[threadIdx.x];
  if (threadIdx.x < a.nb) S.nact[threadIdx.x] = a.slot_info[threadIdx.x].x;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7, grp = lane >> 3;
  const int total = S.start[a.nb] * a.n_chunks * 4;
  const bool gcn = a.kind == XPGNN_CONV_GCN;
  uint32_t* ids = S.ids[warp];
  const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(&S.ring[warp][grp][0][sub]);  // slot stride: 8 x 16 bytes

  int t_cur = 0;
  auto grab = [&]() {
    int it = 0;
    if (lane == 0) it = atomicAdd(a.counter, 1);
    return __shfl_sync(0xffffffffu, it, 0);
  };
  auto load_meta = [&](int item, int& v, uint32_t& e, uint32_t& f) {
    v = -1; e = 0; f = 0;
    if (item >= total) return;
    const int idx = item >> 2;
    while (idx >= S.start[t_cur + 1] * a.n_chunks) ++t_cur;
    const int ntb = S.start[t_cur + 1] - S.start[t_cur];
    const int rem = idx - S.start[t_cur] * a.n_chunks;
    const int c = rem / ntb;
    const int row0 = (rem - c * ntb) * 128 + (item & 3) * 32;
    if (lane == 0) { S.next[warp][0] = t_cur; S.next[warp][1] = c; }
    const int i = row0 + lane;
    if (i < S.nact[t_cur]) {
      v = __ldcs(a.act_list + (int64_t)t_cur * a.N + i);
      const uint32_t* rp = a.rowptr_c + (int64_t)t_cur * (a.N + 1) + i;
      e = __ldcs(rp);
      f = __ldcs(rp + 1);
    }
  };
  int item1 = grab();
  int v1;
  uint32_t e1, f1;
  load_meta(item1, v1, e1, f1);
  int item2 = grab();

  while (item1 < total) {
    __syncwarp();
    const int t = S.next[warp][0], c = S.next[warp][1];
    __syncwarp();
    int nA, n_l, v_l;
    uint32_t e_l, nonempty;
    float* out_c = a.out + (int64_t)t * a.out_s_stride + (int64_t)c * a.out_chunk_stride + sub * 4;
    {
      const uint32_t cnt_l = f1 - e1;
      const bool is_long = a.long_cnt > 0 && cnt_l > (uint32_t)a.long_cnt;
      const bool mine = v1 >= 0 && !is_long;
      n_l = mine ? (int)cnt_l + (gcn ? 1 : 0) : 0;
      v_l = v1; e_l = e1;
      int a_end = n_l;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, a_end, o);
        if (lane >= o) a_end += y;
      }
      S.rowv[warp][lane] = mine ? v1 : -1;
      S.aend[warp][lane] = a_end;
      S.e[warp][lane] = e1;
      S.cnt[warp][lane] = cnt_l;
      nA = __shfl_sync(0xffffffffu, a_end, 31);
      nonempty = __ballot_sync(0xffffffffu, n_l > 0);
      if (n_l > 0) {  // rows close in this order
        const int k = __popc(nonempty & ((1u << lane) - 1u));
        S.outv[warp][k] = v1;
        S.outs[warp][k] = gcn ? gcn_dinv(cnt_l) : 1.0f / (float)cnt_l;
      }
      // SAGE: a row without an active in-edge aggregates to zero and has no stream entry
      uint32_t empties = __ballot_sync(0xffffffffu, mine && n_l == 0);
      while (empties) {
        const int r = __ffs(empties) - 1;
        empties &= empties - 1;
        const int vr = __shfl_sync(0xffffffffu, v1, r);
        if (lane < 8) __stcs(reinterpret_cast<float4*>(out_c + (int64_t)vr * 32), make_float4(0.f, 0.f, 0.f, 0.f));  // lane < 8: sub = lane
      }
      item1 = item2;
      load_meta(item1, v1, e1, f1);
      item2 = grab();
    }
    const int n_max = __reduce_max_sync(0xffffffffu, n_l);
    __syncwarp();
    const int32_t* cc = a.ccol + a.slot_base[t];
    const char* in_c = reinterpret_cast<const char*>(a.in + (int64_t)t * a.in_s_stride + (int64_t)c * a.in_chunk_stride + sub * 4);

    int r_lo = 0;
    while (r_lo < 32) {
      const int base = r_lo ? S.aend[warp][r_lo - 1] : 0;
      if (base >= nA) break;
      const uint32_t over = __ballot_sync(0xffffffffu, S.aend[warp][lane] - base > kSegCap) & ~((1u << r_lo) - 1u);
      int r_hi = over ? __ffs(over) - 1 : 32;
      int len, lone_len = 0;
      uint32_t lone_e = 0;
      int lone_v = 0;
      const bool lone = r_hi == r_lo;  // one row longer than a round: streamed alone, chunk by chunk, by group 0 .. 3 in quarters
      if (lone) {
        lone_len = S.aend[warp][r_lo] - base; lone_e = S.e[warp][r_lo]; lone_v = S.rowv[warp][r_lo];
        r_hi = r_lo + 1;
      }
      const int a_l = (lane ? S.aend[warp][lane - 1] : 0) - base;
      const bool in_round = lane >= r_lo && lane < r_hi;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int trow = __popc(nonempty & ((1u << r_lo) - 1u));
      for (int S0 = 0; S0 < (lone ? lone_len : 1); S0 += kSegCap) {
        if (lone) {
          len = min(kSegCap, lone_len - S0);
          for (int i = lane; i < len; i += 32) {
            const int pos = S0 + i;
            ids[i] = ((gcn && pos == 0) ? (uint32_t)lone_v : (uint32_t)__ldcs(cc + lone_e + (uint32_t)(pos - (gcn ? 1 : 0)))) |
                     (pos == lone_len - 1 ? 0x80000000u : 0u);
          }
        } else {
          len = S.aend[warp][r_hi - 1] - base;
          if (n_max <= kSegRowwise) {
            if (in_round && n_l > 0) {
              int o = a_l;
              if (gcn) ids[o++] = (uint32_t)v_l | (n_l == 1 ? 0x80000000u : 0u);
              const int ne = n_l - (gcn ? 1 : 0);
              for (int k = 0; k < ne; ++k) ids[o + k] = (uint32_t)__ldg(cc + e_l + k) | (k == ne - 1 ? 0x80000000u : 0u);
            }
          } else {
            for (int i = lane; i < len; i += 32) {
              const int pos = base + i;
              int r = 0;
#pragma unroll
              for (int step = 16; step >= 1; step >>= 1)
                if (S.aend[warp][r + step - 1] <= pos) r += step;
              const int ar = r ? S.aend[warp][r - 1] : 0;
              int id;
              if (gcn) id = pos == ar ? S.rowv[warp][r] : __ldcs(cc + S.e[warp][r] + (uint32_t)(pos - ar - 1));
              else id = __ldcs(cc + S.e[warp][r] + (uint32_t)(pos - ar));
              ids[i] = (uint32_t)id | (pos == S.aend[warp][r] - 1 ? 0x80000000u : 0u);
            }
          }
        }
        // pieces: at row boundaries (several rows) or plain quarters of the chunk (one long row; partial sums are combined below)
        const int per = (len + 3) >> 2;
        int q = min(grp * per, len), q1 = min((grp + 1) * per, len);
        if (!lone) {
          q = 0; q1 = len;
#pragma unroll
          for (int g = 1; g < 4; ++g) {
            const uint32_t m = __ballot_sync(0xffffffffu, in_round && n_l > 0 && a_l >= g * per);
            const int first = m ? __ffs(m) - 1 : 32;
            const int start = m ? __shfl_sync(0xffffffffu, a_l, first & 31) : len;
            if (grp == g) { q = start; trow = __popc(nonempty & ((1u << (first & 31)) - 1u)); }
            if (grp == g - 1) q1 = start;
          }
        }
        __syncwarp();
        // ---- stream through the ring: the copy of position p lands in slot p % R ----
        const int steps = __reduce_max_sync(0xffffffffu, q1 - q);
        int slot_w = 0;  // slot the next issued copy goes to
#pragma unroll 1
        for (int k = 0; k < R; ++k) {  // prologue: fill the ring (one committed group per position, empty groups past the end)
          if (q + k < q1) {
            const char* src;
            asm("mad.wide.u32 %0, %1, 128, %2;" : "=l"(src) : "r"(ids[q + k] & 0x7fffffffu), "l"(in_c));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_base + (uint32_t)k * 128u), "l"(src) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        int slot_r = 0;
#pragma unroll 1
        for (int k = 0; k < steps; ++k) {
          asm volatile("cp.async.wait_group %0;" ::"n"(R - 1) : "memory");  // the oldest of this thread's R pending copies has landed
          const int p = q + k;
          if (p < q1) {
            float4 x;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(ring_base + (uint32_t)slot_r * 128u) : "memory");
            const uint32_t wk = ids[p];
            seg_add(acc, x);
            if (p + R < q1) {
              const char* src;
              asm("mad.wide.u32 %0, %1, 128, %2;" : "=l"(src) : "r"(ids[p + R] & 0x7fffffffu), "l"(in_c));
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_base + (uint32_t)slot_r * 128u), "l"(src) : "memory");
            }
            if ((int)wk < 0 && !lone) {  // the row ends here: scale and store it
              const float sc = S.outs[warp][trow];
              __stcs(reinterpret_cast<float4*>(out_c + (int64_t)S.outv[warp][trow] * 32), make_float4(acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc));
              ++trow;
              acc = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          slot_r = slot_r + 1 == R ? 0 : slot_r + 1;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        (void)slot_w;
      }
      if (lone) {  // the four partial sums of the long row, added in group order
        float4* part = reinterpret_cast<float4*>(ids);
        part[grp * 8 + sub] = acc;
        __syncwarp();
        if (grp == 0) {
          float4 tot = part[sub];
          for (int g = 1; g < 4; ++g) { const float4 pg = part[g * 8 + sub]; tot.x += pg.x; tot.y += pg.y; tot.z += pg.z; tot.w += pg.w; }
          const float sc = S.outs[warp][trow];
          __stcs(reinterpret_cast<float4*>(out_c + (int64_t)S.outv[warp][trow] * 32), make_float4(tot.x * sc, tot.y * sc, tot.z * sc, tot.w * sc));
        }
        __syncwarp();
      }
      r_lo = r_hi;
    }
  }
}

// ------------------------------------------------------------------------------------------
// layers >= 1 of a HeteroConv(sum), transform-first: Z_r = H W_r^T has been computed per relation over the active
// rows of its source type; this kernel gathers, per destination row, over the compacted lists of ALL relations into
// the destination type, adds the merged root term and bias, applies the activation and writes the row ONCE
// (aggregate-first wrote an aggregate per relation and read-modify-wrote the output per relation).
// ------------------------------------------------------------------------------------------
struct MRel {
  const uint32_t* rowptr_c;    // [nb][nd + 1] compact in-edge offsets of this relation (nd = rows of the destination range)
  const long long* slot_base;
  const int32_t* ccol;
  const float* z;              // transformed sources, chunk-major, indexed by global source id (biased pointer)
  int64_t z_s_stride, z_chunk_stride;
  const float* wgt;            // GCN: [nb][.] per-source deg^-1/2, indexed by global source id (biased pointer) | NULL
  int kind;
};
struct CspmmMultiArgs {
  MRel rel[kMaxL0Rel];
  int n_rel, nb, nd, n_chunks, act_fn;
  const int2* slot_info;       // of the destination type (identical for all of its relations)
  const int32_t* slot_tile_start;
  const int32_t* act_list;     // [nb][nd]
  const float* zroot;          // merged root term H Wsum^T, chunk-major, indexed by global destination id (biased) | NULL
  int64_t zr_s_stride, zr_chunk_stride;
  const float* bias;           // sum of the relation biases | NULL
  float* out;
  int64_t out_s_stride, out_chunk_stride;
  int32_t* counter;
};

__global__ void __launch_bounds__(256, 4) cspmm_multi_kernel(const CspmmMultiArgs a) {
  constexpr int CW = 32, G = 8;
  __shared__ int s_start[33];
  __shared__ int s_item;
  if (threadIdx.x <= a.nb) s_start[threadIdx.x] = a.slot_tile_start[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7, grp = lane >> 3;
  const int total = s_start[a.nb] * a.n_chunks;
  const uint64_t pol_s = l2_policy(1), pol_g = l2_policy(0);
  int t = 0;
  while (true) {
    if (threadIdx.x == 0) s_item = atomicAdd(a.counter, 1);
    __syncthreads();
    const int idx = s_item;
    __syncthreads();
    if (idx >= total) break;
    while (idx >= s_start[t + 1] * a.n_chunks) ++t;
    const int ntb = s_start[t + 1] - s_start[t];
    const int rem = idx - s_start[t] * a.n_chunks;
    const int c = rem / ntb, tb = rem - c * ntb;
    const int n_act = a.slot_info[t].x;
    const int32_t* al = a.act_list + (int64_t)t * a.nd;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) bias = __ldg(reinterpret_cast<const float4*>(a.bias + c * CW + sub * 4));
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
      const int i = tb * 128 + it * 32 + warp * 4 + grp;
      const bool valid = i < n_act;
      const int v = valid ? ld_hint(al + i, pol_s) : 0;
      float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int ri = 0; ri < a.n_rel; ++ri) {
        const MRel& R = a.rel[ri];
        const uint32_t* rp = R.rowptr_c + (int64_t)t * (a.nd + 1);
        uint32_t e0 = 0, cnt = 0;
        if (valid) {
          e0 = ld_hint(rp + i, pol_s);
          cnt = ld_hint(rp + i + 1, pol_s) - e0;
        }
        const uint32_t maxcnt = __reduce_max_sync(0xffffffffu, cnt);
        const bool gcn = R.kind == XPGNN_CONV_GCN;
        if (maxcnt == 0 && !gcn) continue;  // warp uniform
        const int32_t* cc = R.ccol + R.slot_base[t];
        const float* z_c = R.z + (int64_t)t * R.z_s_stride + (int64_t)c * R.z_chunk_stride + sub * 4;
        const float* wg = R.wgt ? R.wgt + (int64_t)t * a.nd : nullptr;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t base = 0; base < maxcnt; base += G) {
          int my = -1;
          float myw = 1.0f;
          if (base + sub < cnt) {
            my = ld_hint(cc + e0 + base + sub, pol_s);
            if (wg) myw = __ldg(wg + my);
          }
#pragma unroll
          for (int j = 0; j < G; ++j) {
            const int u = __shfl_sync(0xffffffffu, my, grp * G + j);
            const float wj = __shfl_sync(0xffffffffu, myw, grp * G + j);
            if (u >= 0) {
              const float4 x = ld_hint4(z_c + (int64_t)u * CW, pol_g);
              acc.x = fmaf(wj, x.x, acc.x); acc.y = fmaf(wj, x.y, acc.y);
              acc.z = fmaf(wj, x.z, acc.z); acc.w = fmaf(wj, x.w, acc.w);
            }
          }
        }
        if (!valid) continue;
        if (gcn) {  // dinv_v (sum_u dinv_u Z[u] + dinv_v Z[v])
          const float dinv = gcn_dinv(cnt);
          const float4 self = ld_hint4(z_c + (int64_t)v * CW, pol_g);
          tot.x += dinv * fmaf(dinv, self.x, acc.x); tot.y += dinv * fmaf(dinv, self.y, acc.y);
          tot.z += dinv * fmaf(dinv, self.z, acc.z); tot.w += dinv * fmaf(dinv, self.w, acc.w);
        } else {
          const float inv = 1.0f / (float)max(cnt, 1u);
          tot.x = fmaf(inv, acc.x, tot.x); tot.y = fmaf(inv, acc.y, tot.y);
          tot.z = fmaf(inv, acc.z, tot.z); tot.w = fmaf(inv, acc.w, tot.w);
        }
      }
      if (!valid) continue;
      tot.x += bias.x; tot.y += bias.y; tot.z += bias.z; tot.w += bias.w;
      if (a.zroot) {
        const float4 r = ld_hint4(a.zroot + (int64_t)t * a.zr_s_stride + (int64_t)c * a.zr_chunk_stride + (int64_t)v * CW + sub * 4, pol_s);
        tot.x += r.x; tot.y += r.y; tot.z += r.z; tot.w += r.w;
      }
      tot.x = apply_act(tot.x, a.act_fn); tot.y = apply_act(tot.y, a.act_fn);
      tot.z = apply_act(tot.z, a.act_fn); tot.w = apply_act(tot.w, a.act_fn);
      st_hint4(a.out + (int64_t)t * a.out_s_stride + (int64_t)c * a.out_chunk_stride + (int64_t)v * CW + sub * 4, tot, pol_s);
    }
  }
}

// bf16 activation storage: the same list-driven pass over 64-element (128-byte) chunks of bf16 rows; 8 lanes x 8
// elements per row, fp32 accumulation, aggregate written back as bf16 (layers >= 1 only: no weights, addend or activation).
__device__ __forceinline__ void add_bf16x8(const uint4& q, float (&acc)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    acc[2 * i] += f.x;
    acc[2 * i + 1] += f.y;
  }
}

template <int OCC>
__global__ void __launch_bounds__(256, OCC) cspmm16_kernel(const CspmmArgs a) {
  __shared__ int s_start[33];
  __shared__ int s_item;
  if (threadIdx.x <= a.nb) s_start[threadIdx.x] = a.slot_tile_start[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7, grp = lane >> 3;
  const int total = s_start[a.nb] * a.n_chunks;
  const __nv_bfloat16* in16 = reinterpret_cast<const __nv_bfloat16*>(a.in);
  __nv_bfloat16* out16 = reinterpret_cast<__nv_bfloat16*>(a.out);
  int t = 0;
  while (true) {
    if (threadIdx.x == 0) s_item = atomicAdd(a.counter, 1);
    __syncthreads();
    const int idx = s_item;
    __syncthreads();
    if (idx >= total) break;
    while (idx >= s_start[t + 1] * a.n_chunks) ++t;
    const int ntb = s_start[t + 1] - s_start[t];
    const int rem = idx - s_start[t] * a.n_chunks;
    const int c = rem / ntb, tb = rem - c * ntb;
    const int n_act = a.slot_info[t].x;
    const int32_t* al = a.act_list + (int64_t)t * a.N;
    const uint32_t* rp = a.rowptr_c + (int64_t)t * (a.N + 1);
    const int32_t* cc = a.ccol + a.slot_base[t];
    const __nv_bfloat16* in_c = in16 + (int64_t)t * a.in_s_stride + (int64_t)c * a.in_chunk_stride + sub * 8;
    __nv_bfloat16* out_c = out16 + (int64_t)t * a.out_s_stride + (int64_t)c * a.out_chunk_stride + sub * 8;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
      const int i = tb * 128 + it * 32 + warp * 4 + grp;
      const bool valid = i < n_act;
      int v = 0;
      uint32_t e0 = 0, cnt = 0;
      if (valid) {
        v = __ldcs(al + i);
        e0 = __ldcs(rp + i);
        cnt = __ldcs(rp + i + 1) - e0;
      }
      const uint32_t maxcnt = __reduce_max_sync(0xffffffffu, cnt);
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
      for (uint32_t base = 0; base < maxcnt; base += 8) {
        const int my = base + sub < cnt ? __ldcs(cc + e0 + base + sub) : -1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int u = __shfl_sync(0xffffffffu, my, grp * 8 + j);
          if (u >= 0) add_bf16x8(__ldg(reinterpret_cast<const uint4*>(in_c + (int64_t)u * 64)), acc);
        }
      }
      if (valid) {
        float scale;
        if (a.kind == XPGNN_CONV_GCN) {  // operands are already scaled by deg^-1/2: agg = dinv (sum + self)
          scale = gcn_dinv(cnt);
          add_bf16x8(__ldg(reinterpret_cast<const uint4*>(in_c + (int64_t)v * 64)), acc);
        } else {
          scale = 1.0f / (float)max(cnt, 1u);
        }
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) oh[k] = __floats2bfloat162_rn(acc[2 * k] * scale, acc[2 * k + 1] * scale);
        __stcs(reinterpret_cast<uint4*>(out_c + (int64_t)v * 64), o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// head on the query rows; an inactive query takes the isolated-node chain (coalition invariant)
// ------------------------------------------------------------------------------------------
struct CHeadArgs {
  int n_iso;                        // conv layers >= 1 of the isolated chain: x = act(W x + b)
  xpgnn_dense_t iso[kMaxConvIso];
  int kind0, act0, h0;              // layer 0 of the isolated chain: GCN act(R0[q] + Z[q]) | SAGE act(R0[q])
  const float* zc;                  // Z = X W^T, chunk-major or (z_rowmajor) [N][h0]
  int z_rowmajor;
  const float* r0c;                 // SAGE: b + X W_root^T (chunk-major) | NULL
  const float* bias0;               // GCN: bias vector | NULL
  int64_t z_chunk_stride;
  int n_head;
  xpgnn_dense_t head[kMaxHeadC];
  const float* in;                  // last conv output, chunk-major
  int64_t in_s_stride, in_chunk_stride;
  int dim0, cw, cw_lg, in16;        // in16: bf16 activations in 64-element chunks
  const float* iso_out;             // hetero: isolated value of the last conv layer per query [n_query][dim0] (precomputed)
  const unsigned long long* tile_active;  // zero-edge rule (model.py:213-215): active in-edges of the slot over all relations
  const int32_t* query;
  int n_query, out_col;
  float* y;                         // y[slot * n_query + q]
  const uint32_t* act;
  int W, w, b0;
};

__device__ __forceinline__ void head_dense(const xpgnn_dense_t& L, const float* x0, float* x1) {
  for (int n = threadIdx.x; n < L.out; n += blockDim.x) {
    float s = 0.0f;
    const float* wr = L.w + (int64_t)n * L.in;
    if (L.w)
      for (int k = 0; k < L.in; ++k) s = fmaf(x0[k], __ldg(wr + k), s);
    if (L.b) s += __ldg(L.b + n);
    x1[n] = apply_act(s, L.act);
  }
}

__global__ void __launch_bounds__(128) compact_head_kernel(const CHeadArgs a, int max_dim) {
  extern __shared__ float sm[];
  float* x0 = sm;
  float* x1 = sm + max_dim;
  const int slot = blockIdx.x / a.n_query, q = blockIdx.x % a.n_query;
  const int qv = a.query[q];
  const bool active = (a.act[(int64_t)qv * a.W + a.w] >> (a.b0 + slot)) & 1u;
  if (active) {
    if (a.in16) {
      const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(a.in) + (int64_t)slot * a.in_s_stride + (int64_t)qv * 64;
      for (int i = threadIdx.x; i < a.dim0; i += blockDim.x) x0[i] = __bfloat162float(src[(int64_t)(i >> 6) * a.in_chunk_stride + (i & 63)]);
    } else {
      const float* src = a.in + (int64_t)slot * a.in_s_stride + (int64_t)qv * a.cw;
      for (int i = threadIdx.x; i < a.dim0; i += blockDim.x) x0[i] = src[(int64_t)(i >> a.cw_lg) * a.in_chunk_stride + (i & (a.cw - 1))];
    }
    __syncthreads();
  } else if (a.iso_out) {
    for (int i = threadIdx.x; i < a.dim0; i += blockDim.x) x0[i] = a.iso_out[(int64_t)q * a.dim0 + i];
    __syncthreads();
  } else {
    for (int i = threadIdx.x; i < a.h0; i += blockDim.x) {
      const int64_t off = (int64_t)(i >> a.cw_lg) * a.z_chunk_stride + (int64_t)qv * a.cw + (i & (a.cw - 1));
      const float zq = a.kind0 == XPGNN_CONV_GCN ? (a.z_rowmajor ? a.zc[(int64_t)qv * a.h0 + i] : a.zc[off]) : 0.0f;
      const float r = (a.r0c ? a.r0c[off] : 0.0f) + (a.bias0 ? a.bias0[i] : 0.0f) + zq;
      x0[i] = apply_act(r, a.act0);
    }
    __syncthreads();
    for (int li = 0; li < a.n_iso; ++li) {
      head_dense(a.iso[li], x0, x1);
      __syncthreads();
      float* tp = x0; x0 = x1; x1 = tp;
    }
  }
  for (int li = 0; li < a.n_head; ++li) {
    head_dense(a.head[li], x0, x1);
    __syncthreads();
    float* tp = x0; x0 = x1; x1 = tp;
  }
  if (threadIdx.x == 0) {
    float r = x0[a.out_col];
    if (a.tile_active && a.tile_active[slot] == 0ull) r = 0.0f;
    a.y[(int64_t)slot * a.n_query + q] = r;
  }
}

// ------------------------------------------------------------------------------------------ host
static bool compact_enabled() { return knobs().compact != 0; }

static int compact_cw(const xpgnn_plan_t* p) {
  int cw = knobs().cw == 16 ? 16 : 32;
  for (int l = 0; l < p->n_layers; ++l)
    if (p->layers_host[l].h_out % cw) cw = 16;
  return cw;
}

bool compact_eligible(const xpgnn_plan_t* p) {
  if (!compact_enabled() || p->prune || p->zero_edge_rule || p->n_layers < 1 || p->n_layers > std::min(kMaxConvIso, 16)) return false;
  if (p->n_head > kMaxHeadC || p->n_nodes >= (1 << kPackShift)) return false;
  const xpgnn_relation_t& R0 = p->layers_host[0].rel_host[0];
  for (int l = 0; l < p->n_layers; ++l) {
    const xpgnn_layer_t& L = p->layers_host[l];
    if (L.n_rel != 1 || L.h_out % 16) return false;
    const xpgnn_relation_t& R = L.rel_host[0];
    if (R.src_lo != 0 || R.src_hi != p->n_nodes || R.dst_lo != 0 || R.dst_hi != p->n_nodes) return false;
    if (R.rowptr != R0.rowptr || R.col != R0.col || R.conv_kind != R0.conv_kind) return false;
    if (l > 0 && L.h_in != p->layers_host[l - 1].h_out) return false;
  }
  return true;
}

// bf16 activation storage (plan precision 2): GCN stacks whose widths are multiples of 64 (one bf16 chunk = 64 elements);
// anything else keeps fp32 storage with bf16 tensor-core transforms (precision 1 semantics)
static bool compact_act16(const xpgnn_plan_t* p) {
  if (p->precision != 2 || p->layers_host[0].rel_host[0].conv_kind != XPGNN_CONV_GCN) return false;
  for (int l = 0; l < p->n_layers; ++l)
    if (p->layers_host[l].h_out % 64 || p->layers_host[l].h_out > 256) return false;
  return !knobs().l0_lists;
}

struct CLayout {
  float *zc, *r0c;
  uint32_t* ebits;
  unsigned long long *keys, *scanned;
  void* cub_tmp;
  size_t cub_bytes;
  int32_t* act_list;
  uint32_t* rowptr_c;
  float* wgt;
  float* scale;
  int2* slot_info;
  long long* slot_base;
  int32_t *slot_tile_start, *n_tiles, *counters;
  int32_t *long_rows, *n_long;
  int32_t *long_list, *n_long_list;
  int long_cap;  // entries per slot segment of long_list
  int32_t *item_row, *item_slice, *row_item0, *n_items;  // sliced hub rows of layer 0 (compact_l0.cu)
  uint32_t* act_word;  // [N] the tile's word of the coalition bits (extract_word_kernel)
  float2* slice_scratch;
  int32_t* rows_packed;
  float* rs_packed;
  int32_t* ccol;
  float *hbuf[2], *agg;
  int64_t bytes;
};

static CLayout compact_carve(const xpgnn_plan_t* p, void* ws, int64_t cap, int tile) {
  CLayout c{};
  Bump b(ws, cap);
  const int64_t N = p->n_nodes, E = std::max(p->layers_host[0].rel_host[0].n_edges, 1);
  const int NL = p->n_layers;
  int hmax = 0;
  for (int l = 0; l < NL; ++l) hmax = std::max(hmax, p->layers_host[l].h_out);
  const int h0 = p->layers_host[0].h_out;
  c.zc = b.take<float>(N * h0);
  c.r0c = b.take<float>(N * h0);
  c.ebits = b.take<uint32_t>(E);
  c.keys = b.take<unsigned long long>((int64_t)tile * N + 1);
  c.scanned = b.take<unsigned long long>((int64_t)tile * N + 1);
  c.cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, c.cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int64_t)tile * N + 1);
  c.cub_tmp = b.take<char>((int64_t)c.cub_bytes + 256);
  c.act_list = b.take<int32_t>((int64_t)tile * N);
  c.rowptr_c = b.take<uint32_t>((int64_t)tile * (N + 1));
  c.wgt = b.take<float>((int64_t)tile * N);
  c.scale = b.take<float>(N * 32);
  c.slot_info = b.take<int2>(32);
  c.slot_base = b.take<long long>(32);
  c.slot_tile_start = b.take<int32_t>(33);
  c.n_tiles = b.take<int32_t>(1);
  c.counters = b.take<int32_t>(32);
  c.long_rows = b.take<int32_t>(E / kLongRow + 1);
  c.n_long = b.take<int32_t>(1);
  c.item_row = b.take<int32_t>(l0_long_items_max(E));
  c.item_slice = b.take<int32_t>(l0_long_items_max(E));
  c.row_item0 = b.take<int32_t>(E / kLongRow + 2);
  c.n_items = b.take<int32_t>(1);
  c.act_word = b.take<uint32_t>(N);
  c.slice_scratch = b.take<float2>(l0_slice_scratch_items(E) * (h0 / 64 + 1) * 1024);
  c.long_cap = (int)(E / kLongCompact + 1);
  c.long_list = b.take<int32_t>((int64_t)tile * c.long_cap);
  c.n_long_list = b.take<int32_t>(32);
  c.rows_packed = b.take<int32_t>((int64_t)tile * ceil_div(N, 128) * 128);
  c.rs_packed = b.take<float>((int64_t)tile * ceil_div(N, 128) * 128);
  c.ccol = b.take<int32_t>((int64_t)tile * E);
  const int64_t act_floats = compact_act16(p) ? ((int64_t)tile * N * hmax + 1) / 2 : (int64_t)tile * N * hmax;  // bf16: half
  c.hbuf[0] = b.take<float>(act_floats);
  if (NL > 1) {
    c.hbuf[1] = b.take<float>(act_floats);
    c.agg = b.take<float>(act_floats);
  }
  c.bytes = (b.off + 255) & ~255ll;
  return c;
}

int64_t compact_workspace_bytes(const xpgnn_plan_t* p, int tile) { return compact_carve(p, nullptr, 0, tile).bytes; }

static int launch_cspmm(const CspmmArgs& a, int cw, cudaStream_t st) {
  const int occ = knobs().occ;
  void (*k)(const CspmmArgs);
  // aggregate-first layers >= 1 (plain scaled sums over 32-float chunks): the segmented kernel
  int seg = knobs().seg;  // 0: row-lockstep kernel | 4 / 6 / 8: gathers in flight per lane
  if ((seg == 7 || seg == 8) && a.long_cnt > 0 && knobs().seg_skew > 0) seg = knobs().seg_skew;  // hub rows in the graph: skewed degrees
  if (seg > 0 && cw == 32 && a.counter && !a.wgt && !a.addend && !a.bias && !a.layer0 && !a.prescale && a.act_fn == XPGNN_ACT_NONE) {
    const int socc = knobs().seg_occ;
    if (seg >= 200) {  // warp-specialised TMA bulk-copy variant (compact_bulk.cu): seg = 200 + 100 * mode + ring stages (16 | 32): 216 / 232 bulk copies, 316 / 332 cp.async.cg, 432 cp.async.ca, 516 / 532 TMA gather4
      ProfScope ps(a.prof_cat > 0 ? a.prof_cat : PROF_SPMM_TILE, st);
      if (launch_cspmm_bulk(a, seg - 200, st)) return 1;
      if (a.long_cnt > 0) XP_LAUNCH((cspmm_long_kernel<32, false>), kNumSMs * 8, 256, 0, st, a);
      return 0;
    }
    if (seg >= 100) {  // shared-memory ring variant: seg = 100 + slots per group (116 | 124)
      void (*kr)(const CspmmArgs) = seg >= 124 ? cspmm_ring_kernel<24, 4> : cspmm_ring_kernel<16, 5>;
      const int smem = seg >= 124 ? (int)sizeof(RingSmem<24>) : (int)sizeof(RingSmem<16>);
      XP_CHECK(cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      int per_sm = 0;
      XP_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kr, 32 * kRingWarps, smem));
      ProfScope ps(a.prof_cat > 0 ? a.prof_cat : PROF_SPMM_TILE, st);
      XP_LAUNCH(kr, kNumSMs * std::max(per_sm, 1), 32 * kRingWarps, smem, st, a);
      if (a.long_cnt > 0) XP_LAUNCH((cspmm_long_kernel<32, false>), kNumSMs * 8, 256, 0, st, a);
      return 0;
    }
    const int pf = knobs().seg_pf;
    if (seg >= 16) k = cspmm_seg_kernel<16, 4, 0>;       // 16 warps / SM x 16 gathers in flight per lane (128 registers)
    else if (seg >= 12) k = cspmm_seg_kernel<12, 5, 0>;  // 20 warps / SM x 12
    else if (seg == 7) k = cspmm_seg_kernel<7, 7, 0>;     // 28 warps / SM x 7 gathers in flight (73 registers at most)
    else if (seg >= 8) {
      if (socc == 8) k = cspmm_seg_kernel<8, 8, 0>;
      else if (socc == 7) k = cspmm_seg_kernel<8, 7, 0>;
      else k = pf >= 16 ? cspmm_seg_kernel<8, 6, 16> : (pf >= 12 ? cspmm_seg_kernel<8, 6, 12> : (pf >= 8 ? cspmm_seg_kernel<8, 6, 8> : cspmm_seg_kernel<8, 6, 0>));
    } else if (seg >= 6) k = socc == 6 ? cspmm_seg_kernel<6, 6, 0> : (socc == 7 ? cspmm_seg_kernel<6, 7, 0> : cspmm_seg_kernel<6, 8, 0>);
    else k = socc == 10 ? cspmm_seg_kernel<4, 10, 0> : (pf >= 8 ? cspmm_seg_kernel<4, 8, 8> : cspmm_seg_kernel<4, 8, 0>);
    if (knobs().seg_carve > 0)  // experiment: shared-memory carve-out in percent (what is left of the 256 KB is L1)
      XP_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, knobs().seg_carve));
    int per_sm = 0;
    XP_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 128, 0));
    ProfScope ps(a.prof_cat > 0 ? a.prof_cat : PROF_SPMM_TILE, st);
    if (knobs().seg_tma) {
      // Hybrid: the small TMA gather4 kernel goes first on a second stream (one CTA per SM), the register-queue kernel fills
      // the rest of every SM; both take 32-row blocks from a.counter in order, so the TMA unit's bytes per cycle add to the
      // LSU path's.  Fork / join with events: everything after this call on `st` sees both kernels finished.
      static cudaStream_t aux = nullptr;
      static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
      if (!aux) {
        int prio_lo = 0, prio_hi = 0;  // highest priority: its CTAs are placed before the register-queue kernel's
        XP_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        XP_CHECK(cudaStreamCreateWithPriority(&aux, cudaStreamNonBlocking, prio_hi));
        XP_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        XP_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
      }
      XP_CHECK(cudaEventRecord(ev_fork, st));
      XP_CHECK(cudaStreamWaitEvent(aux, ev_fork, 0));
      if (launch_cspmm_bulk(a, 400 + 16, aux)) return 1;
      XP_LAUNCH(k, kNumSMs * std::max(per_sm - 1, 1), 128, 0, st, a);  // one CTA less per SM: room for the gather4 kernel's registers
      XP_CHECK(cudaEventRecord(ev_join, aux));
      XP_CHECK(cudaStreamWaitEvent(st, ev_join, 0));
    } else {
      XP_LAUNCH(k, kNumSMs * std::max(per_sm, 1), 128, 0, st, a);
    }
    if (a.long_cnt > 0) XP_LAUNCH((cspmm_long_kernel<32, false>), kNumSMs * 8, 256, 0, st, a);
    return 0;
  }
  if (occ >= 8)
    k = cw == 32 ? (a.wgt ? cspmm_kernel<32, true, 8> : cspmm_kernel<32, false, 8>) : (a.wgt ? cspmm_kernel<16, true, 8> : cspmm_kernel<16, false, 8>);
  else
    k = cw == 32 ? (a.wgt ? cspmm_kernel<32, true, 6> : cspmm_kernel<32, false, 6>) : (a.wgt ? cspmm_kernel<16, true, 6> : cspmm_kernel<16, false, 6>);
  // every CTA must be resident: the round-robin deal of the items keeps the whole grid inside one
  // (slot, chunk) pass only if no CTA starts late
  int per_sm = 0;
  XP_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 256, 0));
  const int grid = kNumSMs * std::max(per_sm, 1);
  ProfScope ps(a.prof_cat > 0 ? a.prof_cat : (a.layer0 ? PROF_SPMM_INVARIANT : PROF_SPMM_TILE), st);
  XP_LAUNCH(k, grid, 256, 0, st, a);
  if (a.long_cnt > 0) {
    void (*kl)(const CspmmArgs) = cw == 32 ? (a.wgt ? cspmm_long_kernel<32, true> : cspmm_long_kernel<32, false>)
                                           : (a.wgt ? cspmm_long_kernel<16, true> : cspmm_long_kernel<16, false>);
    XP_LAUNCH(kl, kNumSMs * 8, 256, 0, st, a);
  }
  return 0;
}

int forward_compact(const xpgnn_plan_t* p, const uint32_t* act_in, int32_t W_in, int32_t s0, int32_t n_s, float* y, void* workspace,
                    int64_t workspace_bytes, int64_t* stats, cudaStream_t st, int dense_prec) {
  const int N = p->n_nodes, NL = p->n_layers;
  const xpgnn_relation_t& R0 = p->layers_host[0].rel_host[0];
  const int kind = R0.conv_kind;
  const int cw = compact_cw(p), cw_lg = cw == 32 ? 5 : 4;
  int tile = 32;
  while (tile > 1 && compact_carve(p, nullptr, 0, tile).bytes > workspace_bytes) tile >>= 1;
  XP_REQUIRE(compact_carve(p, nullptr, 0, tile).bytes <= workspace_bytes, "workspace too small even for one coalition per tile");
  XP_REQUIRE((int64_t)tile * std::max(R0.n_edges, 1) < (1ll << kKeyShift), "tile x edges exceeds the packed scan key");
  CLayout lay = compact_carve(p, workspace, workspace_bytes, tile);
  int hmax = 0;
  for (int l = 0; l < NL; ++l) hmax = std::max(hmax, p->layers_host[l].h_out);
  const int64_t hstride = (int64_t)N * hmax;       // per-slot stride of the activation buffers
  const int64_t cstride = (int64_t)N * cw;         // chunk stride

  // ---- coalition-invariant part of layer 0: Z = X W^T, R0 = b (+ X W_root^T for SAGE) ----
  const xpgnn_layer_t& L0 = p->layers_host[0];
  XP_REQUIRE(L0.h_in == p->f_in, "layer 0 input width != feature width");
  // row-outer layer 0 (Z row-major) when a warp can own 128 columns; else the list-driven kernel (Z chunk-major)
  const bool l0_lists = knobs().l0_lists != 0;
  const bool l0_rows = !l0_lists && cw == 32 && L0.h_out % 64 == 0;
  const bool act16 = compact_act16(p) && l0_rows;   // bf16 activation storage
  const int64_t cstride16 = (int64_t)N * 64;        // chunk stride of the bf16 layout (elements)
  {
    DenseArgs z{};
    z.in = p->x; z.ld_in = p->f_in; z.k = p->f_in; z.w = R0.w_nbr; z.n_out = L0.h_out; z.out = lay.zc;
    z.ld_out = cw; z.cw_out = cw; z.cw_out_lg = cw_lg; z.out_chunk_stride = cstride;
    z.rows_per_s = N; z.row_lo = 0; z.M = N; z.dst_lo = 0; z.dst_hi = N;
    DenseArgs zz = z;
    if (l0_rows) { zz.ld_out = L0.h_out; zz.cw_out = 0; zz.cw_out_lg = 0; zz.out_chunk_stride = 0; }
    if (launch_dense(zz, st, dense_prec)) return 1;
    if (kind == XPGNN_CONV_SAGE_MEAN) {  // GCN only adds its bias vector (in the SpMM epilogue)
      const bool sage_root = R0.w_root != nullptr;
      DenseArgs rt = z;  // k = 0 degenerates to "write the bias"
      rt.k = sage_root ? p->f_in : 0; rt.w = sage_root ? R0.w_root : R0.w_nbr; rt.b = R0.b_nbr; rt.out = lay.r0c;
      if (launch_dense(rt, st, dense_prec)) return 1;
    }
  }
  const int l2_stream = knobs().l2_stream;
  const int l2_gather = knobs().l2_gather;
  const bool dyn_sched = !knobs().sched_static;
  XP_CHECK(cudaMemsetAsync(lay.keys + (int64_t)tile * N, 0, sizeof(unsigned long long), st));
  // hub rows get CTA-per-row variants of the row-per-warp kernels
  int n_long = 0, n_l0_items = 0;
  XP_CHECK(cudaMemsetAsync(lay.counters, 0, 32 * sizeof(int32_t), st));  // later zeroed per tile by compact_tilemap_kernel
  XP_CHECK(cudaMemsetAsync(lay.n_long, 0, sizeof(int32_t), st));
  XP_LAUNCH(find_long_rows_kernel, (int)ceil_div(N, 256), 256, 0, st, R0.rowptr, N, kLongRow, lay.long_rows, lay.n_long);
  // layer 0: hub rows in slices of kL0Slice in-edges, one CTA each (compact_l0.cu)
  if (build_l0_long_items(R0.rowptr, lay.long_rows, lay.n_long, lay.item_row, lay.item_slice, lay.row_item0, lay.n_items, st)) return 1;
  XP_CHECK(cudaMemcpyAsync(&n_long, lay.n_long, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  XP_CHECK(cudaMemcpyAsync(&n_l0_items, lay.n_items, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  XP_CHECK(cudaStreamSynchronize(st));
  if (!knobs().long_rows) n_long = 0;
  const bool l0_slices = knobs().l0_slices != 0 && p->layers_host[0].h_out / 64 <= 4 &&
                         n_l0_items <= l0_slice_scratch_items(std::max(R0.n_edges, 1));

  const int w_first = s0 / 32, w_last = (s0 + n_s - 1) / 32;
  const int grid_rows = (int)std::min<int64_t>(std::max<int64_t>(ceil_div(N, 8), 1), (int64_t)kNumSMs * 8);
  for (int w_idx = w_first; w_idx <= w_last; ++w_idx) {
    const int bits_in_word = std::min(32, s0 + n_s - w_idx * 32);
    // the kernels of this word read the bits as act[v * W + w]: from the caller's matrix, or from the 4 MB column of this word
    const uint32_t* act = act_in;
    int W = W_in, w = w_idx;
    if (W_in > 1 && knobs().act_column) {
      XP_LAUNCH(extract_word_kernel, (int)ceil_div(N, 256), 256, 0, st, act_in, W_in, w_idx, (int)N, lay.act_word);
      act = lay.act_word; W = 1; w = 0;
    }
    for (int b0 = 0; b0 < bits_in_word; b0 += tile) {
      const int nb = std::min(tile, bits_in_word - b0);
      // ---- per-tile compaction ----
      {
        ProfScope ps(PROF_SCALE, st);
        XP_LAUNCH(compact_degree_kernel, grid_rows, 256, 0, st, R0.rowptr, R0.col, act, W, w, b0, nb, N, lay.ebits, lay.keys,
                  l0_rows ? lay.scale : nullptr, kind, lay.counters + 15, n_long > 0 ? kLongRow : 0, 0);
        if (n_long > 0)
          XP_LAUNCH(compact_degree_long_kernel, n_long, 256, 0, st, R0.rowptr, R0.col, act, W, w, b0, nb, N, lay.ebits, lay.keys,
                    l0_rows ? lay.scale : nullptr, kind, lay.long_rows, 0);
      }
      {
        ProfScope ps(PROF_COMPACT, st);
        if (nb < tile) XP_CHECK(cudaMemsetAsync(lay.keys + (int64_t)nb * N, 0, sizeof(unsigned long long), st));
        XP_CHECK(cudaMemsetAsync(lay.n_long_list, 0, 32 * sizeof(int32_t), st));
        size_t tmp = lay.cub_bytes;
        XP_CHECK(cub::DeviceScan::ExclusiveSum(lay.cub_tmp, tmp, lay.keys, lay.scanned, (int64_t)nb * N + 1, st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        XP_LAUNCH(compact_finalize_kernel, dim3((unsigned)ceil_div(N, 256), (unsigned)nb), 256, 0, st, lay.keys, lay.scanned, N, 0,
                  lay.act_list, lay.rowptr_c, (kind == XPGNN_CONV_GCN && !l0_rows) ? lay.wgt : nullptr, lay.slot_info, lay.slot_base,
                  lay.rows_packed, lay.rs_packed, n_long > 0 ? kLongCompact : 0, lay.long_list, lay.n_long_list, lay.long_cap);
        XP_LAUNCH(compact_tilemap_kernel, 1, 256, 0, st, lay.slot_info, nb, lay.slot_tile_start, lay.n_tiles, NL, stats, lay.counters, 1);
        XP_LAUNCH(compact_edges_kernel, grid_rows, 256, 0, st, R0.rowptr, R0.col, lay.ebits, act, W, w, b0, nb, N, lay.scanned, lay.ccol,
                  lay.counters + 14, n_long > 0 ? kLongRow : 0, 0);
        if (n_long > 0)
          XP_LAUNCH(compact_edges_long_kernel, n_long, 256, 0, st, R0.rowptr, R0.col, lay.ebits, act, W, w, b0, nb, N, lay.scanned,
                    lay.ccol, lay.long_rows, 0);
      }
      float* cur = lay.hbuf[0];
      float* nxt = lay.hbuf[1];
      for (int l = 0; l < NL; ++l) {
        const xpgnn_layer_t& L = p->layers_host[l];
        const xpgnn_relation_t& R = L.rel_host[0];
        const bool next_gcn = (l + 1 < NL) && kind == XPGNN_CONV_GCN;
        CspmmArgs s{};
        s.nb = nb; s.N = N; s.kind = kind; s.layer0 = l == 0;
        s.slot_info = lay.slot_info; s.slot_tile_start = lay.slot_tile_start; s.act_list = lay.act_list;
        s.rowptr_c = lay.rowptr_c; s.slot_base = lay.slot_base; s.ccol = lay.ccol;
        s.in_chunk_stride = cstride; s.out_s_stride = hstride; s.out_chunk_stride = cstride;
        s.counter = dyn_sched ? lay.counters + l : nullptr;
        s.l2_stream = l2_stream; s.l2_gather = l2_gather;
        s.long_cnt = n_long > 0 ? kLongCompact : 0; s.long_list = lay.long_list; s.n_long_list = lay.n_long_list;
        s.long_cap = lay.long_cap; s.counter_long = dyn_sched ? lay.counters + 16 + std::min(l, 12) : nullptr;
        if (l == 0 && l0_rows) {
          L0RowsArgs r{};
          r.rowptr = R.rowptr; r.col = R.col; r.ebits = lay.ebits; r.act = act; r.W = W; r.w = w; r.b0 = b0; r.nb = nb; r.N = N;
          r.scale = lay.scale; r.z = lay.zc; r.h0 = L.h_out;
          if (kind == XPGNN_CONV_SAGE_MEAN) { r.r0c = lay.r0c; r.r0_chunk_stride = cstride; }
          else r.bias = R.b_nbr;
          r.out = cur; r.out_s_stride = hstride; r.out_chunk_stride = cstride; r.out_row_stride = 32; r.r0_row_stride = 32;
          r.kind = kind; r.act_fn = L.act; r.prescale = next_gcn;
          ProfScope ps(PROF_SPMM_INVARIANT, st);
          r.long_rows = lay.long_rows; r.long_threshold = n_long > 0 ? kLongRow : 0; r.counter = lay.counters + 13;
          r.row_lo = 0; r.row_hi = N; r.accumulate = 0; r.finish = 1;
          const bool sg = L.act == XPGNN_ACT_SIGMOID;
          if (act16) r.out_chunk_stride = cstride16;
          if (launch_l0_rows(r, sg, act16, N, st)) return 1;
          if (n_long > 0) {
            if (l0_slices) {
              r.item_row = lay.item_row; r.item_slice = lay.item_slice; r.row_item0 = lay.row_item0; r.slice_scratch = lay.slice_scratch;
            }
            if (launch_l0_long_rows(r, sg, act16, n_long, st, n_l0_items)) return 1;
          }
        } else if (l == 0) {  // transform-first: gather the coalition-invariant Z
          s.n_chunks = L.h_out / cw;
          s.in = lay.zc; s.in_s_stride = 0; s.wgt = kind == XPGNN_CONV_GCN ? lay.wgt : nullptr;
          if (kind == XPGNN_CONV_SAGE_MEAN) { s.addend = lay.r0c; s.add_chunk_stride = cstride; }
          else s.bias = R.b_nbr;
          s.out = cur; s.act_fn = L.act; s.prescale = next_gcn;
          if (launch_cspmm(s, cw, st)) return 1;
        } else if (act16) {  // bf16 storage: aggregate (bf16 in, fp32 accumulate, bf16 out), then the bf16 tensor-core transform
          s.n_chunks = L.h_in / 64;
          s.in = cur; s.in_s_stride = hstride; s.in_chunk_stride = cstride16;
          s.out = lay.agg; s.out_chunk_stride = cstride16;
          {
            ProfScope ps(PROF_SPMM_TILE, st);
            const int occ16 = knobs().occ16;
            void (*k16)(const CspmmArgs) = occ16 >= 8 ? cspmm16_kernel<8> : (occ16 <= 4 ? cspmm16_kernel<4> : cspmm16_kernel<6>);
            int per_sm = 0;
            XP_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k16, 256, 0));
            XP_LAUNCH(k16, kNumSMs * std::max(per_sm, 1), 256, 0, st, s);
          }
          DenseArgs d{};
          d.in = lay.agg; d.in_s_stride = hstride; d.ld_in = 64; d.k = L.h_in; d.cw_in = 64; d.cw_in_lg = 6; d.in_chunk_stride = cstride16;
          d.w = R.w_nbr; d.b = R.b_nbr; d.n_out = L.h_out;
          d.out = nxt; d.out_s_stride = hstride; d.ld_out = 64; d.cw_out = 64; d.cw_out_lg = 6; d.out_chunk_stride = cstride16;
          d.rows_packed = lay.rows_packed; d.n_tiles_dev = lay.n_tiles;
          d.rows_per_s = N; d.M = (int64_t)nb * ceil_div(N, 128) * 128; d.dst_lo = 0; d.dst_hi = N;
          d.act_fn = L.act; d.rs_packed = next_gcn ? lay.rs_packed : nullptr;
          d.in16 = 1; d.out16 = 1;
          if (launch_dense(d, st, DENSE_TC_BF16)) return 1;
          std::swap(cur, nxt);
        } else {       // aggregate-first, then the dense transform of the active rows
          s.n_chunks = L.h_in / cw;
          s.in = cur; s.in_s_stride = hstride; s.wgt = nullptr; s.addend = nullptr;
          s.out = lay.agg; s.act_fn = XPGNN_ACT_NONE; s.prescale = 0;
          if (launch_cspmm(s, cw, st)) return 1;
          const bool sage_root = kind == XPGNN_CONV_SAGE_MEAN && R.w_root;
          DenseArgs d{};
          d.in = lay.agg; d.in_s_stride = hstride; d.ld_in = cw; d.k = L.h_in; d.cw_in = cw; d.cw_in_lg = cw_lg; d.in_chunk_stride = cstride;
          d.w = R.w_nbr; d.b = R.b_nbr; d.n_out = L.h_out;
          d.out = nxt; d.out_s_stride = hstride; d.ld_out = cw; d.cw_out = cw; d.cw_out_lg = cw_lg; d.out_chunk_stride = cstride;
          d.rows_packed = lay.rows_packed; d.n_tiles_dev = lay.n_tiles;
          d.rows_per_s = N; d.M = (int64_t)nb * ceil_div(N, 128) * 128; d.dst_lo = 0; d.dst_hi = N;
          d.act_fn = sage_root ? XPGNN_ACT_NONE : L.act; d.rs_packed = (next_gcn && !sage_root) ? lay.rs_packed : nullptr;
          if (launch_dense(d, st, dense_prec)) return 1;
          if (sage_root) {
            DenseArgs rt = d;
            rt.in = cur; rt.w = R.w_root; rt.b = nullptr; rt.accumulate = 1; rt.act_fn = L.act; rt.rs_packed = next_gcn ? lay.rs_packed : nullptr;
            if (launch_dense(rt, st, dense_prec)) return 1;
          }
          std::swap(cur, nxt);
        }
      }
      // ---- head on the query rows ----
      CHeadArgs h{};
      int max_dim = L0.h_out;
      h.n_iso = NL - 1;
      for (int l = 1; l < NL; ++l) {
        const xpgnn_layer_t& L = p->layers_host[l];
        const xpgnn_relation_t& R = L.rel_host[0];
        xpgnn_dense_t& d = h.iso[l - 1];
        d.in = L.h_in; d.out = L.h_out; d.act = L.act; d.b = R.b_nbr;
        d.w = kind == XPGNN_CONV_GCN ? R.w_nbr : R.w_root;  // SAGE: the empty mean contributes nothing
        max_dim = std::max(max_dim, std::max(L.h_in, L.h_out));
      }
      h.kind0 = kind; h.act0 = L0.act; h.h0 = L0.h_out; h.zc = lay.zc; h.z_chunk_stride = cstride; h.z_rowmajor = l0_rows;
      if (kind == XPGNN_CONV_SAGE_MEAN) h.r0c = lay.r0c; else h.bias0 = R0.b_nbr;
      h.n_head = p->n_head;
      for (int i = 0; i < p->n_head; ++i) {
        h.head[i] = p->head_host[i];
        max_dim = std::max(max_dim, std::max(p->head_host[i].in, p->head_host[i].out));
      }
      h.in = cur; h.in_s_stride = hstride; h.in_chunk_stride = cstride; h.dim0 = p->layers_host[NL - 1].h_out; h.cw = cw; h.cw_lg = cw_lg; h.in16 = act16; if (act16) h.in_chunk_stride = cstride16;
      h.query = p->query; h.n_query = p->n_query; h.out_col = p->out_col;
      h.y = y + ((int64_t)(w_idx * 32 + b0) - s0) * p->n_query;
      h.act = act; h.W = W; h.w = w; h.b0 = b0;
      {
        ProfScope ps(PROF_HEAD, st);
        XP_LAUNCH(compact_head_kernel, nb * p->n_query, 128, sizeof(float) * 2 * max_dim, st, h, max_dim);
      }
    }
  }
  return 0;
}


// ==========================================================================================
// Hetero compact path: HeteroConv(sum) stacks of SAGEConv(mean) / same-type GCNConv relations (BASELINE C4).
// Every relation is a bipartite graph with its own CSR, masked degrees and per-coalition compacted lists; the
// relations into one destination type add into the same activation rows (first writes, others add, last
// finishes).  Same kernels as the homogeneous driver above, which stays the path of single-relation models.
// ==========================================================================================
constexpr int kMaxIsoGroups = 8;   // destination groups (node types) per layer in the isolated chain
constexpr int kMaxIsoSelf = 16;    // layer-0 GCN relations whose Z_r[q] joins the isolated value

struct IsoArgs {
  int n_layers, h0, act0;
  const float* r0c;
  int64_t r0_chunk_stride;
  int n_self;
  const float* self_z[kMaxIsoSelf];  // biased row-major Z_r (index by global node id)
  int self_lo[kMaxIsoSelf], self_hi[kMaxIsoSelf];
  int n_groups[kMaxConvIso];
  int g_lo[kMaxConvIso][kMaxIsoGroups], g_hi[kMaxConvIso][kMaxIsoGroups];
  const float* g_w[kMaxConvIso][kMaxIsoGroups];  // sum over the group of (GCN: lin.weight | SAGE: lin_r.weight), may be NULL
  const float* g_b[kMaxConvIso][kMaxIsoGroups];  // sum of the biases, may be NULL
  int h_in[kMaxConvIso], h_out[kMaxConvIso], act[kMaxConvIso];
  const int32_t* query;
  float* iso_out;  // [n_query][h_out of the last layer]
};

// value of a query node in a coalition in which it is inactive (no messages reach it in any layer): coalition invariant
__global__ void __launch_bounds__(128) hetero_iso_kernel(const IsoArgs a, int max_dim) {
  extern __shared__ float sm[];
  float* x0 = sm;
  float* x1 = sm + max_dim;
  const int q = blockIdx.x, qv = a.query[q];
  for (int i = threadIdx.x; i < a.h0; i += blockDim.x) {
    float r = a.r0c[(int64_t)(i >> 5) * a.r0_chunk_stride + (int64_t)qv * 32 + (i & 31)];
    for (int k = 0; k < a.n_self; ++k)
      if (qv >= a.self_lo[k] && qv < a.self_hi[k]) r += a.self_z[k][(int64_t)qv * a.h0 + i];
    x0[i] = apply_act(r, a.act0);
  }
  __syncthreads();
  int dim = a.h0;
  for (int l = 1; l < a.n_layers; ++l) {
    int g = -1;
    for (int k = 0; k < a.n_groups[l]; ++k)
      if (qv >= a.g_lo[l][k] && qv < a.g_hi[l][k]) g = k;
    for (int n = threadIdx.x; n < a.h_out[l]; n += blockDim.x) {
      float acc = 0.0f;
      if (g >= 0) {
        if (a.g_w[l][g]) {
          const float* wr = a.g_w[l][g] + (int64_t)n * a.h_in[l];
          for (int k = 0; k < a.h_in[l]; ++k) acc = fmaf(x0[k], __ldg(wr + k), acc);
        }
        if (a.g_b[l][g]) acc += __ldg(a.g_b[l][g] + n);
      }
      x1[n] = apply_act(acc, a.act[l]);
    }
    __syncthreads();
    float* tp = x0; x0 = x1; x1 = tp;
    dim = a.h_out[l];
  }
  for (int i = threadIdx.x; i < dim; i += blockDim.x) a.iso_out[(int64_t)q * dim + i] = x0[i];
}

__global__ void compact_add_into_kernel(float* __restrict__ acc, const float* __restrict__ x, int n, int first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acc[i] = first ? x[i] : acc[i] + x[i];
}

__global__ void compact_sum_active_kernel(const int2* __restrict__ slot_info, int nb, unsigned long long* __restrict__ tile_active, int first) {
  const int t = threadIdx.x;
  if (t < nb) tile_active[t] = (first ? 0ull : tile_active[t]) + (unsigned long long)slot_info[t].y;
}

struct HCsr {  // one unique relation CSR and its per-tile compaction
  const int32_t *rowptr, *col;
  int kind, lo, hi, n_edges, n_long;
  int n_layers_using;  // conv layers whose relations map to this CSR (active-edge statistics)
  uint32_t* ebits;
  float *scale, *wgt, *rs_packed;
  int32_t *act_list, *slot_tile_start, *n_tiles, *counters, *long_rows, *n_long_dev, *long_list, *n_long_list, *rows_packed, *ccol;
  int long_cap;
  uint32_t* rowptr_c;
  int2* slot_info;
  long long* slot_base;
};

struct HLayout {
  std::vector<HCsr> csr;
  std::vector<std::vector<int>> map;       // [layer][relation] -> csr
  std::vector<float*> zr;                  // layer-0 relations: biased row-major Z_r
  float* r0c;
  unsigned long long *keys, *scanned, *tile_active;
  void* cub_tmp;
  size_t cub_bytes;
  std::vector<std::vector<float*>> wroot;  // [layer >= 1][relation]: summed SAGE root weights of its destination group
  std::vector<std::vector<float*>> iso_w, iso_b;  // [layer >= 1][relation]: merged isolated-chain weights / biases of its group
  float *iso_out, *hbuf[2], *agg;
  int64_t bytes;
};

static bool same_dst(const xpgnn_relation_t& a, const xpgnn_relation_t& b) { return a.dst_lo == b.dst_lo && a.dst_hi == b.dst_hi; }

bool compact_hetero_eligible(const xpgnn_plan_t* p) {
  if (!compact_enabled() || !knobs().compact_hetero) return false;
  if (p->prune || p->n_layers < 1 || p->n_layers > kMaxConvIso || p->n_head > kMaxHeadC || p->n_nodes >= (1 << kPackShift)) return false;
  if (p->precision == 2) return false;  // bf16 storage cannot accumulate relation by relation
  int n_self = 0;
  for (int l = 0; l < p->n_layers; ++l) {
    const xpgnn_layer_t& L = p->layers_host[l];
    if (L.n_rel < 1 || L.h_out % 64 || L.h_out > 256) return false;
    if (l > 0 && L.h_in != p->layers_host[l - 1].h_out) return false;
    int groups = 0;
    for (int r = 0; r < L.n_rel; ++r) {
      const xpgnn_relation_t& R = L.rel_host[r];
      if (R.conv_kind == XPGNN_CONV_GCN && (R.src_lo != R.dst_lo || R.src_hi != R.dst_hi)) return false;
      if (R.dst_hi <= R.dst_lo || R.src_hi <= R.src_lo) return false;
      bool first = true;
      for (int q = 0; q < r; ++q) first = first && !same_dst(L.rel_host[q], R);
      groups += first;
      if (l == 0 && R.conv_kind == XPGNN_CONV_GCN) ++n_self;
    }
    if (groups > kMaxIsoGroups) return false;
  }
  return n_self <= kMaxIsoSelf;
}

static HLayout hetero_carve(const xpgnn_plan_t* p, void* ws, int64_t cap, int tile) {
  HLayout h{};
  Bump b(ws, cap);
  const int64_t N = p->n_nodes;
  const int NL = p->n_layers;
  h.map.resize(NL);
  for (int l = 0; l < NL; ++l) {
    const xpgnn_layer_t& L = p->layers_host[l];
    h.map[l].resize(L.n_rel);
    for (int r = 0; r < L.n_rel; ++r) {
      const xpgnn_relation_t& R = L.rel_host[r];
      int id = -1;
      for (size_t i = 0; i < h.csr.size(); ++i)
        if (h.csr[i].rowptr == R.rowptr && h.csr[i].col == R.col && h.csr[i].kind == R.conv_kind) id = (int)i;
      if (id < 0) {
        HCsr c{};
        c.rowptr = R.rowptr; c.col = R.col; c.kind = R.conv_kind; c.lo = R.dst_lo; c.hi = R.dst_hi; c.n_edges = R.n_edges;
        h.csr.push_back(c);
        id = (int)h.csr.size() - 1;
      }
      h.map[l][r] = id;
      h.csr[id].n_layers_using += 1;
    }
  }
  const xpgnn_layer_t& L0 = p->layers_host[0];
  for (int r = 0; r < L0.n_rel; ++r) {
    const xpgnn_relation_t& R = L0.rel_host[r];
    float* z = b.take<float>((int64_t)(R.src_hi - R.src_lo) * L0.h_out);
    h.zr.push_back(z ? z - (int64_t)R.src_lo * L0.h_out : nullptr);
  }
  h.r0c = b.take<float>(N * L0.h_out);
  int64_t nd_max = 1;
  for (auto& c : h.csr) nd_max = std::max<int64_t>(nd_max, c.hi - c.lo);
  h.keys = b.take<unsigned long long>((int64_t)tile * nd_max + 1);
  h.scanned = b.take<unsigned long long>((int64_t)tile * nd_max + 1);
  h.cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, h.cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int64_t)tile * nd_max + 1);
  h.cub_tmp = b.take<char>((int64_t)h.cub_bytes + 256);
  h.tile_active = b.take<unsigned long long>(32);
  for (auto& c : h.csr) {
    const int64_t E = std::max(c.n_edges, 1), nd = c.hi - c.lo;
    c.ebits = b.take<uint32_t>(E);
    // everything below is indexed by the position in the destination range; `scale` and `wgt` are read by global node
    // id (destination row, GCN: source of the same type), so their pointers are biased by the range start
    float* sc = b.take<float>(nd * 32);
    c.scale = sc ? sc - (int64_t)c.lo * 32 : nullptr;
    float* wg = c.kind == XPGNN_CONV_GCN ? b.take<float>((int64_t)tile * nd) : nullptr;
    c.wgt = wg;
    c.act_list = b.take<int32_t>((int64_t)tile * nd);
    c.rowptr_c = b.take<uint32_t>((int64_t)tile * (nd + 1));
    c.slot_info = b.take<int2>(32);
    c.slot_base = b.take<long long>(32);
    c.slot_tile_start = b.take<int32_t>(33);
    c.n_tiles = b.take<int32_t>(1);
    c.counters = b.take<int32_t>(32);
    c.long_rows = b.take<int32_t>(E / kLongRow + 1);
    c.n_long_dev = b.take<int32_t>(1);
    c.long_cap = (int)(E / kLongCompact + 1);
  c.long_list = b.take<int32_t>((int64_t)tile * c.long_cap);
    c.n_long_list = b.take<int32_t>(32);
    c.rows_packed = b.take<int32_t>((int64_t)tile * ceil_div(nd, 128) * 128);
    c.rs_packed = b.take<float>((int64_t)tile * ceil_div(nd, 128) * 128);
    c.ccol = b.take<int32_t>((int64_t)tile * E);
  }
  h.wroot.resize(NL); h.iso_w.resize(NL); h.iso_b.resize(NL);
  int hmax = 0;
  for (int l = 0; l < NL; ++l) {
    const xpgnn_layer_t& L = p->layers_host[l];
    hmax = std::max(hmax, L.h_out);
    if (l == 0) continue;
    for (int r = 0; r < L.n_rel; ++r) {
      h.wroot[l].push_back(b.take<float>((int64_t)L.h_out * L.h_in));
      h.iso_w[l].push_back(b.take<float>((int64_t)L.h_out * L.h_in));
      h.iso_b[l].push_back(b.take<float>(L.h_out));
    }
  }
  h.iso_out = b.take<float>((int64_t)p->n_query * hmax);
  h.hbuf[0] = b.take<float>((int64_t)tile * N * hmax);
  if (NL > 1) {
    h.hbuf[1] = b.take<float>((int64_t)tile * N * hmax);
    // aggregate buffer of the relation-by-relation path == pool of the transformed sources Z_r of one destination group
    // (transform-first path): the larger of the two
    int64_t pool = (int64_t)tile * N * hmax;
    for (int l = 1; l < NL; ++l) {
      const xpgnn_layer_t& L = p->layers_host[l];
      for (int r = 0; r < L.n_rel; ++r) {
        int64_t need = (int64_t)tile * (L.rel_host[r].dst_hi - L.rel_host[r].dst_lo) * L.h_out;  // merged root term
        for (int q = 0; q < L.n_rel; ++q)
          if (same_dst(L.rel_host[q], L.rel_host[r])) need += (int64_t)tile * (L.rel_host[q].src_hi - L.rel_host[q].src_lo) * L.h_out;
        pool = std::max(pool, need);
      }
    }
    h.agg = b.take<float>(pool);
  }
  h.bytes = (b.off + 255) & ~255ll;
  return h;
}

int64_t compact_hetero_workspace_bytes(const xpgnn_plan_t* p, int tile) { return hetero_carve(p, nullptr, 0, tile).bytes; }

int forward_compact_hetero(const xpgnn_plan_t* p, const uint32_t* act, int32_t W, int32_t s0, int32_t n_s, float* y, void* workspace,
                           int64_t workspace_bytes, int64_t* stats, cudaStream_t st, int dense_prec) {
  const int N = p->n_nodes, NL = p->n_layers, cw = 32, cw_lg = 5;
  int tile = 32;
  while (tile > 1 && hetero_carve(p, nullptr, 0, tile).bytes > workspace_bytes) tile >>= 1;
  XP_REQUIRE(hetero_carve(p, nullptr, 0, tile).bytes <= workspace_bytes, "workspace too small even for one coalition per tile");
  HLayout lay = hetero_carve(p, workspace, workspace_bytes, tile);
  int hmax = 0;
  for (int l = 0; l < NL; ++l) hmax = std::max(hmax, p->layers_host[l].h_out);
  const int64_t hstride = (int64_t)N * hmax, cstride = (int64_t)N * cw;
  const xpgnn_layer_t& L0 = p->layers_host[0];
  XP_REQUIRE(L0.h_in == p->f_in, "layer 0 input width != feature width");
  for (auto& c : lay.csr) XP_REQUIRE((int64_t)tile * std::max(c.n_edges, 1) < (1ll << kKeyShift), "tile x edges exceeds the packed scan key");

  // first / last relation into every destination group, per layer
  std::vector<std::vector<char>> first(NL), last(NL), group_root(NL), group_bias(NL);
  for (int l = 0; l < NL; ++l) {
    const xpgnn_layer_t& L = p->layers_host[l];
    first[l].assign(L.n_rel, 1); last[l].assign(L.n_rel, 1); group_root[l].assign(L.n_rel, 0); group_bias[l].assign(L.n_rel, 0);
    for (int r = 0; r < L.n_rel; ++r)
      for (int q = 0; q < L.n_rel; ++q)
        if (q != r && same_dst(L.rel_host[q], L.rel_host[r])) {
          if (q < r) first[l][r] = 0;
          if (q > r) last[l][r] = 0;
        }
  }

  // ---- coalition-invariant part: Z_r = X W_r^T (row-major), R0 = sum_r (b_r + X W_root,r^T) (chunk-major) ----
  XP_CHECK(cudaMemsetAsync(lay.r0c, 0, sizeof(float) * (int64_t)N * L0.h_out, st));
  for (int r = 0; r < L0.n_rel; ++r) {
    const xpgnn_relation_t& R = L0.rel_host[r];
    DenseArgs z{};
    z.in = p->x; z.ld_in = p->f_in; z.k = p->f_in; z.w = R.w_nbr; z.n_out = L0.h_out; z.out = lay.zr[r]; z.ld_out = L0.h_out;
    z.rows_per_s = R.src_hi - R.src_lo; z.row_lo = R.src_lo; z.M = z.rows_per_s; z.dst_lo = 0; z.dst_hi = N;
    if (launch_dense(z, st, dense_prec)) return 1;
    const bool sage_root = R.conv_kind == XPGNN_CONV_SAGE_MEAN && R.w_root;
    if (sage_root || R.b_nbr) {
      DenseArgs rt = z;  // k = 0 degenerates to "add the bias"
      rt.k = sage_root ? p->f_in : 0; rt.w = sage_root ? R.w_root : R.w_nbr; rt.b = R.b_nbr; rt.out = lay.r0c;
      rt.ld_out = cw; rt.cw_out = cw; rt.cw_out_lg = cw_lg; rt.out_chunk_stride = cstride;
      rt.rows_per_s = R.dst_hi - R.dst_lo; rt.row_lo = R.dst_lo; rt.M = rt.rows_per_s; rt.accumulate = 1;
      if (launch_dense(rt, st, dense_prec)) return 1;
    }
  }
  // merged weights: SAGE root transform per destination group, isolated chain per group
  IsoArgs iso{};
  iso.n_layers = NL; iso.h0 = L0.h_out; iso.act0 = L0.act; iso.r0c = lay.r0c; iso.r0_chunk_stride = cstride;
  for (int r = 0; r < L0.n_rel; ++r) {
    const xpgnn_relation_t& R = L0.rel_host[r];
    if (R.conv_kind != XPGNN_CONV_GCN) continue;
    iso.self_z[iso.n_self] = lay.zr[r]; iso.self_lo[iso.n_self] = R.dst_lo; iso.self_hi[iso.n_self] = R.dst_hi;
    ++iso.n_self;
  }
  int max_dim = L0.h_out;
  for (int l = 1; l < NL; ++l) {
    const xpgnn_layer_t& L = p->layers_host[l];
    iso.h_in[l] = L.h_in; iso.h_out[l] = L.h_out; iso.act[l] = L.act;
    max_dim = std::max(max_dim, std::max(L.h_in, L.h_out));
    const int nw = L.h_out * L.h_in;
    for (int r = 0; r < L.n_rel; ++r) {
      if (!last[l][r]) continue;
      int n_root = 0, n_w = 0, n_b = 0;
      for (int q = 0; q <= r; ++q) {
        const xpgnn_relation_t& Q = L.rel_host[q];
        if (!same_dst(Q, L.rel_host[r])) continue;
        const float* wq = Q.conv_kind == XPGNN_CONV_GCN ? Q.w_nbr : Q.w_root;
        if (Q.conv_kind == XPGNN_CONV_SAGE_MEAN && Q.w_root) {
          XP_LAUNCH(compact_add_into_kernel, (int)ceil_div(nw, 256), 256, 0, st, lay.wroot[l][r], Q.w_root, nw, n_root == 0);
          ++n_root;
        }
        if (wq) {
          XP_LAUNCH(compact_add_into_kernel, (int)ceil_div(nw, 256), 256, 0, st, lay.iso_w[l][r], wq, nw, n_w == 0);
          ++n_w;
        }
        if (Q.b_nbr) {
          XP_LAUNCH(compact_add_into_kernel, (int)ceil_div(L.h_out, 256), 256, 0, st, lay.iso_b[l][r], Q.b_nbr, L.h_out, n_b == 0);
          ++n_b;
        }
      }
      group_root[l][r] = n_root > 0;
      group_bias[l][r] = n_b > 0;
      const int g = iso.n_groups[l]++;
      iso.g_lo[l][g] = L.rel_host[r].dst_lo; iso.g_hi[l][g] = L.rel_host[r].dst_hi;
      iso.g_w[l][g] = n_w ? lay.iso_w[l][r] : nullptr; iso.g_b[l][g] = n_b ? lay.iso_b[l][r] : nullptr;
    }
  }
  iso.query = p->query; iso.iso_out = lay.iso_out;
  XP_LAUNCH(hetero_iso_kernel, p->n_query, 128, sizeof(float) * 2 * max_dim, st, iso, max_dim);

  // hub rows per relation CSR
  for (auto& c : lay.csr) {
    XP_CHECK(cudaMemsetAsync(c.counters, 0, 32 * sizeof(int32_t), st));
    XP_CHECK(cudaMemsetAsync(c.n_long_dev, 0, sizeof(int32_t), st));
    XP_LAUNCH(find_long_rows_kernel, (int)ceil_div(N, 256), 256, 0, st, c.rowptr, N, kLongRow, c.long_rows, c.n_long_dev);
  }
  {
    std::vector<int32_t> nl(lay.csr.size());
    for (size_t i = 0; i < lay.csr.size(); ++i)
      XP_CHECK(cudaMemcpyAsync(&nl[i], lay.csr[i].n_long_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    XP_CHECK(cudaStreamSynchronize(st));
    const bool off = !knobs().long_rows;
    for (size_t i = 0; i < lay.csr.size(); ++i) lay.csr[i].n_long = off ? 0 : nl[i];
  }

  const int w_first = s0 / 32, w_last = (s0 + n_s - 1) / 32;
  for (int w = w_first; w <= w_last; ++w) {
    const int bits_in_word = std::min(32, s0 + n_s - w * 32);
    for (int b0 = 0; b0 < bits_in_word; b0 += tile) {
      const int nb = std::min(tile, bits_in_word - b0);
      // ---- per-tile compaction of every relation ----
      for (size_t i = 0; i < lay.csr.size(); ++i) {
        HCsr& c = lay.csr[i];
        const int lt = c.n_long > 0 ? kLongRow : 0;
        const int nd = c.hi - c.lo;  // rows of the destination range: every per-relation structure is indexed by them
        const int grid_nd = (int)std::min<int64_t>(std::max<int64_t>(ceil_div(nd, 8 * kRowGrab), 1), (int64_t)kNumSMs * 8);
        {
          ProfScope ps(PROF_SCALE, st);
          XP_LAUNCH(compact_degree_kernel, grid_nd, 256, 0, st, c.rowptr, c.col, act, W, w, b0, nb, nd, c.ebits, lay.keys, c.scale,
                    c.kind, c.counters + 15, lt, c.lo);
          if (c.n_long > 0)
            XP_LAUNCH(compact_degree_long_kernel, c.n_long, 256, 0, st, c.rowptr, c.col, act, W, w, b0, nb, nd, c.ebits, lay.keys,
                      c.scale, c.kind, c.long_rows, c.lo);
        }
        ProfScope ps(PROF_COMPACT, st);
        XP_CHECK(cudaMemsetAsync(lay.keys + (int64_t)nb * nd, 0, sizeof(unsigned long long), st));  // sentinel of the scan
        XP_CHECK(cudaMemsetAsync(c.n_long_list, 0, 32 * sizeof(int32_t), st));
        size_t tmp = lay.cub_bytes;
        XP_CHECK(cub::DeviceScan::ExclusiveSum(lay.cub_tmp, tmp, lay.keys, lay.scanned, (int64_t)nb * nd + 1, st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        XP_LAUNCH(compact_finalize_kernel, dim3((unsigned)ceil_div(nd, 256), (unsigned)nb), 256, 0, st, lay.keys, lay.scanned, nd, c.lo,
                  c.act_list, c.rowptr_c, c.wgt, c.slot_info, c.slot_base, c.rows_packed, c.rs_packed, lt ? kLongCompact : 0,
                  c.long_list, c.n_long_list, c.long_cap);
        XP_LAUNCH(compact_tilemap_kernel, 1, 256, 0, st, c.slot_info, nb, c.slot_tile_start, c.n_tiles, c.n_layers_using, stats,
                  c.counters, i == 0 ? 1 : 0);
        XP_LAUNCH(compact_edges_kernel, grid_nd, 256, 0, st, c.rowptr, c.col, c.ebits, act, W, w, b0, nb, nd, lay.scanned, c.ccol,
                  c.counters + 14, lt, c.lo);
        if (c.n_long > 0)
          XP_LAUNCH(compact_edges_long_kernel, c.n_long, 256, 0, st, c.rowptr, c.col, c.ebits, act, W, w, b0, nb, nd, lay.scanned, c.ccol,
                    c.long_rows, c.lo);
        XP_LAUNCH(compact_sum_active_kernel, 1, 32, 0, st, c.slot_info, nb, lay.tile_active, i == 0);
      }
      float* cur = lay.hbuf[0];
      float* nxt = lay.hbuf[1];
      for (int l = 0; l < NL; ++l) {
        const xpgnn_layer_t& L = p->layers_host[l];
        // layer 0, one pass per destination type over all of its incoming relations (no hub rows, <= kMaxL0Rel relations)
        bool multi_l0 = l == 0 && knobs().l0_multi != 0;
        if (multi_l0) {
          for (auto& c : lay.csr) multi_l0 = multi_l0 && c.n_long == 0;
          for (int r = 0; r < L.n_rel && multi_l0; ++r) {
            int members = 0;
            for (int q = 0; q < L.n_rel; ++q) members += same_dst(L.rel_host[q], L.rel_host[r]);
            multi_l0 = members <= kMaxL0Rel;
          }
        }
        if (multi_l0) {
          for (int r = 0; r < L.n_rel; ++r) {
            if (!first[l][r]) continue;
            const xpgnn_relation_t& R = L.rel_host[r];
            L0MultiArgs a{};
            for (int q = r; q < L.n_rel; ++q) {
              if (!same_dst(L.rel_host[q], R)) continue;
              const HCsr& cq = lay.csr[lay.map[l][q]];
              L0Rel& x = a.rel[a.n_rel++];
              x.rowptr = cq.rowptr; x.col = cq.col; x.ebits = cq.ebits; x.scale = cq.scale; x.z = lay.zr[q]; x.kind = L.rel_host[q].conv_kind;
            }
            a.act = act; a.W = W; a.w = w; a.b0 = b0; a.nb = nb; a.h0 = L.h_out; a.r0c = lay.r0c; a.r0_chunk_stride = cstride;
            a.out = cur; a.out_s_stride = hstride; a.out_chunk_stride = cstride; a.act_fn = L.act;
            a.counter = lay.csr[lay.map[l][r]].counters + 13; a.row_lo = R.dst_lo; a.row_hi = R.dst_hi;
            ProfScope ps(PROF_SPMM_INVARIANT, st);
            if (launch_l0_multi(a, L.act == XPGNN_ACT_SIGMOID, st)) return 1;
          }
          continue;
        }
        // layers >= 1 transform-first: Z_q = H W_q^T per relation over the active rows of its source type, then one gather
        // pass per destination type over all of its relations (no hub rows, <= kMaxL0Rel relations per type)
        bool multi_l1 = l > 0 && knobs().l1_multi == 1;  // opt-in: measured slower on C4 (200 vs 224 evals/s)
        if (multi_l1) {
          for (auto& c : lay.csr) multi_l1 = multi_l1 && c.n_long == 0;
          for (int r = 0; r < L.n_rel && multi_l1; ++r) {
            int members = 0;
            for (int q = 0; q < L.n_rel; ++q) members += same_dst(L.rel_host[q], L.rel_host[r]);
            multi_l1 = members <= kMaxL0Rel;
          }
        }
        if (multi_l1) {
          auto type_csr = [&](int lo, int hi) {  // a relation whose destination range is this node type: its tile table lists the type's active rows
            for (size_t i = 0; i < lay.csr.size(); ++i)
              if (lay.csr[i].lo == lo && lay.csr[i].hi == hi) return (int)i;
            return -1;
          };
          for (int r = 0; r < L.n_rel; ++r) {
            if (!first[l][r]) continue;
            const xpgnn_relation_t& R = L.rel_host[r];
            const HCsr& cd = lay.csr[lay.map[l][r]];
            const int nd = cd.hi - cd.lo;
            int r_last = r;
            for (int q = r; q < L.n_rel; ++q)
              if (same_dst(L.rel_host[q], R)) r_last = q;
            float* pool = lay.agg;
            CspmmMultiArgs m{};
            DenseArgs d{};
            d.in = cur; d.in_s_stride = hstride; d.ld_in = cw; d.k = L.h_in; d.cw_in = cw; d.cw_in_lg = cw_lg; d.in_chunk_stride = cstride;
            d.n_out = L.h_out; d.ld_out = cw; d.cw_out = cw; d.cw_out_lg = cw_lg; d.dst_lo = 0; d.dst_hi = N; d.act_fn = XPGNN_ACT_NONE;
            for (int q = r; q < L.n_rel; ++q) {
              const xpgnn_relation_t& Q = L.rel_host[q];
              if (!same_dst(Q, R)) continue;
              const HCsr& cq = lay.csr[lay.map[l][q]];
              const int ts = type_csr(Q.src_lo, Q.src_hi);
              if (ts < 0) continue;  // the source type never receives messages: it has no rows in layers >= 1
              const int ns = Q.src_hi - Q.src_lo;
              float* zq = pool;
              pool += (int64_t)nb * ns * L.h_out;
              DenseArgs dz = d;
              dz.w = Q.w_nbr; dz.b = nullptr; dz.out = zq - (int64_t)Q.src_lo * cw; dz.out_s_stride = (int64_t)ns * L.h_out;
              dz.out_chunk_stride = (int64_t)ns * cw; dz.rows_packed = lay.csr[ts].rows_packed; dz.n_tiles_dev = lay.csr[ts].n_tiles;
              dz.rows_per_s = ns; dz.M = (int64_t)nb * ceil_div(ns, 128) * 128;
              if (launch_dense(dz, st, dense_prec)) return 1;
              MRel& x = m.rel[m.n_rel++];
              x.rowptr_c = cq.rowptr_c; x.slot_base = cq.slot_base; x.ccol = cq.ccol; x.z = dz.out; x.z_s_stride = dz.out_s_stride;
              x.z_chunk_stride = dz.out_chunk_stride; x.kind = Q.conv_kind;
              x.wgt = Q.conv_kind == XPGNN_CONV_GCN ? cq.wgt - cq.lo : nullptr;
            }
            if (group_root[l][r_last]) {  // merged root term of the destination type
              DenseArgs dr = d;
              dr.w = lay.wroot[l][r_last]; dr.b = nullptr; dr.out = pool - (int64_t)cd.lo * cw; dr.out_s_stride = (int64_t)nd * L.h_out;
              dr.out_chunk_stride = (int64_t)nd * cw; dr.rows_packed = cd.rows_packed; dr.n_tiles_dev = cd.n_tiles;
              dr.rows_per_s = nd; dr.M = (int64_t)nb * ceil_div(nd, 128) * 128;
              if (launch_dense(dr, st, dense_prec)) return 1;
              m.zroot = dr.out; m.zr_s_stride = dr.out_s_stride; m.zr_chunk_stride = dr.out_chunk_stride;
            }
            m.nb = nb; m.nd = nd; m.n_chunks = L.h_out / cw; m.act_fn = L.act;
            m.slot_info = cd.slot_info; m.slot_tile_start = cd.slot_tile_start; m.act_list = cd.act_list;
            m.bias = group_bias[l][r_last] ? lay.iso_b[l][r_last] : nullptr;
            m.out = nxt; m.out_s_stride = hstride; m.out_chunk_stride = cstride; m.counter = lay.csr[lay.map[l][r]].counters + std::min(l, 12);
            ProfScope ps(PROF_SPMM_TILE, st);
            int per_sm = 0;
            XP_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cspmm_multi_kernel, 256, 0));
            XP_LAUNCH(cspmm_multi_kernel, kNumSMs * std::max(per_sm, 1), 256, 0, st, m);
          }
          std::swap(cur, nxt);
          continue;
        }
        for (int r = 0; r < L.n_rel; ++r) {
          const xpgnn_relation_t& R = L.rel_host[r];
          HCsr& c = lay.csr[lay.map[l][r]];
          const bool gcn = R.conv_kind == XPGNN_CONV_GCN;
          if (l == 0) {
            L0RowsArgs a{};
            a.rowptr = c.rowptr; a.col = c.col; a.ebits = c.ebits; a.act = act; a.W = W; a.w = w; a.b0 = b0; a.nb = nb; a.N = N;
            a.scale = c.scale; a.z = lay.zr[r]; a.h0 = L.h_out; a.r0c = lay.r0c; a.r0_chunk_stride = cstride;
            a.out = cur; a.out_s_stride = hstride; a.out_chunk_stride = cstride; a.out_row_stride = 32; a.r0_row_stride = 32;
            a.kind = R.conv_kind; a.act_fn = L.act; a.prescale = 0;
            a.long_rows = c.long_rows; a.long_threshold = c.n_long > 0 ? kLongRow : 0; a.counter = c.counters + 13;
            a.row_lo = R.dst_lo; a.row_hi = R.dst_hi; a.accumulate = !first[l][r]; a.finish = last[l][r];
            ProfScope ps(PROF_SPMM_INVARIANT, st);
            const bool sg = L.act == XPGNN_ACT_SIGMOID;
            if (launch_l0_rows(a, sg, false, R.dst_hi - R.dst_lo, st)) return 1;
            if (c.n_long > 0) {
              if (launch_l0_long_rows(a, sg, false, c.n_long, st)) return 1;
            }
            // the row counter of this relation's layer-0 launch is used once per tile; compact_tilemap_kernel re-zeroes it
          } else {
            CspmmArgs s{};
            s.nb = nb; s.N = c.hi - c.lo; s.kind = R.conv_kind; s.layer0 = gcn; s.prof_cat = PROF_SPMM_TILE; s.n_chunks = L.h_in / cw;
            s.slot_info = c.slot_info; s.slot_tile_start = c.slot_tile_start; s.act_list = c.act_list; s.rowptr_c = c.rowptr_c;
            s.slot_base = c.slot_base; s.ccol = c.ccol;
            s.in = cur; s.in_s_stride = hstride; s.in_chunk_stride = cstride; s.wgt = gcn ? c.wgt - c.lo : nullptr;  // read by global source id
            s.out = lay.agg; s.out_s_stride = hstride; s.out_chunk_stride = cstride; s.act_fn = XPGNN_ACT_NONE;
            s.counter = c.counters + std::min(l, 12); s.l2_stream = 1; s.l2_gather = 0;
            s.long_cnt = c.n_long > 0 ? kLongCompact : 0; s.long_list = c.long_list; s.n_long_list = c.n_long_list;
            s.long_cap = c.long_cap; s.counter_long = c.counters + 16 + std::min(l, 12);
            if (launch_cspmm(s, cw, st)) return 1;
            const bool root = last[l][r] && group_root[l][r];
            DenseArgs d{};
            d.in = lay.agg; d.in_s_stride = hstride; d.ld_in = cw; d.k = L.h_in; d.cw_in = cw; d.cw_in_lg = cw_lg; d.in_chunk_stride = cstride;
            d.w = R.w_nbr; d.b = R.b_nbr; d.n_out = L.h_out;
            d.out = nxt; d.out_s_stride = hstride; d.ld_out = cw; d.cw_out = cw; d.cw_out_lg = cw_lg; d.out_chunk_stride = cstride;
            d.rows_packed = c.rows_packed; d.n_tiles_dev = c.n_tiles;
            d.rows_per_s = R.dst_hi - R.dst_lo; d.M = (int64_t)nb * ceil_div(R.dst_hi - R.dst_lo, 128) * 128; d.dst_lo = 0; d.dst_hi = N;
            d.accumulate = !first[l][r]; d.act_fn = (last[l][r] && !root) ? L.act : XPGNN_ACT_NONE;
            if (launch_dense(d, st, dense_prec)) return 1;
            if (root) {  // one root transform per destination group, with the summed lin_r weights
              DenseArgs rt = d;
              rt.in = cur; rt.w = lay.wroot[l][r]; rt.b = nullptr; rt.accumulate = 1; rt.act_fn = L.act;
              if (launch_dense(rt, st, dense_prec)) return 1;
            }
          }
        }
        if (l > 0) std::swap(cur, nxt);
      }
      // ---- head on the query rows ----
      CHeadArgs h{};
      int hd = p->layers_host[NL - 1].h_out;
      h.n_iso = 0; h.iso_out = lay.iso_out;
      h.n_head = p->n_head;
      for (int i = 0; i < p->n_head; ++i) {
        h.head[i] = p->head_host[i];
        hd = std::max(hd, std::max(p->head_host[i].in, p->head_host[i].out));
      }
      h.in = cur; h.in_s_stride = hstride; h.in_chunk_stride = cstride; h.dim0 = p->layers_host[NL - 1].h_out; h.cw = cw; h.cw_lg = cw_lg;
      h.query = p->query; h.n_query = p->n_query; h.out_col = p->out_col;
      h.y = y + ((int64_t)(w * 32 + b0) - s0) * p->n_query;
      h.act = act; h.W = W; h.w = w; h.b0 = b0;
      h.tile_active = p->zero_edge_rule ? lay.tile_active : nullptr;
      {
        ProfScope ps(PROF_HEAD, st);
        XP_LAUNCH(compact_head_kernel, nb * p->n_query, 128, sizeof(float) * 2 * hd, st, h, hd);
      }
    }
  }
  return 0;
}

}  // namespace xpgnn
