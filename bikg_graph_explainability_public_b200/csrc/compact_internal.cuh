// Shared by the compact-path translation units: compact.cu (per-tile compaction, list-driven SpMM, head, the two
// drivers) and compact_l0.cu (layer-0 row kernels).
#pragma once
#include <algorithm>

#include "engine_internal.cuh"

namespace xpgnn {

// Dynamic row scheduling for the row-per-warp kernels: a warp takes kRowGrab consecutive rows at a time from a
// global counter.  (A static stride is badly unbalanced on power-law graphs: with R-MAT ids the in-degree is a
// function of the id's bit pattern, and a stride that is a multiple of 64 hands some CTAs rows ~1000x heavier.)
constexpr int kRowGrab = 4;
__device__ __forceinline__ int grab_rows(int32_t* counter, int lane) {
  int base = 0;
  if (lane == 0) base = atomicAdd(counter, kRowGrab);
  return __shfl_sync(0xffffffffu, base, 0);
}


// ------------------------------------------------------------------------------------------
// layer 0, row-outer: the gathered operand Z = X W^T is coalition invariant, so a neighbour row crosses
// the L2 fabric ONCE per destination row and is reused from registers for every coalition slot of the tile
// (the list-driven kernel below moves it once per (slot, edge): 82 GB per C3 tile against ~12 GB here).
// ------------------------------------------------------------------------------------------
struct L0RowsArgs {
  const int32_t* rowptr;
  const int32_t* col;
  const uint32_t* ebits;
  const uint32_t* act;
  int W, w, b0, nb, N;
  const float* scale;        // [N][32] by bit of the word
  const float* z;            // [N][h0] row-major
  int h0;
  const float* bias;         // GCN bias or NULL
  const float* r0c;          // SAGE: b + X W_root^T, chunk-major (cw = 32) or NULL
  int64_t r0_chunk_stride;
  float* out;                // chunk-major (cw = 32) activations of the tile
  int64_t out_s_stride, out_chunk_stride;
  int kind, act_fn, prescale;
  const int32_t* long_rows;  // rows with more than long_threshold in-edges (hub rows): one CTA each in the LONG variant
  int long_threshold;        // 0: no splitting
  int32_t* counter;          // row counter of the dynamic schedule (zero at launch)
  // several relations into one destination type (HeteroConv sum): the first writes, the others add, the last finishes
  int row_lo, row_hi;        // destination range of the relation (rows outside are skipped)
  int accumulate;            // 1: add the partial sum already stored in `out`
  int finish;                // 1: add the bias / addend, apply the activation (and pre-scale); 0: store the partial sum
  // fp32 outputs / addends: element (chunk, row v, column c of the chunk) sits at chunk * chunk_stride + v * row_stride + c --
  // chunk-major activations of the compact path: row_stride 32; a row-major [N][ld] buffer (tile path): chunk_stride 32, row_stride ld
  int out_row_stride, r0_row_stride;
  const int32_t* rows;       // optional row list (pruned mode of the tile path): row i of the launch is rows[i], i < n_list
  int n_list;
  unsigned long long* dbg;   // XPGNN_L0_DBG: cycle counters of warp pair 0 of CTA 0 (diagnostics only)
  // LONG launch, optional: hub rows cut into slices of kL0Slice in-edges, one CTA per slice (an 82 K-in-edge row of an R-MAT
  // graph kept one CTA busy for ~2 ms per tile while the rest of the GPU had finished).  Item i = slice item_slice[i] of hub row
  // long_rows[item_row[i]]; the items of row r are row_item0[r] .. row_item0[r + 1] - 1.  A row with several slices leaves
  // its partial accumulators in slice_scratch and l0_long_reduce_kernel adds them in slice order and runs the epilogue.
  const int32_t* item_row;
  const int32_t* item_slice;
  const int32_t* row_item0;
  float2* slice_scratch;     // [item][column block][32 accumulators][32 lanes]
};
constexpr int kL0Slice = 4096;  // in-edges of a hub-row slice

constexpr int kLongRow = 1024;    // in-edges above which a destination row is processed by a whole CTA
constexpr int kLongCompact = 256; // active in-edges above which a compact row is processed by a whole CTA (SpMM)

constexpr int kMaxL0Rel = 12;
struct L0Rel {
  const int32_t* rowptr;
  const int32_t* col;
  const uint32_t* ebits;
  const float* scale;  // [.][32] by bit of the word, indexed by global node id (biased pointer)
  const float* z;      // [.][h0] row-major Z_r = X W_r^T, indexed by global source id (biased pointer)
  int kind;
};
struct L0MultiArgs {
  L0Rel rel[kMaxL0Rel];
  int n_rel;
  const uint32_t* act;
  int W, w, b0, nb, h0;
  const float* r0c;          // sum over the relations of b_r (+ X W_root,r^T), chunk-major
  int64_t r0_chunk_stride;
  float* out;
  int64_t out_s_stride, out_chunk_stride;
  int act_fn;
  int32_t* counter;
  int row_lo, row_hi;
};

// Arguments of the list-driven masked SpMM kernels of the compact path (compact.cu: row-lockstep / segmented / ring
// variants; compact_bulk.cu: the warp-specialised TMA bulk-copy variant).
struct CspmmArgs {
  int nb, N, n_chunks, kind, layer0, act_fn, prescale;  // layer0: the self term is weighted by deg^-1/2 as well (operand not pre-scaled)
  int prof_cat;               // -1: by layer0
  const int2* slot_info;
  const int32_t* slot_tile_start;
  const int32_t* act_list;    // [nb][N]
  const uint32_t* rowptr_c;   // [nb][N+1]
  const long long* slot_base; // [nb] first entry of the slot's list in ccol
  const int32_t* ccol;
  const float* in;            // chunk-major operand, rows by original node id
  int64_t in_s_stride;        // 0: coalition invariant (layer 0)
  int64_t in_chunk_stride;
  const float* wgt;           // [nb][N] per-source weight (layer-0 GCN) or NULL
  const float* addend;        // chunk-major coalition-invariant addend or NULL
  int64_t add_chunk_stride;
  const float* bias;          // per-column addend or NULL
  int32_t* counter;           // work counter (zeroed per launch); NULL: static round-robin
  int long_cnt;               // compact rows with more active in-edges are left to cspmm_long_kernel (0: none)
  const int32_t* long_list;   // [nb][long_cap] compact rows of the slot, written by compact_finalize_kernel
  const int32_t* n_long_list; // [nb]
  int long_cap;
  int32_t* counter_long;      // work counter of cspmm_long_kernel (zeroed per tile); NULL: static round-robin
  int l2_stream, l2_gather;   // eviction priority of the streamed / gathered accesses (l2_policy kinds)
  float* out;
  int64_t out_s_stride, out_chunk_stride;
};

// layers >= 1, aggregate-first: warp-specialised kernel whose gathers are TMA bulk copies into a shared-memory ring (compact_bulk.cu)
int launch_cspmm_bulk(const CspmmArgs& a, int variant, cudaStream_t st);

// ---- launchers of compact_l0.cu (return non-zero on a CUDA error, like every launch helper of the engine) ----
// rows that are not hub rows: warp specialised by default (XPGNN_L0_WS selects the variant)
int launch_l0_rows(const L0RowsArgs& r, bool sigmoid, bool out16, int n_rows, cudaStream_t st);
// hub rows (r.long_rows, more than r.long_threshold in-edges): one CTA per row, or per slice when r.item_row is set
// (n_items = number of slices over all hub rows)
int launch_l0_long_rows(const L0RowsArgs& r, bool sigmoid, bool out16, int n_long, cudaStream_t st, int n_items = 0);
// slices of the hub rows (graph property: once per forward call); *n_items_dev = number of items
int build_l0_long_items(const int32_t* rowptr, const int32_t* long_rows, const int32_t* n_long_dev, int32_t* item_row, int32_t* item_slice,
                        int32_t* row_item0, int32_t* n_items_dev, cudaStream_t st);
inline int64_t l0_long_items_max(int64_t n_edges) { return n_edges / kLongRow + n_edges / kL0Slice + 2; }
// The scratch of the sliced hub rows (8 KB per item and column block) is sized for at most this many items -- the worst case of
// l0_long_items_max would be 2.9 GB at 100 M edges for graphs that may have no hub row at all; a call with more items runs its
// hub rows unsliced (one CTA per row)
constexpr int64_t kL0SliceScratchItems = 16384;
inline int64_t l0_slice_scratch_items(int64_t n_edges) { return std::min<int64_t>(l0_long_items_max(n_edges), kL0SliceScratchItems); }
// all relations into one destination type in one pass (HeteroConv sum)
int launch_l0_multi(const L0MultiArgs& a, bool sigmoid, cudaStream_t st);

}  // namespace xpgnn
