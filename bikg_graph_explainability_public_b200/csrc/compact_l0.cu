// Layer 0 of the compact path: the gathered operand Z = X W^T is coalition invariant, so the kernels here go row-outer --
// a neighbour row crosses the L2 fabric ONCE per destination row and is reused for every coalition slot of the tile
// (the list-driven SpMM of compact.cu moves it once per (slot, edge): 82 GB per C3 tile against ~12 GB here).
//   l0_rows_kernel   one warp per row (hub rows: one CTA per row); also the XPGNN_L0_WS=0 variant
//   l0_ws_kernel     warp specialised (default): producer warps stage weights / Z pieces, consumer warps do the FMAs
//   l0_ws2_kernel    warp specialised with slot x column register tiles (opt-in)
//   l0_multi_kernel  HeteroConv(sum): all relations into one destination type in one pass
// Replaces, together with compact.cu, data.py:390-648 + model.py:62-328 of the reference for layer 0.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "compact_internal.cuh"

namespace xpgnn {

constexpr int kL0WStride = 36;  // floats per staged weight row (32 + pad: 16-byte aligned, 4-way instead of 32-way store conflicts)
constexpr int kL0SmemBytes = 8 * 32 * kL0WStride * 4 + 8 * 32 * 64 * 4 + 8 * 32 * 8;  // weights, Z pieces, source ids / row pointers

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// A warp owns one destination row and walks its 64-column blocks; a lane owns 2 columns of the block and the
// accumulators of the row's ACTIVE slots (the slots in which the destination node is inactive produce nothing:
// about half of them), at most 32 float2 = 64 registers.  Per batch of <= 32 in-edges:
//   weights : lane j loads the 32 per-slot scales of source u_j (8 independent 128-bit loads), zeroes the
//             slots in which edge j is inactive, stores the row to shared memory and compacts it in place to
//             the destination's active slots (k-th active slot -> position k), once per row;
//   operand : the 256-byte pieces of Z[u_j] go to shared memory with cp.async, all in flight together;
//   FMAs    : per edge 1 LDS.64 (z) + NQ broadcast LDS.128 (weights) + 8 NQ FFMA, NQ = ceil(active slots / 4)
//             a compile-time constant of the specialised loop (a zero weight is cheaper than a branch).
// Packed fp32 FMA of sm_100 (FFMA2): d.{x,y} += a.{x,y} * b.{x,y} in ONE instruction.  The accumulator pair is (slot 2i,
// slot 2i+1) of one column, so the weight pair comes straight out of the LDS.128 and only the two Z values of the edge
// have to be duplicated: half the FMA-pipe instructions of the scalar loop.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}
// accumulator layout: acc[2i] = (slot 2i, slot 2i+1) of the lane's first column, acc[2i+1] = the same slots of its second
__device__ __forceinline__ float2 l0_acc_get(const float2 (&acc)[32], int k) {
  const float2 px = acc[(k >> 1) * 2], py = acc[(k >> 1) * 2 + 1];
  return (k & 1) ? make_float2(px.y, py.y) : make_float2(px.x, py.x);
}

template <int NQ>
__device__ __forceinline__ void l0_fma(const float* __restrict__ w_rows, const float* __restrict__ z_rows, int n, int lane, float2 (&acc)[32]) {
  // two edges in flight while the weight registers allow it: the LDS latency of edge j + 1 hides behind the FMAs of edge j
#pragma unroll(NQ <= 4 ? 2 : 1)
  for (int j = 0; j < n; ++j) {
    const float2 zv = *reinterpret_cast<const float2*>(z_rows + j * 64 + lane * 2);
    const float2 zx = make_float2(zv.x, zv.x), zy = make_float2(zv.y, zv.y);
    const float4* wr = reinterpret_cast<const float4*>(w_rows + j * kL0WStride);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float4 w4 = wr[q];
      const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);
      ffma2(acc[4 * q + 0], w01, zx);
      ffma2(acc[4 * q + 1], w01, zy);
      ffma2(acc[4 * q + 2], w23, zx);
      ffma2(acc[4 * q + 3], w23, zy);
    }
  }
}

// ---- epilogue of one column block of a destination row (shared by the two layer-0 row kernels) ----
// operands that do not depend on the accumulators: the row's own Z piece (GCN self loop) and bias / SAGE root addend
__device__ __forceinline__ void l0_epilogue_operands(const L0RowsArgs& a, int v, int cb, int lane, bool gcn, float2& self, float2& add) {
  self = make_float2(0.f, 0.f);
  add = self;
  if (gcn) self = __ldg(reinterpret_cast<const float2*>(a.z + (int64_t)v * a.h0 + cb * 64 + lane * 2));
  if (a.bias && a.finish) add = __ldg(reinterpret_cast<const float2*>(a.bias + cb * 64 + lane * 2));
  if (a.r0c && a.finish) {
    const float2 r = __ldg(reinterpret_cast<const float2*>(a.r0c + (int64_t)(cb * 2 + (lane >> 4)) * a.r0_chunk_stride + (int64_t)v * a.r0_row_stride + (lane & 15) * 2));
    add.x += r.x; add.y += r.y;
  }
}
// position k of the accumulators is the k-th active slot; sc_v = lane b holds the destination's scale of slot b.
// Kept tiny (the kernels must stay inside the instruction cache): GCN and SAGE share one formula (SAGE: self weight 0),
// ReLU / identity are a max with 0 / -inf.
template <bool SIGMOID, bool OUT16>
__device__ __forceinline__ void l0_epilogue(const L0RowsArgs& a, int v, uint32_t av, int n_slots, int cb, int lane, float sc_v, bool gcn,
                                            float lower, const float2 (&acc)[32], const float2 self, const float2 add) {
  const int64_t cm_off = (int64_t)(cb * 2 + (lane >> 4)) * a.out_chunk_stride + (int64_t)v * a.out_row_stride + (lane & 15) * 2;
  float* outp = a.out + cm_off - (int64_t)a.b0 * a.out_s_stride;
  // bf16 layout: one 64-column block is one chunk, a lane's 2 columns are one bf162
  __nv_bfloat16* outp16 = reinterpret_cast<__nv_bfloat16*>(a.out) + (int64_t)cb * a.out_chunk_stride + (int64_t)v * 64 + lane * 2 -
                          (int64_t)a.b0 * a.out_s_stride;
  uint32_t m = av;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (k < n_slots) {  // warp uniform
      const int b = __ffs(m) - 1;
      m &= m - 1;
      const float dv = __shfl_sync(0xffffffffu, sc_v, b);
      const float sw = gcn ? dv : 0.0f, pv = a.prescale ? dv : 1.0f;
      const float2 ak = l0_acc_get(acc, k);
      float2 o;
      o.x = dv * fmaf(self.x, sw, ak.x) + add.x;
      o.y = dv * fmaf(self.y, sw, ak.y) + add.y;
      if (!OUT16 && a.accumulate) {  // partial sum of the relations before this one
        const float2 pr = *reinterpret_cast<const float2*>(outp + (int64_t)b * a.out_s_stride);
        o.x += pr.x; o.y += pr.y;
      }
      if (a.finish) {
        if (SIGMOID) {
          o.x = apply_act(o.x, XPGNN_ACT_SIGMOID); o.y = apply_act(o.y, XPGNN_ACT_SIGMOID);
        } else {
          o.x = fmaxf(o.x, lower); o.y = fmaxf(o.y, lower);
        }
        o.x *= pv; o.y *= pv;
      }
      if (OUT16) *reinterpret_cast<__nv_bfloat162*>(outp16 + (int64_t)b * a.out_s_stride) = __floats2bfloat162_rn(o.x, o.y);
      else __stcs(reinterpret_cast<float2*>(outp + (int64_t)b * a.out_s_stride), o);
    }
  }
}

// weight rows of a batch of <= 32 in-edges: lane j loads the 32 per-slot scales of source u_j (8 independent 128-bit loads),
// zeroes the slots in which edge j is inactive, stores the row and compacts it in place to the destination's active slots
__device__ __forceinline__ void l0_stage_weights(const L0RowsArgs& a, float* wrow, int u, uint32_t bits, uint32_t av, int nq, bool gcn) {
  float4 wq[8];
#pragma unroll
  for (int q = 0; q < 8; ++q)
    wq[q] = gcn ? __ldg(reinterpret_cast<const float4*>(a.scale + (int64_t)u * 32 + q * 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    wq[q].x = (bits >> (4 * q + 0)) & 1u ? wq[q].x : 0.0f;
    wq[q].y = (bits >> (4 * q + 1)) & 1u ? wq[q].y : 0.0f;
    wq[q].z = (bits >> (4 * q + 2)) & 1u ? wq[q].z : 0.0f;
    wq[q].w = (bits >> (4 * q + 3)) & 1u ? wq[q].w : 0.0f;
    *reinterpret_cast<float4*>(wrow + 4 * q) = wq[q];
  }
  // in place: position k <- bit b_k (k <= b_k and the bits ascend, so nothing unread is overwritten)
  int k = 0;
  for (uint32_t m = av; m; m &= m - 1, ++k) wrow[k] = wrow[__ffs(m) - 1];
  for (; k < 4 * nq; ++k) wrow[k] = 0.0f;
}

// FMA loop specialised on ceil(active slots / 4)
__device__ __forceinline__ void l0_fma_switch(int nq, const float* wr, const float* zr, int n, int lane, float2 (&acc)[32]) {
  switch (nq) {
    case 1: l0_fma<1>(wr, zr, n, lane, acc); break;
    case 2: l0_fma<2>(wr, zr, n, lane, acc); break;
    case 3: l0_fma<3>(wr, zr, n, lane, acc); break;
    case 4: l0_fma<4>(wr, zr, n, lane, acc); break;
    case 5: l0_fma<5>(wr, zr, n, lane, acc); break;
    case 6: l0_fma<6>(wr, zr, n, lane, acc); break;
    case 7: l0_fma<7>(wr, zr, n, lane, acc); break;
    default: l0_fma<8>(wr, zr, n, lane, acc); break;
  }
}

// LONG: one CTA per hub row -- the 8 warps take contiguous slices of the row's in-edges, the partial accumulators
// are summed through shared memory in a fixed order (deterministic) and warp 0 runs the epilogue.
// OUT16: the activations of the tile are stored as bf16 in 64-element chunks (precision = bf16 activation storage).
template <bool SIGMOID, bool LONG, bool OUT16>
__global__ void __launch_bounds__(256, 2) l0_rows_kernel(const L0RowsArgs a) {
  extern __shared__ __align__(16) uint8_t l0_smem[];
  float(*s_w)[32][kL0WStride] = reinterpret_cast<float(*)[32][kL0WStride]>(l0_smem);
  float(*s_z)[32][64] = reinterpret_cast<float(*)[32][64]>(l0_smem + 8 * 32 * kL0WStride * 4);
  int(*s_u)[32] = reinterpret_cast<int(*)[32]>(l0_smem + 8 * 32 * kL0WStride * 4 + 8 * 32 * 64 * 4);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t live = (a.nb == 32 ? 0xffffffffu : ((1u << a.nb) - 1u)) << a.b0;  // bits of the word in this tile
  const bool gcn = a.kind == XPGNN_CONV_GCN;
  const int ncb = a.h0 / 64;
  const float lower = a.act_fn == XPGNN_ACT_RELU ? 0.0f : -INFINITY;
  const int n_rows = a.rows ? a.n_list : a.row_hi - a.row_lo;  // row list (pruned mode) | destination range of the relation
  for (int vb = LONG ? 0 : grab_rows(a.counter, lane); vb < n_rows; vb = LONG ? n_rows : grab_rows(a.counter, lane))
  for (int vi = LONG ? 0 : vb; vi < (LONG ? 1 : min(n_rows, vb + kRowGrab)); ++vi) {
    int slice = 0, n_slices = 1;  // LONG with items: this CTA sums one slice of the hub row
    if (LONG && a.item_row) {
      const int ri = a.item_row[blockIdx.x];
      slice = a.item_slice[blockIdx.x];
      n_slices = a.row_item0[ri + 1] - a.row_item0[ri];
    }
    const int v = LONG ? a.long_rows[a.item_row ? a.item_row[blockIdx.x] : blockIdx.x] : (a.rows ? a.rows[vi] : a.row_lo + vi);
    const uint32_t av = a.act[(int64_t)v * a.W + a.w] & live;
    if (!av) continue;  // LONG: the same row for the whole CTA, so every warp leaves together
    const int n_slots = __popc(av), nq = (n_slots + 3) >> 2;
    int e0 = a.rowptr[v], e1 = a.rowptr[v + 1];
    if (LONG && a.item_row) {
      e0 += slice * kL0Slice;
      e1 = min(e1, e0 + kL0Slice);
    }
    if (!LONG && a.long_threshold > 0 && e1 - e0 > a.long_threshold) continue;  // hub row: left to the LONG launch
    const bool short_row = !LONG && e1 - e0 <= 32;
    const float sc_v = a.scale[(int64_t)v * 32 + lane];
    int bs = e0, be = e1;  // this warp's slice of the in-edges
    if (LONG) {
      const int nbatch = (e1 - e0 + 31) >> 5;
      bs = e0 + 32 * ((wib * nbatch) >> 3);
      be = min(e1, e0 + 32 * (((wib + 1) * nbatch) >> 3));
    }
    for (int cb = 0; cb < ncb; ++cb) {
      float2 acc[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[k] = make_float2(0.f, 0.f);
      for (int base = bs; base < be; base += 32) {
        const int n = min(32, be - base);
        __syncwarp();
        if (cb == 0 || !short_row) {  // weights of the batch (kept across column blocks for short rows)
          if (lane < n) {
            const int u = __ldg(a.col + base + lane);
            const uint32_t bits = __ldg(a.ebits + base + lane) & av;
            s_u[wib][lane] = u;
            l0_stage_weights(a, &s_w[wib][lane][0], u, bits, av, nq, gcn);
          }
          __syncwarp();
        }
        {
          const float* zc = a.z + cb * 64 + lane * 2;
#pragma unroll 4
          for (int j = 0; j < n; ++j) cp_async8(&s_z[wib][j][lane * 2], zc + (int64_t)s_u[wib][j] * a.h0);
        }
        cp_async_wait_all();
        __syncwarp();
        l0_fma_switch(nq, &s_w[wib][0][0], &s_z[wib][0][0], n, lane, acc);
      }
      if (LONG) {  // sum the 8 partial accumulators (fixed order) into warp 0
        float2* red = reinterpret_cast<float2*>(&s_z[0][0][0]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) red[(wib * 32 + k) * 32 + lane] = acc[k];
        __syncthreads();
        if (wib == 0) {
          for (int w8 = 1; w8 < 8; ++w8) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float2 p = red[(w8 * 32 + k) * 32 + lane];
              acc[k].x += p.x; acc[k].y += p.y;
            }
          }
        }
        __syncthreads();
        if (wib != 0) continue;
        if (n_slices > 1) {  // partial sums of this slice: l0_long_reduce_kernel finishes the row
          float2* sc = a.slice_scratch + ((int64_t)blockIdx.x * ncb + cb) * 1024;
#pragma unroll
          for (int k = 0; k < 32; ++k) sc[k * 32 + lane] = acc[k];
          continue;
        }
      }
      float2 self, add;
      l0_epilogue_operands(a, v, cb, lane, gcn, self, add);
      l0_epilogue<SIGMOID, OUT16>(a, v, av, n_slots, cb, lane, sc_v, gcn, lower, acc, self, add);
    }
  }
}

// ------------------------------------------------------------------------------------------
// layer 0, warp specialised.  The row kernel above is latency bound: one warp walks the dependent chain
// rowptr -> col -> scales -> Z of its row and only then runs the FMAs (4 warps per scheduler, 35 % long-scoreboard
// stalls, issue slots 46 % busy -- profiles/r01_summary.md).  Here 8 PRODUCER warps run that chain and stage weight rows
// and Z pieces into a shared-memory ring with cp.async (completion on an mbarrier), 8 CONSUMER warps only see shared
// memory: FFMA2 loop + epilogue.  Pair p = (warp p, warp p + 8) -- both on scheduler p % 4 -- owns two Z stages
// (one (row, column block, <= 32 in-edges) each) and two weight buffers (one batch each, shared by the column blocks of
// a short row), so the chain of row r + 1 and the Z pieces of the next stage are in flight while stage s is consumed.
// Same arithmetic in the same order as l0_rows_kernel: bit-identical results.
// ------------------------------------------------------------------------------------------
constexpr int kWsPairs = 8;
constexpr int kWsGrab = 8;  // rows a producer takes from the row counter at a time
enum { WS_FIRST = 1, WS_LAST = 2, WS_FREE_W = 4, WS_END = 8 };
struct WsHdr {
  int v;
  uint32_t av;
  int n, cb, wbuf, flags, pad0, pad1;
};
struct WsW {                      // one batch of <= 32 in-edges of a row
  float w[32][kL0WStride];        // weight rows, compacted to the destination's active slots
  long long off[32];              // k-th active slot -> element offset of its activation tile (slot bit x out_s_stride)
  float dv[32];                   // k-th active slot -> destination scale
  float pad[32];
};
struct WsPair {
  float z[2][32][64];             // Z pieces of a stage: 256 bytes per in-edge
  WsW wb[2];
  WsHdr hdr[2];
  uint64_t zfull[2], zempty[2], wempty[2];
  uint64_t pad[2];
};
constexpr int kWsSmemBytes = kWsPairs * (int)sizeof(WsPair);
static_assert(kWsSmemBytes <= 227 * 1024, "ring does not fit shared memory");
static_assert(sizeof(WsPair) % 16 == 0 && sizeof(WsW) % 16 == 0, "cp.async destinations must stay 16-byte aligned");

__device__ __forceinline__ uint32_t ws_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ws_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_u32(bar)) : "memory");
}
// the barrier counts one arrival of this thread once all of its cp.async issued so far have landed
__device__ __forceinline__ void ws_cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ws_u32(bar)) : "memory");
}
// > 0: waiting warps suspend inside mbarrier.try_wait (suspend-time hint in ns, woken by the hardware when the phase completes)
// instead of polling with nanosleep -- the r02 source-level profile counted 38 % of the warp-specialised layer-0 kernel's executed
// instructions in the polling loops (option l0_wait_ns)
__constant__ int g_l0_wait_hint_ns = 0;
template <int SLEEP_NS>
__device__ __forceinline__ void ws_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = ws_u32(bar);
  uint32_t done = 0;
  const uint32_t hint = (uint32_t)g_l0_wait_hint_ns;
  if (hint > 0) {
    for (uint32_t spin = 0; !done; ++spin) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(addr), "r"(parity), "r"(hint)
          : "memory");
      if (spin > (1u << 22)) __trap();
    }
    return;
  }
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done) __nanosleep(SLEEP_NS);  // a waiting warp must not eat the issue slots of the three working ones
    if (spin > (1u << 22)) __trap();  // never hang the GPU: a lost arrival becomes an error
  }
}
template <int SLEEP_NS>
__device__ __forceinline__ void ws_mbar_wait_timed(uint64_t* bar, uint32_t parity, bool timed, unsigned long long& acc) {
  if (timed) {
    const long long t0 = clock64();
    ws_mbar_wait<SLEEP_NS>(bar, parity);
    acc += (unsigned long long)(clock64() - t0);
  } else {
    ws_mbar_wait<SLEEP_NS>(bar, parity);
  }
}
constexpr int kWsProducerSleep = 256, kWsConsumerSleep = 32;  // ns between polls: the producers run ahead, the consumers are the critical path
__device__ __forceinline__ void ws_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ws_u32(bar)), "r"(bytes) : "memory");
}
// one bulk copy (TMA unit) global -> shared, size a multiple of 16 bytes, both ends 16-byte aligned
__device__ __forceinline__ void ws_bulk_copy(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ws_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(ws_u32(bar))
               : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {  // through L1: neighbouring 16-byte pieces share a sector
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(ws_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ws_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
// weight rows whose 32 raw scales already sit in shared memory (cp.async one row ahead): mask, compact in place
__device__ __forceinline__ void l0_finish_weights(float* wrow, uint32_t bits, uint32_t av, int nq) {
  float4 wq[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) wq[q] = *reinterpret_cast<const float4*>(wrow + 4 * q);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    wq[q].x = (bits >> (4 * q + 0)) & 1u ? wq[q].x : 0.0f;
    wq[q].y = (bits >> (4 * q + 1)) & 1u ? wq[q].y : 0.0f;
    wq[q].z = (bits >> (4 * q + 2)) & 1u ? wq[q].z : 0.0f;
    wq[q].w = (bits >> (4 * q + 3)) & 1u ? wq[q].w : 0.0f;
    *reinterpret_cast<float4*>(wrow + 4 * q) = wq[q];
  }
  int k = 0;
  for (uint32_t m = av; m; m &= m - 1, ++k) wrow[k] = wrow[__ffs(m) - 1];
  for (; k < 4 * nq; ++k) wrow[k] = 0.0f;
}

// table-driven epilogue of the warp-specialised kernel: scale and BYTE offset of the k-th active slot's activation tile come
// from the producer's tables (broadcast LDS.128 for four slots) instead of a FLO / SHFL / 64-bit IMAD chain per slot, the
// GCN / pre-scale switches are multiplications by uniform 0 / 1 constants (exact), and four slots share one uniform branch
// so that their dependent chains interleave.  PLAIN: single relation (no partial sums to add, always finished here).
template <bool SIGMOID, bool OUT16, bool PLAIN>
__device__ __forceinline__ void l0_epilogue_tab(const L0RowsArgs& a, const WsW& T, int v, int n_slots, int cb, int lane, float gcn1, float pre1,
                                                float lower, const float2 (&acc)[32], const float2 self, const float2 add) {
  char* outp = OUT16 ? reinterpret_cast<char*>(reinterpret_cast<__nv_bfloat16*>(a.out) + (int64_t)cb * a.out_chunk_stride + (int64_t)v * 64 + lane * 2 -
                                               (int64_t)a.b0 * a.out_s_stride)
                     : reinterpret_cast<char*>(a.out + (int64_t)(cb * 2 + (lane >> 4)) * a.out_chunk_stride + (int64_t)v * a.out_row_stride + (lane & 15) * 2 -
                                               (int64_t)a.b0 * a.out_s_stride);
  const int nq = (n_slots + 3) >> 2;
  const float pre0 = 1.0f - pre1;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (q < nq) {  // warp uniform
      const float4 dv4 = *reinterpret_cast<const float4*>(&T.dv[4 * q]);
      const longlong2 o01 = *reinterpret_cast<const longlong2*>(&T.off[4 * q]), o23 = *reinterpret_cast<const longlong2*>(&T.off[4 * q + 2]);
      const float dvs[4] = {dv4.x, dv4.y, dv4.z, dv4.w};
      const long long offs[4] = {o01.x, o01.y, o23.x, o23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = 4 * q + i;
        const float dv = dvs[i];
        const float sw = dv * gcn1, pv = fmaf(dv, pre1, pre0);  // gcn ? dv : 0, prescale ? dv : 1 (exact)
        const float2 ak = l0_acc_get(acc, k);
        float2 o;
        o.x = dv * fmaf(self.x, sw, ak.x) + add.x;
        o.y = dv * fmaf(self.y, sw, ak.y) + add.y;
        const bool on = k < n_slots;
        char* dst = outp + offs[i];
        if (!PLAIN && !OUT16 && a.accumulate && on) {  // partial sum of the relations before this one
          const float2 pr = *reinterpret_cast<const float2*>(dst);
          o.x += pr.x; o.y += pr.y;
        }
        if (PLAIN || a.finish) {
          if (SIGMOID) {
            o.x = apply_act(o.x, XPGNN_ACT_SIGMOID); o.y = apply_act(o.y, XPGNN_ACT_SIGMOID);
          } else {
            o.x = fmaxf(o.x, lower); o.y = fmaxf(o.y, lower);
          }
          o.x *= pv; o.y *= pv;
        }
        if (on) {
          if (OUT16) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(o.x, o.y);
          else __stcs(reinterpret_cast<float2*>(dst), o);
        }
      }
    }
  }
}

// BULK: every producer lane moves its in-edge's 256-byte Z piece with ONE bulk copy (cp.async.bulk = the TMA unit, SASS
// UBLKCP; completion counted in bytes on the stage's mbarrier) instead of 16 lanes x 16-byte cp.async per in-edge: the
// pieces no longer pass through the LSU / shared-memory store pipe the consumers' LDS traffic saturates.
template <bool SIGMOID, bool OUT16, bool PLAIN, bool DBG, bool BULK>
__global__ void __launch_bounds__(2 * kWsPairs * 32, 1) l0_ws_kernel(const L0RowsArgs a) {
  extern __shared__ __align__(128) uint8_t ws_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool producer = wid < kWsPairs;
  WsPair& P = reinterpret_cast<WsPair*>(ws_smem)[producer ? wid : wid - kWsPairs];
  if (producer && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      // 32 cp.async completions (one per producer lane) + the header / weights arrival | BULK: that arrival + the bytes
      ws_mbar_init(&P.zfull[i], BULK ? 1 : 33);
      ws_mbar_init(&P.zempty[i], 1);
      ws_mbar_init(&P.wempty[i], 1);
    }
    if (BULK) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t live = (a.nb == 32 ? 0xffffffffu : ((1u << a.nb) - 1u)) << a.b0;  // bits of the word in this tile
  const bool gcn = a.kind == XPGNN_CONV_GCN;
  const int ncb = a.h0 / 64;
  const bool timed = DBG && a.dbg != nullptr && blockIdx.x == 0 && (wid == 0 || wid == kWsPairs);  // DBG = false: compiled out
  unsigned long long t_wait0 = 0, t_wait1 = 0, t_fma = 0;
  const long long t_start = timed ? clock64() : 0;
  if (producer) {
    const int n_rows = a.rows ? a.n_list : a.row_hi - a.row_lo;  // row list (pruned mode) | destination range
    uint32_t zs = 0, ws = 0;  // stages / weight buffers handed over so far
    // Row iterator with the metadata one row ahead: (active word, row pointers) of the NEXT row are loaded when the
    // current row starts, its first 32 (source, edge-activity) pairs once the current row's first stage is out, so the
    // per-row chain of dependent loads shrinks to scales -> Z.
    int g_lo = 0, g_hi = 0, g_base = 0, g_row = 0;
    auto next_v = [&]() -> int {
      if (g_lo >= g_hi) {
        int base = 0;
        if (lane == 0) base = atomicAdd(a.counter, kWsGrab);
        g_lo = g_base = __shfl_sync(0xffffffffu, base, 0);
        g_hi = min(n_rows, g_lo + kWsGrab);
        if (g_lo >= n_rows) return -1;
        if (a.rows) g_row = lane < g_hi - g_lo ? __ldg(a.rows + g_lo + lane) : 0;  // the grab's list entries in one load
      }
      const int i = g_lo++;
      return a.rows ? __shfl_sync(0xffffffffu, g_row, i - g_base) : a.row_lo + i;
    };
    const int elt = OUT16 ? 2 : 4;  // bytes per stored activation element
    int v_n = next_v();
    uint32_t av_n = 0;
    int e0_n = 0, e1_n = 0, u_n = 0;
    uint32_t bits_n = 0;
    bool raw_n = false;  // the next row's raw scale rows are on their way into its weight buffer
    if (v_n >= 0) {
      av_n = a.act[(int64_t)v_n * a.W + a.w];
      e0_n = a.rowptr[v_n]; e1_n = a.rowptr[v_n + 1];
      if (lane < e1_n - e0_n) { u_n = __ldg(a.col + e0_n + lane); bits_n = __ldg(a.ebits + e0_n + lane); }
    }
    while (v_n >= 0) {
      const int v = v_n, e0 = e0_n, e1 = e1_n;
      const uint32_t av = av_n & live;
      const bool raw = raw_n;
      int u = u_n;
      const uint32_t bits0 = bits_n & av;
      v_n = next_v();
      raw_n = false;
      if (v_n >= 0) {  // in flight while this row is staged (consumed by the prefetch steps / the next iteration only)
        av_n = a.act[(int64_t)v_n * a.W + a.w];
        e0_n = a.rowptr[v_n]; e1_n = a.rowptr[v_n + 1];
      }
      const bool skip = !av || (a.long_threshold > 0 && e1 - e0 > a.long_threshold);  // nothing active / hub row (LONG launch)
      const int n_slots = __popc(av), nq = (n_slots + 3) >> 2;
      const bool short_row = e1 - e0 <= 32;
      const float sc_v = skip ? 0.0f : a.scale[(int64_t)v * 32 + lane];
      int pre_step = 0;
      // the next row, one step per posted stage: (1) its first 32 (source, edge-activity) pairs -- the row pointers have
      // landed by then; (2) the raw scale rows of those sources, cp.async into the weight buffer the row will use
      auto prefetch_next = [&](uint32_t ws_next) {
        if (v_n < 0 || pre_step >= 2) return;
        if (pre_step == 0) {
          u_n = 0; bits_n = 0;
          if (lane < e1_n - e0_n) { u_n = __ldg(a.col + e0_n + lane); bits_n = __ldg(a.ebits + e0_n + lane); }
        } else if (gcn && (av_n & live) && !(a.long_threshold > 0 && e1_n - e0_n > a.long_threshold)) {
          const int wb = ws_next & 1;
          ws_mbar_wait_timed<kWsProducerSleep>(&P.wempty[wb], ((ws_next >> 1) & 1) ^ 1, timed, t_wait0);
          if (lane < e1_n - e0_n) {
            float* wrow = &P.wb[wb].w[lane][0];
            const float* src = a.scale + (int64_t)u_n * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q) cp_async16(wrow + 4 * q, src + 4 * q);
          }
          cp_async_commit();
          raw_n = true;
        }
        ++pre_step;
      };
      if (skip) {
        prefetch_next(ws);
        prefetch_next(ws);
        continue;
      }
      int wbuf = 0;
      for (int cb = 0; cb < ncb; ++cb) {
        int base = e0;
        do {  // a row without in-edges still gets one (empty) stage per column block: self loop / bias / root term
          const int n = max(0, min(32, e1 - base));
          const bool stage_w = cb == 0 || !short_row;  // weights of the batch (kept across column blocks for short rows)
          if (stage_w) {
            wbuf = ws & 1;
            WsW& T = P.wb[wbuf];
            const bool first = base == e0 && cb == 0;
            if (first && raw) {  // acquired and filled one row ahead
              cp_async_wait0();
              if (lane < n) l0_finish_weights(&T.w[lane][0], bits0, av, nq);
            } else {
              ws_mbar_wait_timed<kWsProducerSleep>(&P.wempty[wbuf], ((ws >> 1) & 1) ^ 1, timed, t_wait0);
              if (lane < n) {
                uint32_t bits = bits0;
                if (!first) {  // beyond the prefetched first batch
                  u = __ldg(a.col + base + lane);
                  bits = __ldg(a.ebits + base + lane) & av;
                }
                l0_stage_weights(a, &T.w[lane][0], u, bits, av, nq, gcn);
              }
            }
            if ((av >> lane) & 1u) {  // tables of the epilogue: the k-th active slot's scale and byte offset
              const int k = __popc(av & ((1u << lane) - 1u));
              T.dv[k] = sc_v;
              T.off[k] = (long long)lane * a.out_s_stride * elt;
            }
            ++ws;
          }
          const int stage = zs & 1;
          ws_mbar_wait_timed<kWsProducerSleep>(&P.zempty[stage], ((zs >> 1) & 1) ^ 1, timed, t_wait1);
          if (BULK) {  // Z pieces: one 256-byte bulk copy per in-edge, issued by the lane that owns the edge
            if (lane < n) ws_bulk_copy(&P.z[stage][lane][0], a.z + (int64_t)u * a.h0 + cb * 64, 256, &P.zfull[stage]);
          } else {     // 16 lanes x 16 bytes per in-edge, two in-edges per instruction (coalesced 256-byte requests)
            const int uo = u * a.h0;  // element offset of the lane's own source row (N * h0 < 2^31: checked on the host)
            const float* zc = a.z + cb * 64 + (lane & 15) * 4;
            float* zd = &P.z[stage][lane >> 4][(lane & 15) * 4];
#pragma unroll 4
            for (int jj = 0; jj < n; jj += 2) {
              const int o = __shfl_sync(0xffffffffu, uo, (jj + (lane >> 4)) & 31);
              if (jj + (lane >> 4) < n) cp_async16_cg(zd + jj * 64, zc + o);
            }
          }
          if (!BULK) ws_cp_async_arrive(&P.zfull[stage]);
          const bool last_batch = base + 32 >= e1;
          if (lane == 0) {
            WsHdr h;
            h.v = v; h.av = av; h.n = n; h.cb = cb; h.wbuf = wbuf;
            h.flags = (base == e0 ? WS_FIRST : 0) | (last_batch ? WS_LAST : 0) | ((short_row ? cb == ncb - 1 : true) ? WS_FREE_W : 0);
            h.pad0 = h.pad1 = 0;
            P.hdr[stage] = h;
          }
          __syncwarp();
          if (lane == 0) {  // releases the header, the tables and the weight rows of all lanes
            if (BULK) ws_mbar_arrive_expect_tx(&P.zfull[stage], (uint32_t)n * 256u);
            else ws_mbar_arrive(&P.zfull[stage]);
          }
          ++zs;
          base += 32;
          // the row's remaining weight buffers come before the next row's: only prefetch into a buffer once this row
          // needs no further one (short row, or the last batch of the last column block)
          if (pre_step == 0 || short_row || (last_batch && cb == ncb - 1)) prefetch_next(ws);
        } while (base < e1);
      }
      while (pre_step < 2 && v_n >= 0) prefetch_next(ws);
    }
    const int stage = zs & 1;  // end marker
    ws_mbar_wait<kWsProducerSleep>(&P.zempty[stage], ((zs >> 1) & 1) ^ 1);
    if (lane == 0) P.hdr[stage].flags = WS_END;
    if (!BULK) ws_cp_async_arrive(&P.zfull[stage]);
    __syncwarp();
    if (lane == 0) ws_mbar_arrive(&P.zfull[stage]);
    if (timed && lane == 0) {
      a.dbg[0] = (unsigned long long)(clock64() - t_start); a.dbg[1] = t_wait0; a.dbg[2] = t_wait1; a.dbg[3] = zs;
    }
  } else {
    const float lower = a.act_fn == XPGNN_ACT_RELU ? 0.0f : -INFINITY;
    const float gcn1 = gcn ? 1.0f : 0.0f, pre1 = a.prescale ? 1.0f : 0.0f;
    float2 acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = make_float2(0.f, 0.f);
    for (uint32_t zs = 0;; ++zs) {
      const int stage = zs & 1;
      ws_mbar_wait_timed<kWsConsumerSleep>(&P.zfull[stage], (zs >> 1) & 1, timed, t_wait0);
      const WsHdr h = P.hdr[stage];
      if (h.flags & WS_END) {
        if (timed && lane == 0) {
          a.dbg[4] = (unsigned long long)(clock64() - t_start); a.dbg[5] = t_wait0; a.dbg[6] = t_fma; a.dbg[7] = zs;
        }
        break;
      }
      // epilogue operands: loads only, in flight during the FMAs (nothing may consume them before the loop)
      float2 self = make_float2(0.f, 0.f), add = self, root = self;
      if (h.flags & WS_LAST) {
        if (gcn) self = __ldg(reinterpret_cast<const float2*>(a.z + (int64_t)h.v * a.h0 + h.cb * 64 + lane * 2));
        if (a.bias && (PLAIN || a.finish)) add = __ldg(reinterpret_cast<const float2*>(a.bias + h.cb * 64 + lane * 2));
        if (a.r0c && (PLAIN || a.finish))
          root = __ldg(reinterpret_cast<const float2*>(a.r0c + (int64_t)(h.cb * 2 + (lane >> 4)) * a.r0_chunk_stride + (int64_t)h.v * a.r0_row_stride + (lane & 15) * 2));
      }
      if (h.flags & WS_FIRST) {
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = make_float2(0.f, 0.f);
      }
      const int n_slots = __popc(h.av);
      const WsW& T = P.wb[h.wbuf];
      const long long t_f0 = timed ? clock64() : 0;
      l0_fma_switch((n_slots + 3) >> 2, &T.w[0][0], &P.z[stage][0][0], h.n, lane, acc);
      if (timed) t_fma += (unsigned long long)(clock64() - t_f0);
      __syncwarp();
      if (lane == 0) ws_mbar_arrive(&P.zempty[stage]);  // before the epilogue: the next stage's pieces fly during it
      if (h.flags & WS_LAST) {
        add.x += root.x; add.y += root.y;
        l0_epilogue_tab<SIGMOID, OUT16, PLAIN>(a, T, h.v, n_slots, h.cb, lane, gcn1, pre1, lower, acc, self, add);
      }
      if (h.flags & WS_FREE_W) {
        __syncwarp();
        if (lane == 0) ws_mbar_arrive(&P.wempty[h.wbuf]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// layer 0, warp specialised, slot x column tiling (widths % 128 == 0).  The kernel above is bound by the shared-memory
// pipe (68 % busy): with a lane owning 2 columns x all active slots, every in-edge costs 2 x (4 broadcast LDS.128 of
// weights + 1 LDS.64 of Z) = 20 wavefronts.  Here a consumer lane (g, c) = (lane / 16, lane % 16) owns 8 columns
// {4c..4c+3, 64+4c..64+4c+3} x 8 of the 16 slots of a SLOT block (slot pairs g, g+2, g+4, g+6 of the block, so that few
// active slots still spread over both half warps): 2 LDS.128 of Z + NP <= 4 LDS.64 of weights + 8 NP FFMA2 per in-edge
// for all 128 columns -- 8 wavefronts.  A destination with <= 16 active slots is ONE pass over its in-edges (the usual
// case: about half of the 32 slots are active), otherwise two.  Stages hold 16 in-edges x 512 bytes.
// ------------------------------------------------------------------------------------------
constexpr int kW2EB = 16;  // in-edges per stage
struct W2Hdr {
  int v;
  uint32_t av;
  int n, cs, sb, wbuf, flags, pad;
};
struct W2W {
  float w[kW2EB][kL0WStride];     // weight rows of a batch, compacted to the destination's active slots
  long long off[32];              // k-th active slot -> byte offset of its activation tile
  float dv[32];                   // k-th active slot -> destination scale
};
struct W2Pair {
  float z[2][kW2EB][128];         // Z rows (one 128-column super block) of a stage
  W2W wb[4];                      // four: the next row's raw scales land in one while this row's batches use theirs
  W2Hdr hdr[2];
  uint64_t zfull[2], zempty[2], wempty[4];
};
constexpr int kW2SmemBytes = kWsPairs * (int)sizeof(W2Pair);
static_assert(kW2SmemBytes <= 227 * 1024, "ring does not fit shared memory");
static_assert(sizeof(W2Pair) % 16 == 0 && sizeof(W2W) % 16 == 0, "cp.async destinations must stay 16-byte aligned");

// acc[p * 8 + i] = (slot pair 2p + g of the block) x column i of the lane (i < 4: 4c + i, else 64 + 4c + i - 4)
template <int NP>
__device__ __forceinline__ void ws2_fma(const float* __restrict__ w_rows, const float* __restrict__ z_rows, int n, float2 (&acc)[32]) {
#pragma unroll(NP <= 2 ? 2 : 1)
  for (int j = 0; j < n; ++j) {
    const float4 z0 = *reinterpret_cast<const float4*>(z_rows + j * 128), z1 = *reinterpret_cast<const float4*>(z_rows + j * 128 + 64);
    const float zs[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const float2 wp = *reinterpret_cast<const float2*>(w_rows + j * kL0WStride + 4 * p);
#pragma unroll
      for (int i = 0; i < 8; ++i) ffma2(acc[p * 8 + i], wp, make_float2(zs[i], zs[i]));
    }
  }
}
__device__ __forceinline__ void ws2_fma_switch(int np, const float* wr, const float* zr, int n, float2 (&acc)[32]) {
  switch (np) {
    case 1: ws2_fma<1>(wr, zr, n, acc); break;
    case 2: ws2_fma<2>(wr, zr, n, acc); break;
    case 3: ws2_fma<3>(wr, zr, n, acc); break;
    default: ws2_fma<4>(wr, zr, n, acc); break;
  }
}

template <bool SIGMOID, bool OUT16, bool PLAIN>
__device__ __forceinline__ void ws2_epilogue(const L0RowsArgs& a, const W2W& T, int v, int n_slots, int cs, int sb, int lane, float gcn1, float pre1,
                                             float lower, const float2 (&acc)[32], const float4 (&ex)[2], const float4 (&add)[2]) {
  const int g = lane >> 4, c = lane & 15;
  char* outp[2];
#pragma unroll
  for (int b = 0; b < 2; ++b)
    outp[b] = OUT16 ? reinterpret_cast<char*>(reinterpret_cast<__nv_bfloat16*>(a.out) + (int64_t)(cs * 2 + b) * a.out_chunk_stride + (int64_t)v * 64 + 4 * c -
                                              (int64_t)a.b0 * a.out_s_stride)
                    : reinterpret_cast<char*>(a.out + (int64_t)(cs * 4 + b * 2 + (c >> 3)) * a.out_chunk_stride + (int64_t)v * a.out_row_stride + ((4 * c) & 31) -
                                              (int64_t)a.b0 * a.out_s_stride);
  const int np = (min(16, n_slots - 16 * sb) + 3) >> 2;
  const float pre0 = 1.0f - pre1;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (p < np) {  // warp uniform
      const int k0 = 16 * sb + 4 * p + 2 * g;
      const float2 dv2 = *reinterpret_cast<const float2*>(&T.dv[k0]);
      const longlong2 off2 = *reinterpret_cast<const longlong2*>(&T.off[k0]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float dv = e ? dv2.y : dv2.x;
        const long long off = e ? off2.y : off2.x;
        const bool on = k0 + e < n_slots;
        const float sw = dv * gcn1, pv = fmaf(dv, pre1, pre0);  // gcn ? dv : 0, prescale ? dv : 1 (exact)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const float se[4] = {ex[b].x, ex[b].y, ex[b].z, ex[b].w}, ad[4] = {add[b].x, add[b].y, add[b].z, add[b].w};
          float o[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 ap = acc[p * 8 + b * 4 + t];
            o[t] = dv * fmaf(se[t], sw, e ? ap.y : ap.x) + ad[t];
          }
          char* dst = outp[b] + off;
          if (!PLAIN && !OUT16 && a.accumulate && on) {  // partial sum of the relations before this one
            const float4 pr = *reinterpret_cast<const float4*>(dst);
            o[0] += pr.x; o[1] += pr.y; o[2] += pr.z; o[3] += pr.w;
          }
          if (PLAIN || a.finish) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              o[t] = SIGMOID ? apply_act(o[t], XPGNN_ACT_SIGMOID) : fmaxf(o[t], lower);
              o[t] *= pv;
            }
          }
          if (on) {
            if (OUT16) {
              __nv_bfloat162 h01 = __floats2bfloat162_rn(o[0], o[1]), h23 = __floats2bfloat162_rn(o[2], o[3]);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&h01); pk.y = *reinterpret_cast<uint32_t*>(&h23);
              *reinterpret_cast<uint2*>(dst) = pk;
            } else {
              __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
            }
          }
        }
      }
    }
  }
}

template <bool SIGMOID, bool OUT16, bool PLAIN, bool DBG>
__global__ void __launch_bounds__(2 * kWsPairs * 32, 1) l0_ws2_kernel(const L0RowsArgs a) {
  extern __shared__ __align__(128) uint8_t ws_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool producer = wid < kWsPairs;
  W2Pair& P = reinterpret_cast<W2Pair*>(ws_smem)[producer ? wid : wid - kWsPairs];
  if (producer && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ws_mbar_init(&P.zfull[i], 33);  // 32 cp.async completions (one per producer lane) + the header / weights arrival
      ws_mbar_init(&P.zempty[i], 1);
    }
    for (int i = 0; i < 4; ++i) ws_mbar_init(&P.wempty[i], 1);
  }
  __syncthreads();
  const uint32_t live = (a.nb == 32 ? 0xffffffffu : ((1u << a.nb) - 1u)) << a.b0;  // bits of the word in this tile
  const bool gcn = a.kind == XPGNN_CONV_GCN;
  const int ncs = a.h0 / 128;
  const bool timed = DBG && a.dbg != nullptr && blockIdx.x == 0 && (wid == 0 || wid == kWsPairs);  // DBG = false: compiled out
  unsigned long long t_wait0 = 0, t_wait1 = 0, t_fma = 0;
  const long long t_start = timed ? clock64() : 0;
  if (producer) {
    const int n_rows = a.row_hi - a.row_lo;
    uint32_t zs = 0, ws = 0;  // stages / weight buffers handed over so far
    int g_lo = 0, g_hi = 0;
    auto next_v = [&]() -> int {
      if (g_lo >= g_hi) {
        int base = 0;
        if (lane == 0) base = atomicAdd(a.counter, kWsGrab);
        g_lo = __shfl_sync(0xffffffffu, base, 0);
        g_hi = min(n_rows, g_lo + kWsGrab);
        if (g_lo >= n_rows) return -1;
      }
      return a.row_lo + g_lo++;
    };
    const int elt = OUT16 ? 2 : 4;  // bytes per stored activation element
    // Software pipeline over the rows, one step per iteration, so that no load is consumed in the iteration that issues it:
    //   M (row t+3) active word + row pointers -> E (row t+2) first 16 (source, edge-activity) pairs -> S (row t+1) raw scale
    //   rows of those sources by cp.async into the weight buffer the row will use + the row's own scales -> P (row t) staging.
    int m_v = -1, m_e0 = 0, m_e1 = 0;            uint32_t m_act = 0;
    int e_v = -1, e_e0 = 0, e_e1 = 0, e_u = 0;   uint32_t e_act = 0, e_bits = 0;
    int s_v = -1, s_e0 = 0, s_e1 = 0, s_u = 0;   uint32_t s_act = 0, s_bits = 0;  bool s_raw = false;  float s_sc = 0.0f;
    bool more = true;
    for (int drain = 0; drain < 3;) {
      // ---- shift ----
      const int v = s_v, e0 = s_e0, e1 = s_e1;
      int u = s_u;
      const uint32_t av = s_act & live, bits0 = s_bits & av;
      const bool raw = s_raw;
      const float sc_v = s_sc;
      s_v = e_v; s_e0 = e_e0; s_e1 = e_e1; s_u = e_u; s_act = e_act; s_bits = e_bits; s_raw = false;
      e_v = m_v; e_e0 = m_e0; e_e1 = m_e1; e_act = m_act;
      // ---- M: row t+3 ----
      m_v = more ? next_v() : -1;
      if (m_v < 0) { more = false; ++drain; }
      else {
        m_act = a.act[(int64_t)m_v * a.W + a.w];
        m_e0 = a.rowptr[m_v]; m_e1 = a.rowptr[m_v + 1];
      }
      // ---- E: row t+2 (its pointers were loaded one iteration ago) ----
      e_u = 0; e_bits = 0;
      if (e_v >= 0 && lane < min(kW2EB, e_e1 - e_e0)) { e_u = __ldg(a.col + e_e0 + lane); e_bits = __ldg(a.ebits + e_e0 + lane); }
      // ---- this row ----
      const bool valid = v >= 0 && av && !(a.long_threshold > 0 && e1 - e0 > a.long_threshold);  // else: nothing active / hub row (LONG launch)
      const int n_slots = __popc(av), nq = (n_slots + 3) >> 2, nsb = n_slots > 16 ? 2 : 1;
      const bool short_row = e1 - e0 <= kW2EB;
      const int nbuf = !valid ? 0 : (short_row ? 1 : ((e1 - e0 + kW2EB - 1) / kW2EB) * ncs * nsb);  // weight buffers this row takes
      // ---- S: row t+1 (its sources were loaded one iteration ago) ----
      const bool s_valid = s_v >= 0 && (s_act & live) && !(a.long_threshold > 0 && s_e1 - s_e0 > a.long_threshold);
      if (s_valid) s_sc = a.scale[(int64_t)s_v * 32 + lane];
      auto stage_raw = [&](uint32_t ws_next) {  // raw scale rows of row t+1's first batch into the buffer it will use
        const int wb = ws_next & 3;
        ws_mbar_wait_timed<kWsProducerSleep>(&P.wempty[wb], ((ws_next >> 2) & 1) ^ 1, timed, t_wait0);
        if (lane < min(kW2EB, s_e1 - s_e0)) {
          float* wrow = &P.wb[wb].w[lane][0];
          const float* src = a.scale + (int64_t)s_u * 32;
#pragma unroll
          for (int q = 0; q < 8; ++q) cp_async16(wrow + 4 * q, src + 4 * q);
        }
        cp_async_commit();
        s_raw = true;
      };
      // early unless the buffer it needs is one this row still has to fill (4 buffers; use k waits for the release of use k - 4)
      const bool early = gcn && s_valid && nbuf < 4;
      if (early) stage_raw(ws + nbuf);
      if (valid) {
        int wbuf = 0;
        for (int cs = 0; cs < ncs; ++cs)
        for (int sb = 0; sb < nsb; ++sb) {
          const bool last_pass = cs == ncs - 1 && sb == nsb - 1;
          int base = e0;
          do {  // a row without in-edges still gets one (empty) stage per pass: self loop / bias / root term
            const int n = max(0, min(kW2EB, e1 - base));
            const bool first = base == e0 && cs == 0 && sb == 0;
            const bool stage_w = first || !short_row;  // weights of the batch (kept across the passes of a short row)
            if (stage_w) {
              wbuf = ws & 3;
              W2W& T = P.wb[wbuf];
              if (first && raw) {  // acquired and filled one iteration ago; the group issued just now may stay in flight
                if (early) cp_async_wait1(); else cp_async_wait0();
                if (lane < n) l0_finish_weights(&T.w[lane][0], bits0, av, nq);
              } else {
                ws_mbar_wait_timed<kWsProducerSleep>(&P.wempty[wbuf], ((ws >> 2) & 1) ^ 1, timed, t_wait0);
                if (lane < n) {
                  uint32_t bits = bits0;
                  if (!first) {  // beyond the prefetched first batch
                    u = __ldg(a.col + base + lane);
                    bits = __ldg(a.ebits + base + lane) & av;
                  }
                  l0_stage_weights(a, &T.w[lane][0], u, bits, av, nq, gcn);
                }
              }
              if ((av >> lane) & 1u) {  // tables of the epilogue: the k-th active slot's scale and byte offset
                const int k = __popc(av & ((1u << lane) - 1u));
                T.dv[k] = sc_v;
                T.off[k] = (long long)lane * a.out_s_stride * elt;
              }
              ++ws;
            }
            const int stage = zs & 1;
            ws_mbar_wait_timed<kWsProducerSleep>(&P.zempty[stage], ((zs >> 1) & 1) ^ 1, timed, t_wait1);
            {  // Z rows of the super block: one coalesced 512-byte request per in-edge
              const int uo = u * a.h0;  // element offset of the lane's own source row (N * h0 < 2^31: checked on the host)
              const float* zc = a.z + cs * 128 + lane * 4;
              float* zd = &P.z[stage][0][lane * 4];
#pragma unroll 4
              for (int j = 0; j < n; ++j) cp_async16_cg(zd + j * 128, zc + __shfl_sync(0xffffffffu, uo, j));
            }
            ws_cp_async_arrive(&P.zfull[stage]);
            const bool last_batch = base + kW2EB >= e1;
            if (lane == 0) {
              W2Hdr h;
              h.v = v; h.av = av; h.n = n; h.cs = cs; h.sb = sb; h.wbuf = wbuf;
              h.flags = (base == e0 ? WS_FIRST : 0) | (last_batch ? WS_LAST : 0) | ((short_row ? last_pass : true) ? WS_FREE_W : 0);
              h.pad = 0;
              P.hdr[stage] = h;
            }
            __syncwarp();
            if (lane == 0) ws_mbar_arrive(&P.zfull[stage]);  // releases the header, the tables and the weight rows of all lanes
            ++zs;
            base += kW2EB;
          } while (base < e1);
        }
      }
      if (gcn && s_valid && !early) stage_raw(ws);  // row with many batches: its successor's buffer only now
    }
    const int stage = zs & 1;  // end marker
    ws_mbar_wait<kWsProducerSleep>(&P.zempty[stage], ((zs >> 1) & 1) ^ 1);
    if (lane == 0) P.hdr[stage].flags = WS_END;
    ws_cp_async_arrive(&P.zfull[stage]);
    __syncwarp();
    if (lane == 0) ws_mbar_arrive(&P.zfull[stage]);
    if (timed && lane == 0) {
      a.dbg[0] = (unsigned long long)(clock64() - t_start); a.dbg[1] = t_wait0; a.dbg[2] = t_wait1; a.dbg[3] = zs;
    }
  } else {
    const float lower = a.act_fn == XPGNN_ACT_RELU ? 0.0f : -INFINITY;
    const float gcn1 = gcn ? 1.0f : 0.0f, pre1 = a.prescale ? 1.0f : 0.0f;
    const int g = lane >> 4, c = lane & 15;
    float2 acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = make_float2(0.f, 0.f);
    for (uint32_t zs = 0;; ++zs) {
      const int stage = zs & 1;
      ws_mbar_wait_timed<kWsConsumerSleep>(&P.zfull[stage], (zs >> 1) & 1, timed, t_wait0);
      const W2Hdr h = P.hdr[stage];
      if (h.flags & WS_END) {
        if (timed && lane == 0) {
          a.dbg[4] = (unsigned long long)(clock64() - t_start); a.dbg[5] = t_wait0; a.dbg[6] = t_fma; a.dbg[7] = zs;
        }
        break;
      }
      // epilogue operands: loads only, in flight during the FMAs.  ex = the row's own Z (GCN self loop) | the SAGE root term
      float4 ex[2], add[2];
#pragma unroll
      for (int b = 0; b < 2; ++b) ex[b] = add[b] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (h.flags & WS_LAST) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          if (gcn) ex[b] = __ldg(reinterpret_cast<const float4*>(a.z + (int64_t)h.v * a.h0 + h.cs * 128 + b * 64 + 4 * c));
          else if (a.r0c && (PLAIN || a.finish))
            ex[b] = __ldg(reinterpret_cast<const float4*>(a.r0c + (int64_t)(h.cs * 4 + b * 2 + (c >> 3)) * a.r0_chunk_stride + (int64_t)h.v * a.r0_row_stride + ((4 * c) & 31)));
          if (a.bias && (PLAIN || a.finish)) add[b] = __ldg(reinterpret_cast<const float4*>(a.bias + h.cs * 128 + b * 64 + 4 * c));
        }
      }
      if (h.flags & WS_FIRST) {
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = make_float2(0.f, 0.f);
      }
      const int n_slots = __popc(h.av);
      const W2W& T = P.wb[h.wbuf];
      const long long t_f0 = timed ? clock64() : 0;
      ws2_fma_switch((min(16, n_slots - 16 * h.sb) + 3) >> 2, &T.w[0][16 * h.sb + 2 * g], &P.z[stage][0][4 * c], h.n, acc);
      if (timed) t_fma += (unsigned long long)(clock64() - t_f0);
      __syncwarp();
      if (lane == 0) ws_mbar_arrive(&P.zempty[stage]);  // before the epilogue: the next stage's rows fly during it
      if (h.flags & WS_LAST) {
        if (!gcn) {  // SAGE: no self term (its weight is 0), the root term joins the addend
#pragma unroll
          for (int b = 0; b < 2; ++b) { add[b].x += ex[b].x; add[b].y += ex[b].y; add[b].z += ex[b].z; add[b].w += ex[b].w; }
        }
        ws2_epilogue<SIGMOID, OUT16, PLAIN>(a, T, h.v, n_slots, h.cs, h.sb, lane, gcn1, pre1, lower, acc, ex, add);
      }
      if (h.flags & WS_FREE_W) {
        __syncwarp();
        if (lane == 0) ws_mbar_arrive(&P.wempty[h.wbuf]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// layer 0 of a HeteroConv(sum): all relations into one destination type in ONE pass over its rows.  The warp loops
// over the incoming relations of its row; the destination-side normalisation of each relation (SAGE 1 / count, GCN
// deg^-1/2) is folded into the staged weights, GCN's unit self loop is one more staged edge, so every relation adds
// into the same accumulators and the row is written once (relation by relation it was read-modify-written per relation).
// ------------------------------------------------------------------------------------------
template <bool SIGMOID>
__global__ void __launch_bounds__(256, 2) l0_multi_kernel(const L0MultiArgs a) {
  extern __shared__ __align__(16) uint8_t l0_smem[];
  float(*s_w)[32][kL0WStride] = reinterpret_cast<float(*)[32][kL0WStride]>(l0_smem);
  float(*s_z)[32][64] = reinterpret_cast<float(*)[32][64]>(l0_smem + 8 * 32 * kL0WStride * 4);
  const float*(*s_zp)[32] = reinterpret_cast<const float*(*)[32]>(l0_smem + 8 * 32 * kL0WStride * 4 + 8 * 32 * 64 * 4);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t live = (a.nb == 32 ? 0xffffffffu : ((1u << a.nb) - 1u)) << a.b0;
  const int ncb = a.h0 / 64, n_rows = a.row_hi - a.row_lo;
  const float lower = a.act_fn == XPGNN_ACT_RELU ? 0.0f : -INFINITY;
  for (int vb = grab_rows(a.counter, lane); vb < n_rows; vb = grab_rows(a.counter, lane))
  for (int v = a.row_lo + vb; v < a.row_lo + min(n_rows, vb + kRowGrab); ++v) {
    const uint32_t av = a.act[(int64_t)v * a.W + a.w] & live;
    if (!av) continue;
    const int n_slots = __popc(av), nq = (n_slots + 3) >> 2;
    for (int cb = 0; cb < ncb; ++cb) {
      float2 acc[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[k] = make_float2(0.f, 0.f);
      for (int ri = 0; ri < a.n_rel; ++ri) {
        const L0Rel& R = a.rel[ri];
        const bool gcn = R.kind == XPGNN_CONV_GCN;
        const int e0 = R.rowptr[v], e1 = R.rowptr[v + 1];
        const int n_ent = e1 - e0 + (gcn ? 1 : 0);  // GCN: its unit self loop is entry e1 - e0
        if (n_ent == 0) continue;
        // destination-side scale of this relation for every bit of the word (same addresses in all lanes: broadcast loads)
        float4 dq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) dq[q] = __ldg(reinterpret_cast<const float4*>(R.scale + (int64_t)v * 32 + q * 4));
        for (int base = 0; base < n_ent; base += 32) {
          const int n = min(32, n_ent - base);
          __syncwarp();
          if (lane < n) {
            const int ent = base + lane;
            const bool self = ent == e1 - e0;  // only for GCN
            const int u = self ? v : __ldg(R.col + e0 + ent);
            const uint32_t bits = self ? av : (__ldg(R.ebits + e0 + ent) & av);
            s_zp[wib][lane] = R.z + (int64_t)u * a.h0;
            float* wrow = &s_w[wib][lane][0];
            float4 wq[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              wq[q] = gcn ? __ldg(reinterpret_cast<const float4*>(R.scale + (int64_t)u * 32 + q * 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
              wq[q].x = (bits >> (4 * q + 0)) & 1u ? wq[q].x * dq[q].x : 0.0f;
              wq[q].y = (bits >> (4 * q + 1)) & 1u ? wq[q].y * dq[q].y : 0.0f;
              wq[q].z = (bits >> (4 * q + 2)) & 1u ? wq[q].z * dq[q].z : 0.0f;
              wq[q].w = (bits >> (4 * q + 3)) & 1u ? wq[q].w * dq[q].w : 0.0f;
              *reinterpret_cast<float4*>(wrow + 4 * q) = wq[q];
            }
            int k = 0;  // in place: position k <- bit b_k of the destination's active slots
            for (uint32_t m = av; m; m &= m - 1, ++k) wrow[k] = wrow[__ffs(m) - 1];
            for (; k < 4 * nq; ++k) wrow[k] = 0.0f;
          }
          __syncwarp();
#pragma unroll 4
          for (int j = 0; j < n; ++j) cp_async8(&s_z[wib][j][lane * 2], s_zp[wib][j] + cb * 64 + lane * 2);
          cp_async_wait_all();
          __syncwarp();
          l0_fma_switch(nq, &s_w[wib][0][0], &s_z[wib][0][0], n, lane, acc);
        }
      }
      // ---- epilogue: every normalisation is already inside the weights ----
      const int64_t cm_off = (int64_t)(cb * 2 + (lane >> 4)) * a.out_chunk_stride + (int64_t)v * 32 + (lane & 15) * 2;
      const float2 add = __ldg(reinterpret_cast<const float2*>(a.r0c + (int64_t)(cb * 2 + (lane >> 4)) * a.r0_chunk_stride + (int64_t)v * 32 + (lane & 15) * 2));
      float* outp = a.out + cm_off - (int64_t)a.b0 * a.out_s_stride;
      uint32_t m = av;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        if (k < n_slots) {  // warp uniform
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const float2 ak = l0_acc_get(acc, k);
          float2 o = make_float2(ak.x + add.x, ak.y + add.y);
          if (SIGMOID) {
            o.x = apply_act(o.x, XPGNN_ACT_SIGMOID); o.y = apply_act(o.y, XPGNN_ACT_SIGMOID);
          } else {
            o.x = fmaxf(o.x, lower); o.y = fmaxf(o.y, lower);
          }
          __stcs(reinterpret_cast<float2*>(outp + (int64_t)b * a.out_s_stride), o);
        }
      }
    }
  }
}

// layer-0 row kernel over the rows that are not hub rows: warp specialised by default, XPGNN_L0_WS=0 selects the
// one-warp-per-row kernel (same arithmetic, same order: bit-identical)
int launch_l0_rows(const L0RowsArgs& r, bool sigmoid, bool out16, int n_rows, cudaStream_t st) {
  {
    static int hint_set = 0;  // value of g_l0_wait_hint_ns on the device
    const int want = knobs().l0_wait_ns;
    if (want != hint_set) {
      XP_CHECK(cudaMemcpyToSymbolAsync(g_l0_wait_hint_ns, &want, sizeof(int), 0, cudaMemcpyHostToDevice, st));
      hint_set = want;
    }
  }
  // l0_ws option: 3 = 1 with the Z pieces moved by TMA bulk copies (UBLKCP) instead of cp.async;
  // 1 (default) warp specialised, column-block tiling; 2 warp specialised, slot x column tiling (widths % 128 == 0:
  // 40 % fewer shared-memory wavefronts and fewer cycles, but a lower clock under the power cap -- 6.2 against 5.3 ms at C3);
  // 0 one warp per row.  Read per call so that the tests can compare the three.
  const int ws_env = knobs().l0_ws;
  int ws = (int64_t)r.N * r.h0 < (int64_t(1) << 31) ? ws_env : 0;  // 32-bit source-row offsets in the producers
  if (ws == 2 && (r.h0 % 128 != 0 || r.rows)) ws = 1;  // the slot x column kernel takes whole ranges only
  const bool plain = !r.accumulate && r.finish;
  if (ws) {
#ifdef XPGNN_EXPERIMENTS
    static const bool dbg_on = getenv("XPGNN_L0_DBG") != nullptr;  // cycle counters; synchronises; diagnostics only (fp32, ReLU / none)
#else
    constexpr bool dbg_on = false;
#endif
    const bool dbg = dbg_on && !out16 && plain && !sigmoid;
    void (*k)(const L0RowsArgs);
    if (ws == 2)
      k = dbg    ? l0_ws2_kernel<false, false, true, true>
        : out16  ? (sigmoid ? l0_ws2_kernel<true, true, true, false> : l0_ws2_kernel<false, true, true, false>)
        : plain ? (sigmoid ? l0_ws2_kernel<true, false, true, false> : l0_ws2_kernel<false, false, true, false>)
                : (sigmoid ? l0_ws2_kernel<true, false, false, false> : l0_ws2_kernel<false, false, false, false>);
    else
      k = dbg    ? l0_ws_kernel<false, false, true, true, false>
        : ws == 3 ? (out16  ? (sigmoid ? l0_ws_kernel<true, true, true, false, true> : l0_ws_kernel<false, true, true, false, true>)
                    : plain ? (sigmoid ? l0_ws_kernel<true, false, true, false, true> : l0_ws_kernel<false, false, true, false, true>)
                            : (sigmoid ? l0_ws_kernel<true, false, false, false, true> : l0_ws_kernel<false, false, false, false, true>))
        : out16  ? (sigmoid ? l0_ws_kernel<true, true, true, false, false> : l0_ws_kernel<false, true, true, false, false>)
        : plain ? (sigmoid ? l0_ws_kernel<true, false, true, false, false> : l0_ws_kernel<false, false, true, false, false>)
                : (sigmoid ? l0_ws_kernel<true, false, false, false, false> : l0_ws_kernel<false, false, false, false, false>);
    const int smem = ws == 2 ? kW2SmemBytes : kWsSmemBytes;
    XP_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    L0RowsArgs rr = r;
    if (dbg) {
      XP_CHECK(cudaMalloc(&rr.dbg, 8 * sizeof(unsigned long long)));
      XP_CHECK(cudaMemsetAsync(rr.dbg, 0, 8 * sizeof(unsigned long long), st));
    }
    XP_LAUNCH(k, (int)std::min<int64_t>(std::max<int64_t>(ceil_div(n_rows, kWsPairs * kWsGrab), 1), kNumSMs), 2 * kWsPairs * 32, smem, st, rr);
    if (dbg) {
      unsigned long long h[8];
      XP_CHECK(cudaStreamSynchronize(st));
      XP_CHECK(cudaMemcpy(h, rr.dbg, sizeof h, cudaMemcpyDeviceToHost));
      XP_CHECK(cudaFree(rr.dbg));
      fprintf(stderr, "[l0_ws%d] producer total %llu wait_wempty %llu wait_zempty %llu stages %llu | consumer total %llu wait_zfull %llu fma %llu stages %llu (cycles)\n",
              ws, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    }
  } else {
    const int64_t row_groups = std::max<int64_t>(ceil_div(n_rows, 8 * kRowGrab), 1);
    void (*k)(const L0RowsArgs) = out16 ? (sigmoid ? l0_rows_kernel<true, false, true> : l0_rows_kernel<false, false, true>)
                                        : (sigmoid ? l0_rows_kernel<true, false, false> : l0_rows_kernel<false, false, false>);
    XP_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kL0SmemBytes));
    XP_LAUNCH(k, (int)std::min<int64_t>(row_groups, (int64_t)kNumSMs * 2), 256, kL0SmemBytes, st, r);
  }
  return 0;
}

// hub rows cut into slices: adds the slices' partial accumulators in slice order (deterministic) and runs the row's epilogue.
// One CTA per hub row, warp cb = column block cb; rows with a single slice were finished by l0_rows_kernel<., LONG> itself.
template <bool SIGMOID, bool OUT16>
__global__ void __launch_bounds__(128) l0_long_reduce_kernel(const L0RowsArgs a) {
  const int lane = threadIdx.x & 31, cb = threadIdx.x >> 5, ncb = a.h0 / 64, ri = blockIdx.x;
  const int i0 = a.row_item0[ri], n_slices = a.row_item0[ri + 1] - i0;
  if (n_slices <= 1 || cb >= ncb) return;
  const uint32_t live = (a.nb == 32 ? 0xffffffffu : ((1u << a.nb) - 1u)) << a.b0;
  const int v = a.long_rows[ri];
  const uint32_t av = a.act[(int64_t)v * a.W + a.w] & live;
  if (!av) return;
  const bool gcn = a.kind == XPGNN_CONV_GCN;
  const float lower = a.act_fn == XPGNN_ACT_RELU ? 0.0f : -INFINITY;
  float2 acc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) acc[k] = make_float2(0.f, 0.f);
  for (int sl = 0; sl < n_slices; ++sl) {
    const float2* sc = a.slice_scratch + ((int64_t)(i0 + sl) * ncb + cb) * 1024;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float2 p = sc[k * 32 + lane];
      acc[k].x += p.x; acc[k].y += p.y;
    }
  }
  const float sc_v = a.scale[(int64_t)v * 32 + lane];
  float2 self, add;
  l0_epilogue_operands(a, v, cb, lane, gcn, self, add);
  l0_epilogue<SIGMOID, OUT16>(a, v, av, __popc(av), cb, lane, sc_v, gcn, lower, acc, self, add);
}

// items of the sliced hub-row launch: one CTA, chunks of 256 rows with a running offset
__global__ void __launch_bounds__(256) l0_long_items_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ long_rows,
                                                            const int32_t* __restrict__ n_long_dev, int32_t* __restrict__ item_row,
                                                            int32_t* __restrict__ item_slice, int32_t* __restrict__ row_item0,
                                                            int32_t* __restrict__ n_items_dev) {
  __shared__ int s_scan[256];
  __shared__ int s_carry;
  const int n_long = *n_long_dev;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_long; base += 256) {
    const int i = base + threadIdx.x;
    int ns = 0;
    if (i < n_long) {
      const int v = long_rows[i];
      ns = (rowptr[v + 1] - rowptr[v] + kL0Slice - 1) / kL0Slice;
    }
    s_scan[threadIdx.x] = ns;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
      const int y = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
      __syncthreads();
      s_scan[threadIdx.x] += y;
      __syncthreads();
    }
    const int off = s_carry + s_scan[threadIdx.x] - ns;
    if (i < n_long) {
      row_item0[i] = off;
      for (int sl = 0; sl < ns; ++sl) { item_row[off + sl] = i; item_slice[off + sl] = sl; }
    }
    __syncthreads();
    if (threadIdx.x == 255) s_carry += s_scan[255];
    __syncthreads();
  }
  if (threadIdx.x == 0) { row_item0[n_long] = s_carry; *n_items_dev = s_carry; }
}

int build_l0_long_items(const int32_t* rowptr, const int32_t* long_rows, const int32_t* n_long_dev, int32_t* item_row, int32_t* item_slice,
                        int32_t* row_item0, int32_t* n_items_dev, cudaStream_t st) {
  XP_LAUNCH(l0_long_items_kernel, 1, 256, 0, st, rowptr, long_rows, n_long_dev, item_row, item_slice, row_item0, n_items_dev);
  return 0;
}

int launch_l0_long_rows(const L0RowsArgs& r, bool sigmoid, bool out16, int n_long, cudaStream_t st, int n_items) {
  void (*k)(const L0RowsArgs) = out16 ? (sigmoid ? l0_rows_kernel<true, true, true> : l0_rows_kernel<false, true, true>)
                                      : (sigmoid ? l0_rows_kernel<true, true, false> : l0_rows_kernel<false, true, false>);
  XP_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kL0SmemBytes));
  if (!r.item_row) {
    XP_LAUNCH(k, n_long, 256, kL0SmemBytes, st, r);
    return 0;
  }
  XP_REQUIRE(r.h0 / 64 <= 4, "sliced hub rows: at most 4 column blocks");
  if (n_items > 0) XP_LAUNCH(k, n_items, 256, kL0SmemBytes, st, r);
  if (n_items > n_long) {  // some row has several slices
    void (*kr)(const L0RowsArgs) = out16 ? (sigmoid ? l0_long_reduce_kernel<true, true> : l0_long_reduce_kernel<false, true>)
                                         : (sigmoid ? l0_long_reduce_kernel<true, false> : l0_long_reduce_kernel<false, false>);
    XP_LAUNCH(kr, n_long, 128, 0, st, r);
  }
  return 0;
}

int launch_l0_multi(const L0MultiArgs& a, bool sigmoid, cudaStream_t st) {
  void (*k)(const L0MultiArgs) = sigmoid ? l0_multi_kernel<true> : l0_multi_kernel<false>;
  XP_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kL0SmemBytes));
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(ceil_div(a.row_hi - a.row_lo, 8 * kRowGrab), 1), (int64_t)kNumSMs * 2);
  XP_LAUNCH(k, grid, 256, kL0SmemBytes, st, a);
  return 0;
}

}  // namespace xpgnn
