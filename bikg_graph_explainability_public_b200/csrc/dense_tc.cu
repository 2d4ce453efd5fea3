// Dense row transform on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// out[m][n] (+)= act( sum_k in[m][k] * w[n][k] + b[n] )      (the X*W^T transforms of the GNN layers)
//
// Two arithmetic modes:
//   TF32X3 : fp32 operands split on the fly into hi + lo TF32 parts, three MMAs per K step
//            (hi*hi + lo*hi + hi*lo).  Error ~2^-20 relative: keeps the fp32 parity bar (1e-4).
//   BF16   : operands rounded to bf16, one MMA per K step (the "bf16 transforms" mode, 2e-2 bar).
// Operands are staged in shared memory in the canonical no-swizzle K-major core-matrix layout
// (8 rows x 16 bytes per core matrix) by plain vectorised loads -- the A operand is produced by the
// masked SpMM in a fused pipeline, so TMA tensor maps would not apply to it.  One elected thread
// issues the MMAs; completion is tracked with tcgen05.commit -> mbarrier; the epilogue reads the
// accumulator tile with tcgen05.ld (32 lanes x 32 columns per warp instruction).
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "dense_args.cuh"

namespace xpgnn {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();  // never hang the GPU: a lost arrival becomes an error
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int KIND>  // 0: tf32, 1: f16 (bf16 operands)
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (KIND == 0)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 B stored contiguously
// (128 B); LBO = byte stride between core matrices adjacent in K, SBO = between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base offset 0, lbo mode 0, layout type 0 (SWIZZLE_NONE)
}
__host__ __device__ constexpr uint32_t make_idesc(int kind_fmt /*1 bf16, 2 tf32*/, int M, int N) {
  return (1u << 4) | ((uint32_t)kind_fmt << 7) | ((uint32_t)kind_fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------ kernel
constexpr int TM = 128;          // rows per tile == UMMA M == TMEM lanes
constexpr int LOADER_WARPS = 8;
constexpr int TC_THREADS = 32 * (4 + LOADER_WARPS + 1);  // warps 0-3 epilogue, 4-11 loaders / converters, 12 MMA issue
constexpr int KC = 32;           // K elements per stage
constexpr int MAX_STAGES = 4;

// MODE 0: TF32x3 (element 4 B, hi/lo copies)   MODE 1: BF16 (element 2 B)
template <int MODE>
struct TcCfg {
  static constexpr int ELT = MODE == 0 ? 4 : 2;
  static constexpr int UK = 32 / ELT;                // K elements per MMA (32 bytes)
  static constexpr int PARTS = MODE == 0 ? 2 : 1;    // hi/lo
  static constexpr int K16 = KC * ELT / 16;          // 16-byte columns of a stage row
  static constexpr int PART_BYTES = TM * KC * ELT;
  static constexpr int A_STAGE_BYTES = PART_BYTES * PARTS;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// offset (bytes) of element group (row r, 16-byte column j) inside an operand block whose K extent is
// `k16` sixteen-byte columns: core matrices contiguous along K (LBO = 128), SBO = k16 * 128
__device__ __forceinline__ uint32_t core_off(int r, int j, int k16) { return (uint32_t)((r >> 3) * k16 * 128 + j * 128 + (r & 7) * 16); }

__device__ __forceinline__ void split_tf32(const float4& v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
}
__device__ __forceinline__ uint2 pack_bf16(const float4& v) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&p0);
  r.y = *reinterpret_cast<uint32_t*>(&p1);
  return r;
}

// out of line on purpose (rare paths; the hot epilogue must stay small)
__device__ __noinline__ void sigmoid4(float4& v) {
  v.x = apply_act(v.x, XPGNN_ACT_SIGMOID); v.y = apply_act(v.y, XPGNN_ACT_SIGMOID);
  v.z = apply_act(v.z, XPGNN_ACT_SIGMOID); v.w = apply_act(v.w, XPGNN_ACT_SIGMOID);
}
__device__ __noinline__ void scalar_epilogue(const DenseArgs& a, const uint32_t (&r)[32], int c0, int64_t oo, float rs, const float* s_bias) {
#pragma unroll 1
  for (int c = 0; c < 32; ++c) {
    const int n = c0 + c;
    if (n >= a.n_out) break;
    float x = __uint_as_float(r[c]) + s_bias[n];
    float* op = a.out + oo + dense_out_off(a, n);
    if (a.accumulate) x += *op;
    *op = apply_act(x, a.act_fn) * rs;
  }
}

// out of line on purpose: called once per (tile, row) by the loaders, keeps their unrolled pipeline small
__device__ __noinline__ const float* resolve_in_row(const DenseArgs& a, int64_t tile, int r) {
  DenseRow row;
  return dense_resolve_row(a, tile, r, row) ? a.in + row.io : nullptr;
}

// Warp-specialised persistent kernel, one CTA per SM, tiles dealt round-robin:
//   loaders  (4 warps): global -> registers three K chunks ahead -> hi/lo (or bf16) operand stage in shared
//                       memory (canonical no-swizzle K-major core matrices) -> mbarrier "stage full";
//   MMA      (1 lane) : waits "stage full", issues the tcgen05.mma's of the chunk into one of TWO accumulator
//                       buffers in TMEM, commits to "stage empty" and, after the last chunk, "accumulator full";
//   epilogue (4 warps): waits "accumulator full", tcgen05.ld, bias / accumulate / activation / row scale,
//                       stores, then "accumulator empty".  It overlaps the next tile's loads and MMAs.
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) dense_tc_kernel(const DenseArgs a, int n_pad, uint32_t tmem_cols, int n_stages) {
  using Cfg = TcCfg<MODE>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES];
  __shared__ uint64_t bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_sh;
  __shared__ __align__(16) float s_bias[256];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.k;                                  // multiple of KC
  const int kb16 = K * Cfg::ELT / 16;                 // 16-byte columns of a full-K row
  const uint32_t b_part_bytes = (uint32_t)n_pad * K * Cfg::ELT;
  uint8_t* sB = smem;                                      // [PARTS][n_pad x K]
  uint8_t* sA = smem + (size_t)b_part_bytes * Cfg::PARTS;  // [n_stages][PARTS][TM x KC]

  if (warp == 0) tmem_alloc(&tmem_base_sh, tmem_cols);
  if (tid < 256) s_bias[tid] = (a.b && tid < a.n_out) ? a.b[tid] : 0.0f;
  if (tid == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&bar_full[s], 32 * LOADER_WARPS);
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_acc_full[s], 1);
      mbar_init(&bar_acc_empty[s], 128);
    }
    fence_mbar_init();
  }
  // ---- B operand (weights) resident in shared memory for the whole kernel ----
  for (int f = tid; f < n_pad * kb16; f += TC_THREADS) {
    // unit f -> (row group, 16B column): lanes of a quarter warp take the 8 rows of one core matrix
    const int r = ((f >> 3) / kb16) * 8 + (f & 7);
    const int j = (f >> 3) % kb16;
    const uint32_t off = core_off(r, j, kb16);
    if (MODE == 0) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
      if (r < a.n_out) v = __ldg(reinterpret_cast<const float4*>(a.w + (int64_t)r * K + j * 4));
      split_tf32(v, hi, lo);
      *reinterpret_cast<float4*>(sB + off) = hi;
      *reinterpret_cast<float4*>(sB + b_part_bytes + off) = lo;
    } else {
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (r < a.n_out) {
        v0 = __ldg(reinterpret_cast<const float4*>(a.w + (int64_t)r * K + j * 8));
        v1 = __ldg(reinterpret_cast<const float4*>(a.w + (int64_t)r * K + j * 8 + 4));
      }
      const uint2 p0 = pack_bf16(v0), p1 = pack_bf16(v1);
      *reinterpret_cast<uint4*>(sB + off) = make_uint4(p0.x, p0.y, p1.x, p1.y);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;
  const int n_chunks = K / KC;
  const int64_t n_tiles = a.tile_map ? (int64_t)*a.n_tiles_dev : (a.M + TM - 1) / TM;
  const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp >= 4 && warp < 4 + LOADER_WARPS) {
    // ================= loaders / converters =================
    // a warp owns 16 rows of the tile; per K chunk a thread moves 4 float4 (2 rows x 2 sixteen-byte columns)
    const int lw = warp - 4, lr = lane & 7, lj = lane >> 3;
    const int64_t n_items = my_tiles * n_chunks;
    // MODE 0: 16-byte columns lj and lj + 4; MODE 1: the adjacent pair 2 lj, 2 lj + 1 (one bf16 column)
    const int j0 = MODE == 0 ? lj : 2 * lj, j1 = MODE == 0 ? lj + 4 : 2 * lj + 1;
    // ---- load cursor: rows are resolved once per tile, when its first chunk is requested ----
    int64_t ld_tile = (int64_t)blockIdx.x - gridDim.x;
    int ld_kc = n_chunks - 1;
    const float* ld_row[2] = {nullptr, nullptr};
    auto load_next = [&](float4 (&buf)[4]) {
      if (++ld_kc == n_chunks) {
        ld_kc = 0;
        ld_tile += gridDim.x;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          ld_row[p] = resolve_in_row(a, ld_tile, lw * 16 + p * 8 + lr);
        }
      }
      const int64_t o0 = dense_in_off(a, ld_kc * KC + j0 * 4), o1 = dense_in_off(a, ld_kc * KC + j1 * 4);
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        buf[2 * p] = ld_row[p] ? __ldg(reinterpret_cast<const float4*>(ld_row[p] + o0)) : zero;
        buf[2 * p + 1] = ld_row[p] ? __ldg(reinterpret_cast<const float4*>(ld_row[p] + o1)) : zero;
      }
    };
    // ---- stage cursor ----
    int st = 0;
    uint32_t st_use = 0;  // completed rounds over the stage ring
    auto stage_next = [&](const float4 (&buf)[4]) {
      if (st_use > 0) mbar_wait(&bar_empty[st], (st_use - 1) & 1u);  // the MMAs that read this stage are done
      uint8_t* stage = sA + (size_t)st * Cfg::A_STAGE_BYTES;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int r = lw * 16 + p * 8 + lr;
        if (MODE == 0) {
          float4 hi, lo;
          split_tf32(buf[2 * p], hi, lo);
          *reinterpret_cast<float4*>(stage + core_off(r, lj, Cfg::K16)) = hi;
          *reinterpret_cast<float4*>(stage + Cfg::PART_BYTES + core_off(r, lj, Cfg::K16)) = lo;
          split_tf32(buf[2 * p + 1], hi, lo);
          *reinterpret_cast<float4*>(stage + core_off(r, lj + 4, Cfg::K16)) = hi;
          *reinterpret_cast<float4*>(stage + Cfg::PART_BYTES + core_off(r, lj + 4, Cfg::K16)) = lo;
        } else {
          const uint2 p0 = pack_bf16(buf[2 * p]), p1 = pack_bf16(buf[2 * p + 1]);
          *reinterpret_cast<uint4*>(stage + core_off(r, lj, Cfg::K16)) = make_uint4(p0.x, p0.y, p1.x, p1.y);
        }
      }
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&bar_full[st]);
      if (++st == n_stages) {
        st = 0;
        ++st_use;
      }
    };
    // three chunks in flight ahead of the one being staged; the four register buffers rotate by unrolling
    // (the bodies are kept small -- the row lookup is a call -- so that the kernel stays inside the instruction cache)
    float4 b0[4], b1[4], b2[4], b3[4];
    int64_t loaded = 0, staged = 0;
    auto step = [&](float4 (&pre)[4], const float4 (&use)[4]) {
      if (loaded < n_items) { load_next(pre); ++loaded; }
      if (staged < n_items) { stage_next(use); ++staged; }
    };
    if (loaded < n_items) { load_next(b0); ++loaded; }
    if (loaded < n_items) { load_next(b1); ++loaded; }
    if (loaded < n_items) { load_next(b2); ++loaded; }
    while (staged < n_items) {
      step(b3, b0);
      step(b0, b1);
      step(b1, b2);
      step(b2, b3);
    }
  } else if (warp == 4 + LOADER_WARPS) {
    // ================= MMA issue (one lane) =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(MODE == 0 ? 2 : 1, TM, n_pad);
      const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
      int64_t item = 0;
      for (int64_t ti = 0; ti < my_tiles; ++ti) {
        const int ab = (int)(ti & 1);
        if (ti >= 2) {  // the epilogue has drained this accumulator buffer
          mbar_wait(&bar_acc_empty[ab], (uint32_t)((ti >> 1) - 1) & 1u);
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * n_pad);
        for (int kc = 0; kc < n_chunks; ++kc, ++item) {
          const int st = (int)(item % n_stages);
          mbar_wait(&bar_full[st], (uint32_t)(item / n_stages) & 1u);
          tc_fence_after();
          const uint32_t a_hi = sA_addr + st * Cfg::A_STAGE_BYTES;
          const uint32_t a_lo = a_hi + Cfg::PART_BYTES;
#pragma unroll
          for (int ks = 0; ks < KC / Cfg::UK; ++ks) {
            // one MMA consumes 2 sixteen-byte columns (32 bytes of K)
            const uint32_t a_off = ks * 2 * 128;
            const uint32_t b_off = (kc * Cfg::K16 + ks * 2) * 128;
            const uint64_t da_hi = smem_desc(a_hi + a_off, 128, Cfg::K16 * 128);
            const uint64_t db_hi = smem_desc(sB_addr + b_off, 128, kb16 * 128);
            const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
            umma<MODE>(d_tmem, da_hi, db_hi, idesc, first);
            if (MODE == 0) {
              const uint64_t da_lo = smem_desc(a_lo + a_off, 128, Cfg::K16 * 128);
              const uint64_t db_lo = smem_desc(sB_addr + b_part_bytes + b_off, 128, kb16 * 128);
              umma<MODE>(d_tmem, da_lo, db_hi, idesc, 1u);
              umma<MODE>(d_tmem, da_hi, db_lo, idesc, 1u);
            }
          }
          umma_commit(&bar_empty[st]);
          if (kc == n_chunks - 1) umma_commit(&bar_acc_full[ab]);
        }
      }
    }
  } else if (warp < 4) {
    // ================= epilogue: TMEM -> registers -> bias / accumulate / activation / row scale -> global =================
    const int row_in_tile = warp * 32 + lane;  // TMEM lane == tile row; warp w may touch lanes 32 w .. 32 w + 31
    // ReLU and identity share one branch-free path (max with 0 or -inf); sigmoid is a separate loop
    const float lower = a.act_fn == XPGNN_ACT_RELU ? 0.0f : -INFINITY;
    const bool vec_ok = (a.n_out & 3) == 0 && (a.ld_out & 3) == 0 && (a.out_s_stride & 3) == 0 && (a.out_chunk_stride & 3) == 0;
    for (int64_t ti = 0; ti < my_tiles; ++ti) {
      const int ab = (int)(ti & 1);
      const int64_t tile = (int64_t)blockIdx.x + ti * gridDim.x;
      DenseRow row;
      const bool ok = dense_resolve_row(a, tile, row_in_tile, row);
      const int64_t oo = row.oo;
      const float rs = row.rs;
      mbar_wait(&bar_acc_full[ab], (uint32_t)(ti >> 1) & 1u);
      tc_fence_after();
      for (int c0 = 0; c0 < n_pad; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ab * n_pad + c0), r);
        if (!ok) continue;
        if (vec_ok) {
          float4 prev[8];
          if (a.accumulate) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              prev[c] = (c0 + 4 * c < a.n_out) ? *reinterpret_cast<const float4*>(a.out + oo + dense_out_off(a, c0 + 4 * c))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const int n = c0 + c;
            if (n >= a.n_out) break;
            const float4 bb = *reinterpret_cast<const float4*>(s_bias + n);
            float4 v = make_float4(__uint_as_float(r[c]) + bb.x, __uint_as_float(r[c + 1]) + bb.y, __uint_as_float(r[c + 2]) + bb.z,
                                   __uint_as_float(r[c + 3]) + bb.w);
            if (a.accumulate) {
              v.x += prev[c >> 2].x; v.y += prev[c >> 2].y; v.z += prev[c >> 2].z; v.w += prev[c >> 2].w;
            }
            if (a.act_fn == XPGNN_ACT_SIGMOID) {
              sigmoid4(v);
            } else {
              v.x = fmaxf(v.x, lower); v.y = fmaxf(v.y, lower); v.z = fmaxf(v.z, lower); v.w = fmaxf(v.w, lower);
            }
            v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
            *reinterpret_cast<float4*>(a.out + oo + dense_out_off(a, n)) = v;
          }
        } else {
          scalar_epilogue(a, r, c0, oo, rs, s_bias);
        }
      }
      tc_fence_before();
      mbar_arrive(&bar_acc_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------ host
static size_t tc_b_bytes(const DenseArgs& d, int mode) {
  const int n_pad = (d.n_out + 15) / 16 * 16;
  return (size_t)n_pad * d.k * (mode == 0 ? 8 : 2);
}
static int tc_stages(const DenseArgs& d, int mode) {
  const size_t stage = mode == 0 ? TcCfg<0>::A_STAGE_BYTES : TcCfg<1>::A_STAGE_BYTES;
  const size_t budget = 200 * 1024;
  const size_t b = tc_b_bytes(d, mode);
  if (b + 2 * stage > budget) return 0;
  return (int)std::min<size_t>(MAX_STAGES, (budget - b) / stage);
}

static bool tc_eligible(const DenseArgs& d, int mode) {
  if (d.k <= 0 || d.k % KC != 0 || d.n_out < 8 || d.n_out > 256) return false;
  if (d.M >= (1ll << 31) - 256 || d.rows_per_s <= 0) return false;
  if (d.ld_in % 4 != 0 || d.in_s_stride % 4 != 0 || d.in_chunk_stride % 4 != 0 || ((uintptr_t)d.in & 15) != 0 || ((uintptr_t)d.w & 15) != 0)
    return false;
  return tc_stages(d, mode) >= 2;
}

int launch_dense_tc(const DenseArgs& d, int mode, cudaStream_t st) {
  XP_REQUIRE(tc_eligible(d, mode), "shape not eligible for the tensor-core dense path");
  const int n_pad = (d.n_out + 15) / 16 * 16;
  uint32_t cols = 32;
  while ((int)cols < 2 * n_pad) cols <<= 1;  // two accumulator buffers
  const int stages = tc_stages(d, mode);
  const size_t smem = tc_b_bytes(d, mode) + (size_t)stages * (mode == 0 ? TcCfg<0>::A_STAGE_BYTES : TcCfg<1>::A_STAGE_BYTES);
  const int64_t tiles = ceil_div(d.M, TM);
  const int grid = (int)std::min<int64_t>(tiles, kNumSMs);
  if (mode == 0) {
    XP_CHECK(cudaFuncSetAttribute(dense_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XP_LAUNCH(dense_tc_kernel<0>, grid, TC_THREADS, smem, st, d, n_pad, cols, stages);
  } else {
    XP_CHECK(cudaFuncSetAttribute(dense_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XP_LAUNCH(dense_tc_kernel<1>, grid, TC_THREADS, smem, st, d, n_pad, cols, stages);
  }
  return 0;
}

bool dense_tc_eligible(const DenseArgs& d, int mode) { return tc_eligible(d, mode); }

}  // namespace xpgnn
