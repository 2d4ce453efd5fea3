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
constexpr int TM = 128;       // rows per tile == UMMA M == TMEM lanes
constexpr int TC_THREADS = 256;
constexpr int KC_BYTES = 128; // bytes of K per stage row (32 fp32 or 64 bf16): 8 core matrices along K

// MODE 0: TF32x3 (element 4 B, hi/lo copies)   MODE 1: BF16 (element 2 B)
template <int MODE>
struct TcCfg {
  static constexpr int ELT = MODE == 0 ? 4 : 2;
  static constexpr int KC = KC_BYTES / ELT;          // K elements per stage
  static constexpr int UK = 32 / ELT;                // K elements per MMA (32 bytes)
  static constexpr int PARTS = MODE == 0 ? 2 : 1;    // hi/lo
  static constexpr int A_STAGE_BYTES = TM * KC_BYTES * PARTS;
  static constexpr int STAGES = 2;
};

// offset (bytes) of element group (row r, 16-byte column j) inside an operand block whose K extent is
// `k16` sixteen-byte columns: core matrices contiguous along K (LBO = 128), SBO = k16 * 128
__device__ __forceinline__ uint32_t core_off(int r, int j, int k16) { return (uint32_t)((r >> 3) * k16 * 128 + j * 128 + (r & 7) * 16); }

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) dense_tc_kernel(const DenseArgs a, int n_pad, uint32_t tmem_cols) {
  using Cfg = TcCfg<MODE>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_stage[Cfg::STAGES];
  __shared__ uint64_t bar_acc;
  __shared__ uint32_t tmem_base_sh;
  __shared__ int64_t out_off[TM];
  __shared__ float row_scale[TM];
  __shared__ __align__(16) float s_bias[256];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.k;                                  // multiple of Cfg::KC
  const int kb16 = K * Cfg::ELT / 16;                 // 16-byte columns of a full-K row
  const uint32_t b_part_bytes = (uint32_t)n_pad * K * Cfg::ELT;
  uint8_t* sB = smem;                                  // [PARTS][n_pad x K]
  uint8_t* sA = smem + (size_t)b_part_bytes * Cfg::PARTS;  // [STAGES][PARTS][TM x KC]

  if (warp == 0) tmem_alloc(&tmem_base_sh, tmem_cols);
  s_bias[tid] = (a.b && tid < a.n_out) ? a.b[tid] : 0.0f;  // TC_THREADS == 256 >= n_pad
  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) mbar_init(&bar_stage[s], 1);
    mbar_init(&bar_acc, 1);
    fence_mbar_init();
  }
  // ---- B operand (weights) resident in shared memory for the whole kernel ----
  for (int f = tid; f < n_pad * kb16; f += TC_THREADS) {
    // unit f -> (row group, 16B column): lanes of a quarter warp take the 8 rows of one core matrix
    const int r = ((f >> 3) / kb16) * 8 + (f & 7);
    const int j = (f >> 3) % kb16;
    const uint32_t off = core_off(r, j, kb16);
    if (MODE == 0) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < a.n_out) v = __ldg(reinterpret_cast<const float4*>(a.w + (int64_t)r * K + j * 4));
      float4 hi, lo;
      hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
      hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
      hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
      hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
      *reinterpret_cast<float4*>(sB + off) = hi;
      *reinterpret_cast<float4*>(sB + b_part_bytes + off) = lo;
    } else {
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (r < a.n_out) {
        v0 = __ldg(reinterpret_cast<const float4*>(a.w + (int64_t)r * K + j * 8));
        v1 = __ldg(reinterpret_cast<const float4*>(a.w + (int64_t)r * K + j * 8 + 4));
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v0.x, v0.y), p1 = __floats2bfloat162_rn(v0.z, v0.w);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(v1.x, v1.y), p3 = __floats2bfloat162_rn(v1.z, v1.w);
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
      pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
      *reinterpret_cast<uint4*>(sB + off) = pk;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;
  const uint32_t idesc = make_idesc(MODE == 0 ? 2 : 1, TM, n_pad);
  const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
  const int n_chunks = K / Cfg::KC;
  const int64_t n_tiles = a.tile_map ? (int64_t)*a.n_tiles_dev : (a.M + TM - 1) / TM;
  uint32_t uses0 = 0u, uses1 = 0u;  // per-stage use counters (scalars: a dynamically indexed array would live in local memory)
  uint32_t chunk_ctr = 0, tile_ctr = 0;

  // loader mapping: per warp instruction 8 rows x 4 sixteen-byte columns; 32 units of (row group, half)
  const int lr = lane & 7, lj = lane >> 3;
  constexpr int NV = MODE == 0 ? 4 : 8;  // float4 registers per thread per K chunk

  // global -> registers for one (tile, K chunk); rows beyond M / outside the destination range read as 0
  auto load_item = [&](int64_t tile, int kc, float4 (&buf)[NV]) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int unit = warp * 4 + it;
      const int r = (unit >> 1) * 8 + lr;
      const int j = (unit & 1) * 4 + lj;
      DenseRow row;
      const bool ok = dense_resolve_row(a, tile, r, row);
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 0) {
        buf[it] = ok ? __ldg(reinterpret_cast<const float4*>(a.in + row.io + dense_in_off(a, kc * Cfg::KC + j * 4))) : zero;
      } else {
        buf[2 * it] = ok ? __ldg(reinterpret_cast<const float4*>(a.in + row.io + dense_in_off(a, kc * Cfg::KC + j * 8))) : zero;
        buf[2 * it + 1] = ok ? __ldg(reinterpret_cast<const float4*>(a.in + row.io + dense_in_off(a, kc * Cfg::KC + j * 8 + 4))) : zero;
      }
    }
  };
  // item i (flattened over this CTA's tiles and K chunks) -> (tile, kc); tile < 0 when past the end
  auto item_tile = [&](int64_t i) -> int64_t {
    const int64_t t = (int64_t)blockIdx.x + (i / n_chunks) * gridDim.x;
    return t < n_tiles ? t : -1;
  };

  // Two chunks are always in flight ahead of the one being staged (1 CTA/SM: latency hiding is explicit).
  // The three register buffers rotate by unrolling, not by moves -- a move would wait for the load.
  float4 b0[NV], b1[NV], b2[NV];
  int64_t item = 0, tile = blockIdx.x;
  int kc = 0;
  if (item_tile(0) >= 0) load_item(item_tile(0), 0, b0);
  if (item_tile(1) >= 0) load_item(item_tile(1), 1 % n_chunks, b1);

  auto step = [&](float4 (&use)[NV], float4 (&pre)[NV]) -> bool {
    if (tile >= n_tiles) return false;
    if (kc == 0 && tid < TM) {
      DenseRow row;
      const bool ok = dense_resolve_row(a, tile, tid, row);
      out_off[tid] = ok ? row.oo : -1;
      row_scale[tid] = row.rs;
    }
    {
      const int64_t t2 = item_tile(item + 2);
      if (t2 >= 0) load_item(t2, (int)((item + 2) % n_chunks), pre);
    }
    const int st = chunk_ctr & 1;
    const uint32_t used = st ? uses1 : uses0;
    if (used > 0) mbar_wait(&bar_stage[st], (used - 1) & 1);  // MMAs that read this stage are done
    uint8_t* stage = sA + (size_t)st * Cfg::A_STAGE_BYTES;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int unit = warp * 4 + it;
      const int r = (unit >> 1) * 8 + lr;
      const int j = (unit & 1) * 4 + lj;
      const uint32_t off = core_off(r, j, 8);
      if (MODE == 0) {
        const float4 v = use[it];
        float4 hi, lo;
        hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
        hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
        hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
        hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
        *reinterpret_cast<float4*>(stage + off) = hi;
        *reinterpret_cast<float4*>(stage + TM * KC_BYTES + off) = lo;
      } else {
        const float4 v0 = use[(2 * it) % NV], v1 = use[(2 * it + 1) % NV];
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v0.x, v0.y), p1 = __floats2bfloat162_rn(v0.z, v0.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v1.x, v1.y), p3 = __floats2bfloat162_rn(v1.z, v1.w);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
        pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
        *reinterpret_cast<uint4*>(stage + off) = pk;
      }
    }
    fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_hi = sA_addr + st * Cfg::A_STAGE_BYTES;
      const uint32_t a_lo = a_hi + TM * KC_BYTES;
#pragma unroll
      for (int ks = 0; ks < Cfg::KC / Cfg::UK; ++ks) {
        // one MMA consumes 2 sixteen-byte columns (32 bytes of K)
        const uint32_t a_off = ks * 2 * 128;
        const uint32_t b_off = (kc * (KC_BYTES / 16) + ks * 2) * 128;
        const uint64_t da_hi = smem_desc(a_hi + a_off, 128, 8 * 128);
        const uint64_t db_hi = smem_desc(sB_addr + b_off, 128, kb16 * 128);
        const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
        umma<MODE>(tmem_base, da_hi, db_hi, idesc, first);
        if (MODE == 0) {
          const uint64_t da_lo = smem_desc(a_lo + a_off, 128, 8 * 128);
          const uint64_t db_lo = smem_desc(sB_addr + b_part_bytes + b_off, 128, kb16 * 128);
          umma<MODE>(tmem_base, da_lo, db_hi, idesc, 1u);
          umma<MODE>(tmem_base, da_hi, db_lo, idesc, 1u);
        }
      }
      umma_commit(&bar_stage[st]);
      if (kc == n_chunks - 1) umma_commit(&bar_acc);
    }
    if (st) ++uses1; else ++uses0;
    ++chunk_ctr;
    ++item;
    if (kc != n_chunks - 1) {
      ++kc;
      return true;
    }
    // ---- epilogue: TMEM -> registers -> bias / accumulate / activation -> global ----
    mbar_wait(&bar_acc, tile_ctr & 1);
    tc_fence_after();
    {
      const int row = (warp & 3) * 32 + lane;          // TMEM lane == tile row; warp w may touch lanes 32*(w%4)..
      const int64_t oo = out_off[row];
      const float rs = row_scale[row];
      // ReLU and identity share one branch-free path (max with 0 or -inf); sigmoid is a separate loop
      const float lower = a.act_fn == XPGNN_ACT_RELU ? 0.0f : -INFINITY;
      const bool vec_ok = (a.n_out & 3) == 0 && (a.ld_out & 3) == 0 && (a.out_s_stride & 3) == 0 && (a.out_chunk_stride & 3) == 0;
      for (int c0 = (warp >> 2) * 32; c0 < n_pad; c0 += 64) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, r);
        if (oo < 0) continue;
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
              v.x = apply_act(v.x, XPGNN_ACT_SIGMOID); v.y = apply_act(v.y, XPGNN_ACT_SIGMOID);
              v.z = apply_act(v.z, XPGNN_ACT_SIGMOID); v.w = apply_act(v.w, XPGNN_ACT_SIGMOID);
            } else {
              v.x = fmaxf(v.x, lower); v.y = fmaxf(v.y, lower); v.z = fmaxf(v.z, lower); v.w = fmaxf(v.w, lower);
            }
            v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
            *reinterpret_cast<float4*>(a.out + oo + dense_out_off(a, n)) = v;
          }
        } else {
          for (int c = 0; c < 32; ++c) {
            const int n = c0 + c;
            if (n >= a.n_out) break;
            float x = __uint_as_float(r[c]) + s_bias[n];
            float* op = a.out + oo + dense_out_off(a, n);
            if (a.accumulate) x += *op;
            *op = apply_act(x, a.act_fn) * rs;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // accumulator tile and the row offsets may be reused
    kc = 0;
    tile += gridDim.x;
    ++tile_ctr;
    return true;
  };
  while (true) {
    if (!step(b0, b2)) break;
    if (!step(b1, b0)) break;
    if (!step(b2, b1)) break;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------ host
static bool tc_eligible(const DenseArgs& d, int mode) {
  const int kc = mode == 0 ? 32 : 64;
  if (d.k <= 0 || d.k % kc != 0 || d.n_out < 8 || d.n_out > 256) return false;
  if (d.M >= (1ll << 31) - 256 || d.rows_per_s <= 0) return false;
  if (d.ld_in % 4 != 0 || d.in_s_stride % 4 != 0 || d.in_chunk_stride % 4 != 0 || ((uintptr_t)d.in & 15) != 0 || ((uintptr_t)d.w & 15) != 0)
    return false;
  const int n_pad = (d.n_out + 15) / 16 * 16;
  const size_t smem = (size_t)n_pad * d.k * (mode == 0 ? 8 : 2) + (size_t)TcCfg<0>::STAGES * TM * KC_BYTES * (mode == 0 ? 2 : 1);
  return smem <= 200 * 1024;
}

int launch_dense_tc(const DenseArgs& d, int mode, cudaStream_t st) {
  XP_REQUIRE(tc_eligible(d, mode), "shape not eligible for the tensor-core dense path");
  const int n_pad = (d.n_out + 15) / 16 * 16;
  uint32_t cols = 32;
  while ((int)cols < n_pad) cols <<= 1;
  const size_t smem = (size_t)n_pad * d.k * (mode == 0 ? 8 : 2) + (size_t)2 * TM * KC_BYTES * (mode == 0 ? 2 : 1);
  const int64_t tiles = ceil_div(d.M, TM);
  const int grid = (int)std::min<int64_t>(tiles, kNumSMs);
  if (mode == 0) {
    XP_CHECK(cudaFuncSetAttribute(dense_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XP_LAUNCH(dense_tc_kernel<0>, grid, TC_THREADS, smem, st, d, n_pad, cols);
  } else {
    XP_CHECK(cudaFuncSetAttribute(dense_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XP_LAUNCH(dense_tc_kernel<1>, grid, TC_THREADS, smem, st, d, n_pad, cols);
  }
  return 0;
}

bool dense_tc_eligible(const DenseArgs& d, int mode) { return tc_eligible(d, mode); }

}  // namespace xpgnn
