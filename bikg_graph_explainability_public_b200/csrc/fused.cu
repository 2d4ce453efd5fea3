// Fused masked SpMM + dense transform for conv layers >= 1 (aggregate-first):
//
//   out_s[v] = act( W * ( dinv_s[v] * ( sum_e a_e dinv_s[u] H_s[u] + dinv_s[v] H_s[v] ) ) + b )      (GCN)
//   out_s[v] = act( W * ( inv_s[v]  *   sum_e a_e H_s[u] ) + b )                                    (SAGE, no root)
//
// One CTA builds a 128-row A tile (RB node rows x SB coalition slots) in shared memory: its 16 warps
// gather the active in-neighbours' activation rows (128-bit loads, fp32 accumulation), split every
// aggregated row into TF32 hi + lo parts and store them in the canonical K-major core-matrix layout.
// One thread then issues the tcgen05 MMAs (hi*hi + lo*hi + hi*lo per K step, fp32-grade accuracy)
// against W chunks that a bulk async copy (UBLKCP, mbarrier complete_tx) streams from a pre-formatted
// L2-resident image, accumulators in TMEM; the epilogue reads them with tcgen05.ld, adds the bias, applies
// the activation, transposes through shared memory and writes full 512-byte rows.
// The aggregate tile never goes to HBM (the unfused path writes and re-reads 16 GB per 32-coalition tile).
#include "common.cuh"
#include "dense_args.cuh"
#include "fused.cuh"

namespace xpgnn {

// ------------------------------------------------------------------ PTX wrappers (see dense_tc.cu)
__device__ __forceinline__ uint32_t f_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void f_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = f_smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void f_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(f_smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(f_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void f_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void f_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void f_tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void f_tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void f_tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(f_smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void f_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void f_umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void f_umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(f_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void f_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ uint64_t f_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// ------------------------------------------------------------------ geometry
constexpr int FM = 128;                 // A rows per tile == UMMA M == TMEM lanes
constexpr int F_THREADS = 512;          // 16 warps: 8 (row, slot) pairs each
constexpr int A_LBO = 144;              // padded stride between K-adjacent core matrices: conflict-free row stores
constexpr int W_CHUNK_K = 32;           // K elements (fp32) per streamed W chunk = 8 sixteen-byte columns
constexpr int W_STAGES = 2;

__host__ __device__ inline int fused_w_chunk_bytes(int n_pad) { return 2 * n_pad * W_CHUNK_K * 4; }  // hi + lo

// W [n_out][K] fp32 -> image [K/32 chunks][hi|lo][n_pad x 32] in core-matrix layout (LBO 128, SBO 1024)
__global__ void __launch_bounds__(256) fused_w_image_kernel(const float* __restrict__ w, int n_out, int n_pad, int K, float* __restrict__ img) {
  const int chunks = K / W_CHUNK_K;
  const int total = chunks * n_pad * W_CHUNK_K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i / (n_pad * W_CHUNK_K);
    const int rem = i - c * n_pad * W_CHUNK_K;
    const int n = rem / W_CHUNK_K, kk = rem % W_CHUNK_K;
    const float v = n < n_out ? w[(int64_t)n * K + c * W_CHUNK_K + kk] : 0.0f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const float lo = v - hi;
    const int off = (n >> 3) * 8 * 32 + (kk >> 2) * 32 + (n & 7) * 4 + (kk & 3);  // in floats
    float* base = img + (int64_t)c * (2 * n_pad * W_CHUNK_K);
    base[off] = hi;
    base[n_pad * W_CHUNK_K + off] = lo;
  }
}

// ------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(F_THREADS, 1) spmm_dense_fused_kernel(const FusedArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[W_STAGES], bar_free[W_STAGES], bar_acc;
  __shared__ uint32_t tmem_base_sh;
  __shared__ __align__(16) float s_bias[128];
  __shared__ int64_t s_out_off[FM];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, k16 = K / 4;                       // sixteen-byte columns of an A row
  const uint32_t a_sbo = (uint32_t)k16 * A_LBO;         // stride between 8-row groups of the A tile
  const uint32_t a_part = 16u * a_sbo;                  // bytes of one A part (hi or lo)
  const uint32_t w_chunk = (uint32_t)fused_w_chunk_bytes(a.n_pad);
  uint8_t* sA = smem;                                   // [hi | lo]
  uint8_t* sW = smem + a.w_off;                         // [W_STAGES][hi | lo]
  float* sT = reinterpret_cast<float*>(smem);           // epilogue transpose tile aliases the A tile
  const int n_chunks = K / W_CHUNK_K;

  if (warp == 0) f_tmem_alloc(&tmem_base_sh, 128);
  if (tid == 0) {
    for (int s = 0; s < W_STAGES; ++s) {
      f_mbar_init(&bar_full[s], 1);
      f_mbar_init(&bar_free[s], 1);
    }
    f_mbar_init(&bar_acc, 1);
    f_fence_mbar_init();
  }
  if (tid < 128) s_bias[tid] = (a.b && tid < a.n_out) ? a.b[tid] : 0.0f;
  f_tc_before();
  __syncthreads();
  f_tc_after();
  const uint32_t tmem_base = tmem_base_sh;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.n_pad >> 3) << 17) | ((uint32_t)(FM >> 4) << 24);
  const uint32_t sA_addr = f_smem_u32(sA), sW_addr = f_smem_u32(sW);

  // W chunk stream (thread 0): chunk index g runs over all tiles of this CTA; stage = g & 1
  uint32_t w_issued = 0, w_consumed = 0;  // thread 0 only
  const int SB = a.SB, RB = FM / SB;                    // slots x node rows per tile
  const int n_rb = (a.n_rows + RB - 1) / RB;
  const int n_sb = (a.n_bits + SB - 1) / SB;
  const int64_t n_tiles = (int64_t)n_rb * n_sb;
  const int64_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t w_total = (uint32_t)(my_tiles * n_chunks);  // never stream more chunks than will be consumed
  auto issue_w = [&]() {  // stage must be free
    const uint32_t st = w_issued & 1u;
    f_mbar_expect_tx(&bar_full[st], w_chunk);
    f_bulk_g2s(sW + (size_t)st * w_chunk, reinterpret_cast<const uint8_t*>(a.w_image) + (size_t)(w_issued % n_chunks) * w_chunk,
               w_chunk, &bar_full[st]);
    ++w_issued;
  };
  if (tid == 0) {
    if (w_issued < w_total) issue_w();
    if (w_issued < w_total) issue_w();
  }
  uint32_t tile_ctr = 0;

  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_ctr) {
    const int rb = a.slot_major ? (int)(t % n_rb) : (int)(t / n_sb);
    const int sb = a.slot_major ? (int)(t / n_rb) : (int)(t % n_sb);
    // ------------------------------------------------ gather phase: 16 warps x 8 (row, slot) pairs
    // SB >= 8 (host guarantees it): the 8 pairs of a warp share ONE node row and differ in the coalition slot,
    // so the row's column indices / activity words are loaded once and the 8 slots advance in lockstep with
    // 8 independent 512-byte gathers in flight per warp (1 CTA/SM: memory-level parallelism must be explicit).
    {
      const int p0 = warp * 8;
      const int rl = p0 / SB, sl0 = p0 - rl * SB;
      const int r = rb * RB + rl;
      const bool row_ok = r < a.n_rows;
      const int v = a.row_lo + r;
      float4 acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      int e0 = 0, e1 = 0;
      if (row_ok) {
        e0 = a.rowptr[v];
        e1 = a.rowptr[v + 1];
      }
      const int s_base = sb * SB + sl0;                 // first slot of this warp inside the tile
      const int bbase = a.b0 + s_base;
      for (int base = e0; base < e1; base += 32) {
        const int e = base + lane;
        int u_l = -1;
        uint32_t bits_l = 0;
        if (e < e1) {
          u_l = __ldg(a.col + e);
          bits_l = __ldg(a.ebits + e) >> bbase;         // bit i: edge active in slot s_base + i
        }
        uint32_t m[8];
        uint32_t any = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          m[i] = (s_base + i < a.n_bits) ? __ballot_sync(0xffffffffu, (bits_l >> i) & 1u) : 0u;
          any |= m[i];
        }
        while (any) {
          int u[8];
          float kq[8];
          float4 x[8];
          any = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int l = m[i] ? __ffs(m[i]) - 1 : 0;
            u[i] = __shfl_sync(0xffffffffu, u_l, l);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            kq[i] = 0.f;
            x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m[i]) {
              kq[i] = a.kind == XPGNN_CONV_GCN ? __ldg(a.scale + (int64_t)u[i] * 32 + bbase + i) : 1.0f;
              if (lane < k16)
                x[i] = __ldg(reinterpret_cast<const float4*>(a.in + (int64_t)(s_base + i) * a.in_s_stride + (int64_t)u[i] * a.ld_in) + lane);
            }
            m[i] &= m[i] - 1;
            any |= m[i];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[i].x = fmaf(kq[i], x[i].x, acc[i].x);
            acc[i].y = fmaf(kq[i], x[i].y, acc[i].y);
            acc[i].z = fmaf(kq[i], x[i].z, acc[i].z);
            acc[i].w = fmaf(kq[i], x[i].w, acc[i].w);
          }
        }
      }
      // finalise: normalisation (+ GCN self loop), TF32 hi/lo split, store to the A tile
      float4 self[8];
      float dv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool valid = row_ok && (s_base + i < a.n_bits);
        dv[i] = valid ? __ldg(a.scale + (int64_t)v * 32 + bbase + i) : 0.f;
        self[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && a.kind == XPGNN_CONV_GCN && lane < k16)
          self[i] = __ldg(reinterpret_cast<const float4*>(a.in + (int64_t)(s_base + i) * a.in_s_stride + (int64_t)v * a.ld_in) + lane);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int p = p0 + i;
        const bool valid = row_ok && (s_base + i < a.n_bits);
        float4 o;
        if (a.kind == XPGNN_CONV_GCN) {
          o.x = dv[i] * fmaf(dv[i], self[i].x, acc[i].x);
          o.y = dv[i] * fmaf(dv[i], self[i].y, acc[i].y);
          o.z = dv[i] * fmaf(dv[i], self[i].z, acc[i].z);
          o.w = dv[i] * fmaf(dv[i], self[i].w, acc[i].w);
        } else {
          o.x = dv[i] * acc[i].x; o.y = dv[i] * acc[i].y; o.z = dv[i] * acc[i].z; o.w = dv[i] * acc[i].w;
        }
        if (lane == 0) s_out_off[p] = valid ? (int64_t)(s_base + i) * a.out_s_stride + (int64_t)v * a.ld_out : -1;
        if (lane < k16) {  // lane == sixteen-byte column of the row
          float4 hi, lo;
          hi.x = __uint_as_float(__float_as_uint(o.x) & 0xffffe000u); lo.x = o.x - hi.x;
          hi.y = __uint_as_float(__float_as_uint(o.y) & 0xffffe000u); lo.y = o.y - hi.y;
          hi.z = __uint_as_float(__float_as_uint(o.z) & 0xffffe000u); lo.z = o.z - hi.z;
          hi.w = __uint_as_float(__float_as_uint(o.w) & 0xffffe000u); lo.w = o.w - hi.w;
          const uint32_t off = (uint32_t)(p >> 3) * a_sbo + (uint32_t)lane * A_LBO + (uint32_t)(p & 7) * 16;
          *reinterpret_cast<float4*>(sA + off) = hi;
          *reinterpret_cast<float4*>(sA + a_part + off) = lo;
        }
      }
    }
    f_fence_proxy_async();
    f_tc_before();
    __syncthreads();
    // ------------------------------------------------ MMA phase (one thread)
    if (tid == 0) {
      f_tc_after();
      for (int c = 0; c < n_chunks; ++c, ++w_consumed) {
        const uint32_t st = w_consumed & 1u;
        f_mbar_wait(&bar_full[st], (w_consumed >> 1) & 1u);   // W chunk landed
        f_tc_after();
        const uint32_t w_hi = sW_addr + st * w_chunk, w_lo = w_hi + (uint32_t)a.n_pad * W_CHUNK_K * 4;
#pragma unroll
        for (int ks = 0; ks < W_CHUNK_K / 8; ++ks) {
          const uint32_t a_off = (uint32_t)(c * (W_CHUNK_K / 4) + ks * 2) * A_LBO;
          const uint64_t da_hi = f_smem_desc(sA_addr + a_off, A_LBO, a_sbo);
          const uint64_t da_lo = f_smem_desc(sA_addr + a_part + a_off, A_LBO, a_sbo);
          const uint64_t db_hi = f_smem_desc(w_hi + ks * 256, 128, 1024);
          const uint64_t db_lo = f_smem_desc(w_lo + ks * 256, 128, 1024);
          f_umma_tf32(tmem_base, da_hi, db_hi, idesc, (c == 0 && ks == 0) ? 0u : 1u);
          f_umma_tf32(tmem_base, da_lo, db_hi, idesc, 1u);
          f_umma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
        }
        f_umma_commit(&bar_free[st]);
        if (c == n_chunks - 1) f_umma_commit(&bar_acc);
        // refill the stage of the PREVIOUS chunk (its MMAs were issued one iteration ago, so this wait is short)
        if (w_consumed >= 1 && w_issued < w_total) {
          const uint32_t prev = w_consumed - 1;
          f_mbar_wait(&bar_free[prev & 1u], (prev >> 1) & 1u);
          issue_w();
        }
      }
    }
    // ------------------------------------------------ epilogue: TMEM -> smem transpose -> coalesced rows
    f_mbar_wait(&bar_acc, tile_ctr & 1u);
    f_tc_after();
    {
      const int row = (warp & 3) * 32 + lane;
      const int c0 = (warp >> 2) * 32;
      if (c0 < a.n_pad) {
        uint32_t r[32];
        f_tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, r);
        const float lower = a.act_fn == XPGNN_ACT_RELU ? 0.0f : -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float4 v;
          v.x = __uint_as_float(r[c]) + s_bias[c0 + c];
          v.y = __uint_as_float(r[c + 1]) + s_bias[c0 + c + 1];
          v.z = __uint_as_float(r[c + 2]) + s_bias[c0 + c + 2];
          v.w = __uint_as_float(r[c + 3]) + s_bias[c0 + c + 3];
          if (a.act_fn == XPGNN_ACT_SIGMOID) {
            v.x = apply_act(v.x, XPGNN_ACT_SIGMOID); v.y = apply_act(v.y, XPGNN_ACT_SIGMOID);
            v.z = apply_act(v.z, XPGNN_ACT_SIGMOID); v.w = apply_act(v.w, XPGNN_ACT_SIGMOID);
          } else {
            v.x = fmaxf(v.x, lower); v.y = fmaxf(v.y, lower); v.z = fmaxf(v.z, lower); v.w = fmaxf(v.w, lower);
          }
          *reinterpret_cast<float4*>(sT + row * 132 + c0 + c) = v;  // 132-float pitch: conflict-free float4 stores
        }
      }
    }
    f_tc_before();
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      const int p = warp * 8 + i;
      const int64_t oo = s_out_off[p];
      if (oo >= 0 && lane * 4 < a.n_out)
        *reinterpret_cast<float4*>(a.out + oo + lane * 4) = *reinterpret_cast<const float4*>(sT + p * 132 + lane * 4);
    }
    __syncthreads();  // the A tile (aliased by sT), s_out_off and the accumulators are reused by the next tile
  }
  __syncthreads();
  if (warp == 0) f_tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------ host
bool fused_eligible(int K, int n_out, int ld_in, int ld_out, int64_t in_s_stride, int64_t out_s_stride, const void* in, const void* out) {
  if (K % W_CHUNK_K != 0 || K < 32 || K > 128) return false;
  if (n_out < 8 || n_out > 128 || n_out % 4 != 0) return false;
  if (ld_in % 4 || ld_out % 4 || in_s_stride % 4 || out_s_stride % 4) return false;
  if (((uintptr_t)in | (uintptr_t)out) & 15) return false;
  return true;
}

int64_t fused_w_image_bytes(int K, int n_out) {
  const int n_pad = (n_out + 15) / 16 * 16;
  return (int64_t)(K / W_CHUNK_K) * fused_w_chunk_bytes(n_pad);
}

int fused_build_w_image(const float* w, int n_out, int K, float* img, cudaStream_t st) {
  const int n_pad = (n_out + 15) / 16 * 16;
  XP_LAUNCH(fused_w_image_kernel, 64, 256, 0, st, w, n_out, n_pad, K, img);
  return 0;
}

int launch_fused(FusedArgs a, cudaStream_t st) {
  a.n_pad = (a.n_out + 15) / 16 * 16;
  int sb = 8;  // the 8 pairs of a warp must share a node row
  while (sb < a.n_bits && sb < 32) sb <<= 1;
  if (a.SB < 8 || a.SB > 32 || (a.SB & (a.SB - 1))) a.SB = sb;
  const int k16 = a.K / 4;
  const size_t a_bytes = 2ull * 16 * k16 * A_LBO;
  const size_t t_bytes = (size_t)FM * 132 * 4;
  a.w_off = (uint32_t)((std::max(a_bytes, t_bytes) + 127) / 128 * 128);
  const size_t smem = a.w_off + (size_t)W_STAGES * fused_w_chunk_bytes(a.n_pad);
  XP_REQUIRE(smem <= 227 * 1024, "fused tile does not fit shared memory");
  const int RB = FM / a.SB;
  const int64_t tiles = (int64_t)ceil_div(a.n_rows, RB) * ceil_div(a.n_bits, a.SB);
  if (tiles <= 0) return 0;
  const int grid = (int)std::min<int64_t>(tiles, kNumSMs);
  XP_CHECK(cudaFuncSetAttribute(spmm_dense_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  XP_LAUNCH(spmm_dense_fused_kernel, grid, F_THREADS, smem, st, a);
  return 0;
}

}  // namespace xpgnn
