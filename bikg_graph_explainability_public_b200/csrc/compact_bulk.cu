// Masked SpMM of layers >= 1 (aggregate-first: unweighted sums of pre-scaled 128-byte row pieces, out = scale * sum),
// warp-specialised: the gathers are ASYNCHRONOUS COPIES into a shared-memory ring (TMA bulk copies or cp.async), so the data
// in flight is bounded by shared memory instead of registers and the load issue is decoupled from the arithmetic.
//
// STATUS: opt-in (option seg = 2xx / 3xx / 4xx), bit-identical to cspmm_seg_kernel and covered by the parity tests, but
// SLOWER on B200 -- 20.5 ms per C3 tile with cp.async, 35.6 ms with bulk copies, against 11.1 ms for the register-queue
// kernel.  The role cycle counters (make EXPERIMENTS=1, XPGNN_BG_DBG=1) and tools/tma_gather4_probe.cu say why: for 128-byte
// rows the asynchronous copy engines are slower than LDG.128 -- LDGSTS moves 16 bytes per cycle per SM (one 16-byte lane
// request per cycle), a per-lane cp.async.bulk (UBLKCP takes its operands from uniform registers, so a warp issues its 32
// copies one by one) 9 - 15, and the TMA gather4 form (UTMALDG.2D.GATHER4, 4 rows per instruction) 29.7 = 8.4 TB/s per GPU,
// TMA-unit bound (same rate from a 16 MB table) -- against > 40 bytes per cycle per SM for plain LDG.128
// (tools/gather_probe.cu, 11.3 TB/s).  profiles/r02_summary.md section 1b has the numbers.  Kept as the measured answer
// to "stage the gathered feature rows through TMA": on this part the register path is the fast one.
//
//   * one SCHEDULER warp takes 128-row items (slot, chunk, row tile) IN ORDER from the global work counter (the whole GPU
//     stays inside one (slot, chunk) pass, whose gathered operand fits L2), loads the item's row ids / list offsets -- two
//     items ahead of their use -- and publishes a descriptor in shared memory: per row the inclusive end of its part of
//     the gather STREAM (for GCN the row itself first -- the operands are pre-scaled by deg^-1/2, so the unit self loop is
//     one more unweighted term -- then its active sources in list order), the row id and its first list entry;
//   * kBgProd PRODUCER warps walk the stream 32 positions (= one ring stage) at a time, across item boundaries, with the
//     source ids of their next three stages already loading: lane = position, its row by binary search over the row ends,
//     its source id from the compacted list; then either ONE bulk copy per lane (mode 0: cp.async.bulk = the TMA unit,
//     SASS UBLKCP, completion counted in bytes on the stage's `full` mbarrier) or 8 x cp.async (modes 1 / 2: LDGSTS, 8 lanes
//     x 16 bytes per row piece, completion by cp.async.mbarrier.arrive.noinc);
//   * kBgCons CONSUMER warps only read shared memory: a group of 8 lanes (float4 each) owns a row, adds its pieces in stream
//     order (self first: the same order as the other SpMM kernels), scales and stores it.  Rounds of 4 rows are dealt
//     round-robin to the warps; control flow is warp-uniform (groups waiting or looping on their own serialised the warp);
//     a warp releases a stage (one arrival per warp on its `empty` mbarrier) after having seen it full -- that makes an
//     early arrival for a later use of the slot impossible -- and consumes rounds longer than NS / 4 stages window by
//     window, so a row may be longer than the whole ring.
//
// Ring: NS stages x 32 slots x 128 bytes (NS = 32: 128 KB per SM).  Rows with more than long_cnt active in-edges are not
// part of the stream (cspmm_long_kernel sums them, as for the other variants).
#include <algorithm>
#include <cstdlib>

#include "compact_internal.cuh"

namespace xpgnn {

namespace {

constexpr int kBgProd = 8;   // producer warps
constexpr int kBgCons = 8;   // consumer warps (16 measured slower: 26.0 vs 20.5 ms)
constexpr int kBgDesc = 4;   // item descriptors in flight
constexpr int kBgRows = 128; // rows per item
constexpr int kBgThreads = 32 * (1 + kBgProd + kBgCons);

struct BgDesc {
  int aend[kBgRows];     // stream positions of rows 0 .. r (inclusive prefix), relative to the item
  int v[kBgRows];        // node id | -1: no row here / hub row
  uint32_t e[kBgRows];   // first list entry of the row
  int t, c, len, n_rows; // n_rows < 0: no more items
  uint32_t stage0;       // global stage index (per CTA) of the item's first stage
  int pad[3];
};

// cycle counters of CTA 0 (scheduler, producer 0, consumer 0): only in the experiment build (make EXPERIMENTS=1, XPGNN_BG_DBG=1)
#ifdef XPGNN_EXPERIMENTS
__device__ unsigned long long g_bg_dbg[16];
#define BG_TIMED(acc, stmt)                                  \
  do {                                                       \
    const long long _t0 = clock64();                         \
    stmt;                                                    \
    acc += (unsigned long long)(clock64() - _t0);            \
  } while (0)
#else
#define BG_TIMED(acc, stmt) \
  do {                      \
    stmt;                   \
  } while (0)
#endif

template <int NS>
struct BgSmem {
  alignas(128) char ring[NS][32][128];
  BgDesc desc[kBgDesc];
  uint64_t full[NS], empty[NS], dfull[kBgDesc], dempty[kBgDesc];
  int start[33];
  int nact[32];
};

__device__ __forceinline__ uint32_t bg_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bg_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bg_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bg_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bg_u32(bar)) : "memory");
}
__device__ __forceinline__ void bg_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bg_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool bg_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bg_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
template <int SLEEP_NS>
__device__ __forceinline__ void bg_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !bg_test(bar, parity); ++spin) {
    if (SLEEP_NS) __nanosleep(SLEEP_NS);
    if (spin > (1u << 22)) __trap();  // never hang the GPU: a lost arrival becomes an error
  }
}
// one bulk copy (TMA unit) global -> shared; both ends 16-byte aligned; completion in bytes on `bar`
__device__ __forceinline__ void bg_bulk_copy(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gmem_src),
               "r"(bytes), "r"(bg_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bg_add(float4& acc, const float4& x) {  // two packed adds
  asm("{\n\t.reg .b64 a0, a1, x0, x1, one;\n\t"
      "mov.b64 a0, {%0, %1};\n\tmov.b64 a1, {%2, %3};\n\tmov.b64 x0, {%4, %5};\n\tmov.b64 x1, {%6, %7};\n\t"
      "mov.b64 one, {0f3F800000, 0f3F800000};\n\t"
      "fma.rn.f32x2 a0, x0, one, a0;\n\tfma.rn.f32x2 a1, x1, one, a1;\n\t"
      "mov.b64 {%0, %1}, a0;\n\tmov.b64 {%2, %3}, a1;\n\t}"
      : "+f"(acc.x), "+f"(acc.y), "+f"(acc.z), "+f"(acc.w)
      : "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w));
}

// MODE 0: bulk copies (UBLKCP) | 1: cp.async.cg (LDGSTS, 8 lanes x 16 bytes per row piece) | 2: cp.async.ca
template <int NS, int MODE>  // NS: ring stages, a power of two
__global__ void __launch_bounds__(kBgThreads, 1) cspmm_bulk_kernel(const CspmmArgs a) {
  extern __shared__ __align__(128) unsigned char bg_raw[];
  BgSmem<NS>& S = *reinterpret_cast<BgSmem<NS>*>(bg_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x <= a.nb) S.start[threadIdx.x] = a.slot_tile_start[threadIdx.x];
  if (threadIdx.x < a.nb) S.nact[threadIdx.x] = a.slot_info[threadIdx.x].x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { bg_init(&S.full[i], MODE == 0 ? 1 : 32); bg_init(&S.empty[i], kBgCons); }
    for (int i = 0; i < kBgDesc; ++i) { bg_init(&S.dfull[i], 1); bg_init(&S.dempty[i], kBgProd + kBgCons); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool gcn = a.kind == XPGNN_CONV_GCN;
  const int gself = gcn ? 1 : 0;

  if (warp == 0) {
    // ================= scheduler: items in global order -> descriptors =================
    // Software pipeline: while item n is turned into a descriptor, the metadata loads of n + 1 and n + 2 and the counter
    // fetch of n + 3 are in flight (under load a DRAM round trip is several microseconds; an item lasts about two).
    const int total = S.start[a.nb] * a.n_chunks;
    struct Meta {
      int item, t, c, row0;
      int v[4];
      uint32_t e[4], f[4];
    };
    int t_ld = 0;
    auto grab_raw = [&]() {  // lane 0 holds the value; broadcast when it is needed
      int it = 0;
      if (lane == 0) it = atomicAdd(a.counter, 1);
      return it;
    };
    auto load_meta = [&](int item, Meta& m) {
      m.item = item;
      m.t = 0; m.c = 0; m.row0 = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) { m.v[k] = -1; m.e[k] = 0; m.f[k] = 0; }
      if (item >= total) return;
      while (item >= S.start[t_ld + 1] * a.n_chunks) ++t_ld;
      const int ntb = S.start[t_ld + 1] - S.start[t_ld];
      const int rem = item - S.start[t_ld] * a.n_chunks;
      m.t = t_ld;
      m.c = rem / ntb;
      m.row0 = (rem - m.c * ntb) * kBgRows;
      const int n_act = S.nact[t_ld];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = m.row0 + k * 32 + lane;
        if (i < n_act) {
          m.v[k] = __ldcs(a.act_list + (int64_t)t_ld * a.N + i);
          const uint32_t* rp = a.rowptr_c + (int64_t)t_ld * (a.N + 1) + i;
          m.e[k] = __ldcs(rp);
          m.f[k] = __ldcs(rp + 1);
        }
      }
    };
    Meta A, B;
    unsigned long long t_w0 = 0;
    const long long t_begin = clock64();
    load_meta(__shfl_sync(0xffffffffu, grab_raw(), 0), A);
    load_meta(__shfl_sync(0xffffffffu, grab_raw(), 0), B);
    int raw = grab_raw();
    uint32_t stage_next = 0;
    for (uint32_t n = 0;; ++n) {
      const int d = n % kBgDesc;
      BgDesc& D = S.desc[d];
      if (n >= kBgDesc) BG_TIMED(t_w0, bg_wait<64>(&S.dempty[d], ((n / kBgDesc) - 1) & 1));
      if (A.item >= total) {
        if (lane == 0) D.n_rows = -1;
        __syncwarp();
        if (lane == 0) bg_arrive(&S.dfull[d]);
#ifdef XPGNN_EXPERIMENTS
        if (blockIdx.x == 0 && lane == 0) { g_bg_dbg[0] = (unsigned long long)(clock64() - t_begin); g_bg_dbg[1] = t_w0; g_bg_dbg[2] = n; }
#endif
        break;
      }
      int carry = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t cnt = A.f[k] - A.e[k];
        const bool mine = A.v[k] >= 0 && !(a.long_cnt > 0 && cnt > (uint32_t)a.long_cnt);
        int x = mine ? (int)cnt + gself : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += y;
        }
        D.aend[k * 32 + lane] = carry + x;
        D.v[k * 32 + lane] = mine ? A.v[k] : -1;
        D.e[k * 32 + lane] = A.e[k];
        carry += __shfl_sync(0xffffffffu, x, 31);
      }
      if (lane == 0) {
        D.t = A.t; D.c = A.c; D.len = carry; D.n_rows = 1; D.stage0 = stage_next;
      }
      stage_next += (uint32_t)((carry + 31) >> 5);
      __syncwarp();
      if (lane == 0) bg_arrive(&S.dfull[d]);
      A = B;
      load_meta(__shfl_sync(0xffffffffu, raw, 0), B);
      raw = grab_raw();
    }
  } else if (warp <= kBgProd) {
    // ================= producers: stream positions -> copies into the ring =================
    // Producer pw owns the stages whose GLOBAL index is pw mod kBgProd.  It walks them as one sequence across item
    // boundaries with the source ids of the next kBgAhead stages already loading (a FETCH cursor runs ahead of the
    // ISSUE cursor), so neither a stage nor the first stage of an item waits for a list load.
    constexpr int kAhead = 3;
    const int pw = warp - 1;
    const int sub = lane & 7, grp = lane >> 3;
    // ---- fetch cursor ----
    uint32_t fn = 0;        // item the cursor is in
    int fk = 0, f_nst = -1; // next stage of the item for this warp; f_nst < 0: item not opened yet
    bool f_end = false;
#ifdef XPGNN_EXPERIMENTS
    const bool bg_fake = a.l2_gather == 77;  // set through the l2_gather knob in the experiment build
#endif
    unsigned long long t_w0 = 0, t_w1 = 0, t_f = 0, n_st = 0;
    const long long t_begin = clock64();
    // 0: produced the next stage of this warp | 1: the next item's descriptor is not there yet (only when !may_block) | 2: end
    auto f_next = [&](int& id, uint32_t& g, uint32_t& item_no, bool may_block) -> int {
      for (;;) {
        if (f_end) return 2;
        const BgDesc& D = S.desc[fn % kBgDesc];
        if (f_nst < 0) {  // open item fn
          if (may_block) BG_TIMED(t_w0, bg_wait<32>(&S.dfull[fn % kBgDesc], (fn / kBgDesc) & 1));
          else if (!bg_test(&S.dfull[fn % kBgDesc], (fn / kBgDesc) & 1)) return 1;
          if (D.n_rows < 0) { f_end = true; return 2; }
          f_nst = (D.len + 31) >> 5;
          fk = (int)(((uint32_t)pw + (uint32_t)kBgProd - (D.stage0 % (uint32_t)kBgProd)) % (uint32_t)kBgProd);
        }
        if (fk < f_nst) {
          const int len = D.len, pos = fk * 32 + lane;
          id = -1;
#ifdef XPGNN_EXPERIMENTS
          if (bg_fake && pos < len) {  // ceiling experiment: pseudo-random source ids without the search / the list load (WRONG sums)
            id = (int)(((uint32_t)pos * 2654435761u + D.stage0 * 40503u + (uint32_t)fk * 7919u) % (uint32_t)a.N);
          } else
#endif
          if (pos < len) {
            int r = 0;  // rows whose stream ends at or before pos
#pragma unroll
            for (int step = 64; step >= 1; step >>= 1)
              if (D.aend[r + step - 1] <= pos) r += step;
            const int ar = r ? D.aend[r - 1] : 0;
            if (gcn && pos == ar) id = D.v[r];
            else id = __ldcs(a.ccol + a.slot_base[D.t] + D.e[r] + (uint32_t)(pos - ar - gself));
          }
          g = D.stage0 + (uint32_t)fk;
          item_no = fn;
          fk += kBgProd;
          return 0;
        }
        ++fn;
        f_nst = -1;
      }
    };
    uint32_t done_n = 0;  // this warp has arrived on dempty of every item below done_n
    auto finish_items = [&](uint32_t upto) {
      while (done_n < upto) {
        __syncwarp();
        if (lane == 0) bg_arrive(&S.dempty[done_n % kBgDesc]);
        ++done_n;
      }
    };
    auto issue = [&](int id, uint32_t g, uint32_t item_no) {
      finish_items(item_no);  // every stage of the earlier items has been issued
      const BgDesc& D = S.desc[item_no % kBgDesc];
      const uint32_t slot = g % NS, use = g / NS;
      const int k = (int)(g - D.stage0);
      const char* in_c = reinterpret_cast<const char*>(a.in + (int64_t)D.t * a.in_s_stride + (int64_t)D.c * a.in_chunk_stride);
      if (use > 0) BG_TIMED(t_w1, bg_wait<64>(&S.empty[slot], (use - 1) & 1));
      ++n_st;
      if (MODE == 0) {
        if (lane == 0) bg_arrive_expect_tx(&S.full[slot], (uint32_t)min(32, D.len - k * 32) * 128u);
        __syncwarp();
        if (id >= 0) {
          const char* src;
          asm("mad.wide.u32 %0, %1, 128, %2;" : "=l"(src) : "r"((uint32_t)id), "l"(in_c));
          bg_bulk_copy(bg_u32(&S.ring[slot][lane][0]), src, 128u, &S.full[slot]);
        }
      } else {
        const uint32_t dst0 = bg_u32(&S.ring[slot][0][0]) + (uint32_t)sub * 16u;
        const char* in_l = in_c + sub * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int pj = j * 4 + grp;  // slot of the stage: 8 lanes x 16 bytes each
          const int idj = __shfl_sync(0xffffffffu, id, pj);
          if (idj >= 0) {
            const char* src;
            asm("mad.wide.u32 %0, %1, 128, %2;" : "=l"(src) : "r"((uint32_t)idj), "l"(in_l));
            if (MODE == 1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)pj * 128u), "l"(src) : "memory");
            else asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)pj * 128u), "l"(src) : "memory");
          }
        }
        // the stage is full once every producer lane's copies have landed (32 arrivals, none added by the instruction)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bg_u32(&S.full[slot])) : "memory");
      }
    };
    // The queue is a FIFO over kAhead register slots used round-robin (the loop is unrolled, so slot indices are compile-time
    // constants and the ids stay in registers -- their loads may still be in flight).  Entries are issued strictly in order:
    // once a refill finds the next descriptor missing, nothing is appended until the queue has drained; only an EMPTY queue
    // blocks on a descriptor, after this warp has let go of every earlier item (so the scheduler can always publish it).
    int q_id[kAhead];
    uint32_t q_g[kAhead], q_n[kAhead];
    bool q_ok[kAhead];
#pragma unroll
    for (int i = 0; i < kAhead; ++i) { q_ok[i] = false; q_id[i] = -1; q_g[i] = 0; q_n[i] = 0; }
    bool stalled = true, ended = false, live = true;  // stalled: no appends until the head finds the queue empty
    while (live) {
#pragma unroll
      for (int i = 0; i < kAhead; ++i) {
        if (live && !q_ok[i]) {  // in-order FIFO: an empty head means an empty queue
          if (ended) {
            live = false;
          } else {
            finish_items(fn);  // all of this warp's stages of the items before the cursor have been issued
            const int rc = f_next(q_id[i], q_g[i], q_n[i], true);
            if (rc == 2) { ended = true; live = false; }
            else {
              q_ok[i] = true;
              stalled = false;
#pragma unroll
              for (int j = 1; j < kAhead; ++j) {  // top the queue up behind the new head
                const int sj = (i + j) % kAhead;
                if (!stalled && !ended) {
                  const int r2 = f_next(q_id[sj], q_g[sj], q_n[sj], false);
                  q_ok[sj] = r2 == 0;
                  if (r2 == 1) stalled = true;
                  if (r2 == 2) ended = true;
                }
              }
            }
          }
        }
        if (live) {
          issue(q_id[i], q_g[i], q_n[i]);
          q_ok[i] = false;
          if (!stalled && !ended) {  // slot i becomes the tail
            int r2;
            BG_TIMED(t_f, r2 = f_next(q_id[i], q_g[i], q_n[i], false));
            q_ok[i] = r2 == 0;
            if (r2 == 1) stalled = true;
            if (r2 == 2) ended = true;
          }
        }
      }
    }
    finish_items(fn);  // fn = the item that carried the end marker
#ifdef XPGNN_EXPERIMENTS
    if (blockIdx.x == 0 && pw == 0 && lane == 0) {
      g_bg_dbg[4] = (unsigned long long)(clock64() - t_begin); g_bg_dbg[5] = t_w0; g_bg_dbg[6] = t_w1; g_bg_dbg[7] = t_f; g_bg_dbg[8] = n_st;
    }
#endif
  } else {
    // ================= consumers: a group of 8 lanes sums a row from the ring =================
    // Round k of an item = rows 4k .. 4k + 3 (one per group), dealt round-robin to the consumer warps, so all warps work
    // next to the fill front.  Releases are per WARP (kBgCons arrivals per stage): before a round the warp releases every
    // stage below the round's first one -- lanes take different stages; each stage is seen full first, which makes an early
    // arrival for a later use of the slot impossible -- and a round longer than kWin stages is consumed window by window, so
    // no row length can hold more than kWin + 1 stages unreleased (a row may be longer than the whole ring).
    constexpr int kWin = NS / 4;
    constexpr int kPar = NS < 32 ? NS : 32;
    const int cw = warp - 1 - kBgProd, sub = lane & 7, grp = lane >> 3;
    uint32_t rel = 0;   // every stage below `rel` has been released by this warp (warp-uniform)
    unsigned long long t_w0 = 0, t_w1 = 0, t_w2 = 0, n_rd = 0;
    const long long t_begin = clock64();
    auto release_to = [&](uint32_t target) {
      while (rel < target) {
        const uint32_t g = rel + (uint32_t)lane;
        if (lane < kPar && g < target) {  // no lane may wait on a use of a slot that is two uses ahead of the barrier's phase
          bg_wait<20>(&S.full[g % NS], (g / NS) & 1);
          bg_arrive(&S.empty[g % NS]);
        }
        rel = min(target, rel + (uint32_t)kPar);
      }
    };
    for (uint32_t n = 0;; ++n) {
      const int d = n % kBgDesc;
      const BgDesc& D = S.desc[d];
      BG_TIMED(t_w0, bg_wait<32>(&S.dfull[d], (n / kBgDesc) & 1));
      if (D.n_rows < 0) break;
      const uint32_t stage0 = D.stage0;
      const int nst = (D.len + 31) >> 5;
      float* out_c = a.out + (int64_t)D.t * a.out_s_stride + (int64_t)D.c * a.out_chunk_stride + sub * 4;
      for (int k = cw; k < kBgRows / 4; k += kBgCons) {
        const int P0 = k ? D.aend[4 * k - 1] : 0, P1 = D.aend[4 * k + 3];
        if (P1 == P0 && D.v[4 * k] < 0 && D.v[4 * k + 1] < 0 && D.v[4 * k + 2] < 0 && D.v[4 * k + 3] < 0) continue;  // no rows here
        const int r = 4 * k + grp;
        const int v = D.v[r];
        const int a0 = r ? D.aend[r - 1] : 0, a1 = D.aend[r];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int sa = P0 >> 5, sb = (P1 + 31) >> 5;  // the round's positions sit in stages [sa, sb) of the item
        BG_TIMED(t_w1, release_to(stage0 + (uint32_t)sa));
        ++n_rd;
        // Warp-uniform control flow (groups that wait or loop on their own serialise the warp: 2900 cycles per round
        // measured): every lane waits for every stage of the window, then all groups run the same number of predicated steps.
        for (int ws = sa; ws < sb; ws += kWin) {
          const int we = min(ws + kWin, sb);
          for (int gg = ws; gg < we; ++gg) {
            const uint32_t g = stage0 + (uint32_t)gg;
            BG_TIMED(t_w2, bg_wait<20>(&S.full[g % NS], (g / NS) & 1));
          }
          const int lo = max(a0, ws << 5), hi = min(a1, we << 5);
          const int n_it = __reduce_max_sync(0xffffffffu, max(hi - lo, 0));
          const uint32_t pos0 = stage0 * 32u + (uint32_t)lo;  // ring slot of position q: (stage0 * 32 + q) mod (NS * 32)
          for (int i = 0; i < n_it; i += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (lo + i + j < hi) {
                const uint32_t slot = (pos0 + (uint32_t)(i + j)) & (uint32_t)(NS * 32 - 1);
                bg_add(acc, *reinterpret_cast<const float4*>(&S.ring[0][0][0] + (size_t)slot * 128 + sub * 16));
              }
            }
          }
          __syncwarp();
          if (we < sb) release_to(stage0 + (uint32_t)we);
        }
        if (v >= 0) {
          const uint32_t cnt = (uint32_t)(a1 - a0 - gself);
          float sc = 0.f;  // SAGE row without an active in-edge: empty mean
          if (gcn) sc = gcn_dinv(cnt);
          else if (cnt) sc = 1.0f / (float)cnt;
          __stcs(reinterpret_cast<float4*>(out_c + (int64_t)v * 32), make_float4(acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc));
        }
        __syncwarp();
      }
      BG_TIMED(t_w1, release_to(stage0 + (uint32_t)nst));
      __syncwarp();
      if (lane == 0) bg_arrive(&S.dempty[d]);
    }
#ifdef XPGNN_EXPERIMENTS
    if (blockIdx.x == 0 && cw == 0 && lane == 0) {
      g_bg_dbg[10] = (unsigned long long)(clock64() - t_begin); g_bg_dbg[11] = t_w0; g_bg_dbg[12] = t_w1; g_bg_dbg[13] = t_w2; g_bg_dbg[14] = n_rd;
    }
#endif
  }
}

}  // namespace

int launch_cspmm_bulk(const CspmmArgs& a, int variant, cudaStream_t st) {
  // variant = 100 * mode + ring stages: mode 0 bulk copies | 1 cp.async.cg | 2 cp.async.ca; stages 16 | 32
  const int mode = variant / 100, ns = variant % 100;
  void (*k)(const CspmmArgs);
  int smem;
#define BG_PICK(NS_)                                                                                          \
  do {                                                                                                        \
    k = mode == 0 ? cspmm_bulk_kernel<NS_, 0> : (mode == 1 ? cspmm_bulk_kernel<NS_, 1> : cspmm_bulk_kernel<NS_, 2>); \
    smem = (int)sizeof(BgSmem<NS_>);                                                                          \
  } while (0)
  if (ns >= 32) BG_PICK(32);
  else BG_PICK(16);
#undef BG_PICK
  XP_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  XP_LAUNCH(k, kNumSMs, kBgThreads, smem, st, a);
#ifdef XPGNN_EXPERIMENTS
  static const bool dbg_on = getenv("XPGNN_BG_DBG") != nullptr;
  static int n_launch = 0;
  if (dbg_on && (n_launch++ % 8) == 7) {
    unsigned long long h[16];
    XP_CHECK(cudaStreamSynchronize(st));
    XP_CHECK(cudaMemcpyFromSymbol(h, g_bg_dbg, sizeof h));
    fprintf(stderr,
            "bg dbg variant %d: sched total %llu wait_dempty %llu items %llu | prod0 total %llu wait_desc %llu wait_empty %llu fetch %llu stages %llu | "
            "cons0 total %llu wait_desc %llu release %llu wait_full %llu rounds %llu\n",
            variant, h[0], h[1], h[2], h[4], h[5], h[6], h[7], h[8], h[10], h[11], h[12], h[13], h[14]);
  }
#endif
  return 0;
}

}  // namespace xpgnn
