// pipeline.cuh — producer/consumer plumbing shared by the blend kernels: mbarrier helpers, the
// shared-memory ring geometry, footprint -> warp-block mask, and the producer warp's batch gather.
#pragma once
#include "common.cuh"

namespace tgr {

constexpr int BL_BATCH = 128;                // list entries per ring stage
constexpr int BL_CHUNKS = BL_BATCH / 32;     // 32-entry chunks (one ballot word each)
constexpr int BL_THREADS = 9 * 32;           // 8 consumer warps + 1 producer warp

// ---- SM-affine work queues ------------------------------------------------------------------------------
// Tiles are sorted heaviest-first (tile_order_kernel).  All non-empty tiles of an avatar view fit in the first
// wave of CTAs, so WHICH SM a tile lands on decides the balance, and the hardware's CTA->SM placement is not
// a simple modulo (measured on B200: tools/cuda/smid_map.cu).  Instead of trusting blockIdx, every CTA reads
// %smid and pops the next rank of that SM's own queue; queue q holds the ranks dealt to it boustrophedon-wise
// (q, 2n-1-q, 2n+q, 4n-1-q, ...) so per-SM sums of a sorted list are even.  A CTA whose queue is drained
// steals from the others (warp-parallel scan of the counters), which also absorbs dynamic imbalance.
constexpr uint32_t NO_TILE = 0xffffffffu;

__device__ __forceinline__ uint32_t queue_rank(uint32_t q, uint32_t k, uint32_t nq) {
  return k * nq + ((k & 1u) ? (nq - 1u - q) : q);
}

// Called by warp 0 (all 32 lanes).  counters[nq] must be zero at kernel start.  Returns a rank < T or NO_TILE.
__device__ __forceinline__ uint32_t fetch_tile_rank(uint32_t* __restrict__ counters, uint32_t T, uint32_t nq, int lane) {
  uint32_t smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const uint32_t q = smid % nq;
  uint32_t r = NO_TILE;
  if (lane == 0) {
    const uint32_t k = atomicAdd(&counters[q], 1u);
    r = queue_rank(q, k, nq);
    if (r >= T) r = NO_TILE;
  }
  r = __shfl_sync(0xffffffffu, r, 0);
  if (r != NO_TILE) return r;
  // own queue drained: steal.  Each lane inspects queues q+1+lane, q+1+lane+32, ...
  // (#CTAs == #ranks, so a CTA may only give up once every queue is observed drained)
  while (true) {
    uint32_t found = NO_TILE;
    for (uint32_t i = lane; i < nq && found == NO_TILE; i += 32) {
      const uint32_t qq = (q + 1u + i) % nq;
      const uint32_t k = *(volatile uint32_t*)&counters[qq];
      if (queue_rank(qq, k, nq) < T) found = qq;
    }
    const uint32_t have = __ballot_sync(0xffffffffu, found != NO_TILE);
    if (!have) return NO_TILE;  // every queue is drained
    const int src = __ffs(have) - 1;
    const uint32_t qq = __shfl_sync(0xffffffffu, found, src);
    if (lane == 0) {
      const uint32_t k = atomicAdd(&counters[qq], 1u);
      r = queue_rank(qq, k, nq);
      if (r >= T) r = NO_TILE;
    }
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r != NO_TILE) return r;
  }
}

// ---- mbarrier (shared::cta) ------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// Blocks until the phase with the given parity completes.  try_wait parks the warp only for a few dozen cycles
// (whatever the suspend hint says), so a waiting warp re-issues it: ncu attributed 26 % of all instructions the
// forward issued to these polling loops once the fused multi-view launches had made the kernel issue-bound.
// Failed tries therefore back off with nanosleep (doubling, capped): a warp that waits long polls rarely, and the
// ring's slack of several batches hides the coarser wake-up.
__device__ __forceinline__ bool mbar_try(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(a), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  if (mbar_try(a, parity)) return;
  uint32_t ns = 64;
  while (true) {
    __nanosleep(ns);
    if (mbar_try(a, parity)) return;
    if (ns < 512) ns *= 2;
  }
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- cp.async (LDGSTS) -----------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- producer warp ------------------------------------------------------------------------------------
// Batch entry e maps to list position e (forward) or total-1-e (reverse, for the back-to-front replay).
// The producer is software-pipelined: ids of batch b+2 are prefetched into registers and the 16-byte
// cp.async gathers of batch b+1 are in flight while the footprints of batch b are classified.
__device__ __forceinline__ void prod_load_ids(const uint32_t* __restrict__ list, int total, int first, bool reverse,
                                              int lane, uint32_t (&ids)[BL_CHUNKS]) {
#pragma unroll
  for (int c = 0; c < BL_CHUNKS; ++c) {
    const int e = first + c * 32 + lane;
    ids[c] = 0xffffffffu;
    if (e < total) ids[c] = __ldg(list + (reverse ? (total - 1 - e) : e));
  }
}

// Records go global -> shared with cp.async (no register staging).  The backward also needs the Gaussian id of
// every entry next to its record; it travels as a 4-byte cp.async from the list itself so that EVERYTHING a
// consumer reads is covered by the asynchronous arrive on full[] (producer_loop).
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void prod_issue(const uint32_t (&ids)[BL_CHUNKS], const uint32_t* __restrict__ list, int total,
                                           int first, bool reverse, const float4* __restrict__ xy_ext,
                                           const float4* __restrict__ conic_opacity,
                                           const float4* __restrict__ rgb_depth, float4* s_xy, float4* s_co,
                                           float4* s_cd, uint32_t* s_id, int lane) {
#pragma unroll
  for (int c = 0; c < BL_CHUNKS; ++c) {
    if (ids[c] != 0xffffffffu) {
      const int j = c * 32 + lane;
      cp_async16(&s_xy[j], &xy_ext[ids[c]]);
      cp_async16(&s_co[j], &conic_opacity[ids[c]]);
      cp_async16(&s_cd[j], &rgb_depth[ids[c]]);
      if (s_id) {
        const int e = first + j;
        cp_async4(&s_id[j], list + (reverse ? (total - 1 - e) : e));
      }
    }
  }
}

// ---- consumer-side classification -------------------------------------------------------------------
// Every consumer warp tests the landed footprints of a batch against ITS OWN 8x4 pixel block and compacts the
// hits into an ordered list of batch-local indices (bytes), padded to a multiple of CAND_GROUP with PAD_ENTRY.
// PAD_ENTRY indexes a dummy record with opacity 0 (it fails the alpha >= 1/255 test), so the evaluation loop
// walks full groups without validity bookkeeping: one 32-bit shared load yields four candidates.
// History (ncu-driven): v1 had the producer warp classify all eight blocks (8 ballots per 32 entries) — it became
// the serial bottleneck of the ring, a third of all issued instructions were consumers polling full[]; v2 moves
// the test to the eight consumers (one ballot per 32 entries each, in parallel) and leaves the producer a pure
// data mover (ids -> cp.async gathers), like a TMA producer.
// The block test is conservative: pixel centres are integers (forward.cu:282), pixel p can reach alpha >= 1/255
// only if x-hx <= p <= x+hx, so a block [x0,x1] is skipped when x-hx > x1 or x+hx < x0 (same in y).
constexpr int CAND_GROUP = 4;
constexpr int PAD_ENTRY = BL_BATCH;                 // dummy record slot; record arrays hold BL_BATCH + 1 entries
constexpr int LIST_BYTES = BL_BATCH + CAND_GROUP;   // per consumer warp

// Sub-blocks.  The splats of an avatar are tiny: ncu showed the blend instructions of the forward running with 5
// of 32 lanes active on average, i.e. a candidate of an 8x4 block touches ~3-5 of its pixels.  A warp therefore
// splits into SUB_GROUPS lane groups, each owning a SUB_W x SUB_H sub-block of the warp's 8x4 block and walking
// ITS OWN candidate list: in one warp instruction up to SUB_GROUPS different Gaussians are evaluated, each only
// on the pixels of a sub-block it can reach.  The warp iterates to the longest of its group lists; shorter lists
// read PAD_ENTRY.
// Measured on the fused 8-view C3 batch (blend_fwd / blend_bwd ms): 2 of 4x4 1.35 / 2.15,
// 4 of 4x2 1.29 / 2.17, 8 of 2x2 1.38 / 2.52 (classification costs one ballot per group and 32 entries).
constexpr int SUB_GROUPS = 4;
constexpr int SUB_W = 4, SUB_H = 2;
constexpr int SUB_GX = 8 / SUB_W;                 // sub-blocks across the warp's 8x4 block
constexpr int SUB_LANES = 32 / SUB_GROUPS;
static_assert(SUB_W * SUB_H == SUB_LANES && (8 % SUB_W) == 0 && (4 % SUB_H) == 0, "sub-block geometry");

// pixel (relative to the tile origin) owned by `lane` of consumer warp `warp`; shared by forward and backward
__device__ __forceinline__ void lane_pixel(int warp, int lane, int& x, int& y, int& group) {
  group = lane / SUB_LANES;
  const int t = lane % SUB_LANES;
  x = (warp & 1) * 8 + (group % SUB_GX) * SUB_W + (t % SUB_W);
  y = (warp >> 1) * 4 + (group / SUB_GX) * SUB_H + (t / SUB_W);
}

// first_excluded: entries >= this batch-local index are ignored (ragged last batch; in the backward also the
// entries behind the block's last contributor, see blend_bwd.cu).  first_included: entries below are ignored.
// (bx0, by0) is the pixel origin of the warp's 8x4 block.  Writes SUB_GROUPS lists (lists[g*LIST_BYTES ...]) and
// returns, per lane, the padded length of ITS group's list; `longest` is the warp-wide maximum.
// group_live: bit g set = sub-block g still has a pixel that can take contributions (the forward switches a sub-block
// off once all of its pixels have terminated: a tile on the silhouette keeps a few pixels alive deep into its list).
__device__ __forceinline__ int cons_classify(int first_included, int first_excluded, const float4* s_xy, uint8_t* lists,
                                             float bx0, float by0, int lane, int& longest, uint32_t group_live = 0xfu) {
  uint32_t cnt[SUB_GROUPS];
#pragma unroll
  for (int g = 0; g < SUB_GROUPS; ++g) cnt[g] = 0;
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < BL_CHUNKS; ++c) {
    const int e = c * 32 + lane;
    float x_lo = 1e30f, x_hi = -1e30f, y_lo = 1e30f, y_hi = -1e30f;  // empty interval: hits nothing
    if (e >= first_included && e < first_excluded) {
      const float4 q = s_xy[e];
      if (q.z >= 0.f) { x_lo = q.x - q.z - bx0; x_hi = q.x + q.z - bx0; y_lo = q.y - q.w - by0; y_hi = q.y + q.w - by0; }
    }
#pragma unroll
    for (int g = 0; g < SUB_GROUPS; ++g) {
      const float sx0 = (float)((g % SUB_GX) * SUB_W), sy0 = (float)((g / SUB_GX) * SUB_H);
      const bool hit = ((group_live >> g) & 1u) && (x_lo <= sx0 + (float)(SUB_W - 1)) && (x_hi >= sx0) &&
                       (y_lo <= sy0 + (float)(SUB_H - 1)) && (y_hi >= sy0);
      const uint32_t bal = __ballot_sync(0xffffffffu, hit);
      if (hit) lists[g * LIST_BYTES + cnt[g] + __popc(bal & lt)] = (uint8_t)e;
      cnt[g] += __popc(bal);
    }
  }
  int mine = 0;
  longest = 0;
#pragma unroll
  for (int g = 0; g < SUB_GROUPS; ++g) {
    const uint32_t padded = (cnt[g] + CAND_GROUP - 1) & ~(uint32_t)(CAND_GROUP - 1);
    if (cnt[g] + lane < padded) lists[g * LIST_BYTES + cnt[g] + lane] = (uint8_t)PAD_ENTRY;
    if (lane / SUB_LANES == g) mine = (int)padded;
    longest = max(longest, (int)padded);
  }
  __syncwarp();
  return mine;
}

// ---- pair-centric evaluation: hits -> candidates -> pairs ------------------------------------------------
// The splats of a dense avatar are a few pixels wide and mostly thin (surface-aligned discs seen at an angle): with
// pixels mapped to lanes (above), ncu showed the blend instructions of the backward running with 5.8 of 32 lanes
// active inside the gradient block and ~13 % of the evaluated (pixel, Gaussian) slots passing the alpha test.  The
// pair-centric kernels turn the loop inside out: a consumer warp still owns 32 pixels of the tile, but its lanes no
// longer stand for pixels.  Per batch of staged list entries it
//   1. classifies the entries against its pixels and compacts the HITS — entry index plus, per pixel row, the span of
//      integer pixels inside the ellipse {alpha >= 1/255} (not its bounding box: a thin ellipse at 45 degrees fills
//      a fifth of its box) — into shared memory;
//   2. expands the hits into CANDIDATES, one (entry, pixel) per lane, 32 per round, every lane busy: the spans are
//      laid end to end by a warp scan and a lane finds its hit with a reduce-or / popc over the span starts;
//   3. evaluates alpha per candidate and compacts the survivors into a ring of PAIRS;
//   4. commits 32 pairs at a time to the per-pixel recurrence state (two scalars per pixel, in shared memory); pairs of
//      one round that fall on the same pixel compose their updates in list order by pointer jumping over the peer
//      lanes (match.any -> previous peer, log2 steps of shuffles; blend_bwd.cu), everything else runs at full width.
// Consumer warp w owns tile rows 2w and 2w+1 (16 x 2 pixels), so the 16 row spans of an entry are computed exactly
// once per tile — by the warp that owns the row.  Pixel index inside the warp: (row << 4) | x; lane l owns pixel l for
// the initialisation / write-out of the state.
constexpr int PB_CAP = 64;                                  // pair ring per warp (>= 32 carried + 32 new)

// hit word: entry 0..127 | x0 of row 0 << 7 (4) | w of row 0 << 11 (5, 0..16) | x0 of row 1 << 16 (4) | w of row 1 << 20 (5)
__device__ __forceinline__ int hit_pixels(uint32_t hw) { return (int)((hw >> 11) & 31u) + (int)((hw >> 20) & 31u); }

// Span of integer pixel columns of tile row `v` (relative to the Gaussian's centre: v = row - y) that lie inside the
// ellipse A u^2 + 2 B u v + C v^2 <= 2 tau:  u in kB v -+ sqrt(s (hy^2 - v^2)),  kB = -B / A,  s = det / A^2, and
// hy^2 = 2 tau A / det is the half-height of the ellipse's bounding box as stored by the preprocess kernel (with its
// margins: tau inflated by 1 %, hy by 0.1 % + 0.01 px); 0.02 px are added to the half-width for the fp32 rounding of
// centre and root.  Returns the width (0 = row not reached) and the first column, both clipped to the pixel columns
// [lo, hi] (the live part of the row, see RowWindow).
__device__ __forceinline__ int row_span(float x, float v, float hy, float kB, float s, int lo, int hi, int& x0) {
  const float t = hy * hy - v * v;
  if (!(t >= 0.f)) return 0;
  const float hw = sqrtf(s * t) + 0.02f;
  const float c = fmaf(kB, v, x);
  x0 = max(__float2int_ru(c - hw), lo);
  const int x1 = min(__float2int_rd(c + hw), hi);
  return max(x1 - x0 + 1, 0);
}

// The columns of the warp's two rows that can still take part in the current batch: [lo, hi] per row in absolute pixel
// columns, lo > hi = row switched off.  The forward passes the columns whose pixels have not terminated yet, the
// backward those whose last contributor lies at or before the batch: a tile on the silhouette keeps a few columns
// alive deep into its list, and only those are enumerated.
struct RowWindow {
  int lo0, hi0, lo1, hi1;
};
__device__ __forceinline__ RowWindow row_window(uint32_t live_mask, int tx0) {   // bit (row << 4 | x) set = pixel live
  RowWindow w;
  const uint32_t m0 = live_mask & 0xffffu, m1 = live_mask >> 16;
  w.lo0 = m0 ? tx0 + (__ffs(m0) - 1) : 1;
  w.hi0 = m0 ? tx0 + (31 - __clz(m0)) : 0;
  w.lo1 = m1 ? tx0 + (__ffs(m1) - 1) : 1;
  w.hi1 = m1 ? tx0 + (31 - __clz(m1)) : 0;
  return w;
}

// Entries [first_included, first_excluded) of the landed batch against tile rows ty0, ty0 + 1 (pixel rows) of the tile
// whose first pixel column is tx0, restricted to the live columns `rw`.  Writes the hit words in list order and returns
// their number.
__device__ __forceinline__ int classify_hits(int first_included, int first_excluded, const float4* s_xy, const float4* s_co,
                                             uint32_t* hits, int tx0, int ty0, const RowWindow rw, int lane) {
  int cnt = 0;
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < BL_CHUNKS; ++c) {
    const int e = c * 32 + lane;
    uint32_t word = 0;
    bool hit = false;
    if (e >= first_included && e < first_excluded) {
      const float4 q = s_xy[e];
      const float v0 = (float)ty0 - q.y;                       // row 0 relative to the centre; row 1 is v0 + 1
      if (q.z >= 0.f && v0 + 1.f >= -q.w && v0 <= q.w) {       // opacity >= 1/255 and the box reaches one of the rows
        int x00 = tx0, x01 = tx0, w0, w1;
        if (q.w < 1e29f) {
          const float4 co = s_co[e];
          const float invA = 1.0f / co.x;
          const float kB = -co.y * invA;
          const float s = (co.x * co.z - co.y * co.y) * invA * invA;
          w0 = row_span(q.x, v0, q.w, kB, s, rw.lo0, rw.hi0, x00);
          w1 = row_span(q.x, v0 + 1.f, q.w, kB, s, rw.lo1, rw.hi1, x01);
        } else {                                               // ill-conditioned conic: no culling (preprocess.cu)
          x00 = rw.lo0; w0 = max(rw.hi0 - rw.lo0 + 1, 0);
          x01 = rw.lo1; w1 = max(rw.hi1 - rw.lo1 + 1, 0);
        }
        hit = (w0 + w1) > 0;
        word = (uint32_t)e | ((uint32_t)((x00 - tx0) & 15) << 7) | ((uint32_t)w0 << 11) | ((uint32_t)((x01 - tx0) & 15) << 16) |
               ((uint32_t)w1 << 20);
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, hit);
    if (hit) hits[cnt + __popc(bal & lt)] = word;
    cnt += __popc(bal);
  }
  __syncwarp();
  return cnt;
}

// One expansion round.  Lane i holds hit i of the current group (`hw`, n = 0 beyond the group), `start` = exclusive
// prefix of the hits' pixel counts.  Candidate `base + lane` -> entry index `j` and pixel (row << 4 | x) `pix`.
__device__ __forceinline__ bool expand_candidate(uint32_t hw, int n, int start, int total, int base, int lane, int& j, int& pix) {
  const bool starts_here = n > 0 && start >= base && start < base + 32;
  const uint32_t m = __reduce_or_sync(0xffffffffu, starts_here ? (1u << (start - base)) : 0u);
  const int before = __popc(__ballot_sync(0xffffffffu, n > 0 && start < base));
  const int idx = before + __popc(m & (0xffffffffu >> (31 - lane))) - 1;   // last hit starting at or before this candidate
  const uint32_t w_hit = __shfl_sync(0xffffffffu, hw, idx & 31);
  const int hs = __shfl_sync(0xffffffffu, start, idx & 31);
  const int k = base + lane - hs;
  const int w0 = (int)((w_hit >> 11) & 31u);
  j = (int)(w_hit & 127u);
  pix = (k < w0) ? ((int)((w_hit >> 7) & 15u) + k) : (16 + (int)((w_hit >> 16) & 15u) + (k - w0));
  return base + lane < total;
}

constexpr uint32_t PAD_WORD = 0x01010101u * (uint32_t)PAD_ENTRY;  // four PAD_ENTRY indices

// ---- producer loop -----------------------------------------------------------------------------------
// The producer never waits for its own gathers: after issuing a batch every lane executes
// cp.async.mbarrier.arrive.noinc on full[slot], so the hardware arrives on the barrier when that lane's copies
// have landed (the Ampere-era equivalent of a TMA complete_tx).  The only thing the producer blocks on is
// empty[slot] — the back-pressure of the ring — so up to STAGES batches of gathers are in flight and a landed
// batch is never held back by the producer being busy elsewhere.  full[] is initialised with 32 expected arrivals.
// (Tried and measured slower: hardware named barriers, bar.arrive / bar.sync.  They do not poll, but the producer
// then has to wait for its copies itself and can only signal a landed batch between two blocking waits, which
// couples the fast consumers to the slowest one: forward 1.36 -> 1.62 ms per 8-view batch.)
// `stop` (forward only) is polled once per batch: when it returns true the producer hands over a stop marker
// instead of data and leaves.
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
constexpr int BAR_EPILOGUE = 1;

#ifdef TGR_MEASURE_STAGING
static __device__ unsigned long long g_producer_empty_wait;   // cycles producers spent blocked on empty[] (per translation unit)
#endif

template <int STAGES, bool REVERSE, bool WITH_IDS, typename StopFn>
__device__ __forceinline__ void producer_loop(const uint32_t* __restrict__ list, int total, int rounds,
                                              const float4* __restrict__ xy_ext, const float4* __restrict__ conic_opacity,
                                              const float4* __restrict__ rgb_depth, float4 (*s_xy)[BL_BATCH + 1],
                                              float4 (*s_co)[BL_BATCH + 1], float4 (*s_cd)[BL_BATCH + 1],
                                              uint32_t (*s_id)[BL_BATCH + 1], uint64_t* s_full, uint64_t* s_empty,
                                              int lane, StopFn stop) {
  uint32_t ids[BL_CHUNKS];
  prod_load_ids(list, total, 0, REVERSE, lane, ids);
  for (int k = 0; k < rounds; ++k) {
    const int st = k % STAGES;
#ifdef TGR_MEASURE_STAGING
    const long long t_e0 = clock64();
#endif
    if (k >= STAGES) mbar_wait(&s_empty[st], ((k / STAGES) - 1) & 1);
#ifdef TGR_MEASURE_STAGING
    if (lane == 0) atomicAdd(&g_producer_empty_wait, (unsigned long long)(clock64() - t_e0));
#endif
    if (stop(st)) {
      cp_async_wait<0>();
      __syncwarp();
      mbar_arrive(&s_full[st]);
      return;
    }
    prod_issue(ids, list, total, k * BL_BATCH, REVERSE, xy_ext, conic_opacity, rgb_depth, s_xy[st], s_co[st], s_cd[st],
               WITH_IDS ? s_id[st] : nullptr, lane);
    cp_async_mbar_arrive(&s_full[st]);
    prod_load_ids(list, total, (k + 1) * BL_BATCH, REVERSE, lane, ids);
  }
  cp_async_wait<0>();  // do not leave with copies in flight
}

// The dummy record PAD_ENTRY points at (written once per CTA, for every ring stage).
__device__ __forceinline__ void init_pad_record(float4* s_xy_stage, float4* s_co_stage, float4* s_cd_stage) {
  s_xy_stage[PAD_ENTRY] = make_float4(0.f, 0.f, 0.f, 0.f);
  s_co_stage[PAD_ENTRY] = make_float4(0.f, 0.f, 0.f, 0.f);   // opacity 0 -> alpha 0 -> never blended
  s_cd_stage[PAD_ENTRY] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace tgr
