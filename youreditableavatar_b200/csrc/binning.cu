// binning.cu — instance emission and tile ranges.
//
//   emit_*        replace duplicateWithKeys (rasterizer_impl.cu:70-111) *and* the InclusiveSum that
//                 feeds it (rasterizer_impl.cu:277): Gaussians are visited in (depth, id) order; emit_count sums
//                 the instances of every 1024-Gaussian tile, emit_scan turns the totals into output offsets,
//                 emit_kernel writes each (Gaussian, tile) instance as a 32-bit tile id + 32-bit Gaussian id.
//                 The tile rectangle was stored by the preprocess kernel, so getRect is not re-evaluated.
//   ranges_kernel replaces identifyTileRanges (rasterizer_impl.cu:116-138) on the tile-sorted ids.
#include <algorithm>
#include "common.cuh"

namespace tgr {

constexpr uint32_t COOP_THRESHOLD = 24;  // rectangles with more tiles than this are written by the whole warp
constexpr int EMIT_STAGE = 4096;         // instances of one CTA assembled in shared memory before the flush (32 KB)

// Per-Gaussian instance counts of one emission tile (1024 Gaussians in depth order), shared by the two passes.
struct EmitItems {
  uint32_t gid[EMIT_IPT];
  ushort4 rc[EMIT_IPT];
  uint32_t cnt[EMIT_IPT];
  uint32_t tsum;
};
__device__ __forceinline__ EmitItems emit_load(const RenderView& rv, uint32_t tile, int tid) {
  EmitItems it;
  it.tsum = 0;
  const uint32_t slot0 = tile * EMIT_TILE + tid * EMIT_IPT;
#pragma unroll
  for (int i = 0; i < EMIT_IPT; ++i) {
    const uint32_t slot = slot0 + i;
    if (slot < (uint32_t)rv.P) {
      it.gid[i] = rv.order[slot];
      it.rc[i] = rv.rect[it.gid[i]];
      it.cnt[i] = (uint32_t)(it.rc[i].z - it.rc[i].x) * (uint32_t)(it.rc[i].w - it.rc[i].y);
    } else {
      it.gid[i] = 0; it.rc[i] = make_ushort4(0, 0, 0, 0); it.cnt[i] = 0;
    }
    it.tsum += it.cnt[i];
  }
  return it;
}

// Pass 1: instances per emission tile -> scan_state[tile].  (The first version of the emission was ONE kernel with
// a chained scan — decoupled look-back, one word per CTA; ncu showed it latency-bound on that chain: 22 of 45 stall
// cycles at the barrier behind the look-back, 0.22 ms for 8 views.  Three dependency-free launches are faster.)
__global__ void __launch_bounds__(EMIT_THREADS) emit_count_kernel(const __grid_constant__ RenderBatch rb) {
  const RenderView& rv = rb.v[blockIdx.y];
  // on the side: clear the temp area of the tile sort that follows (histograms, tickets, look-back flags), CTA b its
  // slice b — a memset node per view less (the grid covers the largest view: every CTA takes part, also the surplus ones)
  {
    const uint32_t words = rv.tile_sort_zero_words;
    const uint32_t per = ((words + gridDim.x - 1) / gridDim.x + 3u) & ~3u;
    const uint32_t lo = blockIdx.x * per, hi = min(lo + per, words);
    uint4* t4 = reinterpret_cast<uint4*>(rv.tile_sort_temp);
    for (uint32_t i = lo / 4 + threadIdx.x; i * 4 < hi; i += blockDim.x) t4[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if ((uint32_t)blockIdx.x * EMIT_TILE >= (uint32_t)rv.P) return;   // the grid is sized for the largest view
  __shared__ uint32_t s_warp[EMIT_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const EmitItems it = emit_load(rv, blockIdx.x, tid);
  const uint32_t wsum = __reduce_add_sync(0xffffffffu, it.tsum);
  if (lane == 0) s_warp[warp] = wsum;
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < EMIT_THREADS / 32; ++w) t += s_warp[w];
    rv.scan_state[blockIdx.x] = t;
  }
}

// Pass 2: exclusive scan of the per-tile totals in place (one CTA per view), overflow check against the capacity.
__global__ void __launch_bounds__(1024) emit_scan_kernel(const __grid_constant__ RenderBatch rb) {
  const RenderView& rv = rb.v[blockIdx.x];
  // this CTA is the one per-view launch that precedes the tile sort: it also clears the view's tile ranges
  // (ranges_kernel only writes the boundaries it finds), which saves a memset node per view
  for (uint32_t t = threadIdx.x; t < rv.T; t += blockDim.x) rv.ranges[t] = make_uint2(0u, 0u);
  const uint32_t ntiles = ((uint32_t)rv.P + EMIT_TILE - 1) / EMIT_TILE;
  if (ntiles == 0) return;
  uint32_t* __restrict__ st = rv.scan_state;
  __shared__ uint32_t s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t per = (ntiles + 1023) / 1024;           // consecutive entries per thread
  const uint32_t first = tid * per;
  uint32_t sum = 0;
  for (uint32_t k = 0; k < per; ++k) sum += (first + k < ntiles) ? st[first + k] : 0u;
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_warp[lane] = winc - w;                             // exclusive warp offsets
    if (lane == 31 && winc > rv.cap) rv.header->overflow = 1u;
  }
  __syncthreads();
  uint32_t run = s_warp[warp] + inc - sum;
  for (uint32_t k = 0; k < per; ++k) {
    if (first + k < ntiles) {
      const uint32_t c = st[first + k];
      st[first + k] = run;
      run += c;
    }
  }
}

// Pass 3: the instances.  A CTA owns emission tile blockIdx.x; its output offset is scan_state[tile].
__global__ void __launch_bounds__(EMIT_THREADS, 4) emit_kernel(const __grid_constant__ RenderBatch rb) {
  const RenderView& rv = rb.v[blockIdx.y];
  const int P = rv.P;
  if ((uint32_t)blockIdx.x * EMIT_TILE >= (uint32_t)P) return;   // the grid is sized for the largest view
  const uint32_t grid_w = (rv.W + TILE - 1) / TILE;
  uint32_t* __restrict__ keys = rv.key_a;
  uint32_t* __restrict__ vals = rv.val_a;
  const uint32_t cap = rv.cap;
  __shared__ uint32_t s_warp[EMIT_THREADS / 32];
  __shared__ uint32_t s_stage_k[EMIT_STAGE], s_stage_v[EMIT_STAGE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = blockIdx.x;
  const uint32_t s_prefix = rv.scan_state[tile];

  const EmitItems it = emit_load(rv, tile, tid);
  const uint32_t tsum = it.tsum;
  // block exclusive scan of tsum
  uint32_t inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t woff = 0, btotal = 0;
#pragma unroll
  for (int w = 0; w < EMIT_THREADS / 32; ++w) {
    if (w < warp) woff += s_warp[w];
    btotal += s_warp[w];
  }
  // A CTA's instances form one contiguous run [s_prefix, s_prefix + btotal) of the output.  Threads own short,
  // unaligned pieces of it (1.9 instances per Gaussian on the avatar scenes), so writing them directly costs one
  // partial 32-byte sector per lane and store; instead the run is assembled in shared memory and flushed with
  // coalesced stores.  Runs longer than the staging buffer (close-up views: huge rectangles) go straight to global.
  const bool staged = btotal <= EMIT_STAGE;
  uint32_t* __restrict__ kdst = staged ? s_stage_k : keys;
  uint32_t* __restrict__ vdst = staged ? s_stage_v : vals;
  const uint32_t lim = staged ? (uint32_t)EMIT_STAGE : cap;
  uint32_t off = (staged ? 0u : s_prefix) + woff + inc - tsum;

#pragma unroll
  for (int i = 0; i < EMIT_IPT; ++i) {
    const uint32_t c = it.cnt[i];
    const uint32_t w = it.rc[i].z - it.rc[i].x;
    if (c != 0 && c <= COOP_THRESHOLD) {
      uint32_t o = off;
      for (uint32_t y = it.rc[i].y; y < it.rc[i].w; ++y)
        for (uint32_t x = it.rc[i].x; x < it.rc[i].z; ++x) {
          if (o < lim) { kdst[o] = y * grid_w + x; vdst[o] = it.gid[i]; }
          ++o;
        }
    }
    // large rectangles: the whole warp writes them with lane-strided, coalesced stores
    uint32_t big = __ballot_sync(0xffffffffu, c > COOP_THRESHOLD);
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      const uint32_t bc = __shfl_sync(0xffffffffu, c, src);
      const uint32_t bw = __shfl_sync(0xffffffffu, w, src);
      const uint32_t bx = __shfl_sync(0xffffffffu, (uint32_t)it.rc[i].x, src);
      const uint32_t by = __shfl_sync(0xffffffffu, (uint32_t)it.rc[i].y, src);
      const uint32_t bo = __shfl_sync(0xffffffffu, off, src);
      const uint32_t bg = __shfl_sync(0xffffffffu, it.gid[i], src);
      for (uint32_t k = lane; k < bc; k += 32) {
        const uint32_t o = bo + k;
        if (o < lim) { kdst[o] = (by + k / bw) * grid_w + (bx + k % bw); vdst[o] = bg; }
      }
    }
    off += c;
  }
  if (staged) {
    __syncthreads();
    const uint32_t base = s_prefix;
    for (uint32_t k = tid; k < btotal; k += EMIT_THREADS) {
      const uint32_t o = base + k;
      if (o < cap) { keys[o] = s_stage_k[k]; vals[o] = s_stage_v[k]; }
    }
  }
}

int launch_emit(const RenderBatch& rb, cudaStream_t s) {
  int ntiles = 0;
  for (int v = 0; v < rb.V; ++v) ntiles = std::max(ntiles, (rb.v[v].P + EMIT_TILE - 1) / EMIT_TILE);
  if (ntiles == 0) return 0;
  emit_count_kernel<<<dim3(ntiles, rb.V), EMIT_THREADS, 0, s>>>(rb);
  emit_scan_kernel<<<rb.V, 1024, 0, s>>>(rb);
  emit_kernel<<<dim3(ntiles, rb.V), EMIT_THREADS, 0, s>>>(rb);
  count_launch(3);
  return check_launch("emit", false, s);
}

// per-tile [start,end) in the tile-sorted instance list; ranges are zero-initialised by emit_scan_kernel (or a
// memset when nothing is emitted).  Four list entries per thread (one 16-byte load): a quarter of the CTAs of the
// one-entry-per-thread version, whose cost was CTA scheduling, not memory.
constexpr int RANGE_IPT = 4;
__global__ void __launch_bounds__(256) ranges_kernel(const __grid_constant__ RenderBatch rb) {
  const RenderView& rv = rb.v[blockIdx.y];
  const uint32_t* __restrict__ keys = rv.sorted_keys;
  const uint32_t cap = rv.cap;
  const GeomHeader* __restrict__ header = rv.header;
  uint2* __restrict__ ranges = rv.ranges;
  const uint32_t L = min(header->num_rendered, cap);
  if (header->overflow) return;
  const uint32_t idx0 = (blockIdx.x * blockDim.x + threadIdx.x) * RANGE_IPT;
  if (idx0 >= L) return;
  uint32_t k[RANGE_IPT];
  if (idx0 + RANGE_IPT <= L) {
    const uint4 q = *reinterpret_cast<const uint4*>(keys + idx0);
    k[0] = q.x; k[1] = q.y; k[2] = q.z; k[3] = q.w;
  } else {
#pragma unroll
    for (int i = 0; i < RANGE_IPT; ++i) k[i] = idx0 + i < L ? keys[idx0 + i] : 0u;
  }
  uint32_t prev = idx0 ? keys[idx0 - 1] : 0u;
#pragma unroll
  for (int i = 0; i < RANGE_IPT; ++i) {
    const uint32_t idx = idx0 + i;
    if (idx < L) {
      const uint32_t cur = k[i];
      if (idx == 0) {
        ranges[cur].x = 0;
      } else if (cur != prev) {
        ranges[prev].y = idx;
        ranges[cur].x = idx;
      }
      if (idx == L - 1) ranges[cur].y = L;
      prev = cur;
    }
  }
}

int launch_ranges(const RenderBatch& rb, bool zero_first, cudaStream_t s) {
  uint32_t cap_max = 0;
  for (int v = 0; v < rb.V; ++v) {
    if (zero_first) cudaMemsetAsync(rb.v[v].ranges, 0, (size_t)rb.v[v].T * sizeof(uint2), s);
    cap_max = std::max(cap_max, rb.v[v].cap);
  }
  if (cap_max == 0) return 0;
  const uint32_t per_cta = 256 * RANGE_IPT;
  ranges_kernel<<<dim3((cap_max + per_cta - 1) / per_cta, rb.V), 256, 0, s>>>(rb);
  count_launch();
  return check_launch("ranges", false, s);
}

// Heaviest-first tile order.  The block scheduler hands consecutive CTAs to different SMs, so rendering tiles
// in descending list-length order deals every SM a similar mix of long and short tiles (the avatar covers
// only ~15-20% of the tiles; with row-major order ncu showed SMs idle ~48% of the blend kernels).
// One CTA, bucketed counting sort on weight/8 (4096 buckets, exact order inside a bucket is irrelevant).
constexpr int TO_BUCKETS = 4096;
__global__ void __launch_bounds__(1024) tile_order_kernel(const __grid_constant__ RenderBatch rb) {
  const RenderView& rv = rb.v[blockIdx.x];   // one CTA per view
  const uint2* __restrict__ ranges = rv.ranges;
  const uint32_t* tile_last = nullptr;
  const uint32_t T = rv.T;
  uint32_t* __restrict__ order = rv.order_fwd;
  uint32_t* __restrict__ queue_counters = blockIdx.x == 0 ? rb.queue_counters : nullptr;
  uint32_t* __restrict__ seg_base = rv.seg_base;
  __shared__ uint32_t s_cnt[TO_BUCKETS];
  __shared__ uint32_t s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < TO_BUCKETS; i += 1024) s_cnt[i] = 0;
  if (tid == 0) rv.unit_count[3] = 0;   // the forward that follows clears the packed gradient rows (blend_fwd.cu)
  if (queue_counters != nullptr && tid < MAX_QUEUES) queue_counters[tid] = 0;
  __syncthreads();
  auto bucket_of = [&](uint32_t t) {
    const uint2 r = ranges[t];
    uint32_t w = r.y - r.x;
    if (tile_last) w = min(w, tile_last[t]);
    // descending: heaviest -> bucket 0
    return (uint32_t)(TO_BUCKETS - 1) - min(w >> 3, (uint32_t)(TO_BUCKETS - 1));
  };
  for (uint32_t t = tid; t < T; t += 1024) atomicAdd(&s_cnt[bucket_of(t)], 1u);
  __syncthreads();
  // exclusive scan of the 4096 counts: 4 per thread
  uint32_t c[4], sum = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { c[k] = s_cnt[tid * 4 + k]; sum += c[k]; }
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    s_warp[lane] = wi - w;
  }
  __syncthreads();
  uint32_t base = s_warp[warp] + inc - sum;
#pragma unroll
  for (int k = 0; k < 4; ++k) { s_cnt[tid * 4 + k] = base; base += c[k]; }
  __syncthreads();
  for (uint32_t t = tid; t < T; t += 1024) order[atomicAdd(&s_cnt[bucket_of(t)], 1u)] = t;

  // checkpoint slots: seg_base[t] = exclusive scan (tile-index order) of ceil(len_t / SEG); seg_base[T] = total
  if (seg_base != nullptr) {
    __shared__ uint32_t s_carry;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < T; t0 += 1024) {
      const uint32_t t = t0 + tid;
      uint32_t n = 0;
      if (t < T) { const uint2 r = ranges[t]; n = (r.y - r.x + SEG - 1) / SEG; }
      uint32_t incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      uint32_t woff = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < 32; ++w) { if (w < warp) woff += s_warp[w]; tot += s_warp[w]; }
      const uint32_t carry = s_carry;
      if (t < T) seg_base[t] = carry + woff + incl - n;
      __syncthreads();
      if (tid == 0) s_carry = carry + tot;
      __syncthreads();
    }
    if (tid == 0) seg_base[T] = s_carry;
  }
}

// Backward work units: tile t contributes ceil(min(tile_last_t, len_t) / SEG) units (tile, segment).
__global__ void __launch_bounds__(1024) unit_build_kernel(const __grid_constant__ RenderBatch rb) {
  const RenderView& rv = rb.v[blockIdx.x];   // one CTA per view
  const uint2* __restrict__ ranges = rv.ranges;
  const uint32_t* __restrict__ tile_last = rv.tile_last;
  const uint32_t T = rv.T;
  uint2* __restrict__ units = rv.units;
  const uint32_t units_cap = rv.units_cap;
  uint32_t* __restrict__ unit_count = rv.unit_count;
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  // The packed gradient rows are cleared by the forward (blend_fwd.cu).  A second backward on the same forward
  // (retain_graph) finds them used: it clears them here — one CTA, slow, but only on that rare path.
  if (unit_count[3] != 0u) {
    float4* g4 = reinterpret_cast<float4*>(rv.grad_acc);
    const size_t n4 = (size_t)rv.P * GRAD_ACC / 4;
    for (size_t i = tid; i < n4; i += 1024) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  if (tid == 0) {
    unit_count[3] = 1u;
    unit_count[1] = 0u;   // ticket counter of blend_bwd's persistent CTAs (the one of view 0 serves the batch)
  }
  for (uint32_t t0 = 0; t0 < T; t0 += 1024) {
    const uint32_t t = t0 + tid;
    uint32_t n = 0;
    if (t < T) { const uint2 r = ranges[t]; n = (min(r.y - r.x, tile_last[t]) + SEG - 1) / SEG; }
    uint32_t incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) { if (w < warp) woff += s_warp[w]; tot += s_warp[w]; }
    const uint32_t carry = s_carry;
    uint32_t u = carry + woff + incl - n;
    // heaviest segment (the deepest one is usually shorter) first is irrelevant: units are <= SEG entries each
    for (uint32_t k = 0; k < n; ++k, ++u)
      if (u < units_cap) units[u] = make_uint2(t, k);
    __syncthreads();
    if (tid == 0) s_carry = carry + tot;
    __syncthreads();
  }
  if (tid == 0) *unit_count = min(s_carry, units_cap);
}

int launch_unit_build(const RenderBatch& rb, cudaStream_t s) {
  unit_build_kernel<<<rb.V, 1024, 0, s>>>(rb);
  count_launch();
  return check_launch("unit_build", false, s);
}

uint32_t num_queues() {
  static thread_local int cached_dev = -1;
  static thread_local uint32_t cached = NUM_SM;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    int n = NUM_SM;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached = (uint32_t)std::min(std::max(n, 1), MAX_QUEUES);
    cached_dev = dev;
  }
  return cached;
}

int launch_tile_order(const RenderBatch& rb, cudaStream_t s) {
  tile_order_kernel<<<rb.V, 1024, 0, s>>>(rb);
  count_launch();
  return check_launch("tile_order", false, s);
}

}  // namespace tgr
