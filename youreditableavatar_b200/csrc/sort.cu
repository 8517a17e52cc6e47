// sort.cu — hand-written onesweep LSD radix sort for (u32 key, u32 value) pairs on sm_100a.
//
// Replaces cub::DeviceRadixSort::SortPairs at rasterizer_impl.cu:303-308 and simple_knn.cu:207-213.
// The reference sorts 64-bit (tile<<32 | depth) keys in six 8-bit passes over all R instances.  Here
// the same total order (tile, depth bits, Gaussian id) is produced by two *stable* 32-bit sorts:
// P Gaussians by depth bits, then R instances by tile id (ceil(log2 T) bits) — see DESIGN.md.
//
// One kernel per digit pass ("onesweep"): a CTA takes a 4096-pair tile via an atomic ticket, ranks
// its keys with warp-level match.any multi-split (stable in (warp, item, lane) order), obtains its
// per-digit global offset with a decoupled look-back over the preceding tiles, reorders the tile in
// shared memory and writes digit-contiguous runs with coalesced stores.  An upfront histogram kernel
// produces the per-pass digit totals for all passes in one read of the keys.
#include <algorithm>
#include "common.cuh"

namespace tgr {

constexpr uint32_t LB_FLAG_AGG = 1u << 30;   // tile aggregate available
constexpr uint32_t LB_FLAG_INC = 2u << 30;   // inclusive prefix available
constexpr uint32_t LB_VALUE = (1u << 30) - 1;
constexpr int LB_WINDOW = 8;

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- upfront digit histograms for every pass -------------------------------------------------
// Both kernels serve a BATCH of independent sorts (blockIdx.y = segment): the sorts of a multi-view batch are
// each ~1 M pairs — alone they fill the GPU for a few microseconds per pass and are bound by launch + look-back
// latency; eight of them in one launch stream at HBM speed.
__global__ void __launch_bounds__(256) radix_hist_kernel(const __grid_constant__ SortBatch sb, SortPlan plan) {
  const SortSeg& seg = sb.s[blockIdx.y];
  const uint32_t* __restrict__ keys = seg.keys_a;
  const uint32_t n_host = seg.n_host;
  const uint32_t* __restrict__ n_dev = seg.n_dev;
  uint32_t* __restrict__ ghist = seg.temp;
  __shared__ uint32_t sh[RS_MAX_PASSES * RS_BINS];
  for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  uint32_t n = n_dev ? min(*n_dev, n_host) : n_host;
  const uint32_t stride = gridDim.x * blockDim.x * 4;
  for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * 4; base < n; base += stride) {
    uint32_t k[4];
    int cnt;
    if (base + 4 <= n) {
      uint4 v = *reinterpret_cast<const uint4*>(keys + base);
      k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
      cnt = 4;
    } else {
      cnt = n - base;
      for (int j = 0; j < cnt; ++j) k[j] = keys[base + j];
    }
    for (int j = 0; j < cnt; ++j) {
#pragma unroll
      for (int ps = 0; ps < RS_MAX_PASSES; ++ps)
        if (ps < plan.npasses) atomicAdd(&sh[ps * RS_BINS + ((k[j] >> plan.begin[ps]) & ((1u << plan.bits[ps]) - 1))], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < plan.npasses * RS_BINS; i += blockDim.x)
    if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

// clears the leading words of every segment's temp area (histograms, tickets, look-back): one launch for the batch
// where V cudaMemsetAsync nodes used to be.  The rasterizer's own sorts do not even need it: the kernel in front of
// them clears the area on the side (preprocess_kernel for the depth sort, emit_count_kernel for the tile sort).
__global__ void __launch_bounds__(256) radix_zero_kernel(const __grid_constant__ SortBatch sb, int npasses) {
  const SortSeg& seg = sb.s[blockIdx.y];
  if (seg.n_host == 0) return;
  const uint32_t words = (uint32_t)sort_zero_words(seg.n_host, npasses);
  uint4* t4 = reinterpret_cast<uint4*>(seg.temp);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i * 4 < words; i += gridDim.x * blockDim.x) t4[i] = make_uint4(0u, 0u, 0u, 0u);
}

// ---- one digit pass ------------------------------------------------------------------------------
template <bool IOTA>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_CTAS) radix_pass_kernel(const __grid_constant__ SortBatch sb, int pass, int flip,
                                                                   int begin_bit, int nbits, uint32_t ntiles_max) {
  // per-segment buffers: pass `pass` reads A (flip = 0) or B (flip = 1) and writes the other
  const SortSeg& seg = sb.s[blockIdx.y];
  const uint32_t* __restrict__ keys_in = flip ? seg.keys_b : seg.keys_a;
  const uint32_t* __restrict__ vals_in = flip ? seg.vals_b : seg.vals_a;
  uint32_t* __restrict__ keys_out = flip ? seg.keys_a : seg.keys_b;
  uint32_t* __restrict__ vals_out = flip ? seg.vals_a : seg.vals_b;
  const uint32_t n_host = seg.n_host;
  const uint32_t* __restrict__ n_dev = seg.n_dev;
  const uint32_t* __restrict__ ghist = seg.temp + pass * RS_BINS;
  uint32_t* __restrict__ ticket = seg.temp + RS_MAX_PASSES * RS_BINS + pass;
  uint32_t* __restrict__ lookback = seg.temp + RS_MAX_PASSES * RS_BINS + 32 + (uint64_t)pass * sort_ntiles(n_host) * RS_BINS;
  extern __shared__ uint32_t smem[];
  uint32_t* s_keys = smem;                          // [RS_TILE]
  uint32_t* s_vals = s_keys + RS_TILE;              // [RS_TILE]
  uint32_t* s_whist = s_vals + RS_TILE;             // [8][256]
  uint32_t* s_goff = s_whist + (RS_THREADS / 32) * RS_BINS;  // [256]
  uint32_t* s_lstart = s_goff + RS_BINS;            // [256]
  uint32_t* s_scan = s_lstart + RS_BINS;            // [16]
  __shared__ uint32_t s_tile;

  const uint32_t n = n_dev ? min(*n_dev, n_host) : n_host;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = RS_THREADS / 32;

  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < NW * RS_BINS; i += RS_THREADS) s_whist[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t base64 = (uint64_t)tile * RS_TILE;
  if (base64 >= n) return;
  const uint32_t base = (uint32_t)base64;
  const uint32_t count = min((uint32_t)RS_TILE, n - base);
  const uint32_t mask = (1u << nbits) - 1u;

  // ---- load keys, warp-striped: item i of lane l is element wbase + 32*i + l -----------------
  const uint32_t wbase = base + warp * (32 * RS_IPT);
  uint32_t key[RS_IPT];
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    const uint32_t idx = wbase + i * 32 + lane;
    key[i] = idx < n ? keys_in[idx] : 0xffffffffu;
  }

  // ---- stable rank within the warp ---------------------------------------------------------------
  // All match.any are issued first (independent, their latency overlaps); only the shared-memory
  // running-count update is a serial chain over the items.
  uint16_t rnk[RS_IPT];
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t* whist = s_whist + warp * RS_BINS;
  uint32_t peers[RS_IPT];
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) peers[i] = __match_any_sync(0xffffffffu, (key[i] >> begin_bit) & mask);
  // elements beyond n exist only in the last tile and are the highest positions: lanes >= nvalid_i
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    const uint32_t first = wbase + i * 32;
    const uint32_t nv = first >= n ? 0u : min(32u, n - first);
    const uint32_t vm = nv >= 32u ? 0xffffffffu : ((1u << nv) - 1u);
    const uint32_t d = (key[i] >> begin_bit) & mask;
    const uint32_t old = whist[d];
    __syncwarp();
    rnk[i] = (uint16_t)(old + __popc(peers[i] & lt_mask));
    if (lane == (__ffs(peers[i]) - 1)) whist[d] = old + __popc(peers[i] & vm);
    __syncwarp();
  }
  __syncthreads();

  // ---- per digit (thread == digit): exclusive scan over the warps, tile total -------------------
  uint32_t total = 0;
  {
    const int d = tid;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const uint32_t t = s_whist[w * RS_BINS + d];
      s_whist[w * RS_BINS + d] = total;
      total += t;
    }
  }

  // ---- decoupled look-back: exclusive prefix of this digit over preceding tiles ------------------
  // Windowed: LB_WINDOW predecessors are read at once (independent loads, one L2 round trip per window).
  // With every tile of a 1M-key sort resident in a single wave, a one-predecessor-per-hop walk made the
  // pass a ~120-hop serial chain (ncu: 12-18 % issue utilisation, ~20 us per pass regardless of bytes).
  uint32_t excl = 0;
  {
    uint32_t* lb = lookback + (uint64_t)tile * RS_BINS + tid;
    if (tile == 0) {
      st_volatile_u32(lb, LB_FLAG_INC | total);
    } else {
      st_volatile_u32(lb, LB_FLAG_AGG | total);
      int t = (int)tile - 1;  // nearest predecessor not yet accounted for
      bool done = false;
      while (!done) {
        uint32_t v[LB_WINDOW];
#pragma unroll
        for (int i = 0; i < LB_WINDOW; ++i) {
          const int idx = t - i;
          v[i] = idx >= 0 ? ld_volatile_u32(lookback + (uint64_t)idx * RS_BINS + tid) : LB_FLAG_INC;
        }
        int used = 0;
#pragma unroll
        for (int i = 0; i < LB_WINDOW; ++i) {
          if (done || used != i) continue;
          if ((v[i] & ~LB_VALUE) == 0) continue;  // not published yet: re-read from here
          excl += v[i] & LB_VALUE;
          used = i + 1;
          if (v[i] & LB_FLAG_INC) done = true;
        }
        t -= used;
      }
      st_volatile_u32(lb, LB_FLAG_INC | (excl + total));
    }
  }

  // ---- block exclusive scans over the 256 digits: tile-local start, global digit base ------------
  {
    const uint32_t gh = ghist[tid];
    uint32_t a = total, b = gh;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t ta = __shfl_up_sync(0xffffffffu, a, o);
      uint32_t tb = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= o) { a += ta; b += tb; }
    }
    if (lane == 31) { s_scan[warp] = a; s_scan[8 + warp] = b; }
    __syncthreads();
    uint32_t offa = 0, offb = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
      if (w < warp) { offa += s_scan[w]; offb += s_scan[8 + w]; }
    const uint32_t lstart = offa + a - total;   // exclusive
    const uint32_t gbase = offb + b - gh;       // exclusive
    s_lstart[tid] = lstart;
    s_goff[tid] = gbase + excl - lstart;
  }
  __syncthreads();

  // ---- scatter keys into tile-sorted order in shared memory --------------------------------------
  uint16_t pos[RS_IPT];
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    const uint32_t idx = wbase + i * 32 + lane;
    const uint32_t d = (key[i] >> begin_bit) & mask;
    const uint32_t ps = s_lstart[d] + s_whist[warp * RS_BINS + d] + rnk[i];
    pos[i] = (uint16_t)ps;
    if (idx < n) s_keys[ps] = key[i];
  }
  // values ride along
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    const uint32_t idx = wbase + i * 32 + lane;
    if (idx < n) s_vals[pos[i]] = IOTA ? idx : vals_in[idx];
  }
  __syncthreads();

  // ---- coalesced write-out: position p of the tile goes to s_goff[digit] + p --------------------
#pragma unroll
  for (int j = 0; j < RS_IPT; ++j) {
    const uint32_t ps = j * RS_THREADS + tid;
    if (ps < count) {
      const uint32_t k = s_keys[ps];
      const uint32_t d = (k >> begin_bit) & mask;
      const uint32_t o = s_goff[d] + ps;
      keys_out[o] = k;
      vals_out[o] = s_vals[ps];
    }
  }
}

constexpr size_t RS_SMEM = (size_t)(2 * RS_TILE + (RS_THREADS / 32) * RS_BINS + 2 * RS_BINS + 16) * 4;

// Sorts every segment of `sb` (same bit range for all).  Results land in the B buffers when the plan has an odd
// number of passes (*result_in_b), else in A.
int launch_sort_pairs_batch(const SortBatch& sb, bool iota_vals, int begin_bit, int end_bit, cudaStream_t s,
                            bool* result_in_b, bool temp_is_zero) {
  SortPlan plan = make_sort_plan(begin_bit, end_bit);
  if (result_in_b) *result_in_b = (plan.npasses & 1) != 0;
  uint64_t n_max = 0;
  for (int i = 0; i < sb.V; ++i) n_max = std::max<uint64_t>(n_max, sb.s[i].n_host);
  if (sb.V <= 0 || n_max == 0 || plan.npasses == 0) return 0;
  if (n_max >= (1ull << 30)) { set_error("sort: n=%llu exceeds 2^30", (unsigned long long)n_max); return 1; }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(radix_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM);
    cudaFuncSetAttribute(radix_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM);
    attr_set = true;
  }
  const uint64_t ntiles = sort_ntiles(n_max);
  if (!temp_is_zero) {
    radix_zero_kernel<<<dim3(64, sb.V), 256, 0, s>>>(sb, plan.npasses);
    count_launch();
  }
  const int hist_blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((n_max + 256 * 16 - 1) / (256 * 16),
                                                                         (uint64_t)NUM_SM * 8 / sb.V));
  radix_hist_kernel<<<dim3(hist_blocks, sb.V), 256, 0, s>>>(sb, plan);
  count_launch();
  for (int ps = 0; ps < plan.npasses; ++ps) {
    const dim3 grid((unsigned)ntiles, sb.V);
    if (ps == 0 && iota_vals)
      radix_pass_kernel<true><<<grid, RS_THREADS, RS_SMEM, s>>>(sb, ps, ps & 1, plan.begin[ps], plan.bits[ps], (uint32_t)ntiles);
    else
      radix_pass_kernel<false><<<grid, RS_THREADS, RS_SMEM, s>>>(sb, ps, ps & 1, plan.begin[ps], plan.bits[ps], (uint32_t)ntiles);
    count_launch();
  }
  return check_launch("radix_sort", false, s);
}

int launch_sort_pairs(uint64_t n_host, const uint32_t* n_dev, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b,
                      uint32_t* vals_b, bool iota_vals, int begin_bit, int end_bit, uint32_t* temp, cudaStream_t s,
                      bool* result_in_b) {
  SortBatch sb{};
  sb.V = 1;
  sb.s[0] = SortSeg{keys_a, vals_a, keys_b, vals_b, temp, n_dev, (uint32_t)std::min<uint64_t>(n_host, 0xffffffffull)};
  if (n_host >= (1ull << 30)) { set_error("sort: n=%llu exceeds 2^30", (unsigned long long)n_host); return 1; }
  return launch_sort_pairs_batch(sb, iota_vals, begin_bit, end_bit, s, result_in_b);
}

}  // namespace tgr
