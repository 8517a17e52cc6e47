// c_abi.cu — extern "C" boundary (include/tetgs_rast.h) and stage orchestration.  No torch types.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <nvtx3/nvToolsExt.h>   // header-only (dlopens the injection library of an attached profiler, else no-ops)
#include "common.cuh"

namespace tgr {

static thread_local char g_err[512] = "";

static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what, bool debug, cudaStream_t s) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && debug) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

// ---- per-stage profiling ---------------------------------------------------------------------------
constexpr int PROF_MAX = 4096;  // launches per stage between two collects
struct ProfState {
  bool on = false;
  cudaEvent_t ev[TGR_NUM_STAGES][PROF_MAX][2];
  int made[TGR_NUM_STAGES] = {};
  int used[TGR_NUM_STAGES] = {};
};
static ProfState g_prof;

// NVTX range per stage (the reference has none): `nsys` / `ncu --nvtx` group the launches of a batch by
// stage name; without a profiler attached the calls return immediately.
static const char* const kStageNames[TGR_NUM_STAGES] = {"tgr:preprocess", "tgr:depth_sort", "tgr:emit", "tgr:tile_sort",
                                                        "tgr:ranges", "tgr:blend_fwd", "tgr:blend_bwd", "tgr:preprocess_bwd"};

void prof_begin(int stage, cudaStream_t s) {
  nvtxRangePushA(kStageNames[stage]);
  if (!g_prof.on) return;
  int& u = g_prof.used[stage];
  if (u >= PROF_MAX) return;
  if (u >= g_prof.made[stage]) {
    cudaEventCreate(&g_prof.ev[stage][u][0]);
    cudaEventCreate(&g_prof.ev[stage][u][1]);
    g_prof.made[stage] = u + 1;
  }
  cudaEventRecord(g_prof.ev[stage][u][0], s);
}
void prof_end(int stage, cudaStream_t s) {
  nvtxRangePop();
  if (!g_prof.on) return;
  int& u = g_prof.used[stage];
  if (u >= PROF_MAX) return;
  cudaEventRecord(g_prof.ev[stage][u][1], s);
  ++u;
}

static cudaEvent_t count_event() {
  static thread_local cudaEvent_t ev[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (!ev[dev]) cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming);
  return ev[dev];
}

static int validate(const tgr_params* p, bool need_bin, uint64_t cap) {
  if (!p) { set_error("null params"); return 1; }
  if (p->P < 0 || p->W <= 0 || p->H <= 0) { set_error("bad sizes P=%d W=%d H=%d", p->P, p->W, p->H); return 1; }
  if (p->D < 0 || p->D > 3) { set_error("SH degree %d not in 0..3", p->D); return 1; }
  if (!p->geom_buffer || p->geom_bytes < tgr_geom_bytes(p->P)) { set_error("geom buffer too small"); return 1; }
  if (!p->image_buffer || p->image_bytes < tgr_image_bytes(p->W, p->H)) { set_error("image buffer too small"); return 1; }
  if (need_bin && (!p->binning_buffer || p->binning_bytes < tgr_binning_bytes(p->P, cap, p->W, p->H))) {
    set_error("binning buffer too small for capacity %llu", (unsigned long long)cap);
    return 1;
  }
  if ((p->W + TILE - 1) / TILE > 65535 || (p->H + TILE - 1) / TILE > 65535) { set_error("image too large"); return 1; }
  return 0;
}

}  // namespace tgr

using namespace tgr;

extern "C" {

int tgr_abi_version(void) { return TGR_ABI_VERSION; }

uint64_t tgr_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

int tgr_profile_enable(int on) {
  g_prof.on = on != 0;
  for (int i = 0; i < TGR_NUM_STAGES; ++i) g_prof.used[i] = 0;
  return 0;
}

int tgr_profile_collect(float* sum_ms, int32_t* launches) {
  for (int st = 0; st < TGR_NUM_STAGES; ++st) {
    float sum = 0.f;
    for (int i = 0; i < g_prof.used[st]; ++i) {
      cudaError_t e = cudaEventSynchronize(g_prof.ev[st][i][1]);
      float ms = 0.f;
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, g_prof.ev[st][i][0], g_prof.ev[st][i][1]);
      if (e != cudaSuccess) { set_error("profile_collect: %s", cudaGetErrorString(e)); return 2; }
      sum += ms;
    }
    if (sum_ms) sum_ms[st] = sum;
    if (launches) launches[st] = g_prof.used[st];
    g_prof.used[st] = 0;
  }
  return 0;
}
const char* tgr_last_error(void) { return g_err; }

uint64_t tgr_geom_bytes(int32_t P) { return carve_geom(nullptr, P).bytes; }
uint64_t tgr_image_bytes(int32_t W, int32_t H) { return carve_image(nullptr, W, H).bytes; }
uint64_t tgr_binning_bytes(int32_t P, uint64_t cap, int32_t W, int32_t H) { return carve_bin(nullptr, P, cap, W, H).bytes; }
uint64_t tgr_sort_temp_bytes(uint64_t n) { return sort_temp_bytes(n); }
uint64_t tgr_binning_capacity(int32_t P, uint64_t bytes, int32_t W, int32_t H) {
  // tgr_binning_bytes is non-decreasing in the capacity: binary search for the largest one that fits
  if (carve_bin(nullptr, P, 0, W, H).bytes > bytes) return 0;
  uint64_t lo = 0, hi = (1ull << 30) - 1;
  while (lo < hi) {
    const uint64_t mid = lo + (hi - lo + 1) / 2;
    if (carve_bin(nullptr, P, mid, W, H).bytes <= bytes) lo = mid; else hi = mid - 1;
  }
  return lo;
}

static int check_gaussians(const tgr_params* p, const tgr_binding* bind) {
  if (!bind) {
    if (!p->means3D || !p->opacities) { set_error("means3D/opacities missing"); return 1; }
    if (!p->cov3D_precomp && (!p->scales || !p->rotations)) { set_error("need scales+rotations or cov3D_precomp"); return 1; }
  }
  if (!p->colors_precomp && (!p->shs || p->M < (p->D + 1) * (p->D + 1))) { set_error("need colors_precomp or shs with M >= (D+1)^2"); return 1; }
  return 0;
}

// every view of a batch must describe the same Gaussians (only cameras, sizes, workspaces and outputs differ)
static int check_same_gaussians(const tgr_params* a, const tgr_params* b) {
  if (a->P != b->P || a->D != b->D || a->M != b->M || a->means3D != b->means3D || a->shs != b->shs ||
      a->colors_precomp != b->colors_precomp || a->opacities != b->opacities || a->scales != b->scales ||
      a->rotations != b->rotations || a->cov3D_precomp != b->cov3D_precomp || a->scale_modifier != b->scale_modifier) {
    set_error("batch: all views must share the same Gaussian tensors, P, D, M and scale_modifier");
    return 1;
  }
  return 0;
}

// (depth bits, id) order of the Gaussians of every view: positive floats compare like their bit patterns (bit 31
// is 0).  One batched launch sequence for all views.
static int depth_bits_of(const tgr_params& p) { return (p.depth_key_bits > 0 && p.depth_key_bits < 32) ? p.depth_key_bits : 32; }
// where the (depth, id) order of the Gaussians ends up: an odd number of radix passes leaves it in the alternate buffer
static bool depth_order_in_alt(const tgr_params& p) { return (make_sort_plan(0, depth_bits_of(p)).npasses & 1) != 0; }

static int depth_sort_views(const tgr_params* views, int32_t n, cudaStream_t s, bool temp_is_zero) {
  int bits = 1;
  for (int32_t v = 0; v < n; ++v) bits = std::max(bits, depth_bits_of(views[v]));
  for (int32_t v = 0; v < n; ++v)
    if (depth_bits_of(views[v]) != bits) { set_error("batch: all views must use the same depth_key_bits"); return 1; }
  for (int32_t v0 = 0; v0 < n; v0 += MAX_BATCH) {
    SortBatch sb{};
    sb.V = std::min<int32_t>(MAX_BATCH, n - v0);
    for (int32_t k = 0; k < sb.V; ++k) {
      const tgr_params& p = views[v0 + k];
      GeomView g = carve_geom(p.geom_buffer, p.P);
      sb.s[k] = SortSeg{g.depth_key, g.order, g.key_alt, g.val_alt, g.sort_temp, nullptr, (uint32_t)p.P};
    }
    bool in_b = false;
    prof_begin(TGR_STAGE_DEPTH_SORT, s);
    if (int rc = launch_sort_pairs_batch(sb, true, 0, bits, s, &in_b, temp_is_zero)) return rc;   // cleared by preprocess_kernel
    prof_end(TGR_STAGE_DEPTH_SORT, s);
    if (in_b != depth_order_in_alt(views[v0])) { set_error("internal: depth order is not where the binning expects it"); return 3; }
  }
  return 0;
}

static RenderView make_render_view(const tgr_params& p, uint64_t cap, bool tile_sorted_in_b) {
  GeomView g = carve_geom(p.geom_buffer, p.P);
  BinView b = carve_bin(p.binning_buffer, p.P, cap, p.W, p.H);
  ImageView im = carve_image(p.image_buffer, p.W, p.H);
  RenderView r{};
  r.P = p.P; r.W = p.W; r.H = p.H;
  r.T = (uint32_t)((p.W + TILE - 1) / TILE) * ((p.H + TILE - 1) / TILE);
  r.cap = (uint32_t)std::min<uint64_t>(cap, 0xffffffffull);
  r.units_cap = (uint32_t)std::min<uint64_t>(b.units_cap, 0x7fffffffull);
  r.header = g.header; r.order = depth_order_in_alt(p) ? g.val_alt : g.order; r.rect = g.rect;
  r.xy_ext = g.xy_ext; r.conic_opacity = g.conic_opacity; r.rgb_depth = g.rgb_depth; r.scan_state = g.scan_state;
  r.key_a = b.key_a; r.val_a = b.val_a;
  r.tile_sort_temp = b.sort_temp;
  r.tile_sort_zero_words = (uint32_t)sort_zero_words(cap, make_sort_plan(0, tile_bits(r.T)).npasses);
  r.sorted_keys = tile_sorted_in_b ? b.key_b : b.key_a;
  r.point_list = tile_sorted_in_b ? b.val_b : b.val_a;
  r.grad_acc = b.grad_acc; r.ckpt = b.ckpt; r.ckpt_z = b.ckpt_z; r.units = b.units;
  r.ranges = im.ranges; r.tile_last = im.tile_last; r.order_fwd = im.order_fwd; r.seg_base = im.seg_base;
  r.unit_count = im.unit_count; r.final_T = im.final_T; r.n_contrib = im.n_contrib; r.final_state = im.final_state;
  r.final_z = im.final_z;
  r.bg = p.background; r.out_color = p.out_color; r.out_depth = p.out_depth; r.out_alpha = p.out_alpha;
  r.dL_dpix = p.dL_dout_color; r.dL_ddepth = p.dL_dout_depth; r.dL_dalpha = p.dL_dout_alpha;
  return r;
}

// The tile sort of a batch must use one bit range; views whose tile counts need a different number of bits are
// rendered in separate groups (square same-size batches — the normal case — form one group).
static bool same_group(const tgr_params& a, const tgr_params& b) {
  auto tb = [](const tgr_params& p) { return tile_bits((uint32_t)((p.W + TILE - 1) / TILE) * ((p.H + TILE - 1) / TILE)); };
  const bool ea = a.extras && a.out_depth && a.out_alpha, eb = b.extras && b.out_depth && b.out_alpha;
  return tb(a) == tb(b) && ea == eb;
}

// emit -> tile sort -> ranges -> blend for a group of <= MAX_BATCH views, every stage ONE launch for the group
static int render_group(const tgr_params* views, const uint64_t* caps, int32_t n, cudaStream_t s) {
  const tgr_params& p0 = views[0];
  const uint32_t T0 = (uint32_t)((p0.W + TILE - 1) / TILE) * ((p0.H + TILE - 1) / TILE);
  SortPlan plan = make_sort_plan(0, tile_bits(T0));
  const bool in_b = (plan.npasses & 1) != 0;
  RenderBatch rb{};
  rb.V = n;
  bool any = false;
  for (int32_t k = 0; k < n; ++k) {
    rb.v[k] = make_render_view(views[k], caps[k], in_b);
    rb.T_max = std::max(rb.T_max, rb.v[k].T);
    any = any || (views[k].P > 0 && caps[k] > 0);
  }
  rb.queue_counters = carve_image(p0.image_buffer, p0.W, p0.H).queue_counters;
  if (any) {
    prof_begin(TGR_STAGE_EMIT, s);
    if (int rc = launch_emit(rb, s)) return rc;
    prof_end(TGR_STAGE_EMIT, s);
    SortBatch sb{};
    sb.V = n;
    for (int32_t k = 0; k < n; ++k) {
      BinView b = carve_bin(views[k].binning_buffer, views[k].P, caps[k], views[k].W, views[k].H);
      GeomView g = carve_geom(views[k].geom_buffer, views[k].P);
      sb.s[k] = SortSeg{b.key_a, b.val_a, b.key_b, b.val_b, b.sort_temp, &g.header->num_rendered, rb.v[k].cap};
    }
    prof_begin(TGR_STAGE_TILE_SORT, s);
    if (int rc = launch_sort_pairs_batch(sb, false, 0, tile_bits(T0), s, nullptr, /*temp_is_zero=*/true)) return rc;   // emit_count_kernel
    prof_end(TGR_STAGE_TILE_SORT, s);
  }
  prof_begin(TGR_STAGE_RANGES, s);
  if (int rc = launch_ranges(rb, !any, s)) return rc;   // emit_scan_kernel cleared the ranges when anything was emitted
  prof_end(TGR_STAGE_RANGES, s);
  const bool extras = p0.extras && p0.out_depth && p0.out_alpha;
  prof_begin(TGR_STAGE_BLEND_FWD, s);
  if (int rc = launch_blend_fwd(rb, extras, p0.debug != 0, s)) return rc;
  prof_end(TGR_STAGE_BLEND_FWD, s);
  return 0;
}

extern "C++" {
template <typename GroupFn>
static int for_each_group(const tgr_params* views, const uint64_t* caps, int32_t n, GroupFn fn) {
  int32_t v0 = 0;
  while (v0 < n) {
    int32_t v1 = v0 + 1;
    while (v1 < n && v1 - v0 < MAX_BATCH && same_group(views[v0], views[v1])) ++v1;
    if (int rc = fn(views + v0, caps + v0, v1 - v0)) return rc;
    v0 = v1;
  }
  return 0;
}
}  // extern "C++"

struct HeaderBatch {
  int32_t V;
  GeomHeader* h[MAX_BATCH];
};
__global__ void header_init_kernel(const __grid_constant__ HeaderBatch hb) {
  const int v = threadIdx.x >> 5, w = threadIdx.x & 31;
  if (v < hb.V) reinterpret_cast<uint32_t*>(hb.h[v])[w] = (w == 5) ? 0xffffffffu : 0u;   // 128-byte header = 32 words; [5] = key_and
}

// One preprocess launch per chunk of <= TGR_MAX_BATCH views; instance counts go to each view's pinned slot.
static int preprocess_views(const tgr_params* views, int32_t n, const tgr_binding* bind, cudaStream_t s) {
  for (int32_t v = 0; v < n; ++v) {
    if (int rc = validate(&views[v], false, 0)) return rc;
    if (v && check_same_gaussians(&views[0], &views[v])) return 1;
  }
  const tgr_params* p0 = &views[0];
  if (p0->P == 0) {
    for (int32_t v = 0; v < n; ++v)
      if (views[v].host_num_rendered) for (int k = 0; k < 8; ++k) views[v].host_num_rendered[k] = 0;
    return 0;
  }
  if (int rc = check_gaussians(p0, bind)) return rc;
  for (int32_t v0 = 0; v0 < n; v0 += MAX_BATCH) {
    ViewBatch vb{};
    vb.V = std::min<int32_t>(MAX_BATCH, n - v0);
    vb.first = 0;
    vb.end = p0->P;
    HeaderBatch hb{};
    hb.V = vb.V;
    for (int32_t k = 0; k < vb.V; ++k) {
      GeomView g = carve_geom(views[v0 + k].geom_buffer, p0->P);
      hb.h[k] = g.header;
      vb.v[k] = make_view_desc(views[v0 + k], g, nullptr);
    }
    header_init_kernel<<<1, 32 * MAX_BATCH, 0, s>>>(hb);   // one launch where V memset nodes used to be
    count_launch();
    prof_begin(TGR_STAGE_PREPROCESS, s);
    if (int rc = launch_preprocess(*p0, bind, vb, s)) return rc;
    prof_end(TGR_STAGE_PREPROCESS, s);
  }
  bool any = false;
  for (int32_t v = 0; v < n; ++v) {
    if (!views[v].host_num_rendered) continue;
    GeomView g = carve_geom(views[v].geom_buffer, p0->P);
    // {num_rendered, overflow (still 0 here), num_visible, prefilter_violation, key_or, key_and, -, -}
    cudaMemcpyAsync(views[v].host_num_rendered, &g.header->num_rendered, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
    any = true;
  }
  if (any) cudaEventRecord(count_event(), s);
  return 0;
}

int tgr_forward_preprocess(const tgr_params* p, const tgr_binding* bind, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!p) { set_error("null params"); return 1; }
  if (int rc = preprocess_views(p, 1, bind, s)) return rc;
  if (p->P == 0) return 0;
  if (int rc = depth_sort_views(p, 1, s, true)) return rc;
  return check_launch("forward_preprocess", p->debug != 0, s);
}

int tgr_forward_preprocess_batch(const tgr_params* views, int32_t n_views, const tgr_binding* bind, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!views || n_views <= 0) { set_error("batch: no views"); return 1; }
  if (int rc = preprocess_views(views, n_views, bind, s)) return rc;
  return check_launch("forward_preprocess_batch", views[0].debug != 0, s);
}

int tgr_forward_depth_sort(const tgr_params* p, void* stream) {
  if (int rc = validate(p, false, 0)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->P == 0) return 0;
  if (int rc = depth_sort_views(p, 1, s, true)) return rc;    // temp cleared by the preprocess kernel
  return check_launch("forward_depth_sort", p->debug != 0, s);
}

int tgr_wait_num_rendered(void) {
  cudaError_t e = cudaEventSynchronize(count_event());
  if (e != cudaSuccess) { set_error("wait_num_rendered: %s", cudaGetErrorString(e)); return 2; }
  return 0;
}

static const uint32_t* sorted_vals(const tgr_params* p, const BinView& b, bool* in_b_out = nullptr) {
  const uint32_t T = (uint32_t)((p->W + TILE - 1) / TILE) * ((p->H + TILE - 1) / TILE);
  SortPlan plan = make_sort_plan(0, tile_bits(T));
  bool in_b = (plan.npasses & 1) != 0;
  if (in_b_out) *in_b_out = in_b;
  return in_b ? b.val_b : b.val_a;
}

int tgr_forward_render(const tgr_params* p, uint64_t cap, void* stream) {
  if (int rc = validate(p, true, cap)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = render_group(p, &cap, 1, s)) return rc;
  return check_launch("forward_render", p->debug != 0, s);
}

int tgr_forward_render_batch(const tgr_params* views, const uint64_t* caps, int32_t n_views, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!views || !caps || n_views <= 0) { set_error("batch: no views"); return 1; }
  for (int32_t v = 0; v < n_views; ++v)
    if (int rc = validate(&views[v], true, caps[v])) return rc;
  if (views[0].P > 0)
    if (int rc = depth_sort_views(views, n_views, s, true)) return rc;
  if (int rc = for_each_group(views, caps, n_views,
                              [&](const tgr_params* g, const uint64_t* c, int32_t n) { return render_group(g, c, n, s); }))
    return rc;
  return check_launch("forward_render_batch", views[0].debug != 0, s);
}

// memset of the packed 2-D gradient rows + unit build + blend backward for a group of views (one launch each)
static int backward_blend_group(const tgr_params* views, const uint64_t* caps, int32_t n, cudaStream_t s) {
  const tgr_params& p0 = views[0];
  const uint32_t T0 = (uint32_t)((p0.W + TILE - 1) / TILE) * ((p0.H + TILE - 1) / TILE);
  const bool in_b = (make_sort_plan(0, tile_bits(T0)).npasses & 1) != 0;
  RenderBatch rb{};
  rb.V = n;
  bool extras = false;
  for (int32_t k = 0; k < n; ++k) {
    if (!views[k].dL_dout_color) { set_error("dL_dout_color missing"); return 1; }
    rb.v[k] = make_render_view(views[k], caps[k], in_b);
    rb.T_max = std::max(rb.T_max, rb.v[k].T);
    extras = extras || (views[k].extras && (views[k].dL_dout_depth || views[k].dL_dout_alpha));
  }
  prof_begin(TGR_STAGE_BLEND_BWD, s);
  if (int rc = launch_blend_bwd(rb, extras, p0.debug != 0, s)) return rc;
  prof_end(TGR_STAGE_BLEND_BWD, s);
  return 0;
}

int tgr_backward(const tgr_params* p, const tgr_binding* bind, uint64_t cap, void* stream) {
  if (int rc = validate(p, true, cap)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->P == 0) return 0;
  if (int rc = backward_blend_group(p, &cap, 1, s)) return rc;
  ViewBatch vb{};
  vb.V = 1;
  vb.first = 0;
  vb.end = p->P;
  vb.v[0] = make_view_desc(*p, carve_geom(p->geom_buffer, p->P), carve_bin(p->binning_buffer, p->P, cap, p->W, p->H).grad_acc);
  prof_begin(TGR_STAGE_PREPROCESS_BWD, s);
  if (int rc = launch_preprocess_bwd(*p, bind, vb, s)) return rc;
  prof_end(TGR_STAGE_PREPROCESS_BWD, s);
  return check_launch("backward", p->debug != 0, s);
}

int tgr_backward_blend(const tgr_params* p, uint64_t cap, void* stream) {
  if (int rc = validate(p, true, cap)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->P == 0) return 0;
  if (int rc = backward_blend_group(p, &cap, 1, s)) return rc;
  return check_launch("backward_blend", p->debug != 0, s);
}

int tgr_backward_blend_batch(const tgr_params* views, const uint64_t* caps, int32_t n_views, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!views || !caps || n_views <= 0) { set_error("batch: no views"); return 1; }
  for (int32_t v = 0; v < n_views; ++v)
    if (int rc = validate(&views[v], true, caps[v])) return rc;
  if (views[0].P == 0) return 0;
  if (int rc = for_each_group(views, caps, n_views, [&](const tgr_params* g, const uint64_t* c, int32_t n) {
        return backward_blend_group(g, c, n, s);
      }))
    return rc;
  return check_launch("backward_blend_batch", views[0].debug != 0, s);
}

int tgr_backward_preprocess_batch(const tgr_params* views, const uint64_t* caps, int32_t n_views, const tgr_binding* bind,
                                  int32_t gaussian_first, int32_t gaussian_count, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!views || !caps || n_views <= 0) { set_error("batch: no views"); return 1; }
  if (gaussian_count <= 0) { gaussian_first = 0; gaussian_count = views[0].P; }
  if (gaussian_first < 0 || (gaussian_first % 256) != 0 || gaussian_first + (int64_t)gaussian_count > views[0].P) {
    set_error("batch: Gaussian range [%d, +%d) must start at a multiple of 256 and lie inside [0, P)", gaussian_first, gaussian_count);
    return 1;
  }
  for (int32_t v = 0; v < n_views; ++v) {
    if (int rc = validate(&views[v], true, caps[v])) return rc;
    if (v && check_same_gaussians(&views[0], &views[v])) return 1;
  }
  if (views[0].P == 0) return 0;
  tgr_params p0 = views[0];  // Gaussians + output gradient tensors + accumulate flag of the batch
  for (int32_t v0 = 0; v0 < n_views; v0 += MAX_BATCH) {
    ViewBatch vb{};
    vb.V = std::min<int32_t>(MAX_BATCH, n_views - v0);
    vb.first = gaussian_first;
    vb.end = gaussian_first + gaussian_count;
    for (int32_t k = 0; k < vb.V; ++k) {
      const tgr_params& pv = views[v0 + k];
      vb.v[k] = make_view_desc(pv, carve_geom(pv.geom_buffer, pv.P),
                               carve_bin(pv.binning_buffer, pv.P, caps[v0 + k], pv.W, pv.H).grad_acc);
    }
    prof_begin(TGR_STAGE_PREPROCESS_BWD, s);
    if (int rc = launch_preprocess_bwd(p0, bind, vb, s)) return rc;
    prof_end(TGR_STAGE_PREPROCESS_BWD, s);
    p0.accumulate = 1;  // later chunks add to what the first one wrote
  }
  return check_launch("backward_preprocess_batch", views[0].debug != 0, s);
}

int tgr_read_header(const void* geom_buffer, uint32_t out[4], void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GeomView g = carve_geom(const_cast<void*>(geom_buffer), 0);
  cudaError_t e = cudaMemcpyAsync(out, g.header, 16, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { set_error("read_header: %s", cudaGetErrorString(e)); return 2; }
  return 0;
}

int tgr_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream) {
  return launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, static_cast<cudaStream_t>(stream));
}

int tgr_sort_pairs_u32(uint64_t n, uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                       int begin_bit, int end_bit, void* temp, uint64_t temp_bytes, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (temp_bytes < sort_temp_bytes(n)) { set_error("sort temp too small"); return 1; }
  bool in_b = false;
  if (int rc = launch_sort_pairs(n, nullptr, keys_in, vals_in, keys_out, vals_out, false, begin_bit, end_bit,
                                 static_cast<uint32_t*>(temp), s, &in_b)) return rc;
  if (!in_b && n > 0) {  // even number of passes: result sits in the input buffers
    cudaMemcpyAsync(keys_out, keys_in, n * 4, cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(vals_out, vals_in, n * 4, cudaMemcpyDeviceToDevice, s);
  }
  return check_launch("sort_pairs_u32", false, s);
}

int tgr_sort_pairs_u32_batch(int32_t n_segments, const uint64_t* n, uint32_t* const* keys_a, uint32_t* const* vals_a,
                             uint32_t* const* keys_b, uint32_t* const* vals_b, int begin_bit, int end_bit,
                             void* const* temps, int32_t* result_in_b, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_segments <= 0 || n_segments > MAX_BATCH || !n || !keys_a || !vals_a || !keys_b || !vals_b || !temps) {
    set_error("sort_pairs_u32_batch: 1..%d segments and non-null tables", MAX_BATCH);
    return 1;
  }
  SortBatch sb{};
  sb.V = n_segments;
  for (int i = 0; i < n_segments; ++i) {
    if (n[i] >= (1ull << 30)) { set_error("sort: n=%llu exceeds 2^30", (unsigned long long)n[i]); return 1; }
    sb.s[i] = SortSeg{keys_a[i], vals_a[i], keys_b[i], vals_b[i], static_cast<uint32_t*>(temps[i]), nullptr, (uint32_t)n[i]};
  }
  bool in_b = false;
  if (int rc = launch_sort_pairs_batch(sb, false, begin_bit, end_bit, s, &in_b)) return rc;
  if (result_in_b) *result_in_b = in_b ? 1 : 0;
  return check_launch("sort_pairs_u32_batch", false, s);
}

// ---- parity helpers -------------------------------------------------------------------------------
__global__ void export_keys_kernel(uint32_t R, const uint32_t* __restrict__ tile_keys, const uint32_t* __restrict__ ids,
                                   const float4* __restrict__ rgb_depth, uint64_t* __restrict__ keys_out,
                                   uint32_t* __restrict__ ids_out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  uint32_t id = ids[i];
  if (keys_out) keys_out[i] = ((uint64_t)tile_keys[i] << 32) | (uint64_t)__float_as_uint(rgb_depth[id].w);
  if (ids_out) ids_out[i] = id;
}

int tgr_export_binning(const tgr_params* p, uint64_t cap, uint64_t R, uint64_t* keys, uint32_t* ids, uint32_t* ranges,
                       void* stream) {
  if (R > cap) { set_error("export_binning: num_rendered %llu exceeds the capacity %llu", (unsigned long long)R, (unsigned long long)cap); return 1; }
  if (int rc = validate(p, true, cap)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GeomView g = carve_geom(p->geom_buffer, p->P);
  BinView b = carve_bin(p->binning_buffer, p->P, cap, p->W, p->H);
  ImageView im = carve_image(p->image_buffer, p->W, p->H);
  bool in_b = false;
  const uint32_t* vals = sorted_vals(p, b, &in_b);
  const uint32_t* tk = in_b ? b.key_b : b.key_a;
  if (R > 0 && (keys || ids)) {
    export_keys_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>((uint32_t)R, tk, vals, g.rgb_depth, keys, ids);
    count_launch();
  }
  if (ranges) {
    const uint32_t T = (uint32_t)((p->W + TILE - 1) / TILE) * ((p->H + TILE - 1) / TILE);
    cudaMemcpyAsync(ranges, im.ranges, (size_t)T * 8, cudaMemcpyDeviceToDevice, s);
  }
  return check_launch("export_binning", true, s);
}

__global__ void export_geom_kernel(int P, GeomView g, float* depth, float* xy, float* co, float* rgb, uint32_t* tiles) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const ushort4 rc = g.rect[i];
  const bool vis = (rc.z - rc.x) * (rc.w - rc.y) != 0;
  const float4 cd = g.rgb_depth[i];
  const float4 c = g.conic_opacity[i];
  const float4 m = g.xy_ext[i];
  if (depth) depth[i] = vis ? cd.w : 0.f;
  if (xy) { xy[2 * i] = vis ? m.x : 0.f; xy[2 * i + 1] = vis ? m.y : 0.f; }
  if (co) { co[4 * i] = vis ? c.x : 0.f; co[4 * i + 1] = vis ? c.y : 0.f; co[4 * i + 2] = vis ? c.z : 0.f; co[4 * i + 3] = vis ? c.w : 0.f; }
  if (rgb) { rgb[3 * i] = vis ? cd.x : 0.f; rgb[3 * i + 1] = vis ? cd.y : 0.f; rgb[3 * i + 2] = vis ? cd.z : 0.f; }
  if (tiles) tiles[i] = (uint32_t)(rc.z - rc.x) * (uint32_t)(rc.w - rc.y);
}

int tgr_export_geom(const tgr_params* p, float* depth, float* xy, float* conic_opacity, float* rgb,
                    uint32_t* tiles_touched, void* stream) {
  if (int rc = validate(p, false, 0)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->P == 0) return 0;
  GeomView g = carve_geom(p->geom_buffer, p->P);
  export_geom_kernel<<<(p->P + 255) / 256, 256, 0, s>>>(p->P, g, depth, xy, conic_opacity, rgb, tiles_touched);
  count_launch();
  return check_launch("export_geom", true, s);
}

int tgr_export_image_state(const tgr_params* p, float* final_T, uint32_t* n_contrib, void* stream) {
  if (int rc = validate(p, false, 0)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ImageView im = carve_image(p->image_buffer, p->W, p->H);
  const size_t N = (size_t)p->W * p->H;
  if (final_T) cudaMemcpyAsync(final_T, im.final_T, N * 4, cudaMemcpyDeviceToDevice, s);
  if (n_contrib) cudaMemcpyAsync(n_contrib, im.n_contrib, N * 4, cudaMemcpyDeviceToDevice, s);
  return check_launch("export_image_state", true, s);
}

}  // extern "C"
