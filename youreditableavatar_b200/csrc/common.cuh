// common.cuh — shared declarations for the sm_100a TetGS rasterizer kernels.
// Workspace layouts, small fp32 3x3 algebra and launch helpers.  No torch, no glm, no CUB.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tetgs_rast.h"

namespace tgr {

constexpr int TILE = TGR_TILE;          // 16x16 pixel tiles (reference config.h:16-17)
constexpr int TILE_PIX = TILE * TILE;   // 256
constexpr int NUM_SM = 148;             // B200

__host__ __device__ inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// Radix sort plan / temp layout (sort.cu)
// ---------------------------------------------------------------------------------------------
#ifndef RS_THREADS_
#define RS_THREADS_ 256
#endif
#ifndef RS_IPT_
#define RS_IPT_ 16
#endif
constexpr int RS_THREADS = RS_THREADS_;
constexpr int RS_IPT = RS_IPT_;
constexpr int RS_TILE = RS_THREADS * RS_IPT;  // 4096 pairs per CTA
constexpr int RS_BINS = 256;
constexpr int RS_MAX_PASSES = 4;
#ifndef RS_MIN_CTAS
#define RS_MIN_CTAS 3   // resident CTAs per SM the pass kernel is compiled for: 80 registers.  Measured with 4 (64
                        // registers, 52 B of spills, max shared-memory carve-out): depth sort 0.412 -> 0.436 ms, tile
                        // sort 0.316 -> 0.334 ms per 8-view batch — slower, so 3 stays
#endif

struct SortPlan {
  int npasses;
  int begin[RS_MAX_PASSES];
  int bits[RS_MAX_PASSES];
};

__host__ __device__ inline SortPlan make_sort_plan(int begin_bit, int end_bit) {
  SortPlan pl{};
  int nb = end_bit - begin_bit;
  if (nb <= 0) { pl.npasses = 0; return pl; }
  int np = (nb + 7) / 8;
  int per = (nb + np - 1) / np;
  pl.npasses = np;
  int b = begin_bit;
  for (int i = 0; i < np; ++i) {
    int w = (end_bit - b) < per ? (end_bit - b) : per;
    pl.begin[i] = b;
    pl.bits[i] = w;
    b += w;
  }
  return pl;
}

// temp: [ghist: RS_MAX_PASSES*256 u32][tickets: 32 u32][lookback: RS_MAX_PASSES * ntiles * 256 u32]
__host__ __device__ inline uint64_t sort_ntiles(uint64_t n) { return (n + RS_TILE - 1) / RS_TILE; }
// words of the temp area that must be zero before a sort of n pairs with `npasses` passes starts
__host__ __device__ inline uint64_t sort_zero_words(uint64_t n, int npasses) {
  return (uint64_t)RS_MAX_PASSES * RS_BINS + 32 + (uint64_t)npasses * sort_ntiles(n) * RS_BINS;
}
__host__ __device__ inline uint64_t sort_temp_bytes(uint64_t n) {
  uint64_t words = (uint64_t)RS_MAX_PASSES * RS_BINS + 32 + (uint64_t)RS_MAX_PASSES * sort_ntiles(n) * RS_BINS;
  return align_up(words * 4, 128);
}

// ---------------------------------------------------------------------------------------------
// Workspace layouts.  All sections 128-byte aligned.  The same carve is replayed in the backward,
// so sizes are pure functions of (P), (W,H), (P,R) — as in rasterizer_impl.cu:155-194.
// ---------------------------------------------------------------------------------------------
struct GeomHeader {        // 128 bytes
  uint32_t num_rendered;   // total (Gaussian,tile) instances = sum of tiles touched
  uint32_t overflow;       // set when num_rendered > capacity of the binning buffer
  uint32_t num_visible;    // Gaussians with radius > 0
  uint32_t prefilter_violation;  // set when tgr_params.prefiltered != 0 and a Gaussian failed the near-plane test: the
                                 // reference prints and __trap()s there (auxiliary.h:154-160); here the host raises
  uint32_t key_or;         // OR / AND over the depth keys of the visible Gaussians: the bits in which they differ are
  uint32_t key_and;        // all the depth sort has to look at (initialised to 0 / 0xffffffff by header_init_kernel)
  uint32_t pad[26];
};

struct GeomView {
  GeomHeader* header;
  uint32_t* depth_key;     // [P] fp32 bits of view-space depth (0x7fffffff when culled); sort buffer A keys
  uint32_t* order;         // [P] Gaussian ids in (depth, id) order; sort buffer A values
  uint32_t* key_alt;       // [P] sort buffer B keys
  uint32_t* val_alt;       // [P] sort buffer B values
  ushort4* rect;           // [P] tile rectangle (x0,y0,x1,y1); all zero when culled
  float4* xy_ext;          // [P] pixel-space mean (x,y) + conservative half-extents (hx,hy) of alpha >= 1/255
  float4* conic_opacity;   // [P] inverse 2D covariance (xx,xy,yy) + opacity
  float4* rgb_depth;       // [P] colour + view-space depth
  uint8_t* clamped;        // [P] bit c set when colour channel c was clamped at 0
  uint32_t* scan_state;    // [ceil(P/EMIT_TILE)+1] chained-scan state of the emit kernel
  uint32_t* sort_temp;     // sort_temp_bytes(P)
  uint64_t bytes;
};

constexpr int EMIT_THREADS = 256;
constexpr int EMIT_IPT = 4;
constexpr int EMIT_TILE = EMIT_THREADS * EMIT_IPT;

template <typename T>
__host__ __device__ inline T* carve(char*& p, uint64_t count) {
  uint64_t a = align_up((uint64_t)(uintptr_t)p, 128);
  T* r = reinterpret_cast<T*>(a);
  p = reinterpret_cast<char*>(a) + count * sizeof(T);
  return r;
}

__host__ __device__ inline GeomView carve_geom(void* base, int32_t P) {
  GeomView g;
  char* p = static_cast<char*>(base);
  uint64_t n = (uint64_t)(P > 0 ? P : 0);
  g.header = carve<GeomHeader>(p, 1);
  g.depth_key = carve<uint32_t>(p, n);
  g.order = carve<uint32_t>(p, n);
  g.key_alt = carve<uint32_t>(p, n);
  g.val_alt = carve<uint32_t>(p, n);
  g.rect = carve<ushort4>(p, n);
  g.xy_ext = carve<float4>(p, n);
  g.conic_opacity = carve<float4>(p, n);
  g.rgb_depth = carve<float4>(p, n);
  g.clamped = carve<uint8_t>(p, n);
  g.scan_state = carve<uint32_t>(p, (n + EMIT_TILE - 1) / EMIT_TILE + 1);
  g.sort_temp = carve<uint32_t>(p, sort_temp_bytes(n) / 4);
  g.bytes = (uint64_t)(p - static_cast<char*>(base)) + 128;
  return g;
}

constexpr int MAX_QUEUES = 1024;
constexpr int SEG = 512;  // list entries per backward work unit (a tile's list is replayed in independent segments)

struct ImageView {
  float* final_T;        // [N] transmittance left after blending
  uint32_t* n_contrib;   // [N] 1-based index in the tile list of the last blended Gaussian
  uint2* ranges;         // [T] [start,end) of each tile in the sorted instance list
  uint32_t* tile_last;   // [T] max n_contrib over the pixels of the tile (lets the backward skip the tail)
  uint32_t* order_fwd;   // [T] tile ids, heaviest (longest list) first: CTA i of blend_fwd renders tile order_fwd[i]
  uint32_t* order_bwd;   // [T] same for the backward, by tile_last
  uint32_t* queue_counters;  // [MAX_QUEUES] per-SM work-queue cursors of the blend kernels (zeroed by tile_order)
  float4* final_state;   // [N] (T_final, C_r, C_g, C_b) without background: end state of the forward recurrence
  float* final_z;        // [N] depth accumulator at the end of the forward (extras)
  uint32_t* seg_base;    // [T+1] first checkpoint slot of each tile (exclusive scan of ceil(len/SEG))
  uint32_t* unit_count;  // [32] [0] number of backward work units of this view, [1] ticket counter of blend_bwd's
                         //      persistent CTAs (view 0's serves the batch), [3] set once a backward has used the
                         //      packed gradient rows (a repeated backward has to clear them itself)
  uint64_t bytes;
};

__host__ __device__ inline ImageView carve_image(void* base, int32_t W, int32_t H) {
  ImageView v;
  char* p = static_cast<char*>(base);
  uint64_t N = (uint64_t)W * H;
  uint64_t T = (uint64_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
  v.final_T = carve<float>(p, N);
  v.n_contrib = carve<uint32_t>(p, N);
  v.ranges = carve<uint2>(p, T);
  v.tile_last = carve<uint32_t>(p, T);
  v.order_fwd = carve<uint32_t>(p, T);
  v.order_bwd = carve<uint32_t>(p, T);
  v.queue_counters = carve<uint32_t>(p, MAX_QUEUES);
  v.final_state = carve<float4>(p, N);
  v.final_z = carve<float>(p, N);
  v.seg_base = carve<uint32_t>(p, T + 1);
  v.unit_count = carve<uint32_t>(p, 32);
  v.bytes = (uint64_t)(p - static_cast<char*>(base)) + 128;
  return v;
}

struct BinView {
  uint32_t* key_a;   // [R] tile id per instance (sort buffer A)
  uint32_t* val_a;   // [R] Gaussian id per instance
  uint32_t* key_b;   // [R] sort buffer B
  uint32_t* val_b;
  uint32_t* sort_temp;
  float* grad_acc;   // [P*GRAD_ACC] backward accumulators (see blend_bwd.cu); lives here so it is
                     // allocated with the call that needs it and freed with the autograd node
  float4* ckpt;      // [units_cap*256] forward state (T, C) of every pixel at every SEG-th list position
  float* ckpt_z;     // [units_cap*256] same for the depth accumulator (extras)
  uint2* units;      // [units_cap] backward work units (tile, segment)
  uint64_t units_cap;
  uint64_t bytes;
};

constexpr int GRAD_ACC = 12;  // mean2D.xy, conic.xyz, opacity, rgb, depth, 2 pad

// ---------------------------------------------------------------------------------------------
// Multi-view batches: the per-Gaussian kernels (preprocess forward / backward) visit every Gaussian ONCE and
// loop over the views of the batch, so the 236 B of parameters per Gaussian (SH above all) are read once per
// batch instead of once per view, and the gradients are written once per batch instead of read-modify-written
// once per view.  A single-view call is a batch of one.
// ---------------------------------------------------------------------------------------------
constexpr int MAX_BATCH = TGR_MAX_BATCH;

struct ViewDesc {
  const float* viewmatrix;
  const float* projmatrix;
  const float* campos;
  float tan_fovx, tan_fovy;
  int32_t W, H;
  int32_t prefiltered;
  uint32_t sort_zero_words;   // leading words of sort_temp the preprocess kernel clears for the depth sort
  uint32_t* sort_temp;
  // forward outputs of this view
  int32_t* radii;
  GeomHeader* header;
  uint32_t* depth_key;
  ushort4* rect;
  float4* xy_ext;
  float4* conic_opacity;
  float4* rgb_depth;
  uint8_t* clamped;
  // backward input of this view (packed 2-D gradient rows filled by blend_bwd)
  const float* grad_acc;
};

struct ViewBatch {
  int32_t V;
  int32_t first;   // backward: first Gaussian of the range this launch covers (multiple of 256); forward: 0
  int32_t end;     // backward: one past the last Gaussian of the range; forward: P
  int32_t pad;
  ViewDesc v[MAX_BATCH];
};

// Everything the binning and blend kernels need to know about one view of a batch.  Those kernels take the whole
// table as a __grid_constant__ parameter and pick their view from blockIdx.y (or from the work item), so ONE launch
// serves every view of the batch: the per-view launches of a 1 M-Gaussian view are individually too small to fill
// 148 SMs (sorts, scans) or end in a long single-tile tail (blending).
struct RenderView {
  int32_t P, W, H;
  uint32_t T;               // tiles of this view
  uint32_t cap;             // instance capacity of the binning buffer
  uint32_t units_cap;
  // geometry records
  GeomHeader* header;
  const uint32_t* order;    // Gaussian ids in (depth, id) order
  const ushort4* rect;
  const float4* xy_ext;
  const float4* conic_opacity;
  const float4* rgb_depth;
  uint32_t* scan_state;
  // binning
  uint32_t* key_a;
  uint32_t* val_a;
  uint32_t* tile_sort_temp;      // temp area of the tile sort; its leading tile_sort_zero_words are cleared by emit_count
  uint32_t tile_sort_zero_words;
  const uint32_t* sorted_keys;   // tile ids after the tile sort
  const uint32_t* point_list;    // Gaussian ids after the tile sort
  float* grad_acc;
  float4* ckpt;
  float* ckpt_z;
  uint2* units;
  // image state
  uint2* ranges;
  uint32_t* tile_last;
  uint32_t* order_fwd;
  uint32_t* seg_base;
  uint32_t* unit_count;
  float* final_T;
  uint32_t* n_contrib;
  float4* final_state;
  float* final_z;
  // user tensors
  const float* bg;
  float* out_color;
  float* out_depth;
  float* out_alpha;
  const float* dL_dpix;
  const float* dL_ddepth;
  const float* dL_dalpha;
};
struct RenderBatch {
  int32_t V;
  uint32_t T_max;            // largest tile count of the batch
  uint32_t* queue_counters;  // [MAX_QUEUES] per-SM work-queue cursors shared by the batch (zeroed by tile_order)
  RenderView v[MAX_BATCH];
};

// A batch of independent radix sorts (sort.cu): one segment per view.
struct SortSeg {
  uint32_t* keys_a;
  uint32_t* vals_a;
  uint32_t* keys_b;
  uint32_t* vals_b;
  uint32_t* temp;          // sort_temp_bytes(n_host)
  const uint32_t* n_dev;   // optional device-side element count (min'ed with n_host)
  uint32_t n_host;
};
struct SortBatch {
  int32_t V;
  SortSeg s[MAX_BATCH];
};

// camera block staged in shared memory by the batched kernels: [view 16 | proj 16 | campos 3 | pad] per view
constexpr int CAM_FLOATS = 36;

// upper bound of sum_t ceil(len_t / SEG): every non-empty tile adds at most one partial segment
__host__ __device__ inline uint64_t units_capacity(uint64_t R, int32_t W, int32_t H) {
  const uint64_t T = (uint64_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
  return R / SEG + (T < R ? T : R) + 1;
}

__host__ inline BinView carve_bin(void* base, int32_t P, uint64_t R, int32_t W, int32_t H) {
  BinView b;
  char* p = static_cast<char*>(base);
  b.key_a = carve<uint32_t>(p, R);
  b.val_a = carve<uint32_t>(p, R);
  b.key_b = carve<uint32_t>(p, R);
  b.val_b = carve<uint32_t>(p, R);
  b.sort_temp = carve<uint32_t>(p, sort_temp_bytes(R) / 4);
  b.grad_acc = carve<float>(p, (uint64_t)(P > 0 ? P : 0) * GRAD_ACC);
  b.units_cap = units_capacity(R, W, H);
  b.ckpt = carve<float4>(p, b.units_cap * TILE_PIX);
  b.ckpt_z = carve<float>(p, b.units_cap * TILE_PIX);
  b.units = carve<uint2>(p, b.units_cap);
  b.bytes = (uint64_t)(p - static_cast<char*>(base)) + 128;
  return b;
}

// number of bits needed to represent tile ids 0..T-1
__host__ __device__ inline int tile_bits(uint32_t T) {
  int b = 0;
  while ((1u << b) < T && b < 31) ++b;
  return b < 1 ? 1 : b;
}

// ---------------------------------------------------------------------------------------------
// 3x3 column-major algebra with exactly the scalar expression shapes glm::mat3 expands to
// (third_party/glm/glm/detail/type_mat3x3.inl:486-519), so that nvcc's FMA contraction — and therefore
// every fp32 rounding — matches the reference build.  m[c][r]: column c, row r.
// ---------------------------------------------------------------------------------------------
struct M3 {
  float m[3][3];
};

__device__ __forceinline__ M3 m3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1,
                                 float c2) {
  M3 r;
  r.m[0][0] = a0; r.m[0][1] = a1; r.m[0][2] = a2;
  r.m[1][0] = b0; r.m[1][1] = b1; r.m[1][2] = b2;
  r.m[2][0] = c0; r.m[2][1] = c1; r.m[2][2] = c2;
  return r;
}

__device__ __forceinline__ M3 m3_mul(const M3& A, const M3& B) {
  M3 R;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r)
      R.m[c][r] = A.m[0][r] * B.m[c][0] + A.m[1][r] * B.m[c][1] + A.m[2][r] * B.m[c][2];
  return R;
}

__device__ __forceinline__ M3 m3_T(const M3& A) {
  M3 R;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) R.m[c][r] = A.m[r][c];
  return R;
}

// p' = M p (+ translation) for the column-major 4x4 camera matrices (auxiliary.h:58-77)
__device__ __forceinline__ float3 xform4x3(const float3& p, const float* __restrict__ m) {
  float3 t = {
      m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
      m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
      m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
  };
  return t;
}
__device__ __forceinline__ float4 xform4x4(const float3& p, const float* __restrict__ m) {
  float4 t = {
      m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
      m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
      m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
      m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15],
  };
  return t;
}

// NDC -> pixel; evaluated in double exactly like auxiliary.h:41-44 (its literals are doubles)
__device__ __forceinline__ float ndc2pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// Tile rectangle touched by a splat of integer radius r centred at p (auxiliary.h:46-56)
__device__ __forceinline__ void tile_rect(const float2 p, int r, uint32_t gx, uint32_t gy, uint2& rmin, uint2& rmax) {
  rmin = {min(gx, (uint32_t)max((int)0, (int)((p.x - r) / TILE))), min(gy, (uint32_t)max((int)0, (int)((p.y - r) / TILE)))};
  rmax = {min(gx, (uint32_t)max((int)0, (int)((p.x + r + TILE - 1) / TILE))),
          min(gy, (uint32_t)max((int)0, (int)((p.y + r + TILE - 1) / TILE)))};
}

// SH basis constants (auxiliary.h:22-39)
__device__ constexpr float SH_C0 = 0.28209479177387814f;
__device__ constexpr float SH_C1 = 0.4886025119029199f;
__device__ constexpr float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                       -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                       0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                       -0.5900435899266435f};

// ---------------------------------------------------------------------------------------------
// host-side plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
// per-stage event timing (c_abi.cu); no-ops unless tgr_profile_enable(1)
void prof_begin(int stage, cudaStream_t s);
void prof_end(int stage, cudaStream_t s);
int check_launch(const char* what, bool debug, cudaStream_t s);
void count_launch(int n = 1);  // bookkeeping for tgr_kernel_launches()

// stage launchers (each in its own .cu)
ViewDesc make_view_desc(const tgr_params& p, const GeomView& g, const float* grad_acc);
int launch_preprocess(const tgr_params& p, const tgr_binding* bind, const ViewBatch& vb, cudaStream_t s);
int launch_sort_pairs(uint64_t n_host, const uint32_t* n_dev, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b,
                      uint32_t* vals_b, bool iota_vals, int begin_bit, int end_bit, uint32_t* temp, cudaStream_t s,
                      bool* result_in_b);
int launch_sort_pairs_batch(const SortBatch& sb, bool iota_vals, int begin_bit, int end_bit, cudaStream_t s,
                            bool* result_in_b, bool temp_is_zero = false);
int launch_emit(const RenderBatch& rb, cudaStream_t s);
int launch_ranges(const RenderBatch& rb, bool zero_first, cudaStream_t s);
int launch_tile_order(const RenderBatch& rb, cudaStream_t s);
int launch_unit_build(const RenderBatch& rb, cudaStream_t s);
uint32_t num_queues();  // number of SMs of the current device (one work queue per SM)
int launch_blend_fwd(const RenderBatch& rb, bool extras, bool debug, cudaStream_t s);
int launch_blend_bwd(const RenderBatch& rb, bool extras, bool debug, cudaStream_t s);
int launch_preprocess_bwd(const tgr_params& p, const tgr_binding* bind, const ViewBatch& vb, cudaStream_t s);
int launch_mark_visible(int32_t P, const float* means3D, const float* view, const float* proj, uint8_t* present,
                        cudaStream_t s);

}  // namespace tgr
