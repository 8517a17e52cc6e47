/*
 * tetgs_rast.h — C ABI of the B200-native (sm_100a) differentiable Gaussian rasterizer for TetGS.
 *
 * This is the drop-in boundary for the hot path named in BASELINE.json: every entry point takes
 * plain pointers / sizes / a cudaStream_t (as void*), returns an int status (0 = ok) and never
 * touches torch types.  The reference interfaces each function replaces are cited as
 * file:line relative to /root/reference/Edit_core/thirdparties/.
 *
 * Conventions (same as the reference):
 *   - all tensors are fp32, contiguous, device memory unless stated;
 *   - viewmatrix / projmatrix are the TRANSPOSED 4x4 matrices (flat memory is column-major,
 *     element (row r, col c) at m[4*c + r]; diff-gaussian-rasterization/cuda_rasterizer/auxiliary.h:58-77);
 *   - quaternions are (r,x,y,z) and are NOT normalised by the kernels (forward.cu:127);
 *   - cov3D is packed upper-triangular (xx,xy,xz,yy,yz,zz);
 *   - an absent optional input is a NULL pointer;
 *   - the three workspace buffers (geom / binning / image) are owned by the caller, opaque, and
 *     must be handed back unchanged to tgr_backward (rasterizer_impl.cu:371-373 does the same).
 */
#ifndef TETGS_RAST_H_
#define TETGS_RAST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGR_ABI_VERSION 4
#define TGR_TILE 16 /* 16x16 pixel tiles, config.h:16-17 */
#define TGR_MAX_BATCH 8 /* views served by one launch of the per-Gaussian kernels (longer batches are chunked) */

/* Per-call description of one view and one set of Gaussians.  POD only. */
typedef struct tgr_params {
  /* sizes */
  int32_t P;            /* number of Gaussians */
  int32_t D;            /* active SH degree (0..3) */
  int32_t M;            /* SH coefficients per channel stored in `shs` (0 when colours are precomputed) */
  int32_t W, H;         /* image width / height */
  float tan_fovx, tan_fovy;
  float scale_modifier;
  int32_t prefiltered;  /* 1: the caller asserts that every Gaussian passes the near-plane test (it ran mark_visible).
                           A Gaussian that fails it anyway is skipped and flagged: word [3] of host_num_rendered /
                           tgr_read_header becomes 1 and the Python operator raises the reference's message — the
                           reference prints it and __trap()s the context (auxiliary.h:154-160) */
  int32_t debug;        /* 1: synchronise + report CUDA errors after every stage (auxiliary.h:166-173) */
  int32_t extras;       /* 1: also produce depth/alpha images (new) */
  int32_t accumulate;   /* backward: 1 = add into the output gradient tensors instead of overwriting them
                           (multi-view gradient accumulation without an extra pass; new) */
  int32_t depth_key_bits; /* forward: number of low bits of the fp32 depth keys the depth sort has to look at, 0 = all 32.
                           The visible depths of an avatar share sign, exponent and often leading mantissa bits; words
                           [4], [5] of the header mirror (OR / AND over the visible keys) tell how many bits differ, and a
                           caller that renders the same scene again passes that (+ slack) to save a radix pass.  Too few
                           bits give a wrong order, never a fault: check the mirror after the call (the Python operator
                           does and re-renders) */
  int32_t reserved_;
  /* camera (device pointers) */
  const float* background; /* [3] */
  const float* viewmatrix; /* [16] */
  const float* projmatrix; /* [16] */
  const float* campos;     /* [3] */
  /* Gaussians (device pointers) */
  const float* means3D;        /* [P,3] */
  const float* shs;            /* [P,M,3] or NULL */
  const float* colors_precomp; /* [P,3]   or NULL */
  const float* opacities;      /* [P,1] */
  const float* scales;         /* [P,3]   or NULL */
  const float* rotations;      /* [P,4]   or NULL */
  const float* cov3D_precomp;  /* [P,6]   or NULL */
  /* workspaces (device pointers, sizes from tgr_*_bytes) */
  void* geom_buffer;
  void* binning_buffer;
  void* image_buffer;
  uint64_t geom_bytes, binning_bytes, image_bytes;
  /* forward outputs */
  float* out_color;   /* [3,H,W] */
  int32_t* radii;     /* [P] */
  float* out_depth;   /* [H,W] or NULL (extras) */
  float* out_alpha;   /* [H,W] or NULL (extras) */
  /* backward inputs */
  const float* dL_dout_color; /* [3,H,W] */
  const float* dL_dout_depth; /* [H,W] or NULL */
  const float* dL_dout_alpha; /* [H,W] or NULL */
  /* backward outputs: every element is written by the kernels, no pre-zeroing needed */
  float* dL_dmeans2D;   /* [P,3] */
  float* dL_dcolors;    /* [P,3] */
  float* dL_dopacity;   /* [P,1] */
  float* dL_dmeans3D;   /* [P,3] */
  float* dL_dcov3D;     /* [P,6] */
  float* dL_dsh;        /* [P,M,3] or NULL when M == 0 */
  float* dL_dscales;    /* [P,3] or NULL */
  float* dL_drotations; /* [P,4] or NULL */
  /* pinned host mirror of the geom header, EIGHT words {num_rendered, overflow, num_visible, prefilter_violation,
   * OR of the visible depth keys, AND of the visible depth keys, 0, 0}, written asynchronously by
   * tgr_forward_preprocess (may be NULL) */
  uint32_t* host_num_rendered;
} tgr_params;

/* Optional mesh binding fused into the preprocess kernels (replaces the eager PyTorch binding in
 * Edit_core/tetgs_scene/tetgs_model.py:252-286 — points = ori + n*delta, exp/sigmoid/normalize
 * activations).  When passed (non-NULL) to the *_bound entry points, means3D/opacities/scales/rotations
 * in tgr_params are ignored and derived from these raw parameters instead. */
typedef struct tgr_binding {
  int32_t n_verts, n_faces;
  const float* verts;          /* [n_verts,3] mesh (tet-surface) vertices */
  const float* vert_normals;   /* [n_verts,3] unit vertex normals */
  const int32_t* faces;        /* [n_faces,3] */
  const int32_t* face_index;   /* [P] face each Gaussian is bound to */
  const float* bary;           /* [P,3] barycentric coordinates */
  const float* delta;          /* [P]  learnable offset along the interpolated normal */
  const float* log_scales;     /* [P,3] pre-activation scales   (exp) */
  const float* raw_quats;      /* [P,4] pre-normalisation quats (normalize) */
  const float* opacity_logits; /* [P]   pre-activation opacity  (sigmoid) */
  /* activated values written by the forward (needed by the drop-in callers and the backward) */
  float* out_means3D;          /* [P,3] */
  float* out_scales;           /* [P,3] */
  float* out_rotations;        /* [P,4] */
  float* out_opacities;        /* [P]   */
  /* gradients wrt the raw parameters (backward outputs, fully written) */
  float* dL_ddelta;            /* [P]   */
  float* dL_dlog_scales;       /* [P,3] */
  float* dL_draw_quats;        /* [P,4] */
  float* dL_dopacity_logits;   /* [P]   */
  float* dL_dverts;            /* [n_verts,3] or NULL; accumulated with atomics, caller zeroes */
  /* Direct form (origins != NULL; the mesh fields above are then ignored): mean = origins + normals * delta with
   * per-Gaussian constants, which is how every scene model of the reference holds its binding between re-meshings —
   * TetGS.ori_points / .normals (tetgs_model.py:156-172, points = ori_points + normals * _points, :252-258), the edit
   * models' cat(keep_points, ori_edit_points + _edit_normals * _edit_points) (tetgs_edit_3d.py:272-280) and the fixed
   * points of the flat 2-D Gaussians (tetgs_edit_2d.py:279-282: normals == NULL).  delta may be NULL (zero offsets). */
  const float* origins;        /* [P,3] or NULL */
  const float* normals;        /* [P,3] or NULL */
  /* Gaussians [0, n_frozen) are frozen — the `keep_*` part of the edit models, requires_grad=False
   * (tetgs_edit_2d.py:237-267, tetgs_edit_3d.py:160-200): the backward writes zero gradient rows for them without
   * computing anything. */
  int32_t n_frozen;
  int32_t reserved_;
} tgr_binding;

/* ---- library ---- */
int tgr_abi_version(void);
const char* tgr_last_error(void);     /* thread-local message of the last non-zero status */

/* ---- workspace sizes (pure functions of the arguments; rasterizer_impl.h:66-72 `required<T>`) ---- */
uint64_t tgr_geom_bytes(int32_t P);
uint64_t tgr_image_bytes(int32_t W, int32_t H);
uint64_t tgr_binning_bytes(int32_t P, uint64_t num_rendered_capacity, int32_t W, int32_t H);
/* Inverse of tgr_binning_bytes: the instance capacity a binning buffer of `binning_bytes` bytes was sized for (the
 * largest capacity whose layout fits; every capacity that rounds to the same size has the same layout).  Lets the
 * backward of a forward that was launched from a capacity hint (buffer larger than num_rendered) recover the layout from
 * the buffer alone, the way rasterizer_impl.cu:371-373 recovers it from R. */
uint64_t tgr_binning_capacity(int32_t P, uint64_t binning_bytes, int32_t W, int32_t H);

/* ---- forward: replaces CudaRasterizer::Rasterizer::forward (rasterizer_impl.cu:198-336) ----
 * Stage 1: per-Gaussian preprocess (forward.cu:155-256), depth ordering, total instance count.
 *          Needs geom_buffer (and out radii).  The instance count R lands in the geom header on the
 *          device and, if host_num_rendered != NULL, is copied there asynchronously.
 * Stage 2: key emission + tile sort + tile ranges + alpha blending
 *          (rasterizer_impl.cu:70-138, 289-333; forward.cu:261-374).  `num_rendered_capacity` is the
 *          number of instances binning_buffer was sized for.  If the true R exceeds it the call
 *          renders nothing, sets the overflow flag (tgr_read_header) and returns 0; the caller retries.
 */
int tgr_forward_preprocess(const tgr_params* p, const tgr_binding* bind, void* stream);
int tgr_forward_render(const tgr_params* p, uint64_t num_rendered_capacity, void* stream);
/* Blocks the host until the asynchronous host_num_rendered copy of the calling thread's most recent
 * tgr_forward_preprocess has landed (the GPU keeps running the depth sort meanwhile). This is the only
 * host<->device synchronisation of a forward; the reference has the same one at rasterizer_impl.cu:281. */
int tgr_wait_num_rendered(void);

/* ---- backward: replaces CudaRasterizer::Rasterizer::backward (rasterizer_impl.cu:340-434) ---- */
int tgr_backward(const tgr_params* p, const tgr_binding* bind, uint64_t num_rendered_capacity, void* stream);

/* ---- multi-view batches (new; the reference renders one view per call, tetgs_texture/refine.py:54) ----
 * `views` is an array of n_views tgr_params that describe the SAME Gaussians (identical P, D, M, Gaussian
 * tensors) seen from different cameras, each with its own workspaces and outputs.  The per-Gaussian
 * kernels visit every Gaussian once per batch: parameters (236 B per Gaussian with degree-3 SH) are read once
 * instead of once per view, and the parameter gradients are written once instead of accumulated per view.
 * The per-view stages in between are independent and may be issued on different streams:
 *
 *   tgr_forward_preprocess_batch(views, n, bind, s0)       one launch per <= TGR_MAX_BATCH views; every view's
 *                                                          instance count goes to its host_num_rendered slot
 *   tgr_wait_num_rendered()                                (size the binning buffers)
 *   for v: tgr_forward_depth_sort(&views[v], s_v); tgr_forward_render(&views[v], cap_v, s_v)
 *   ... upstream gradients ...
 *   for v: tgr_backward_blend(&views[v], cap_v, s_v)       2-D gradients of view v (blend stage only)
 *   tgr_backward_preprocess_batch(views, caps, n, bind, 0, 0, s0)
 *                                                          after all s_v joined s0: chain rule to the
 *                                                          parameters, summed over the batch; outputs and the
 *                                                          accumulate flag are taken from views[0]
 * tgr_forward_preprocess == preprocess_batch of one view + depth sort; tgr_backward == blend + batch of one. */
int tgr_forward_preprocess_batch(const tgr_params* views, int32_t n_views, const tgr_binding* bind, void* stream);
int tgr_forward_depth_sort(const tgr_params* p, void* stream);
int tgr_backward_blend(const tgr_params* p, uint64_t num_rendered_capacity, void* stream);
/* Fused schedule: every per-view stage (depth sort, emission, tile sort, ranges, blending / blend backward) is
 * ONE launch for up to TGR_MAX_BATCH views (blockIdx.y or the work queue selects the view), on one stream:
 *   tgr_forward_render_batch  == for v: tgr_forward_depth_sort + tgr_forward_render
 *   tgr_backward_blend_batch  == for v: tgr_backward_blend
 * with identical per-view results.  A single 1 M-Gaussian view cannot fill 148 SMs in its sorts and scans and ends
 * its blending in a single-tile tail; eight views in one launch can, and the heaviest tiles of ALL views are
 * scheduled first. */
int tgr_forward_render_batch(const tgr_params* views, const uint64_t* num_rendered_capacities, int32_t n_views,
                             void* stream);
int tgr_backward_blend_batch(const tgr_params* views, const uint64_t* num_rendered_capacities, int32_t n_views,
                             void* stream);
/* gaussian_first / gaussian_count restrict the launch to a range of Gaussians (first a multiple of 256;
 * count <= 0 = all): a data-parallel caller splits the backward into ranges and all-reduces the gradients of one
 * range while the next range is being computed. */
int tgr_backward_preprocess_batch(const tgr_params* views, const uint64_t* num_rendered_capacities, int32_t n_views,
                                  const tgr_binding* bind, int32_t gaussian_first, int32_t gaussian_count, void* stream);

/* Synchronously reads {num_rendered, overflow, num_visible, prefilter_violation} from a geom buffer header. */
int tgr_read_header(const void* geom_buffer, uint32_t out[4], void* stream);

/* ---- mark_visible: replaces Rasterizer::markVisible (rasterizer_impl.cu:141-153) ---- */
int tgr_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* ---- distCUDA2: replaces SimpleKNN::knn (simple-knn/simple_knn.cu:185-221) ----
 * mean squared distance to the 3 nearest other points; `workspace` of tgr_knn_bytes(P) bytes. */
uint64_t tgr_knn_bytes(int32_t P);
int tgr_dist2(int32_t P, const float* points, float* mean_dist2, void* workspace, uint64_t workspace_bytes,
              void* stream);

/* ---- gradient exchange over NVSwitch (new; the reference is single-GPU) ----
 * Two-shot all-reduce (sum, fp32) of a buffer that lives in symmetric memory with a multicast mapping: rank r
 * reduces slice r of the buffer inside the switch (multimem.ld_reduce) and broadcasts the sum to every GPU
 * (multimem.st).  `multicast_ptr` is the multicast address of the buffer's first element, n_floats a multiple
 * of 4.  The caller brackets the call with cross-rank barriers on the same stream. */
int tgr_multimem_allreduce_f32(void* multicast_ptr, uint64_t n_floats, int32_t rank, int32_t world, void* stream);
/* Same, with the grid capped at max_ctas CTAs (0 = no cap): for slices exchanged on a side stream while the backward
 * still computes the next range of Gaussians (tgr_backward_preprocess_batch with gaussian_first / gaussian_count). */
int tgr_multimem_allreduce_f32_capped(void* multicast_ptr, uint64_t n_floats, int32_t rank, int32_t world,
                                      int32_t max_ctas, void* stream);

/* ---- per-stage device timing (CUDA events recorded on the launching stream) ----
 * tgr_profile_enable(1) makes every subsequent stage launch bracket itself with events;
 * tgr_profile_collect synchronises the recorded events and returns, per stage id (TGR_STAGE_*), the summed
 * milliseconds and the number of launches since the last collect.  Arrays must hold TGR_NUM_STAGES entries. */
#define TGR_STAGE_PREPROCESS 0
#define TGR_STAGE_DEPTH_SORT 1
#define TGR_STAGE_EMIT 2
#define TGR_STAGE_TILE_SORT 3
#define TGR_STAGE_RANGES 4
#define TGR_STAGE_BLEND_FWD 5
#define TGR_STAGE_BLEND_BWD 6
#define TGR_STAGE_PREPROCESS_BWD 7
#define TGR_NUM_STAGES 8
/* Running count of CUDA kernels this library has launched in the calling process (memsets excluded). */
uint64_t tgr_kernel_launches(void);
int tgr_profile_enable(int on);
int tgr_profile_collect(float* sum_ms, int32_t* launches);

/* ---- parity / debugging helpers (used by tests; not on the hot path) ----
 * Reconstructs the reference's sorted 64-bit keys (tile << 32 | depth bits, rasterizer_impl.cu:102-104),
 * the sorted Gaussian ids and the per-tile ranges from the opaque buffers. keys/ids have num_rendered entries,
 * ranges 2*T entries; num_rendered_capacity is what the binning buffer was sized for (>= num_rendered).
 * Any output pointer may be NULL. */
int tgr_export_binning(const tgr_params* p, uint64_t num_rendered_capacity, uint64_t num_rendered, uint64_t* keys,
                       uint32_t* ids, uint32_t* ranges, void* stream);
/* Copies per-Gaussian preprocess records out of the geom buffer: depth[P], xy[2P], conic_opacity[4P],
 * rgb[3P], tiles_touched[P].  Any output pointer may be NULL. */
int tgr_export_geom(const tgr_params* p, float* depth, float* xy, float* conic_opacity, float* rgb,
                    uint32_t* tiles_touched, void* stream);
/* Copies per-pixel blend state: final_T[H*W], n_contrib[H*W]. */
int tgr_export_image_state(const tgr_params* p, float* final_T, uint32_t* n_contrib, void* stream);

/* Stand-alone stable LSD radix sort of (u32 key, u32 value) pairs on bits [begin_bit, end_bit) — the
 * hand-written onesweep that replaces cub::DeviceRadixSort::SortPairs (rasterizer_impl.cu:303-308,
 * simple_knn.cu:207-213).  Exposed for tests and micro-benchmarks. Result is in keys_out/vals_out. */
uint64_t tgr_sort_temp_bytes(uint64_t n);
/* Batched form — what the rasterizer itself uses for the V views of a batch: up to TGR_MAX_BATCH independent sorts, the
 * same bit range for all, every radix pass ONE launch (blockIdx.y = segment).  Segment i sorts n[i] pairs in place
 * between its a / b buffers; *result_in_b tells where the results are (1: b buffers, 0: a buffers). */
int tgr_sort_pairs_u32_batch(int32_t n_segments, const uint64_t* n, uint32_t* const* keys_a, uint32_t* const* vals_a,
                             uint32_t* const* keys_b, uint32_t* const* vals_b, int begin_bit, int end_bit,
                             void* const* temps, int32_t* result_in_b, void* stream);
int tgr_sort_pairs_u32(uint64_t n, uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                       int begin_bit, int end_bit, void* temp, uint64_t temp_bytes, void* stream);

/* ======================= rows "next" of the scope table ======================= */

/* ---- f1: image loss, forward + gradient in two launches ----
 * Replaces Edit_core/utils/loss_utils.py:17-63 (l1_loss, l2_loss, ssim / _ssim: five grouped 11x11 conv2d +
 * ~20 elementwise ATen kernels, and as many again in autograd's backward) and the loss closure built at
 * Edit_core/tetgs_texture/refine.py:241-247 (same code: paint_2dgs.py:341-347, refine_3dgs.py:273-279):
 *
 *   loss_v = l1_weight * mean|p_v - g_v| + l2_weight * mean (p_v - g_v)^2
 *          + dssim_weight * (1 - mean ssim_map(p_v, g_v))                      per view v, means over [3,H,W]
 *   total  = sum_v w_v * loss_v          w_v = view_weights[v], or 1/n_views when view_weights == NULL
 *                                        (the reference's .mean() over its image batch, loss_utils.py:18,61)
 *
 * ssim_map exactly as loss_utils.py:44-59: Gaussian window 11, sigma 1.5 (fp32 weights normalised in fp32,
 * loss_utils.py:23-25), zero padding 5, C1 = 0.01^2, C2 = 0.03^2, variances as E[x^2] - mu^2.
 * `pred` is [V,3,H,W] fp32 (the rasterizer's colour output); `target` is [V,3,H,W] fp32, or uint8 with
 * target_is_u8 != 0 (converted as u/255, general_utils.py:8: the training images are 8-bit).
 * loss_out: [1 + V] floats = {total, loss_0 .. loss_{V-1}}, written on the device (no host sync).
 * workspace: tgr_image_loss_bytes(V, W, H) bytes, 16-byte aligned; the forward leaves there what the backward
 * needs (three fp32 maps per pixel and channel), so both calls get the same workspace, pred and target.
 * backward: dL_dpred [V,3,H,W] = sum_v dL_dloss_view[v] * d loss_v / d pred, fully written; dL_dloss_view is a
 * DEVICE array [V] (what autograd hands down: grad_total * w_v + grad_loss_v), NULL = 1/n_views.
 * tgr_image_loss = forward + (dL_dpred != NULL) backward with dL_dloss_view = view_weights: the training step. */
uint64_t tgr_image_loss_bytes(int32_t n_views, int32_t W, int32_t H);
int tgr_image_loss_forward(int32_t n_views, int32_t W, int32_t H, const float* pred, const void* target,
                           int32_t target_is_u8, const float* view_weights, float l1_weight, float l2_weight,
                           float dssim_weight, float* loss_out, void* workspace, uint64_t workspace_bytes,
                           void* stream);
int tgr_image_loss_backward(int32_t n_views, int32_t W, int32_t H, const float* pred, const void* target,
                            int32_t target_is_u8, const float* dL_dloss_view, float l1_weight, float l2_weight,
                            float dssim_weight, float* dL_dpred, void* workspace, uint64_t workspace_bytes,
                            void* stream);
int tgr_image_loss(int32_t n_views, int32_t W, int32_t H, const float* pred, const void* target,
                   int32_t target_is_u8, const float* view_weights, float l1_weight, float l2_weight,
                   float dssim_weight, float* loss_out, float* dL_dpred, void* workspace,
                   uint64_t workspace_bytes, void* stream);

/* ---- f2: Adam over all parameter groups in ONE launch ----
 * Replaces torch.optim.Adam(l, lr=0.0, eps=1e-15).step() as driven by TetGSOptimizer / EditTetGSOptimizer
 * (Edit_core/tetgs_scene/tetgs_optimizer.py:66-104, 136-170): betas (0.9, 0.999) defaults, no weight decay,
 * no amsgrad, per-group learning rates (points: exponential schedule evaluated on the host,
 * utils/general_utils.py:25-58; SH dc: feature_lr, SH rest: feature_lr / 20).
 *
 *   m <- m + (g - m)(1 - beta1);  v <- beta2 v + (1 - beta2) g^2;
 *   p <- p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)        g = grad_scale * grad
 *
 * A group is a contiguous run of `count` floats in four parallel arrays.  The rasterizer keeps SH as [P,M,3] rows
 * (dc = the first 3 floats of each 3M-float row, rest = the others) where the reference holds two tensors with
 * two learning rates: element i of a group uses `lr` when (i % period) < split and `lr_alt` otherwise
 * (period == 0: always `lr`).  grad may be the flat all-reduced gradient buffer itself (no copy).
 * lr == 0 (and lr_alt == 0) still updates m and v, as torch does. */
#define TGR_ADAM_MAX_GROUPS 16
typedef struct tgr_adam_group {
  float* param;           /* [count] */
  const float* grad;      /* [count] */
  float* exp_avg;         /* [count] */
  float* exp_avg_sq;      /* [count] */
  uint64_t count;
  double lr, lr_alt;      /* doubles, like the Python floats torch divides by the bias correction */
  uint32_t period, split;
} tgr_adam_group;
int tgr_adam_step(const tgr_adam_group* groups, int32_t n_groups, int32_t step /* t >= 1 */, double beta1,
                  double beta2, double eps, float grad_scale, void* stream);

/* ---- f3: device-side camera / settings construction for a batch of views ----
 * Replaces the per-call preamble of render_image_gaussian_rasterizer (Edit_core/tetgs_scene/
 * tetgs_model.py:479-503: cat + axis flip + torch.inverse + getWorld2View + getProjectionMatrix + bmm, about
 * 25 small ATen launches, two .item() syncs and an H2D copy per view; the edit variants add a
 * device->numpy->device round trip and a numpy inverse, tetgs_edit_3d.py:508-518) and
 * utils/graphics_utils.py:39-49,68-86.  One launch, no host synchronisation, for V views:
 *   c2w_v   = [camera_to_worlds[v]; 0 0 0 1] with columns 1 and 2 negated   (OpenGL -> COLMAP axes, :482-485)
 *   w2c_v   = inverse(c2w_v)                                                 (:488; affine inverse, fp32)
 *   viewmatrix = w2c_v^T            (flat memory: element (r,c) of w2c at [4c + r], what the kernels read)
 *   P          = getProjectionMatrix(znear, zfar, fovx, fovy), P[0,2] = -cx, P[1,2] = -cy   (:493-499)
 *   projmatrix = viewmatrix . P^T   (:501)
 *   campos     = camera centre = c2w_v[:3,3]                                 (:502)
 * camera_to_worlds: [V,3,4] row-major (nerfstudio convention, the tensor the reference indexes at :480);
 * intrinsics: [V,4] = {fovx, fovy, cx, cy} (radians; cx, cy = K[0,0,2], K[0,1,2] in NDC units);
 * out_cameras: [V,TGR_CAMERA_FLOATS] = {viewmatrix 16, projmatrix 16, campos 3, tan(fovx/2), tan(fovy/2), pad};
 * all pointers are device memory. */
#define TGR_CAMERA_FLOATS 40
int tgr_build_cameras(int32_t n_views, const float* camera_to_worlds, const float* intrinsics, float znear,
                      float zfar, float* out_cameras, void* stream);

/* ---- f4: marching tetrahedra (re-meshing after a geometry edit) ----
 * Replaces MarchingTetrahedraHelper._forward (Edit_core/tetgs_spatial/models/isosurface.py:112-184: boolean masks,
 * torch.unique(dim=0, return_inverse=True) over the sorted edge pairs, gathers through the triangle table).  Outputs are
 * identical to the reference's, ordering included: one vertex per unique grid edge with exactly one occupied end, in
 * lexicographic (min id, max id) order, position by the reference's fp32 expression; faces of the one-triangle tets
 * first, then of the two-triangle tets; face_to_tet = the tet of every face (what the keep / edit inheritance keys on,
 * tetgs_model.py:679-726); interp_v = the two grid vertices of every mesh vertex.
 * Three phases, because the output sizes only exist on the device (each of the first two synchronises the stream once —
 * this is set-up work per re-meshing, not the per-iteration path):
 *   tgr_mt_classify   counts[4] <- {valid tets, tets with one triangle, tets with two, -}
 *   tgr_mt_edges      counts[2] <- {unique edges of the valid tets, mesh vertices}
 *   tgr_mt_emit       verts [n_mesh_verts,3] f32, interp_v [n_mesh_verts,2] i64 (or NULL), faces [n_one + 2 n_two, 3] i64,
 *                     face_to_tet [n_one + 2 n_two] i64
 * level [n_verts] f32 (> 0 = inside), pos [n_verts,3] f32, tets [n_tets,4] i32; work1 / work2: device workspaces of
 * tgr_mt_classify_bytes(n_tets) / tgr_mt_edges_bytes(valid tets) bytes, handed from phase to phase. */
uint64_t tgr_mt_classify_bytes(int64_t n_tets);
uint64_t tgr_mt_edges_bytes(int64_t n_valid_tets);
int tgr_mt_classify(int32_t n_verts, int64_t n_tets, const float* level, const int32_t* tets, void* work1,
                    uint64_t work1_bytes, uint32_t* counts_host, void* stream);
int tgr_mt_edges(int32_t n_verts, int64_t n_tets, uint32_t n_valid, const float* level, const int32_t* tets, void* work1,
                 void* work2, uint64_t work2_bytes, uint32_t* counts_host, void* stream);
int tgr_mt_emit(int64_t n_tets, uint32_t n_valid, uint32_t n_one, uint32_t n_unique, const float* pos, const float* level,
                void* work1, void* work2, float* verts, int64_t* interp_v, int64_t* faces, int64_t* face_to_tet, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TETGS_RAST_H_ */
