// marching_tets.cu — marching tetrahedra on the GPU: level set on a tetrahedral grid -> triangle mesh + face -> tet map.
//
// Behavioural reference: MarchingTetrahedraHelper._forward, Edit_core/tetgs_spatial/models/isosurface.py:112-184
// (torch: boolean masks, torch.unique(dim=0, return_inverse=True) over the sorted edge pairs, gathers through the
// triangle table) — the step that turns the optimised SDF into the mesh the Gaussians are (re)bound to, and whose
// face_to_tet_idx drives the keep / edit inheritance (tetgs_model.py:679-726).  Outputs are IDENTICAL to the reference's,
// ordering included:
//   verts         one per unique grid edge with exactly one occupied end, in lexicographic (min id, max id) order of the
//                 edges; position = the reference's fp32 expression (two rounded products, one rounded sum — no FMA)
//   faces         tets with one triangle first (tet order), then tets with two (tet order, two consecutive rows)
//   face_to_tet   the tet of every face;  interp_v: the two grid vertices of every mesh vertex
// How: one pass classifies the tets (occupancy code, validity); the 6 edges of every valid tet are sorted as (min, max)
// pairs by two stable 32-bit radix sorts of this library (sort.cu; low word first), run heads give the unique edges
// and the inverse map, three exclusive scans (valid tets, one- / two-triangle tets, crossing edges) give every output
// its slot.  Sizes that only exist on the device are read back by the host between the three phases (this is set-up
// work, once per re-meshing — not the per-iteration path).
#include "common.cuh"

namespace tgr {

__constant__ int8_t MT_TRI[16][6] = {{-1, -1, -1, -1, -1, -1}, {1, 0, 2, -1, -1, -1}, {4, 0, 3, -1, -1, -1}, {1, 4, 2, 1, 3, 4},
                                     {3, 1, 5, -1, -1, -1},    {2, 3, 0, 2, 5, 3},    {1, 4, 0, 1, 5, 4},    {4, 2, 5, -1, -1, -1},
                                     {4, 5, 2, -1, -1, -1},    {4, 1, 0, 4, 5, 1},    {3, 2, 0, 3, 5, 2},    {1, 3, 5, -1, -1, -1},
                                     {4, 1, 2, 4, 3, 1},       {3, 0, 4, -1, -1, -1}, {2, 0, 1, -1, -1, -1}, {-1, -1, -1, -1, -1, -1}};
__constant__ uint8_t MT_NTRI[16] = {0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0};
__constant__ uint8_t MT_EDGE[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

// ---- exclusive scan of u32 (three launches: tile sums, scan of the sums by one CTA, apply) ------------------------
constexpr int SC_THREADS = 256, SC_IPT = 8, SC_TILE = SC_THREADS * SC_IPT;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t off = 0, tot = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { if (w < warp) off += s_warp[w]; tot += s_warp[w]; }
  __syncthreads();
  total = tot;
  return off + inc - v;
}

__global__ void __launch_bounds__(SC_THREADS) scan_tile_sums(const uint32_t* __restrict__ in, uint32_t* __restrict__ sums, uint64_t n) {
  __shared__ uint32_t s_warp[SC_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_IPT;
  uint32_t v = 0;
#pragma unroll
  for (int i = 0; i < SC_IPT; ++i) v += base + i < n ? in[base + i] : 0u;
  uint32_t tot;
  block_exclusive_scan(v, s_warp, tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) scan_sums(uint32_t* __restrict__ sums, uint32_t ntiles, uint32_t* __restrict__ total_out) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t t0 = 0; t0 < ntiles; t0 += 1024) {
    const uint32_t t = t0 + threadIdx.x;
    const uint32_t v = t < ntiles ? sums[t] : 0u;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan(v, s_warp, tot);
    const uint32_t carry = s_carry;
    if (t < ntiles) sums[t] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = s_carry;
}
__global__ void __launch_bounds__(SC_THREADS) scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                         const uint32_t* __restrict__ sums, uint64_t n) {
  __shared__ uint32_t s_warp[SC_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_IPT;
  uint32_t x[SC_IPT], v = 0;
#pragma unroll
  for (int i = 0; i < SC_IPT; ++i) { x[i] = base + i < n ? in[base + i] : 0u; v += x[i]; }
  uint32_t tot;
  uint32_t run = sums[blockIdx.x] + block_exclusive_scan(v, s_warp, tot);
#pragma unroll
  for (int i = 0; i < SC_IPT; ++i) {
    if (base + i < n) out[base + i] = run;
    run += x[i];
  }
}
// out may alias in.  sums: >= ceil(n / SC_TILE) words.
static void exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* sums, uint32_t* total_out, cudaStream_t s) {
  const uint32_t ntiles = (uint32_t)((n + SC_TILE - 1) / SC_TILE);
  if (ntiles == 0) { cudaMemsetAsync(total_out, 0, 4, s); return; }
  scan_tile_sums<<<ntiles, SC_THREADS, 0, s>>>(in, sums, n);
  scan_sums<<<1, 1024, 0, s>>>(sums, ntiles, total_out);
  scan_apply<<<ntiles, SC_THREADS, 0, s>>>(in, out, sums, n);
  count_launch(3);
}

// ---- phase 1: classify the tets -----------------------------------------------------------------------------------
__global__ void mt_classify_kernel(int64_t n_tets, const float* __restrict__ level, const int32_t* __restrict__ tets,
                                   uint8_t* __restrict__ code, uint32_t* __restrict__ f_valid, uint32_t* __restrict__ f_one,
                                   uint32_t* __restrict__ f_two) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tets) return;
  const int4 v = reinterpret_cast<const int4*>(tets)[t];
  const uint32_t c = (level[v.x] > 0.f ? 1u : 0u) | (level[v.y] > 0.f ? 2u : 0u) | (level[v.z] > 0.f ? 4u : 0u) | (level[v.w] > 0.f ? 8u : 0u);
  code[t] = (uint8_t)c;
  const uint32_t nt = MT_NTRI[c];          // 0 exactly for the codes 0 and 15: all out / all in
  f_valid[t] = nt != 0u;
  f_one[t] = nt == 1u;
  f_two[t] = nt == 2u;
}

// ---- phase 2: edges of the valid tets, unique edges, crossing edges -----------------------------------------------
__global__ void mt_edges_kernel(int64_t n_tets, const int32_t* __restrict__ tets, const uint8_t* __restrict__ code,
                                const uint32_t* __restrict__ valid_idx, uint32_t* __restrict__ e_lo, uint32_t* __restrict__ e_hi,
                                uint32_t* __restrict__ e_id) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tets) return;
  const uint32_t c = code[t];
  if (MT_NTRI[c] == 0) return;
  const int4 q = reinterpret_cast<const int4*>(tets)[t];
  const int v[4] = {q.x, q.y, q.z, q.w};
  const uint32_t base = valid_idx[t] * 6u;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int a = v[MT_EDGE[k][0]], b = v[MT_EDGE[k][1]];
    e_hi[base + k] = (uint32_t)min(a, b);   // "hi" = the more significant sort word = the smaller vertex id (sort_edges)
    e_lo[base + k] = (uint32_t)max(a, b);
    e_id[base + k] = base + k;
  }
}
__global__ void mt_gather_kernel(uint32_t n, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[ids[i]];
}
// order[i] = edge slot at sorted position i.  head flag: first of a run of equal (hi, lo)
__global__ void mt_heads_kernel(uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ e_hi,
                                const uint32_t* __restrict__ e_lo, uint32_t* __restrict__ head) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool h = true;
  if (i > 0) {
    const uint32_t a = order[i], b = order[i - 1];
    h = e_hi[a] != e_hi[b] || e_lo[a] != e_lo[b];
  }
  head[i] = h ? 1u : 0u;
}
// uid_excl[i] = exclusive scan of head -> unique id of position i = uid_excl[i] + head[i] - 1
__global__ void mt_unique_kernel(uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ head,
                                 const uint32_t* __restrict__ uid_excl, const uint32_t* __restrict__ e_hi,
                                 const uint32_t* __restrict__ e_lo, const float* __restrict__ level, uint32_t* __restrict__ edge_uid,
                                 uint32_t* __restrict__ u_a, uint32_t* __restrict__ u_b, uint32_t* __restrict__ u_cross) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t e = order[i];
  const uint32_t uid = uid_excl[i] + head[i] - 1u;
  edge_uid[e] = uid;
  if (head[i]) {
    const uint32_t a = e_hi[e], b = e_lo[e];
    u_a[uid] = a; u_b[uid] = b;
    u_cross[uid] = ((level[a] > 0.f) != (level[b] > 0.f)) ? 1u : 0u;   // exactly one end occupied (isosurface.py:124)
  }
}

// ---- phase 3: vertices and faces -------------------------------------------------------------------------------------
__global__ void mt_verts_kernel(uint32_t n_unique, const uint32_t* __restrict__ u_a, const uint32_t* __restrict__ u_b,
                                const uint32_t* __restrict__ u_cross, const uint32_t* __restrict__ vid_excl,
                                const float* __restrict__ pos, const float* __restrict__ level, float* __restrict__ verts,
                                int64_t* __restrict__ interp_v) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_unique || !u_cross[u]) return;
  const uint32_t a = u_a[u], b = u_b[u], vid = vid_excl[u];
  // isosurface.py:135-141: sdf pair (s_a, -s_b), weights = flip / sum, vertex = sum of the two weighted ends; every
  // operation rounded to fp32 on its own, as the eager torch ops do
  const float sa = level[a], sb = -level[b];
  const float den = __fadd_rn(sa, sb);
  const float wa = __fdiv_rn(sb, den), wb = __fdiv_rn(sa, den);
#pragma unroll
  for (int c = 0; c < 3; ++c) verts[3 * (size_t)vid + c] = __fadd_rn(__fmul_rn(pos[3 * (size_t)a + c], wa), __fmul_rn(pos[3 * (size_t)b + c], wb));
  if (interp_v) { interp_v[2 * (size_t)vid] = a; interp_v[2 * (size_t)vid + 1] = b; }
}
__global__ void mt_faces_kernel(int64_t n_tets, const uint8_t* __restrict__ code, const uint32_t* __restrict__ valid_idx,
                                const uint32_t* __restrict__ one_idx, const uint32_t* __restrict__ two_idx, uint32_t n_one,
                                const uint32_t* __restrict__ edge_uid, const uint32_t* __restrict__ u_cross,
                                const uint32_t* __restrict__ vid_excl, int64_t* __restrict__ faces, int64_t* __restrict__ face_to_tet) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tets) return;
  const uint32_t c = code[t];
  const uint32_t nt = MT_NTRI[c];
  if (nt == 0) return;
  const uint32_t ebase = valid_idx[t] * 6u;
  const size_t f0 = nt == 1 ? (size_t)one_idx[t] : (size_t)n_one + 2 * (size_t)two_idx[t];
  for (uint32_t k = 0; k < 3 * nt; ++k) {
    const uint32_t u = edge_uid[ebase + (uint32_t)MT_TRI[c][k]];
    faces[3 * f0 + k] = u_cross[u] ? (int64_t)vid_excl[u] : -1;   // (the triangle table only ever picks crossing edges)
  }
  for (uint32_t k = 0; k < nt; ++k) face_to_tet[f0 + k] = t;
}

struct MtWork1 {   // sized by n_tets
  uint8_t* code; uint32_t *valid_idx, *one_idx, *two_idx, *sums, *counts;
  uint64_t bytes;
};
static MtWork1 carve_mt1(void* base, int64_t n_tets) {
  MtWork1 w;
  char* p = static_cast<char*>(base);
  const uint64_t n = (uint64_t)n_tets;
  w.counts = carve<uint32_t>(p, 32);
  w.code = carve<uint8_t>(p, n);
  w.valid_idx = carve<uint32_t>(p, n);
  w.one_idx = carve<uint32_t>(p, n);
  w.two_idx = carve<uint32_t>(p, n);
  w.sums = carve<uint32_t>(p, (n + SC_TILE - 1) / SC_TILE + 1);
  w.bytes = (uint64_t)(p - static_cast<char*>(base)) + 128;
  return w;
}
struct MtWork2 {   // sized by n_valid
  uint32_t *e_lo, *e_hi, *ka, *va, *kb, *vb, *head, *uid, *edge_uid, *u_a, *u_b, *u_cross, *vid, *sums, *sort_temp;
  uint64_t bytes;
};
static MtWork2 carve_mt2(void* base, uint64_t n_valid) {
  MtWork2 w;
  char* p = static_cast<char*>(base);
  const uint64_t E = 6 * n_valid;
  w.e_lo = carve<uint32_t>(p, E); w.e_hi = carve<uint32_t>(p, E);
  w.ka = carve<uint32_t>(p, E); w.va = carve<uint32_t>(p, E); w.kb = carve<uint32_t>(p, E); w.vb = carve<uint32_t>(p, E);
  w.head = carve<uint32_t>(p, E); w.uid = carve<uint32_t>(p, E); w.edge_uid = carve<uint32_t>(p, E);
  w.u_a = carve<uint32_t>(p, E); w.u_b = carve<uint32_t>(p, E); w.u_cross = carve<uint32_t>(p, E); w.vid = carve<uint32_t>(p, E);
  w.sums = carve<uint32_t>(p, (E + SC_TILE - 1) / SC_TILE + 1);
  w.sort_temp = carve<uint32_t>(p, sort_temp_bytes(E) / 4);
  w.bytes = (uint64_t)(p - static_cast<char*>(base)) + 128;
  return w;
}
static int bits_for(uint64_t n) { int b = 1; while (b < 32 && (1ull << b) < n) ++b; return b; }

}  // namespace tgr

using namespace tgr;

extern "C" uint64_t tgr_mt_classify_bytes(int64_t n_tets) { return carve_mt1(nullptr, n_tets).bytes; }
extern "C" uint64_t tgr_mt_edges_bytes(int64_t n_valid_tets) { return carve_mt2(nullptr, (uint64_t)n_valid_tets).bytes; }

// Phase 1.  counts_host[4] <- {valid tets, tets with one triangle, tets with two, 0} (synchronises the stream).
extern "C" int tgr_mt_classify(int32_t n_verts, int64_t n_tets, const float* level, const int32_t* tets, void* work1,
                               uint64_t work1_bytes, uint32_t* counts_host, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_tets < 0 || n_verts < 0 || (n_tets > 0 && (!level || !tets)) || !work1 || !counts_host) { set_error("mt_classify: bad arguments"); return 1; }
  if (work1_bytes < tgr_mt_classify_bytes(n_tets)) { set_error("mt_classify: workspace too small"); return 1; }
  if ((uint64_t)n_tets * 6 >= (1ull << 30)) { set_error("mt_classify: more than 2^30 / 6 tetrahedra"); return 1; }
  MtWork1 w = carve_mt1(work1, n_tets);
  cudaMemsetAsync(w.counts, 0, 128, s);
  if (n_tets > 0) {
    const unsigned blocks = (unsigned)((n_tets + 255) / 256);
    mt_classify_kernel<<<blocks, 256, 0, s>>>(n_tets, level, tets, w.code, w.valid_idx, w.one_idx, w.two_idx);
    count_launch();
    exclusive_scan_u32(w.valid_idx, w.valid_idx, (uint64_t)n_tets, w.sums, w.counts + 0, s);
    exclusive_scan_u32(w.one_idx, w.one_idx, (uint64_t)n_tets, w.sums, w.counts + 1, s);
    exclusive_scan_u32(w.two_idx, w.two_idx, (uint64_t)n_tets, w.sums, w.counts + 2, s);
  }
  cudaError_t e = cudaMemcpyAsync(counts_host, w.counts, 16, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { set_error("mt_classify: %s", cudaGetErrorString(e)); return 2; }
  return 0;
}

// Phase 2.  counts_host[2] <- {unique edges, crossing edges = mesh vertices} (synchronises the stream).
extern "C" int tgr_mt_edges(int32_t n_verts, int64_t n_tets, uint32_t n_valid, const float* level, const int32_t* tets,
                            void* work1, void* work2, uint64_t work2_bytes, uint32_t* counts_host, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!work1 || !work2 || !counts_host) { set_error("mt_edges: bad arguments"); return 1; }
  if (work2_bytes < tgr_mt_edges_bytes(n_valid)) { set_error("mt_edges: workspace too small"); return 1; }
  MtWork1 w1 = carve_mt1(work1, n_tets);
  MtWork2 w = carve_mt2(work2, n_valid);
  const uint32_t E = 6u * n_valid;
  counts_host[0] = counts_host[1] = 0;
  if (E == 0) return 0;
  const unsigned tb = (unsigned)((n_tets + 255) / 256), eb = (E + 255) / 256;
  mt_edges_kernel<<<tb, 256, 0, s>>>(n_tets, tets, w1.code, w1.valid_idx, w.e_lo, w.e_hi, w.va);
  count_launch();
  // lexicographic (hi, lo) order = stable sort by lo, then stable sort by hi
  const int vbits = bits_for((uint64_t)n_verts);
  bool in_b = false;
  cudaMemcpyAsync(w.ka, w.e_lo, (size_t)E * 4, cudaMemcpyDeviceToDevice, s);
  if (int rc = launch_sort_pairs(E, nullptr, w.ka, w.va, w.kb, w.vb, false, 0, vbits, w.sort_temp, s, &in_b)) return rc;
  uint32_t* order1 = in_b ? w.vb : w.va;
  // second sort: keys = hi of the edges in the order of the first sort; values keep travelling
  uint32_t *k2a = in_b ? w.kb : w.ka, *v2a = order1, *k2b = in_b ? w.ka : w.kb, *v2b = in_b ? w.va : w.vb;
  mt_gather_kernel<<<eb, 256, 0, s>>>(E, order1, w.e_hi, k2a);
  count_launch();
  bool in_b2 = false;
  if (int rc = launch_sort_pairs(E, nullptr, k2a, v2a, k2b, v2b, false, 0, vbits, w.sort_temp, s, &in_b2)) return rc;
  const uint32_t* order = in_b2 ? v2b : v2a;
  mt_heads_kernel<<<eb, 256, 0, s>>>(E, order, w.e_hi, w.e_lo, w.head);
  count_launch();
  exclusive_scan_u32(w.head, w.uid, E, w.sums, w1.counts + 4, s);
  mt_unique_kernel<<<eb, 256, 0, s>>>(E, order, w.head, w.uid, w.e_hi, w.e_lo, level, w.edge_uid, w.u_a, w.u_b, w.u_cross);
  count_launch();
  // crossing edges -> vertex slots.  n_unique is only known on the device: scan over E slots, the tail holds zeros
  cudaError_t e = cudaMemcpyAsync(counts_host, w1.counts + 4, 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { set_error("mt_edges: %s", cudaGetErrorString(e)); return 2; }
  const uint32_t n_unique = counts_host[0];
  exclusive_scan_u32(w.u_cross, w.vid, n_unique, w.sums, w1.counts + 5, s);
  e = cudaMemcpyAsync(counts_host + 1, w1.counts + 5, 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { set_error("mt_edges: %s", cudaGetErrorString(e)); return 2; }
  return check_launch("mt_edges", false, s);
}

// Phase 3.  verts [n_mesh_verts,3] f32, interp_v [n_mesh_verts,2] i64 (may be NULL), faces [n_one + 2 n_two, 3] i64,
// face_to_tet [n_one + 2 n_two] i64.
extern "C" int tgr_mt_emit(int64_t n_tets, uint32_t n_valid, uint32_t n_one, uint32_t n_unique, const float* pos,
                           const float* level, void* work1, void* work2, float* verts, int64_t* interp_v, int64_t* faces,
                           int64_t* face_to_tet, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!work1 || !work2 || !pos || !level) { set_error("mt_emit: bad arguments"); return 1; }
  if (n_valid == 0) return 0;
  MtWork1 w1 = carve_mt1(work1, n_tets);
  MtWork2 w = carve_mt2(work2, n_valid);
  if (n_unique > 0) {
    mt_verts_kernel<<<(n_unique + 255) / 256, 256, 0, s>>>(n_unique, w.u_a, w.u_b, w.u_cross, w.vid, pos, level, verts, interp_v);
    count_launch();
  }
  mt_faces_kernel<<<(unsigned)((n_tets + 255) / 256), 256, 0, s>>>(n_tets, w1.code, w1.valid_idx, w1.one_idx, w1.two_idx, n_one,
                                                                   w.edge_uid, w.u_cross, w.vid, faces, face_to_tet);
  count_launch();
  return check_launch("mt_emit", false, s);
}
