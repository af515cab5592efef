// Iso-surface extraction from a DPSR indicator grid (the step after slide_dpsr_forward).  C ABI: include/slide_sap.h.
//
// The reference hands every grid to scikit-image on the CPU (dpsr_utils/utils.py:246-287 mc_from_psr ->
// skimage.measure.marching_cubes, Lewiner's variant) -- ~0.1 s per 128^3 grid, 7x the whole GPU stage before it.  scikit-image
// is not available offline, so its triangulation cannot be pinned; this file extracts the SAME level set with marching
// tetrahedra (every cell split into the six tetrahedra around its main diagonal, a decomposition that matches across cell
// faces), which needs no 256-case table, has no ambiguous cases and is watertight by construction.  Vertices lie on grid /
// face-diagonal / body-diagonal edges at the linear zero crossing; vertex i of the output is the i-th crossing in (node,
// edge type) order and triangles come in (cell, tetrahedron) order, so the output is deterministic and the numpy restatement in
// oracle/mesh_oracle.py reproduces it bit for bit.
//
// Three launches + two scans per grid: per-node crossing masks -> exclusive scan -> vertices; per-cell triangle counts ->
// exclusive scan -> faces.  All HBM-streaming work: the grid is read ~3 times (L2-resident at 8 MB).
#include "../../include/slide_sap.h"
#include "common.cuh"

namespace slide {

namespace {

// cube corners: v0 (0,0,0) v1 (1,0,0) v2 (1,1,0) v3 (0,1,0) v4 (0,0,1) v5 (1,0,1) v6 (1,1,1) v7 (0,1,1)
__constant__ int c_corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
// the six tetrahedra around the diagonal v0-v6
__constant__ int c_tet[6][4] = {{0, 5, 1, 6}, {0, 1, 2, 6}, {0, 2, 3, 6}, {0, 3, 7, 6}, {0, 7, 4, 6}, {0, 4, 5, 6}};

constexpr int SCAN_THREADS = 1024, SCAN_ITEMS = 4, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ bool inside(float v, float level) { return v < level; }

// mask bit (type-1) of node (x,y,z): the edge from the node in direction type = dx + 2 dy + 4 dz crosses the level
__global__ void mc_node_kernel(const float *__restrict__ phi, int R, float level, unsigned char *__restrict__ mask,
                               int *__restrict__ count) {
  pdl_wait();
  pdl_trigger();
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)R * R * R;
  if (n >= total) return;
  const int z = (int)(n % R), y = (int)((n / R) % R), x = (int)(n / ((long long)R * R));
  const bool a = inside(phi[n], level);
  unsigned m = 0;
#pragma unroll
  for (int t = 1; t <= 7; ++t) {
    const int dx = t & 1, dy = (t >> 1) & 1, dz = (t >> 2) & 1;
    if (x + dx < R && y + dy < R && z + dz < R) {
      const bool b = inside(phi[((long long)(x + dx) * R + (y + dy)) * R + (z + dz)], level);
      if (a != b) m |= 1u << (t - 1);
    }
  }
  mask[n] = (unsigned char)m;
  count[n] = __popc(m);
}

// ---- exclusive scan of int32 (three launches; up to SCAN_TILE^2 = 16.7 M items) ------------------------------------------------
__device__ __forceinline__ int block_exclusive(int v, int *total) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  const int base = warp ? wsum[warp - 1] : 0;
  if (total) *total = wsum[31];
  __syncthreads();
  return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const int *__restrict__ in, long long n, int *__restrict__ sums) {
  pdl_wait();
  pdl_trigger();
  const long long i0 = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (i0 + k < n) s += in[i0 + k];
  int total;
  block_exclusive(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// one block: sums[0..nb) -> exclusive in place, grand total to sums[nb]
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(int *__restrict__ sums, int nb) {
  pdl_wait();
  pdl_trigger();
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int i = threadIdx.x * SCAN_ITEMS + k;
    v[k] = i < nb ? sums[i] : 0;
    s += v[k];
  }
  int total;
  int run = block_exclusive(s, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int i = threadIdx.x * SCAN_ITEMS + k;
    if (i < nb) sums[i] = run;
    run += v[k];
  }
  if (threadIdx.x == 0) sums[nb] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(int *__restrict__ data, long long n, const int *__restrict__ sums) {
  pdl_wait();
  pdl_trigger();
  const long long i0 = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = i0 + k < n ? data[i0 + k] : 0;
    s += v[k];
  }
  int run = block_exclusive(s, nullptr) + sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (i0 + k < n) data[i0 + k] = run;
    run += v[k];
  }
}

int exclusive_scan(int *data, long long n, int *sums, cudaStream_t st) {
  const int nb = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  if (nb > SCAN_TILE) return SLIDE_ERR_UNSUPPORTED;
  int rc;
  launch_k(scan_tile_sums_kernel, dim3(nb), dim3(SCAN_THREADS), 0, st, (const int *)data, n, sums);
  if ((rc = after_launch())) return rc;
  launch_k(scan_sums_kernel, dim3(1), dim3(SCAN_THREADS), 0, st, sums, nb);
  if ((rc = after_launch())) return rc;
  launch_k(scan_apply_kernel, dim3(nb), dim3(SCAN_THREADS), 0, st, data, n, (const int *)sums);
  return after_launch();
}

// ---- per cell: number of triangles -----------------------------------------------------------------------------------------------
__device__ __forceinline__ int tet_triangles(int m4) {
  const int c = __popc(m4);
  return (c == 0 || c == 4) ? 0 : (c == 2 ? 2 : 1);
}

__global__ void mc_cell_count_kernel(const float *__restrict__ phi, int R, float level, int *__restrict__ count) {
  pdl_wait();
  pdl_trigger();
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)R * R * R;
  if (n >= total) return;
  const int z = (int)(n % R), y = (int)((n / R) % R), x = (int)(n / ((long long)R * R));
  int tri = 0;
  if (x + 1 < R && y + 1 < R && z + 1 < R) {
    unsigned in8 = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (inside(phi[((long long)(x + c_corner[c][0]) * R + (y + c_corner[c][1])) * R + (z + c_corner[c][2])], level)) in8 |= 1u << c;
    if (in8 != 0 && in8 != 255) {
#pragma unroll
      for (int t = 0; t < 6; ++t) {
        int m4 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) m4 |= ((in8 >> c_tet[t][k]) & 1) << k;
        tri += tet_triangles(m4);
      }
    }
  }
  count[n] = tri;  // cells on the upper faces of the grid hold 0: one scan array for nodes and cells
}

// ---- vertices ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float grad_axis(const float *phi, int R, int x, int y, int z, int axis) {
  // numpy.gradient: central differences inside, one-sided at the faces
  int c[3] = {x, y, z};
  const int lo = c[axis] > 0 ? c[axis] - 1 : c[axis], hi = c[axis] < R - 1 ? c[axis] + 1 : c[axis];
  int a[3] = {x, y, z}, b[3] = {x, y, z};
  a[axis] = lo;
  b[axis] = hi;
  const float fa = phi[((long long)a[0] * R + a[1]) * R + a[2]], fb = phi[((long long)b[0] * R + b[1]) * R + b[2]];
  return __fdiv_rn(__fsub_rn(fb, fa), (float)(hi - lo));
}

__global__ void mc_vertex_kernel(const float *__restrict__ phi, int R, float level, const unsigned char *__restrict__ mask,
                                 const int *__restrict__ vscan, float scale, float *__restrict__ verts,
                                 float *__restrict__ normals) {
  pdl_wait();
  pdl_trigger();
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)R * R * R;
  if (n >= total) return;
  const unsigned m = mask[n];
  if (!m) return;
  const int z = (int)(n % R), y = (int)((n / R) % R), x = (int)(n / ((long long)R * R));
  const float fa = phi[n];
  int id = vscan[n];
  float ga[3];
  if (normals) {
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) ga[ax] = grad_axis(phi, R, x, y, z, ax);
  }
#pragma unroll
  for (int t = 1; t <= 7; ++t) {
    if (!((m >> (t - 1)) & 1)) continue;
    const int dx = t & 1, dy = (t >> 1) & 1, dz = (t >> 2) & 1;
    const float fb = phi[((long long)(x + dx) * R + (y + dy)) * R + (z + dz)];
    const float s = __fdiv_rn(__fsub_rn(level, fa), __fsub_rn(fb, fa));  // in [0, 1]: fa, fb lie on opposite sides
    float *v = verts + (size_t)id * 3;
    v[0] = __fmul_rn(__fmaf_rn(s, (float)dx, (float)x), scale);
    v[1] = __fmul_rn(__fmaf_rn(s, (float)dy, (float)y), scale);
    v[2] = __fmul_rn(__fmaf_rn(s, (float)dz, (float)z), scale);
    if (normals) {
      float g[3];
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        const float gb = grad_axis(phi, R, x + dx, y + dy, z + dz, ax);
        g[ax] = __fadd_rn(ga[ax], __fmul_rn(s, __fsub_rn(gb, ga[ax])));
      }
      const float len = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(g[0], g[0]), __fmul_rn(g[1], g[1])), __fmul_rn(g[2], g[2])));
      const float inv = len > 0.f ? __fdiv_rn(1.0f, len) : 0.f;
      float *nn = normals + (size_t)id * 3;
      nn[0] = __fmul_rn(g[0], inv);
      nn[1] = __fmul_rn(g[1], inv);
      nn[2] = __fmul_rn(g[2], inv);
    }
    ++id;
  }
}

// ---- faces ---------------------------------------------------------------------------------------------------------------------------
// vertex id of the crossing on the edge between cube corners u and v of the cell at (x,y,z)
__device__ __forceinline__ int edge_vertex(const unsigned char *mask, const int *vscan, int R, int x, int y, int z, int u, int v) {
  const int su = c_corner[u][0] + c_corner[u][1] + c_corner[u][2], sv = c_corner[v][0] + c_corner[v][1] + c_corner[v][2];
  const int lo = su < sv ? u : v, hi = su < sv ? v : u;
  const int dx = c_corner[hi][0] - c_corner[lo][0], dy = c_corner[hi][1] - c_corner[lo][1], dz = c_corner[hi][2] - c_corner[lo][2];
  const int slot = (dx + 2 * dy + 4 * dz) - 1;
  const long long n = ((long long)(x + c_corner[lo][0]) * R + (y + c_corner[lo][1])) * R + (z + c_corner[lo][2]);
  return vscan[n] + __popc((unsigned)mask[n] & ((1u << slot) - 1u));
}

__global__ void mc_face_kernel(const float *__restrict__ phi, int R, float level, const unsigned char *__restrict__ mask,
                               const int *__restrict__ vscan, const int *__restrict__ fscan, const float *__restrict__ verts,
                               int *__restrict__ faces) {
  pdl_wait();
  pdl_trigger();
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)R * R * R;
  if (n >= total) return;
  const int z = (int)(n % R), y = (int)((n / R) % R), x = (int)(n / ((long long)R * R));
  if (!(x + 1 < R && y + 1 < R && z + 1 < R)) return;
  unsigned in8 = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (inside(phi[((long long)(x + c_corner[c][0]) * R + (y + c_corner[c][1])) * R + (z + c_corner[c][2])], level)) in8 |= 1u << c;
  if (in8 == 0 || in8 == 255) return;
  int f = fscan[n];
  for (int t = 0; t < 6; ++t) {
    int cor[4], m4 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cor[k] = c_tet[t][k];
      m4 |= ((in8 >> cor[k]) & 1) << k;
    }
    const int cnt = __popc(m4);
    if (cnt == 0 || cnt == 4) continue;
    int tri[2][3];
    int ntri;
    // direction from the inside corners towards the outside corners (sum of corner offsets, sign only matters)
    float dir[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sgn = ((m4 >> k) & 1) ? -1.f : 1.f;
      const float w = __fdiv_rn(sgn, (float)(((m4 >> k) & 1) ? cnt : 4 - cnt));
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) dir[ax] = __fadd_rn(dir[ax], __fmul_rn(w, (float)c_corner[cor[k]][ax]));
    }
    if (cnt == 1 || cnt == 3) {
      const int lone_bit = cnt == 1 ? m4 : (~m4 & 15);
      const int L = __ffs(lone_bit) - 1;
      int o[3], j = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k != L) o[j++] = k;
      for (int k = 0; k < 3; ++k) tri[0][k] = edge_vertex(mask, vscan, R, x, y, z, cor[L], cor[o[k]]);
      ntri = 1;
    } else {
      int P[2], Q[2], a = 0, b = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((m4 >> k) & 1)
          P[a++] = k;
        else
          Q[b++] = k;
      }
      const int e0 = edge_vertex(mask, vscan, R, x, y, z, cor[P[0]], cor[Q[0]]);
      const int e1 = edge_vertex(mask, vscan, R, x, y, z, cor[P[0]], cor[Q[1]]);
      const int e2 = edge_vertex(mask, vscan, R, x, y, z, cor[P[1]], cor[Q[1]]);
      const int e3 = edge_vertex(mask, vscan, R, x, y, z, cor[P[1]], cor[Q[0]]);
      tri[0][0] = e0, tri[0][1] = e1, tri[0][2] = e2;
      tri[1][0] = e0, tri[1][1] = e2, tri[1][2] = e3;
      ntri = 2;
    }
    for (int q = 0; q < ntri; ++q) {
      const float *p0 = verts + (size_t)tri[q][0] * 3, *p1 = verts + (size_t)tri[q][1] * 3, *p2 = verts + (size_t)tri[q][2] * 3;
      // explicit single-rounding operations: the oracle repeats them, so the winding of (near-)degenerate triangles agrees too
      const float ux = __fsub_rn(p1[0], p0[0]), uy = __fsub_rn(p1[1], p0[1]), uz = __fsub_rn(p1[2], p0[2]);
      const float vx = __fsub_rn(p2[0], p0[0]), vy = __fsub_rn(p2[1], p0[1]), vz = __fsub_rn(p2[2], p0[2]);
      const float nx = __fsub_rn(__fmul_rn(uy, vz), __fmul_rn(uz, vy)), ny = __fsub_rn(__fmul_rn(uz, vx), __fmul_rn(ux, vz)),
                  nz = __fsub_rn(__fmul_rn(ux, vy), __fmul_rn(uy, vx));
      const float dot = __fadd_rn(__fadd_rn(__fmul_rn(nx, dir[0]), __fmul_rn(ny, dir[1])), __fmul_rn(nz, dir[2]));
      const bool flip = dot < 0.f;  // orient: normal from inside (phi < level) to outside
      int *o = faces + (size_t)(f++) * 3;
      o[0] = tri[q][0];
      o[1] = flip ? tri[q][2] : tri[q][1];
      o[2] = flip ? tri[q][1] : tri[q][2];
    }
  }
}

struct McLayout {
  size_t mask, vscan, fscan, sums, total;
};

McLayout mc_layout(int R) {
  const size_t vol = (size_t)R * R * R;
  McLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  L.mask = take(vol);
  L.vscan = take(vol * sizeof(int));
  L.fscan = take(vol * sizeof(int));
  L.sums = take((size_t)(2 * (SCAN_TILE + 1)) * sizeof(int));
  L.total = off;
  return L;
}

}  // namespace

}  // namespace slide

using namespace slide;

extern "C" {

int slide_mc_workspace_bytes(int res, size_t *bytes) {
  if (!bytes || res < 2 || res > 256) return SLIDE_ERR_INVALID;
  *bytes = mc_layout(res).total;
  return SLIDE_OK;
}

int slide_mc_count(const float *phi, int res, float level, void *workspace, size_t workspace_bytes, int *counts,
                   slide_stream_t stream) {
  if (!phi || !workspace || !counts || res < 2 || res > 256) return SLIDE_ERR_INVALID;
  const McLayout L = mc_layout(res);
  if (workspace_bytes < L.total) return SLIDE_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  char *ws = (char *)workspace;
  unsigned char *mask = (unsigned char *)(ws + L.mask);
  int *vscan = (int *)(ws + L.vscan), *fscan = (int *)(ws + L.fscan), *sums = (int *)(ws + L.sums);
  const long long vol = (long long)res * res * res;
  const unsigned grid = (unsigned)ceil_div_ll(vol, 256);
  const int nb = (int)((vol + SCAN_TILE - 1) / SCAN_TILE);
  int rc;
  launch_k(mc_node_kernel, dim3(grid), dim3(256), 0, st, phi, res, level, mask, vscan);
  if ((rc = after_launch())) return rc;
  if ((rc = exclusive_scan(vscan, vol, sums, st))) return rc;
  launch_k(mc_cell_count_kernel, dim3(grid), dim3(256), 0, st, phi, res, level, fscan);
  if ((rc = after_launch())) return rc;
  if ((rc = exclusive_scan(fscan, vol, sums + SCAN_TILE + 1, st))) return rc;
  // grand totals: sums[nb] of each scan
  if ((rc = cuda_rc(cudaMemcpyAsync(counts, sums + nb, sizeof(int), cudaMemcpyDeviceToDevice, st)))) return rc;
  return cuda_rc(cudaMemcpyAsync(counts + 1, sums + SCAN_TILE + 1 + nb, sizeof(int), cudaMemcpyDeviceToDevice, st));
}

int slide_mc_emit(const float *phi, int res, float level, const void *workspace, float vertex_scale, float *verts,
                  float *normals, int *faces, slide_stream_t stream) {
  if (!phi || !workspace || !verts || !faces || res < 2 || res > 256) return SLIDE_ERR_INVALID;
  const McLayout L = mc_layout(res);
  cudaStream_t st = (cudaStream_t)stream;
  const char *ws = (const char *)workspace;
  const unsigned char *mask = (const unsigned char *)(ws + L.mask);
  const int *vscan = (const int *)(ws + L.vscan), *fscan = (const int *)(ws + L.fscan);
  const long long vol = (long long)res * res * res;
  const unsigned grid = (unsigned)ceil_div_ll(vol, 256);
  int rc;
  launch_k(mc_vertex_kernel, dim3(grid), dim3(256), 0, st, phi, res, level, mask, vscan, vertex_scale, verts, normals);
  if ((rc = after_launch())) return rc;
  launch_k(mc_face_kernel, dim3(grid), dim3(256), 0, st, phi, res, level, mask, vscan, fscan, (const float *)verts, faces);
  return after_launch();
}

}  // extern "C"
