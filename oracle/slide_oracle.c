/*
 * slide_oracle.c -- CPU restatement of the native point-set ops on SLIDE's sampling path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under slide_b200/ may link, import or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * What is restated (citations are relative to /root/reference, EXT = pointnet2_ops_lib/pointnet2_ops/_ext-src):
 *   so_fps                 EXT/src/sampling_gpu.cu:69-173   + host init EXT/src/sampling.cpp:66-87
 *   so_gather(_grad)       EXT/src/sampling_gpu.cu:8-57
 *   so_ball_query          EXT/src/ball_query_gpu.cu:9-47   + zero init EXT/src/ball_query.cpp:21-27
 *   so_group(_grad)        EXT/src/group_points_gpu.cu:8-75
 *   so_three_nn            EXT/src/interpolate_gpu.cu:9-59
 *   so_three_interpolate(_grad)  EXT/src/interpolate_gpu.cu:72-154
 *   so_knn / so_fps_p3d    pytorch3d==0.7.0 (environment.yml:117), NOT in the reference tree:
 *                          call sites pointnet2_ops/pointnet2_utils.py:370,506-507 and
 *                          models/point_upsample_decoder.py:178-180; FPS semantics mirrored by the
 *                          vendored data_utils/points_sampling.py:13-118.
 *
 * Floating-point contraction: the reference kernels were compiled here (nvcc 12.9, -O3, sm_100a)
 * and their SASS read back.  `a*a + b*b + c*c` becomes FMUL t=b*b; FFMA t=a*a+t; FFMA t=c*c+t,
 * i.e. fmaf(c,c, fmaf(a,a, b*b)); `p1*w1 + p2*w2 + p3*w3` becomes fmaf(p3,w3, fmaf(p1,w1, p2*w2)).
 * This file must be compiled with -ffp-contract=off so that only the explicit fmaf() calls fuse.
 *
 * Parity status: the _ext ops are pinned against the reference's own kernel source semantics and SASS
 * (and against the reference CUDA build on a GPU box when oracle/_ref exists).  so_fps_p3d is pinned against
 * the copy of pytorch3d's reference implementation vendored in the reference tree
 * (data_utils/points_sampling.py::sample_farthest_points_naive; tests/golden/make_golden_fps.py).  so_knn is
 * "parity unpinned": pytorch3d's source is absent; it restates the published algorithm (brute force,
 * squared L2 accumulated x,y,z with FMA, ascending sort).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SO_OK 0
#define SO_EINVAL (-1)

/* x*x + y*y + z*z exactly as the sm_100a SASS of the reference evaluates it */
static inline float sumsq3_ref(float x, float y, float z) {
  float t = y * y;
  t = fmaf(x, x, t);
  return fmaf(z, z, t);
}

/* EXT/include/cuda_utils.h:15-19 -- block size chosen by the reference launcher (double log!) */
int so_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

/* ------------------------------------------------------------------ FPS (pointnet2_ops._ext) */
/* Literal simulation of one CTA of `S` threads: strided per-thread scan, then the shared-memory
 * tree where the left slot survives ties (sampling_gpu.cu:59-65, 95-170). */
int so_fps(const float *xyz, int B, int N, int m, int *idx) {
  if (B < 0 || N <= 0 || m < 0) return SO_EINVAL;
  const int S = so_opt_n_threads(N);
  float *temp = (float *)malloc(sizeof(float) * (size_t)N);
  float *dv = (float *)malloc(sizeof(float) * (size_t)S);
  int *di = (int *)malloc(sizeof(int) * (size_t)S);
  if (!temp || !dv || !di) { free(temp); free(dv); free(di); return SO_EINVAL; }
  for (int b = 0; b < B; ++b) {
    const float *p = xyz + (size_t)b * N * 3;
    int *out = idx + (size_t)b * m;
    for (int k = 0; k < N; ++k) temp[k] = 1e10f; /* sampling.cpp:74-76 */
    for (int j = 0; j < m; ++j) out[j] = 0;      /* torch::zeros */
    if (m <= 0) continue;
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < m; ++j) {
      const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
      for (int tid = 0; tid < S; ++tid) {
        int besti = 0;
        float best = -1.f;
        for (int k = tid; k < N; k += S) {
          const float x2 = p[k * 3 + 0], y2 = p[k * 3 + 1], z2 = p[k * 3 + 2];
          const float mag = sumsq3_ref(x2, y2, z2);
          if ((double)mag <= 1e-3) continue; /* double literal in the source */
          const float d = sumsq3_ref(x2 - x1, y2 - y1, z2 - z1);
          const float d2 = d < temp[k] ? d : temp[k];
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dv[tid] = best;
        di[tid] = besti;
      }
      for (int stride = S / 2; stride >= 1; stride >>= 1) {
        for (int tid = 0; tid < stride; ++tid) {
          const float v1 = dv[tid], v2 = dv[tid + stride];
          const int i1 = di[tid], i2 = di[tid + stride];
          dv[tid] = v1 > v2 ? v1 : v2;
          di[tid] = v2 > v1 ? i2 : i1;
        }
      }
      old = di[0];
      out[j] = old;
    }
  }
  free(temp); free(dv); free(di);
  return SO_OK;
}

/* ------------------------------------------------------------------ gather / group */
int so_gather(const float *points, const int *idx, int B, int C, int N, int m, float *out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < m; ++j)
        out[((size_t)b * C + c) * m + j] = points[((size_t)b * C + c) * N + idx[(size_t)b * m + j]];
  return SO_OK;
}

int so_gather_grad(const float *grad_out, const int *idx, int B, int C, int N, int m, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * N);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < m; ++j)
        grad_points[((size_t)b * C + c) * N + idx[(size_t)b * m + j]] += grad_out[((size_t)b * C + c) * m + j];
  return SO_OK;
}

int so_group(const float *points, const int *idx, int B, int C, int N, int np, int ns, float *out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < np; ++j)
        for (int k = 0; k < ns; ++k)
          out[(((size_t)b * C + c) * np + j) * ns + k] =
              points[((size_t)b * C + c) * N + idx[((size_t)b * np + j) * ns + k]];
  return SO_OK;
}

int so_group_grad(const float *grad_out, const int *idx, int B, int C, int N, int np, int ns, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * N);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < np; ++j)
        for (int k = 0; k < ns; ++k)
          grad_points[((size_t)b * C + c) * N + idx[((size_t)b * np + j) * ns + k]] +=
              grad_out[(((size_t)b * C + c) * np + j) * ns + k];
  return SO_OK;
}

/* ------------------------------------------------------------------ ball query */
int so_ball_query(const float *new_xyz, const float *xyz, int B, int N, int m, float radius, int nsample,
                  int *idx, int *counts) {
  const float radius2 = radius * radius;
  memset(idx, 0, sizeof(int) * (size_t)B * m * nsample);
  memset(counts, 0, sizeof(int) * (size_t)B * m);
  for (int b = 0; b < B; ++b) {
    const float *q = new_xyz + (size_t)b * m * 3;
    const float *p = xyz + (size_t)b * N * 3;
    for (int j = 0; j < m; ++j) {
      int *row = idx + ((size_t)b * m + j) * nsample;
      const float qx = q[j * 3 + 0], qy = q[j * 3 + 1], qz = q[j * 3 + 2];
      int cnt = 0;
      for (int k = 0; k < N && cnt < nsample; ++k) {
        const float d2 = sumsq3_ref(qx - p[k * 3 + 0], qy - p[k * 3 + 1], qz - p[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) row[l] = k;
          row[cnt] = k;
          ++cnt;
          counts[(size_t)b * m + j] = cnt;
        }
      }
    }
  }
  return SO_OK;
}

/* ------------------------------------------------------------------ three_nn / three_interpolate */
int so_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2, int *idx) {
  for (int b = 0; b < B; ++b) {
    const float *u = unknown + (size_t)b * n * 3;
    const float *kn = known + (size_t)b * m * 3;
    for (int j = 0; j < n; ++j) {
      const float ux = u[j * 3 + 0], uy = u[j * 3 + 1], uz = u[j * 3 + 2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int b1 = 0, b2 = 0, b3 = 0;
      for (int k = 0; k < m; ++k) {
        const float d = sumsq3_ref(ux - kn[k * 3 + 0], uy - kn[k * 3 + 1], uz - kn[k * 3 + 2]);
        if (d < best1) {
          best3 = best2; b3 = b2; best2 = best1; b2 = b1; best1 = d; b1 = k;
        } else if (d < best2) {
          best3 = best2; b3 = b2; best2 = d; b2 = k;
        } else if (d < best3) {
          best3 = d; b3 = k;
        }
      }
      float *dd = dist2 + ((size_t)b * n + j) * 3;
      int *ii = idx + ((size_t)b * n + j) * 3;
      dd[0] = (float)best1; dd[1] = (float)best2; dd[2] = (float)best3;
      ii[0] = b1; ii[1] = b2; ii[2] = b3;
    }
  }
  return SO_OK;
}

int so_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C, int m, int n,
                         float *out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float *p = points + ((size_t)b * C + c) * m;
      for (int j = 0; j < n; ++j) {
        const int *ii = idx + ((size_t)b * n + j) * 3;
        const float *w = weight + ((size_t)b * n + j) * 3;
        float t = p[ii[1]] * w[1];
        t = fmaf(p[ii[0]], w[0], t);
        t = fmaf(p[ii[2]], w[2], t);
        out[((size_t)b * C + c) * n + j] = t;
      }
    }
  return SO_OK;
}

int so_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B, int C, int n,
                              int m, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * m);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < n; ++j) {
        const int *ii = idx + ((size_t)b * n + j) * 3;
        const float *w = weight + ((size_t)b * n + j) * 3;
        const float g = grad_out[((size_t)b * C + c) * n + j];
        float *gp = grad_points + ((size_t)b * C + c) * m;
        gp[ii[0]] += g * w[0];
        gp[ii[1]] += g * w[1];
        gp[ii[2]] += g * w[2];
      }
  return SO_OK;
}

/* ------------------------------------------------------------------ pytorch3d-style kNN */
/* Brute force, squared L2 accumulated over d = 0..D-1 (`dist += diff*diff`, FMA-contracted on the GPU),
 * K smallest, ascending; equal distances keep scan order (smaller index first).  Missing neighbours
 * (K > lengths2) are left 0.  idx is int64 like pytorch3d's. */
int so_knn(const float *p1, const float *p2, int B, int P1, int P2, int D, const int64_t *lengths1,
           const int64_t *lengths2, int K, float *dists, int64_t *idx) {
  if (K <= 0 || D <= 0) return SO_EINVAL;
  float *bd = (float *)malloc(sizeof(float) * (size_t)K);
  int64_t *bi = (int64_t *)malloc(sizeof(int64_t) * (size_t)K);
  if (!bd || !bi) { free(bd); free(bi); return SO_EINVAL; }
  memset(dists, 0, sizeof(float) * (size_t)B * P1 * K);
  memset(idx, 0, sizeof(int64_t) * (size_t)B * P1 * K);
  for (int b = 0; b < B; ++b) {
    const int64_t l1 = lengths1 ? lengths1[b] : P1;
    const int64_t l2 = lengths2 ? lengths2[b] : P2;
    for (int64_t i = 0; i < l1; ++i) {
      const float *q = p1 + ((size_t)b * P1 + i) * D;
      int cnt = 0;
      for (int64_t j = 0; j < l2; ++j) {
        const float *r = p2 + ((size_t)b * P2 + j) * D;
        float dist = 0.f;
        for (int d = 0; d < D; ++d) {
          const float diff = q[d] - r[d];
          dist = fmaf(diff, diff, dist);
        }
        if (cnt < K || dist < bd[cnt - 1]) {
          int pos = cnt < K ? cnt : K - 1;
          while (pos > 0 && dist < bd[pos - 1]) {
            bd[pos] = bd[pos - 1];
            bi[pos] = bi[pos - 1];
            --pos;
          }
          bd[pos] = dist;
          bi[pos] = j;
          if (cnt < K) ++cnt;
        }
      }
      for (int k = 0; k < cnt; ++k) {
        dists[((size_t)b * P1 + i) * K + k] = bd[k];
        idx[((size_t)b * P1 + i) * K + k] = bi[k];
      }
    }
  }
  free(bd); free(bi);
  return SO_OK;
}

/* ------------------------------------------------------------------ pytorch3d-style FPS */
/* closest distances start at +inf, first pick = start_idx[b], each following pick = first arg-max of the
 * running min distance (data_utils/points_sampling.py:73-107); slots beyond min(K_b, length_b) stay -1. */
int so_fps_p3d(const float *points, int B, int P, int D, const int64_t *lengths, const int64_t *K,
               const int64_t *start_idx, int maxK, int64_t *idx) {
  float *md = (float *)malloc(sizeof(float) * (size_t)(P > 0 ? P : 1));
  if (!md) return SO_EINVAL;
  for (int b = 0; b < B; ++b) {
    const int64_t len = lengths ? lengths[b] : P;
    const int64_t kb = K ? K[b] : maxK;
    int64_t *out = idx + (size_t)b * maxK;
    for (int j = 0; j < maxK; ++j) out[j] = -1;
    if (len <= 0 || kb <= 0) continue;
    const float *p = points + (size_t)b * P * D;
    for (int64_t i = 0; i < len; ++i) md[i] = INFINITY;
    int64_t sel = start_idx ? start_idx[b] : 0;
    out[0] = sel;
    const int64_t kn = kb < len ? kb : len;
    for (int64_t j = 1; j < kn; ++j) {
      const float *ps = p + (size_t)sel * D;
      float best = -1.f;
      int64_t besti = 0;
      for (int64_t i = 0; i < len; ++i) {
        float dist = 0.f;
        for (int d = 0; d < D; ++d) {
          const float diff = ps[d] - p[(size_t)i * D + d];
          dist = fmaf(diff, diff, dist);
        }
        const float v = dist < md[i] ? dist : md[i];
        md[i] = v;
        if (v > best) { best = v; besti = i; }
      }
      sel = besti;
      out[j] = sel;
    }
  }
  free(md);
  return SO_OK;
}
