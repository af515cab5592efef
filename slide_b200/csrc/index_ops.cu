// Point-set index ops for sm_100a: FPS, ball query, kNN, gather/group, three_nn/three_interpolate.
//
// These replace the nine kernels of the reference's pointnet2_ops._ext (EXT/src/*.cu) and the pytorch3d
// ops it depends on.  They are integer / byte-movement work bound by latency (FPS) or HBM/L2 bandwidth
// (scans and gathers): coalesced loads, shared-memory staging, warp ballots and REDUX reductions -- no
// tensor cores.  Index results are bit-exact with the reference, including its tie rules.
#include <math.h>

#include <stdlib.h>

#include "common.cuh"

namespace slide {

// =====================================================================================================
// Furthest point sampling
// =====================================================================================================
// One CTA per cloud.  Every thread keeps PPT points and their running min distance in registers for the
// whole m-1 iteration chain; one iteration = PPT distance updates, two REDUX warp reductions, ONE block
// barrier (double-buffered warp slots) and a redundant second-level reduction in every warp.  The
// reference does the same chain with a strided global-memory scan and a 10-barrier shared-memory tree
// (EXT/src/sampling_gpu.cu:89-170).
//
// Tie rule.  The reference's result on equal distances is a property of its launch shape: block size
// S = opt_n_threads(N), thread tid scans k = tid, tid+S, ... keeping the first strict maximum, and the tree
// keeps the left slot.  The global winner among equal maxima is therefore the point with the smallest
// (bitreverse_log2S(k mod S), k div S).  We reduce the pair (distance bits, that rank) so that any
// thread->point mapping reproduces it.
//
// MODE 0: pointnet2_ops._ext semantics (start 0, |p|^2 <= 1e-3 skipped, init 1e10, i32 output).
// MODE 1: pytorch3d semantics (start index given, init +inf, lowest index wins ties, i64 output, -1 pad).
// MODE 2: MODE 1 with i32 start indices / output and no per-cloud lengths (used inside slide programs).
// `ldx` is the row stride of xyz in floats (3 for a packed cloud).
constexpr int FPS_MAX_RESIDENT_N = 16384;  // 1024 threads x 16 points in registers

template <int PPT, int MODE>
__global__ void __launch_bounds__(1024) fps_kernel(const float *__restrict__ xyz, int ldx, int N, int m, int s_log2,
                                                   int nb_log2, const int64_t *__restrict__ lengths,
                                                   const int64_t *__restrict__ Ks,
                                                   const void *__restrict__ start_raw, void *out_raw) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint2 slots[2][32];
  const int b = blockIdx.x;
  const int T = blockDim.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const float *p = xyz + (size_t)b * N * ldx;
  const int64_t *start_idx = (const int64_t *)start_raw;

  int len = N, kn = m;
  if (MODE == 1) {
    if (lengths) len = (int)lengths[b];
    if (Ks) kn = (int)Ks[b];
    if (kn > len) kn = len;
    if (kn > m) kn = m;
  }

  float px[PPT], py[PPT], pz[PPT], md[PPT];
  uint32_t rank[PPT];
  bool ok[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = tid + i * T;
    const bool in = k < len;
    px[i] = in ? p[(size_t)k * ldx + 0] : 0.f;
    py[i] = in ? p[(size_t)k * ldx + 1] : 0.f;
    pz[i] = in ? p[(size_t)k * ldx + 2] : 0.f;
    if (MODE == 0) {
      const float mag = sumsq3_ref(px[i], py[i], pz[i]);
      ok[i] = in && !((double)mag <= 1e-3);
      md[i] = 1e10f;
      const uint32_t lo = (uint32_t)k & ((1u << s_log2) - 1u);
      const uint32_t rev = s_log2 ? (__brev(lo) >> (32 - s_log2)) : 0u;
      rank[i] = (rev << nb_log2) | ((uint32_t)k >> s_log2);
    } else {
      ok[i] = in;
      md[i] = INFINITY;
      rank[i] = (uint32_t)k;
    }
  }

  int old = 0;
  if (MODE == 1 && start_idx) old = (int)start_idx[b];
  if (MODE == 2 && start_raw) old = ((const int *)start_raw)[b];
  int *out32 = (int *)out_raw + (size_t)b * m;
  long long *out64 = (long long *)out_raw + (size_t)b * m;
  if (tid == 0) {
    if (MODE == 0) {
      if (m > 0) out32[0] = 0;
    } else if (MODE == 2) {
      if (m > 0) out32[0] = old;
    } else {
      for (int j = kn > 0 ? kn : 0; j < m; ++j) out64[j] = -1;
      if (kn > 0) out64[0] = old;
    }
  }

  for (int j = 1; j < kn; ++j) {
    const float x1 = __ldg(p + (size_t)old * ldx + 0), y1 = __ldg(p + (size_t)old * ldx + 1),
                z1 = __ldg(p + (size_t)old * ldx + 2);
    uint32_t bhi = 0, brk = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      float d;
      if (MODE == 0)
        d = sumsq3_ref(px[i] - x1, py[i] - y1, pz[i] - z1);
      else
        d = sumsq3_p3d(x1 - px[i], y1 - py[i], z1 - pz[i]);
      const float v = fminf(d, md[i]);
      if (ok[i]) {
        md[i] = v;
        const uint32_t hi = __float_as_uint(v) + 1u;  // v >= 0: bit pattern is monotone; 0 means "none"
        if (hi > bhi || (hi == bhi && rank[i] < brk)) {
          bhi = hi;
          brk = rank[i];
        }
      }
    }
    uint32_t mx = __reduce_max_sync(0xffffffffu, bhi);
    uint32_t mr = __reduce_min_sync(0xffffffffu, bhi == mx ? brk : 0xffffffffu);
    if (lane == 0) slots[j & 1][warp] = make_uint2(mx, mr);
    __syncthreads();
    uint2 s = lane < nwarps ? slots[j & 1][lane] : make_uint2(0u, 0xffffffffu);
    mx = __reduce_max_sync(0xffffffffu, s.x);
    mr = __reduce_min_sync(0xffffffffu, s.x == mx ? s.y : 0xffffffffu);
    if (mx == 0u) {
      old = 0;  // no candidate anywhere: the reference's tree returns slot 0's initial index
    } else if (MODE == 0) {
      const uint32_t rev = mr >> nb_log2;
      const uint32_t lo = s_log2 ? (__brev(rev) >> (32 - s_log2)) : 0u;
      old = (int)((((mr & ((1u << nb_log2) - 1u))) << s_log2) | lo);
    } else {
      old = (int)mr;
    }
    if (tid == 0) {
      if (MODE == 1)
        out64[j] = old;
      else
        out32[j] = old;
    }
  }
}

// Clouds too large for the register-resident kernel (N > 16384): same chain, same tie rule, with the running
// distances in a caller-provided scratch `temp` f32[B,N] (the reference's own `tmp` tensor, sampling.cpp:74-76) and the
// points re-read through L1/L2 every round.  Latency-bound like the reference's kernel, but with 1024 threads per cloud,
// REDUX reductions and one barrier per round.
template <int MODE>
__global__ void __launch_bounds__(1024) fps_global_kernel(const float *__restrict__ xyz, int N, int m, int s_log2,
                                                          int nb_log2, const int64_t *__restrict__ lengths,
                                                          const int64_t *__restrict__ Ks,
                                                          const int64_t *__restrict__ start_idx,
                                                          float *__restrict__ temp, void *out_raw) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint2 slots[2][32];
  const int b = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const float *p = xyz + (size_t)b * N * 3;
  float *md = temp + (size_t)b * N;
  int len = N, kn = m;
  if (MODE == 1) {
    if (lengths) len = (int)lengths[b];
    if (Ks) kn = (int)Ks[b];
    if (kn > len) kn = len;
    if (kn > m) kn = m;
  }
  for (int k = tid; k < len; k += T) md[k] = MODE == 0 ? 1e10f : INFINITY;
  int old = 0;
  if (MODE == 1 && start_idx) old = (int)start_idx[b];
  int *out32 = (int *)out_raw + (size_t)b * m;
  long long *out64 = (long long *)out_raw + (size_t)b * m;
  if (tid == 0) {
    if (MODE == 0) {
      if (m > 0) out32[0] = 0;
    } else {
      for (int j = kn > 0 ? kn : 0; j < m; ++j) out64[j] = -1;
      if (kn > 0) out64[0] = old;
    }
  }
  for (int j = 1; j < kn; ++j) {
    const float x1 = __ldg(p + (size_t)old * 3 + 0), y1 = __ldg(p + (size_t)old * 3 + 1), z1 = __ldg(p + (size_t)old * 3 + 2);
    uint32_t bhi = 0, brk = 0xffffffffu;
    for (int k = tid; k < len; k += T) {
      const float x = __ldg(p + (size_t)k * 3 + 0), y = __ldg(p + (size_t)k * 3 + 1), z = __ldg(p + (size_t)k * 3 + 2);
      float d;
      uint32_t rk;
      if (MODE == 0) {
        if ((double)sumsq3_ref(x, y, z) <= 1e-3) continue;
        d = sumsq3_ref(x - x1, y - y1, z - z1);
        const uint32_t lo = (uint32_t)k & ((1u << s_log2) - 1u);
        rk = ((s_log2 ? (__brev(lo) >> (32 - s_log2)) : 0u) << nb_log2) | ((uint32_t)k >> s_log2);
      } else {
        d = sumsq3_p3d(x1 - x, y1 - y, z1 - z);
        rk = (uint32_t)k;
      }
      const float v = fminf(d, md[k]);
      md[k] = v;
      const uint32_t hi = __float_as_uint(v) + 1u;
      if (hi > bhi || (hi == bhi && rk < brk)) {
        bhi = hi;
        brk = rk;
      }
    }
    uint32_t mx = __reduce_max_sync(0xffffffffu, bhi);
    uint32_t mr = __reduce_min_sync(0xffffffffu, bhi == mx ? brk : 0xffffffffu);
    if (lane == 0) slots[j & 1][warp] = make_uint2(mx, mr);
    __syncthreads();
    uint2 sl = lane < nwarps ? slots[j & 1][lane] : make_uint2(0u, 0xffffffffu);
    mx = __reduce_max_sync(0xffffffffu, sl.x);
    mr = __reduce_min_sync(0xffffffffu, sl.x == mx ? sl.y : 0xffffffffu);
    if (mx == 0u) {
      old = 0;
    } else if (MODE == 0) {
      const uint32_t rev = mr >> nb_log2;
      const uint32_t lo = s_log2 ? (__brev(rev) >> (32 - s_log2)) : 0u;
      old = (int)((((mr & ((1u << nb_log2) - 1u))) << s_log2) | lo);
    } else {
      old = (int)mr;
    }
    if (tid == 0) {
      if (MODE == 1)
        out64[j] = old;
      else
        out32[j] = old;
    }
  }
}

static int host_opt_n_threads(int work_size) {
  // EXT/include/cuda_utils.h:15-19 (double-precision log, truncation) -- decides the reference's tie rule.
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

static int ilog2_ceil(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

template <int MODE>
static int launch_fps(const float *xyz, int ldx, int B, int N, int m, const int64_t *lengths, const int64_t *Ks,
                      const void *start, void *out, cudaStream_t st) {
  if (B == 0 || m == 0) return SLIDE_OK;
  if (N > FPS_MAX_RESIDENT_N) return SLIDE_ERR_UNSUPPORTED;  // use the *_ws entry points (scratch-backed kernel)
  int T = ((N + 31) / 32) * 32;
  if (T > 1024) T = 1024;
  const int ppt = ceil_div(N, T);
  const int S = host_opt_n_threads(N);
  const int s_log2 = ilog2_ceil(S);
  const int nb_log2 = ilog2_ceil(ceil_div(N, S));
  if (s_log2 + nb_log2 > 31) return SLIDE_ERR_UNSUPPORTED;
#define FPS_CASE(P)                                                                                      \
  launch_k(fps_kernel<P, MODE>, B, T, 0, st, xyz, ldx, N, m, s_log2, nb_log2, lengths, Ks, start, out);        \
  break;
  switch (ppt <= 1 ? 1 : ppt <= 2 ? 2 : ppt <= 4 ? 4 : ppt <= 8 ? 8 : 16) {
    case 1: FPS_CASE(1)
    case 2: FPS_CASE(2)
    case 4: FPS_CASE(4)
    case 8: FPS_CASE(8)
    default: FPS_CASE(16)
  }
#undef FPS_CASE
  return after_launch();
}

// =====================================================================================================
// gather / group (pure index gathers, channel-major like the reference: points (B,C,N))
// =====================================================================================================
// out[b,c,j] = points[b,c,idx[b,j]] for j in [0, M) where M = m (gather) or npoint*nsample (group).
// One thread per output column j, a few channels per thread: writes are fully coalesced, the index row is
// read once per thread and reused across channels.
__global__ void gather_cols_kernel(const float *__restrict__ points, const int *__restrict__ idx, int C, int N,
                                   int M, int c_per_block, float *__restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int a = idx[(size_t)b * M + j];
  const int c0 = blockIdx.y * c_per_block;
  const int c1 = min(C, c0 + c_per_block);
  const float *src = points + ((size_t)b * C + c0) * N + a;
  float *dst = out + ((size_t)b * C + c0) * M + j;
  for (int c = c0; c < c1; ++c, src += N, dst += M) *dst = __ldg(src);
}

__global__ void scatter_cols_grad_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, int C,
                                         int N, int M, int c_per_block, float *__restrict__ grad_points) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int a = idx[(size_t)b * M + j];
  const int c0 = blockIdx.y * c_per_block;
  const int c1 = min(C, c0 + c_per_block);
  for (int c = c0; c < c1; ++c)
    atomicAdd(grad_points + ((size_t)b * C + c) * N + a, grad_out[((size_t)b * C + c) * M + j]);
}

static int launch_gather_cols(const float *points, const int *idx, int B, int C, int N, int M, float *out,
                              bool grad, cudaStream_t st) {
  if (B == 0 || C == 0 || M == 0) return SLIDE_OK;
  if (B > 65535) return SLIDE_ERR_UNSUPPORTED;
  const int threads = 256;
  int cpb = 8;
  int cy = ceil_div(C, cpb);
  if (cy > 65535) {
    cpb = ceil_div(C, 65535);
    cy = ceil_div(C, cpb);
  }
  dim3 grid(ceil_div(M, threads), cy, B);
  if (!grad)
    launch_k(gather_cols_kernel, grid, threads, 0, st, points, idx, C, N, M, cpb, out);
  else
    launch_k(scatter_cols_grad_kernel, grid, threads, 0, st, points, idx, C, N, M, cpb, out);
  return after_launch();
}

// =====================================================================================================
// Ball query: one warp per query, ballot over 32 candidate points at a time
// =====================================================================================================
// The cloud is staged through shared memory in tiles (coalesced global reads, each point read once per
// CTA instead of once per query); each lane tests one point, __ballot_sync gives the hit mask in ascending
// point order, and lanes write their hit at cnt + popc(mask below me).  A warp stops scanning as soon as it
// has nsample hits, a CTA as soon as all of its warps have.
constexpr int BQ_WARPS = 8;
constexpr int BQ_TILE = 1024;

__global__ void __launch_bounds__(BQ_WARPS * 32) ball_query_kernel(const float *__restrict__ new_xyz,
                                                                   const float *__restrict__ xyz, int N, int m,
                                                                   float radius2, int nsample,
                                                                   int *__restrict__ idx, int *__restrict__ counts) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[BQ_TILE * 3];
  __shared__ int live_warps;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * BQ_WARPS + warp;
  const bool active = j < m;
  const float *q = new_xyz + ((size_t)b * m + (active ? j : 0)) * 3;
  const float qx = q[0], qy = q[1], qz = q[2];
  const float *p = xyz + (size_t)b * N * 3;
  int *row = idx + ((size_t)b * m + j) * nsample;
  int cnt = 0, first = 0;
  bool done = !active;
  for (int base = 0; base < N; base += BQ_TILE) {
    const int tn = min(BQ_TILE, N - base);
    __syncthreads();
    if (threadIdx.x == 0) live_warps = 0;
    for (int i = threadIdx.x; i < tn * 3; i += blockDim.x) tile[i] = p[(size_t)base * 3 + i];
    __syncthreads();
    if (!done) {
      for (int k0 = 0; k0 < tn; k0 += 32) {
        const int k = k0 + lane;
        bool hit = false;
        if (k < tn) {
          const float d2 = sumsq3_ref(qx - tile[k * 3 + 0], qy - tile[k * 3 + 1], qz - tile[k * 3 + 2]);
          hit = d2 < radius2;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask) {
          if (cnt == 0) first = base + k0 + __ffs(mask) - 1;
          const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
          if (hit && pos < nsample) row[pos] = base + k;
          cnt += __popc(mask);
          if (cnt >= nsample) {
            cnt = nsample;
            done = true;
            break;
          }
        }
      }
      if (!done && lane == 0) atomicAdd(&live_warps, 1);
    }
    __syncthreads();
    if (live_warps == 0) break;
  }
  if (active) {
    // pad with the first hit (all zeros when there was none), like the reference's pre-fill + zero init
    for (int l = cnt + lane; l < nsample; l += 32) row[l] = first;
    if (lane == 0) counts[(size_t)b * m + j] = cnt;
  }
}

// =====================================================================================================
// three_nn / three_interpolate
// =====================================================================================================
constexpr int NN_TILE = 1024;

__global__ void __launch_bounds__(128) three_nn_kernel(const float *__restrict__ unknown,
                                                       const float *__restrict__ known, int n, int m,
                                                       float *__restrict__ dist2, int *__restrict__ idx) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[NN_TILE * 3];
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = j < n;
  const float *u = unknown + ((size_t)b * n + (active ? j : 0)) * 3;
  const float ux = u[0], uy = u[1], uz = u[2];
  const float *kp = known + (size_t)b * m * 3;
  // the reference keeps doubles initialised to 1e40; every real distance is a float, so +inf is equivalent
  float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int base = 0; base < m; base += NN_TILE) {
    const int tn = min(NN_TILE, m - base);
    __syncthreads();
    for (int i = threadIdx.x; i < tn * 3; i += blockDim.x) tile[i] = kp[(size_t)base * 3 + i];
    __syncthreads();
    if (active) {
      for (int k = 0; k < tn; ++k) {
        const float d = sumsq3_ref(ux - tile[k * 3 + 0], uy - tile[k * 3 + 1], uz - tile[k * 3 + 2]);
        if (d < b1) {
          b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = base + k;
        } else if (d < b2) {
          b3 = b2; i3 = i2; b2 = d; i2 = base + k;
        } else if (d < b3) {
          b3 = d; i3 = base + k;
        }
      }
    }
  }
  if (active) {
    float *dd = dist2 + ((size_t)b * n + j) * 3;
    int *ii = idx + ((size_t)b * n + j) * 3;
    dd[0] = b1; dd[1] = b2; dd[2] = b3;
    ii[0] = i1; ii[1] = i2; ii[2] = i3;
  }
}

__global__ void three_interpolate_kernel(const float *__restrict__ points, const int *__restrict__ idx,
                                         const float *__restrict__ weight, int C, int m, int n, int c_per_block,
                                         float *__restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int *ii = idx + ((size_t)b * n + j) * 3;
  const float *w = weight + ((size_t)b * n + j) * 3;
  const int i1 = ii[0], i2 = ii[1], i3 = ii[2];
  const float w1 = w[0], w2 = w[1], w3 = w[2];
  const int c0 = blockIdx.y * c_per_block, c1 = min(C, c0 + c_per_block);
  for (int c = c0; c < c1; ++c) {
    const float *p = points + ((size_t)b * C + c) * m;
    float t = __fmul_rn(__ldg(p + i2), w2);  // reference SASS: FMUL (2nd term), FFMA (1st), FFMA (3rd)
    t = __fmaf_rn(__ldg(p + i1), w1, t);
    t = __fmaf_rn(__ldg(p + i3), w3, t);
    out[((size_t)b * C + c) * n + j] = t;
  }
}

__global__ void three_interpolate_grad_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx,
                                              const float *__restrict__ weight, int C, int n, int m,
                                              int c_per_block, float *__restrict__ grad_points) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int *ii = idx + ((size_t)b * n + j) * 3;
  const float *w = weight + ((size_t)b * n + j) * 3;
  const int c0 = blockIdx.y * c_per_block, c1 = min(C, c0 + c_per_block);
  for (int c = c0; c < c1; ++c) {
    const float g = grad_out[((size_t)b * C + c) * n + j];
    float *gp = grad_points + ((size_t)b * C + c) * m;
    atomicAdd(gp + ii[0], g * w[0]);
    atomicAdd(gp + ii[1], g * w[1]);
    atomicAdd(gp + ii[2], g * w[2]);
  }
}

// =====================================================================================================
// kNN (pytorch3d knn_points semantics, D = 3)
// =====================================================================================================
// One thread per query point; the reference cloud streams through shared memory in tiles.  Each thread
// keeps its K best (distance, index) pairs sorted; a candidate is inserted only if it beats the current
// K-th best (strict <), so equal distances stay in ascending index order.
constexpr int KNN_TILE = 1024;
constexpr int KNN_MAXK = 64;

template <int KCAP>
__global__ void __launch_bounds__(128) knn_kernel(const float *__restrict__ p1, const float *__restrict__ p2,
                                                  int P1, int P2, const int64_t *__restrict__ lengths1,
                                                  const int64_t *__restrict__ lengths2, int K,
                                                  float *__restrict__ dists, long long *__restrict__ idx) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[KNN_TILE * 3];
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int l1 = lengths1 ? (int)lengths1[b] : P1;
  const int l2 = lengths2 ? (int)lengths2[b] : P2;
  const bool active = i < l1 && i < P1;
  const float *q = p1 + ((size_t)b * P1 + (i < P1 ? i : 0)) * 3;
  const float qx = q[0], qy = q[1], qz = q[2];
  const float *r = p2 + (size_t)b * P2 * 3;
  float bd[KCAP];
  int bi[KCAP];
  int cnt = 0;
  float worst = INFINITY;
  for (int base = 0; base < l2; base += KNN_TILE) {
    const int tn = min(KNN_TILE, l2 - base);
    __syncthreads();
    for (int t = threadIdx.x; t < tn * 3; t += blockDim.x) tile[t] = r[(size_t)base * 3 + t];
    __syncthreads();
    if (!active) continue;
    for (int k = 0; k < tn; ++k) {
      const float d = sumsq3_p3d(qx - tile[k * 3 + 0], qy - tile[k * 3 + 1], qz - tile[k * 3 + 2]);
      if (cnt < K || d < worst) {
        int pos = cnt < K ? cnt : K - 1;
        while (pos > 0 && d < bd[pos - 1]) {
          bd[pos] = bd[pos - 1];
          bi[pos] = bi[pos - 1];
          --pos;
        }
        bd[pos] = d;
        bi[pos] = base + k;
        if (cnt < K) ++cnt;
        if (cnt == K) worst = bd[K - 1];
      }
    }
  }
  if (i < P1) {
    float *dd = dists + ((size_t)b * P1 + i) * K;
    long long *ii = idx + ((size_t)b * P1 + i) * K;
    for (int k = 0; k < K; ++k) {
      const bool have = active && k < cnt;
      dd[k] = have ? bd[k] : 0.f;
      ii[k] = have ? bi[k] : 0;
    }
  }
}

// entry points for the program executor (program.cu)
template <int MODE>
static int launch_fps_global(const float *xyz, int B, int N, int m, const int64_t *lengths, const int64_t *Ks,
                             const int64_t *start, float *temp, void *out, cudaStream_t st) {
  if (B == 0 || m == 0) return SLIDE_OK;
  if (!temp) return SLIDE_ERR_INVALID;
  const int S = host_opt_n_threads(N);
  const int s_log2 = ilog2_ceil(S);
  const int nb_log2 = ilog2_ceil(ceil_div(N, S));
  if (s_log2 + nb_log2 > 31) return SLIDE_ERR_UNSUPPORTED;
  launch_k(fps_global_kernel<MODE>, B, 1024, 0, st, xyz, N, m, s_log2, nb_log2, lengths, Ks, start, temp, out);
  return after_launch();
}

int program_fps(int mode, const float *xyz, int ldx, int B, int N, int m, const int *start, int *out, cudaStream_t st) {
  if (mode == 0) return launch_fps<0>(xyz, ldx, B, N, m, nullptr, nullptr, nullptr, out, st);
  return launch_fps<2>(xyz, ldx, B, N, m, nullptr, nullptr, start, out, st);
}

long long g_launch_count = 0;
static int pdl_from_env() {
  const char *e = getenv("SLIDE_PDL");
  return e ? atoi(e) != 0 : 0;
}
int g_pdl_enabled = pdl_from_env();
static thread_local cudaError_t g_last_error = cudaSuccess;
void set_cuda_error(cudaError_t e) { g_last_error = e; }

}  // namespace slide

using namespace slide;

extern "C" {

int slide_abi_version(void) { return 1; }
const char *slide_last_cuda_error(void) { return cudaGetErrorString(g_last_error); }
long long slide_launch_count(void) { return g_launch_count; }
void slide_reset_launch_count(void) { g_launch_count = 0; }

int slide_furthest_point_sampling(const float *xyz, int B, int N, int m, int *idx, slide_stream_t stream) {
  if (!xyz || !idx || B < 0 || N <= 0 || m < 0) return SLIDE_ERR_INVALID;
  return launch_fps<0>(xyz, 3, B, N, m, nullptr, nullptr, nullptr, idx, (cudaStream_t)stream);
}

int slide_sample_farthest_points(const float *points, int B, int P, int D, const int64_t *lengths,
                                 const int64_t *K, const int64_t *start_idx, int maxK, int64_t *idx,
                                 slide_stream_t stream) {
  if (!points || !idx || B < 0 || P <= 0 || maxK < 0) return SLIDE_ERR_INVALID;
  if (D != 3) return SLIDE_ERR_UNSUPPORTED;
  return launch_fps<1>(points, 3, B, P, maxK, lengths, K, start_idx, idx, (cudaStream_t)stream);
}

int slide_furthest_point_sampling_ws(const float *xyz, int B, int N, int m, int *idx, float *temp,
                                     slide_stream_t stream) {
  if (!xyz || !idx || B < 0 || N <= 0 || m < 0) return SLIDE_ERR_INVALID;
  if (N <= FPS_MAX_RESIDENT_N) return launch_fps<0>(xyz, 3, B, N, m, nullptr, nullptr, nullptr, idx, (cudaStream_t)stream);
  return launch_fps_global<0>(xyz, B, N, m, nullptr, nullptr, nullptr, temp, idx, (cudaStream_t)stream);
}

int slide_sample_farthest_points_ws(const float *points, int B, int P, int D, const int64_t *lengths,
                                    const int64_t *K, const int64_t *start_idx, int maxK, int64_t *idx, float *temp,
                                    slide_stream_t stream) {
  if (!points || !idx || B < 0 || P <= 0 || maxK < 0) return SLIDE_ERR_INVALID;
  if (D != 3) return SLIDE_ERR_UNSUPPORTED;
  if (P <= FPS_MAX_RESIDENT_N)
    return launch_fps<1>(points, 3, B, P, maxK, lengths, K, start_idx, idx, (cudaStream_t)stream);
  return launch_fps_global<1>(points, B, P, maxK, lengths, K, start_idx, temp, idx, (cudaStream_t)stream);
}

int slide_fps_resident_max_points(void) { return FPS_MAX_RESIDENT_N; }

int slide_gather_points(const float *points, const int *idx, int B, int C, int N, int m, float *out,
                        slide_stream_t stream) {
  if (!points || !idx || !out || B < 0 || C < 0 || N <= 0 || m < 0) return SLIDE_ERR_INVALID;
  return launch_gather_cols(points, idx, B, C, N, m, out, false, (cudaStream_t)stream);
}

int slide_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int m,
                             float *grad_points, slide_stream_t stream) {
  if (!grad_out || !idx || !grad_points || B < 0 || C < 0 || N <= 0 || m < 0) return SLIDE_ERR_INVALID;
  return launch_gather_cols(grad_out, idx, B, C, N, m, grad_points, true, (cudaStream_t)stream);
}

int slide_group_points(const float *points, const int *idx, int B, int C, int N, int npoint, int nsample,
                       float *out, slide_stream_t stream) {
  if (!points || !idx || !out || B < 0 || C < 0 || N <= 0 || npoint < 0 || nsample < 0) return SLIDE_ERR_INVALID;
  return launch_gather_cols(points, idx, B, C, N, npoint * nsample, out, false, (cudaStream_t)stream);
}

int slide_group_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int npoint,
                            int nsample, float *grad_points, slide_stream_t stream) {
  if (!grad_out || !idx || !grad_points || B < 0 || C < 0 || N <= 0 || npoint < 0 || nsample < 0)
    return SLIDE_ERR_INVALID;
  return launch_gather_cols(grad_out, idx, B, C, N, npoint * nsample, grad_points, true, (cudaStream_t)stream);
}

int slide_ball_query(const float *new_xyz, const float *xyz, int B, int N, int m, float radius, int nsample,
                     int *idx, int *counts, slide_stream_t stream) {
  if (!new_xyz || !xyz || !idx || !counts || B < 0 || N <= 0 || m < 0 || nsample <= 0) return SLIDE_ERR_INVALID;
  if (B == 0 || m == 0) return SLIDE_OK;
  if (B > 65535) return SLIDE_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(m, BQ_WARPS), B);
  launch_k(ball_query_kernel, grid, BQ_WARPS * 32, 0, (cudaStream_t)stream, new_xyz, xyz, N, m, radius * radius,
                                                                      nsample, idx, counts);
  return after_launch();
}

int slide_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2, int *idx,
                   slide_stream_t stream) {
  if (!unknown || !known || !dist2 || !idx || B < 0 || n < 0 || m < 0) return SLIDE_ERR_INVALID;
  if (B == 0 || n == 0) return SLIDE_OK;
  if (B > 65535) return SLIDE_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(n, 128), B);
  launch_k(three_nn_kernel, grid, 128, 0, (cudaStream_t)stream, unknown, known, n, m, dist2, idx);
  return after_launch();
}

static int launch_interp(const float *a, const int *idx, const float *w, int B, int C, int m, int n, float *out,
                         bool grad, cudaStream_t st) {
  if (B == 0 || C == 0 || n == 0) return SLIDE_OK;
  if (B > 65535) return SLIDE_ERR_UNSUPPORTED;
  int cpb = 8, cy = ceil_div(C, cpb);
  if (cy > 65535) {
    cpb = ceil_div(C, 65535);
    cy = ceil_div(C, cpb);
  }
  dim3 grid(ceil_div(n, 256), cy, B);
  if (!grad)
    launch_k(three_interpolate_kernel, grid, 256, 0, st, a, idx, w, C, m, n, cpb, out);
  else
    launch_k(three_interpolate_grad_kernel, grid, 256, 0, st, a, idx, w, C, n, m, cpb, out);
  return after_launch();
}

int slide_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C, int m,
                            int n, float *out, slide_stream_t stream) {
  if (!points || !idx || !weight || !out || B < 0 || C < 0 || m <= 0 || n < 0) return SLIDE_ERR_INVALID;
  return launch_interp(points, idx, weight, B, C, m, n, out, false, (cudaStream_t)stream);
}

int slide_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B, int C,
                                 int n, int m, float *grad_points, slide_stream_t stream) {
  if (!grad_out || !idx || !weight || !grad_points || B < 0 || C < 0 || m <= 0 || n < 0) return SLIDE_ERR_INVALID;
  return launch_interp(grad_out, idx, weight, B, C, m, n, grad_points, true, (cudaStream_t)stream);
}

int slide_knn_points(const float *p1, const float *p2, int B, int P1, int P2, int D, const int64_t *lengths1,
                     const int64_t *lengths2, int K, float *dists, int64_t *idx, slide_stream_t stream) {
  if (!p1 || !p2 || !dists || !idx || B < 0 || P1 < 0 || P2 < 0 || K <= 0) return SLIDE_ERR_INVALID;
  if (D != 3 || K > KNN_MAXK) return SLIDE_ERR_UNSUPPORTED;
  if (B == 0 || P1 == 0) return SLIDE_OK;
  if (B > 65535) return SLIDE_ERR_UNSUPPORTED;
  const int threads = P1 >= 128 ? 128 : ((P1 + 31) / 32) * 32;
  dim3 grid(ceil_div(P1, threads), B);
  cudaStream_t st = (cudaStream_t)stream;
  if (K <= 8)
    launch_k(knn_kernel<8>, grid, threads, 0, st, p1, p2, P1, P2, lengths1, lengths2, K, dists, (long long *)idx);
  else if (K <= 16)
    launch_k(knn_kernel<16>, grid, threads, 0, st, p1, p2, P1, P2, lengths1, lengths2, K, dists, (long long *)idx);
  else if (K <= 32)
    launch_k(knn_kernel<32>, grid, threads, 0, st, p1, p2, P1, P2, lengths1, lengths2, K, dists, (long long *)idx);
  else
    launch_k(knn_kernel<64>, grid, threads, 0, st, p1, p2, P1, P2, lengths1, lengths2, K, dists, (long long *)idx);
  return after_launch();
}

}  // extern "C"
