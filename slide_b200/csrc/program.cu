// Executor of slide programs (include/slide_program.h): one kernel per record, optional CUDA-graph replay.
//
// This file holds the memory-bound "glue" kernels of the fused networks (grouping, soft-max aggregation,
// DDPM update, up-sampling, copies) and the program object; the GEMM kernels live in gemm_simt.cu (fp32 FFMA,
// any shape) and gemm_tc.cu (tcgen05 / TMEM, TF32 operands, fp32 accumulate).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/slide_resident.h"
#include "common.cuh"
#include "program.cuh"

namespace slide {

// =====================================================================================================
// STEP_BEGIN: zero the statistics region, decrement the step counter
// =====================================================================================================
__global__ void step_begin_kernel(uint4 *__restrict__ zero, size_t n16, int *__restrict__ step) {
  pdl_wait();
  pdl_trigger();
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = i0; i < n16; i += stride) zero[i] = make_uint4(0u, 0u, 0u, 0u);
  if (i0 == 0 && step) *step -= 1;
}

// =====================================================================================================
// kNN inside programs: strided point rows, i32 indices, optional squared distances
// =====================================================================================================
// One thread per query; the reference cloud streams through shared memory.  Same insertion rule as
// knn_kernel in index_ops.cu (strict <, ascending index on ties).
constexpr int PK_TILE = 1024;

template <int KCAP>
__global__ void __launch_bounds__(128) knn_prog_kernel(const float *__restrict__ q, int ldq, int P1,
                                                       const float *__restrict__ ref, int ldr, int P2, int K,
                                                       int *__restrict__ idx, float *__restrict__ d2) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[PK_TILE * 3];
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < P1;
  const float *qq = q + ((size_t)b * P1 + (active ? i : 0)) * ldq;
  const float qx = qq[0], qy = qq[1], qz = qq[2];
  const float *r = ref + (size_t)b * P2 * ldr;
  float bd[KCAP];
  int bi[KCAP];
#pragma unroll
  for (int k = 0; k < KCAP; ++k) {
    bd[k] = INFINITY;
    bi[k] = 0;
  }
  for (int base = 0; base < P2; base += PK_TILE) {
    const int tn = min(PK_TILE, P2 - base);
    __syncthreads();
    for (int t = threadIdx.x; t < tn; t += blockDim.x) {
      const float *s = r + (size_t)(base + t) * ldr;
      tile[t * 3 + 0] = s[0];
      tile[t * 3 + 1] = s[1];
      tile[t * 3 + 2] = s[2];
    }
    __syncthreads();
    if (!active) continue;
    for (int k = 0; k < tn; ++k) {
      const float d = sumsq3_p3d(qx - tile[k * 3 + 0], qy - tile[k * 3 + 1], qz - tile[k * 3 + 2]);
      if (d < bd[KCAP - 1]) {
        // branch-free sorted insertion over a fully unrolled register array
        float cd = d;
        int ci = base + k;
#pragma unroll
        for (int s = 0; s < KCAP; ++s) {
          const bool sw = cd < bd[s];
          const float td = bd[s];
          const int ti = bi[s];
          bd[s] = sw ? cd : td;
          bi[s] = sw ? ci : ti;
          cd = sw ? td : cd;
          ci = sw ? ti : ci;
        }
      }
    }
  }
  if (active) {
    int *ii = idx + ((size_t)b * P1 + i) * K;
    float *dd = d2 ? d2 + ((size_t)b * P1 + i) * K : nullptr;
#pragma unroll
    for (int k = 0; k < KCAP; ++k) {
      if (k < K) {
        ii[k] = bi[k];
        if (dd) dd[k] = bd[k];
      }
    }
  }
}

// Large reference clouds (autoencoder / refinement levels: 256..4096 points, K up to 32): ONE WARP PER QUERY.  The sorted
// candidate list lives one entry per lane; 32 reference points are tested per iteration (one per lane) and every point
// that beats the current worst entry is inserted with three shuffles (entries behind it move up one lane).  Points are
// inserted in ascending index order with a strict `<`, i.e. exactly the sequence of the one-thread-per-query kernel
// above -- same neighbours, same order on ties -- at ~1/4 of its instructions and 32x its parallelism (that kernel
// executes its 32-step register insertion whenever ANY lane of the warp inserts: 1.1 ms for 32 x (1024 x 4096), K = 32).
constexpr int KW_WARPS = 8, KW_QPW = 4, KW_TILE = 1024;

__global__ void __launch_bounds__(KW_WARPS * 32) knn_warp_kernel(const float *__restrict__ q, int ldq, int P1,
                                                                   const float *__restrict__ ref, int ldr, int P2, int K,
                                                                   int *__restrict__ idx, float *__restrict__ d2) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tx[KW_TILE], ty[KW_TILE], tz[KW_TILE];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q0 = (blockIdx.x * KW_WARPS + warp) * KW_QPW;
  float qx[KW_QPW], qy[KW_QPW], qz[KW_QPW], bd[KW_QPW];
  int bi[KW_QPW];
#pragma unroll
  for (int j = 0; j < KW_QPW; ++j) {
    const int qi = min(q0 + j, P1 - 1);
    const float *qq = q + ((size_t)b * P1 + qi) * ldq;
    qx[j] = qq[0], qy[j] = qq[1], qz[j] = qq[2];
    bd[j] = INFINITY;
    bi[j] = 0;
  }
  const float *r = ref + (size_t)b * P2 * ldr;
  for (int base = 0; base < P2; base += KW_TILE) {
    const int tn = min(KW_TILE, P2 - base);
    __syncthreads();
    for (int t = threadIdx.x; t < tn; t += blockDim.x) {
      const float *s = r + (size_t)(base + t) * ldr;
      tx[t] = s[0], ty[t] = s[1], tz[t] = s[2];
    }
    __syncthreads();
    for (int k0 = 0; k0 < tn; k0 += 32) {
      const int k = k0 + lane;
      const bool in = k < tn;
      const float x = in ? tx[k] : 0.f, y = in ? ty[k] : 0.f, z = in ? tz[k] : 0.f;
#pragma unroll
      for (int j = 0; j < KW_QPW; ++j) {
        const float d = in ? sumsq3_p3d(qx[j] - x, qy[j] - y, qz[j] - z) : INFINITY;
        const float worst = __shfl_sync(0xffffffffu, bd[j], K - 1);
        unsigned m = __ballot_sync(0xffffffffu, d < worst);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float cd = __shfl_sync(0xffffffffu, d, src);
          const int ci = base + k0 + src;
          const bool flag = cd < bd[j];
          const float ud = __shfl_up_sync(0xffffffffu, bd[j], 1);
          const int ui = __shfl_up_sync(0xffffffffu, bi[j], 1);
          const bool uflag = __shfl_up_sync(0xffffffffu, (int)flag, 1) && lane > 0;
          if (flag) {
            bd[j] = uflag ? ud : cd;
            bi[j] = uflag ? ui : ci;
          }
        }
      }
    }
  }
  if (lane < K) {
#pragma unroll
    for (int j = 0; j < KW_QPW; ++j) {
      const int qi = q0 + j;
      if (qi < P1) {
        idx[((size_t)b * P1 + qi) * K + lane] = bi[j];
        if (d2) d2[((size_t)b * P1 + qi) * K + lane] = bd[j];
      }
    }
  }
}

// =====================================================================================================
// GROUP (wide rows): one warp per grouped row; channels-last makes the feature part a contiguous row copy
// =====================================================================================================
__global__ void __launch_bounds__(256) group_rows_kernel(int mode, const float *__restrict__ F, int ldf, int C,
                                                    const float *__restrict__ xyz, int ldx, int N,
                                                    const float *__restrict__ ctr, int ldc, int np,
                                                    const int *__restrict__ idx, int K, const float *__restrict__ d2,
                                                    float *__restrict__ out, int ldo, int inc_abs, int inc_ctr,
                                                    long long rows) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp0; row < rows; row += nwarps) {
    const long long pi = row / K;  // (sample, point)
    const int s = (int)(pi / np);
    const int j = idx[row];
    float *o = out + row * ldo;
    if (C > 0) {
      const float *f = F + ((size_t)s * N + j) * ldf;
      for (int c = lane; c < C; c += 32) o[c] = __ldg(f + c);
    }
    const float *xj = xyz + ((size_t)s * N + j) * ldx;
    const float *ci = ctr + (size_t)pi * ldc;
    if (mode == 0) {
      // [x_j - c_i | x_j | c_i]
      if (lane < 3) {
        const float a = __ldg(xj + lane), c = __ldg(ci + lane);
        int pos = C;
        o[pos + lane] = __fsub_rn(a, c);
        pos += 3;
        if (inc_abs) {
          o[pos + lane] = a;
          pos += 3;
        }
        if (inc_ctr) o[pos + lane] = c;
      }
    } else {
      // [d2 | w | x_j | x_j - c_i | c_i],  w = (1/(d2+1e-8)) / sum_k (1/(d2+1e-8))  (sum in ascending k like torch.sum)
      if (lane < 3) {
        const float a = __ldg(xj + lane), c = __ldg(ci + lane);
        o[C + 2 + lane] = a;
        o[C + 5 + lane] = __fsub_rn(a, c);
        o[C + 8 + lane] = c;
      } else if (lane == 3) {
        const float *dr = d2 + pi * K;
        float sum = 0.f;
        for (int k = 0; k < K; ++k) sum = __fadd_rn(sum, __fdiv_rn(1.0f, __fadd_rn(dr[k], 1e-8f)));
        const float dk = d2[row];
        o[C] = dk;
        o[C + 1] = __fdiv_rn(__fdiv_rn(1.0f, __fadd_rn(dk, 1e-8f)), sum);
      }
    }
  }
}

// =====================================================================================================
// GROUP (narrow rows, W <= 16: the xyz-only inputs): one thread per output element.  Every element is an independent
// index -> gather -> store chain, so the whole tensor is in flight at once (measured on B200, 65536 rows x 12:
// 24.6 -> 12.3 us; for wide rows the warp-per-row kernel above wins: 28.7 vs 65.5 us at 278 columns).
// =====================================================================================================
__global__ void __launch_bounds__(256) group_elem_kernel(int mode, const float *__restrict__ F, int ldf, int C,
                                                    const float *__restrict__ xyz, int ldx, int N,
                                                    const float *__restrict__ ctr, int ldc, int np,
                                                    const int *__restrict__ idx, int K, const float *__restrict__ d2,
                                                    float *__restrict__ out, int ldo, int inc_abs, int inc_ctr,
                                                    long long rows, int W) {
  pdl_wait();
  pdl_trigger();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * W) return;
  const long long row = e / W;
  const int c = (int)(e - row * W);
  const long long pi = row / K;  // (sample, point)
  const int s = (int)(pi / np);
  const int j = __ldg(idx + row);
  float v;
  if (c < C) {
    v = __ldg(F + ((size_t)s * N + j) * ldf + c);
  } else {
    const int t = c - C;
    const float *xj = xyz + ((size_t)s * N + j) * ldx;
    const float *ci = ctr + (size_t)pi * ldc;
    if (mode == 0) {
      // [x_j - c_i | x_j | c_i]  (the abs / centre blocks are optional)
      const int blk = t / 3, d = t - 3 * blk;
      const int what = blk == 0 ? 0 : (blk == 1 && inc_abs ? 1 : 2);
      const float a = __ldg(xj + d), cc = __ldg(ci + d);
      v = what == 0 ? __fsub_rn(a, cc) : (what == 1 ? a : cc);
    } else {
      // [d2 | w | x_j | x_j - c_i | c_i],  w = (1/(d2+1e-8)) / sum_k (1/(d2+1e-8))  (sum in ascending k like torch.sum)
      if (t == 0) {
        v = __ldg(d2 + row);
      } else if (t == 1) {
        const float *dr = d2 + pi * K;
        float sum = 0.f;
        for (int k = 0; k < K; ++k) sum = __fadd_rn(sum, __fdiv_rn(1.0f, __fadd_rn(__ldg(dr + k), 1e-8f)));
        v = __fdiv_rn(__fdiv_rn(1.0f, __fadd_rn(__ldg(d2 + row), 1e-8f)), sum);
      } else {
        const int blk = (t - 2) / 3, d = (t - 2) - 3 * blk;
        const float a = __ldg(xj + d), cc = __ldg(ci + d);
        v = blk == 0 ? a : (blk == 1 ? __fsub_rn(a, cc) : cc);
      }
    }
  }
  out[row * ldo + c] = v;
}

// =====================================================================================================
// SOFTMAX_WSUM: out[i,c] = sum_k xf(V)[i,k,c] * softmax_k(S[i,k,c])   (one thread per (i,c), coalesced in c)
// =====================================================================================================
__global__ void __launch_bounds__(256) softmax_wsum_kernel(const float *__restrict__ S, int lds,
                                                           const float *__restrict__ Vt, int ldv, XFd xf,
                                                           const int *__restrict__ step_ptr, float *__restrict__ out,
                                                           int ldo, long long rows, int K, int C) {
  pdl_wait();
  pdl_trigger();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * C) return;
  const long long i = e / C;
  const int c = (int)(e - i * C);
  const long long r0 = i * K;
  const int step = step_ptr ? *step_ptr : 0;
  // all K rows of one point belong to one sample
  const int s = xf.R > 0 ? (int)(r0 / xf.R) : 0;
  float mean = 0.f, rstd = 1.f, gam = 1.f, bet = 0.f;
  bool norm = false;
  if (xf.stats) {
    const int ch = xf.choff + c;
    if (ch < xf.nnorm) {
      norm = true;
      const int G = xf.nnorm / xf.cg;
      const double *st = xf.stats + ((size_t)s * G + ch / xf.cg) * 2;
      const double m = st[0] * (double)xf.inv_count;
      double var = st[1] * (double)xf.inv_count - m * m;
      var = var < 0.0 ? 0.0 : var;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)SLIDE_GN_EPS));
      gam = __ldg(xf.gamma + ch);
      bet = __ldg(xf.beta + ch);
    }
  }
  float add = 0.f;
  if (xf.addvec) {
    const long long arow = xf.addmode == 0 ? s : (xf.addmode == 1 ? step : 0);
    add = xf.addvec[arow * xf.addld + c];
  }
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, S[(r0 + k) * lds + c]);
  float den = 0.f, acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float w = expf(S[(r0 + k) * lds + c] - mx);
    float v = Vt[(r0 + k) * ldv + c];
    if (norm) v = (v - mean) * rstd * gam + bet;
    if (xf.relu) v = fmaxf(v, 0.f);
    v += add;
    den += w;
    acc = fmaf(v, w, acc);
  }
  out[i * ldo + c] = acc / den;
}

// =====================================================================================================
// element-wise helpers
// =====================================================================================================
__global__ void copy_cols_kernel(const float *__restrict__ src, int lds, float *__restrict__ dst, int ldd,
                                 long long rows, int n) {
  pdl_wait();
  pdl_trigger();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * n) return;
  const long long r = e / n;
  const int c = (int)(e - r * n);
  dst[r * ldd + c] = src[r * lds + c];
}

// DDPM ancestral update; every multiply / add is a separate IEEE operation in the reference's order (torch
// evaluates the expression op by op), so given the same eps the update is bit-exact.
__global__ void ddpm_update_kernel(int mode, float *__restrict__ x, int ldx, const float *__restrict__ eps, int lde,
                                   const float *__restrict__ noise, long long rows, int ncols, int col0,
                                   const float *__restrict__ table, const int *__restrict__ step_ptr, float clamp,
                                   const float *__restrict__ x0c, int ldx0c, const float *__restrict__ mask) {
  pdl_wait();
  pdl_trigger();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * ncols) return;
  const long long r = e / ncols;
  const int c = (int)(e - r * ncols);
  if (c < col0) return;
  const int t = *step_ptr;
  const float *tab = table + (size_t)t * 8;
  const float xv = x[r * ldx + c], ev = eps[r * lde + c];
  const float nz = noise[((size_t)t * rows + r) * ncols + c];
  float res;
  if (mode == 0) {
    // x = (x - k1*eps) / sqrt_alpha ; if t > 0: x += sigma * z          (pointnet2/util.py:247-253)
    res = __fdiv_rn(__fsub_rn(xv, __fmul_rn(tab[0], ev)), tab[1]);
    if (t > 0) res = __fadd_rn(res, __fmul_rn(tab[2], nz));
  } else if (mode == 2) {
    // x *= a ; x += c*eps + sigma*z                                       (util_fastdpmv2.py:440-443)
    res = __fadd_rn(__fmul_rn(xv, tab[0]), __fadd_rn(__fmul_rn(tab[1], ev), __fmul_rn(tab[2], nz)));
  } else {
    // x0 = c1*x - c2*eps ; clamp ; mean = pm1*x0 + pm2*x ; x = mean + (t != 0) * sig * z   (diffusion.py:71-92)
    float x0 = __fsub_rn(__fmul_rn(tab[0], xv), __fmul_rn(tab[1], ev));
    if (clamp > 0.f) x0 = fminf(fmaxf(x0, -clamp), clamp);
    if (x0c) {
      // local resampling: pred_xstart * keypoint_mask + complete_x0 * (1 - keypoint_mask)   (diffusion.py:76-79)
      const float m = mask[r];
      x0 = __fadd_rn(__fmul_rn(x0, m), __fmul_rn(x0c[r * ldx0c + c], __fsub_rn(1.0f, m)));
    }
    const float mean = __fadd_rn(__fmul_rn(tab[2], x0), __fmul_rn(tab[3], xv));
    const float m = t == 0 ? 0.f : 1.f;
    res = __fadd_rn(mean, __fmul_rn(__fmul_rn(m, tab[4]), nz));
  }
  x[r * ldx + c] = res;
}

__global__ void gather_rows_kernel(const float *__restrict__ src, int lds, int N, const int *__restrict__ idx, int m,
                                   float *__restrict__ dst, int ldd, int ncols, long long total_rows) {
  pdl_wait();
  pdl_trigger();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total_rows * ncols) return;
  const long long r = e / ncols;  // (sample, j)
  const int c = (int)(e - r * ncols);
  const long long s = r / m;
  const int j = idx[r];
  dst[r * ldd + c] = src[((size_t)s * N + j) * lds + c];
}

__global__ void upsample_kernel(const float *__restrict__ coarse, int ldc, int coarse_c,
                                const float *__restrict__ disp, int ldd, float *__restrict__ out, int ldo,
                                long long rows, int factor, int Fd, float inv_sqrt_f, float scale) {
  pdl_wait();
  pdl_trigger();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * factor * Fd) return;
  const long long orow = e / Fd;
  const int c = (int)(e - orow * Fd);
  const long long n = orow / factor;
  const int q = (int)(orow - n * factor);
  const float base = c < coarse_c ? coarse[n * ldc + c] : 0.f;
  const float d = disp[n * ldd + q * Fd + c];
  out[orow * ldo + c] = __fadd_rn(base, __fmul_rn(__fmul_rn(d, inv_sqrt_f), scale));
}

__global__ void temb_kernel(const float *__restrict__ ts, const float *__restrict__ freq, int half,
                            float *__restrict__ out, int ldo, int rows) {
  pdl_wait();
  pdl_trigger();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * half) return;
  const int r = e / half, j = e - r * half;
  const float arg = __fmul_rn(ts[r], freq[j]);
  out[(size_t)r * ldo + j] = sinf(arg);
  out[(size_t)r * ldo + half + j] = cosf(arg);
}

// =====================================================================================================
// PAIR: a 1x1 conv over grouped rows, factored through the gather (see slide_program.h)
// =====================================================================================================
// CTA = (block of PB points of one sample) x all N columns; thread = column (coalesced U gathers / stores / residual
// loads), loop over the block's points and their K neighbours.  Index, coordinate and distance loads are warp-uniform
// broadcasts.  Per-column statistics stay in registers and leave the CTA as one fp64 atomic per (group, moment).
struct PairArgs {
  const float *U;
  int ldu, nsrc;
  const float *xyz;
  int ldx;
  const float *ctr;
  int ldctr, np;
  const int *idx;
  int K;
  const float *d2;
  const float *wx, *wc, *wd, *ww, *bias;
  int N;
  float *out;
  int ldo, act;
  const float *res;
  int ldr;
  double *st_stats;
  int st_cg, st_nnorm, st_choff;
  float st_weight;
  XFd xfr;
  const int *step;
  int pb;  // points per CTA
  int pf;  // pair_smem_kernel: L2 prefetch distance of the residual rows, in loop iterations (0 = none)
};

// RB = rows in flight per thread.  The neighbour indices (and squared distances) of a point are fetched ONCE as a
// lane-parallel row and broadcast with shuffles, so a batch of RB rows costs one memory latency (the U gathers,
// coordinate and residual loads of the whole batch are independent) instead of index -> gather chains per row.
#ifndef PAIR_MIN_BLOCKS
#define PAIR_MIN_BLOCKS 1  // A/B knob (register cap of the gather kernel: 1 -> 116-127 registers, 3 -> 80, 4 -> 64 + spills)
#endif
// LITE: the caller folded the neighbour-coordinate term into U (U'[j] = f_j W_f^T + x_j wx^T, one K = 3 GEMM over the source
// points; a.wx == nullptr), so a row is U'[idx] + (per-point centre term): no coordinate loads, 12 fewer FMAs and 36 fewer
// live registers per thread -- the large-source PAIR records of the autoencoder / refinement levels are bound by exactly that
// (127 registers -> 15 warps per SM, ncu: 48 % SM busy at 0.3-2.5 TB/s).
template <int RB, bool HAS_RES, bool LITE>
__global__ void __launch_bounds__(256, PAIR_MIN_BLOCKS) pair_kernel(PairArgs a) {
  pdl_wait();
  pdl_trigger();
  // thread = 4 consecutive columns (128-bit U gathers / residual loads / stores); the per-row index, coordinate and
  // distance loads are amortised over the 4 outputs
  const int s = blockIdx.y;
  const int p0 = blockIdx.x * a.pb;
  const int p1 = min(p0 + a.pb, a.np);
  const int lane = threadIdx.x & 31;
  const int step = a.step ? *a.step : 0;
  for (int n0 = 0; n0 < a.N; n0 += 4 * blockDim.x) {
    const int n = n0 + 4 * threadIdx.x;
    bool on[4];
    float wx[4][3], wc[4][3], bias[4], wd[4], ww[4], rsc[4], rsh[4], radd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      on[u] = n + u < a.N;
      const int nn = on[u] ? n + u : 0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        wx[u][d] = LITE ? 0.f : __ldg(a.wx + nn * 3 + d);
        wc[u][d] = __ldg(a.wc + nn * 3 + d);
      }
      bias[u] = a.bias ? __ldg(a.bias + nn) : 0.f;
      wd[u] = a.d2 ? __ldg(a.wd + nn) : 0.f;
      ww[u] = a.d2 ? __ldg(a.ww + nn) : 0.f;
      rsc[u] = 1.f;
      rsh[u] = 0.f;
      radd[u] = 0.f;
      if (HAS_RES) {
        if (a.xfr.stats) {
          const int ch = a.xfr.choff + nn;
          if (ch < a.xfr.nnorm) {
            const int G = a.xfr.nnorm / a.xfr.cg;
            const double *st = a.xfr.stats + ((size_t)s * G + ch / a.xfr.cg) * 2;
            const double m = st[0] * (double)a.xfr.inv_count;
            double var = st[1] * (double)a.xfr.inv_count - m * m;
            var = var < 0.0 ? 0.0 : var;
            rsc[u] = (float)(1.0 / sqrt(var + (double)SLIDE_GN_EPS)) * __ldg(a.xfr.gamma + ch);
            rsh[u] = __ldg(a.xfr.beta + ch) - (float)m * rsc[u];
          }
        }
        if (a.xfr.addvec) {
          const long long arow = a.xfr.addmode == 0 ? s : (a.xfr.addmode == 1 ? step : 0);
          radd[u] = a.xfr.addvec[arow * a.xfr.addld + nn];
        }
      }
    }
    const bool any = on[0];
    const bool full = on[3];
    float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = p0; i < p1; ++i) {
      const float *c = a.ctr + ((size_t)s * a.np + i) * a.ldctr;
      const float c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2);
      float vterm[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) vterm[u] = fmaf(c2, wc[u][2], fmaf(c1, wc[u][1], fmaf(c0, wc[u][0], bias[u])));
      const size_t prow = ((size_t)s * a.np + i) * a.K;
      float inv_sum = 0.f;
      if (a.d2) {
        // the reference's left-to-right fp32 sum over the K neighbours (terms computed lane-parallel)
        for (int kc = 0; kc < a.K; kc += 32) {
          const int kn = min(32, a.K - kc);
          const float term = lane < kn ? __fdiv_rn(1.0f, __fadd_rn(__ldg(a.d2 + prow + kc + lane), 1e-8f)) : 0.f;
          for (int k = 0; k < kn; ++k) inv_sum = __fadd_rn(inv_sum, __shfl_sync(0xffffffffu, term, k));
        }
      }
      for (int kc = 0; kc < a.K; kc += 32) {
        const int kn = min(32, a.K - kc);  // warp-uniform
        const int jrow = lane < kn ? __ldg(a.idx + prow + kc + lane) : 0;
        const float drow = (a.d2 && lane < kn) ? __ldg(a.d2 + prow + kc + lane) : 0.f;
        for (int k0 = 0; k0 < kn; k0 += RB) {
          float4 uu[RB], r4[RB];
          float xs[RB][3], dk[RB];
          bool kon[RB];
#pragma unroll
          for (int b = 0; b < RB; ++b) {
            kon[b] = k0 + b < kn;
            const int j = __shfl_sync(0xffffffffu, jrow, (k0 + b) & 31);
            dk[b] = __shfl_sync(0xffffffffu, drow, (k0 + b) & 31);
            const size_t row = prow + kc + k0 + b;
            if (!LITE) {
              const float *x = a.xyz + ((size_t)s * a.nsrc + j) * a.ldx;
              xs[b][0] = __ldg(x);
              xs[b][1] = __ldg(x + 1);
              xs[b][2] = __ldg(x + 2);
            } else {
              xs[b][0] = xs[b][1] = xs[b][2] = 0.f;
            }
            uu[b] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (any) uu[b] = __ldg(reinterpret_cast<const float4 *>(a.U + ((size_t)s * a.nsrc + j) * a.ldu + n));
            r4[b] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (HAS_RES && any && kon[b]) {
              if (full) {
                r4[b] = __ldcs(reinterpret_cast<const float4 *>(a.res + row * a.ldr + n));
              } else {
                float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (on[u]) t[u] = a.res[row * a.ldr + n + u];
                r4[b] = make_float4(t[0], t[1], t[2], t[3]);
              }
            }
          }
#pragma unroll
          for (int b = 0; b < RB; ++b) {
            if (!kon[b]) continue;
            const size_t row = prow + kc + k0 + b;
            float v[4] = {uu[b].x, uu[b].y, uu[b].z, uu[b].w};
            const float rr[4] = {r4[b].x, r4[b].y, r4[b].z, r4[b].w};
            float w = 0.f;
            if (a.d2) w = __fdiv_rn(__fdiv_rn(1.0f, __fadd_rn(dk[b], 1e-8f)), inv_sum);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float t = LITE ? v[u] + vterm[u]
                             : fmaf(xs[b][2], wx[u][2], fmaf(xs[b][1], wx[u][1], fmaf(xs[b][0], wx[u][0], v[u]))) + vterm[u];
              if (a.d2) t = fmaf(w, ww[u], fmaf(dk[b], wd[u], t));
              if (HAS_RES) {
                float r = fmaf(rr[u], rsc[u], rsh[u]);
                if (a.xfr.relu) r = fmaxf(r, 0.f);
                t += r + radd[u];
              }
              if (a.act == 1) t = fmaxf(t, 0.f);
              v[u] = on[u] ? t : 0.f;
              ssum[u] += v[u];
              ssq[u] = fmaf(v[u], v[u], ssq[u]);
            }
            if (full) {
              *reinterpret_cast<float4 *>(a.out + row * a.ldo + n) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (on[u]) a.out[row * a.ldo + n + u] = v[u];
            }
          }
        }
      }
    }
    if (a.st_stats) {
      const int G = a.st_nnorm / a.st_cg;
      const int ch0 = a.st_choff + n;
      const bool pow2 = (a.st_cg & (a.st_cg - 1)) == 0 && (a.st_choff % a.st_cg) == 0 && (n0 % a.st_cg) == 0;
      if (pow2 && a.st_cg >= 4) {
        // the thread's 4 columns lie in one GroupNorm group; the group's threads are st_cg/4 consecutive, aligned lanes
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (on[u] && ch0 + u < a.st_nnorm) {
            ts += ssum[u];
            tq += ssq[u];
          }
        }
        const int seg = min(a.st_cg / 4, 32);
        for (int d = 1; d < seg; d <<= 1) {
          ts += __shfl_xor_sync(0xffffffffu, ts, d);
          tq += __shfl_xor_sync(0xffffffffu, tq, d);
        }
        if ((lane & (seg - 1)) == 0 && any && ch0 < a.st_nnorm) {
          double *slot = a.st_stats + ((size_t)s * G + ch0 / a.st_cg) * 2;
          atomicAdd(slot, (double)ts * (double)a.st_weight);
          atomicAdd(slot + 1, (double)tq * (double)a.st_weight);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (on[u] && ch0 + u < a.st_nnorm) {
            double *slot = a.st_stats + ((size_t)s * G + (ch0 + u) / a.st_cg) * 2;
            atomicAdd(slot, (double)ssum[u] * (double)a.st_weight);
            atomicAdd(slot + 1, (double)ssq[u] * (double)a.st_weight);
          }
        }
      }
    }
  }
}

// ---- PAIR with the source rows staged in shared memory ---------------------------------------------------------------
// For the denoisers the neighbours of a point come from the sample's own 16 points, so the whole gather source fits on
// chip: the CTA first builds  Us[j][n] = U[j][n] + x_j . wx[n]  for every source point j of its sample and
// Vs[i][n] = c_i . wc[n] + bias[n]  for its points (the same fused-multiply-add chains, in the same order, as the
// gather kernel above: results are bit-identical), then every output row is  act(Us[idx] + Vs[i] (+ the two distance
// terms) (+ the transformed residual))  -- one 16-byte shared-memory read, 4 adds and one 16-byte store per 4 outputs, no
// dependent global gather and no per-row address arithmetic (the gather kernel spends ~47 instructions per output).
// Thread = (row group rg, 4 columns ct); rows of a group are processed 4 at a time so that the residual loads and the
// stores of 4 rows are in flight together.
constexpr int PS_THREADS = 256;
constexpr int PS_MAX_SRC = 64;

template <bool HAS_RES, bool HAS_D2>
__global__ void __launch_bounds__(PS_THREADS) pair_smem_kernel(PairArgs a, int CT, int RG, int ldn) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float ps_smem[];
  const int s = blockIdx.y;
  const int p0 = blockIdx.x * a.pb;
  const int pbl = min(a.pb, a.np - p0);
  const int rows = pbl * a.K;
  float *Us = ps_smem;                                   // [nsrc][ldn]
  float *Vs = Us + (size_t)a.nsrc * ldn;                 // [pb][ldn]
  float *part = Vs + (size_t)a.pb * ldn;                 // [RG][ldn][2] column sums per row group
  int *idx_s = reinterpret_cast<int *>(part + (size_t)RG * ldn * 2);  // [pb * K]
  float *dk_s = reinterpret_cast<float *>(idx_s + a.pb * a.K);        // [pb * K] squared distances
  float *w_s = dk_s + a.pb * a.K;                                       // [pb * K] interpolation weights
  const int tid = threadIdx.x;
  const int rg = tid / CT, ct = tid - rg * CT;
  const bool live = rg < RG;
  const int n = 4 * ct;
  const int step = a.step ? *a.step : 0;
  const size_t prow0 = ((size_t)s * a.np + p0) * a.K;

  // ---- per-row scalars (stored together with the staged rows: one barrier for both)
  for (int r = tid; r < rows; r += PS_THREADS) {
    idx_s[r] = __ldg(a.idx + prow0 + r);
    if (HAS_D2) dk_s[r] = __ldg(a.d2 + prow0 + r);
  }
  if (HAS_RES && a.pf > 0 && live && (ct & 7) == 0 && n + 3 < a.N) {
    // the first iterations' residual rows: requested now, they arrive while the source rows are staged below
    for (int r = rg; r < rows && r < rg + 4 * a.pf * RG; r += RG)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + (prow0 + r) * a.ldr + n));
  }
  // ---- per-column constants of this thread's 4 columns
  bool on[4];
  float wx[4][3], wc[4][3], bias[4], wd[4], ww[4], rsc[4], rsh[4], radd[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    on[u] = live && n + u < a.N;
    const int nn = on[u] ? n + u : 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      wx[u][d] = __ldg(a.wx + nn * 3 + d);
      wc[u][d] = __ldg(a.wc + nn * 3 + d);
    }
    bias[u] = a.bias ? __ldg(a.bias + nn) : 0.f;
    wd[u] = HAS_D2 ? __ldg(a.wd + nn) : 0.f;
    ww[u] = HAS_D2 ? __ldg(a.ww + nn) : 0.f;
    rsc[u] = 1.f;
    rsh[u] = 0.f;
    radd[u] = 0.f;
    if (HAS_RES) {
      if (a.xfr.stats) {
        const int ch = a.xfr.choff + nn;
        if (ch < a.xfr.nnorm) {
          const int G = a.xfr.nnorm / a.xfr.cg;
          const double *st = a.xfr.stats + ((size_t)s * G + ch / a.xfr.cg) * 2;
          const double m = st[0] * (double)a.xfr.inv_count;
          double var = st[1] * (double)a.xfr.inv_count - m * m;
          var = var < 0.0 ? 0.0 : var;
          rsc[u] = (float)(1.0 / sqrt(var + (double)SLIDE_GN_EPS)) * __ldg(a.xfr.gamma + ch);
          rsh[u] = __ldg(a.xfr.beta + ch) - (float)m * rsc[u];
        }
      }
      if (a.xfr.addvec) {
        const long long arow = a.xfr.addmode == 0 ? s : (a.xfr.addmode == 1 ? step : 0);
        radd[u] = a.xfr.addvec[arow * a.xfr.addld + nn];
      }
    }
  }
  const bool any = on[0], full = on[3];
  // ---- stage Us (all source points of the sample) and Vs (this CTA's points)
  if (any) {
    for (int j = rg; j < a.nsrc; j += RG) {
      const float *x = a.xyz + ((size_t)s * a.nsrc + j) * a.ldx;
      const float x0 = __ldg(x), x1 = __ldg(x + 1), x2 = __ldg(x + 2);
      const float4 u4 = __ldg(reinterpret_cast<const float4 *>(a.U + ((size_t)s * a.nsrc + j) * a.ldu + n));
      const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
      float t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = fmaf(x2, wx[u][2], fmaf(x1, wx[u][1], fmaf(x0, wx[u][0], uu[u])));
      *reinterpret_cast<float4 *>(Us + (size_t)j * ldn + n) = make_float4(t[0], t[1], t[2], t[3]);
    }
    for (int il = rg; il < pbl; il += RG) {
      const float *c = a.ctr + ((size_t)s * a.np + p0 + il) * a.ldctr;
      const float c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2);
      float t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = fmaf(c2, wc[u][2], fmaf(c1, wc[u][1], fmaf(c0, wc[u][0], bias[u])));
      *reinterpret_cast<float4 *>(Vs + (size_t)il * ldn + n) = make_float4(t[0], t[1], t[2], t[3]);
    }
  }
  __syncthreads();
  if (HAS_D2) {
    if (tid < pbl) {
      // the reference's left-to-right fp32 sum over the K neighbours, then w_k = (1 / (d_k + 1e-8)) / sum
      float inv_sum = 0.f;
      for (int k = 0; k < a.K; ++k) inv_sum = __fadd_rn(inv_sum, __fdiv_rn(1.0f, __fadd_rn(dk_s[tid * a.K + k], 1e-8f)));
      for (int k = 0; k < a.K; ++k)
        w_s[tid * a.K + k] = __fdiv_rn(__fdiv_rn(1.0f, __fadd_rn(dk_s[tid * a.K + k], 1e-8f)), inv_sum);
    }
    __syncthreads();
  }

  // ---- output rows
  // The residual stream is the kernel's HBM read (256 KB per CTA at N = 512).  A thread keeps 4 x 16 bytes of it in flight
  // (registers); at 3 CTAs per SM that is 48 KB per SM, short of what 6.5 TB/s at loaded-DRAM latency needs (ncu: 2.8
  // TB/s, 30 % occupancy, long-scoreboard stalls).  One lane per 128-byte line therefore asks L2 for the rows of the
  // next a.pf iterations -- prefetches hold no registers -- so the loads below hit L2.
  const bool pf_lane = HAS_RES && a.pf > 0 && full && (ct & 7) == 0;
  float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
  if (any) {
    for (int r0 = rg; r0 < rows; r0 += 4 * RG) {
      if (pf_lane) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int r = r0 + (4 * a.pf + b) * RG;
          if (r < rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + (prow0 + r) * a.ldr + n));
        }
      }
      float4 r4[4];
      bool ron[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int r = r0 + b * RG;
        ron[b] = r < rows;
        r4[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (HAS_RES && ron[b]) {
          const float *rp = a.res + (prow0 + r) * a.ldr + n;
          if (full) {
            r4[b] = __ldcs(reinterpret_cast<const float4 *>(rp));
          } else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (on[u]) t[u] = rp[u];
            r4[b] = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (!ron[b]) continue;
        const int r = r0 + b * RG;
        const int il = r / a.K;
        const float4 u4 = *reinterpret_cast<const float4 *>(Us + (size_t)idx_s[r] * ldn + n);
        const float4 v4 = *reinterpret_cast<const float4 *>(Vs + (size_t)il * ldn + n);
        float v[4] = {u4.x + v4.x, u4.y + v4.y, u4.z + v4.z, u4.w + v4.w};
        const float rr[4] = {r4[b].x, r4[b].y, r4[b].z, r4[b].w};
        float dk = 0.f, w = 0.f;
        if (HAS_D2) {
          dk = dk_s[r];
          w = w_s[r];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float t = v[u];
          if (HAS_D2) t = fmaf(w, ww[u], fmaf(dk, wd[u], t));
          if (HAS_RES) {
            float q = fmaf(rr[u], rsc[u], rsh[u]);
            if (a.xfr.relu) q = fmaxf(q, 0.f);
            t += q + radd[u];
          }
          if (a.act == 1) t = fmaxf(t, 0.f);
          v[u] = on[u] ? t : 0.f;
          ssum[u] += v[u];
          ssq[u] = fmaf(v[u], v[u], ssq[u]);
        }
        float *op = a.out + (prow0 + r) * a.ldo + n;
        if (full) {
          *reinterpret_cast<float4 *>(op) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (on[u]) op[u] = v[u];
        }
      }
    }
  }
  // ---- statistics: column sums per row group -> per GroupNorm group -> one fp64 atomic pair per group
  if (a.st_stats) {
    if (any) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        *reinterpret_cast<float2 *>(part + ((size_t)rg * ldn + n + u) * 2) = make_float2(ssum[u], ssq[u]);
    }
    __syncthreads();
    const int G = a.st_nnorm / a.st_cg;
    const int lim = min(a.N, a.st_nnorm - a.st_choff);  // local columns [0, lim) are normalised channels
    if (lim > 0) {
      const int g_first = a.st_choff / a.st_cg, g_last = (a.st_choff + lim - 1) / a.st_cg;
      for (int gi = g_first + tid; gi <= g_last; gi += PS_THREADS) {
        const int c0 = max(0, gi * a.st_cg - a.st_choff), c1 = min(lim, (gi + 1) * a.st_cg - a.st_choff);
        float ts = 0.f, tq = 0.f;
        for (int g = 0; g < RG; ++g)
          for (int c = c0; c < c1; ++c) {
            const float2 v = *reinterpret_cast<const float2 *>(part + ((size_t)g * ldn + c) * 2);
            ts += v.x;
            tq += v.y;
          }
        double *slot = a.st_stats + ((size_t)s * G + gi) * 2;
        atomicAdd(slot, (double)ts * (double)a.st_weight);
        atomicAdd(slot + 1, (double)tq * (double)a.st_weight);
      }
    }
  }
}

template <bool HAS_RES, bool HAS_D2>
static int launch_pair_smem(const PairArgs &a, int B, int CT, int RG, int ldn, size_t smem, cudaStream_t st) {
  static bool configured[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(pair_smem_kernel<HAS_RES, HAS_D2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         160 * 1024);
    if (e != cudaSuccess) return cuda_rc(e);
    configured[dev] = true;
  }
  dim3 grid(ceil_div(a.np, a.pb), B);
  launch_k(pair_smem_kernel<HAS_RES, HAS_D2>, grid, PS_THREADS, smem, st, a, CT, RG, ldn);
  return after_launch();
}

// out[s,c] = max_r xf(X)[s*R + r, c]: thread per (sample, column), coalesced in c (Pnet2Stage's global max-pool)
__global__ void colmax_kernel(const float *__restrict__ X, int ldx, int R, int C, XFd xf,
                              const int *__restrict__ step_ptr, float *__restrict__ out, int ldo, int B) {
  pdl_wait();
  pdl_trigger();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)B * C) return;
  const int s = (int)(e / C), c = (int)(e - (long long)s * C);
  const int step = step_ptr ? *step_ptr : 0;
  float mean = 0.f, rstd = 1.f, gam = 1.f, bet = 0.f;
  bool norm = false;
  if (xf.stats) {
    const int ch = xf.choff + c;
    if (ch < xf.nnorm) {
      norm = true;
      const int G = xf.nnorm / xf.cg;
      const double *st = xf.stats + ((size_t)s * G + ch / xf.cg) * 2;  // xf.R == R: one statistics row per sample
      const double m = st[0] * (double)xf.inv_count;
      double var = st[1] * (double)xf.inv_count - m * m;
      var = var < 0.0 ? 0.0 : var;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)SLIDE_GN_EPS));
      gam = __ldg(xf.gamma + ch);
      bet = __ldg(xf.beta + ch);
    }
  }
  float add = 0.f;
  if (xf.addvec) {
    const long long arow = xf.addmode == 0 ? s : (xf.addmode == 1 ? step : 0);
    add = xf.addvec[arow * xf.addld + c];
  }
  float mx = -INFINITY;
  for (int r = 0; r < R; ++r) {
    float v = X[((size_t)s * R + r) * ldx + c];
    if (norm) v = (v - mean) * rstd * gam + bet;
    if (xf.relu) v = fmaxf(v, 0.f);
    mx = fmaxf(mx, v + add);
  }
  out[(size_t)s * ldo + c] = mx;
}

// DiagonalGaussianDistribution: mode, or mean + exp(0.5 * clamp(logvar, -30, 20)) * noise (op by op like torch)
__global__ void kl_kernel(const float *__restrict__ P, int ldp, int C, const float *__restrict__ noise, int ldn,
                          float *__restrict__ out, int ldo, long long rows) {
  pdl_wait();
  pdl_trigger();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * C) return;
  const long long r = e / C;
  const int c = (int)(e - r * C);
  float v = P[r * ldp + c];
  if (noise) {
    const float lv = fminf(fmaxf(P[r * ldp + C + c], -30.0f), 20.0f);
    v = __fadd_rn(v, __fmul_rn(expf(__fmul_rn(0.5f, lv)), noise[r * ldn + c]));
  }
  out[r * ldo + c] = v;
}

}  // namespace slide

using namespace slide;

// =====================================================================================================
// program object
// =====================================================================================================
struct slide_program {
  std::vector<slide_op> ops;
  char *arena = nullptr;
  char *weights = nullptr;
  size_t arena_bytes = 0, weights_bytes = 0;
  cudaGraphExec_t graphs[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int graph_launches[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int gemm_backend = 0;  // 0 = auto (tcgen05 where eligible), 1 = fp32 FFMA everywhere
  // two-branch regions (SLIDE_OPF_SIDE): a second stream plus fork / join events
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int use_side = 1;  // SLIDE_SIDE_BRANCH=0 runs everything on one stream
  // scratch of the GEMM transform pre-pass (gemm_tc.cu), one buffer per stream so that the two branches never share
  float *prepass[2] = {nullptr, nullptr};
  size_t prepass_bytes = 0;
  // sample-resident plans (resident.cu): a record range [first, first+count) that runs as ONE kernel
  std::vector<ResidentPlan *> resident;
  int use_resident = 1;  // SLIDE_RESIDENT=0 keeps the per-record executor
};

namespace {

// PAIR launch-shape knobs: environment read once (slide_tc_reload_tuning() re-reads)
bool g_pair_tuning_loaded = false;
int g_pair_min_ctas = 2368, g_pair_min_rows = 32, g_pair_smem = 1, g_pair_pb = 8, g_pair_smem_min_ctas = 296;
int g_pair_prefetch = 1;  // SLIDE_PAIR_PREFETCH: residual rows are prefetched into L2 this many loop iterations ahead (0 = off).
                          // A/B on B200, batch 256, feature step: 0 -> 1486 us, 1 -> 1474, 2 -> 1477, 4 -> 1478, 8 -> 1490

template <typename T>
inline T *AP(slide_program *p, int64_t off) {
  return off < 0 ? nullptr : reinterpret_cast<T *>(p->arena + off);
}
template <typename T>
inline const T *WP(slide_program *p, int64_t off) {
  return off < 0 ? nullptr : reinterpret_cast<const T *>(p->weights + off);
}

inline unsigned grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  return (unsigned)(g < 1 ? 1 : g);
}

XFd make_xf(slide_program *p, const int64_t *f) {
  XFd x;
  x.stats = AP<double>(p, f[XF_STATS]);
  x.cg = (int)f[XF_CG];
  x.nnorm = (int)f[XF_NNORM];
  x.choff = (int)f[XF_CHOFF];
  x.gamma = WP<float>(p, f[XF_GAMMA_W]);
  x.beta = WP<float>(p, f[XF_BETA_W]);
  x.R = (int)f[XF_R];
  x.inv_count = f[XF_COUNT] > 0 ? 1.0f / (float)f[XF_COUNT] : 1.0f;
  x.relu = (int)f[XF_RELU];
  x.addvec = AP<float>(p, f[XF_ADDVEC]);
  x.addld = (int)f[XF_ADDLD];
  x.addmode = (int)f[XF_ADDMODE];
  return x;
}

int run_op(slide_program *p, const slide_op &op, cudaStream_t st) {
  const int64_t *q = op.p;
  switch (op.kind) {
    case SLIDE_OP_NOP:
    case SLIDE_OP_JOIN:
      return SLIDE_OK;
    case SLIDE_OP_STEP_BEGIN: {
      const size_t n16 = (size_t)q[SB_ZERO_BYTES] / 16;
      unsigned g = grid_for((long long)n16, 256);
      if (g > 1184) g = 1184;
      launch_k(step_begin_kernel, g, 256, 0, st, AP<uint4>(p, q[SB_ZERO_OFF]), n16, AP<int>(p, q[SB_STEP]));
      return after_launch();
    }
    case SLIDE_OP_KNN: {
      const int B = (int)q[KNN_B], P1 = (int)q[KNN_P1], P2 = (int)q[KNN_P2], K = (int)q[KNN_K];
      if (K > 32 || K > P2 || B > 65535) return SLIDE_ERR_UNSUPPORTED;
      const int threads = P1 >= 128 ? 128 : ((P1 + 31) / 32) * 32;
      dim3 grid(ceil_div(P1, threads), B);
      const float *qq = AP<float>(p, q[KNN_Q]), *rr = AP<float>(p, q[KNN_REF]);
      int *idx = AP<int>(p, q[KNN_IDX]);
      float *d2 = AP<float>(p, q[KNN_D2]);
      const int ldq = (int)q[KNN_LDQ], ldr = (int)q[KNN_LDR];
      if (K > 8 && P2 >= 128 && P1 >= 32) {  // K <= 8: the one-thread-per-query kernel's 8-step insertion is as fast (A/B: 213 vs 255 us)
        launch_k(knn_warp_kernel, dim3(ceil_div(P1, KW_WARPS * KW_QPW), B), dim3(KW_WARPS * 32), 0, st, qq, ldq, P1, rr, ldr, P2, K,
                 idx, d2);
        return after_launch();
      }
      if (K <= 4)
        launch_k(knn_prog_kernel<4>, grid, threads, 0, st, qq, ldq, P1, rr, ldr, P2, K, idx, d2);
      else if (K <= 8)
        launch_k(knn_prog_kernel<8>, grid, threads, 0, st, qq, ldq, P1, rr, ldr, P2, K, idx, d2);
      else if (K <= 16)
        launch_k(knn_prog_kernel<16>, grid, threads, 0, st, qq, ldq, P1, rr, ldr, P2, K, idx, d2);
      else
        launch_k(knn_prog_kernel<32>, grid, threads, 0, st, qq, ldq, P1, rr, ldr, P2, K, idx, d2);
      return after_launch();
    }
    case SLIDE_OP_GROUP: {
      const long long rows = (long long)q[GRP_B] * q[GRP_NP] * q[GRP_K];
      const int mode = (int)q[GRP_MODE];
      const int W = (int)q[GRP_C] + (mode == 0 ? 3 + (q[GRP_ABS] ? 3 : 0) + (q[GRP_CENTER] ? 3 : 0) : 11);
      if (W <= 16) {
        launch_k(group_elem_kernel, grid_for(rows * W, 256), 256, 0, st, 
            mode, AP<float>(p, q[GRP_F]), (int)q[GRP_LDF], (int)q[GRP_C], AP<float>(p, q[GRP_XYZ]), (int)q[GRP_LDX],
            (int)q[GRP_N], AP<float>(p, q[GRP_CTR]), (int)q[GRP_LDCTR], (int)q[GRP_NP], AP<int>(p, q[GRP_IDX]),
            (int)q[GRP_K], AP<float>(p, q[GRP_D2]), AP<float>(p, q[GRP_OUT]), (int)q[GRP_LDO], (int)q[GRP_ABS],
            (int)q[GRP_CENTER], rows, W);
      } else {
        unsigned g = grid_for(rows * 32, 256);
        if (g > 148 * 16) g = 148 * 16;
        launch_k(group_rows_kernel, g, 256, 0, st, 
            mode, AP<float>(p, q[GRP_F]), (int)q[GRP_LDF], (int)q[GRP_C], AP<float>(p, q[GRP_XYZ]), (int)q[GRP_LDX],
            (int)q[GRP_N], AP<float>(p, q[GRP_CTR]), (int)q[GRP_LDCTR], (int)q[GRP_NP], AP<int>(p, q[GRP_IDX]),
            (int)q[GRP_K], AP<float>(p, q[GRP_D2]), AP<float>(p, q[GRP_OUT]), (int)q[GRP_LDO], (int)q[GRP_ABS],
            (int)q[GRP_CENTER], rows);
      }
      return after_launch();
    }
    case SLIDE_OP_GEMM: {
      GemmArgs a;
      a.A = AP<float>(p, q[GEMM_A]);
      a.lda = (int)q[GEMM_LDA];
      a.M = (int)q[GEMM_M];
      a.K = (int)q[GEMM_K];
      a.W = WP<float>(p, q[GEMM_W_W]);
      a.ldw = (int)q[GEMM_LDW];
      a.N = (int)q[GEMM_N];
      a.C = AP<float>(p, q[GEMM_C]);
      a.ldc = (int)q[GEMM_LDC];
      a.bias = WP<float>(p, q[GEMM_BIAS_W]);
      a.act = (int)q[GEMM_ACT];
      a.ev = AP<float>(p, q[GEMM_EV]);
      a.evld = (int)q[GEMM_EVLD];
      a.evdiv = (int)q[GEMM_EVDIV] > 0 ? (int)q[GEMM_EVDIV] : 1;
      a.res = AP<float>(p, q[GEMM_RES]);
      a.ldr = (int)q[GEMM_LDR];
      a.st_stats = AP<double>(p, q[GEMM_ST_STATS]);
      a.st_cg = (int)q[GEMM_ST_CG] > 0 ? (int)q[GEMM_ST_CG] : 1;
      a.st_nnorm = (int)q[GEMM_ST_NNORM];
      a.st_choff = (int)q[GEMM_ST_CHOFF];
      a.st_R = (int)q[GEMM_ST_R] > 0 ? (int)q[GEMM_ST_R] : 1;
      a.st_weight = (float)q[GEMM_ST_WEIGHT];
      a.xfa = make_xf(p, q + GEMM_XFA);
      a.xfr = make_xf(p, q + GEMM_XFR);
      a.step = AP<int>(p, q[GEMM_STEP]);
      a.smk = (int)q[GEMM_SMK];
      if (a.smk > 0 && (!a.res || a.ev || a.st_stats || a.act || a.M % a.smk)) return SLIDE_ERR_INVALID;
      const float *wp = WP<float>(p, q[GEMM_WP_W]);
      if (p->gemm_backend == 0 && wp && gemm_tc_prepass_applicable(a) &&
          gemm_tc_prepass_bytes(a.M, a.K) <= p->prepass_bytes) {
        const int rc = launch_gemm_tc_prepass(a, wp, (int)q[GEMM_WP_NA], p->prepass[st == p->side ? 1 : 0], st);
        if (rc <= 0) return rc;  // > 0: not applicable after all -> the fused-transform paths below
      }
      if (p->gemm_backend == 0 && gemm_tc_eligible(a, wp)) return launch_gemm_tc(a, wp, (int)q[GEMM_WP_NA], st);
      return launch_gemm_simt(a, st);
    }
    case SLIDE_OP_SOFTMAX_WSUM: {
      const long long rows = q[SM_ROWS];
      const int C = (int)q[SM_C];
      launch_k(softmax_wsum_kernel, grid_for(rows * C, 256), 256, 0, st, 
          AP<float>(p, q[SM_S]), (int)q[SM_LDS], AP<float>(p, q[SM_V]), (int)q[SM_LDV], make_xf(p, q + SM_XFV),
          AP<int>(p, q[SM_STEP]), AP<float>(p, q[SM_OUT]), (int)q[SM_LDO], rows, (int)q[SM_K], C);
      return after_launch();
    }
    case SLIDE_OP_COPY_COLS: {
      const long long rows = q[CP_ROWS];
      const int n = (int)q[CP_NCOLS];
      launch_k(copy_cols_kernel, grid_for(rows * n, 256), 256, 0, st, AP<float>(p, q[CP_SRC]), (int)q[CP_LDS],
                                                               AP<float>(p, q[CP_DST]), (int)q[CP_LDD], rows, n);
      return after_launch();
    }
    case SLIDE_OP_DDPM_UPDATE: {
      const long long rows = q[DD_ROWS];
      const int n = (int)q[DD_NCOLS];
      launch_k(ddpm_update_kernel, grid_for(rows * n, 256), 256, 0, st, 
          (int)q[DD_MODE], AP<float>(p, q[DD_X]), (int)q[DD_LDX], AP<float>(p, q[DD_EPS]), (int)q[DD_LDE],
          AP<float>(p, q[DD_NOISE]), rows, n, (int)q[DD_COL0], WP<float>(p, q[DD_TABLE_W]), AP<int>(p, q[DD_STEP]),
          op.f[0], AP<float>(p, q[DD_X0C]), (int)q[DD_LDX0C], AP<float>(p, q[DD_MASK]));
      return after_launch();
    }
    case SLIDE_OP_FPS:
      return program_fps((int)q[FPS_MODE], AP<float>(p, q[FPS_XYZ]), (int)q[FPS_LDX], (int)q[FPS_B], (int)q[FPS_N],
                         (int)q[FPS_M], AP<int>(p, q[FPS_START]), AP<int>(p, q[FPS_OUT]), st);
    case SLIDE_OP_GATHER_ROWS: {
      const long long total = (long long)q[GA_B] * q[GA_M];
      const int n = (int)q[GA_NCOLS];
      launch_k(gather_rows_kernel, grid_for(total * n, 256), 256, 0, st, AP<float>(p, q[GA_SRC]), (int)q[GA_LDS],
                                                                  (int)q[GA_N], AP<int>(p, q[GA_IDX]), (int)q[GA_M],
                                                                  AP<float>(p, q[GA_DST]), (int)q[GA_LDD], n, total);
      return after_launch();
    }
    case SLIDE_OP_UPSAMPLE: {
      const long long rows = q[UP_ROWS];
      const int factor = (int)q[UP_FACTOR], Fd = (int)q[UP_F];
      launch_k(upsample_kernel, grid_for(rows * factor * Fd, 256), 256, 0, st, 
          AP<float>(p, q[UP_COARSE]), (int)q[UP_LDC], (int)q[UP_COARSE_C], AP<float>(p, q[UP_DISP]), (int)q[UP_LDD],
          AP<float>(p, q[UP_OUT]), (int)q[UP_LDO], rows, factor, Fd, op.f[0], op.f[1]);
      return after_launch();
    }
    case SLIDE_OP_TEMB: {
      const int rows = (int)q[TE_ROWS], half = (int)q[TE_HALF];
      launch_k(temb_kernel, grid_for((long long)rows * half, 256), 256, 0, st, 
          AP<float>(p, q[TE_TS]), WP<float>(p, q[TE_FREQ_W]), half, AP<float>(p, q[TE_OUT]), (int)q[TE_LDO], rows);
      return after_launch();
    }
    case SLIDE_OP_PAIR: {
      PairArgs a;
      a.U = AP<float>(p, q[PR_U]);
      a.ldu = (int)q[PR_LDU];
      a.nsrc = (int)q[PR_NSRC];
      a.xyz = AP<float>(p, q[PR_XYZ]);
      a.ldx = (int)q[PR_LDX];
      a.ctr = AP<float>(p, q[PR_CTR]);
      a.ldctr = (int)q[PR_LDCTR];
      a.np = (int)q[PR_NP];
      a.idx = AP<int>(p, q[PR_IDX]);
      a.K = (int)q[PR_K];
      a.d2 = AP<float>(p, q[PR_D2]);
      a.wx = WP<float>(p, q[PR_WX_W]);
      a.wc = WP<float>(p, q[PR_WC_W]);
      a.wd = WP<float>(p, q[PR_WD_W]);
      a.ww = WP<float>(p, q[PR_WW_W]);
      a.bias = WP<float>(p, q[PR_BIAS_W]);
      a.N = (int)q[PR_N];
      a.out = AP<float>(p, q[PR_OUT]);
      a.ldo = (int)q[PR_LDO];
      a.act = (int)q[PR_ACT];
      a.res = AP<float>(p, q[PR_RES]);
      a.ldr = (int)q[PR_LDR];
      a.st_stats = AP<double>(p, q[PR_ST_STATS]);
      a.st_cg = (int)q[PR_ST_CG] > 0 ? (int)q[PR_ST_CG] : 1;
      a.st_nnorm = (int)q[PR_ST_NNORM];
      a.st_choff = (int)q[PR_ST_CHOFF];
      a.st_weight = (float)q[PR_ST_WEIGHT];
      a.xfr = make_xf(p, q + PR_XFR);
      a.step = AP<int>(p, q[PR_STEP]);
      const int B = (int)q[PR_B];
      if (!a.U || !a.out || !a.idx || a.K <= 0 || B <= 0 || B > 65535) return SLIDE_ERR_INVALID;
      if (a.d2 && (!a.wd || !a.ww)) return SLIDE_ERR_INVALID;
      if (a.res && a.xfr.stats && a.xfr.R != a.np * a.K) return SLIDE_ERR_UNSUPPORTED;
      if (((uintptr_t)a.U & 15) || (a.ldu & 3) || ((uintptr_t)a.out & 15) || (a.ldo & 3) ||
          (a.res && (((uintptr_t)a.res & 15) || (a.ldr & 3))))
        return SLIDE_ERR_UNSUPPORTED;  // 128-bit row accesses
      if (!g_pair_tuning_loaded) {
        const char *e_ctas = getenv("SLIDE_PAIR_MIN_CTAS"), *e_rows = getenv("SLIDE_PAIR_MIN_ROWS"),
                   *e_smem = getenv("SLIDE_PAIR_SMEM");
        g_pair_min_ctas = e_ctas ? atoi(e_ctas) : 2368;
        g_pair_min_rows = e_rows ? atoi(e_rows) : 32;
        g_pair_smem = e_smem ? atoi(e_smem) : 1;
        const char *e_pb = getenv("SLIDE_PAIR_PB");
        g_pair_pb = e_pb ? atoi(e_pb) : 8;
        if (g_pair_pb < 1) g_pair_pb = 1;
        const char *e_mc = getenv("SLIDE_PAIR_SMEM_MIN_CTAS");
        g_pair_smem_min_ctas = e_mc ? atoi(e_mc) : 296;
        const char *e_pf = getenv("SLIDE_PAIR_PREFETCH");
        g_pair_prefetch = e_pf ? atoi(e_pf) : 1;
        g_pair_tuning_loaded = true;
      }
      // small gather source (the denoisers: the sample's own 16 points): stage it in shared memory
      if (!a.wc) return SLIDE_ERR_INVALID;
      if (a.wx && g_pair_smem && a.nsrc <= PS_MAX_SRC && a.N <= 4 * PS_THREADS) {
        const int CT = ceil_div(a.N, 4), RG = PS_THREADS / CT, ldn = 4 * CT;
        int pbs = a.np < g_pair_pb ? a.np : g_pair_pb;  // points per CTA (8: 128 rows at K = 16; A/B on B200, batch 256, position / feature step: 2 -> 829 / 1650 us, 4 -> 710 / 1525, 8 -> 658 / 1488, 16 -> 654 / 1491); fewer when the grid would be under 2 CTAs per SM
        while (pbs > 1 && (long long)B * ceil_div(a.np, pbs) < g_pair_smem_min_ctas) pbs = (pbs + 1) / 2;
        const size_t smem = ((size_t)(a.nsrc + pbs + 2 * RG) * ldn + 3 * (size_t)pbs * a.K) * 4;
        if (smem <= 160 * 1024) {
          a.pb = pbs;
          a.pf = g_pair_prefetch;
          if (a.res) return a.d2 ? launch_pair_smem<true, true>(a, B, CT, RG, ldn, smem, st)
                                 : launch_pair_smem<true, false>(a, B, CT, RG, ldn, smem, st);
          return a.d2 ? launch_pair_smem<false, true>(a, B, CT, RG, ldn, smem, st)
                      : launch_pair_smem<false, false>(a, B, CT, RG, ldn, smem, st);
        }
      }
      // points per CTA: aim for >= 16 CTAs per SM overall, at least 32 rows per CTA
      int pb = a.np;
      // A/B on B200 (feature-DDPM step): (592 CTAs, 64 rows) 1764 us, (2368, 32) 1750 us, (4736, 16) 1784 us
      const int min_ctas = g_pair_min_ctas, min_rows = g_pair_min_rows;
      while (pb > 1 && (long long)B * ceil_div(a.np, pb) < min_ctas && pb * a.K > min_rows) pb = (pb + 1) / 2;
      a.pb = pb;
      const int cols4 = ceil_div(a.N, 4);
      const int threads = cols4 >= 256 ? 256 : ((cols4 + 31) / 32) * 32;
      dim3 grid(ceil_div(a.np, pb), B);
      if (!a.wx) {  // coordinate term already inside U
        if (a.res)
          launch_k(pair_kernel<4, true, true>, grid, threads, 0, st, a);
        else
          launch_k(pair_kernel<8, false, true>, grid, threads, 0, st, a);
      } else if (a.res) {
        launch_k(pair_kernel<4, true, false>, grid, threads, 0, st, a);
      } else {
        launch_k(pair_kernel<8, false, false>, grid, threads, 0, st, a);
      }
      return after_launch();
    }
    case SLIDE_OP_COLMAX: {
      const int B = (int)q[CM_B], C = (int)q[CM_C];
      XFd xf = make_xf(p, q + CM_XF);
      if (xf.stats && xf.R != (int)q[CM_R]) return SLIDE_ERR_UNSUPPORTED;
      launch_k(colmax_kernel, grid_for((long long)B * C, 128), 128, 0, st, AP<float>(p, q[CM_X]), (int)q[CM_LDX], (int)q[CM_R], C,
                                                                   xf, AP<int>(p, q[CM_STEP]), AP<float>(p, q[CM_OUT]),
                                                                   (int)q[CM_LDO], B);
      return after_launch();
    }
    case SLIDE_OP_KL: {
      const long long rows = q[KL_ROWS];
      const int C = (int)q[KL_C];
      launch_k(kl_kernel, grid_for(rows * C, 256), 256, 0, st, AP<float>(p, q[KL_P]), (int)q[KL_LDP], C,
                                                         AP<float>(p, q[KL_NOISE]), (int)q[KL_LDN],
                                                         AP<float>(p, q[KL_OUT]), (int)q[KL_LDO], rows);
      return after_launch();
    }
    default:
      return SLIDE_ERR_INVALID;
  }
}

// Main-stream records go to `st`; SIDE records go to p->side.  The side stream forks lazily (it waits for everything
// enqueued on `st` before the region's first SIDE record) and is joined at a JOIN record or at the end of the range,
// so any sub-range (the tests run single records) is self-contained.
const ResidentPlan *resident_for(slide_program *p, int first, int count) {
  if (!p->use_resident || p->gemm_backend != 0) return nullptr;
  for (const ResidentPlan *r : p->resident)
    if (resident_first(r) == first && resident_count(r) == count) return r;
  return nullptr;
}

int run_range(slide_program *p, int first, int count, cudaStream_t st) {
  if (!p || first < 0 || count < 0 || (size_t)(first + count) > p->ops.size()) return SLIDE_ERR_INVALID;
  if (const ResidentPlan *r = resident_for(p, first, count)) return resident_launch(r, p->arena, p->weights, st);
  bool side_live = false;
  auto join = [&]() -> int {
    if (!side_live) return SLIDE_OK;
    side_live = false;
    int rc = cuda_rc(cudaEventRecord(p->ev_join, p->side));
    if (rc == SLIDE_OK) rc = cuda_rc(cudaStreamWaitEvent(st, p->ev_join, 0));
    return rc;
  };
  for (int i = first; i < first + count; ++i) {
    const slide_op &op = p->ops[i];
    int rc = SLIDE_OK;
    if (op.kind == SLIDE_OP_JOIN) {
      rc = join();
    } else if ((op.flags & SLIDE_OPF_SIDE) && p->use_side && p->side) {
      if (!side_live) {
        rc = cuda_rc(cudaEventRecord(p->ev_fork, st));
        if (rc == SLIDE_OK) rc = cuda_rc(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
        side_live = true;
      }
      if (rc == SLIDE_OK) rc = run_op(p, op, p->side);
    } else {
      rc = run_op(p, op, st);
    }
    if (rc != SLIDE_OK) {
      join();
      return rc;
    }
  }
  return join();
}

}  // namespace

extern "C" {

int slide_program_create(const struct slide_op *ops, int n_ops, size_t arena_bytes, const void *weights,
                         size_t weights_bytes, slide_program **out) {
  if (!ops || n_ops < 0 || !out) return SLIDE_ERR_INVALID;
  slide_program *p = new slide_program();
  p->ops.assign(ops, ops + n_ops);
  p->arena_bytes = arena_bytes;
  p->weights_bytes = weights_bytes;
  const char *be = getenv("SLIDE_GEMM_BACKEND");
  if (be && strcmp(be, "simt") == 0) p->gemm_backend = 1;
  const char *sb = getenv("SLIDE_SIDE_BRANCH");
  if (sb && atoi(sb) == 0) p->use_side = 0;
  const char *rs = getenv("SLIDE_RESIDENT");
  if (rs && atoi(rs) == 0) p->use_resident = 0;
  int rc = cuda_rc(cudaMalloc((void **)&p->arena, arena_bytes > 0 ? arena_bytes : 256));
  if (rc == SLIDE_OK) rc = cuda_rc(cudaMemset(p->arena, 0, arena_bytes));
  if (rc == SLIDE_OK) rc = cuda_rc(cudaMalloc((void **)&p->weights, weights_bytes > 0 ? weights_bytes : 256));
  if (rc == SLIDE_OK && weights && weights_bytes)
    rc = cuda_rc(cudaMemcpy(p->weights, weights, weights_bytes, cudaMemcpyHostToDevice));
  // scratch for the GEMM transform pre-pass: the largest point-level A operand that qualifies (cheap: <= a few MB)
  for (const slide_op &op : p->ops) {
    if (op.kind != SLIDE_OP_GEMM || op.p[GEMM_WP_W] < 0) continue;
    const int64_t *xf = op.p + GEMM_XFA;
    const bool has = xf[XF_STATS] >= 0 || xf[XF_ADDVEC] >= 0 || xf[XF_RELU] != 0;
    const int R = (int)xf[XF_R];
    if (!has || R <= 0 || R % 128 == 0 || 128 % R != 0) continue;
    const size_t need = gemm_tc_prepass_bytes((int)op.p[GEMM_M], (int)op.p[GEMM_K]);
    if (need > p->prepass_bytes && need <= ((size_t)256 << 20)) p->prepass_bytes = need;
  }
  for (int i = 0; i < 2 && rc == SLIDE_OK && p->prepass_bytes; ++i)
    rc = cuda_rc(cudaMalloc((void **)&p->prepass[i], p->prepass_bytes));
  if (rc == SLIDE_OK) rc = cuda_rc(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
  if (rc == SLIDE_OK) rc = cuda_rc(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
  if (rc == SLIDE_OK) rc = cuda_rc(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
  if (rc != SLIDE_OK) {
    slide_program_destroy(p);
    return rc;
  }
  *out = p;
  return SLIDE_OK;
}

void slide_program_destroy(slide_program *p) {
  if (!p) return;
  for (int i = 0; i < 8; ++i)
    if (p->graphs[i]) cudaGraphExecDestroy(p->graphs[i]);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  if (p->side) cudaStreamDestroy(p->side);
  if (p->arena) cudaFree(p->arena);
  if (p->weights) cudaFree(p->weights);
  for (int i = 0; i < 2; ++i)
    if (p->prepass[i]) cudaFree(p->prepass[i]);
  for (ResidentPlan *r : p->resident) resident_free(r);
  delete p;
}

int slide_program_set_resident(slide_program *p, const struct slide_resident_plan *plan, const struct slide_rop *rops,
                               int n_rops) {
  if (!p || !plan || !rops || n_rops <= 0) return SLIDE_ERR_INVALID;
  if (plan->first < 0 || plan->count <= 0 || (size_t)(plan->first + plan->count) > p->ops.size()) return SLIDE_ERR_INVALID;
  ResidentPlan *r = nullptr;
  const int rc = resident_create(plan, rops, n_rops, &r);
  if (rc != SLIDE_OK) return rc;
  for (ResidentPlan *&old : p->resident)
    if (resident_first(old) == plan->first && resident_count(old) == plan->count) {
      resident_free(old);
      old = r;
      return SLIDE_OK;
    }
  p->resident.push_back(r);
  return SLIDE_OK;
}

int slide_program_use_resident(slide_program *p, int enable) {
  if (!p) return SLIDE_ERR_INVALID;
  p->use_resident = enable ? 1 : 0;
  return SLIDE_OK;
}

void *slide_program_arena(slide_program *p) { return p ? p->arena : nullptr; }
void *slide_program_weights(slide_program *p) { return p ? p->weights : nullptr; }

int slide_program_set_gemm_backend(slide_program *p, int backend) {
  if (!p || backend < 0 || backend > 1) return SLIDE_ERR_INVALID;
  p->gemm_backend = backend;
  return SLIDE_OK;
}

int slide_program_run(slide_program *p, int first, int count, slide_stream_t stream) {
  return run_range(p, first, count, (cudaStream_t)stream);
}

int slide_program_capture(slide_program *p, int slot, int first, int count, int repeat, slide_stream_t stream) {
  if (!p || slot < 0 || slot >= 8 || repeat < 1) return SLIDE_ERR_INVALID;
  (void)stream;  // capture records work, it does not execute it: use a private stream (the legacy default
                 // stream, which is what callers usually run on, cannot be captured)
  cudaStream_t st = nullptr;
  int rc = cuda_rc(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  if (rc != SLIDE_OK) return rc;
  if (p->graphs[slot]) {
    cudaGraphExecDestroy(p->graphs[slot]);
    p->graphs[slot] = nullptr;
  }
  const long long before = g_launch_count;
  rc = cuda_rc(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  if (rc != SLIDE_OK) {
    cudaStreamDestroy(st);
    return rc;
  }
  int run_rc = SLIDE_OK;
  for (int r = 0; r < repeat && run_rc == SLIDE_OK; ++r) run_rc = run_range(p, first, count, st);
  cudaGraph_t graph = nullptr;
  rc = cuda_rc(cudaStreamEndCapture(st, &graph));
  cudaStreamDestroy(st);
  p->graph_launches[slot] = (int)(g_launch_count - before);
  g_launch_count = before;  // captured, not launched
  if (run_rc != SLIDE_OK || rc != SLIDE_OK) {
    if (graph) cudaGraphDestroy(graph);
    return run_rc != SLIDE_OK ? run_rc : rc;
  }
  rc = cuda_rc(cudaGraphInstantiate(&p->graphs[slot], graph, 0));
  cudaGraphDestroy(graph);
  return rc;
}

int slide_program_replay(slide_program *p, int slot, int times, slide_stream_t stream) {
  if (!p || slot < 0 || slot >= 8 || !p->graphs[slot] || times < 0) return SLIDE_ERR_INVALID;
  for (int i = 0; i < times; ++i) {
    const int rc = cuda_rc(cudaGraphLaunch(p->graphs[slot], (cudaStream_t)stream));
    if (rc != SLIDE_OK) return rc;
    g_launch_count += p->graph_launches[slot];
  }
  return SLIDE_OK;
}

int slide_tc_error(void) { return tc_error_flag(); }
void slide_tc_reset_error(void) { tc_error_reset(); }
void slide_tc_reload_tuning(void) {
  tc_reload_tuning();
  g_pair_tuning_loaded = false;
}

int slide_program_launches(slide_program *p, int first, int count) {
  if (!p || first < 0 || count < 0 || (size_t)(first + count) > p->ops.size()) return SLIDE_ERR_INVALID;
  if (resident_for(p, first, count)) return 1;
  int n = 0;
  for (int i = first; i < first + count; ++i)
    if (p->ops[i].kind != SLIDE_OP_NOP && p->ops[i].kind != SLIDE_OP_JOIN) ++n;
  return n;
}

}  // extern "C"
