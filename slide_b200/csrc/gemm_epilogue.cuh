// Epilogue shared by the 64x64-tile GEMM kernels (gemm_simt.cu: fp32 FFMA, gemm_mma.cu: mma.sync TF32): the accumulator
// tile is staged in shared memory so that the epilogue runs with thread = column -- coalesced stores / residual loads,
// per-column GroupNorm statistics in registers, the fused soft-max down the rows (see slide_program.h, SLIDE_OP_GEMM).
#pragma once
#include "common.cuh"
#include "program.cuh"

namespace slide {

constexpr int EBM = 64, EBN = 64, ETHREADS = 256;

// tile[row][col] holds xfA(A) W^T for rows m0.., columns n0..; tabR: (mean, rstd) table of the residual's transform for
// the samples of this row tile (first sample sR0); stacc: zeroed [XF_MAXS][XF_MAXG][2] scratch; sS0: first sample of the
// statistics rows.  All ETHREADS threads call this after a __syncthreads() that made `tile` visible.
__device__ __forceinline__ void tile_epilogue(const GemmArgs &a, float (*tile)[EBN + 1], const float2 *tabR, float *stacc,
                                              int m0, int n0, int mlast, int sR0, int sS0, int step, int tid) {
  constexpr int SBM = EBM, STHREADS = ETHREADS;
  (void)SBM;
  // epilogue: thread = (column c, row quarter rq)
  const int c = tid & 63, rq = tid >> 6;
  const int n = n0 + c;
  const bool ncol = n < a.N;
  const float bias = (ncol && a.bias) ? __ldg(a.bias + n) : 0.f;
  const int rows_here = mlast - m0 + 1;
  if (a.smk > 0) {
    // fused AttentionModule tail: soft-max over groups of smk rows, reduce the transformed values with it
    const int K = a.smk;
    if (ncol) {
      for (int g0 = rq * K; g0 + K <= rows_here; g0 += 4 * K) {
        float mx = -INFINITY;
        for (int k = 0; k < K; ++k) mx = fmaxf(mx, tile[g0 + k][c]);
        float den = 0.f, o = 0.f;
        for (int k = 0; k < K; ++k) {
          const int m = m0 + g0 + k;
          const float w = expf(tile[g0 + k][c] - mx);  // the bias is common to the group: it cancels in the soft-max
          const float x = xf_apply(a.xfr, tabR, sR0, m / a.xfr.R, n, a.res[(size_t)m * a.ldr + n], step);
          den += w;
          o = fmaf(x, w, o);
        }
        a.C[(size_t)((m0 + g0) / K) * a.ldc + n] = o / den;
      }
    }
    return;
  }
  {
    const int r0 = rq * 16;
    const int ch = a.st_choff + n;
    const bool dost = a.st_stats && ncol && ch < a.st_nnorm;
    float ssum = 0.f, ssq = 0.f;
    int scur = -1;
    for (int rr = 0; rr < 16; ++rr) {
      const int m = m0 + r0 + rr;
      if (m >= a.M) break;
      if (ncol) {
        float v = tile[r0 + rr][c] + bias;
        if (a.ev) v += a.ev[(size_t)(m / a.evdiv) * a.evld + n];
        if (a.res) v += xf_apply(a.xfr, tabR, sR0, m / a.xfr.R, n, a.res[(size_t)m * a.ldr + n], step);
        v = act_apply(a.act, v);
        a.C[(size_t)m * a.ldc + n] = v;
        if (dost) {
          const int sm = m / a.st_R - sS0;
          if (sm != scur) {
            if (scur >= 0) {
              float *slot = stacc + (scur * XF_MAXG + ch / a.st_cg) * 2;
              atomicAdd(slot, ssum);
              atomicAdd(slot + 1, ssq);
            }
            scur = sm;
            ssum = 0.f;
            ssq = 0.f;
          }
          ssum += v;
          ssq = fmaf(v, v, ssq);
        }
      }
    }
    if (dost && scur >= 0) {
      float *slot = stacc + (scur * XF_MAXG + ch / a.st_cg) * 2;
      atomicAdd(slot, ssum);
      atomicAdd(slot + 1, ssq);
    }
  }
  if (a.st_stats) {
    __syncthreads();
    const int G = a.st_nnorm / a.st_cg;
    const int ns = mlast / a.st_R - sS0 + 1;
    for (int e = tid; e < ns * G * 2; e += STHREADS) {
      const int sl = e / (G * 2), rem = e - sl * G * 2;
      const float v = stacc[(sl * XF_MAXG + (rem >> 1)) * 2 + (rem & 1)];
      if (v != 0.f) atomicAdd(a.st_stats + ((size_t)(sS0 + sl) * G) * 2 + rem, (double)v * (double)a.st_weight);
    }
  }
}

}  // namespace slide
