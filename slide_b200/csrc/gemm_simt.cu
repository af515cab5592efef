// fp32 FFMA GEMM with the fused prologue / epilogue of SLIDE_OP_GEMM (see slide_program.h).
//
// This is the any-shape kernel: tiny K (the 12-channel position-DDPM inputs), N that is not a multiple of 8
// (3-channel heads), unaligned column views, ragged M.  Dense, aligned contractions go to gemm_tc.cu (tcgen05).
//   C = act( xfA(A) W^T + bias + ev[row / evdiv] + xfR(resid) ),  GroupNorm statistics of C accumulated in fp64,
//   or (GEMM_SMK) C[g] = sum_k xfR(resid)[g*K+k] * softmax_k(xfA(A) W^T + bias)[g*K+k].
// The 64x64 accumulator tile is staged in shared memory so that the epilogue runs with thread = column:
// coalesced stores / residual loads, per-column statistics in registers, soft-max down the rows.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "program.cuh"

namespace slide {

constexpr int SBM = 64, SBN = 64, SBK = 16, STHREADS = 256;

__global__ void __launch_bounds__(STHREADS) gemm_simt_kernel(GemmArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Ws[SBK][SBN + 4];
  __shared__ float tile[SBM][SBN + 1];
  __shared__ float2 tabA[XF_MAXS * XF_MAXG];
  __shared__ float2 tabR[XF_MAXS * XF_MAXG];
  __shared__ float stacc[XF_MAXS * XF_MAXG * 2];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
  const int mlast = min(m0 + SBM, a.M) - 1;
  const int step = a.step ? *a.step : 0;

  const int sA0 = a.xfa.stats ? m0 / a.xfa.R : 0;
  const int sR0 = a.xfr.stats ? m0 / a.xfr.R : 0;
  const int sS0 = a.st_stats ? m0 / a.st_R : 0;
  if (a.xfa.stats) xf_fill_table(a.xfa, tabA, sA0, mlast / a.xfa.R - sA0 + 1, tid, STHREADS);
  if (a.res && a.xfr.stats) xf_fill_table(a.xfr, tabR, sR0, mlast / a.xfr.R - sR0 + 1, tid, STHREADS);
  if (a.st_stats)
    for (int e = tid; e < XF_MAXS * XF_MAXG * 2; e += STHREADS) stacc[e] = 0.f;
  __syncthreads();

  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: thread -> (row = tid / 4, 4 consecutive k starting at (tid % 4) * 4)
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int lm = m0 + lr;
  const int lsA = (lm < a.M) ? lm / a.xfa.R : 0;  // sample of the row this thread loads (fixed over the K loop)
  for (int k0 = 0; k0 < a.K; k0 += SBK) {
    {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + lk + u;
        float v = 0.f;
        if (lm < a.M && k < a.K) v = xf_apply(a.xfa, tabA, sA0, lsA, k, a.A[(size_t)lm * a.lda + k], step);
        As[lk + u][lr] = v;
      }
      const int n = n0 + lr;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + lk + u;
        Ws[lk + u][lr] = (n < a.N && k < a.K) ? __ldg(a.W + (size_t)n * a.ldw + k) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      const float4 av = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      const float4 wv = *reinterpret_cast<const float4 *>(&Ws[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }

  // stage the accumulators: tile[row][col]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) tile[ty * 4 + i][tx * 4 + j] = acc[i][j];
  __syncthreads();

  tile_epilogue(a, tile, tabR, stacc, m0, n0, mlast, sR0, sS0, step, tid);
}

static bool spans_ok(int R, int bm) { return (bm - 1) / R + 2 <= XF_MAXS; }

int launch_gemm_simt(const GemmArgs &a, cudaStream_t st) {
  if (!a.A || !a.W || !a.C || a.M <= 0 || a.N <= 0 || a.K <= 0) return SLIDE_ERR_INVALID;
  if (a.xfa.stats && (!spans_ok(a.xfa.R, SBM) || a.xfa.nnorm / a.xfa.cg > XF_MAXG)) return SLIDE_ERR_UNSUPPORTED;
  if (a.res && a.xfr.stats && (!spans_ok(a.xfr.R, SBM) || a.xfr.nnorm / a.xfr.cg > XF_MAXG))
    return SLIDE_ERR_UNSUPPORTED;
  if (a.st_stats && (!spans_ok(a.st_R, SBM) || a.st_nnorm / a.st_cg > XF_MAXG)) return SLIDE_ERR_UNSUPPORTED;
  if (a.smk > 0 && (SBM % a.smk != 0)) return SLIDE_ERR_UNSUPPORTED;  // neighbour groups must not straddle row tiles
  dim3 grid(ceil_div(a.M, SBM), ceil_div(a.N, SBN));
  if (grid.y > 65535) return SLIDE_ERR_UNSUPPORTED;
  launch_k(gemm_simt_kernel, grid, STHREADS, 0, st, a);
  return after_launch();
}

}  // namespace slide
