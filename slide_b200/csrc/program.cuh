// Device-side argument blocks shared by the program executor and the GEMM kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slide_program.h"

namespace slide {

// Transform-on-load (XF block of slide_program.h) with pointers resolved.
struct XFd {
  const double *stats;  // [B, nnorm/cg, 2] sums, nullptr = no normalisation
  int cg, nnorm, choff;
  const float *gamma, *beta;
  int R;            // rows per sample of the tensor being loaded
  float inv_count;  // 1 / elements per (sample, group)
  int relu;
  const float *addvec;
  int addld, addmode;
};

struct GemmArgs {
  const float *A;
  int lda, M, K;
  const float *W;
  int ldw, N;
  float *C;
  int ldc;
  const float *bias;
  int act;
  const float *ev;
  int evld, evdiv;
  const float *res;
  int ldr;
  double *st_stats;
  int st_cg, st_nnorm, st_choff, st_R;
  float st_weight;
  XFd xfa, xfr;
  const int *step;
  int smk;  // > 0: fused soft-max over groups of smk rows, applied to xfr(res); C has M / smk rows (GEMM_SMK)
};

// Per-CTA table of (mean, rstd) for the samples a row tile touches: [XF_MAXS][XF_MAXG].
constexpr int XF_MAXS = 8;
constexpr int XF_MAXG = 32;

// Fill `tab` for samples [s0, s0 + ns) (ns <= XF_MAXS).  All threads of the CTA call this; the caller syncs.
__device__ __forceinline__ void xf_fill_table(const XFd &x, float2 *tab, int s0, int ns, int tid, int nthreads) {
  if (!x.stats) return;
  const int G = x.nnorm / x.cg;
  for (int e = tid; e < ns * G; e += nthreads) {
    const int sl = e / G, g = e - sl * G;
    const double *st = x.stats + ((size_t)(s0 + sl) * G + g) * 2;
    const double m = st[0] * (double)x.inv_count;
    double var = st[1] * (double)x.inv_count - m * m;
    var = var < 0.0 ? 0.0 : var;
    tab[sl * XF_MAXG + g] = make_float2((float)m, (float)(1.0 / sqrt(var + (double)SLIDE_GN_EPS)));
  }
}

// Apply the transform to one element of sample `s` (= row / x.R, computed once per row by the caller); `col` is
// the local column of the loaded tensor.
__device__ __forceinline__ float xf_apply(const XFd &x, const float2 *tab, int s0, int s, int col, float v, int step) {
  if (x.stats) {
    const int ch = x.choff + col;
    if (ch < x.nnorm) {
      const float2 mr = tab[(s - s0) * XF_MAXG + ch / x.cg];
      v = (v - mr.x) * mr.y * __ldg(x.gamma + ch) + __ldg(x.beta + ch);
    }
  }
  if (x.relu) v = fmaxf(v, 0.f);
  if (x.addvec) {
    const long long arow = x.addmode == 0 ? s : (x.addmode == 1 ? step : 0);
    v += __ldg(x.addvec + arow * x.addld + col);
  }
  return v;
}

__device__ __forceinline__ float act_apply(int act, float v) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v / (1.0f + expf(-v));  // swish: x * sigmoid(x)
  return v;
}

int launch_gemm_simt(const GemmArgs &a, cudaStream_t st);
bool gemm_tc_eligible(const GemmArgs &a, const float *Wp);
int launch_gemm_tc(const GemmArgs &a, const float *Wp, int wp_na, cudaStream_t st);
// Row tiles of 128 that span several samples (point-level tensors: 16 rows per sample): apply the A transform in a
// small elementwise pre-pass into `scratch` (rows of round4(K) floats) so that the GEMM takes the TMA-fed path.
bool gemm_tc_prepass_applicable(const GemmArgs &a);
size_t gemm_tc_prepass_bytes(int M, int K);
int launch_gemm_tc_prepass(const GemmArgs &a, const float *Wp, int wp_na, float *scratch, cudaStream_t st);
int tc_error_flag();
void tc_error_reset();
void tc_reload_tuning();
// sample-resident plans (resident.cu)
struct ResidentPlan;
}  // namespace slide
struct slide_resident_plan;
struct slide_rop;
namespace slide {
int resident_create(const slide_resident_plan *hdr, const slide_rop *rops, int n_rops, ResidentPlan **out);
void resident_free(ResidentPlan *p);
int resident_launch(const ResidentPlan *p, char *arena, const char *weights, cudaStream_t st);
int resident_first(const ResidentPlan *p);
int resident_count(const ResidentPlan *p);
int program_fps(int mode, const float *xyz, int ldx, int B, int N, int m, const int *start, int *out, cudaStream_t st);

}  // namespace slide
