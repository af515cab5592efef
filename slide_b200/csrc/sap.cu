// SAP mesh-reconstruction path: everything between the refinement network and marching cubes,
// hand-written for sm_100a.  C ABI: include/slide_sap.h.
//
//   slide_sap_mirror_concat   data_utils/mirror_partial.py:8-58 (mirror through the centroid plane, +-1 label, permutation)
//   slide_sap_unit_cube       dpsr_evaluation.py:22-32,72-76   (bounding-box normalisation, x/1.2+0.5, clamp)
//   slide_dpsr_forward        dpsr_utils/dpsr.py:30-77          (differentiable Poisson solver, forward only)
//
// DPSR is HBM-bound integer/float streaming work (a 128^3 grid is 8 MB per channel): the reference runs it as ~60 eager
// torch ops around three cuFFT calls.  Here it is 8 launches per batch:
//   memset -> splat (RED.ADD) -> Z pass (two real lines packed into one complex FFT) -> Y pass -> X pass FUSED with the
//   spectral solve and the inverse X pass (the three normal channels are read once, one potential channel is written)
//   -> inverse Y -> inverse Z (two Hermitian lines per complex FFT, real output) -> trilinear read-back + mean ->
//   shift / scale.  Every FFT is a shared-memory radix-2 pass over a tile of lines; global accesses are contiguous along
//   the innermost (kz) axis in every pass.  Algorithmic bytes per sample at R = 128: 25 (splat target) + 25 + 26 (Z)
//   + 2 x 26 (Y) + 26 + 8.5 (X + solve) + 2 x 8.5 (Y^-1) + 8.5 + 8.4 (Z^-1) + 2 x 8.4 (finalise) = 213 MB.
#include <math.h>

#include "../../include/slide_sap.h"
#include "common.cuh"

namespace slide {

namespace {

constexpr int FFT_THREADS = 256;

// ---- FFT building blocks ------------------------------------------------------------------------------------------------
// A line of R = R1 * R2 complex values is transformed in two register-resident steps (Cooley-Tukey, four-step form):
//   input index  n = R2 n1 + n2   held in shared memory at  n1 * P + n2      ("natural" layout, P = R2 + 1)
//   step 1: thread (line, n2): radix-R1 transform over n1 in registers, times W_R^(n2 k1), back to the same positions
//   step 2: thread (line, k1): radix-R2 transform over n2 in registers, back to the same positions
//   output index k = k1 + R1 k2   held at  k1 * P + k2                          ("transposed" layout)
// -- two shared-memory round trips and two barriers per transform instead of log2(R) of each (the first version of this
// file ran a radix-2 pass per barrier and was bound by instruction issue and bank conflicts: ncu 67-75 % SM busy at
// 1.6 TB/s; profiles/r02_sap_*).  The mirrored form (transposed layout in, natural layout out) exists for the inverse
// transform that directly follows a forward one in solve_x_kernel, so the spectrum is consumed where it lies.
template <int LOGR>
struct Geo {
  static constexpr int R = 1 << LOGR, L1 = LOGR / 2, R1 = 1 << L1, L2 = LOGR - L1, R2 = 1 << L2, P = R2 + 1;
  static constexpr int LS = (R1 * P) | 1;  // odd line stride: the transposing global<->shared copies stay conflict-free
  __device__ static __forceinline__ int in_pos(int i) { return (i >> L2) * P + (i & (R2 - 1)); }
  __device__ static __forceinline__ int out_pos(int k) { return (k & (R1 - 1)) * P + (k >> L1); }
};

__device__ constexpr float COS16[8] = {1.f, 0.92387953251f, 0.70710678119f, 0.38268343237f,
                                       0.f, -0.38268343237f, -0.70710678119f, -0.92387953251f};
__device__ constexpr float SIN16[8] = {0.f, 0.38268343237f, 0.70710678119f, 0.92387953251f,
                                       1.f, 0.92387953251f, 0.70710678119f, 0.38268343237f};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// N-point transform (N <= 16) of a register array, natural order in and out: X[k] = sum_n x[n] W_N^(+-nk).
template <int N, bool INV>
__device__ __forceinline__ void reg_fft(float2 (&v)[N]) {
  if constexpr (N == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = make_float2(a.x + b.x, a.y + b.y);
    v[1] = make_float2(a.x - b.x, a.y - b.y);
  } else if constexpr (N > 2) {
    float2 e[N / 2], o[N / 2];
#pragma unroll
    for (int i = 0; i < N / 2; ++i) e[i] = v[2 * i], o[i] = v[2 * i + 1];
    reg_fft<N / 2, INV>(e);
    reg_fft<N / 2, INV>(o);
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      constexpr int step = 16 / N;
      const int j = k * step;  // W_N^k = W_16^j
      float2 t;
      if (j == 0)
        t = o[k];
      else if (j == 4)
        t = INV ? make_float2(-o[k].y, o[k].x) : make_float2(o[k].y, -o[k].x);  // -+ i
      else
        t = cmul(o[k], make_float2(COS16[j], INV ? SIN16[j] : -SIN16[j]));
      v[k] = make_float2(e[k].x + t.x, e[k].y + t.y);
      v[k + N / 2] = make_float2(e[k].x - t.x, e[k].y - t.y);
    }
  }
}

// 16-byte asynchronous global -> shared copies (LDGSTS): the next tile of a persistent CTA is in flight while the current
// one is transformed; they hold no registers and complete in groups.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// tw[j] = exp(-2 pi i j / R), j < R
template <int LOGR>
__device__ __forceinline__ void fill_twiddles(float2 *tw) {
  constexpr int R = 1 << LOGR;
  for (int j = threadIdx.x; j < R; j += blockDim.x) {
    float sn, cs;
    sincospif(2.0f * (float)j / (float)R, &sn, &cs);
    tw[j] = make_float2(cs, -sn);
  }
}

// natural layout in -> transposed layout out.  Callers synchronise before; ends with a barrier.
template <int LOGR, bool INV>
__device__ __forceinline__ void fft_lines(float2 *s, const float2 *tw, int lines) {
  using G = Geo<LOGR>;
  for (int t = threadIdx.x; t < lines * G::R2; t += blockDim.x) {
    const int line = t >> G::L2, n2 = t & (G::R2 - 1);
    float2 *p = s + line * G::LS + n2;
    float2 v[G::R1];
#pragma unroll
    for (int j = 0; j < G::R1; ++j) v[j] = p[j * G::P];
    reg_fft<G::R1, INV>(v);
    p[0] = v[0];
#pragma unroll
    for (int k1 = 1; k1 < G::R1; ++k1) {
      float2 w = tw[n2 * k1];
      if (INV) w.y = -w.y;
      p[k1 * G::P] = cmul(v[k1], w);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < lines * G::R1; t += blockDim.x) {
    const int line = t >> G::L1, k1 = t & (G::R1 - 1);
    float2 *p = s + line * G::LS + k1 * G::P;
    float2 v[G::R2];
#pragma unroll
    for (int j = 0; j < G::R2; ++j) v[j] = p[j];
    reg_fft<G::R2, INV>(v);
#pragma unroll
    for (int j = 0; j < G::R2; ++j) p[j] = v[j];
  }
  __syncthreads();
}

// inverse transform, transposed layout in -> natural layout out (the mirror image of fft_lines<LOGR, true>).
template <int LOGR>
__device__ __forceinline__ void ifft_lines_mirror(float2 *s, const float2 *tw, int lines) {
  using G = Geo<LOGR>;
  for (int t = threadIdx.x; t < lines * G::R1; t += blockDim.x) {
    const int line = t >> G::L1, k1 = t & (G::R1 - 1);
    float2 *p = s + line * G::LS + k1 * G::P;
    float2 v[G::R2];
#pragma unroll
    for (int j = 0; j < G::R2; ++j) v[j] = p[j];
    reg_fft<G::R2, true>(v);
    p[0] = v[0];
#pragma unroll
    for (int n2 = 1; n2 < G::R2; ++n2) {
      float2 w = tw[n2 * k1];
      w.y = -w.y;
      p[n2] = cmul(v[n2], w);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < lines * G::R2; t += blockDim.x) {
    const int line = t >> G::L2, n2 = t & (G::R2 - 1);
    float2 *p = s + line * G::LS + n2;
    float2 v[G::R1];
#pragma unroll
    for (int j = 0; j < G::R1; ++j) v[j] = p[j * G::P];
    reg_fft<G::R1, true>(v);
#pragma unroll
    for (int j = 0; j < G::R1; ++j) p[j * G::P] = v[j];
  }
  __syncthreads();
}

// ---- mirror + label + permutation ----------------------------------------------------------------------------------
__global__ void centroid_kernel(const float *__restrict__ cloud, int N, float *__restrict__ centre) {
  pdl_wait();
  pdl_trigger();
  const float *src = cloud + (size_t)blockIdx.x * N * 6;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    sx += src[i * 6 + 0];
    sy += src[i * 6 + 1];
    sz += src[i * 6 + 2];
  }
  __shared__ float red[3][32];
  for (int o = 16; o; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = sx;
    red[1][threadIdx.x >> 5] = sy;
    red[2][threadIdx.x >> 5] = sz;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
    centre[blockIdx.x * 3 + threadIdx.x] = s / (float)N;
  }
}

__global__ void mirror_concat_kernel(const float *__restrict__ cloud, const float *__restrict__ centre,
                                     const int *__restrict__ perm, int B, int N, int axis, float *__restrict__ out,
                                     int ldo) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 2 * N) return;
  const int b = i / (2 * N), r = i - b * 2 * N;
  const int src = perm ? perm[r] : r;
  const bool mir = src >= N;
  const float *p = cloud + ((size_t)b * N + (mir ? src - N : src)) * 6;
  float v[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) v[c] = p[c];
  if (mir) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float ctr = centre[b * 3 + c];
      float rel = __fsub_rn(v[c], ctr);
      if (c == axis) rel = -rel;
      v[c] = __fadd_rn(rel, ctr);
    }
    v[3 + axis] = -v[3 + axis];
  }
  float *o = out + (size_t)i * ldo;
#pragma unroll
  for (int c = 0; c < 6; ++c) o[c] = v[c];
  o[6] = mir ? -1.f : 1.f;
}

// ---- bounding-box normalisation + unit-cube map: one CTA per sample -----------------------------------------------------
__global__ void unit_cube_kernel(const float *__restrict__ pts, int ld, int n, int explicit_normalize, float scale,
                                 float *__restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const float *src = pts + (size_t)blockIdx.x * n * ld;
  float *dst = out + (size_t)blockIdx.x * n * 3;
  __shared__ float red[6][32];
  __shared__ float box[4];  // centre xyz, extent
  if (explicit_normalize) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = src[(size_t)i * ld + c];
        lo[c] = fminf(lo[c], v);
        hi[c] = fmaxf(hi[c], v);
      }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      for (int o = 16; o; o >>= 1) {
        lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
        hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
      }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        red[c][threadIdx.x >> 5] = lo[c];
        red[3 + c][threadIdx.x >> 5] = hi[c];
      }
    __syncthreads();
    if (threadIdx.x == 0) {
      float ext = -INFINITY;
      for (int c = 0; c < 3; ++c) {
        float l = INFINITY, h = -INFINITY;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
          l = fminf(l, red[c][w]);
          h = fmaxf(h, red[3 + c][w]);
        }
        box[c] = __fdiv_rn(__fadd_rn(h, l), 2.f);
        ext = fmaxf(ext, __fsub_rn(h, l));
      }
      box[3] = ext;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n * 3; i += blockDim.x) {
    const int r = i / 3, c = i - r * 3;
    float v = src[(size_t)r * ld + c];
    if (explicit_normalize)
      v = __fmul_rn(__fdiv_rn(__fsub_rn(v, box[c]), box[3]), 0.99f);
    else
      v = __fdiv_rn(__fdiv_rn(v, scale), 2.f);
    v = __fadd_rn(__fdiv_rn(v, 1.2f), 0.5f);
    dst[i] = fminf(fmaxf(v, 0.f), 0.99f);
  }
}

// ---- trilinear corners (dpsr_utils/utils.py:155-172 / :87-111) -----------------------------------------------------------
struct Corners {
  int i0[3], i1[3];
  float w0[3], w1[3];  // weight of the low / high node along each axis
};

__device__ __forceinline__ Corners corners_of(const float *p, int R) {
  Corners c;
  const float size = (float)R;
  const float cell = __fdiv_rn(1.0f, size);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float q = __fdiv_rn(p[d], cell);
    const float f = floorf(q);
    c.i0[d] = (int)f;
    c.i1[d] = (int)fmodf(ceilf(q), size);
    // coordinates outside [0, 1) (the reference raises an index error there): wrap instead of touching memory outside the
    // grid; no effect on valid input
    if ((unsigned)c.i0[d] >= (unsigned)R) c.i0[d] = ((c.i0[d] % R) + R) % R;
    if ((unsigned)c.i1[d] >= (unsigned)R) c.i1[d] = ((c.i1[d] % R) + R) % R;
    const float lo = __fmul_rn(f, cell), hi = __fmul_rn(__fadd_rn(f, 1.0f), cell);
    c.w0[d] = __fdiv_rn(fabsf(__fsub_rn(p[d], hi)), cell);  // low node <- distance to the opposite (high) corner
    c.w1[d] = __fdiv_rn(fabsf(__fsub_rn(p[d], lo)), cell);
  }
  return c;
}

__global__ void splat_kernel(const float *__restrict__ V, int ldv, const float *__restrict__ Nr, int ldn, int B, int n,
                             int R, float *__restrict__ raster) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n) return;
  const int b = i / n;
  const float *p = V + (size_t)i * ldv;
  const float *nv = Nr + (size_t)i * ldn;
  const float pv[3] = {p[0], p[1], p[2]};
  const float val[3] = {nv[0], nv[1], nv[2]};
  const Corners c = corners_of(pv, R);
  const size_t vol = (size_t)R * R * R;
  float *base = raster + (size_t)b * 3 * vol;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int cx = k >> 2, cy = (k >> 1) & 1, cz = k & 1;
    const int x = cx ? c.i1[0] : c.i0[0], y = cy ? c.i1[1] : c.i0[1], z = cz ? c.i1[2] : c.i0[2];
    const float w = __fmul_rn(__fmul_rn(cx ? c.w1[0] : c.w0[0], cy ? c.w1[1] : c.w0[1]), cz ? c.w1[2] : c.w0[2]);
    const size_t cell = ((size_t)x * R + y) * R + z;
#pragma unroll
    for (int f = 0; f < 3; ++f) atomicAdd(base + f * vol + cell, __fmul_rn(w, val[f]));
  }
}

// All transform kernels are PERSISTENT and double-buffered: a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the
// raw bytes of tile i+1 are requested (cp.async) before tile i is repacked into the padded transform layout, transformed
// and stored, so every SM keeps 48-64 KB of reads in flight all the time (the one-tile-per-CTA version alternated load /
// compute / store phases and sat at 2.4 TB/s with the SMs 40-57 % busy: latency-bound).

// ---- Z pass: real lines -> half spectra, two lines per complex transform --------------------------------------------------
// raster f32 [lines, R]  ->  spec float2 [lines, Hp]
template <int LOGR, int LINES>
__global__ void __launch_bounds__(FFT_THREADS) fft_z_forward_kernel(const float *__restrict__ raster, float2 *__restrict__ spec,
                                                                    long long n_pairs, int Hp) {
  pdl_wait();
  pdl_trigger();
  using G = Geo<LOGR>;
  constexpr int R = G::R, H = R / 2 + 1;
  extern __shared__ __align__(16) float2 smem[];
  float2 *tw = smem;
  float2 *s = smem + R;
  float *raw0 = reinterpret_cast<float *>(s + LINES * G::LS + (((LINES * G::LS) & 1) ? 1 : 0));  // 16-byte aligned
  constexpr int RAW_STRIDE = LINES * 2 * R;
  fill_twiddles<LOGR>(tw);
  const long long n_tiles = (n_pairs + LINES - 1) / LINES;
  auto issue = [&](long long tile, float *dst) {
    const long long pair0 = tile * LINES;
    const int np = (int)min((long long)LINES, n_pairs - pair0);
    const float *src = raster + pair0 * 2 * R;
    for (int c = threadIdx.x; c < np * 2 * R / 4; c += blockDim.x) cp_async16(dst + 4 * c, src + 4 * c);
  };
  long long tile = blockIdx.x;
  int buf = 0;
  if (tile < n_tiles) issue(tile, raw0);
  cp_async_commit();
  for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    if (tile + gridDim.x < n_tiles) issue(tile + gridDim.x, raw0 + (buf ^ 1) * RAW_STRIDE);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const long long pair0 = tile * LINES;
    const int np = (int)min((long long)LINES, n_pairs - pair0);
    const float *rw = raw0 + buf * RAW_STRIDE;
#pragma unroll 4
    for (int t = threadIdx.x; t < np * R; t += blockDim.x) {
      const int p = t >> LOGR, i = t & (R - 1);
      s[p * G::LS + G::in_pos(i)] = make_float2(rw[p * 2 * R + i], rw[p * 2 * R + R + i]);
    }
    __syncthreads();
    fft_lines<LOGR, false>(s, tw, np);
    for (int t = threadIdx.x; t < np * 2 * H; t += blockDim.x) {
      const int p = t / (2 * H), r = t - p * 2 * H;
      const int which = r >= H, k = which ? r - H : r;
      const float2 zk = s[p * G::LS + G::out_pos(k)], zr = s[p * G::LS + G::out_pos((R - k) & (R - 1))];
      float2 o;
      if (!which)
        o = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y - zr.y));
      else
        o = make_float2(0.5f * (zk.y + zr.y), -0.5f * (zk.x - zr.x));
      spec[((pair0 + p) * 2 + which) * Hp + k] = o;
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

// ---- generic in-place complex pass along a strided axis ---------------------------------------------------------------------
// data float2; a line set o = (o_hi, o_lo), o_lo < n_lo: base = o_hi*stride_hi + o_lo*stride_lo; element i of the line at
// kz is base + i*es + kz.  One tile: KZT neighbouring kz of one line set (KZT * 8 contiguous bytes per row).
template <int LOGR, bool INV, int KZT>
__global__ void __launch_bounds__(FFT_THREADS) fft_axis_kernel(float2 *__restrict__ data, int n_sets, int n_lo, long long stride_hi,
                                                               long long stride_lo, long long es, int H, int Hp) {
  pdl_wait();
  pdl_trigger();
  using G = Geo<LOGR>;
  constexpr int R = G::R, CH = KZT / 2;  // 16-byte chunks per row
  extern __shared__ __align__(16) float2 smem[];
  float2 *tw = smem;
  float2 *s = smem + R;
  float2 *raw0 = s + KZT * G::LS + (((KZT * G::LS) & 1) ? 1 : 0);
  constexpr int RAW_STRIDE = R * KZT;
  fill_twiddles<LOGR>(tw);
  const int kchunks = (H + KZT - 1) / KZT;
  const long long n_tiles = (long long)n_sets * kchunks;
  auto base_of = [&](long long tile, int &kz0) -> float2 * {
    const int o = (int)(tile / kchunks);
    kz0 = (int)(tile % kchunks) * KZT;
    return data + (long long)(o / n_lo) * stride_hi + (long long)(o % n_lo) * stride_lo + kz0;
  };
  auto issue = [&](long long tile, float2 *dst) {
    int kz0;
    const float2 *base = base_of(tile, kz0);
    for (int c = threadIdx.x; c < R * CH; c += blockDim.x) {
      const int i = c / CH, j = c % CH;
      if (kz0 + 2 * j < Hp) cp_async16(dst + i * KZT + 2 * j, base + (long long)i * es + 2 * j);
    }
  };
  long long tile = blockIdx.x;
  int buf = 0;
  if (tile < n_tiles) issue(tile, raw0);
  cp_async_commit();
  for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    if (tile + gridDim.x < n_tiles) issue(tile + gridDim.x, raw0 + (buf ^ 1) * RAW_STRIDE);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    int kz0;
    float2 *base = base_of(tile, kz0);
    const int nk = min(KZT, H - kz0);
    const float2 *rw = raw0 + buf * RAW_STRIDE;
#pragma unroll 4
    for (int t = threadIdx.x; t < R * KZT; t += blockDim.x) {
      const int i = t / KZT, k = t % KZT;
      if (k < nk) s[k * G::LS + G::in_pos(i)] = rw[t];
    }
    __syncthreads();
    fft_lines<LOGR, INV>(s, tw, nk);
#pragma unroll 4
    for (int t = threadIdx.x; t < R * KZT; t += blockDim.x) {
      const int i = t / KZT, k = t % KZT;
      if (k < nk) base[(long long)i * es + k] = s[k * G::LS + G::out_pos(i)];
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

// spec_gaussian_filter (dpsr_utils/utils.py:65-71) as a table over |k|^2 = 0 .. 3 (R/2)^2: float32(exp(-0.5 (2 sig |k| / R)^2)),
// evaluated in float64 like the reference.  12 289 entries at R = 128; keeps the fp64 exp / sqrt out of the solve kernel.
__global__ void gauss_table_kernel(float *__restrict__ gtab, int n, float sig, int R) {
  pdl_wait();
  pdl_trigger();
  const int k2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (k2 >= n) return;
  const double dis = sqrt((double)k2);
  const double q = (double)sig * 2.0 * dis / (double)R;
  gtab[k2] = (float)exp(-0.5 * (q * q));
}

// ---- X pass + spectral solve + inverse X pass --------------------------------------------------------------------------------
// spec float2 [B,3,R(x),R(y),Hp] (Z and Y already transformed)  ->  pot float2 [B,R(x),R(y),Hp] (X already inverted)
//   Phi = sum_d (-i G N_d) w_d / (-(|w|^2) + 1e-6),  w = 2 pi k,  G = exp(-0.5 (2 sig |k| / R)^2),  Phi(0) = 0
template <int LOGR, int KZT>
__global__ void __launch_bounds__(FFT_THREADS) solve_x_kernel(const float2 *__restrict__ spec, float2 *__restrict__ pot, int B,
                                                              int H, int Hp, const float *__restrict__ gtab) {
  pdl_wait();
  pdl_trigger();
  using G = Geo<LOGR>;
  constexpr int R = G::R, CH = KZT / 2;
  extern __shared__ __align__(16) float2 smem[];
  float2 *tw = smem;
  float2 *s = smem + R;  // [3 channels][nk][LS]
  float2 *raw0 = s + 3 * KZT * G::LS + (((3 * KZT * G::LS) & 1) ? 1 : 0);
  constexpr int RAW_STRIDE = 3 * R * KZT;
  fill_twiddles<LOGR>(tw);
  const int kchunks = (H + KZT - 1) / KZT;
  const long long n_tiles = (long long)B * R * kchunks;
  const long long plane = (long long)R * Hp, vol = plane * R;
  auto issue = [&](long long tile, float2 *dst) {
    const int by = (int)(tile / kchunks), kz0 = (int)(tile % kchunks) * KZT;
    const int y = by & (R - 1), b = by >> LOGR;
    for (int c = threadIdx.x; c < 3 * R * CH; c += blockDim.x) {
      const int ch = c / (R * CH), rem = c - ch * (R * CH);
      const int i = rem / CH, j = rem % CH;
      if (kz0 + 2 * j < Hp)
        cp_async16(dst + (ch * R + i) * KZT + 2 * j,
                   spec + ((long long)b * 3 + ch) * vol + (long long)i * plane + (long long)y * Hp + kz0 + 2 * j);
    }
  };
  long long tile = blockIdx.x;
  int buf = 0;
  if (tile < n_tiles) issue(tile, raw0);
  cp_async_commit();
  for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    if (tile + gridDim.x < n_tiles) issue(tile + gridDim.x, raw0 + (buf ^ 1) * RAW_STRIDE);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int by = (int)(tile / kchunks), kz0 = (int)(tile % kchunks) * KZT;
    const int y = by & (R - 1), b = by >> LOGR;
    const int nk = min(KZT, H - kz0);
    const float2 *rw = raw0 + buf * RAW_STRIDE;
#pragma unroll 4
    for (int t = threadIdx.x; t < 3 * R * KZT; t += blockDim.x) {
      const int ch = t / (R * KZT), rem = t - ch * (R * KZT);
      const int i = rem / KZT, k = rem % KZT;
      if (k < nk) s[(ch * nk + k) * G::LS + G::in_pos(i)] = rw[t];
    }
    __syncthreads();
    fft_lines<LOGR, false>(s, tw, 3 * nk);
    // frequency i of line (c, k) now lies at out_pos(i); the potential replaces channel 0 in place (same position, same thread)
    const int fy = y < (R >> 1) ? y : y - R;
    const float TWO_PI = 6.2831855f;  // float32(2 pi): the reference scales a float32 frequency tensor in place
    const float wy = __fmul_rn((float)fy, TWO_PI);
    for (int t = threadIdx.x; t < nk * R; t += blockDim.x) {
      const int k = t >> LOGR, i = t & (R - 1);
      const int fx = i < (R >> 1) ? i : i - R;
      const int fz = kz0 + k;
      const float Gf = __ldg(gtab + fx * fx + fy * fy + fz * fz);  // the filter depends on |k|^2 only (gauss_table_kernel)
      const float wx = __fmul_rn((float)fx, TWO_PI), wz = __fmul_rn((float)fz, TWO_PI);
      const int pos = G::out_pos(i);
      const float2 nx = s[(0 * nk + k) * G::LS + pos], ny = s[(1 * nk + k) * G::LS + pos], nz = s[(2 * nk + k) * G::LS + pos];
      // -(i * (G z)) = (G im, -(G re)); summed over the axes in order, each term rounded as the reference's product
      float re = __fmul_rn(__fmul_rn(nx.y, Gf), wx);
      re = __fadd_rn(re, __fmul_rn(__fmul_rn(ny.y, Gf), wy));
      re = __fadd_rn(re, __fmul_rn(__fmul_rn(nz.y, Gf), wz));
      float im = __fmul_rn(-__fmul_rn(nx.x, Gf), wx);
      im = __fadd_rn(im, __fmul_rn(-__fmul_rn(ny.x, Gf), wy));
      im = __fadd_rn(im, __fmul_rn(-__fmul_rn(nz.x, Gf), wz));
      const float lap = -__fadd_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)), __fmul_rn(wz, wz));
      const float den = __fadd_rn(lap, 1e-6f);
      float2 o = make_float2(__fdiv_rn(re, den), __fdiv_rn(im, den));
      if (fx == 0 && fy == 0 && fz == 0) o = make_float2(0.f, 0.f);
      s[k * G::LS + pos] = o;
    }
    __syncthreads();
    ifft_lines_mirror<LOGR>(s, tw, nk);
    float2 *dst = pot + (long long)b * vol + (long long)y * Hp + kz0;
#pragma unroll 4
    for (int t = threadIdx.x; t < R * KZT; t += blockDim.x) {
      const int i = t / KZT, k = t % KZT;
      if (k < nk) dst[(long long)i * plane + k] = s[k * G::LS + G::in_pos(i)];
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

// ---- inverse Z pass: two Hermitian half spectra per complex transform -> two real lines ------------------------------------
template <int LOGR, int LINES>
__global__ void __launch_bounds__(FFT_THREADS) fft_z_inverse_kernel(const float2 *__restrict__ pot, float *__restrict__ phi,
                                                                    long long n_pairs, float norm, int Hp) {
  pdl_wait();
  pdl_trigger();
  using G = Geo<LOGR>;
  constexpr int R = G::R;
  extern __shared__ __align__(16) float2 smem[];
  float2 *tw = smem;
  float2 *s = smem + R;
  float2 *raw0 = s + LINES * G::LS + (((LINES * G::LS) & 1) ? 1 : 0);
  const int RAW_STRIDE = LINES * 2 * Hp;
  fill_twiddles<LOGR>(tw);
  const long long n_tiles = (n_pairs + LINES - 1) / LINES;
  auto issue = [&](long long tile, float2 *dst) {
    const long long pair0 = tile * LINES;
    const int np = (int)min((long long)LINES, n_pairs - pair0);
    const float2 *src = pot + pair0 * 2 * Hp;
    for (int c = threadIdx.x; c < np * Hp; c += blockDim.x) cp_async16(dst + 2 * c, src + 2 * c);  // Hp is even
  };
  long long tile = blockIdx.x;
  int buf = 0;
  if (tile < n_tiles) issue(tile, raw0);
  cp_async_commit();
  for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    if (tile + gridDim.x < n_tiles) issue(tile + gridDim.x, raw0 + (buf ^ 1) * RAW_STRIDE);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const long long pair0 = tile * LINES;
    const int np = (int)min((long long)LINES, n_pairs - pair0);
    const float2 *rw = raw0 + buf * RAW_STRIDE;
#pragma unroll 4
    for (int t = threadIdx.x; t < np * R; t += blockDim.x) {
      const int p = t >> LOGR, k = t & (R - 1);
      const float2 *A = rw + p * 2 * Hp;
      const float2 *Bv = A + Hp;
      float2 z;
      if (k <= (R >> 1)) {
        float2 a = A[k], c = Bv[k];
        if (k == 0 || k == (R >> 1)) a.y = 0.f, c.y = 0.f;  // a real-output transform ignores these imaginary parts
        z = make_float2(a.x - c.y, a.y + c.x);
      } else {
        const float2 a = A[R - k], c = Bv[R - k];
        z = make_float2(a.x + c.y, c.x - a.y);
      }
      s[p * G::LS + G::in_pos(k)] = z;
    }
    __syncthreads();
    fft_lines<LOGR, true>(s, tw, np);
    for (int t = threadIdx.x; t < np * 2 * R; t += blockDim.x) {
      const int p = t / (2 * R), r = t - p * 2 * R;
      const int which = r >= R, i = which ? r - R : r;
      const float2 z = s[p * G::LS + G::out_pos(i)];
      phi[((pair0 + p) * 2 + which) * R + i] = (which ? z.y : z.x) * norm;
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

// ---- read-back at the points, mean per sample -----------------------------------------------------------------------------------
__global__ void interp_sum_kernel(const float *__restrict__ phi, const float *__restrict__ V, int ldv, int n, int R,
                                  float *__restrict__ acc) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = 0.f;
  if (i < n) {
    const float *p = V + ((size_t)b * n + i) * ldv;
    const float pv[3] = {p[0], p[1], p[2]};
    const Corners c = corners_of(pv, R);
    const float *g = phi + (size_t)b * R * R * R;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int cx = k >> 2, cy = (k >> 1) & 1, cz = k & 1;
      const int x = cx ? c.i1[0] : c.i0[0], y = cy ? c.i1[1] : c.i0[1], z = cz ? c.i1[2] : c.i0[2];
      const float w = __fmul_rn(__fmul_rn(cx ? c.w1[0] : c.w0[0], cy ? c.w1[1] : c.w0[1]), cz ? c.w1[2] : c.w0[2]);
      v = __fadd_rn(v, __fmul_rn(g[((size_t)x * R + y) * R + z], w));
    }
  }
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    atomicAdd(acc + b, s);
  }
}

// scal[b] = (offset, |phi[b,0,0,0] - offset|)
__global__ void scalars_kernel(const float *__restrict__ phi, const float *__restrict__ acc, int B, int n, size_t vol,
                               int shift, float2 *__restrict__ scal) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float off = shift ? __fdiv_rn(acc[b], (float)n) : 0.f;
  scal[b] = make_float2(off, fabsf(__fsub_rn(phi[(size_t)b * vol], off)));
}

__global__ void finalize_kernel(float4 *__restrict__ phi, const float2 *__restrict__ scal, size_t vol4, int scale) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const float2 sc = scal[b];
  float4 *g = phi + (size_t)b * vol4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < vol4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = g[i];
    float *e = reinterpret_cast<float *>(&v);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float x = __fsub_rn(e[c], sc.x);
      if (scale) x = __fmul_rn(__fdiv_rn(-x, sc.y), 0.5f);
      e[c] = x;
    }
    g[i] = v;
  }
}

int log2_exact(int R) {
  int l = 0;
  while ((1 << l) < R) ++l;
  return (1 << l) == R ? l : -1;
}

// row pitch (in complex values) of the half spectra: R/2+1 rounded up to 64 bytes, so that every row and every 8-value
// chunk of a row is a whole number of 64-byte DRAM bursts (65 -> 72 at R = 128; the pad columns are never read or
// written).  With a 32-byte pitch (68) ncu showed 1.4-1.6x the algorithmic DRAM reads in the strided passes.
int half_pitch(int R) { return (R / 2 + 1 + 7) & ~7; }

struct DpsrLayout {
  size_t raster, spec, pot, acc, scal, gtab, total;
};

int gauss_table_size(int R) { return 3 * (R / 2) * (R / 2) + 1; }

DpsrLayout layout_of(int B, int R) {
  const size_t vol = (size_t)R * R * R, hvol = (size_t)R * R * half_pitch(R);
  DpsrLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  L.raster = take((size_t)B * 3 * vol * sizeof(float));
  L.spec = take((size_t)B * 3 * hvol * sizeof(float2));
  L.pot = take((size_t)B * hvol * sizeof(float2));
  L.acc = take((size_t)B * sizeof(float));
  L.scal = take((size_t)B * sizeof(float2));
  L.gtab = take((size_t)gauss_table_size(R) * sizeof(float));
  L.total = off;
  return L;
}


// Opt a kernel instantiation into > 48 KB of dynamic shared memory, once per device (the attribute is per function and
// device; re-setting it on every launch stalls the stream: 33 ms per batch instead of 3).
template <auto Kern>
int opt_in_smem(size_t bytes) {
  static bool done[64] = {false};
  if (bytes <= 48 * 1024) return SLIDE_OK;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (done[dev]) return SLIDE_OK;
  const int rc = cuda_rc(cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  if (rc == SLIDE_OK) done[dev] = true;
  return rc;
}

int sm_count_cached() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!sms[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

// persistent grid: as many CTAs as fit (shared memory / 8 per SM), never more than there are tiles
unsigned persistent_grid(long long n_tiles, size_t smem_bytes) {
  int per_sm = (int)((220 * 1024) / (smem_bytes + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  const long long g = (long long)sm_count_cached() * per_sm;
  return (unsigned)(n_tiles < g ? n_tiles : g);
}

// The five transform launches of one batch at R = 2^LOGR.
template <int LOGR>
int run_transforms_t(int B, float sig, float *raster, float2 *spec, float2 *pot, float *phi, float *gtab, cudaStream_t st) {
  using G = Geo<LOGR>;
  constexpr int R = G::R, H = R / 2 + 1;
  constexpr int KZA = R <= 128 ? 16 : 8;  // kz per tile of the axis passes (128 / 64 contiguous bytes per row)
  constexpr int KZS = R <= 128 ? 8 : 4;   // ... of the fused X pass (three channels resident)
  constexpr int LINES = (2048 / R) < 8 ? 8 : ((2048 / R) > 64 ? 64 : (2048 / R));  // line pairs per tile of the Z passes
  const int Hp = half_pitch(R);
  const size_t tw_bytes = (size_t)R * sizeof(float2), line_bytes = (size_t)G::LS * sizeof(float2);
  const long long plane = (long long)R * Hp, hvol = plane * R;
  int rc;
  {
    const long long n_pairs = (long long)B * 3 * R * R / 2;
    const size_t smem = tw_bytes + LINES * line_bytes + 16 + 2 * (size_t)LINES * 2 * R * sizeof(float);
    if ((rc = opt_in_smem<fft_z_forward_kernel<LOGR, LINES>>(smem))) return rc;
    launch_k(fft_z_forward_kernel<LOGR, LINES>, dim3(persistent_grid(ceil_div_ll(n_pairs, LINES), smem)), dim3(FFT_THREADS), smem,
             st, (const float *)raster, spec, n_pairs, Hp);
    if ((rc = after_launch())) return rc;
  }
  const int kca = ceil_div(H, KZA);
  const size_t smem_axis = tw_bytes + KZA * line_bytes + 16 + 2 * (size_t)R * KZA * sizeof(float2);
  // Y pass over the three normal channels: line set (b*3+c, x), element stride Hp
  if ((rc = opt_in_smem<fft_axis_kernel<LOGR, false, KZA>>(smem_axis))) return rc;
  launch_k(fft_axis_kernel<LOGR, false, KZA>, dim3(persistent_grid((long long)B * 3 * R * kca, smem_axis)), dim3(FFT_THREADS),
           smem_axis, st, spec, B * 3 * R, R, hvol, plane, (long long)Hp, H, Hp);
  if ((rc = after_launch())) return rc;
  {
    const size_t smem = tw_bytes + 3 * KZS * line_bytes + 16 + 2 * (size_t)3 * R * KZS * sizeof(float2);
    if ((rc = opt_in_smem<solve_x_kernel<LOGR, KZS>>(smem))) return rc;
    launch_k(gauss_table_kernel, dim3(ceil_div(gauss_table_size(R), 256)), dim3(256), 0, st, gtab, gauss_table_size(R), sig, R);
    if ((rc = after_launch())) return rc;
    launch_k(solve_x_kernel<LOGR, KZS>, dim3(persistent_grid((long long)B * R * ceil_div(H, KZS), smem)), dim3(FFT_THREADS), smem,
             st, (const float2 *)spec, pot, B, H, Hp, (const float *)gtab);
    if ((rc = after_launch())) return rc;
  }
  if ((rc = opt_in_smem<fft_axis_kernel<LOGR, true, KZA>>(smem_axis))) return rc;
  launch_k(fft_axis_kernel<LOGR, true, KZA>, dim3(persistent_grid((long long)B * R * kca, smem_axis)), dim3(FFT_THREADS),
           smem_axis, st, pot, B * R, R, hvol, plane, (long long)Hp, H, Hp);
  if ((rc = after_launch())) return rc;
  {
    const long long n_pairs = (long long)B * R * R / 2;
    const float norm = 1.0f / ((float)R * (float)R * (float)R);
    const size_t smem = tw_bytes + LINES * line_bytes + 16 + 2 * (size_t)LINES * 2 * Hp * sizeof(float2);
    if ((rc = opt_in_smem<fft_z_inverse_kernel<LOGR, LINES>>(smem))) return rc;
    launch_k(fft_z_inverse_kernel<LOGR, LINES>, dim3(persistent_grid(ceil_div_ll(n_pairs, LINES), smem)), dim3(FFT_THREADS), smem,
             st, (const float2 *)pot, phi, n_pairs, norm, Hp);
    if ((rc = after_launch())) return rc;
  }
  return SLIDE_OK;
}

int run_transforms(int logR, int B, float sig, float *raster, float2 *spec, float2 *pot, float *phi, float *gtab, cudaStream_t st) {
  switch (logR) {
    case 3: return run_transforms_t<3>(B, sig, raster, spec, pot, phi, gtab, st);
    case 4: return run_transforms_t<4>(B, sig, raster, spec, pot, phi, gtab, st);
    case 5: return run_transforms_t<5>(B, sig, raster, spec, pot, phi, gtab, st);
    case 6: return run_transforms_t<6>(B, sig, raster, spec, pot, phi, gtab, st);
    case 7: return run_transforms_t<7>(B, sig, raster, spec, pot, phi, gtab, st);
    case 8: return run_transforms_t<8>(B, sig, raster, spec, pot, phi, gtab, st);
  }
  return SLIDE_ERR_UNSUPPORTED;
}

}  // namespace

}  // namespace slide

using namespace slide;

extern "C" {

int slide_sap_mirror_concat(const float *cloud, int B, int N, int axis, const int *perm, float *centre_scratch, float *out,
                            int ldo, slide_stream_t stream) {
  if (!cloud || !out || !centre_scratch || B <= 0 || N <= 0 || axis < 0 || axis > 2 || ldo < 7) return SLIDE_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  launch_k(centroid_kernel, dim3(B), dim3(512), 0, st, cloud, N, centre_scratch);
  int rc = after_launch();
  if (rc) return rc;
  launch_k(mirror_concat_kernel, dim3(ceil_div(B * 2 * N, 256)), dim3(256), 0, st, cloud, (const float *)centre_scratch, perm,
           B, N, axis, out, ldo);
  return after_launch();
}

int slide_sap_unit_cube(const float *pts, int ld, int B, int n, int explicit_normalize, float dataset_scale, float *out,
                        slide_stream_t stream) {
  if (!pts || !out || B <= 0 || n <= 0 || ld < 3) return SLIDE_ERR_INVALID;
  if (!explicit_normalize && !(dataset_scale > 0.f)) return SLIDE_ERR_INVALID;
  launch_k(unit_cube_kernel, dim3(B), dim3(1024), 0, (cudaStream_t)stream, pts, ld, n, explicit_normalize, dataset_scale, out);
  return after_launch();
}

int slide_dpsr_workspace_bytes(int B, int res, size_t *bytes) {
  if (!bytes || B <= 0 || log2_exact(res) < 3 || res > 256) return SLIDE_ERR_INVALID;
  *bytes = layout_of(B, res).total;
  return SLIDE_OK;
}

int slide_dpsr_forward(const float *V, int ldv, const float *Nrm, int ldn, int B, int n, int res, float sig, int shift,
                       int scale, float *phi, void *workspace, size_t workspace_bytes, slide_stream_t stream) {
  const int R = res, logR = log2_exact(res);
  if (!V || !Nrm || !phi || !workspace || B <= 0 || n <= 0 || ldv < 3 || ldn < 3) return SLIDE_ERR_INVALID;
  if (logR < 3 || R > 256) return SLIDE_ERR_UNSUPPORTED;  // power-of-two grids 8..256 (the shipped configs use 128)
  const DpsrLayout L = layout_of(B, R);
  if (workspace_bytes < L.total) return SLIDE_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  char *ws = (char *)workspace;
  float *raster = (float *)(ws + L.raster);
  float2 *spec = (float2 *)(ws + L.spec);
  float2 *pot = (float2 *)(ws + L.pot);
  float *acc = (float *)(ws + L.acc);
  float2 *scal = (float2 *)(ws + L.scal);
  const size_t vol = (size_t)R * R * R;
  int rc;
  if ((rc = cuda_rc(cudaMemsetAsync(raster, 0, (size_t)B * 3 * vol * sizeof(float), st)))) return rc;
  if ((rc = cuda_rc(cudaMemsetAsync(acc, 0, (size_t)B * sizeof(float), st)))) return rc;
  launch_k(splat_kernel, dim3(ceil_div(B * n, 128)), dim3(128), 0, st, V, ldv, Nrm, ldn, B, n, R, raster);
  if ((rc = after_launch())) return rc;

  if ((rc = run_transforms(logR, B, sig, raster, spec, pot, phi, (float *)(ws + L.gtab), st))) return rc;
  if (!shift && !scale) return SLIDE_OK;
  if (shift) {
    launch_k(interp_sum_kernel, dim3(ceil_div(n, 256), B), dim3(256), 0, st, (const float *)phi, V, ldv, n, R, acc);
    if ((rc = after_launch())) return rc;
  }
  launch_k(scalars_kernel, dim3(ceil_div(B, 128)), dim3(128), 0, st, (const float *)phi, (const float *)acc, B, n, vol, shift, scal);
  if ((rc = after_launch())) return rc;
  launch_k(finalize_kernel, dim3(296, B), dim3(256), 0, st, (float4 *)phi, (const float2 *)scal, vol / 4, scale);
  return after_launch();
}

}  // extern "C"
