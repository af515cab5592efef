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

__device__ __forceinline__ int bitrev(int i, int logR) { return (int)(__brev((unsigned)i) >> (32 - logR)); }

// tw[k] = exp(-2 pi i k / R), k < R/2
__device__ __forceinline__ void fill_twiddles(float2 *tw, int R) {
  for (int k = threadIdx.x; k < (R >> 1); k += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)k / (float)R, &s, &c);
    tw[k] = make_float2(c, -s);
  }
}

// In-place radix-2 decimation-in-time FFT of `lines` lines of R complex values held bit-reversed in shared memory
// (line l at s + l * ldl).  Unnormalised; INV uses the conjugate twiddles.  Ends with a barrier.
template <bool INV>
__device__ __forceinline__ void fft_tile(float2 *s, const float2 *tw, int lines, int R, int logR, int ldl) {
  const int per_line = R >> 1;
  const int total = lines * per_line;
  for (int st = 0; st < logR; ++st) {
    const int half = 1 << st;
    __syncthreads();
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
      const int line = t >> (logR - 1);
      const int j = t & (per_line - 1);
      const int pos = j & (half - 1);
      const int i0 = ((j >> st) << (st + 1)) + pos;
      float2 w = tw[pos << (logR - 1 - st)];
      if (INV) w.y = -w.y;
      float2 *p = s + line * ldl + i0;
      const float2 a = p[0], b = p[half];
      const float tr = w.x * b.x - w.y * b.y;
      const float ti = w.x * b.y + w.y * b.x;
      p[0] = make_float2(a.x + tr, a.y + ti);
      p[half] = make_float2(a.x - tr, a.y - ti);
    }
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

// ---- Z pass: real lines -> half spectra, two lines per complex transform --------------------------------------------------
// raster f32 [lines, R]  ->  spec float2 [lines, H]
__global__ void fft_z_forward_kernel(const float *__restrict__ raster, float2 *__restrict__ spec, long long n_pairs, int R,
                                     int logR, int pairs_per_cta, int Hp) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float2 smem[];
  const int ldl = R + 1, H = (R >> 1) + 1;
  float2 *tw = smem;
  float2 *s = smem + (R >> 1);
  fill_twiddles(tw, R);
  const long long pair0 = (long long)blockIdx.x * pairs_per_cta;
  const int np = (int)min((long long)pairs_per_cta, n_pairs - pair0);
#pragma unroll 4
  for (int t = threadIdx.x; t < np * R; t += blockDim.x) {
    const int p = t >> logR, i = t & (R - 1);
    const float *a = raster + (pair0 + p) * 2 * R;
    s[p * ldl + bitrev(i, logR)] = make_float2(a[i], a[R + i]);
  }
  fft_tile<false>(s, tw, np, R, logR, ldl);
  for (int t = threadIdx.x; t < np * 2 * H; t += blockDim.x) {
    const int p = t / (2 * H), r = t - p * 2 * H;
    const int which = r >= H, k = which ? r - H : r;
    const float2 zk = s[p * ldl + k], zr = s[p * ldl + ((R - k) & (R - 1))];
    float2 o;
    if (!which)
      o = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y - zr.y));
    else
      o = make_float2(0.5f * (zk.y + zr.y), -0.5f * (zk.x - zr.x));
    spec[((pair0 + p) * 2 + which) * Hp + k] = o;
  }
}

// ---- generic in-place complex pass along a strided axis ---------------------------------------------------------------------
// data float2; a line set o = (o_hi, o_lo), o_lo < n_lo: base = o_hi*stride_hi + o_lo*stride_lo; element i of the line at
// kz is base + i*es + kz.  One CTA: KZT neighbouring kz of one line set.
template <bool INV>
__global__ void fft_axis_kernel(float2 *__restrict__ data, int n_lo, long long stride_hi, long long stride_lo, long long es,
                                int H, int R, int logR, int KZT) {  // H: valid kz per row (the row pitch is inside the strides)
  pdl_wait();
  pdl_trigger();
  extern __shared__ float2 smem[];
  const int ldl = R + 1;
  float2 *tw = smem;
  float2 *s = smem + (R >> 1);
  fill_twiddles(tw, R);
  const int o = blockIdx.x;
  const int kz0 = blockIdx.y * KZT;
  const int nk = min(KZT, H - kz0);
  float2 *base = data + (long long)(o / n_lo) * stride_hi + (long long)(o % n_lo) * stride_lo + kz0;
#pragma unroll 4
  for (int t = threadIdx.x; t < R * KZT; t += blockDim.x) {
    const int i = t / KZT, k = t - i * KZT;
    if (k < nk) s[k * ldl + bitrev(i, logR)] = base[(long long)i * es + k];
  }
  fft_tile<INV>(s, tw, nk, R, logR, ldl);
#pragma unroll 4
  for (int t = threadIdx.x; t < R * KZT; t += blockDim.x) {
    const int i = t / KZT, k = t - i * KZT;
    if (k < nk) base[(long long)i * es + k] = s[k * ldl + i];
  }
}

// ---- X pass + spectral solve + inverse X pass --------------------------------------------------------------------------------
// spec float2 [B,3,R(x),R(y),H] (Z and Y already transformed)  ->  pot float2 [B,R(x),R(y),H] (X already inverted)
//   Phi = sum_d (-i G N_d) w_d / (-(|w|^2) + 1e-6),  w = 2 pi k,  G = exp(-0.5 (2 sig |k| / R)^2),  Phi(0) = 0
__global__ void solve_x_kernel(const float2 *__restrict__ spec, float2 *__restrict__ pot, int R, int logR, int H, int Hp,
                               int KZT, float sig) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float2 smem[];
  const int ldl = R + 1;
  float2 *tw = smem;
  float2 *s = smem + (R >> 1);  // [3 channels + 1 potential][KZT][ldl]
  fill_twiddles(tw, R);
  const int y = blockIdx.x & (R - 1), b = blockIdx.x >> logR;
  const int kz0 = blockIdx.y * KZT;
  const int nk = min(KZT, H - kz0);
  const long long plane = (long long)R * Hp, vol = plane * R;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float2 *src = spec + ((long long)b * 3 + c) * vol + (long long)y * Hp + kz0;
#pragma unroll 4
    for (int t = threadIdx.x; t < R * KZT; t += blockDim.x) {
      const int i = t / KZT, k = t - i * KZT;
      if (k < nk) s[(c * nk + k) * ldl + bitrev(i, logR)] = src[(long long)i * plane + k];
    }
  }
  fft_tile<false>(s, tw, 3 * nk, R, logR, ldl);
  float2 *ph = s + 3 * nk * ldl;
  const int fy = y < (R >> 1) ? y : y - R;
  const float TWO_PI = 6.2831855f;  // float32(2 pi): the reference scales a float32 frequency tensor in place
  const float wy = __fmul_rn((float)fy, TWO_PI);
  for (int t = threadIdx.x; t < nk * R; t += blockDim.x) {
    const int k = t >> logR, i = t & (R - 1);
    const int fx = i < (R >> 1) ? i : i - R;
    const int fz = kz0 + k;
    const double dis = sqrt((double)(fx * fx + fy * fy + fz * fz));
    const double q = (double)sig * 2.0 * dis / (double)R;
    const float G = (float)exp(-0.5 * (q * q));
    const float wx = __fmul_rn((float)fx, TWO_PI), wz = __fmul_rn((float)fz, TWO_PI);
    const float2 nx = s[(0 * nk + k) * ldl + i], ny = s[(1 * nk + k) * ldl + i], nz = s[(2 * nk + k) * ldl + i];
    // -(i * (G z)) = (G im, -(G re)); summed over the axes in order, each term rounded as the reference's product
    float re = __fmul_rn(__fmul_rn(nx.y, G), wx);
    re = __fadd_rn(re, __fmul_rn(__fmul_rn(ny.y, G), wy));
    re = __fadd_rn(re, __fmul_rn(__fmul_rn(nz.y, G), wz));
    float im = __fmul_rn(-__fmul_rn(nx.x, G), wx);
    im = __fadd_rn(im, __fmul_rn(-__fmul_rn(ny.x, G), wy));
    im = __fadd_rn(im, __fmul_rn(-__fmul_rn(nz.x, G), wz));
    const float lap = -__fadd_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)), __fmul_rn(wz, wz));
    const float den = __fadd_rn(lap, 1e-6f);
    float2 o = make_float2(__fdiv_rn(re, den), __fdiv_rn(im, den));
    if (fx == 0 && fy == 0 && fz == 0) o = make_float2(0.f, 0.f);
    ph[k * ldl + bitrev(i, logR)] = o;
  }
  fft_tile<true>(ph, tw, nk, R, logR, ldl);
  float2 *dst = pot + (long long)b * vol + (long long)y * Hp + kz0;
#pragma unroll 4
  for (int t = threadIdx.x; t < R * KZT; t += blockDim.x) {
    const int i = t / KZT, k = t - i * KZT;
    if (k < nk) dst[(long long)i * plane + k] = ph[k * ldl + i];
  }
}

// ---- inverse Z pass: two Hermitian half spectra per complex transform -> two real lines ------------------------------------
__global__ void fft_z_inverse_kernel(const float2 *__restrict__ pot, float *__restrict__ phi, long long n_pairs, int R,
                                     int logR, int pairs_per_cta, float norm, int Hp) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float2 smem[];
  const int ldl = R + 1;
  float2 *tw = smem;
  float2 *s = smem + (R >> 1);
  fill_twiddles(tw, R);
  const long long pair0 = (long long)blockIdx.x * pairs_per_cta;
  const int np = (int)min((long long)pairs_per_cta, n_pairs - pair0);
#pragma unroll 4
  for (int t = threadIdx.x; t < np * R; t += blockDim.x) {
    const int p = t >> logR, k = t & (R - 1);
    const float2 *A = pot + (pair0 + p) * 2 * Hp;
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
    s[p * ldl + bitrev(k, logR)] = z;
  }
  fft_tile<true>(s, tw, np, R, logR, ldl);
  for (int t = threadIdx.x; t < np * 2 * R; t += blockDim.x) {
    const int p = t / (2 * R), r = t - p * 2 * R;
    const int which = r >= R, i = which ? r - R : r;
    const float2 z = s[p * ldl + i];
    phi[((pair0 + p) * 2 + which) * R + i] = (which ? z.y : z.x) * norm;
  }
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

// row pitch (in complex values) of the half spectra: R/2+1 rounded up to 32 bytes, so that every row and every 8-value
// chunk of a row starts on a sector boundary (65 -> 68 at R = 128; the pad columns are never read or written)
int half_pitch(int R) { return (R / 2 + 1 + 3) & ~3; }

struct DpsrLayout {
  size_t raster, spec, pot, acc, scal, total;
};

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
  L.total = off;
  return L;
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
  const int H = R / 2 + 1, Hp = half_pitch(R);
  const size_t vol = (size_t)R * R * R;
  int rc;
  if ((rc = cuda_rc(cudaMemsetAsync(raster, 0, (size_t)B * 3 * vol * sizeof(float), st)))) return rc;
  if ((rc = cuda_rc(cudaMemsetAsync(acc, 0, (size_t)B * sizeof(float), st)))) return rc;
  launch_k(splat_kernel, dim3(ceil_div(B * n, 128)), dim3(128), 0, st, V, ldv, Nrm, ldn, B, n, R, raster);
  if ((rc = after_launch())) return rc;

  const int KZT = R <= 128 ? 8 : 4;
  const int PAIRS = max(1, 1024 / R);  // 8 pairs (16 lines) per CTA at R = 128
  const size_t tw_bytes = (size_t)(R / 2) * sizeof(float2), line_bytes = (size_t)(R + 1) * sizeof(float2);
  {
    const long long n_pairs = (long long)B * 3 * R * R / 2;
    launch_k(fft_z_forward_kernel, dim3((unsigned)ceil_div_ll(n_pairs, PAIRS)), dim3(FFT_THREADS), tw_bytes + PAIRS * line_bytes,
             st, (const float *)raster, spec, n_pairs, R, logR, PAIRS, Hp);
    if ((rc = after_launch())) return rc;
  }
  const long long plane = (long long)R * Hp, hvol = plane * R;
  // Y pass over the three normal channels: line set (b*3+c, x), element stride H
  launch_k(fft_axis_kernel<false>, dim3(B * 3 * R, ceil_div(H, KZT)), dim3(FFT_THREADS), tw_bytes + KZT * line_bytes, st, spec, R,
           hvol, plane, (long long)Hp, H, R, logR, KZT);
  if ((rc = after_launch())) return rc;
  launch_k(solve_x_kernel, dim3(R * B, ceil_div(H, KZT)), dim3(FFT_THREADS), tw_bytes + 4 * KZT * line_bytes, st,
           (const float2 *)spec, pot, R, logR, H, Hp, KZT, sig);
  if ((rc = after_launch())) return rc;
  launch_k(fft_axis_kernel<true>, dim3(B * R, ceil_div(H, KZT)), dim3(FFT_THREADS), tw_bytes + KZT * line_bytes, st, pot, R, hvol,
           plane, (long long)Hp, H, R, logR, KZT);
  if ((rc = after_launch())) return rc;
  {
    const long long n_pairs = (long long)B * R * R / 2;
    const float norm = 1.0f / ((float)R * (float)R * (float)R);
    launch_k(fft_z_inverse_kernel, dim3((unsigned)ceil_div_ll(n_pairs, PAIRS)), dim3(FFT_THREADS), tw_bytes + PAIRS * line_bytes,
             st, (const float2 *)pot, phi, n_pairs, R, logR, PAIRS, norm, Hp);
    if ((rc = after_launch())) return rc;
  }
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
