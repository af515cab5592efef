// Sample-resident executor (include/slide_resident.h): one kernel launch per diffusion step for the denoisers over
// 16 latent points.  One thread-block cluster owns one sample; its activations stay in shared memory from the first
// layer to the update of x, GroupNorm statistics are reduced in shared memory and completed across the cluster through
// distributed shared memory, weights stream from L2 through a 3-stage cp.async ring (with the next layer's first chunks
// already in flight while the current layer's epilogue runs).
//
// Arithmetic: the contractions here are small (K, N <= 139 for the position DDPM) and every operand is on chip, so the
// work is issue/latency bound, not tensor-pipe bound; the dense products run on the warp-level tensor path (mma.sync
// m16n8k8, TF32 operands rounded to nearest, fp32 accumulate -- the numerics class of the reference's cuDNN convs) fed by
// ldmatrix from padded shared-memory rows (row stride = 4 mod 8 floats: conflict-free).  PRECISE plans split both
// operands (3xTF32) and match the fp32 oracle to ~1e-6.  Everything else (grouping, GroupNorm, soft-max, update) is
// fp32 on the CUDA cores.  Replaces 65 launches per position-DDPM step of the per-record executor (program.cu).
#include <math.h>
#include <string.h>

#include "../../include/slide_resident.h"
#include "common.cuh"
#include "program.cuh"

namespace slide {

namespace {

constexpr int RT = SLIDE_RES_THREADS;        // 512 threads = 16 warps
constexpr int NW = RT / 32;
constexpr int WCH = SLIDE_RES_WCHUNK;        // K columns per staged weight chunk
constexpr int WLD = WCH + SLIDE_RES_WPAD;    // staged row stride in floats (4 mod 8)
constexpr int WSTAGES = SLIDE_RES_WSTAGES;
constexpr int NTMAX = 5;                     // n-tiles per warp

struct ResArgs {
  const slide_rop *rops;
  int n_rops;
  char *arena;
  const char *weights;
  char *scratch;  // [grid][scratch_bytes]
  int *done;      // CTAs finished (for the step-counter update)
  slide_resident_plan plan;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_peer(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_peer(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ uint32_t tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- operands ------------------------------------------------------------------------------------------------------
struct Ctx {
  float *sm;            // dynamic shared memory
  char *arena;
  const char *weights;
  int sample, rank, cl, np, npl, p0, t;
};

struct Opnd {
  float *ptr;  // resolved base pointer (shared or global), nullptr = absent
  int ld;
  bool shared;
};

__device__ __forceinline__ Opnd resolve(const Ctx &c, const int64_t *f) {
  Opnd o;
  o.ld = (int)f[RO_LD];
  o.shared = false;
  switch ((int)f[RO_SPACE]) {
    case 1:
      o.ptr = c.sm + f[RO_OFF];
      o.shared = true;
      break;
    case 2:
      o.ptr = reinterpret_cast<float *>(c.arena + f[RO_OFF] + (int64_t)c.sample * f[RO_SSTRIDE]);
      break;
    case 3:
      o.ptr = reinterpret_cast<float *>(c.arena + f[RO_OFF]);
      break;
    case 4:
      o.ptr = reinterpret_cast<float *>(const_cast<char *>(c.weights) + f[RO_OFF]);
      break;
    default:
      o.ptr = nullptr;
  }
  return o;
}
__device__ __forceinline__ const float *wptr(const Ctx &c, int64_t off) {
  return off < 0 ? nullptr : reinterpret_cast<const float *>(c.weights + off);
}

// ---- weight-chunk ring ---------------------------------------------------------------------------------------------
// Chunk q of a GEMM's packed weight copy ([nchunk][npad][WLD] floats) -> stage (issue index % WSTAGES).  Every thread
// commits exactly one group per issued chunk, so cp.async.wait_group counts are uniform across the CTA.
__device__ __forceinline__ void issue_chunk(const Ctx &c, float *wst, int stage_floats, int slot, int64_t wch, int npad, int q,
                                            int tid) {
  const float4 *src = reinterpret_cast<const float4 *>(c.weights + wch) + (size_t)q * npad * (WLD / 4);
  const uint32_t dst = smem_u32(wst + (size_t)slot * stage_floats);
  const int n16 = npad * (WLD / 4);
  for (int i = tid; i < n16; i += RT) cp_async16(dst + i * 16, src + i);
  cp_commit();
}

// ---- RS_GEMM -------------------------------------------------------------------------------------------------------
// Warp grid over (row tiles of 16, column tiles of 8): 8 row tiles -> 4 x 4 warps with 2 row tiles each, 4 -> 4 x 4,
// 2 -> 2 x 8, 1 -> 1 x 16.  More than 128 rows: passes of 128 rows (the weight chunks are streamed again).
template <bool PRECISE>
__device__ void rs_gemm(const Ctx &c, const slide_rop &r, float *wst, int stage_floats, int &ring, int tid) {
  const int64_t *f = r.i;
  const Opnd A = resolve(c, f + RG_A), C = resolve(c, f + RG_C), EV = resolve(c, f + RG_EV), RES = resolve(c, f + RG_RES);
  const int M = (int)f[RG_M], K = (int)f[RG_K], N = (int)f[RG_N];
  const int pairrows = (int)f[RG_PAIRROWS], rpp = (int)f[RG_RPP];
  const int nchunk = (int)f[RG_NCHUNK], npad = (int)f[RG_NPAD];
  const int64_t wch = f[RG_WCH];
  const float *bias = wptr(c, f[RG_BIAS]);
  const int act = (int)f[RG_ACT], smk = (int)f[RG_SMK];
  const int st_off = (int)f[RG_ST], st_cg = (int)f[RG_ST_CG], st_nnorm = (int)f[RG_ST_NNORM], st_choff = (int)f[RG_ST_CHOFF];
  const int st_owned = (int)f[RG_ST_OWNED];
  const float st_w = r.f[0];
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int n_tiles = (N + 7) >> 3;

  for (int m_pass = 0; m_pass < M; m_pass += 128) {
    const int m_rows = min(128, M - m_pass);
    const int m_tiles = m_rows >> 4;
    int wgm, mt_n;
    if (m_tiles >= 8) { wgm = 4; mt_n = 2; }
    else if (m_tiles >= 4) { wgm = 4; mt_n = 1; }
    else if (m_tiles >= 2) { wgm = 2; mt_n = 1; }
    else { wgm = 1; mt_n = 1; }
    const int wgn = NW / wgm;
    const int wm = warp % wgm, wn = warp / wgm;
    const int nt_n = (n_tiles + wgn - 1) / wgn;  // <= NTMAX by construction (N <= 128 -> 16 tiles / >= 4 column warps)
    const int tile0 = wn * nt_n;                  // this warp's first column tile
    const int row0 = m_pass + wm * mt_n * 16;     // this warp's first row

    float acc[2][NTMAX][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < NTMAX; ++b)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;

    // chunks already in flight for this GEMM (issued by the previous GEMM rop's tail): only on the first row pass
    int issued = (m_pass == 0) ? min(ring, nchunk) : 0;
    if (m_pass == 0) ring = 0;
    // ldmatrix lane addressing
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = ((lane >> 4) & 1) * 4;
    const int b_row = (lane & 7) + ((lane >> 4) & 1) * 8, b_col = ((lane >> 3) & 1) * 4;
    for (int q = 0; q < nchunk; ++q) {
      // keep two chunks in flight
      while (issued < nchunk && issued < q + 2) {
        issue_chunk(c, wst, stage_floats, issued % WSTAGES, wch, npad, issued, tid);
        ++issued;
      }
      if (issued - q >= 2) cp_wait<1>(); else cp_wait<0>();
      __syncthreads();  // chunk q visible to all warps; everyone is done with chunk q-1 (its stage may be refilled)
      const float *ws = wst + (size_t)(q % WSTAGES) * stage_floats;
#pragma unroll
      for (int ks = 0; ks < WCH / 8; ++ks) {
        const int k0 = q * WCH + ks * 8;
        if (k0 >= K) break;
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (mt < mt_n) {
            uint32_t r0, r1, r2, r3;
            ldsm4(smem_u32(A.ptr + (size_t)(row0 + mt * 16 + a_row) * A.ld + k0 + a_col), r0, r1, r2, r3);
            float v[4] = {__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3)};
            if (k0 + 8 > K) {  // K tail: columns >= K of A are not part of the operand (and may hold anything)
              if (k0 + tq >= K) v[0] = v[1] = 0.f;
              if (k0 + 4 + tq >= K) v[2] = v[3] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              ah[mt][e] = tf32_rna(v[e]);
              if (PRECISE) al[mt][e] = tf32_rna(v[e] - __uint_as_float(ah[mt][e]));
            }
          }
        }
#pragma unroll
        for (int jp = 0; jp < (NTMAX + 1) / 2; ++jp) {
          const int j0 = jp * 2;
          if (j0 < nt_n && tile0 + j0 < n_tiles) {
            uint32_t b[4];
            ldsm4(smem_u32(ws + (size_t)((tile0 + j0) * 8 + b_row) * WLD + ks * 8 + b_col), b[0], b[1], b[2], b[3]);
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = j0 + jj;
              if (j < NTMAX && j < nt_n && tile0 + j < n_tiles) {
                uint32_t bh0 = b[jj * 2], bh1 = b[jj * 2 + 1];
                if (PRECISE) {
                  const float f0 = __uint_as_float(bh0), f1 = __uint_as_float(bh1);
                  bh0 = tf32_rna(f0);
                  bh1 = tf32_rna(f1);
                  const uint32_t bl0 = tf32_rna(f0 - __uint_as_float(bh0)), bl1 = tf32_rna(f1 - __uint_as_float(bh1));
#pragma unroll
                  for (int mt = 0; mt < 2; ++mt)
                    if (mt < mt_n) {
                      mma_tf32(acc[mt][j], al[mt][0], al[mt][1], al[mt][2], al[mt][3], bh0, bh1);
                      mma_tf32(acc[mt][j], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bl0, bl1);
                    }
                }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
                  if (mt < mt_n) mma_tf32(acc[mt][j], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bh0, bh1);
              }
            }
          }
        }
      }
    }
    __syncthreads();  // every warp is done with the last chunk: the ring may be refilled
    // the next GEMM's first two chunks go in flight under this epilogue (and any rops in between)
    if (m_pass + 128 >= M && f[RG_NEXT_WCH] >= 0) {
      const int nn = (int)f[RG_NEXT_NPAD];
      const int nc = (int)f[RG_NEXT_NCHUNK];
      ring = min(2, nc);
      for (int q = 0; q < ring; ++q) issue_chunk(c, wst, stage_floats, q % WSTAGES, f[RG_NEXT_WCH], nn, q, tid);
    }

    // ---- epilogue ----
    float *st = st_off >= 0 ? c.sm + st_off : nullptr;
#pragma unroll
    for (int j = 0; j < NTMAX; ++j) {
      const int tile = tile0 + j;
      if (j >= nt_n || tile >= n_tiles) continue;
      const int n0 = tile * 8 + tq * 2;
      const bool v0 = n0 < N, v1 = n0 + 1 < N;
      const float b0 = (bias && v0) ? __ldg(bias + n0) : 0.f, b1 = (bias && v1) ? __ldg(bias + n0 + 1) : 0.f;
      if (smk > 0) {
        // fused soft-max over the smk (8 or 16) rows of each point, applied to the value rows RES
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (mt >= mt_n) continue;
          const int rbase = row0 + mt * 16;
          float s[4] = {acc[mt][j][0] + b0, acc[mt][j][1] + b1, acc[mt][j][2] + b0, acc[mt][j][3] + b1};
          float val[4];
          val[0] = v0 ? RES.ptr[(size_t)(rbase + g) * RES.ld + n0] : 0.f;
          val[1] = v1 ? RES.ptr[(size_t)(rbase + g) * RES.ld + n0 + 1] : 0.f;
          val[2] = v0 ? RES.ptr[(size_t)(rbase + g + 8) * RES.ld + n0] : 0.f;
          val[3] = v1 ? RES.ptr[(size_t)(rbase + g + 8) * RES.ld + n0 + 1] : 0.f;
          float mx[4] = {s[0], s[1], s[2], s[3]};
          if (smk == 16) {
            mx[0] = mx[2] = fmaxf(s[0], s[2]);
            mx[1] = mx[3] = fmaxf(s[1], s[3]);
          }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1)
#pragma unroll
            for (int e = 0; e < 4; ++e) mx[e] = fmaxf(mx[e], __shfl_xor_sync(0xffffffffu, mx[e], o));
          float ex[4], num[4], den[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ex[e] = expf(s[e] - mx[e]);
            num[e] = ex[e] * val[e];
            den[e] = ex[e];
          }
          if (smk == 16) {
            num[0] += num[2]; num[1] += num[3]; den[0] += den[2]; den[1] += den[3];
          }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              num[e] += __shfl_xor_sync(0xffffffffu, num[e], o);
              den[e] += __shfl_xor_sync(0xffffffffu, den[e], o);
            }
          if (g == 0) {
            const int npts = smk == 16 ? 1 : 2;
            for (int h = 0; h < npts; ++h) {
              const int prow = c.p0 + (rbase + h * 8) / smk;
              float *dst = C.ptr + (size_t)prow * C.ld + n0;
              const float o0 = num[h * 2] / den[h * 2], o1 = num[h * 2 + 1] / den[h * 2 + 1];
              if (v0) dst[0] = o0;
              if (v1) dst[1] = o1;
              if (c.cl > 1) {
                const uint32_t la = smem_u32(dst);
                for (int pr = 0; pr < c.cl; ++pr) {
                  if (pr == c.rank) continue;
                  const uint32_t ra = map_peer(la, pr);
                  if (v0) st_peer(ra, o0);
                  if (v1) st_peer(ra + 4, o1);
                }
              }
            }
          }
        }
        continue;
      }
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (mt >= mt_n) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = row0 + mt * 16 + g + h * 8;
          float x0 = acc[mt][j][h * 2] + b0, x1 = acc[mt][j][h * 2 + 1] + b1;
          const int point = pairrows ? c.p0 + row / rpp : row;
          if (EV.ptr) {
            if (v0) x0 += EV.ptr[(size_t)point * EV.ld + n0];
            if (v1) x1 += EV.ptr[(size_t)point * EV.ld + n0 + 1];
          }
          if (RES.ptr) {
            if (v0) x0 += RES.ptr[(size_t)row * RES.ld + n0];
            if (v1) x1 += RES.ptr[(size_t)row * RES.ld + n0 + 1];
          }
          if (act == 1) {
            x0 = fmaxf(x0, 0.f);
            x1 = fmaxf(x1, 0.f);
          }
          if (v0) C.ptr[(size_t)row * C.ld + n0] = x0;
          if (v1) C.ptr[(size_t)row * C.ld + n0 + 1] = x1;
          const bool counted = !st_owned || (point >= c.p0 && point < c.p0 + c.npl);
          if (st && counted) {
            if (v0) { s0 += x0; q0 += x0 * x0; }
            if (v1) { s1 += x1; q1 += x1 * x1; }
          }
        }
      }
      if (st) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          q0 += __shfl_xor_sync(0xffffffffu, q0, o);
          q1 += __shfl_xor_sync(0xffffffffu, q1, o);
        }
        if (g == 0) {
          const int ch0 = st_choff + n0;
          if (v0 && ch0 < st_nnorm) {
            atomicAdd(st + 2 * (ch0 / st_cg), st_w * s0);
            atomicAdd(st + 2 * (ch0 / st_cg) + 1, st_w * q0);
          }
          if (v1 && ch0 + 1 < st_nnorm) {
            atomicAdd(st + 2 * ((ch0 + 1) / st_cg), st_w * s1);
            atomicAdd(st + 2 * ((ch0 + 1) / st_cg) + 1, st_w * q1);
          }
        }
      }
    }
  }
  __syncthreads();
}

// ---- RS_PAIR -------------------------------------------------------------------------------------------------------
__device__ void rs_pair(const Ctx &c, const slide_rop &r, int tid) {
  const int64_t *f = r.i;
  const Opnd U = resolve(c, f + RP_U), X = resolve(c, f + RP_XYZ), CT = resolve(c, f + RP_CTR), O = resolve(c, f + RP_OUT),
             RES = resolve(c, f + RP_RES);
  const int K = (int)f[RP_K], N = (int)f[RP_N], act = (int)f[RP_ACT];
  const int *idx = reinterpret_cast<const int *>(c.sm + f[RP_IDX]);
  const float *d2 = f[RP_D2] >= 0 ? c.sm + f[RP_D2] : nullptr;
  const float *wx = wptr(c, f[RP_WX]), *wc = wptr(c, f[RP_WC]), *wd = wptr(c, f[RP_WD]), *ww = wptr(c, f[RP_WW]),
              *bias = wptr(c, f[RP_BIAS]);
  const int st_off = (int)f[RP_ST], st_cg = (int)f[RP_ST_CG], st_nnorm = (int)f[RP_ST_NNORM], st_choff = (int)f[RP_ST_CHOFF];
  float *st = st_off >= 0 ? c.sm + st_off : nullptr;
  const int warp = tid >> 5, lane = tid & 31;
  const int rows = c.npl * K;
  for (int n0 = 0; n0 < N; n0 += 32) {
    const int n = n0 + lane;
    const bool v = n < N;
    float wxx = 0.f, wxy = 0.f, wxz = 0.f, wcx = 0.f, wcy = 0.f, wcz = 0.f, bb = 0.f, wdd = 0.f, www = 0.f;
    if (v) {
      wxx = __ldg(wx + n * 3); wxy = __ldg(wx + n * 3 + 1); wxz = __ldg(wx + n * 3 + 2);
      wcx = __ldg(wc + n * 3); wcy = __ldg(wc + n * 3 + 1); wcz = __ldg(wc + n * 3 + 2);
      if (bias) bb = __ldg(bias + n);
      if (d2) { wdd = __ldg(wd + n); www = __ldg(ww + n); }
    }
    float s = 0.f, q = 0.f;
    for (int lr = warp; lr < rows; lr += NW) {
      const int il = lr / K, k = lr - il * K;
      const int pi = c.p0 + il;
      const int j = idx[pi * K + k];
      const float xj0 = X.ptr[(size_t)j * X.ld], xj1 = X.ptr[(size_t)j * X.ld + 1], xj2 = X.ptr[(size_t)j * X.ld + 2];
      const float c0 = CT.ptr[(size_t)pi * CT.ld], c1 = CT.ptr[(size_t)pi * CT.ld + 1], c2 = CT.ptr[(size_t)pi * CT.ld + 2];
      float val = 0.f;
      if (v) {
        val = U.ptr[(size_t)j * U.ld + n];
        val += xj0 * wxx + xj1 * wxy + xj2 * wxz;
        val += c0 * wcx + c1 * wcy + c2 * wcz;
        val += bb;
      }
      if (d2) {
        float sum = 0.f;
        for (int kk = 0; kk < K; ++kk) sum += 1.0f / (d2[pi * K + kk] + 1e-8f);
        const float dd = d2[pi * K + k];
        const float w = (1.0f / (dd + 1e-8f)) / sum;
        val += dd * wdd + w * www;
      }
      if (v) {
        if (RES.ptr) val += RES.ptr[(size_t)lr * RES.ld + n];
        if (act == 1) val = fmaxf(val, 0.f);
        O.ptr[(size_t)lr * O.ld + n] = val;
        s += val;
        q += val * val;
      }
    }
    if (st) {
      const int ch = st_choff + n;
      const bool in = v && ch < st_nnorm;
      if (!in) s = q = 0.f;
      // lanes of one group are adjacent when the group size is a power of two that divides the chunk start
      if ((st_cg & (st_cg - 1)) == 0 && st_cg <= 32 && ((st_choff + n0) % st_cg) == 0) {
        for (int o = 1; o < st_cg; o <<= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o);
          q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (in && (lane % st_cg) == 0) {
          atomicAdd(st + 2 * (ch / st_cg), s);
          atomicAdd(st + 2 * (ch / st_cg) + 1, q);
        }
      } else if (in) {
        atomicAdd(st + 2 * (ch / st_cg), s);
        atomicAdd(st + 2 * (ch / st_cg) + 1, q);
      }
    }
  }
  __syncthreads();
}

// ---- RS_XFORM ------------------------------------------------------------------------------------------------------
__device__ void rs_xform(const Ctx &c, const slide_rop &r, int tid) {
  const int64_t *f = r.i;
  const Opnd X = resolve(c, f + RX_X);
  const int rows = (int)f[RX_ROWS], C = (int)f[RX_C];
  const int st_off = (int)f[RX_ST], cg = (int)f[RX_CG], nnorm = (int)f[RX_NNORM], choff = (int)f[RX_CHOFF];
  const float *gamma = wptr(c, f[RX_GAMMA]), *beta = wptr(c, f[RX_BETA]);
  const int relu = (int)f[RX_RELU], addmode = (int)f[RX_ADDMODE];
  const float inv_count = r.f[0];
  const float *st = st_off >= 0 ? c.sm + st_off : nullptr;
  const float *add = nullptr;
  if ((int)f[RX_ADD + RO_SPACE] != 0) {
    const float *base = reinterpret_cast<const float *>(c.arena + f[RX_ADD + RO_OFF]);
    const int64_t arow = addmode == 0 ? c.sample : (addmode == 1 ? c.t : 0);
    add = base + arow * f[RX_ADD + RO_LD];
  }
  const int warp = tid >> 5, lane = tid & 31;
  for (int n0 = 0; n0 < C; n0 += 32) {
    const int n = n0 + lane;
    if (n >= C) continue;
    float a = 1.f, b = 0.f;
    const int ch = choff + n;
    if (st && ch < nnorm) {
      const float sum = st[2 * (ch / cg)], sq = st[2 * (ch / cg) + 1];
      const float mean = sum * inv_count;
      float var = sq * inv_count - mean * mean;
      var = var < 0.f ? 0.f : var;
      const float rstd = 1.0f / sqrtf(var + SLIDE_GN_EPS);
      a = rstd * __ldg(gamma + ch);
      b = __ldg(beta + ch) - mean * a;
    }
    const float av = add ? __ldg(add + n) : 0.f;
    for (int row = warp; row < rows; row += NW) {
      float v = X.ptr[(size_t)row * X.ld + n];
      v = v * a + b;
      if (relu) v = fmaxf(v, 0.f);
      X.ptr[(size_t)row * X.ld + n] = v + av;
    }
  }
  __syncthreads();
}

// ---- RS_KNN --------------------------------------------------------------------------------------------------------
__device__ void rs_knn(const Ctx &c, const slide_rop &r, int tid) {
  const int64_t *f = r.i;
  const Opnd Q = resolve(c, f + RK_Q), R = resolve(c, f + RK_REF);
  const int P1 = (int)f[RK_P1], P2 = (int)f[RK_P2], K = (int)f[RK_K];
  int *idx = reinterpret_cast<int *>(c.sm + f[RK_IDX]);
  float *d2 = f[RK_D2] >= 0 ? c.sm + f[RK_D2] : nullptr;
  if (tid < P1) {
    const float qx = Q.ptr[(size_t)tid * Q.ld], qy = Q.ptr[(size_t)tid * Q.ld + 1], qz = Q.ptr[(size_t)tid * Q.ld + 2];
    float bd[16];
    int bi[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      bd[k] = INFINITY;
      bi[k] = 0;
    }
    for (int p = 0; p < P2; ++p) {
      const float d = sumsq3_p3d(qx - R.ptr[(size_t)p * R.ld], qy - R.ptr[(size_t)p * R.ld + 1], qz - R.ptr[(size_t)p * R.ld + 2]);
      float cd = d;
      int ci = p;
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        const bool sw = cd < bd[s];
        const float td = bd[s];
        const int ti = bi[s];
        bd[s] = sw ? cd : td;
        bi[s] = sw ? ci : ti;
        cd = sw ? td : cd;
        ci = sw ? ti : ci;
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K) {
        idx[tid * K + k] = bi[k];
        if (d2) d2[tid * K + k] = bd[k];
      }
  }
  __syncthreads();
}

// ---- RS_COPY -------------------------------------------------------------------------------------------------------
__device__ void rs_copy(const Ctx &c, const slide_rop &r, int tid) {
  const int64_t *f = r.i;
  const Opnd S = resolve(c, f + RC_SRC), D = resolve(c, f + RC_DST);
  int rows = (int)f[RC_ROWS];
  const int cols = (int)f[RC_COLS];
  int r0 = 0;
  if (f[RC_OWNED]) {
    r0 = c.p0 * (int)f[RC_RPP];
    rows = c.npl * (int)f[RC_RPP];
  }
  for (int e = tid; e < rows * cols; e += RT) {
    const int row = r0 + e / cols, col = e % cols;
    D.ptr[(size_t)row * D.ld + col] = S.ptr[(size_t)row * S.ld + col];
  }
  __syncthreads();
}

// ---- RS_DDPM -------------------------------------------------------------------------------------------------------
__device__ void rs_ddpm(const Ctx &c, const slide_rop &r, int tid) {
  const int64_t *f = r.i;
  const Opnd X = resolve(c, f + RD_X), XG = resolve(c, f + RD_XG), E = resolve(c, f + RD_EPS), X0C = resolve(c, f + RD_X0C),
             MK = resolve(c, f + RD_MASK);
  const int mode = (int)f[RD_MODE], ncols = (int)f[RD_NCOLS], col0 = (int)f[RD_COL0];
  const int64_t brows = f[RD_BROWS];  // rows of the whole batch (B * NP)
  const float *tab = wptr(c, f[RD_TABLE]) + (size_t)c.t * 8;
  const float *noise = reinterpret_cast<const float *>(c.arena + f[RD_NOISE]);
  const float clamp = r.f[0];
  const int t = c.t;
  for (int e = tid; e < c.npl * ncols; e += RT) {
    const int row = c.p0 + e / ncols, col = e % ncols;
    if (col < col0) continue;
    const int64_t grow = (int64_t)c.sample * c.np + row;
    const float xv = X.ptr[(size_t)row * X.ld + col], ev = E.ptr[(size_t)row * E.ld + col];
    const float nz = noise[((size_t)t * brows + grow) * ncols + col];
    float res;
    if (mode == 0) {
      res = __fdiv_rn(__fsub_rn(xv, __fmul_rn(tab[0], ev)), tab[1]);
      if (t > 0) res = __fadd_rn(res, __fmul_rn(tab[2], nz));
    } else if (mode == 2) {
      res = __fadd_rn(__fmul_rn(xv, tab[0]), __fadd_rn(__fmul_rn(tab[1], ev), __fmul_rn(tab[2], nz)));
    } else {
      float x0 = __fsub_rn(__fmul_rn(tab[0], xv), __fmul_rn(tab[1], ev));
      if (clamp > 0.f) x0 = fminf(fmaxf(x0, -clamp), clamp);
      if (X0C.ptr) {
        const float m = MK.ptr[row];
        x0 = __fadd_rn(__fmul_rn(x0, m), __fmul_rn(X0C.ptr[(size_t)row * X0C.ld + col], __fsub_rn(1.0f, m)));
      }
      const float mean = __fadd_rn(__fmul_rn(tab[2], x0), __fmul_rn(tab[3], xv));
      const float m = t == 0 ? 0.f : 1.f;
      res = __fadd_rn(mean, __fmul_rn(__fmul_rn(m, tab[4]), nz));
    }
    XG.ptr[(size_t)row * XG.ld + col] = res;
  }
  __syncthreads();
}

// ---- the kernel ----------------------------------------------------------------------------------------------------
template <int CL, bool PRECISE>
__global__ void __launch_bounds__(RT, 1) resident_kernel(const ResArgs a) {
  extern __shared__ __align__(16) float sm[];
  __shared__ slide_rop rop_s[2];
  const int tid = threadIdx.x;
  Ctx c;
  c.sm = sm;
  c.arena = a.arena;
  c.weights = a.weights;
  c.cl = CL;
  c.rank = CL > 1 ? (int)cluster_rank() : 0;
  c.sample = blockIdx.x / CL;
  c.np = a.plan.np;
  c.npl = a.plan.np / CL;
  c.p0 = c.rank * c.npl;
  c.t = *reinterpret_cast<const volatile int *>(a.arena + a.plan.step_off) - 1;
  float *wst = sm + a.plan.wstage_off;
  const int stage_floats = a.plan.wstage_floats;
  int ring = 0;  // chunks of the next GEMM already in flight

  // statistics region = 0; first rop into shared memory
  for (int i = tid; i < a.plan.stats_floats; i += RT) sm[a.plan.stats_off + i] = 0.f;
  constexpr int ROP_WORDS = sizeof(slide_rop) / 4;
  if (tid < ROP_WORDS) reinterpret_cast<uint32_t *>(&rop_s[0])[tid] = reinterpret_cast<const uint32_t *>(a.rops)[tid];
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // every CTA of the cluster is resident before any remote access

  for (int ri = 0; ri < a.n_rops; ++ri) {
    const slide_rop &r = rop_s[ri & 1];
    if (ri + 1 < a.n_rops && tid < ROP_WORDS)  // next record (visible after this rop's trailing barrier)
      reinterpret_cast<uint32_t *>(&rop_s[(ri + 1) & 1])[tid] =
          reinterpret_cast<const uint32_t *>(a.rops + ri + 1)[tid];
    switch (r.kind) {
      case RS_COPY: rs_copy(c, r, tid); break;
      case RS_KNN: rs_knn(c, r, tid); break;
      case RS_GEMM: rs_gemm<PRECISE>(c, r, wst, stage_floats, ring, tid); break;
      case RS_PAIR: rs_pair(c, r, tid); break;
      case RS_XFORM: rs_xform(c, r, tid); break;
      case RS_STATSX: {
        // partial sums of every CTA are final -> totals (separate buffer: partials are never rewritten)
        const int src = (int)r.i[RT_ST], dst = (int)r.i[RT_DST], n = (int)r.i[RT_NFLOATS];
        if (CL > 1) cluster_sync_all(); else __syncthreads();
        if (tid < n) {
          float v = sm[src + tid];
          if (CL > 1) {
            const uint32_t la = smem_u32(sm + src + tid);
            for (int pr = 0; pr < CL; ++pr)
              if (pr != c.rank) v += ld_peer(map_peer(la, pr));
          }
          sm[dst + tid] = v;
        }
        __syncthreads();
        break;
      }
      case RS_CSYNC:
        if (CL > 1) cluster_sync_all(); else __syncthreads();
        break;
      case RS_DDPM: rs_ddpm(c, r, tid); break;
      case RS_SPILL:
      case RS_FILL: {
        float4 *s4 = reinterpret_cast<float4 *>(sm + r.i[RL_SMEM]);
        float4 *g4 = reinterpret_cast<float4 *>(a.scratch + (size_t)blockIdx.x * a.plan.scratch_bytes + r.i[RL_SCRATCH]);
        const int n4 = (int)(r.i[RL_NFLOATS] / 4);
        if (r.kind == RS_SPILL)
          for (int i = tid; i < n4; i += RT) g4[i] = s4[i];
        else
          for (int i = tid; i < n4; i += RT) s4[i] = g4[i];
        __syncthreads();
        break;
      }
      default:
        __syncthreads();
    }
  }
  cp_wait<0>();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still touch its shared memory
  // SLIDE_OP_STEP_BEGIN's counter update: every CTA has read the counter before it started, so the last one to finish
  // may store t (= counter - 1) for the next launch
  if (tid == 0) {
    __threadfence();
    const int prev = atomicAdd(a.done, 1);
    if (prev == (int)gridDim.x - 1) {
      *reinterpret_cast<volatile int *>(a.arena + a.plan.step_off) = c.t;
      *a.done = 0;
      __threadfence();
    }
  }
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------------------------
struct ResidentPlan {
  slide_resident_plan hdr;
  slide_rop *rops = nullptr;  // device
  int n_rops = 0;
  char *scratch = nullptr;
  int *done = nullptr;
  size_t smem_bytes = 0;
};

void resident_free(ResidentPlan *p) {
  if (!p) return;
  if (p->rops) cudaFree(p->rops);
  if (p->scratch) cudaFree(p->scratch);
  if (p->done) cudaFree(p->done);
  delete p;
}

int resident_create(const slide_resident_plan *hdr, const slide_rop *rops, int n_rops, ResidentPlan **out) {
  if (!hdr || !rops || n_rops <= 0 || !out) return SLIDE_ERR_INVALID;
  if (hdr->cluster != 1 && hdr->cluster != 2 && hdr->cluster != 4) return SLIDE_ERR_UNSUPPORTED;
  if (hdr->np % hdr->cluster) return SLIDE_ERR_INVALID;
  ResidentPlan *p = new ResidentPlan();
  p->hdr = *hdr;
  p->n_rops = n_rops;
  p->smem_bytes = (size_t)hdr->smem_floats * 4;
  if (p->smem_bytes > 227 * 1024 - sizeof(slide_rop) * 2 - 64) {
    delete p;
    return SLIDE_ERR_UNSUPPORTED;
  }
  int rc = cuda_rc(cudaMalloc((void **)&p->rops, sizeof(slide_rop) * n_rops));
  if (rc == SLIDE_OK) rc = cuda_rc(cudaMemcpy(p->rops, rops, sizeof(slide_rop) * n_rops, cudaMemcpyHostToDevice));
  if (rc == SLIDE_OK) rc = cuda_rc(cudaMalloc((void **)&p->done, 256));
  if (rc == SLIDE_OK) rc = cuda_rc(cudaMemset(p->done, 0, 256));
  const size_t sbytes = (size_t)hdr->scratch_bytes * hdr->batch * hdr->cluster;
  if (rc == SLIDE_OK && sbytes) rc = cuda_rc(cudaMalloc((void **)&p->scratch, sbytes));
  if (rc != SLIDE_OK) {
    resident_free(p);
    return rc;
  }
  *out = p;
  return SLIDE_OK;
}

template <int CL, bool PRECISE>
static int launch_resident_t(const ResidentPlan *p, const ResArgs &a, cudaStream_t st) {
  static bool configured[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  auto kern = resident_kernel<CL, PRECISE>;
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)sizeof(slide_rop) * 2 - 64);
    if (e != cudaSuccess) return cuda_rc(e);
    configured[dev] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(p->hdr.batch * CL);
  cfg.blockDim = dim3(RT);
  cfg.dynamicSmemBytes = p->smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) return cuda_rc(e);
  return after_launch();
}

int resident_launch(const ResidentPlan *p, char *arena, const char *weights, cudaStream_t st) {
  ResArgs a;
  a.rops = p->rops;
  a.n_rops = p->n_rops;
  a.arena = arena;
  a.weights = weights;
  a.scratch = p->scratch;
  a.done = p->done;
  a.plan = p->hdr;
  const bool pr = p->hdr.precise != 0;
  switch (p->hdr.cluster) {
    case 1: return pr ? launch_resident_t<1, true>(p, a, st) : launch_resident_t<1, false>(p, a, st);
    case 2: return pr ? launch_resident_t<2, true>(p, a, st) : launch_resident_t<2, false>(p, a, st);
    case 4: return pr ? launch_resident_t<4, true>(p, a, st) : launch_resident_t<4, false>(p, a, st);
  }
  return SLIDE_ERR_UNSUPPORTED;
}

int resident_first(const ResidentPlan *p) { return p->hdr.first; }
int resident_count(const ResidentPlan *p) { return p->hdr.count; }

}  // namespace slide
