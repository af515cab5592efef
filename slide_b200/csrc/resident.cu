// Sample-resident executor (include/slide_resident.h): one kernel launch per diffusion step for the denoisers over
// 16 latent points.  One thread-block cluster owns one sample; its activations stay in shared memory from the first
// layer to the update of x, GroupNorm statistics are reduced in shared memory and completed across the cluster through
// distributed shared memory, weights stream from L2 through a 3-stage cp.async ring (with the next layer's first chunks
// already in flight while the current layer's epilogue runs).
//
// Arithmetic: the contractions here are small (K, N <= 139 for the position DDPM) and every operand is on chip, so the
// kernel is bound by INSTRUCTION ISSUE, not by the tensor pipe or memory (ncu: tensor pipe < 5 %, issue slots ~ 50 %
// busy with 4 warps per scheduler).  Hence: the dense products run on the warp-level tensor path (mma.sync m16n8k8, TF32
// operands rounded to nearest, fp32 accumulate -- the numerics class of the reference's cuDNN convs) fed by ldmatrix from
// padded shared-memory rows (row stride = 4 mod 8 floats: conflict-free); tile counts are compile-time so the inner loop
// is ldmatrix + cvt + mma only; records are fully resolved on the host (32-bit offsets, shifts instead of divisions);
// element-wise passes use 16-byte accesses; statistics go through per-warp column partials instead of shared-memory
// float atomics (CAS loops on this architecture).  PRECISE plans split both operands (3xTF32) and match the fp32 oracle
// to ~1e-6.  Replaces 62 launches per position-DDPM step of the per-record executor (program.cu).
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
constexpr int NBLK_PAD = SLIDE_RES_NBLK + 8;  // column slots of the statistics scratch (and rows of a ring stage)

struct ResArgs {
  const slide_rop *rops;
  int n_rops;
  float *arena;
  const float *weights;
  float *scratch;  // [grid][scratch_floats]
  int *done;       // CTAs finished (for the step-counter update)
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

struct Ctx {
  float *arena;
  const float *weights;
  int sample, rank, cl, np, npl, p0, t;
};

// ---- weight-chunk ring ---------------------------------------------------------------------------------------------
// Chunk q of a GEMM's packed weight copy ([nchunk][npad][WLD] floats) -> stage q % WSTAGES.  Every thread commits exactly
// one group per issued chunk, so cp.async.wait_group counts are uniform across the CTA.
__device__ __forceinline__ void issue_chunk(const float *wsrc, uint32_t wst_u, int stage_bytes, int n16, int q, int tid) {
  const float4 *src = reinterpret_cast<const float4 *>(wsrc) + q * n16;
  const uint32_t dst = wst_u + (uint32_t)((q % WSTAGES) * stage_bytes);
  if (tid < n16) cp_async16(dst + tid * 16, src + tid);
  if (tid + RT < n16) cp_async16(dst + (tid + RT) * 16, src + tid + RT);
  cp_commit();
}

// ---- column statistics ---------------------------------------------------------------------------------------------
// Per-(slot, column) partial sums {sum, sum of squares} parked in `scr` ([nslots][NBLK_PAD][2] floats; a free stage of the
// weight ring) are folded into the statistics buffer by ONE thread per group.  Group gi covers channels
// [gi*cg, (gi+1)*cg) = local columns [gi*cg - choff, ...) clipped to [0, lim): the first and last group of the range may
// be shared with another producer (q | k concatenations, column blocks), which ran in an earlier rop.
__device__ __forceinline__ void fold_colsum(float *st, const float *scr, int nslots, int ncols, int cg, int nnorm, int choff,
                                            float weight, int tid) {
  const int lim = min(ncols, nnorm - choff);
  if (tid >= 64 || lim <= 0) return;  // <= 32 groups per tensor (+ a shared edge group)
  const int g_first = choff / cg, g_last = (choff + lim - 1) / cg;
  if (tid <= g_last - g_first) {
    const int gi = g_first + tid;
    const int c0 = max(0, gi * cg - choff), c1 = min(lim, (gi + 1) * cg - choff);
    float s = 0.f, q = 0.f;
    for (int slot = 0; slot < nslots; ++slot)
      for (int col = c0; col < c1; ++col) {
        const float2 v = *reinterpret_cast<const float2 *>(scr + (slot * NBLK_PAD + col) * 2);
        s += v.x;
        q += v.y;
      }
    st[2 * gi] += weight * s;
    st[2 * gi + 1] += weight * q;
  }
}

// ---- RS_GEMM -------------------------------------------------------------------------------------------------------
// Warp grid over (row tiles of 16, column tiles of 8): 8 row tiles -> 4 x 4 warps with 2 row tiles each, 4 -> 4 x 4,
// 2 -> 2 x 8, 1 -> 1 x 16.  More than 128 rows: passes of 128 rows (the weight chunks are streamed again).
// MT / NT (row / column tiles per warp) are compile-time: the inner loop is branch-free (ldmatrix + cvt + mma only).
template <int MT, int NT, bool PRECISE>
__device__ __forceinline__ void gemm_pass(const Ctx &c, const int *f, float stw, float *sm, float *wst, int stage_floats,
                                          int &ring, int tid, int m_pass, int wgm, bool last_pass) {
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int K = f[RG_K], N = f[RG_N], nchunk = f[RG_NCHUNK];
  const int n_tiles = (N + 7) >> 3;
  const int wm = warp % wgm, wn = warp / wgm;
  const int tile0 = wn * NT;
  const int row0 = m_pass + wm * MT * 16;
  const bool active = tile0 < n_tiles;
  const int lda = f[RG_ALD];

  float acc[MT][NT][4];
#pragma unroll
  for (int a = 0; a < MT; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;

  const float *wsrc = c.weights + f[RG_WCH];
  const int n16 = f[RG_NPAD] * (WLD / 4);
  const uint32_t wst_u = smem_u32(wst);
  const int stage_bytes = stage_floats * 4;
  int issued = (m_pass == 0) ? min(ring, nchunk) : 0;
  if (m_pass == 0) ring = 0;
  while (issued < nchunk && issued < 2) {  // chunks 0 and 1 (unless the previous GEMM's tail already sent them)
    issue_chunk(wsrc, wst_u, stage_bytes, n16, issued, tid);
    ++issued;
  }
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = ((lane >> 4) & 1) * 4;
  const int b_row = (lane & 7) + ((lane >> 4) & 1) * 8, b_col = ((lane >> 3) & 1) * 4;
  uint32_t a_addr = smem_u32(sm + f[RG_A] + (row0 + a_row) * lda + a_col);
  const uint32_t a_mt = (uint32_t)(16 * lda * 4);
  const uint32_t b_off = (uint32_t)(((tile0 * 8 + b_row) * WLD + b_col) * 4);
  for (int q = 0; q < nchunk; ++q) {
    // pending groups here: chunk q, and chunk q+1 if it exists (issued one iteration ago / by the prologue)
    if (q + 1 < nchunk) cp_wait<1>(); else cp_wait<0>();
    __syncthreads();  // chunk q visible to all warps; everyone is done with chunk q-1, whose stage takes chunk q+2
    if (q + 2 < nchunk && q + 2 >= issued) issue_chunk(wsrc, wst_u, stage_bytes, n16, q + 2, tid);
    if (active) {
      const uint32_t ws = wst_u + (uint32_t)((q % WSTAGES) * stage_bytes) + b_off;
#pragma unroll
      for (int ks = 0; ks < WCH / 8; ++ks) {
        const int k0 = q * WCH + ks * 8;
        if (k0 < K) {
          uint32_t ah[MT][4], al[MT][4];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            uint32_t r0, r1, r2, r3;
            ldsm4(a_addr + mt * a_mt + ks * 32, r0, r1, r2, r3);
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
#pragma unroll
          for (int jp = 0; jp < (NT + 1) / 2; ++jp) {
            uint32_t b[4];
            ldsm4(ws + (uint32_t)((jp * 16 * WLD + ks * 8) * 4), b[0], b[1], b[2], b[3]);
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = jp * 2 + jj;
              if (j < NT) {
                uint32_t bh0 = b[jj * 2], bh1 = b[jj * 2 + 1];
                if (PRECISE) {
                  const float f0 = __uint_as_float(bh0), f1 = __uint_as_float(bh1);
                  bh0 = tf32_rna(f0);
                  bh1 = tf32_rna(f1);
                  const uint32_t bl0 = tf32_rna(f0 - __uint_as_float(bh0)), bl1 = tf32_rna(f1 - __uint_as_float(bh1));
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt) {
                    mma_tf32(acc[mt][j], al[mt][0], al[mt][1], al[mt][2], al[mt][3], bh0, bh1);
                    mma_tf32(acc[mt][j], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bl0, bl1);
                  }
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][j], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bh0, bh1);
              }
            }
          }
        }
      }
      a_addr += WCH * 4;
    }
  }
  __syncthreads();  // every warp is done with the last chunk: the ring may be refilled
  // the next GEMM's first chunks go in flight under this epilogue (and any rops in between)
  if (last_pass && f[RG_NEXT_WCH] >= 0) {
    ring = min(f[RG_NEXT_PF], f[RG_NEXT_NCHUNK]);
    const float *nsrc = c.weights + f[RG_NEXT_WCH];
    const int nn16 = f[RG_NEXT_NPAD] * (WLD / 4);
    for (int q = 0; q < ring; ++q) issue_chunk(nsrc, wst_u, stage_bytes, nn16, q, tid);
  }

  // ---- epilogue ----
  float *scr = wst + 2 * stage_floats;  // stage 2 is free between the main loop and the next GEMM's third chunk
  const int st_off = f[RG_ST];
  const bool stats = st_off >= 0;
  if (active) {
    const float *bias = f[RG_BIAS] >= 0 ? c.weights + f[RG_BIAS] : nullptr;
    const int smk = f[RG_SMK];
    float *Cp = sm + f[RG_C];
    const int ldc = f[RG_CLD];
    const float *Rp = f[RG_RES] >= 0 ? sm + f[RG_RES] : nullptr;
    const int ldr = f[RG_RESLD];
    const float *Ep = f[RG_EV] >= 0 ? sm + f[RG_EV] : nullptr;
    const int lde = f[RG_EVLD];
    const int pairrows = f[RG_PAIRROWS], rshift = f[RG_RPP_SHIFT], act = f[RG_ACT], st_owned = f[RG_ST_OWNED];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int tile = tile0 + j;
      if (tile >= n_tiles) continue;
      const int n0 = tile * 8 + tq * 2;
      const bool v0 = n0 < N, v1 = n0 + 1 < N;
      const float b0 = (bias && v0) ? __ldg(bias + n0) : 0.f, b1 = (bias && v1) ? __ldg(bias + n0 + 1) : 0.f;
      if (smk > 0) {
        // fused soft-max over the smk (8 or 16) rows of each point, applied to the value rows RES
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int rbase = row0 + mt * 16;
          float s[4] = {acc[mt][j][0] + b0, acc[mt][j][1] + b1, acc[mt][j][2] + b0, acc[mt][j][3] + b1};
          float val[4];
          const float *vr = Rp + (rbase + g) * ldr + n0;
          val[0] = v0 ? vr[0] : 0.f;
          val[1] = v1 ? vr[1] : 0.f;
          val[2] = v0 ? vr[8 * ldr] : 0.f;
          val[3] = v1 ? vr[8 * ldr + 1] : 0.f;
          float mx[4] = {s[0], s[1], s[2], s[3]};
          if (smk == 16) {
            mx[0] = mx[2] = fmaxf(s[0], s[2]);
            mx[1] = mx[3] = fmaxf(s[1], s[3]);
          }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1)
#pragma unroll
            for (int e = 0; e < 4; ++e) mx[e] = fmaxf(mx[e], __shfl_xor_sync(0xffffffffu, mx[e], o));
          float num[4], den[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            den[e] = expf(s[e] - mx[e]);
            num[e] = den[e] * val[e];
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
              const int prow = c.p0 + ((rbase + h * 8) >> rshift);
              float *dst = Cp + prow * ldc + n0;
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
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = row0 + mt * 16 + g + h * 8;
          float x0 = acc[mt][j][h * 2] + b0, x1 = acc[mt][j][h * 2 + 1] + b1;
          const int point = pairrows ? c.p0 + (row >> rshift) : row;
          if (Ep) {
            const float *e = Ep + point * lde + n0;
            if (v0) x0 += e[0];
            if (v1) x1 += e[1];
          }
          if (Rp) {
            const float *e = Rp + row * ldr + n0;
            if (v0) x0 += e[0];
            if (v1) x1 += e[1];
          }
          if (act == 1) {
            x0 = fmaxf(x0, 0.f);
            x1 = fmaxf(x1, 0.f);
          }
          float *d = Cp + row * ldc + n0;
          if (v1) *reinterpret_cast<float2 *>(d) = make_float2(x0, x1);  // n0 even, ld even, base 16-byte aligned
          else if (v0) d[0] = x0;
          const bool counted = !st_owned || (point >= c.p0 && point < c.p0 + c.npl);
          if (stats && counted) {
            if (v0) { s0 += x0; q0 += x0 * x0; }
            if (v1) { s1 += x1; q1 += x1 * x1; }
          }
        }
      }
      if (stats) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          q0 += __shfl_xor_sync(0xffffffffu, q0, o);
          q1 += __shfl_xor_sync(0xffffffffu, q1, o);
        }
        if (g == 0) *reinterpret_cast<float4 *>(scr + (wm * NBLK_PAD + n0) * 2) = make_float4(s0, q0, s1, q1);
      }
    }
  }
  if (stats) {
    __syncthreads();
    fold_colsum(sm + st_off, scr, wgm, N, f[RG_ST_CG], f[RG_ST_NNORM], f[RG_ST_CHOFF], stw, tid);
  }
}

template <bool PRECISE>
__device__ void rs_gemm(const Ctx &c, const slide_rop &r, float *sm, float *wst, int stage_floats, int &ring, int tid) {
  const int *f = r.i;
  const int M = f[RG_M];
  const int n_tiles = (f[RG_N] + 7) >> 3;
  for (int m_pass = 0; m_pass < M; m_pass += 128) {
    const int m_tiles = min(128, M - m_pass) >> 4;
    const bool last = m_pass + 128 >= M;
#define GP(MT_, NT_, WGM_) gemm_pass<MT_, NT_, PRECISE>(c, f, r.f[0], sm, wst, stage_floats, ring, tid, m_pass, WGM_, last)
    if (m_tiles >= 8) {         // 4 x 4 warps, 2 row tiles each
      const int nt = (n_tiles + 3) >> 2;
      if (nt <= 1) GP(2, 1, 4); else if (nt == 2) GP(2, 2, 4); else if (nt == 3) GP(2, 3, 4); else GP(2, 4, 4);
    } else if (m_tiles >= 4) {  // 4 x 4 warps
      const int nt = (n_tiles + 3) >> 2;
      if (nt <= 1) GP(1, 1, 4); else if (nt == 2) GP(1, 2, 4); else if (nt == 3) GP(1, 3, 4); else GP(1, 4, 4);
    } else if (m_tiles >= 2) {  // 2 x 8 warps
      if (n_tiles <= 8) GP(1, 1, 2); else GP(1, 2, 2);
    } else {                    // 1 x 16 warps
      GP(1, 1, 1);
    }
#undef GP
  }
  __syncthreads();
}

// ---- RS_PAIR -------------------------------------------------------------------------------------------------------
// Lanes own 4 consecutive output columns (16-byte accesses), warps own row slices; per-row scalars (neighbour index,
// coordinates, d2, interpolation weight) are computed once into the scratch stage.
__device__ void rs_pair(const Ctx &c, const slide_rop &r, float *sm, float *scr, int tid) {
  const int *f = r.i;
  const float *Up = sm + f[RP_U], *Xp = sm + f[RP_XYZ], *Ctp = sm + f[RP_CTR];
  float *Op = sm + f[RP_OUT];
  const float *Rp = f[RP_RES] >= 0 ? sm + f[RP_RES] : nullptr;
  const int ldu = f[RP_ULD], ldx = f[RP_XLD], ldct = f[RP_CLD], ldo = f[RP_OLD], ldr = f[RP_RLD];
  const int K = f[RP_K], N = f[RP_N], act = f[RP_ACT];
  const int *idx = reinterpret_cast<const int *>(sm + f[RP_IDX]);
  const float *d2 = f[RP_D2] >= 0 ? sm + f[RP_D2] : nullptr;
  const float *wx = c.weights + f[RP_WX], *wc = c.weights + f[RP_WC];
  const float *wd = f[RP_WD] >= 0 ? c.weights + f[RP_WD] : nullptr, *ww = f[RP_WW] >= 0 ? c.weights + f[RP_WW] : nullptr;
  const float *bias = f[RP_BIAS] >= 0 ? c.weights + f[RP_BIAS] : nullptr;
  const int st_off = f[RP_ST];
  const int warp = tid >> 5, lane = tid & 31;
  const int rows = c.npl * K;  // <= 128
  // row table: [rows][12] = {j, xj(3), ci(3), d2, w, -, -, -}
  float *rowtab = scr;               // 128 * 12 = 1536 floats
  float *colsum = scr + 128 * 12;    // [8][NBLK_PAD][2] = 2176 floats
  if (tid < rows) {
    const int il = tid / K, k = tid - il * K;
    const int pi = c.p0 + il;
    const int j = idx[pi * K + k];
    float *t = rowtab + tid * 12;
    t[0] = __int_as_float(j);
    t[1] = Xp[j * ldx];
    t[2] = Xp[j * ldx + 1];
    t[3] = Xp[j * ldx + 2];
    t[4] = Ctp[pi * ldct];
    t[5] = Ctp[pi * ldct + 1];
    t[6] = Ctp[pi * ldct + 2];
    float dd = 0.f, w = 0.f;
    if (d2) {
      float sum = 0.f;
      for (int kk = 0; kk < K; ++kk) sum += 1.0f / (d2[pi * K + kk] + 1e-8f);
      dd = d2[pi * K + k];
      w = (1.0f / (dd + 1e-8f)) / sum;
    }
    t[7] = dd;
    t[8] = w;
  }
  __syncthreads();
  const bool stats = st_off >= 0;
  for (int n0 = 0; n0 < N; n0 += 128) {
    const int n = n0 + lane * 4;
    float wxx[4], wxy[4], wxz[4], wcx[4], wcy[4], wcz[4], bb[4], wdd[4], www[4];
    bool v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[e] = n + e < N;
      const int ne = v[e] ? n + e : 0;
      wxx[e] = __ldg(wx + ne * 3); wxy[e] = __ldg(wx + ne * 3 + 1); wxz[e] = __ldg(wx + ne * 3 + 2);
      wcx[e] = __ldg(wc + ne * 3); wcy[e] = __ldg(wc + ne * 3 + 1); wcz[e] = __ldg(wc + ne * 3 + 2);
      bb[e] = bias ? __ldg(bias + ne) : 0.f;
      wdd[e] = d2 ? __ldg(wd + ne) : 0.f;
      www[e] = d2 ? __ldg(ww + ne) : 0.f;
    }
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (v[0]) {
      for (int lr = warp; lr < rows; lr += NW) {
        const float4 t0 = *reinterpret_cast<const float4 *>(rowtab + lr * 12);
        const float4 t1 = *reinterpret_cast<const float4 *>(rowtab + lr * 12 + 4);
        const float t8 = rowtab[lr * 12 + 8];
        const int j = __float_as_int(t0.x);
        const float4 u = *reinterpret_cast<const float4 *>(Up + j * ldu + n);
        float val[4] = {u.x, u.y, u.z, u.w};
        float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (Rp) rs = *reinterpret_cast<const float4 *>(Rp + lr * ldr + n);
        const float rr[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float a = val[e];
          a += t0.y * wxx[e] + t0.z * wxy[e] + t0.w * wxz[e];
          a += t1.x * wcx[e] + t1.y * wcy[e] + t1.z * wcz[e];
          a += bb[e];
          a += t1.w * wdd[e] + t8 * www[e];
          a += rr[e];
          if (act == 1) a = fmaxf(a, 0.f);
          val[e] = a;
          if (v[e]) {
            s[e] += a;
            q[e] += a * a;
          }
        }
        float *o = Op + lr * ldo + n;
        if (v[3]) {
          *reinterpret_cast<float4 *>(o) = make_float4(val[0], val[1], val[2], val[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 3; ++e)
            if (v[e]) o[e] = val[e];
        }
      }
    }
    if (stats) {
      // 16 row slices -> 8 (warps 8..15 hand their partials to warps 0..7) -> one thread per group
      const int nl = n - n0;  // local column 0..127
      if (warp >= 8 && v[0]) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          *reinterpret_cast<float2 *>(colsum + ((warp - 8) * NBLK_PAD + nl + e) * 2) = make_float2(s[e], q[e]);
      }
      __syncthreads();
      if (warp < 8 && v[0]) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 *slot = reinterpret_cast<float2 *>(colsum + (warp * NBLK_PAD + nl + e) * 2);
          const float2 o = *slot;
          *slot = make_float2(o.x + s[e], o.y + q[e]);
        }
      }
      __syncthreads();
      fold_colsum(sm + st_off, colsum, 8, min(128, N - n0), f[RP_ST_CG], f[RP_ST_NNORM], f[RP_ST_CHOFF] + n0, 1.0f, tid);
      __syncthreads();
    }
  }
  __syncthreads();
}

// ---- RS_XFORM ------------------------------------------------------------------------------------------------------
// Per-column constants (scale, shift, additive vector) are computed ONCE by the first C threads into the scratch stage;
// then lanes own 4 consecutive columns, warps own row slices.
__device__ void rs_xform(const Ctx &c, const slide_rop &r, float *sm, float *scr, int tid) {
  const int *f = r.i;
  float *Xp = sm + f[RX_X];
  const int ldx = f[RX_XLD], rows = f[RX_ROWS], C = f[RX_C];
  float *ca = scr, *cb = scr + 640, *cadd = scr + 1280;  // C <= 640 columns
  for (int n = tid; n < C; n += RT) {
    float a = 1.f, b = 0.f, av = 0.f;
    const int st_off = f[RX_ST];
    const int ch = f[RX_CHOFF] + n;
    if (st_off >= 0 && ch < f[RX_NNORM]) {
      const float inv_count = r.f[0];
      const float2 sq = *reinterpret_cast<const float2 *>(sm + st_off + 2 * (ch / f[RX_CG]));
      const float mean = sq.x * inv_count;
      float var = sq.y * inv_count - mean * mean;
      var = var < 0.f ? 0.f : var;
      const float rstd = 1.0f / sqrtf(var + SLIDE_GN_EPS);
      a = rstd * __ldg(c.weights + f[RX_GAMMA] + ch);
      b = __ldg(c.weights + f[RX_BETA] + ch) - mean * a;
    }
    if (f[RX_ADD] >= 0) {
      const int mode = f[RX_ADDMODE];
      const long long arow = mode == 0 ? c.sample : (mode == 1 ? c.t : 0);
      av = __ldg(c.arena + f[RX_ADD] + arow * f[RX_ADDLD] + n);
    }
    ca[n] = a;
    cb[n] = b;
    cadd[n] = av;
  }
  __syncthreads();
  const int relu = f[RX_RELU];
  const int warp = tid >> 5, lane = tid & 31;
  for (int n0 = 0; n0 < C; n0 += 128) {
    const int n = n0 + lane * 4;
    if (n >= C) continue;
    const float4 a4 = *reinterpret_cast<const float4 *>(ca + n), b4 = *reinterpret_cast<const float4 *>(cb + n),
                 d4 = *reinterpret_cast<const float4 *>(cadd + n);
    const bool full = n + 3 < C;
    for (int row = warp; row < rows; row += NW) {
      float *px = Xp + row * ldx + n;
      float4 x = *reinterpret_cast<const float4 *>(px);
      x.x = x.x * a4.x + b4.x;
      x.y = x.y * a4.y + b4.y;
      x.z = x.z * a4.z + b4.z;
      x.w = x.w * a4.w + b4.w;
      if (relu) {
        x.x = fmaxf(x.x, 0.f);
        x.y = fmaxf(x.y, 0.f);
        x.z = fmaxf(x.z, 0.f);
        x.w = fmaxf(x.w, 0.f);
      }
      x.x += d4.x;
      x.y += d4.y;
      x.z += d4.z;
      x.w += d4.w;
      if (full) {
        *reinterpret_cast<float4 *>(px) = x;
      } else {
        px[0] = x.x;
        if (n + 1 < C) px[1] = x.y;
        if (n + 2 < C) px[2] = x.z;
      }
    }
  }
  __syncthreads();
}

// ---- RS_KNN --------------------------------------------------------------------------------------------------------
// One thread per (query, reference) pair: its rank among the query's distances (strict <, ties by ascending index -- the
// order pytorch3d's sorted insertion produces) is the slot it is written to.
__device__ void rs_knn(const Ctx &c, const slide_rop &r, float *sm, float *scr, int tid) {
  const int *f = r.i;
  const float *Q = sm + f[RK_Q], *R = sm + f[RK_REF];
  const int ldq = f[RK_QLD], ldr = f[RK_RLD];
  const int P1 = f[RK_P1], P2 = f[RK_P2], K = f[RK_K];
  int *idx = reinterpret_cast<int *>(sm + f[RK_IDX]);
  float *d2 = f[RK_D2] >= 0 ? sm + f[RK_D2] : nullptr;
  float *dist = scr;  // [P1][P2] <= 16 x 32
  const int q = tid / P2, p = tid - q * P2;
  float d = 0.f;
  if (tid < P1 * P2) {
    d = sumsq3_p3d(Q[q * ldq] - R[p * ldr], Q[q * ldq + 1] - R[p * ldr + 1], Q[q * ldq + 2] - R[p * ldr + 2]);
    dist[tid] = d;
  }
  __syncthreads();
  if (tid < P1 * P2) {
    int rank = 0;
    for (int o = 0; o < P2; ++o) {
      const float od = dist[q * P2 + o];
      rank += (od < d || (od == d && o < p)) ? 1 : 0;
    }
    if (rank < K) {
      idx[q * K + rank] = p;
      if (d2) d2[q * K + rank] = d;
    }
  }
  __syncthreads();
}

// ---- RS_COPY -------------------------------------------------------------------------------------------------------
__device__ void rs_copy(const Ctx &c, const slide_rop &r, float *sm, int tid) {
  const int *f = r.i;
  const float *S = f[RC_SRC_G] ? c.arena + f[RC_SRC] + (long long)c.sample * f[RC_SSTRIDE] : sm + f[RC_SRC];
  float *D = f[RC_DST_G] ? c.arena + f[RC_DST] + (long long)c.sample * f[RC_DSTRIDE] : sm + f[RC_DST];
  const int lds = f[RC_SLD], ldd = f[RC_DLD], cols = f[RC_COLS];
  int rows = f[RC_ROWS], r0 = 0;
  if (f[RC_OWNED]) {
    r0 = c.p0;
    rows = c.npl;
  }
  for (int e = tid; e < rows * cols; e += RT) {
    const int row = r0 + e / cols, col = e % cols;
    D[row * ldd + col] = S[row * lds + col];
  }
  __syncthreads();
}

// ---- RS_DDPM -------------------------------------------------------------------------------------------------------
__device__ void rs_ddpm(const Ctx &c, const slide_rop &r, float *sm, int tid) {
  const int *f = r.i;
  const float *X = sm + f[RD_X], *E = sm + f[RD_EPS];
  float *XG = c.arena + f[RD_XG] + (long long)c.sample * f[RD_XGSTRIDE];
  const float *X0C = f[RD_X0C] >= 0 ? c.arena + f[RD_X0C] + (long long)c.sample * f[RD_X0CSTRIDE] : nullptr;
  const float *MK = f[RD_MASK] >= 0 ? c.arena + f[RD_MASK] + (long long)c.sample * f[RD_MASKSTRIDE] : nullptr;
  const int mode = f[RD_MODE], ncols = f[RD_NCOLS], col0 = f[RD_COL0];
  const long long brows = f[RD_BROWS];  // rows of the whole batch (B * NP)
  const int t = c.t;
  const float *tab = c.weights + f[RD_TABLE] + t * 8;
  const float *noise = c.arena + f[RD_NOISE] + ((long long)t * brows + (long long)c.sample * c.np) * ncols;
  const float clamp = r.f[0];
  for (int e = tid; e < c.npl * ncols; e += RT) {
    const int row = c.p0 + e / ncols, col = e % ncols;
    if (col < col0) continue;
    const float xv = X[row * f[RD_XLD] + col], ev = E[row * f[RD_ELD] + col];
    const float nz = noise[row * ncols + col];
    float res;
    if (mode == 0) {
      res = __fdiv_rn(__fsub_rn(xv, __fmul_rn(tab[0], ev)), tab[1]);
      if (t > 0) res = __fadd_rn(res, __fmul_rn(tab[2], nz));
    } else if (mode == 2) {
      res = __fadd_rn(__fmul_rn(xv, tab[0]), __fadd_rn(__fmul_rn(tab[1], ev), __fmul_rn(tab[2], nz)));
    } else {
      float x0 = __fsub_rn(__fmul_rn(tab[0], xv), __fmul_rn(tab[1], ev));
      if (clamp > 0.f) x0 = fminf(fmaxf(x0, -clamp), clamp);
      if (X0C) {
        const float m = MK[row];
        x0 = __fadd_rn(__fmul_rn(x0, m), __fmul_rn(X0C[row * f[RD_X0CLD] + col], __fsub_rn(1.0f, m)));
      }
      const float mean = __fadd_rn(__fmul_rn(tab[2], x0), __fmul_rn(tab[3], xv));
      const float m = t == 0 ? 0.f : 1.f;
      res = __fadd_rn(mean, __fmul_rn(__fmul_rn(m, tab[4]), nz));
    }
    XG[row * f[RD_XGLD] + col] = res;
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
  c.arena = a.arena;
  c.weights = a.weights;
  c.cl = CL;
  c.rank = CL > 1 ? (int)cluster_rank() : 0;
  c.sample = blockIdx.x / CL;
  c.np = a.plan.np;
  c.npl = a.plan.np / CL;
  c.p0 = c.rank * c.npl;
  int *step_ptr = reinterpret_cast<int *>(reinterpret_cast<char *>(a.arena) + a.plan.step_off);
  c.t = *reinterpret_cast<const volatile int *>(step_ptr) - 1;
  float *wst = sm + a.plan.wstage_off;
  const int stage_floats = a.plan.wstage_floats;
  int ring = 0;  // chunks of the next GEMM already in flight

  // statistics region = 0; first rop into shared memory
  for (int i = tid; i < a.plan.stats_floats; i += RT) sm[a.plan.stats_off + i] = 0.f;
  constexpr int ROP_WORDS = sizeof(slide_rop) / 4;
  if (tid < ROP_WORDS) reinterpret_cast<uint32_t *>(&rop_s[0])[tid] = reinterpret_cast<const uint32_t *>(a.rops)[tid];
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // every CTA of the cluster is resident before any remote access

  const int n_rops = a.n_rops;
  for (int ri = 0; ri < n_rops; ++ri) {
    const slide_rop &r = rop_s[ri & 1];
    if (ri + 1 < n_rops && tid < ROP_WORDS)  // next record (visible after this rop's trailing barrier)
      reinterpret_cast<uint32_t *>(&rop_s[(ri + 1) & 1])[tid] =
          reinterpret_cast<const uint32_t *>(a.rops + ri + 1)[tid];
    switch (r.kind) {
      case RS_GEMM: rs_gemm<PRECISE>(c, r, sm, wst, stage_floats, ring, tid); break;
      case RS_XFORM: rs_xform(c, r, sm, wst + 2 * stage_floats, tid); break;
      case RS_PAIR: rs_pair(c, r, sm, wst + stage_floats, tid); break;  // stages 1 + 2
      case RS_COPY: rs_copy(c, r, sm, tid); break;
      case RS_KNN: rs_knn(c, r, sm, wst + 2 * stage_floats, tid); break;
      case RS_STATSX: {
        // partial sums of every CTA are final -> totals (separate buffer: partials are never rewritten)
        const int src = r.i[RT_ST], dst = r.i[RT_DST], n = r.i[RT_NFLOATS];
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
      case RS_DDPM: rs_ddpm(c, r, sm, tid); break;
      case RS_SPILL:
      case RS_FILL: {
        float4 *s4 = reinterpret_cast<float4 *>(sm + r.i[RL_SMEM]);
        float4 *g4 = reinterpret_cast<float4 *>(a.scratch + (size_t)blockIdx.x * (a.plan.scratch_bytes / 4) + r.i[RL_SCRATCH]);
        const int n4 = r.i[RL_NFLOATS] / 4;
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
      *reinterpret_cast<volatile int *>(step_ptr) = c.t;
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
  float *scratch = nullptr;
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

constexpr int RES_MAX_DYN_SMEM = 227 * 1024 - (int)sizeof(slide_rop) * 2 - 64;

int resident_create(const slide_resident_plan *hdr, const slide_rop *rops, int n_rops, ResidentPlan **out) {
  if (!hdr || !rops || n_rops <= 0 || !out) return SLIDE_ERR_INVALID;
  if (hdr->cluster != 2 && hdr->cluster != 4) return SLIDE_ERR_UNSUPPORTED;
  if (hdr->np % hdr->cluster) return SLIDE_ERR_INVALID;
  ResidentPlan *p = new ResidentPlan();
  p->hdr = *hdr;
  p->n_rops = n_rops;
  p->smem_bytes = (size_t)hdr->smem_floats * 4;
  if (p->smem_bytes > (size_t)RES_MAX_DYN_SMEM) {
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
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RES_MAX_DYN_SMEM);
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
  a.arena = reinterpret_cast<float *>(arena);
  a.weights = reinterpret_cast<const float *>(weights);
  a.scratch = p->scratch;
  a.done = p->done;
  a.plan = p->hdr;
  const bool pr = p->hdr.precise != 0;
  switch (p->hdr.cluster) {
    case 2: return pr ? launch_resident_t<2, true>(p, a, st) : launch_resident_t<2, false>(p, a, st);
    case 4: return pr ? launch_resident_t<4, true>(p, a, st) : launch_resident_t<4, false>(p, a, st);
  }
  return SLIDE_ERR_UNSUPPORTED;
}

int resident_first(const ResidentPlan *p) { return p->hdr.first; }
int resident_count(const ResidentPlan *p) { return p->hdr.count; }

}  // namespace slide
