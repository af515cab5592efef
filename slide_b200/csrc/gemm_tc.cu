// tcgen05 GEMM for sm_100a: C = act( xfA(A) W^T + bias + ev + xfR(resid) ) with TF32 operands, fp32 accumulation
// in tensor memory, and the fused GroupNorm prologue / statistics epilogue of SLIDE_OP_GEMM.
//
// Two kernels share the operand formats and the epilogue arithmetic:
//   gemm_tcp_kernel  persistent and warp-specialised (copy warp, MMA warp, 8 epilogue warps, optionally 8 transform
//                    warps), double-buffered TMEM accumulator; taken when a launch has >= 148 tiles that each lie
//                    inside one sample (the pair-level layers of the DDPMs at batch 256, most decode / encode layers);
//   gemm_tc_kernel   one tile per CTA, two CTAs per SM; takes everything else (point-level rows, ragged tiles).
//
// Operand paths
//   W (weights)  : a TF32-rounded copy tiled on the host exactly as the B operand sits in shared memory
//                  (slide_program.h, GEMM_WP_W), so one cp.async.bulk per (N tile, K block) lands a whole
//                  SWIZZLE_128B tile and signals the stage's mbarrier with its byte count -- no thread touches it.
//   A (activations): NOT stored in the form the tensor core consumes: the producing layer's GroupNorm + ReLU +
//                  timestep/condition vector are applied WHILE loading (this is what removes the separate
//                  normalisation pass over HBM).
//                  - no transform needed (grouped inputs, residual-summed tensors: q/k/v projections, first convs,
//                    residual branches): one cp.async.bulk.tensor.2d per stage (TFLOAT32 tensor map: the copy engine
//                    rounds fp32 to TF32 and writes the SWIZZLE_128B layout);
//                  - persistent kernel, transform needed: the same tensor copy with a plain fp32 map, then the
//                    transform warps rewrite the tile in place (shared -> registers -> shared, cvt.rna.tf32);
//                  - one-tile kernel, transform needed: eight producer warps move A global -> registers (transform)
//                    -> shared memory in the K-major SWIZZLE_128B layout, two K blocks of loads in flight;
//                  - point-level operands with K > 128 (a 128-row tile spans 8 samples): normalised once by
//                    xf_prepass_kernel into scratch, then TMA-fed.
//   D            : one elected thread issues tcgen05.mma (M=128, N=BN, K=8 per instruction); the accumulator lives
//                  in TMEM and is drained with tcgen05.ld (32 lanes x 32 columns per warp), transposed through shared
//                  memory so that every global store / residual load is a coalesced 128-byte row segment.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "program.cuh"

namespace slide {

constexpr int TBM = 128;  // rows per CTA tile (UMMA M)
constexpr int TBK = 32;   // fp32 elements per K block = one 128-byte swizzle row
constexpr int TC_PWARPS = 8;                    // producer / epilogue warps
constexpr int TC_PRODUCERS = TC_PWARPS * 32;    // 256 threads: 8 chunks x 32 row slots, 4 rows each
constexpr int TC_THREADS = TC_PRODUCERS + 32;   // + warp 8: TMEM allocation and MMA issue
constexpr int TC_TABLE_BUDGET = 40 * 1024;  // bytes of shared memory for the A transform table
constexpr int TC_MAX_DYN_SMEM = 227 * 1024;

__device__ int g_tc_error = 0;  // set when a barrier wait times out (never in a correct run)

// Phase stamps of every CTA (tuning builds only: -DTC_TIMELINE, see tools/tc_timeline.py).  Slot layout per CTA:
// 0 start | 1 prologue done | 2 first A loads issued | 3 first stage published | 4 last stage published |
// 5 first stage seen by the MMA thread | 6 last commit issued | 7 accumulator ready | 8 resid table done |
// 9 chunks drained | 10 after the closing barrier | 11 end | 12 globaltimer at start | 13 SM id |
// 14 first chunk read from tensor memory (warp 0) | 15 first chunk done (warp 0)
#ifdef TC_TIMELINE
constexpr int TL_SLOTS = 16, TL_CTAS = 8192;
__device__ unsigned long long g_tc_tl[TL_SLOTS * TL_CTAS];
__device__ __forceinline__ void tl_stamp(int slot) {
  const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
  if (cta < TL_CTAS) g_tc_tl[cta * TL_SLOTS + slot] = (unsigned long long)clock64();
}
#define TL(slot, cond) do { if (cond) tl_stamp(slot); } while (0)
// persistent kernel: per-role cycle accumulators (slot layout in tools/tc_timeline.py)
#define TLP_DECL long long tlp_t0 = 0, tlp_acc[4] = {0, 0, 0, 0}
#define TLP_BEGIN() (tlp_t0 = clock64())
#define TLP_END(i) do { const long long now_ = clock64(); tlp_acc[i] += now_ - tlp_t0; tlp_t0 = now_; } while (0)
#define TLP_FLUSH(slot, i, cond) do { if ((cond) && blockIdx.x < TL_CTAS) g_tc_tl[blockIdx.x * TL_SLOTS + (slot)] = (unsigned long long)tlp_acc[i]; } while (0)
#else
#define TL(slot, cond) do { } while (0)
#define TLP_DECL
#define TLP_BEGIN()
#define TLP_END(i)
#define TLP_FLUSH(slot, i, cond)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU.  After ~2 s (or once any thread has given up) the wait
// returns false and the kernel drains without using tensor-memory results.
__device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((it & 255u) == 0u) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (*(volatile int *)&g_tc_error != 0 || t1 - t0 > 2000000000ull) {
        atomicExch(&g_tc_error, 1);
        return false;
      }
    }
  }
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  return mbar_wait_slow(bar, parity);
}

// x / d for x >= 0, d > 0.  The divisors on the epilogue's per-chunk path (neighbours per point, rows per sample,
// channels per GroupNorm group) are powers of two in every shipped configuration: a shift instead of the ~20-instruction
// division sequence whose I2F / MUFU.RCP / F2I go through the quarter-rate XU pipe (ncu: XU pipe 60 % busy in the
// epilogue-bound kernels).  d is warp-uniform, so the branch does not diverge.
__device__ __forceinline__ int idiv(int x, int d) { return (d & (d - 1)) == 0 ? x >> (31 - __clz(d)) : x / d; }

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// leading byte offset 1 (unused for swizzled K-major), stride byte offset 1024 B (8 rows x 128 B), version 1,
// layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, TF32 x TF32, both K-major, M=128, N=bn.
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
}

// Shared-memory carve-up (bytes from the 1024-aligned base): operand stages | A transform table | control.
// After the main loop the stage area is reused for the epilogue's transpose buffers and resid transform table.
__host__ __device__ constexpr int tc_stage_bytes(int BN) { return (TBM + BN) * TBK * 4; }
__host__ __device__ constexpr int tc_epi_bytes(int BN) { return TC_PWARPS * 32 * 33 * 4 + XF_MAXS * BN * 16; }
__host__ __device__ constexpr int tc_stages_bytes(int BN, int STAGES) {
  return STAGES * tc_stage_bytes(BN) > tc_epi_bytes(BN) ? STAGES * tc_stage_bytes(BN) : tc_epi_bytes(BN);
}
constexpr int TC_CTRL_BYTES = 256 + XF_MAXS * XF_MAXG * 2 * 4 + XF_MAXS * XF_MAXG * 8;  // barriers | stats | (mean, rstd)

// Fill a transform table: entry (sl, k) = (scale, shift, add, 0) so that y = relu?(x*scale + shift) + add.
// Two phases so that the fp64 part runs once per (sample, group), not once per column: `mr` (>= ns * XF_MAXG
// float2, shared) receives (mean, rstd); the caller's CTA must call this with all `nthreads` threads.
__device__ __forceinline__ void fill_xf_table(const XFd &x, float4 *tab, float2 *mr, int s0, int ns, int ncols,
                                              int col0, int stride, int step, int tid, int nthreads) {
  const int G = x.stats ? x.nnorm / x.cg : 1;
  if (x.stats) {
    for (int e = tid; e < ns * G; e += nthreads) {
      const int sl = e / G, g = e - sl * G;
      const double *st = x.stats + ((size_t)(s0 + sl) * G + g) * 2;
      const double m = st[0] * (double)x.inv_count;
      double var = st[1] * (double)x.inv_count - m * m;
      var = var < 0.0 ? 0.0 : var;
      mr[sl * XF_MAXG + g] = make_float2((float)m, (float)(1.0 / sqrt(var + (double)SLIDE_GN_EPS)));
    }
    __syncthreads();
  }
  for (int e = tid; e < ns * ncols; e += nthreads) {
    const int sl = e / ncols, kk = e - sl * ncols;
    const int k = col0 + kk;
    float scale = 1.f, shift = 0.f, add = 0.f;
    if (x.stats) {
      const int ch = x.choff + k;
      if (ch < x.nnorm) {
        const float2 v = mr[sl * XF_MAXG + ch / x.cg];
        scale = v.y * __ldg(x.gamma + ch);
        shift = __ldg(x.beta + ch) - v.x * scale;
      }
    }
    if (x.addvec) {
      const long long arow = x.addmode == 0 ? (s0 + sl) : (x.addmode == 1 ? step : 0);
      add = __ldg(x.addvec + arow * x.addld + k);
    }
    tab[sl * stride + kk] = make_float4(scale, shift, add, 0.f);
  }
}

// Fused AttentionModule tail for one 32-row x 32-column chunk with a compile-time neighbour count.  x = the 32
// value rows of the chunk, fetched by the caller with 32 independent loads issued BEFORE the accumulator chunk is
// read from tensor memory (the dependent per-neighbour loads of the generic loop left ~0.5 KB per warp in flight
// and made the epilogue latency-bound).
__device__ __forceinline__ void load_rows32(float (&x)[32], const float *__restrict__ p, int ld, bool on) {
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = on ? __ldcs(p + (size_t)i * ld) : 0.f;
}

template <int KK>
__device__ __forceinline__ void smk_chunk(const float *tbuf, int lane, const float (&x)[32], bool has_xfr, bool xr_relu,
                                          const float4 *tabR, const int (&xr_off)[4], int ccol,
                                          float *__restrict__ outp, int ldc) {
#pragma unroll
  for (int g0 = 0; g0 < 32; g0 += KK) {
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < KK; ++k) mx = fmaxf(mx, tbuf[(g0 + k) * 33 + lane]);
    float4 c = make_float4(1.f, 0.f, 0.f, 0.f);
    if (has_xfr) c = tabR[xr_off[g0 >> 3] + ccol];
    float den = 0.f, o = 0.f;
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      const float w = __expf(tbuf[(g0 + k) * 33 + lane] - mx);
      float xv = x[g0 + k];
      if (has_xfr) {
        xv = fmaf(xv, c.x, c.y);
        if (xr_relu) xv = fmaxf(xv, 0.f);
        xv += c.z;
      }
      den += w;
      o = fmaf(xv, w, o);
    }
    outp[(size_t)(g0 / KK) * ldc] = o / den;
  }
}

#ifndef TC_MAXNREG
#define TC_MAXNREG 96  // measured on B200: 96 keeps two 288-thread CTAs per SM (112 and 128 run ~25% slower)
#endif
template <int BN, int STAGES, bool SMK, bool TMA_A>
__global__ void __maxnreg__(TC_MAXNREG) gemm_tc_kernel(GemmArgs a, const float *__restrict__ Wp, int wp_na,
                                                       int table_stride, int table_rows, int dbg,
                                                       const __grid_constant__ CUtensorMap tmA) {
  // dbg (SLIDE_TC_DEBUG, profiling only -- results are wrong when set): 1 = no W copies, 2 = no A stores,
  // 4 = no epilogue, 8 = no MMA
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment (pointer arithmetic on the __shared__ array keeps the
  // address space known to the compiler: LDS/STS instead of generic accesses)
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int STAGE_BYTES = tc_stage_bytes(BN);
  constexpr int W_BYTES = BN * TBK * 4;
  const bool has_xfa = a.xfa.stats != nullptr || a.xfa.addvec != nullptr || a.xfa.relu != 0;
  float4 *tabA = reinterpret_cast<float4 *>(smem + tc_stages_bytes(BN, STAGES));
  const int tabA_bytes = has_xfa ? table_rows * table_stride * 16 : 0;
  uint8_t *ctrl = smem + tc_stages_bytes(BN, STAGES) + tabA_bytes;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(ctrl);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *accum_bar = empty_bar + STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);
  float *stacc = reinterpret_cast<float *>(ctrl + 256);
  float2 *mrbuf = reinterpret_cast<float2 *>(ctrl + 256 + XF_MAXS * XF_MAXG * 2 * 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * BN;  // N tiles of one row tile are adjacent: A stays in L2
  TL(0, tid == 0);
#ifdef TC_TIMELINE
  if (tid == 0) {
    unsigned long long gt;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (cta < TL_CTAS) {
      g_tc_tl[cta * TL_SLOTS + 12] = gt;
      g_tc_tl[cta * TL_SLOTS + 13] = smid;
    }
  }
#endif
  const int mlast = min(m0 + TBM, a.M) - 1;
  const int num_kb = (a.K + TBK - 1) / TBK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      // A producers + the W bulk copy's expect_tx arrive; with TMA for A only the expect_tx arrive
      mbar_init(smem_u32(full_bar + s), TMA_A ? 1 : TC_PRODUCERS + 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    mbar_init(smem_u32(accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_PWARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Programmatic dependent launch: everything above is on-chip set-up (barriers, tensor-memory allocation) and may
  // overlap the previous kernel's tail; global memory is touched only from here on.
  pdl_wait();
  pdl_trigger();
  const int step = a.step ? *a.step : 0;
  const int sA0 = m0 / a.xfa.R;
  if (has_xfa)
    fill_xf_table(a.xfa, tabA, mrbuf, sA0, mlast / a.xfa.R - sA0 + 1, a.K, 0, table_stride, step, tid, TC_THREADS);
  if (a.st_stats)
    for (int e = tid; e < XF_MAXS * XF_MAXG * 2; e += TC_THREADS) stacc[e] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  TL(1, tid == 0);

  bool ok = true;
  if (TMA_A && warp < TC_PWARPS) {
    // ------------------------------------------------------------------------------------- TMA producer (one thread)
    if (tid == 0) {
      constexpr uint32_t A_BYTES = TBM * TBK * 4;
      const uint64_t tmap = reinterpret_cast<uint64_t>(&tmA);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        if (ok) ok = mbar_wait(smem_u32(empty_bar + s), ph ^ 1u);
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t bar = smem_u32(full_bar + s);
        mbar_arrive_expect_tx(bar, A_BYTES + (uint32_t)W_BYTES);
        // A tile: rows m0.., columns kb*32..; out-of-range rows / columns are zero-filled by the copy engine
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sa),
            "l"(tmap), "r"(kb * TBK), "r"(m0), "r"(bar)
            : "memory");
        const float *src = Wp + ((size_t)kb * wp_na + (size_t)(n0 >> 3)) * 256;
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                sa + TBM * TBK * 4),
            "l"(src), "r"((uint32_t)W_BYTES), "r"(bar)
            : "memory");
      }
    }
    __syncwarp();
  } else if (warp < TC_PWARPS) {
    // ------------------------------------------------------------------------------------- A producers
    const int chunk = tid & 7;   // 16-byte chunk within the 128-byte K row
    const int rbase = tid >> 3;  // 0..31; this thread owns rows rbase + 32 i, i = 0..3
    const float *arow[4];
    int trow[4];  // offset of the row's sample in the transform table
    bool rvalid[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + rbase + 32 * i;
      rvalid[i] = m < a.M;
      arow[i] = a.A + (size_t)(rvalid[i] ? m : 0) * a.lda + chunk * 4;
      trow[i] = has_xfa && rvalid[i] ? (m / a.xfa.R - sA0) * table_stride + chunk * 4 : 0;
    }
    const bool relu = a.xfa.relu != 0;
    // global loads run two K blocks ahead of the shared-memory stores (two register buffers)
    auto load = [&](float4(&buf)[4], int kb) {
      const int k = kb * TBK + chunk * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kb < num_kb && rvalid[i] && k < a.K) buf[i] = *reinterpret_cast<const float4 *>(arow[i] + (size_t)kb * TBK);
      }
    };
    auto process = [&](float4(&buf)[4], int kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      if (ok) ok = mbar_wait(smem_u32(empty_bar + s), ph ^ 1u);
      uint8_t *sa = smem + (size_t)s * STAGE_BYTES;
      if (tid == 0) {
        const uint32_t bar = smem_u32(full_bar + s);
        if (dbg & 1) {
          mbar_arrive(bar);
        } else {
          // weights: one bulk copy of the pre-tiled, pre-swizzled (N tile, K block); completes on full_bar[s]
          mbar_arrive_expect_tx(bar, (uint32_t)W_BYTES);
          const float *src = Wp + ((size_t)kb * wp_na + (size_t)(n0 >> 3)) * 256;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  smem_u32(sa + TBM * TBK * 4)),
              "l"(src), "r"((uint32_t)W_BYTES), "r"(bar)
              : "memory");
        }
      }
      const int k = kb * TBK + chunk * 4;
      float4 tc[4];
      if (has_xfa && table_rows == 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) tc[u] = tabA[k + u];  // the table is padded to a multiple of TBK entries
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rbase + 32 * i;
        float v[4] = {buf[i].x, buf[i].y, buf[i].z, buf[i].w};
        if (has_xfa && table_rows == 1) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float y = fmaf(v[u], tc[u].x, tc[u].y);
            if (relu) y = fmaxf(y, 0.f);
            v[u] = y + tc[u].z;
          }
        } else if (has_xfa && rvalid[i]) {
          const float4 *t = tabA + trow[i] + kb * TBK;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (k + u < a.K) {
              const float4 c = t[u];
              float y = fmaf(v[u], c.x, c.y);
              if (relu) y = fmaxf(y, 0.f);
              v[u] = y + c.z;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (k + u >= a.K) v[u] = 0.f;
        const uint4 o = make_uint4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
        if (!(dbg & 2)) *reinterpret_cast<uint4 *>(sa + r * 128 + ((chunk ^ (r & 7)) << 4)) = o;
      }
    };
    auto publish = [&](int kb) {
      // make the generic-proxy writes visible to the tensor core (async proxy), then signal
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(smem_u32(full_bar + (kb % STAGES)));
    };
    float4 bufA[4], bufB[4];
    load(bufA, 0);
    load(bufB, 1);
    TL(2, tid == 0);
    for (int kb = 0; kb < num_kb; kb += 2) {
      process(bufA, kb);
      load(bufA, kb + 2);
      publish(kb);
      TL(3, tid == 0 && kb == 0);
      if (kb + 1 < num_kb) {
        process(bufB, kb + 1);
        load(bufB, kb + 3);
        publish(kb + 1);
      }
    }
    TL(4, tid == 0);
  } else {
    // ------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        if (ok) ok = mbar_wait(smem_u32(full_bar + s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        TL(5, kb == 0);
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint64_t da = make_smem_desc(sa);
        const uint64_t db = make_smem_desc(sa + TBM * TBK * 4);
        if (ok && !(dbg & 8)) {
#pragma unroll
          for (int kk = 0; kk < TBK / 8; ++kk) {
            const uint32_t accum = (kb > 0 || kk > 0) ? 1u : 0u;
            // advance 8 TF32 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
                "}\n" ::"r"(tmem_base),
                "l"(da + (uint64_t)(kk * 2)), "l"(db + (uint64_t)(kk * 2)), "r"(idesc), "r"(accum)
                : "memory");
          }
        }
        // release the stage to the producers once the MMAs that read it have completed
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(empty_bar + s))
                     : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(accum_bar))
                   : "memory");
      TL(6, true);
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------- epilogue
  // bias and the ev rows of every chunk this warp will drain are requested BEFORE the wait for the accumulator:
  // inside the chunk loop each of them would cost a serialised L2 round trip (measured: ~1000 cycles apiece)
  // (the chunk loop stays rolled -- unrolling it made the kernel instruction-fetch bound -- so the values for chunk
  // i+1 are requested at the top of chunk i and the first chunk's before the wait)
  float bias_nx = 0.f, ev_nx[4] = {0.f, 0.f, 0.f, 0.f};
  auto prefetch_vec = [&](int c0_) {
    const int mw_ = m0 + (warp & 3) * 32;
    const int n_ = n0 + c0_ + lane;
    const bool on_ = warp < TC_PWARPS && c0_ < BN && n_ < a.N;
    bias_nx = (on_ && a.bias) ? __ldg(a.bias + n_) : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      ev_nx[q] = (!SMK && on_ && a.ev && mw_ + 32 <= a.M) ? __ldg(a.ev + (size_t)idiv(mw_ + 8 * q, a.evdiv) * a.evld + n_)
                                                         : 0.f;
  };
  prefetch_vec(32 * (warp >> 2));
  if (warp < TC_PWARPS) {
    if (ok) ok = mbar_wait(smem_u32(accum_bar), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  TL(7, tid == 0);
  // every stage buffer is dead now (all MMAs completed before accum_bar fired): reuse the area
  float *tbuf = reinterpret_cast<float *>(smem) + (warp < TC_PWARPS ? warp : 0) * (32 * 33);
  float4 *tabR = reinterpret_cast<float4 *>(smem + TC_PWARPS * 32 * 33 * 4);
  const bool has_xfr = a.res && (a.xfr.stats != nullptr || a.xfr.addvec != nullptr || a.xfr.relu != 0);
  const int sR0 = idiv(m0, a.xfr.R);
  // (no CTA barrier without a residual transform: a warp that has seen accum_bar knows every MMA -- every reader of the
  // stage area -- has completed, and its staging tile is private; the barrier cost ~800 cycles of wake-up skew per record)
  if (has_xfr) {
    __syncthreads();
    const int ncols = min(BN, a.N - n0);
    fill_xf_table(a.xfr, tabR, mrbuf, sR0, mlast / a.xfr.R - sR0 + 1, ncols, n0, BN, step, tid, TC_THREADS);
    __syncthreads();  // has_xfr is uniform over the CTA
  }
  TL(8, tid == 0);

  if (warp < TC_PWARPS) {
    // warp w drains TMEM lanes 32 (w & 3) .. +31 (the hardware's lane window of warp w % 4); the two warps that share
    // a lane window split the 32-column chunks: even chunks to warps 0-3, odd chunks to warps 4-7
    const int lq = warp & 3, cpar = warp >> 2;
    const int mw = m0 + lq * 32;           // first row of this warp
    const int rr_end = min(32, a.M - mw);  // rows of this warp that exist (warp-uniform, may be <= 0)
    const int sS0 = m0 / a.st_R;
    const bool xr_relu = a.xfr.relu != 0;
    // Fast path: blocks of 8 rows never straddle an ev row / resid sample / statistics sample.
    const bool fast = rr_end == 32 && (!a.ev || a.evdiv % 8 == 0) && (!has_xfr || a.xfr.R % 8 == 0) &&
                      (!a.st_stats || a.st_R % 8 == 0);
    // GroupNorm groups are st_cg consecutive channels.  When st_cg is a power of two and the group grid is aligned
    // with this warp's 32-column chunks, the lanes of a group are reduced with shuffles before the shared atomics.
    const int st_seg = a.st_cg < 32 ? a.st_cg : 32;
    const bool st_pow2 = a.st_stats && (a.st_cg & (a.st_cg - 1)) == 0 && ((a.st_choff + n0) % st_seg) == 0 &&
                         (a.st_nnorm % st_seg) == 0;
    // per 8-row block: ev row, resid-table row, statistics row (hoisted: no division inside the loops)
    int ev_off[4], xr_off[4], st_off[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int mb = mw + 8 * q;
      ev_off[q] = a.ev ? idiv(mb, a.evdiv) * a.evld : 0;
      xr_off[q] = has_xfr ? (idiv(mb, a.xfr.R) - sR0) * BN : 0;
      st_off[q] = a.st_stats ? (idiv(mb, a.st_R) - sS0) * XF_MAXG * 2 : 0;
    }
    for (int c0 = 32 * cpar; c0 < BN; c0 += 64) {
      if (n0 + c0 >= a.N || (dbg & 4)) break;
      const float bias = bias_nx;
      const float evc[4] = {ev_nx[0], ev_nx[1], ev_nx[2], ev_nx[3]};
      prefetch_vec(c0 + 64);
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0;
      if (ok) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
              "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
              "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      TL(14, tid == 0 && c0 == 0);
      // transpose: thread (= row) writes its 32 columns; afterwards lane = column
#pragma unroll
      for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = __uint_as_float(r[j]);
      __syncwarp();
      const int n = n0 + c0 + lane;
      const bool ncol = n < a.N;
      const int stch = a.st_choff + n;
      const bool dost = a.st_stats && ncol && stch < a.st_nnorm;
      const int stg = dost ? idiv(stch, a.st_cg) : 0;
      if (SMK) {
        // fused AttentionModule tail: soft-max down each group of smk rows (the neighbours of one point), applied
        // to the transformed value rows; one output row per group.  (bias is constant down a column: it cancels.)
        const int K = a.smk;
        const bool smk_fast = rr_end == 32 && (K == 4 || K == 8 || K == 16 || K == 32);  // warp-uniform
        if (ncol && smk_fast) {
          float x[32];
          load_rows32(x, a.res + (size_t)mw * a.ldr + n, a.ldr, true);
          float *op = a.C + (size_t)idiv(mw, K) * a.ldc + n;
          if (K == 16) smk_chunk<16>(tbuf, lane, x, has_xfr, xr_relu, tabR, xr_off, c0 + lane, op, a.ldc);
          else if (K == 8) smk_chunk<8>(tbuf, lane, x, has_xfr, xr_relu, tabR, xr_off, c0 + lane, op, a.ldc);
          else if (K == 32) smk_chunk<32>(tbuf, lane, x, has_xfr, xr_relu, tabR, xr_off, c0 + lane, op, a.ldc);
          else smk_chunk<4>(tbuf, lane, x, has_xfr, xr_relu, tabR, xr_off, c0 + lane, op, a.ldc);
        } else if (ncol) {
          for (int g0 = 0; g0 < 32; g0 += K) {
            float mx = -INFINITY;
            for (int k = 0; k < K; ++k) mx = fmaxf(mx, tbuf[(g0 + k) * 33 + lane]);
            const float *vp = a.res + (size_t)(mw + g0) * a.ldr + n;
            float4 c = make_float4(1.f, 0.f, 0.f, 0.f);
            if (has_xfr) c = tabR[xr_off[g0 >> 3] + c0 + lane];
            float den = 0.f, o = 0.f;
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
              const float w = __expf(tbuf[(g0 + k) * 33 + lane] - mx);
              float x = vp[(size_t)k * a.ldr];
              if (has_xfr) {
                x = fmaf(x, c.x, c.y);
                if (xr_relu) x = fmaxf(x, 0.f);
                x += c.z;
              }
              den += w;
              o = fmaf(x, w, o);
            }
            a.C[(size_t)((mw + g0) / K) * a.ldc + n] = o / den;
          }
        }
      } else if (!SMK && fast) {
        float ssum = 0.f, ssq = 0.f;
        {
          // all 32 residual rows of the chunk are requested up front (32 independent 128-byte row segments per warp)
          float xres[32];
          if (a.res) {
            const float *rp0 = a.res + (size_t)mw * a.ldr + n;
#pragma unroll
            for (int i = 0; i < 32; ++i) xres[i] = ncol ? __ldcs(rp0 + (size_t)i * a.ldr) : 0.f;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int rb = 8 * q;
            const int mb = mw + rb;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = tbuf[(rb + i) * 33 + lane] + bias;
            if (a.ev) {
              const float e = evc[q];  // fast path: the warp's 32 rows exist, 8-row blocks lie inside one ev row
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += e;
            }
            if (a.res) {
              float x[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = xres[rb + i];
              if (has_xfr) {
                const float4 c = tabR[xr_off[q] + c0 + lane];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float y = fmaf(x[i], c.x, c.y);
                  if (xr_relu) y = fmaxf(y, 0.f);
                  x[i] = y + c.z;
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += x[i];
            }
            if (a.act == 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (a.act == 2) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = act_apply(2, v[i]);
            }
            if (ncol) {
              float *cp = a.C + (size_t)mb * a.ldc + n;
#pragma unroll
              for (int i = 0; i < 8; ++i) cp[(size_t)i * a.ldc] = v[i];
            }
            if (a.st_stats) {  // warp-uniform
              if (dost) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  ssum += v[i];
                  ssq = fmaf(v[i], v[i], ssq);
                }
              }
              // flush when the next block belongs to another sample (or this is the last block)
              if (q == 3 || st_off[q + 1] != st_off[q]) {
                float rs = ssum, rq = ssq;
                bool leader = dost;
                if (st_pow2) {
                  // lanes of one GroupNorm group are st_cg consecutive columns (aligned): segmented butterfly sum
                  for (int d = 1; d < st_seg; d <<= 1) {
                    rs += __shfl_xor_sync(0xffffffffu, rs, d);
                    rq += __shfl_xor_sync(0xffffffffu, rq, d);
                  }
                  leader = dost && (lane & (st_seg - 1)) == 0;
                }
                if (leader) {
                  float *slot = stacc + st_off[q] + stg * 2;
                  atomicAdd(slot, rs);
                  atomicAdd(slot + 1, rq);
                }
                ssum = 0.f;
                ssq = 0.f;
              }
            }
          }
        }
      } else if (!SMK) {
        // general path (ragged tiles / odd sample sizes): one row at a time
        for (int rr = 0; rr < rr_end; ++rr) {
          const int m = mw + rr;
          if (ncol) {
            float v = tbuf[rr * 33 + lane] + bias;
            if (a.ev) v += a.ev[(size_t)(m / a.evdiv) * a.evld + n];
            if (a.res) {
              float x = a.res[(size_t)m * a.ldr + n];
              if (has_xfr) {
                const float4 c = tabR[(m / a.xfr.R - sR0) * BN + c0 + lane];
                x = fmaf(x, c.x, c.y);
                if (xr_relu) x = fmaxf(x, 0.f);
                x += c.z;
              }
              v += x;
            }
            v = act_apply(a.act, v);
            a.C[(size_t)m * a.ldc + n] = v;
            if (dost) {
              float *slot = stacc + ((m / a.st_R - sS0) * XF_MAXG + stg) * 2;
              atomicAdd(slot, v);
              atomicAdd(slot + 1, v * v);
            }
          }
        }
      }
      __syncwarp();
      TL(15, tid == 0 && c0 == 0);
    }
  }
  TL(9, tid == 0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  TL(10, tid == 0);
  if (a.st_stats) {
    const int G = a.st_nnorm / a.st_cg;
    const int sS0 = m0 / a.st_R;
    const int ns = mlast / a.st_R - sS0 + 1;
    for (int e = tid; e < ns * G * 2; e += TC_THREADS) {
      const int sl = e / (G * 2), rem = e - sl * G * 2;
      const float v = stacc[(sl * XF_MAXG + (rem >> 1)) * 2 + (rem & 1)];
      if (v != 0.f) atomicAdd(a.st_stats + ((size_t)(sS0 + sl) * G) * 2 + rem, (double)v * (double)a.st_weight);
    }
  }
  if (warp == TC_PWARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
  }
  TL(11, tid == 0);
}

// =====================================================================================================
// Persistent, warp-specialised variant: one CTA per SM loops over output tiles; an operand ring and a
// DOUBLE-BUFFERED accumulator in tensor memory let the loads and MMAs of tile i+1 run while eight epilogue warps
// drain tile i, so the L2->SM operand stream, the tensor pipe and the HBM write burst of the epilogue overlap
// (the one-tile-per-CTA kernel above runs them one after another, and its co-resident CTAs move in lock step).
//   warp 0     : bulk-copy producer (W per K block; with !XFA also the A tensor copy)
//   warp 1     : TMEM allocation (2 x BN columns) + tcgen05.mma issue
//   warps 2-9  : epilogue (TMEM -> registers -> smem transpose -> coalesced stores, statistics; SMK: fused soft-max)
//   warps 10-17: (XFA only) A producers: global -> registers (GroupNorm / ReLU / add, TF32 rounding) -> swizzled smem
// Used when every row tile lies inside one sample (R % 128 == 0) and M % 128 == 0.
// =====================================================================================================
constexpr int TCP_EPI_WARPS = 8;
constexpr int TCP_THREADS = 32 * (2 + TCP_EPI_WARPS);
constexpr int TCP_XFA_THREADS = TCP_THREADS + TC_PRODUCERS;

__host__ __device__ constexpr int tcp_smem_bytes(int BN, int STAGES) {
  // operand ring | per-warp transpose tiles | resid table x2 | barriers | per-lane-window column sums [4][BN] float2
  return STAGES * tc_stage_bytes(BN) + TCP_EPI_WARPS * 32 * 33 * 4 + 2 * BN * 16 + 512 + 4 * BN * 8;
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TCP_EPI_WARPS * 32) : "memory"); }
__device__ __forceinline__ void prod_bar_sync() { asm volatile("bar.sync 2, %0;" ::"n"(TC_PRODUCERS) : "memory"); }

template <int BN, int STAGES, bool XFA, bool SMK>
__global__ void __launch_bounds__(XFA ? TCP_XFA_THREADS : TCP_THREADS, 1)
    gemm_tcp_kernel(GemmArgs a, const float *__restrict__ Wp, int wp_na, int table_stride,
                    const __grid_constant__ CUtensorMap tmA) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int STAGE_BYTES = tc_stage_bytes(BN);
  constexpr int W_BYTES = BN * TBK * 4;
  constexpr uint32_t A_BYTES = TBM * TBK * 4;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // BN in {32,...,256}: a power of two
  float *tbuf_all = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES);
  float4 *tabR_all = reinterpret_cast<float4 *>(smem + STAGES * STAGE_BYTES + TCP_EPI_WARPS * 32 * 33 * 4);
  uint8_t *ctrl = reinterpret_cast<uint8_t *>(tabR_all) + 2 * BN * 16;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(ctrl);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *raw_bar = empty_bar + STAGES;  // XFA: the untransformed A tile (+ W) of a stage has landed
  uint64_t *tfull_bar = raw_bar + STAGES;
  uint64_t *tempty_bar = tfull_bar + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);
  float2 *spart = reinterpret_cast<float2 *>(ctrl + 512);  // (sum, sum of squares) per (lane window, tile column)
  float2 *mrA = reinterpret_cast<float2 *>(ctrl + 512 + 4 * BN * 8);  // XFA: (mean, rstd) per group
  float4 *tabA = reinterpret_cast<float4 *>(ctrl + 512 + 4 * BN * 8 + XF_MAXG * 8);  // XFA: per K column

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int num_kb = (a.K + TBK - 1) / TBK;
  const int tiles_n = (a.N + BN - 1) / BN;
  const int tiles = (a.M / TBM) * tiles_n;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      // XFA: the 256 transform threads arrive; otherwise only the copy engine's expect_tx arrive
      mbar_init(smem_u32(full_bar + s), XFA ? TC_PRODUCERS : 1);
      mbar_init(smem_u32(raw_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(tfull_bar + b), 1);
      mbar_init(smem_u32(tempty_bar + b), TCP_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: barriers and tensor memory are set up while the previous kernel drains; global
  // memory (operands, statistics, the step counter) is touched only after this point.
  pdl_wait();
  pdl_trigger();
  const int step = a.step ? *a.step : 0;
  TL(0, tid == 0);
#ifdef TC_TIMELINE
  if (tid == 0 && blockIdx.x < TL_CTAS) {
    unsigned long long gt;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_tc_tl[blockIdx.x * TL_SLOTS + 12] = gt;
    g_tc_tl[blockIdx.x * TL_SLOTS + 13] = smid;
    g_tc_tl[blockIdx.x * TL_SLOTS + 10] = (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  }
#endif
  TLP_DECL;

  bool ok = true;
  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint64_t tmap = reinterpret_cast<uint64_t>(&tmA);
      uint32_t it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * TBM, n0 = (t % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          TLP_BEGIN();
          if (ok) ok = mbar_wait(smem_u32(empty_bar + s), ph ^ 1u);
          TLP_END(0);
          const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
          // XFA: the tensor map is plain fp32 and the tile lands on raw_bar: the transform warps rewrite it in place
          const uint32_t bar = smem_u32((XFA ? raw_bar : full_bar) + s);
          mbar_arrive_expect_tx(bar, A_BYTES + (uint32_t)W_BYTES);
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sa),
              "l"(tmap), "r"(kb * TBK), "r"(m0), "r"(bar)
              : "memory");
          const float *src = Wp + ((size_t)kb * wp_na + (size_t)(n0 >> 3)) * 256;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  sa + TBM * TBK * 4),
              "l"(src), "r"((uint32_t)W_BYTES), "r"(bar)
              : "memory");
        }
      }
      TLP_FLUSH(9, 0, true);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BN);
      uint32_t it = 0, tc = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++tc) {
        const uint32_t buf = tc & 1u, tph = (tc >> 1) & 1u;
        TLP_BEGIN();
        if (ok) ok = mbar_wait(smem_u32(tempty_bar + buf), tph ^ 1u);  // epilogue has drained this accumulator
        TLP_END(1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          TLP_BEGIN();
          if (ok) ok = mbar_wait(smem_u32(full_bar + s), ph);
          TLP_END(0);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint64_t da = make_smem_desc(sa);
          const uint64_t db = make_smem_desc(sa + TBM * TBK * 4);
          if (ok) {
#pragma unroll
            for (int kk = 0; kk < TBK / 8; ++kk) {
              const uint32_t accum = (kb > 0 || kk > 0) ? 1u : 0u;
              asm volatile(
                  "{\n"
                  ".reg .pred p;\n"
                  "setp.ne.b32 p, %4, 0;\n"
                  "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
                  "}\n" ::"r"(tacc),
                  "l"(da + (uint64_t)(kk * 2)), "l"(db + (uint64_t)(kk * 2)), "r"(idesc), "r"(accum)
                  : "memory");
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_u32(empty_bar + s))
                       : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(tfull_bar + buf))
                     : "memory");
      }
      TLP_FLUSH(1, 0, true);
      TLP_FLUSH(2, 1, true);
    }
    __syncwarp();
  } else if (XFA && warp >= 2 + TCP_EPI_WARPS) {
    // ------------------------------------------------------------------------------------------ A transform warps
    // The copy engine has written the raw fp32 tile in the SWIZZLE_128B layout; each thread rewrites 4 of its
    // 16-byte chunks in place: y = relu?(x * scale + shift) + add, rounded to TF32.  No global-memory latency sits
    // on this path: the tensor copies run up to STAGES K blocks ahead.
    const int ptid = tid - TCP_THREADS;  // 0..255
    const int pchunk = ptid & 7;         // physical 16-byte chunk within the 128-byte row
    const int rbase = ptid >> 3;         // 0..31; this thread owns rows rbase + 32 i, i = 0..3
    const int lchunk = pchunk ^ (rbase & 7);  // logical K chunk stored there ((r & 7) is the same for the 4 rows)
    const bool relu = a.xfa.relu != 0;
    const int G = a.xfa.stats ? a.xfa.nnorm / a.xfa.cg : 0;
    uint32_t it = 0;
    int cur_m0 = -1;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int m0 = (t / tiles_n) * TBM;
      if (m0 != cur_m0) {
        // (scale, shift, add) per K column for this tile's sample; fp64 only once per group.  Columns >= K get
        // (0, 0, 0): the copy engine zero-fills them and they must stay zero.
        cur_m0 = m0;
        const int sA = m0 / a.xfa.R;
        prod_bar_sync();  // everyone has finished reading the previous table
        if (ptid < G) {
          const double *st = a.xfa.stats + ((size_t)sA * G + ptid) * 2;
          const double m = st[0] * (double)a.xfa.inv_count;
          double var = st[1] * (double)a.xfa.inv_count - m * m;
          var = var < 0.0 ? 0.0 : var;
          mrA[ptid] = make_float2((float)m, (float)(1.0 / sqrt(var + (double)SLIDE_GN_EPS)));
        }
        prod_bar_sync();
        for (int k = ptid; k < table_stride; k += TC_PRODUCERS) {
          float scale = 0.f, shift = 0.f, add = 0.f;
          if (k < a.K) {
            scale = 1.f;
            if (a.xfa.stats) {
              const int ch = a.xfa.choff + k;
              if (ch < a.xfa.nnorm) {
                const float2 v = mrA[ch / a.xfa.cg];
                scale = v.y * __ldg(a.xfa.gamma + ch);
                shift = __ldg(a.xfa.beta + ch) - v.x * scale;
              }
            }
            if (a.xfa.addvec) {
              const long long arow_ = a.xfa.addmode == 0 ? sA : (a.xfa.addmode == 1 ? step : 0);
              add = __ldg(a.xfa.addvec + arow_ * a.xfa.addld + k);
            }
          }
          tabA[k] = make_float4(scale, shift, add, 0.f);
        }
        prod_bar_sync();
      }
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        float4 tc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) tc[u] = tabA[kb * TBK + lchunk * 4 + u];
        TLP_BEGIN();
        if (ok) ok = mbar_wait(smem_u32(raw_bar + s), ph);
        TLP_END(0);
        uint8_t *sa = smem + (size_t)s * STAGE_BYTES + rbase * 128 + (pchunk << 4);
        float4 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4 *>(sa + i * (32 * 128));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float y = fmaf(v[u], tc[u].x, tc[u].y);
            if (relu) y = fmaxf(y, 0.f);
            v[u] = y + tc[u].z;
          }
          *reinterpret_cast<uint4 *>(sa + i * (32 * 128)) =
              make_uint4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
        }
        // make the generic-proxy writes visible to the tensor core (async proxy), then signal
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(smem_u32(full_bar + s));
      }
    }
    TLP_FLUSH(8, 0, ptid == 0);
  } else {
    // ------------------------------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;              // 0..7
    const int etid = tid - 64;            // 0..255
    const int lq = warp & 3;              // TMEM lane window of this warp (hardware: warp id % 4)
    const int cpar = ew >> 2;             // which half of the 32-column chunks
    float *tbuf = tbuf_all + ew * (32 * 33);
    const bool has_xfr = a.res && (a.xfr.stats != nullptr || a.xfr.addvec != nullptr || a.xfr.relu != 0);
    const bool xr_relu = a.xfr.relu != 0;
    const int st_seg = a.st_cg < 32 ? a.st_cg : 32;
    uint32_t tc = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++tc) {
      const int m0 = (t / tiles_n) * TBM, n0 = (t % tiles_n) * BN;
      const uint32_t buf = tc & 1u, tph = (tc >> 1) & 1u;
      float4 *tabR = tabR_all + buf * BN;
      // per-tile set-up (runs while the tile's MMAs are still in flight): statistics accumulators, and one
      // (scale, shift, add, bias) entry per column: the resid transform of this (sample, column range) + the bias
      if (etid < BN) {
        const int n = n0 + etid;
        float scale = 1.f, shift = 0.f, add = 0.f, bias = 0.f;
        if (n < a.N) {
          if (a.bias) bias = __ldg(a.bias + n);
          if (has_xfr) {
            const int sR = m0 / a.xfr.R;
            if (a.xfr.stats) {
              const int ch = a.xfr.choff + n;
              if (ch < a.xfr.nnorm) {
                const int G = a.xfr.nnorm / a.xfr.cg;
                const double *st = a.xfr.stats + ((size_t)sR * G + ch / a.xfr.cg) * 2;
                const double m = st[0] * (double)a.xfr.inv_count;
                double var = st[1] * (double)a.xfr.inv_count - m * m;
                var = var < 0.0 ? 0.0 : var;
                scale = (float)(1.0 / sqrt(var + (double)SLIDE_GN_EPS)) * __ldg(a.xfr.gamma + ch);
                shift = __ldg(a.xfr.beta + ch) - (float)m * scale;
              }
            }
            if (a.xfr.addvec) {
              const long long arow = a.xfr.addmode == 0 ? sR : (a.xfr.addmode == 1 ? step : 0);
              add = __ldg(a.xfr.addvec + arow * a.xfr.addld + n);
            }
          }
        }
        tabR[etid] = make_float4(scale, shift, add, bias);
      }
      const int mw = m0 + lq * 32;
      // the ev rows (one per 8-row block) of the first chunk are requested now and used after the wait; inside
      // the (rolled) chunk loop the rows of chunk i+1 are requested at the top of chunk i
      float ev_nx[4];
      auto prefetch_ev = [&](int c0_) {
        const int n_ = n0 + c0_ + lane;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          ev_nx[q] = (!SMK && a.ev && c0_ < BN && n_ < a.N) ? __ldg(a.ev + (size_t)idiv(mw + 8 * q, a.evdiv) * a.evld + n_) : 0.f;
      };
      prefetch_ev(32 * cpar);
      epi_bar_sync();
      TLP_BEGIN();
      if (ok) ok = mbar_wait(smem_u32(tfull_bar + buf), tph);
      TLP_END(0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool st_pow2 = a.st_stats && (a.st_cg & (a.st_cg - 1)) == 0 && ((a.st_choff + n0) % st_seg) == 0 &&
                           (a.st_nnorm % st_seg) == 0;
      for (int c0 = 32 * cpar; c0 < BN; c0 += 64) {
        if (n0 + c0 >= a.N) break;
        const int n = n0 + c0 + lane;
        const bool ncol = n < a.N;
        const float evc[4] = {ev_nx[0], ev_nx[1], ev_nx[2], ev_nx[3]};
        prefetch_ev(c0 + 64);
        // the chunk's 32 value / residual rows: in flight while the accumulator chunk is read and transposed
        float xres[32];
        if (SMK || a.res) load_rows32(xres, a.res + (size_t)mw * a.ldr + n, a.ldr, ncol);
        uint32_t r[32];
        const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0;
        if (ok) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                "=r"(r[30]), "=r"(r[31])
              : "r"(taddr)
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) r[jj] = 0u;
        }
        TLP_END(1);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) tbuf[lane * 33 + jj] = __uint_as_float(r[jj]);
        __syncwarp();
        TLP_END(2);
        const float4 c = tabR[c0 + lane];  // (resid scale, shift, add, bias)
        const float bias = c.w;
        const int stch = a.st_choff + n;
        const bool dost = a.st_stats && ncol && stch < a.st_nnorm;
        if (SMK) {
          // fused AttentionModule tail (see smk_chunk); the whole tile is one sample: a single resid-table row
          if (ncol) {
            const int zoff[4] = {0, 0, 0, 0};
            float *op = a.C + (size_t)idiv(mw, a.smk) * a.ldc + n;
            if (a.smk == 16) smk_chunk<16>(tbuf, lane, xres, has_xfr, xr_relu, tabR, zoff, c0 + lane, op, a.ldc);
            else if (a.smk == 8) smk_chunk<8>(tbuf, lane, xres, has_xfr, xr_relu, tabR, zoff, c0 + lane, op, a.ldc);
            else if (a.smk == 32) smk_chunk<32>(tbuf, lane, xres, has_xfr, xr_relu, tabR, zoff, c0 + lane, op, a.ldc);
            else smk_chunk<4>(tbuf, lane, xres, has_xfr, xr_relu, tabR, zoff, c0 + lane, op, a.ldc);
          }
          __syncwarp();
          TLP_END(3);
          continue;
        }
        float ssum = 0.f, ssq = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int mb = mw + 8 * q;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = tbuf[(8 * q + i) * 33 + lane] + bias;
          if (a.ev) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += evc[q];
          }
          if (a.res) {
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = xres[8 * q + i];
            if (has_xfr) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float y = fmaf(x[i], c.x, c.y);
                if (xr_relu) y = fmaxf(y, 0.f);
                x[i] = y + c.z;
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += x[i];
          }
          if (a.act == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (a.act == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = act_apply(2, v[i]);
          }
          if (ncol) {
            float *cp = a.C + (size_t)mb * a.ldc + n;
#pragma unroll
            for (int i = 0; i < 8; ++i) cp[(size_t)i * a.ldc] = v[i];
          }
          if (dost) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              ssum += v[i];
              ssq = fmaf(v[i], v[i], ssq);
            }
          }
        }
        // column sums of this warp's 32 rows: parked per (lane window, column); reduced once per tile below (the
        // per-chunk shuffle tree + shared-memory float atomics -- CAS loops -- cost 17 % of the epilogue's samples)
        if (a.st_stats) spart[lq * BN + c0 + lane] = make_float2(dost ? ssum : 0.f, dost ? ssq : 0.f);
        __syncwarp();
        TLP_END(3);
      }
      // this warp has finished reading the accumulator: hand the TMEM buffer back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tempty_bar + buf));
      if (a.st_stats) {
        epi_bar_sync();  // every warp has parked its column sums
        // thread = tile column: add the four lane windows, reduce the columns of a GroupNorm group with a segmented
        // butterfly, one fp64 atomic pair per (segment, moment)
        const int G = a.st_nnorm / a.st_cg;
        const int ncols = min(BN, a.N - n0);
        float rs = 0.f, rq = 0.f;
        const int stch = a.st_choff + n0 + etid;
        const bool on = etid < ncols && stch < a.st_nnorm;
        if (on) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float2 v = spart[w * BN + etid];
            rs += v.x;
            rq += v.y;
          }
        }
        bool leader = on;
        if (st_pow2) {
          for (int d = 1; d < st_seg; d <<= 1) {
            rs += __shfl_xor_sync(0xffffffffu, rs, d);
            rq += __shfl_xor_sync(0xffffffffu, rq, d);
          }
          leader = on && (lane & (st_seg - 1)) == 0;
        }
        if (leader && etid < BN) {
          double *slot = a.st_stats + ((size_t)(m0 / a.st_R) * G + idiv(stch, a.st_cg)) * 2;
          atomicAdd(slot, (double)rs * (double)a.st_weight);
          atomicAdd(slot + 1, (double)rq * (double)a.st_weight);
        }
        // (the parking area is rewritten only after the next tile's set-up barrier, which every warp reaches after
        // these reads)
      }
    }
  }
  if (warp == 2) {
    TLP_FLUSH(3, 0, lane == 0);
    TLP_FLUSH(4, 1, lane == 0);
    TLP_FLUSH(5, 2, lane == 0);
    TLP_FLUSH(6, 3, lane == 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
  TL(11, tid == 0);
}

// A row tile must map onto whole samples (or lie inside one): then it touches at most XF_MAXS of them.
static bool spans_ok_tc(int R) { return R % TBM == 0 || (TBM % R == 0 && TBM / R <= XF_MAXS); }
static int tc_rows_per_tile(int R) { return R % TBM == 0 ? 1 : TBM / R; }
static int tc_table_stride(const GemmArgs &a) { return ((a.K + TBK - 1) / TBK) * TBK; }
static bool has_xf(const XFd &x) { return x.stats || x.addvec || x.relu; }

#ifndef TC_MIN_MACS
#define TC_MIN_MACS 0  // A/B on B200: the tcgen05 kernel beats the FFMA kernel even for the 4096-row point-level GEMMs
#endif

bool gemm_tc_eligible(const GemmArgs &a, const float *Wp) {
  if (!Wp) return false;
  if (a.M < TBM || a.K < 32 || a.N < 32) return false;
  if ((long long)a.M * a.N * a.K < TC_MIN_MACS) return false;
  if (((uintptr_t)a.A & 15) || (a.lda & 3) || ((uintptr_t)Wp & 15)) return false;
  if (has_xf(a.xfa)) {
    if (!spans_ok_tc(a.xfa.R)) return false;
    if (tc_rows_per_tile(a.xfa.R) * tc_table_stride(a) * 16 > TC_TABLE_BUDGET) return false;
    if (a.xfa.stats && a.xfa.nnorm / a.xfa.cg > XF_MAXG) return false;
  }
  if (a.res && has_xf(a.xfr)) {
    if (!spans_ok_tc(a.xfr.R)) return false;
    if (a.xfr.stats && a.xfr.nnorm / a.xfr.cg > XF_MAXG) return false;
  }
  if (a.st_stats && (!spans_ok_tc(a.st_R) || a.st_nnorm / a.st_cg > XF_MAXG)) return false;
  // fused soft-max: neighbour groups must tile a warp's 32 rows, full row tiles only
  if (a.smk > 0 && (32 % a.smk != 0 || a.M % TBM != 0 || a.xfr.R % 8 != 0)) return false;
  return true;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// A as a 2-D TFLOAT32 tensor [M rows, K columns] with row pitch lda: box = 128 rows x 32 columns, SWIZZLE_128B.
static bool make_a_map(const GemmArgs &a, CUtensorMap *m, bool raw = false) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)a.M};
  const cuuint64_t strides[1] = {(cuuint64_t)a.lda * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)TBM};
  const cuuint32_t estr[2] = {1, 1};
  // raw: plain fp32 (the transform warps round to TF32 themselves); otherwise the copy engine rounds
  return fn(m, raw ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float *>(a.A), dims,
            strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Tuning knobs (A/B runs and the tests that force the persistent kernel onto small problems).  The environment is
// read ONCE, on first use; slide_tc_reload_tuning() re-reads it (tests call it after changing a variable).
struct TcTuning {
  int use_tma, persist, persist_min_tiles, persist_min_k, persist_min_bn, persist_grid, prepass, debug;
};
static TcTuning g_tuning;
static bool g_tuning_loaded = false;
static void load_tuning() {
  g_tuning.use_tma = env_int("SLIDE_TC_TMA", 1);
  // persistent kernel: 0 = never, 1 = TMA-fed operands only, 2 = also with transform producers / fused soft-max
  g_tuning.persist = env_int("SLIDE_TC_PERSIST", 2);
  g_tuning.persist_min_tiles = env_int("SLIDE_TC_PERSIST_MIN_TILES", 148);
  g_tuning.persist_min_k = env_int("SLIDE_TC_PERSIST_MIN_K", 32);
  g_tuning.persist_min_bn = env_int("SLIDE_TC_PERSIST_MIN_BN", 128);
  g_tuning.persist_grid = env_int("SLIDE_TC_PERSIST_GRID", 0);  // 0 = one CTA per SM of the current device
  g_tuning.prepass = env_int("SLIDE_TC_PREPASS", 1);
  g_tuning.debug = env_int("SLIDE_TC_DEBUG", 0);
  g_tuning_loaded = true;
}
static inline const TcTuning &tuning() {
  if (!g_tuning_loaded) load_tuning();
  return g_tuning;
}
void tc_reload_tuning() { load_tuning(); }

// Per-DEVICE caches (one process may drive several GPUs): SM count and "dynamic shared memory limit raised" flags.
constexpr int MAX_DEVICES = 64;
static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) dev = 0;
  return dev;
}
static int sm_count() {
  static int n[MAX_DEVICES] = {0};
  const int dev = current_device();
  if (n[dev] == 0) {
    if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

template <int BN, int STAGES, bool SMK, bool TMA_A>
static int launch_tc_impl(const GemmArgs &a, const float *Wp, int wp_na, cudaStream_t st, const CUtensorMap &tm) {
  const int stride = tc_table_stride(a);
  const int rows = has_xf(a.xfa) ? tc_rows_per_tile(a.xfa.R) : 0;
  const int total = tc_stages_bytes(BN, STAGES) + rows * stride * 16 + TC_CTRL_BYTES + 1024 /* alignment slack */;
  if (total > TC_MAX_DYN_SMEM) return SLIDE_ERR_UNSUPPORTED;
  static bool configured[MAX_DEVICES] = {false};  // function attributes are per device
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, SMK, TMA_A>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         TC_MAX_DYN_SMEM);
    if (e != cudaSuccess) return cuda_rc(e);
    configured[dev] = true;
  }
  dim3 grid(ceil_div(a.N, BN), ceil_div(a.M, TBM));
  if (grid.y > 65535) return SLIDE_ERR_UNSUPPORTED;
  const int dbg = tuning().debug;
  launch_k(gemm_tc_kernel<BN, STAGES, SMK, TMA_A>, grid, TC_THREADS, total, st, a, Wp, wp_na, stride, rows, dbg, tm);
  return after_launch();
}

constexpr int TCP_NOT_APPLICABLE = 1;  // internal: take the one-tile-per-CTA kernel instead

template <int BN, bool XFA, bool SMK>
static int launch_tcp(const GemmArgs &a, const float *Wp, int wp_na, cudaStream_t st, const CUtensorMap &tm) {
  constexpr int PSTAGES = BN == 256 ? 3 : 5;
  const int stride = XFA ? tc_table_stride(a) : 0;
  const int total = tcp_smem_bytes(BN, PSTAGES) + (XFA ? XF_MAXG * 8 + stride * 16 : 0) + 1024 /* alignment slack */;
  if (total > TC_MAX_DYN_SMEM) return TCP_NOT_APPLICABLE;
  static bool configured[MAX_DEVICES] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcp_kernel<BN, PSTAGES, XFA, SMK>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_DYN_SMEM);
    if (e != cudaSuccess) return cuda_rc(e);
    configured[dev] = true;
  }
  const int tiles = (a.M / TBM) * ceil_div(a.N, BN);
  const int max_grid = tuning().persist_grid > 0 ? tuning().persist_grid : sm_count();  // tests shrink it: many tiles per CTA
  const int grid = tiles < max_grid ? tiles : max_grid;
  launch_k(gemm_tcp_kernel<BN, PSTAGES, XFA, SMK>, grid, XFA ? TCP_XFA_THREADS : TCP_THREADS, total, st, a, Wp, wp_na, stride, tm);
  return after_launch();
}

template <int BN, int STAGES>
static int launch_tc(const GemmArgs &a, const float *Wp, int wp_na, cudaStream_t st) {
  const TcTuning &tn = tuning();
  const int use_tma = tn.use_tma, persist = tn.persist, persist_min_tiles = tn.persist_min_tiles,
            persist_min_k = tn.persist_min_k;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  // every row tile inside one sample, whole row tiles, 8-row blocks inside one ev row
  const bool one_sample = (!a.res || !has_xf(a.xfr) || a.xfr.R % TBM == 0) && (!a.st_stats || a.st_R % TBM == 0) &&
                          a.M % TBM == 0 && (!a.ev || a.evdiv % 8 == 0);
  const int tiles = (a.M / TBM) * ceil_div(a.N, BN);
  const bool xfa = has_xf(a.xfa);
  if (persist >= 2 && one_sample && tiles >= persist_min_tiles && (xfa || a.smk > 0) &&
      (!xfa || a.xfa.R % TBM == 0) && (a.smk == 0 || (xfa && a.res)) && use_tma && make_a_map(a, &tm, true)) {
    const int rc = a.smk > 0 ? launch_tcp<BN, true, true>(a, Wp, wp_na, st, tm) : launch_tcp<BN, true, false>(a, Wp, wp_na, st, tm);
    if (rc != TCP_NOT_APPLICABLE) return rc;
  }
  if (a.smk > 0) return launch_tc_impl<BN, STAGES, true, false>(a, Wp, wp_na, st, tm);
  // A needs no transform and its rows are 16-byte aligned with a 16-byte pitch: let TMA fetch it
  if (use_tma && !xfa && make_a_map(a, &tm)) {
    // A/B on B200 after the per-tile statistics reduction (feature / position DDPM step, us): min K 256: 1702 / 861,
    // 64: 1703 / 835, 32: 1677 / 836; narrower tiles (BN < 128) gain nothing.  (Before that change short K loops were
    // epilogue-bound and the persistent kernel lost: K=60 N=256 60 -> 85 us.)
    if (persist >= 1 && one_sample && tiles >= persist_min_tiles && BN >= tn.persist_min_bn &&
        a.K >= persist_min_k) {
      const int rc = launch_tcp<BN, false, false>(a, Wp, wp_na, st, tm);
      if (rc != TCP_NOT_APPLICABLE) return rc;
    }
    return launch_tc_impl<BN, STAGES, false, true>(a, Wp, wp_na, st, tm);
  }
  return launch_tc_impl<BN, STAGES, false, false>(a, Wp, wp_na, st, tm);
}

// =====================================================================================================
// Transform pre-pass for row tiles that span several samples.  With 16 rows per sample a 128-row tile needs 8 table
// rows and the in-loop transform reads one table entry per ELEMENT (measured: M=4096 K=256 N=256 53 us with the
// fused transform vs 16 us TMA-fed); the tensors are tiny (M x K x 4 B <= a few MB, L2-resident), so normalising
// them once (~4 us) and feeding the GEMM by TMA is the faster order.  Same (scale, shift, add) arithmetic as the
// fused path; the copy engine rounds to TF32.
// =====================================================================================================
constexpr int PP_COLS = 64;

__global__ void __launch_bounds__(256) xf_prepass_kernel(GemmArgs a, float *__restrict__ out, int ldo) {
  pdl_wait();
  pdl_trigger();
  __shared__ float4 tab[XF_MAXS * PP_COLS];
  __shared__ float2 mr[XF_MAXS * XF_MAXG];
  const int m0 = blockIdx.y * TBM, k0 = blockIdx.x * PP_COLS;
  const int mlast = min(m0 + TBM, a.M) - 1;
  const int s0 = m0 / a.xfa.R, ns = mlast / a.xfa.R - s0 + 1;
  const int ncols = min(PP_COLS, a.K - k0);
  const int step = a.step ? *a.step : 0;
  fill_xf_table(a.xfa, tab, mr, s0, ns, ncols, k0, PP_COLS, step, threadIdx.x, 256);
  __syncthreads();
  const bool relu = a.xfa.relu != 0;
  for (int e = threadIdx.x; e < TBM * (PP_COLS / 4); e += 256) {
    const int r = e / (PP_COLS / 4), c4 = (e - r * (PP_COLS / 4)) * 4;
    const int m = m0 + r, k = k0 + c4;
    if (m >= a.M || c4 >= ncols) continue;
    const float4 x = *reinterpret_cast<const float4 *>(a.A + (size_t)m * a.lda + k);
    const float4 *t = tab + (m / a.xfa.R - s0) * PP_COLS + c4;
    float v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (c4 + u < ncols) {
        const float4 c = t[u];
        float y = fmaf(v[u], c.x, c.y);
        if (relu) y = fmaxf(y, 0.f);
        v[u] = y + c.z;
      } else {
        v[u] = 0.f;
      }
    }
    *reinterpret_cast<float4 *>(out + (size_t)m * ldo + k) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

static GemmArgs prepass_args(const GemmArgs &a, float *scratch) {
  GemmArgs b = a;
  b.A = scratch;
  b.lda = (a.K + 3) & ~3;
  b.xfa.stats = nullptr;
  b.xfa.addvec = nullptr;
  b.xfa.relu = 0;
  return b;
}

bool gemm_tc_prepass_applicable(const GemmArgs &a) {
  if (tuning().prepass == 0) return false;
  // A/B on B200 (M = 4096): K = 256 55 -> 25 us; K <= 128 18 -> 22 us (the extra launch costs more than it saves)
  if (a.K <= 128) return false;
  if (!has_xf(a.xfa) || a.xfa.R % TBM == 0 || TBM % a.xfa.R != 0 || TBM / a.xfa.R > XF_MAXS) return false;
  if (a.xfa.stats && a.xfa.nnorm / a.xfa.cg > XF_MAXG) return false;
  if (((uintptr_t)a.A & 15) || (a.lda & 3)) return false;
  return true;
}

size_t gemm_tc_prepass_bytes(int M, int K) { return (size_t)M * ((K + 3) & ~3) * 4; }

int launch_gemm_tc_prepass(const GemmArgs &a, const float *Wp, int wp_na, float *scratch, cudaStream_t st) {
  const GemmArgs b = prepass_args(a, scratch);
  if (!gemm_tc_eligible(b, Wp)) return TCP_NOT_APPLICABLE;
  dim3 grid(ceil_div(a.K, PP_COLS), ceil_div(a.M, TBM));
  launch_k(xf_prepass_kernel, grid, 256, 0, st, a, scratch, b.lda);
  const int rc = after_launch();
  if (rc != SLIDE_OK) return rc;
  return launch_gemm_tc(b, Wp, wp_na, st);
}

int launch_gemm_tc(const GemmArgs &a, const float *Wp, int wp_na, cudaStream_t st) {
  // Tile width.  Large problems (more tiles than one wave of 2 CTAs x 148 SMs) take the widest tile that does not
  // pad N by more than 2x: fewest re-reads of A.  Small problems are latency-bound per CTA (prologue + a short K
  // loop), so they take the narrowest tile that still fits in ONE wave: same critical path, less epilogue each.
  const int mt = ceil_div(a.M, TBM);
  const int wave = 2 * 148;
#ifndef TC_MAX_BN
#define TC_MAX_BN 256
#endif
#ifndef TC_STAGES_256
#define TC_STAGES_256 2
#endif
#ifndef TC_STAGES_128
#define TC_STAGES_128 3
#endif
  int widest = a.N > 128 ? 256 : a.N > 64 ? 128 : a.N > 32 ? 64 : 32;
  if (widest > TC_MAX_BN) widest = TC_MAX_BN;
  int bn = widest;
  if (mt * ceil_div(a.N, widest) < wave) {
    for (int c = 32; c <= widest; c <<= 1) {
      if (mt * ceil_div(a.N, c) <= wave) {
        bn = c;
        break;
      }
    }
  }
  switch (bn) {
    case 256: return launch_tc<256, TC_STAGES_256>(a, Wp, wp_na, st);
    case 128: return launch_tc<128, TC_STAGES_128>(a, Wp, wp_na, st);
    case 64: return launch_tc<64, 4>(a, Wp, wp_na, st);
    default: return launch_tc<32, 4>(a, Wp, wp_na, st);
  }
}

#ifdef TC_TIMELINE
extern "C" int slide_debug_tc_timeline(unsigned long long *out, int n_ctas, int clear) {
  if (n_ctas > TL_CTAS) n_ctas = TL_CTAS;
  cudaError_t e = cudaMemcpyFromSymbol(out, g_tc_tl, (size_t)n_ctas * TL_SLOTS * 8);
  if (e == cudaSuccess && clear) {
    void *p = nullptr;
    e = cudaGetSymbolAddress(&p, g_tc_tl);
    if (e == cudaSuccess) e = cudaMemset(p, 0, sizeof(unsigned long long) * TL_SLOTS * TL_CTAS);
  }
  return e == cudaSuccess ? 0 : -2;
}
#endif

int tc_error_flag() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_tc_error, sizeof(int));
  return v;
}

void tc_error_reset() {
  const int z = 0;
  cudaMemcpyToSymbol(g_tc_error, &z, sizeof(int));
}

}  // namespace slide
