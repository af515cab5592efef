// Shared helpers for the slide_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slide_b200.h"

namespace slide {

extern long long g_launch_count;
void set_cuda_error(cudaError_t e);

// Record the launch and surface launch errors as a return code (the reference exits the process instead).
inline int after_launch() {
  ++g_launch_count;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_cuda_error(e);
    return SLIDE_ERR_CUDA;
  }
  return SLIDE_OK;
}

inline int cuda_rc(cudaError_t e) {
  if (e != cudaSuccess) {
    set_cuda_error(e);
    return SLIDE_ERR_CUDA;
  }
  return SLIDE_OK;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// Every kernel of this library starts with pdl_wait(): when it was launched with the programmatic-serialization
// attribute (launch_k below) its CTAs may become resident while the previous kernel of the stream is still running, and
// this instruction blocks them until that kernel has COMPLETED and its writes are visible -- so nothing a kernel does
// after its first instruction can race with its predecessor.  pdl_trigger() (placed right after the wait) lets the NEXT
// kernel's CTAs be scheduled as soon as all of this kernel's CTAs are running: launch latency and CTA start-up overlap
// the kernel's execution instead of following it.  Without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_pdl_enabled;  // SLIDE_PDL (default 0: measured neutral-to-negative under graph replay, profiles/r02_pdl_ab.txt); read once

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_enabled ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // errors surface through after_launch()
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// a*a + b*b + c*c in the order the reference's sm_100a SASS evaluates it (see oracle/slide_oracle.c):
// FMUL t=b*b ; FFMA t=a*a+t ; FFMA t=c*c+t.  Explicit intrinsics so the compiler cannot re-associate.
__device__ __forceinline__ float sumsq3_ref(float a, float b, float c) {
  float t = __fmul_rn(b, b);
  t = __fmaf_rn(a, a, t);
  return __fmaf_rn(c, c, t);
}

// pytorch3d's `dist += diff*diff` over x,y,z as nvcc contracts it.
__device__ __forceinline__ float sumsq3_p3d(float a, float b, float c) {
  float t = __fmul_rn(a, a);
  t = __fmaf_rn(b, b, t);
  return __fmaf_rn(c, c, t);
}

}  // namespace slide
