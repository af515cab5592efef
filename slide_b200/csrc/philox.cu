// Seeked Philox draw of a SLICE of torch's CUDA normal_() stream (include/slide_b200.h: slide_philox_normal_slice).
//
// The reference draws the feature DDPM's per-step noise with torch.randn_like on the GPU
// (pointnet2/diffusion_utils/diffusion.py:88): T calls of normal_() on a (B,16,C) fp32 tensor.  A rank that owns the
// samples [lo, lo+Bl) of a batch sharded over W GPUs needs only its rows of every call, bit-identical to the full draw
// (results must not depend on W).  ATen's kernel (distribution_nullary_kernel, unroll 4, block 256) gives thread `idx`
// of a grid of G blocks the Philox4_32_10 subsequence `idx`; on its k-th loop iteration it draws ONE curand_normal4 and
// writes component ii to element  li = k * 4 * 256 G + ii * 256 G + idx.  Inverting that map, element li of call s is
//   component (li / 256G) % 4  of  curand_normal4(state(seed, subsequence = li % 256G, offset0 + s * inc + 4 * (li / (4 * 256G))))
// so any slice of any call can be produced directly: ONE launch for all T calls instead of T launches of the full batch
// plus T slice copies per rank.  curand's own device functions are used (header-only), so the arithmetic (Box-Muller
// with logf / sqrtf / __sincosf) is the one torch runs.
#include <curand_kernel.h>

#include "common.cuh"

namespace slide {

__global__ void __launch_bounds__(256) philox_normal_slice_kernel(float *__restrict__ out, long long out_call_stride,
                                                                  int n_calls, int reverse, unsigned long long seed,
                                                                  unsigned long long offset0, unsigned long long inc,
                                                                  long long slice_begin, long long slice_len,
                                                                  long long threads_full) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n_calls * slice_len;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long s = e / slice_len, r = e - s * slice_len;
    const long long li = slice_begin + r;
    const long long k = li / (4 * threads_full);
    const long long rem = li - k * 4 * threads_full;
    const int ii = (int)(rem / threads_full);
    const long long idx = rem - ii * threads_full;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)idx, offset0 + (unsigned long long)s * inc + 4ull * (unsigned long long)k, &st);
    const float4 v = curand_normal4(&st);
    const float x = ii == 0 ? v.x : ii == 1 ? v.y : ii == 2 ? v.z : v.w;
    const long long row = reverse ? (n_calls - 1 - s) : s;
    out[row * out_call_stride + r] = x * 1.0f + 0.0f;  // normal_(mean = 0, std = 1)'s transform
  }
}

}  // namespace slide

extern "C" int slide_philox_normal_slice(float *out, long long out_call_stride, int n_calls, int reverse,
                                         unsigned long long seed, unsigned long long offset,
                                         unsigned long long offset_increment, long long numel, long long slice_begin,
                                         long long slice_len, int grid_full, slide_stream_t stream) {
  using namespace slide;
  if (!out || n_calls <= 0 || numel <= 0 || slice_begin < 0 || slice_len <= 0 || slice_begin + slice_len > numel ||
      grid_full <= 0 || out_call_stride < slice_len || (offset & 3ull) || (offset_increment & 3ull))
    return SLIDE_ERR_INVALID;
  const long long total = (long long)n_calls * slice_len;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  launch_k(philox_normal_slice_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, 
      out, out_call_stride, n_calls, reverse, seed, offset, offset_increment, slice_begin, slice_len, 256LL * grid_full);
  return after_launch();
}
