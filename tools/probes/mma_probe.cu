// Micro-probe: throughput of the legacy mma.sync path on sm_100a (TF32 m16n8k8, BF16 m16n8k16) per SM,
// used to size the sample-resident position-DDPM kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int NACC, int KIND>
__global__ void probe(float *out, int iters) {
  float acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i)
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = 11, b1 = 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 12345.f) out[0] = s;
}

template <int NACC, int KIND>
void run(int warps, const char *name) {
  float *out;
  cudaMalloc(&out, 4);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<NACC, KIND><<<148, warps * 32>>>(out, 100);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<NACC, KIND><<<148, warps * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double macs = (double)148 * warps * iters * NACC * (KIND == 0 ? 16 * 8 * 8 : 16 * 8 * 16);
  printf("%s warps/SM=%2d acc=%d: %.3f ms  %.1f TFLOP/s  %.0f MAC/clk/SM @1.9GHz\n", name, warps, NACC, ms,
         2 * macs / ms / 1e9, macs / 148 / (ms * 1e-3 * 1.9e9));
  cudaFree(out);
}

int main() {
  run<8, 0>(4, "tf32 m16n8k8 ");
  run<8, 0>(8, "tf32 m16n8k8 ");
  run<8, 0>(16, "tf32 m16n8k8 ");
  run<4, 0>(16, "tf32 m16n8k8 ");
  run<16, 0>(16, "tf32 m16n8k8 ");
  run<8, 0>(32, "tf32 m16n8k8 ");
  run<8, 1>(8, "bf16 m16n8k16");
  run<8, 1>(16, "bf16 m16n8k16");
  run<8, 1>(32, "bf16 m16n8k16");
  return 0;
}
