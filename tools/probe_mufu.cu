// Hardware probe (B200): throughput of the MUFU variants the edge kernels use, per SM sub-partition.
//   tanh.approx.f32 (MUFU.TANH), tanh.approx.f16x2 (2 x MUFU.TANH.F16), tanh.approx.bf16x2 (2 x MUFU.TANH.BF16),
//   ex2.approx.f32 (MUFU.EX2), ex2.approx.f16x2.
// One CTA of 512 threads (4 warps per sub-partition) on one SM; each thread runs 8 independent chains.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x) {
  uint32_t r;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == 1) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == 2) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == 3) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == 4) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == 5) asm volatile("tanh.approx.f16 %0, %1;" : "=h"(*(uint16_t*)&r) : "h"((uint16_t)x));
  return r;
}

template <int OP>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cyc, int iters) {
  uint32_t v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 8 + i + 0x3c003c00u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = op<OP>(v[i]);
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= v[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
static void run(const char* name, int results_per_op, uint32_t* out, long long* cyc) {
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) k<OP><<<1, 512>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double warp_instr_per_smsp = 4.0 * 8 * iters;  // 4 warps per sub-partition
  printf("%-22s %7.2f cycles per warp-instruction per sub-partition  (%5.2f results / clk / SM)\n", name, h / warp_instr_per_smsp,
         512.0 * 8 * iters * results_per_op / h);
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 512 * 4);
  cudaMalloc(&cyc, 8);
  run<0>("tanh.approx.f32", 1, out, cyc);
  run<1>("tanh.approx.f16x2", 2, out, cyc);
  run<2>("tanh.approx.bf16x2", 2, out, cyc);
  run<3>("ex2.approx.ftz.f32", 1, out, cyc);
  run<4>("ex2.approx.f16x2", 2, out, cyc);
  run<5>("tanh.approx.f16", 1, out, cyc);
  return 0;
}
