// Hardware probe (B200): issue rate per SM sub-partition of the non-MUFU instructions of the A generation
// (cvt.rn.f16x2.f32 = F2FP.PACK_AB, f16 -> f32 widening, HFMA2, HMUL2, PRMT, FFMA, LDS.128, indexed LDC.64).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

struct P { float c[1024]; };

template <int OP>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cyc, int iters, const __grid_constant__ P prm) {
  __shared__ uint4 sm[512];
  uint32_t v[8];
  float f[8];
  for (int i = 0; i < 8; ++i) { v[i] = threadIdx.x * 8 + i + 0x3c003c00u; f[i] = 1.0f + i + threadIdx.x; }
  sm[threadIdx.x] = make_uint4(v[0], v[1], v[2], v[3]);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(v[i]) : "f"(f[i]), "f"(f[(i + 1) & 7]));
      if (OP == 1) { asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(f[i]) : "r"(v[i])); }
      if (OP == 2) asm volatile("fma.rn.f16x2 %0, %0, %1, %0;" : "+r"(v[i]) : "r"(v[(i + 1) & 7]));
      if (OP == 3) asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(v[i]) : "r"(0x54005400u));
      if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(v[i]) : "r"(v[(i + 1) & 7]));
      if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]));
      if (OP == 6) { uint4 q = sm[(threadIdx.x + i * 37 + it) & 511]; v[i] ^= q.x ^ q.y ^ q.z ^ q.w; }
      if (OP == 7) { const int idx = ((it + i) & 63) * 2 + (threadIdx.x >> 5 & 1) * 128; float2 c2 = *reinterpret_cast<const float2*>(&prm.c[idx]); f[i] += c2.x + c2.y; }
      if (OP == 8) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(v[i]) : "r"(v[(i + 1) & 7]));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= v[i] ^ __float_as_uint(f[i]);
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
static void run(const char* name, uint32_t* out, long long* cyc, const P& prm) {
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) k<OP><<<1, 512>>>(out, cyc, iters, prm);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s %6.2f cycles per warp-instruction per sub-partition\n", name, h / (4.0 * 8 * iters));
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 512 * 4);
  cudaMalloc(&cyc, 8);
  static P prm;
  for (int i = 0; i < 1024; ++i) prm.c[i] = 1e-3f * i;
  run<0>("cvt.rn.f16x2.f32 (F2FP.PACK_AB)", out, cyc, prm);
  run<1>("cvt.f32.f16 (widen)", out, cyc, prm);
  run<2>("fma.rn.f16x2 (HFMA2)", out, cyc, prm);
  run<3>("mul.rn.f16x2 (HMUL2)", out, cyc, prm);
  run<4>("prmt.b32", out, cyc, prm);
  run<5>("fma.rn.f32 (FFMA)", out, cyc, prm);
  run<6>("ld.shared.v4 + 4 xor", out, cyc, prm);
  run<7>("indexed ld.param f32x2 (LDC.64) + 2 add", out, cyc, prm);
  run<8>("add.rn.f16x2 (HADD2)", out, cyc, prm);
  return 0;
}
