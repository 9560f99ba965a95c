// Hardware probe (B200), second version: cycles per tcgen05.mma.cta_group::2 (M = 256, K = 16, kind::f16, TS mode) as a
// function of N with a straight-line issue sequence (no per-MMA address arithmetic), so that the single issuing thread is
// not the limit.  Prints issue time and completion time per MMA.
#include <cstdio>
#include "mlcg_common.cuh"
using namespace mlcg;

__device__ __forceinline__ void mma_ts_pair(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void mma_ss_pair(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc)
      : "memory");
}

template <int N, int TS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_rate(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + 49152;
  uint32_t* slot = reinterpret_cast<uint32_t*>(g + 49152 + 64);
  const int warp = threadIdx.x >> 5;
  const uint32_t cr = cluster_ctarank();
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_pair<512>(smem_u32(slot));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = *slot;
  {
    float z[32];
    for (int e = 0; e < 32; ++e) z[e] = 0.f;
    const uint32_t prow = tb + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 512; c0 += 32) tmem_st32(prow + c0, z);
    tmem_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (cr == 0 && threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc(0, 256, N);
    const uint64_t adesc = umma_desc_sw128(base);
    const uint64_t bdesc = umma_desc_sw128(base + 16384);
    const uint32_t d = tb + 224, a = tb;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
#pragma unroll
      for (int u = 0; u < 28; ++u) {
        if (TS) mma_ts_pair(d, a + 8 * (u & 3) + 32 * (u >> 2), bdesc + 2 * (u & 3), idesc);
        else mma_ss_pair(d, adesc + 2 * (u & 3), bdesc + 2 * (u & 3), idesc);
      }
    }
    const long long t1 = clock64();
    umma_commit_pair(bar, 3);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  } else {
    mbar_wait(bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tb);
}

template <int N, int TS>
static void run(long long* d) {
  cudaFuncSetAttribute(k_rate<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 52000);
  for (int rep = 0; rep < 2; ++rep) {
    k_rate<N, TS><<<2, 128, 52000>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("N %d: CUDA error\n", N); return; }
  }
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%4d  %s  | issue %7.1f   total %7.1f   floor N/2 = %5.1f\n", N, TS ? "TS" : "SS", h[0] / 448.0, h[1] / 448.0, N / 2.0);
}

int main() {
  long long* d = nullptr;
  cudaMalloc(&d, 4 * sizeof(long long));
  printf("   N mode | cycles per MMA (448 back-to-back MMAs, M = 256 pair, K = 16, fp16, D at column 224)\n");
  run<16, 1>(d); run<32, 1>(d); run<64, 1>(d); run<96, 1>(d); run<112, 1>(d); run<128, 1>(d); run<144, 1>(d); run<160, 1>(d);
  run<192, 1>(d); run<224, 1>(d); run<256, 1>(d);
  run<32, 0>(d); run<64, 0>(d); run<128, 0>(d); run<144, 0>(d); run<224, 0>(d); run<256, 0>(d);
  return 0;
}
