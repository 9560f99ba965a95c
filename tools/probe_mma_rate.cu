// Hardware probe (B200): cycles per tcgen05.mma.cta_group::2 (M = 256, K = 16, kind::f16, A from TMEM = TS mode, or from
// shared memory = SS mode) as a function of N, of the accumulator's column offset and of whether consecutive MMAs
// accumulate into the same TMEM region.  Background: k_tc_edge3 (resident A, three N = 144 passes) measured ~270 cycles
// per MMA where the N / 2 floor says 72.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I ml_conformer_generator_b200/csrc \
//        -o build_ab/probe_mma_rate tools/probe_mma_rate.cu && build_ab/probe_mma_rate
#include <cstdio>
#include <vector>

#include "mlcg_common.cuh"

using namespace mlcg;

__device__ __forceinline__ void mma_ss_pair(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts_pair(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ bool mbar_test_done(uint32_t bar) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(0u)
      : "memory");
  return done != 0;
}

struct Cfg {
  int n, ts, dcol0, dcol1, alt, acol, reps, a_span;  // a_span: A column advances by 8 per MMA modulo a_span (TS)
  int commit_every;  // > 0: tcgen05.commit (multicast to both CTAs) on a scratch barrier after every commit_every-th MMA
  int ld_traffic;    // 1: the other three warps of both CTAs stream tcgen05.ld from columns 384.. while the MMAs run; 2: tcgen05.st
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_rate(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + 49152;
  uint32_t* slot = reinterpret_cast<uint32_t*>(g + 49152 + 64);
  const int warp = threadIdx.x >> 5;
  const uint32_t cr = cluster_ctarank();
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0;
  const uint32_t bar2 = bar + 8, bar3 = bar + 16;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1 << 20);  // scratch target of the extra commits: never completes a phase
    mbar_init(bar3, 1);
    fence_barrier_init();
  }
  volatile int* stop = reinterpret_cast<volatile int*>(g + 49152 + 128);
  if (threadIdx.x == 0) *stop = 0;
  if (warp == 0) tmem_alloc_pair<512>(smem_u32(slot));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = *slot;
  {  // zero all of TMEM so that the accumulators stay finite
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
    const uint32_t idesc = umma_idesc(0, 256, c.n);
    const uint64_t adesc = umma_desc_sw128(base);           // SS: A chunk [128 x 128 B]
    const uint64_t bdesc = umma_desc_sw128(base + 16384);   // B: this CTA's n/2 rows x 128 B
    const long long t0 = clock64();
    // alt = log2 of the run length on one accumulator (0: never switch); a_span is a power of two
    const uint32_t amask = (uint32_t)c.a_span - 1u, dsel = c.alt > 0 ? 1u : 0u;
    const uint32_t d0 = tb + c.dcol0, d1 = tb + c.dcol1, a0 = tb + c.acol;
#pragma unroll 4
    for (int i = 0; i < c.reps; ++i) {
      const int ks = i & 3;
      const uint32_t d = (dsel & ((uint32_t)i >> c.alt)) ? d1 : d0;
      if (c.ts) mma_ts_pair(d, a0 + (((uint32_t)i * 8u) & amask), bdesc + 2 * ks, idesc, 1);
      else mma_ss_pair(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, 1);
      if (c.commit_every > 0 && (i % c.commit_every) == c.commit_every - 1) umma_commit_pair(bar2, 3);
    }
    const long long t1 = clock64();
    umma_commit_pair(bar, 3);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  } else if (c.ld_traffic && warp > 0) {
    // background TMEM traffic from three warps per CTA until the MMAs are done
    const uint32_t prow = tb + ((uint32_t)(warp * 32) << 16) + 400;
    float v[32];
    for (int e = 0; e < 32; ++e) v[e] = 0.f;
    while (!mbar_test_done(bar)) {
      if (c.ld_traffic == 1) {
        tmem_ld32(prow, v);
        tmem_wait_ld();
      } else {
        tmem_st32(prow, v);
        tmem_wait_st();
      }
    }
    if (v[3] == 123.f) out[2] = 1;
  } else {
    mbar_wait(bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tb);
}

int main() {
  long long* d = nullptr;
  cudaMalloc(&d, 4 * sizeof(long long));
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 52000);
  std::vector<Cfg> cfgs;
  const int R = 448;
  for (int n : {32, 144, 224, 256}) cfgs.push_back({n, 1, 256, 256, 0, 0, R, 32, 0, 0});
  for (int ce : {1, 2, 4, 8}) cfgs.push_back({144, 1, 224, 224, 0, 0, R, 128, ce, 0});
  for (int ce : {1, 4, 8}) cfgs.push_back({224, 1, 0, 224, 2, 448, R, 64, ce, 0});
  for (int lt : {1, 2}) cfgs.push_back({144, 1, 224, 224, 0, 0, R, 128, 0, lt});
  for (int lt : {1, 2}) cfgs.push_back({224, 1, 0, 224, 2, 448, R, 64, 0, lt});
  for (int lt : {1, 2}) cfgs.push_back({144, 1, 224, 224, 0, 0, R, 128, 4, lt});
  // the edge3 layout: A at columns 0..223, accumulators at 224 / 368
  cfgs.push_back({144, 1, 224, 224, 0, 0, R, 128, 0, 0});
  cfgs.push_back({144, 1, 368, 368, 0, 0, R, 128, 0, 0});
  cfgs.push_back({144, 1, 224, 368, 2, 0, R, 128, 0, 0});
  cfgs.push_back({144, 1, 224, 368, 5, 0, R, 128, 0, 0});
  // the k_tc_edge layout: accumulator halves at 0 / 224, A ring at 448.., alternate every 4 MMAs
  cfgs.push_back({224, 1, 0, 224, 2, 448, R, 64, 0, 0});
  cfgs.push_back({224, 1, 0, 0, 0, 448, R, 64, 0, 0});
  cfgs.push_back({144, 1, 0, 144, 2, 448, R, 64, 0, 0});
  cfgs.push_back({144, 1, 0, 0, 0, 448, R, 64, 0, 0});
  cfgs.push_back({144, 1, 256, 256, 0, 0, R, 128, 0, 0});
  cfgs.push_back({128, 1, 224, 224, 0, 0, R, 128, 0, 0});
  cfgs.push_back({128, 1, 256, 384, 5, 0, R, 128, 0, 0});
  // SS mode for comparison
  for (int n : {128, 144, 224, 256}) cfgs.push_back({n, 0, 256, 256, 0, 0, R, 32, 0, 0});
  printf("   N mode  dcol0 dcol1 alt acol a_span commit traffic | issue cyc/MMA  total cyc/MMA   (floor N/2)\n");
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {  // second run: warm
      k_rate<<<2, 128, 52000>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("N %d: CUDA error %s\n", c.n, cudaGetErrorString(e));
        return 1;
      }
    }
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%4d  %s   %4d  %4d  %3d %4d  %4d    %3d    %d    |  %8.1f      %8.1f        %5.1f\n", c.n, c.ts ? "TS" : "SS", c.dcol0, c.dcol1,
           c.alt, c.acol, c.a_span, c.commit_every, c.ld_traffic, (double)h[0] / c.reps, (double)h[1] / c.reps, c.n / 2.0);
  }
  return 0;
}
