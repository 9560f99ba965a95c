// Hardware probe (B200): where does tcgen05.mma.cta_group::2 with M = 128 (64 rows per CTA) put its accumulator, and where
// does it expect a TS-mode A operand?  DESIGN.md section 9 needs this for the double-buffered-accumulator plan: with
// M = 256 per pair (the shape the edge kernel uses) one 128 x 448 fp32 accumulator fills TMEM; if an M = 128 pair tile
// occupies 128 lanes x N/2 columns (or 64 lanes x N columns) two accumulators fit.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I ml_conformer_generator_b200/csrc \
//        -o build_ab/probe_tmem tools/probe_tmem_pair_m128.cu && build_ab/probe_tmem
//
// Test 1 (SS mode): A[m][k], B[n][k] chosen so that D[m][n] = m + 256 n exactly; both CTAs dump all 128 lanes x N columns
//                   of their TMEM and the host prints which (m, n) each (cta, lane, column) holds.
// Test 2 (TS mode): B = identity (N = K = 16), A written to TMEM with tcgen05.st as packed bf16 that encodes first the
//                   lane and then the (column, half) it was written to; D[m][n] = A[m][n] then tells which TMEM cell the
//                   hardware reads for logical A[m][k].
#include <cstdio>
#include <cstring>
#include <vector>

#include "mlcg_common.cuh"

using namespace mlcg;

constexpr int PN = 64;  // N of the probe MMA

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
__device__ __forceinline__ void st_bf16(uint8_t* chunk, int row, int k, float v) {
  reinterpret_cast<__nv_bfloat16*>(chunk + sw128_offset(row, k >> 3))[k & 7] = __float2bfloat16_rn(v);
}

// mode 0: SS probe (M total = m_total: 128 or 256).  mode 1 / 2: TS probe, A cells encode lane / (column*2 + half).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_probe(int mode, int m_total, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* As = g;            // [128 rows x 128 B] operand chunk (K = 64 bf16), this CTA's rows of A
  uint8_t* Bs = g + 16384;    // [N/2 rows x 128 B], this CTA's rows of B
  const uint32_t bar = base + 32768;
  uint32_t* slot = reinterpret_cast<uint32_t*>(g + 32768 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cr = cluster_ctarank();
  const int m_cta = m_total / 2;  // rows of A / D owned by this CTA
  const int n_mma = (mode == 0) ? PN : 16;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0;
  __syncthreads();
  if (mode == 0) {
    // D[m][n] = m + 256 n:  A[m][0] = m, A[m][1] = 1;  B[n][0] = 1, B[n][1] = 128 n
    for (int r = threadIdx.x; r < m_cta; r += blockDim.x) {
      st_bf16(As, r, 0, (float)(cr * m_cta + r));
      st_bf16(As, r, 1, 1.0f);
    }
    for (int r = threadIdx.x; r < n_mma / 2; r += blockDim.x) {
      st_bf16(Bs, r, 0, 1.0f);
      st_bf16(Bs, r, 1, 256.0f * (float)(cr * (n_mma / 2) + r));
    }
  } else {
    // B = identity over (n, k), n, k < 16: this CTA holds rows n = 8 cr .. 8 cr + 7
    for (int r = threadIdx.x; r < 8; r += blockDim.x) st_bf16(Bs, r, cr * 8 + r, 1.0f);
  }
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
  {  // poison the dumped region so that cells the MMA does not write are recognisable
    float neg[32];
    for (int e = 0; e < 32; ++e) neg[e] = -1.0f;
    const uint32_t prow = tb + ((uint32_t)(warp * 32) << 16);
    tmem_st32(prow, neg);
    tmem_st32(prow + 32, neg);
    tmem_wait_st();
    tc_fence_before();
  }
  if (mode != 0) {
    // fill TMEM columns 256..263 of every lane with packed bf16 (2 per column): K = 16 elements of A
    const uint32_t trow = tb + ((uint32_t)(warp * 32) << 16) + 256;
    float v[16];
    for (int c = 0; c < 8; ++c) {
      const float lo = (mode == 1) ? (float)(warp * 32 + lane) : (float)(2 * c);
      const float hi = (mode == 1) ? (float)(warp * 32 + lane) : (float)(2 * c + 1);
      v[c] = __uint_as_float(pack_bf16x2(lo, hi));
    }
    for (int c = 8; c < 16; ++c) v[c] = 0.f;
    tmem_st16(trow, v);
    tmem_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (cr == 0 && threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(1, m_total, n_mma);
    if (mode == 0) mma_ss_pair(tb, umma_desc_sw128(base), umma_desc_sw128(base + 16384), idesc, 0);
    else mma_ts_pair(tb, tb + 256, umma_desc_sw128(base + 16384), idesc, 0);
    umma_commit_pair(bar, 3);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  // dump lanes 32*warp .. +31, columns 0 .. 63 of this CTA's TMEM
  const uint32_t trow = tb + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 64; c0 += 32) {
    float v[32];
    tmem_ld32(trow + c0, v);
    tmem_wait_ld();
    for (int e = 0; e < 32; ++e) out[((size_t)cr * 128 + warp * 32 + lane) * 64 + c0 + e] = v[e];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tb);
}

static void run(int mode, int m_total, std::vector<float>& host) {
  float* d = nullptr;
  cudaMalloc(&d, 2 * 128 * 64 * sizeof(float));
  cudaMemset(d, 0xff, 2 * 128 * 64 * sizeof(float));  // NaN pattern = "never written by the MMA"
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  k_probe<<<2, 128, 40000>>>(mode, m_total, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mode %d M %d: CUDA error %s\n", mode, m_total, cudaGetErrorString(e));
    exit(1);
  }
  host.resize(2 * 128 * 64);
  cudaMemcpy(host.data(), d, host.size() * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d);
}

int main() {
  std::vector<float> h;
  for (int m_total : {256, 128}) {
    // first poison TMEM contents so stale data is recognisable: run the probe twice with different shapes is enough here
    run(0, m_total, h);
    printf("== SS probe, cta_group::2, M = %d, N = %d: D[m][n] = m + 256 n ==\n", m_total, PN);
    for (int cta = 0; cta < 2; ++cta)
      for (int lane = 0; lane < 128; lane += 8) {
        printf("cta %d lane %3d:", cta, lane);
        for (int col : {0, 1, 31, 32, 63}) {
          const float v = h[((size_t)cta * 128 + lane) * 64 + col];
          if (v != v || v < 0 || v > 256 * 64) printf("  c%-2d=   --    ", col);
          else printf("  c%-2d=(m%3d,n%2d)", col, (int)v % 256, (int)v / 256);
        }
        printf("\n");
      }
  }
  for (int mode : {1, 2}) {
    run(mode, 128, h);
    printf("== TS probe, cta_group::2, M = 128, K = 16: D[m][k] = %s of the TMEM cell read for A[m][k] (A written at columns "
           "256..263 of every lane) ==\n", mode == 1 ? "LANE" : "(column-256)*2 + half");
    for (int cta = 0; cta < 2; ++cta)
      for (int lane = 0; lane < 128; lane += 8) {
        printf("cta %d D-lane %3d:", cta, lane);
        for (int k = 0; k < 16; ++k) {
          const float v = h[((size_t)cta * 128 + lane) * 64 + k];
          if (v != v || v < 0) printf("  --");
          else printf(" %3d", (int)v);
        }
        printf("\n");
      }
  }
  return 0;
}
