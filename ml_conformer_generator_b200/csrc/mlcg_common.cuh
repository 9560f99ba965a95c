// Shared device helpers for the sm_100a kernels: PTX wrappers (mbarrier, bulk async copy, tcgen05 / TMEM),
// operand-tile geometry, activation functions.  Everything here is hand-written for sm_100a; there is no
// fallback path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace mlcg {

// ---------------------------------------------------------------------------------------------------------------
// Geometry of the hot path (reference conformer_generator.py:67-88)
// ---------------------------------------------------------------------------------------------------------------
constexpr int HID = 420;        // hidden_nf
constexpr int HP = 448;         // HID padded to 7 x 64 (clean 128-byte swizzle atoms)
constexpr int BIAS_COL = 420;   // spare K column carrying the folded bias of the 2nd edge layer
constexpr int IN_NF = 12;       // 8 classes + time + 3 context
constexpr int ZC = 11;          // latent channels: 3 coords + 8 classes
constexpr int TILE_M = 128;     // rows per tensor-core tile (= TMEM lanes)
constexpr int CHUNK_BYTES = 128;                      // K bytes per operand chunk row (one SWIZZLE_128B atom row)
constexpr int A_CHUNK_BYTES = TILE_M * CHUNK_BYTES;   // 16 KB: one [128 x 128 B] A-operand chunk
constexpr int SEER_D = 42;
constexpr int SEER_H = 2048;
constexpr int SEER_E = 64;
constexpr int SEER_NB = 5;

enum Precision { PREC_FP32_SIMT = 0, PREC_TF32 = 1, PREC_BF16 = 2, PREC_FP16 = 3 };
// the two 16-bit tensor-core modes (tcgen05 kind::f16) share every kernel; they differ in the operand format only
__host__ __device__ constexpr bool is16(int mode) { return mode == PREC_BF16 || mode == PREC_FP16; }

// elements per 128-byte operand chunk row
__host__ __device__ constexpr int epc(int mode) { return is16(mode) ? 64 : 32; }
// elements per 16-byte piece
__host__ __device__ constexpr int epp(int mode) { return is16(mode) ? 8 : 4; }

// ---------------------------------------------------------------------------------------------------------------
// fp16 mode: range management.  fp16 has tf32's 10-bit mantissa (8x finer than bf16) at bf16's tensor-core rate, but
// only 5 exponent bits (max 65504).  Every 16-bit quantity of the hot path is therefore stored with an exact
// power-of-two scale whose inverse is folded into the consumer (weights packed at load time, fp32 epilogue constants):
//   edge pre-activations / activations / messages   x ACT  (2^-6):  |value| up to 4e6 representable
//   squared distances (packed evaluation)           x DIST (2^-10): d^2 up to 6.7e7
//   node-level GEMM operands h, agg, t              x OP   (2^-4)
// bf16 mode uses scale 1 everywhere (fp32 exponent range).  Scales are powers of two: no rounding is introduced.
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr float act_scale(int mode) { return mode == PREC_FP16 ? 0.015625f : 1.0f; }
__host__ __device__ constexpr float dist_scale(int mode) { return mode == PREC_FP16 ? 0.0009765625f : 1.0f; }
__host__ __device__ constexpr float op_scale(int mode) { return mode == PREC_FP16 ? 0.0625f : 1.0f; }
// neighbour aggregate: the segment sum holds ACT * sum(e); stored value = OP * sum(e) / 100
__host__ __device__ constexpr float agg_out_scale(int mode) { return 0.01f / act_scale(mode) * op_scale(mode); }
// tcgen05 instruction-descriptor operand format: kind::f16 -> 0 = F16, 1 = BF16; kind::tf32 -> 2
__host__ __device__ constexpr int umma_fmt(int mode) { return mode == PREC_FP16 ? 0 : (mode == PREC_BF16 ? 1 : 2); }

// byte offset of (row r, 16-byte piece p) inside a SWIZZLE_128B K-major chunk whose base is 1024-byte aligned:
// Swizzle<3,4,3>: address bits [4,7) ^= bits [7,10).
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t p) {
  return r * 128u + ((p ^ (r & 7u)) << 4);
}

// fp32 residual stream h: index of (node, channel).  ldh > 0: row-major [node][ldh] (exact-fp32 SIMT path).
// ldh == 0: tiled [node/128][channel/32][node%128][channel%32] -- in the tensor-core GEMM epilogue a thread owns one row
// and 32 consecutive columns, so a warp touches 32 consecutive 128-byte segments (fully coalesced).
__host__ __device__ __forceinline__ size_t hres_index(int node, int c, int ldh) {
  return ldh > 0 ? (size_t)node * ldh + c
                 : ((size_t)(node >> 7) * (HP / 32) + (c >> 5)) * (TILE_M * 32) + (size_t)(node & 127) * 32 + (c & 31);
}

// ---------------------------------------------------------------------------------------------------------------
// activations
// ---------------------------------------------------------------------------------------------------------------
template <bool kFast>
__device__ __forceinline__ float silu(float x) {
  if constexpr (kFast) {
    // x*sigmoid(x) = h + h*tanh(h), h = x/2 : one MUFU op
    float h = 0.5f * x, t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  } else {
    return __fdividef(x, 1.0f + __expf(-x));
  }
}
// SiLU when the argument has already been halved (bf16 fast mode folds the 1/2 into the packed weights, exactly):
// silu(2h) = h + h*tanh(h).  In precise mode the argument is not scaled and the exact-ish form is used.
template <bool kFast>
__device__ __forceinline__ float silu_scaled(float h) {
  if constexpr (kFast) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  } else {
    return __fdividef(h, 1.0f + __expf(-h));
  }
}
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// full-precision SiLU for the fp32 SIMT path
__device__ __forceinline__ float silu_ref(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ uint32_t f32_to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// packed 16-bit arithmetic of the two kind::f16 modes (bf16x2 / f16x2), selected at compile time
template <int kMode>
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  if constexpr (kMode == PREC_FP16) return pack_f16x2(lo, hi);
  else return pack_bf16x2(lo, hi);
}
template <int kMode>
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
  uint32_t r;
  if constexpr (kMode == PREC_FP16) asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  else asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
template <int kMode>
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) {
  uint32_t r;
  if constexpr (kMode == PREC_FP16) asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  else asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
template <int kMode>
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  if constexpr (kMode == PREC_FP16) asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  else asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
template <int kMode>
__device__ __forceinline__ uint32_t htanh2(uint32_t a) {
  uint32_t r;
  if constexpr (kMode == PREC_FP16) asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a));
  else asm("tanh.approx.bf16x2 %0, %1;" : "=r"(r) : "r"(a));
  return r;
}
// SiLU of a packed, halved and ACT-scaled pre-activation pair: h = ACT * x / 2 -> ACT * x * sigmoid(x) = h + h * tanh(h / ACT).
// fp16: h / ACT may overflow to +-inf, for which MUFU.TANH returns +-1 -- exactly the saturated value.
template <int kMode>
__device__ __forceinline__ uint32_t hsilu2(uint32_t h2) {
  uint32_t arg = h2;
  if constexpr (kMode == PREC_FP16) arg = hmul2<kMode>(h2, 0x54005400u);  // x 64 = 1 / ACT (f16 64.0 = 0x5400)
  const uint32_t t2 = htanh2<kMode>(arg);
  return hfma2<kMode>(h2, t2, h2);
}
template <int kMode>
__device__ __forceinline__ float h2_lo(uint32_t v) {
  if constexpr (kMode == PREC_FP16) return __low2float(*reinterpret_cast<const __half2*>(&v));
  else return __uint_as_float(v << 16);
}
template <int kMode>
__device__ __forceinline__ float h2_hi(uint32_t v) {
  if constexpr (kMode == PREC_FP16) return __high2float(*reinterpret_cast<const __half2*>(&v));
  else return __uint_as_float(v & 0xffff0000u);
}
// bits of the 16-bit value 1.0 * ACT (the constant column that carries the folded bias of the second edge layer)
template <int kMode>
__device__ __forceinline__ constexpr uint32_t h_act_one_bits() { return kMode == PREC_FP16 ? 0x2400u : 0x3f80u; }
template <int kMode>
__device__ __forceinline__ void store_h(void* dst, float v) {
  if constexpr (kMode == PREC_FP16) *reinterpret_cast<__half*>(dst) = __float2half_rn(v);
  else *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------------------------
// shared-memory addresses, mbarriers, proxies
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Wait for the phase with the given parity to complete.  A protocol bug must not hang the GPU: after ~4 s of
// spinning the kernel traps, which surfaces as a CUDA error on the host.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 1023u) == 1023u) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) {
        printf("mlcg: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// thread-block clusters (CTA pairs): remote mbarrier arrives, cluster barrier
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // relaxed: the data handed over lives in TMEM (ordered by tcgen05.fence / wait::st) or goes through
  // fence.proxy.async; a cluster-scope release would add a full cluster fence (L1 flush) to every arrival.
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire (the arrivals may come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 1023u) == 1023u) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) {
        printf("mlcg: cluster mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// CTA-pair (cta_group::2) variants: executed by the same warp of both CTAs of the pair
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in every CTA of `mask` once all previously issued
// tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major, SWIZZLE_128B operand whose rows are 128 bytes (one swizzle atom
// wide in K): 8-row groups are 1024 bytes apart (SBO); LBO is unused for swizzled K-major; version = 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                        // leading byte offset (unused)
  d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B
  return d;
}

// MN-major, SWIZZLE_128B operand: 64 MN-contiguous elements (128 B) per K row, 8-row K groups `1024 B` apart (SBO),
// further 64-element MN blocks `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor: D = fp32, A/B K-major, dense.  fmt: 1 = BF16 (kind::f16), 2 = TF32 (kind::tf32)
__host__ __device__ constexpr uint32_t umma_idesc(int fmt, int m, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

template <int kMode>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (is16(kMode)) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// TMEM -> registers / registers -> TMEM: 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: this warp's 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// Operand-format stores: activations that feed a tensor-core GEMM live in HBM as [m_tile][k_chunk][128 x 128 B]
// blocks that are byte-for-byte the SWIZZLE_128B shared-memory image, so a stage is one contiguous bulk copy.
// ---------------------------------------------------------------------------------------------------------------
template <int kMode>
__device__ __forceinline__ size_t op_tile_bytes(int n_chunks) { return (size_t)n_chunks * A_CHUNK_BYTES; }

// store `kCount` (multiple of epp) consecutive K elements of operand row `grow`, starting at K index `k0`
// (k0 % epp == 0), into an operand-format array with `n_chunks` chunks per tile.
template <int kMode, int kCount>
__device__ __forceinline__ void op_store(uint8_t* base, int n_chunks, int grow, int k0, const float* v) {
  constexpr int EPC = epc(kMode), EPP = epp(kMode);
  const int mt = grow >> 7, r = grow & 127;
#pragma unroll
  for (int e = 0; e < kCount; e += EPP) {
    const int k = k0 + e;
    const int kc = k / EPC, p = (k % EPC) / EPP;
    uint8_t* dst = base + ((size_t)mt * n_chunks + kc) * A_CHUNK_BYTES + sw128_offset(r, p);
    uint4 w;
    if constexpr (is16(kMode)) {
      w.x = pack_h2<kMode>(v[e + 0], v[e + 1]);
      w.y = pack_h2<kMode>(v[e + 2], v[e + 3]);
      w.z = pack_h2<kMode>(v[e + 4], v[e + 5]);
      w.w = pack_h2<kMode>(v[e + 6], v[e + 7]);
    } else {
      w.x = f32_to_tf32(v[e + 0]);
      w.y = f32_to_tf32(v[e + 1]);
      w.z = f32_to_tf32(v[e + 2]);
      w.w = f32_to_tf32(v[e + 3]);
    }
    *reinterpret_cast<uint4*>(dst) = w;
  }
}
// single element store (used by the segment-sum readers, one column per lane)
template <int kMode>
__device__ __forceinline__ void op_store1(uint8_t* base, int n_chunks, int grow, int k, float v) {
  constexpr int EPC = epc(kMode), EPP = epp(kMode);
  const int mt = grow >> 7, r = grow & 127;
  const int kc = k / EPC, within = k % EPC;
  uint8_t* dst = base + ((size_t)mt * n_chunks + kc) * A_CHUNK_BYTES + sw128_offset(r, within / EPP);
  if constexpr (is16(kMode)) {
    store_h<kMode>(dst + 2 * (within % EPP), v);
  } else {
    reinterpret_cast<uint32_t*>(dst)[within % EPP] = f32_to_tf32(v);
  }
}

}  // namespace mlcg
