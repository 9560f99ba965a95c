// k_tc_edge3: the fused all-pairs edge MLP of one EGNN sub-layer with a RESIDENT A operand and two ping-pong
// accumulators (16-bit tensor-core modes, CTA pairs).  Same contract, tile table, weights and outputs as k_tc_edge
// (mlcg_tc.cuh; reference egnn.py:38-52 + 418-437 for GCL, egnn.py:111-135 for EquivariantUpdate); what changes is how
// the 512 TMEM columns are spent, and with it the schedule.
//
// k_tc_edge keeps ONE 128 x 448 fp32 accumulator (448 columns) + a 2-stage A ring (64 columns): the tile is a chain
//   main MMAs -> pass 1 (SiLU, MUFU bound, tensor core idle) -> segment-sum MMAs -> readout -> next tile's MMAs,
// and the A generation is coupled to the MMAs chunk by chunk through the 2-stage ring.
//
// Here the N = 432 (>= 420) output channels of the second layer are computed in three passes of N = 144 over the SAME A
// operand, which therefore stays in TMEM for the whole tile:
//   columns   0..223   A = SiLU(first layer), 128 rows x 448 K as packed 16-bit (7 chunks x 32 columns)
//   columns 224..367   accumulator 0 (128 x 144 fp32)
//   columns 368..511   accumulator 1
// The passes ("thirds") alternate between the two accumulators, so while the tensor core computes third g+1 the compute
// warps run pass 1 of third g; the A chunks of the NEXT tile are generated as soon as the last third has consumed the
// corresponding chunk of this one (per-chunk a_free barriers), i.e. the A generation of tile t+1 overlaps the last third
// and the segment sum of tile t and is never throttled by a ring.  Steady state per tile (C = the 16 compute warps,
// T = tensor core):
//   C: A-gen(t+1) chunks 2..6 | readout(t-1).. | pass1(t,0) | pass1(t,1) | A-gen(t+1) 0..1 | pass1(t,2) gate selector |
//   T: third 0 (paced by A-gen)  seg-sum(t-1)  | third 1    | third 2    |                 | ...
// The compute warps are busy all the time (MUFU / issue bound); the tensor core waits for them, not the other way round.
//
// The segment sum (GCL) stays on the tensor core: the messages of all three thirds are staged in shared memory in natural
// channel order ([channels x rows], MN-major A operand, 7 x 16 KB), the gate enters through the selector, and the result
// D2 lands in the accumulator that the tile's last third used (free until the next tile's second third needs it).
//
// Warp roles (18 warps per CTA, CTA pairs):
//   warp 0        bulk-copy producer (P/Q rows, 21 weight blocks per tile); in the leader CTA it also issues the 32
//                 segment-sum MMAs of a GCL tile -- one thread cannot issue them AND the thirds at the tensor core's rate
//   warp 1        leader CTA: tcgen05.mma issuer of the thirds (cta_group::2, M = 256); peer CTA: relays "my half of the
//                 weight block has landed" onto the leader's barrier
//   warps 2..17   compute: A generation, pass 1 per third, gate / selector / D2 readout or coordinate sums
// All three service loops run with the whole warp converged and issue their tcgen05 / bulk-copy instructions from one elected
// lane (elect.sync inside the asm); the equivariant variant multiplies the second third alongside the first while the A
// operand is generated (e3_blk) and sums the coordinate messages after the next tile's A generation.
// DESIGN.md 4.1b has the measurements behind each of these choices.
#pragma once
#include <type_traits>
#include <utility>
#include "mlcg_tc.cuh"

namespace mlcg {

constexpr int E3_NT = 144;                   // output channels per third (accumulator width)
constexpr int E3_NKC = 7;                    // K chunks of 64
constexpr int E3_DCOL0 = 224;                // first accumulator column (A occupies 0..223)
constexpr int E3_WSLOT = 72 * CHUNK_BYTES;   // 9,216 B: this CTA's 72 rows of one (K chunk, third) block of W2
constexpr int E3_NW_GCL = 5;                 // W ring slots next to the 112 KB message staging (GCL)
constexpr int E3_NW_EQ = 16;                 // W ring slots of the equivariant variant (no staging)
constexpr int E3_EARLY = 2;                  // chunks of the next tile generated before pass 1 of the last third

template <int kMode, bool kEquiv>
struct Edge3Smem {
  static constexpr int E3_NW = kEquiv ? E3_NW_EQ : E3_NW_GCL;
  static constexpr int NSTG = kEquiv ? 0 : 7;
  static constexpr int PQ_PITCH = 912;                                // bytes (== 16 mod 128: conflict-free LDS.128)
  static constexpr int PQ_ROW = HP * 2;                               // bytes copied per P / Q row
  static constexpr int W_OFF = 0;
  static constexpr int STG_OFF = W_OFF + E3_NW * E3_WSLOT;            // 7 x 16 KB message staging (MN-major, 64 channels each)
  static constexpr int SEL_OFF = STG_OFF + NSTG * A_CHUNK_BYTES;         // 4 KB selector + carried partial sums
  static constexpr int Q_OFF = SEL_OFF + 2 * 4096;
  static constexpr int P_OFF = Q_OFF + ((EDGE_MAXN * PQ_PITCH + 127) / 128) * 128;
  static constexpr int DOT_OFF = P_OFF + ((EDGE_MAXG * PQ_PITCH + 127) / 128) * 128;
  static constexpr int TRS_OFF = DOT_OFF + 4 * TILE_M * 4;
  static constexpr int RID_OFF = TRS_OFF + 2 * TILE_M * 3 * 4;
  static constexpr int RIG_OFF = RID_OFF + 2 * TILE_M * 8;
  static constexpr int WV_OFF = RIG_OFF + 2 * TILE_M * 4;
  static constexpr int WCD_OFF = WV_OFF + HP * 4;   // first-layer columns of the two distance features: wc[448] | wd[448] fp32, or
                                                   // the packed 16-bit table wcd_h[448]
  static constexpr int BAR_OFF = WCD_OFF + 2 * HP * 4;
  static constexpr int XT_OFF = BAR_OFF + 1024;    // barriers: 3 per W slot + 2 per K chunk + 10, then the TMEM slot.
                                                   // XT: [2][12 x 3] coordinates of the tile's targets (equivariant layer)
  static constexpr int PROF_OFF = XT_OFF + 512;   // 16 x int64 phase counters (diagnostics)
  static constexpr int TOTAL = PROF_OFF + 128;
  static constexpr int ALLOC = TOTAL + 1024;
  static_assert(STG_OFF % 1024 == 0 && SEL_OFF % 1024 == 0, "operand blocks must be 1024-byte aligned");
  static_assert(8 * (3 * E3_NW + 2 * E3_NKC + 10) + 8 <= 1024, "barrier area");
  static_assert(ALLOC <= 232448, "shared memory budget");
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// tcgen05 instructions for a convergent warp: one elected lane issues (kind::f16 covers bf16 and fp16; the operand
// format is in the instruction descriptor)
__device__ __forceinline__ void umma_ts_pair_e(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ss_pair_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_e(uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar),
      "h"(mask)
      : "memory");
}
// one elected lane of a convergent warp: expect_tx + 1-D bulk copy global -> shared
__device__ __forceinline__ void bulk_g2s_expect_e(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
      "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_e(uint32_t cluster_addr) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n\t}" ::"r"(cluster_addr)
      : "memory");
}
// mbarrier wait whose spin loop (with the timeout bookkeeping of mbar_wait) is out of line: the fast path is one try_wait
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  if (!done) mbar_wait_slow(bar, parity);
}
// the same for a convergent warp: the branch to the slow path is taken by all lanes or none (warp vote), which keeps the
// control flow -- and with it the loop variables of the caller -- uniform
__device__ __forceinline__ void mbar_wait_lean_w(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  if (!__all_sync(0xffffffffu, done != 0)) mbar_wait_slow(bar, parity);
}
// non-blocking test of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}

// Order in which the 21 (third, K chunk) blocks of a tile are streamed and multiplied; returns third * 8 + chunk.
//   sequential : third 0, third 1, third 2, chunks ascending.
//   interleaved: the second third runs alongside the first WHILE the A operand is being generated --
//                (0,0) (0,1) (1,0) (1,1) (0,2) (1,2) (0,3) (1,3) ... (0,6) (1,6), then third 2.  Both accumulators are then final
//                right after the last A chunk, and the tensor core has worked through the stretch in which it otherwise only
//                waits for A chunks.  Needs the second accumulator free by the time the tile's third A chunk is generated: true
//                for the equivariant variant (released after pass 1 of the previous tile's last third, which sits between the
//                second and the third chunk), not for GCL (it holds the segment sums until the readout at the end of the A
//                generation).
__host__ __device__ constexpr int e3_blk(bool interleave, int b) {
  if (!interleave || b >= 14) return (b / 7) * 8 + (b % 7);
  if (b < 2) return b;
  if (b < 4) return 8 + (b - 2);
  return (b & 1) ? 8 + b / 2 : b / 2;
}
template <class F, int... Bs>
__device__ __forceinline__ void e3_for_blocks(F&& f, std::integer_sequence<int, Bs...>) {
  (f(std::integral_constant<int, Bs>{}), ...);
}

// mbarrier slots of k_tc_edge3 (8 bytes each, from BAR_OFF)
template <int NW>
struct E3Bar {
  static constexpr int W_FULL = 0, W_EMPTY = NW, W_PEER = 2 * NW, A_FULL = 3 * NW, A_FREE = 3 * NW + E3_NKC, PQ_FULL = 3 * NW + 2 * E3_NKC,
                       PQ_EMPTY = PQ_FULL + 1, D_FULL = PQ_FULL + 2, D_FREE = PQ_FULL + 4, E_FULL = PQ_FULL + 6, E_DONE = PQ_FULL + 7,
                       TMEM_SLOT = PQ_FULL + 10;
};

// The tcgen05.mma issuer of k_tc_edge3: warp 1 of the leader CTA, all 32 lanes convergent, the tcgen05 instructions issued
// by one elected lane.  A straight-line tcgen05.mma costs ~45 issue cycles and an N = 144 MMA executes in 72
// (tools/probe_mma_rate2.cu), so the bookkeeping per MMA has to stay within a handful of instructions.  That only works if the
// compiler keeps descriptors, TMEM addresses and barrier addresses in UNIFORM registers: hence a separate (non-inlined)
// function whose every input is provably warp-uniform -- the tile count re-broadcast with a shuffle, the shared-memory base
// re-derived from the symbol, no calls inside the loops (a call spills the uniform registers), warp votes on every barrier test
// -- and whose register allocation is independent of the compute warps' code.  The K-chunk and k-step loops are unrolled
// (constant TMEM / descriptor offsets); a W slot is ONE barrier (the peer's relay arrives on the leader's w_full).
template <int kMode, bool kSeg, class S>
__device__ __forceinline__ void edge3_issuer(int n_iter_in) {
  constexpr int NW = S::E3_NW;
  using B = E3Bar<NW>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base + S::BAR_OFF;
  const int n_iter = __shfl_sync(0xffffffffu, n_iter_in, 0);
  constexpr uint32_t tmem_base = 0u;  // checked by the kernel
  constexpr uint32_t idesc = umma_idesc(umma_fmt(kMode), 2 * TILE_M, E3_NT);
  // (the segment-sum MMAs of the GCL variant are issued by warp 0, see edge3_issue_seg: they are independent of the thirds
  // -- other accumulator, released by d_free -- and this warp is the bottleneck of the tensor pipe)
  const uint64_t w_desc0 = umma_desc_sw128(base + S::W_OFF);
  uint32_t ws = 0, wph = 0;  // current W ring slot and its phase parity
  for (int it = 0; it < n_iter; ++it) {
    e3_for_blocks(
        [&](auto bc) {
          constexpr int blk = e3_blk(!kSeg, decltype(bc)::value), n3 = blk >> 3, kc = blk & 7;
          const int G = 3 * it + n3, acc = G & 1, u = G >> 1;
          // first block of a third: its accumulator has been read out (pass 1 / segment-sum readout of its previous content)
          if (kc == 0 && G >= 2) mbar_wait_lean_w(bar0 + 8u * (B::D_FREE + acc), (uint32_t)((u - 1) & 1));
          // first third: the A chunk has just been generated; the later thirds reuse it
          if (n3 == 0) mbar_wait_lean_w(bar0 + 8u * (B::A_FULL + kc), (uint32_t)(it & 1));
          mbar_wait_lean_w(bar0 + 8u * (B::W_FULL + ws), wph);  // both halves of the block: own bulk copy + the peer's relay
          tc_fence_after();
          const uint32_t dcol = tmem_base + E3_DCOL0 + acc * E3_NT;
          const uint64_t bdesc = w_desc0 + (uint64_t)(ws * (E3_WSLOT >> 4));
          // the last chunk's fourth k-step (K 432..447) is all padding: A and W2 are zero there
#pragma unroll
          for (int ks = 0; ks < ((kc == E3_NKC - 1) ? 3 : 4); ++ks)
            umma_ts_pair_e(dcol, tmem_base + kc * 32 + ks * 8, bdesc + 2 * ks, idesc, (kc | ks) != 0);
          umma_commit_pair_e(bar0 + 8u * (B::W_EMPTY + ws), 3);
          if (n3 == 2) umma_commit_pair_e(bar0 + 8u * (B::A_FREE + kc), 3);  // the next tile's A chunk kc may be written
          if (kc == E3_NKC - 1) umma_commit_pair_e(bar0 + 8u * (B::D_FULL + acc), 3);
          if (++ws == NW) { ws = 0; wph ^= 1u; }
        },
        std::make_integer_sequence<int, 3 * E3_NKC>{});
  }
}

// Segment sum of tile j on the tensor core (GCL): D2[128 channels x 32 (16 groups of CTA 0 | 16 of CTA 1)] per 128-channel
// block = E^T . S'^T with E the staged messages (MN-major A operand) and S' the gate-weighted selector, into the accumulator
// that tile j's last third used.  Issued by warp 0 of the leader CTA (convergent, one elected lane).
template <int kMode, class S>
__device__ __forceinline__ void edge3_issue_seg(uint32_t base, int j) {
  using B = E3Bar<S::E3_NW>;
  constexpr uint32_t idesc2 = umma_idesc(umma_fmt(kMode), 2 * TILE_M, 32) | (1u << 15);  // A is MN-major
  const uint64_t seg_a0 = umma_desc_mn_sw128(base + S::STG_OFF, A_CHUNK_BYTES);
  const uint64_t seg_b0 = umma_desc_sw128(base + S::SEL_OFF);
  const uint32_t dcol = E3_DCOL0 + ((3 * j + 2) & 1) * E3_NT;
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
      umma_ss_pair_e(dcol + cb * 32, seg_a0 + (uint64_t)((cb * 2 * A_CHUNK_BYTES + ks * 2048) >> 4),
                     seg_b0 + (uint64_t)(((ks >> 2) * 2048) >> 4) + 2 * (ks & 3), idesc2, ks != 0);
  }
  umma_commit_pair_e(base + S::BAR_OFF + 8u * B::E_DONE, 3);
}

// kProf: phase cycle counters (mlcg_edge_phase_profile): thread ct == 0 of every CTA accumulates [0] wait third 0, [1] pass 1
// of third 0, [2] wait third 1, [3] pass 1 of third 1, [4] barrier + P/Q wait, [5] early A chunks of the next tile, [6] tiles,
// [7] wait third 2, [8] pass 1 of third 2, [9] gate + selector + publish (equivariant: remaining A chunks of the next tile), [10]
// remaining A chunks of the next tile (equivariant: coordinate sums), [11] wait for the segment-sum MMAs, [12] readout, [13] end-of-tile barrier; [14], [15] unused (the
// service warps carry no counters: their instruction streams are the thing to keep short).
template <int kMode, bool kEquiv, bool kDistF32, bool kProf = false>
__global__ void __launch_bounds__(EDGE_THREADS, 1) k_tc_edge3(const __grid_constant__ EdgeArgs p) {
  static_assert(is16(kMode), "k_tc_edge3 serves the 16-bit tensor-core modes");
  constexpr int EPC = 64, ELEMS = 16;
  constexpr bool kSeg = !kEquiv;
  constexpr float kDistScale = dist_scale(kMode);
  using S = Edge3Smem<kMode, kEquiv>;
  constexpr int E3_NW = S::E3_NW;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint8_t* Qs = gbase + S::Q_OFF;
  const uint8_t* Ps = gbase + S::P_OFF;
  float* dots = reinterpret_cast<float*>(gbase + S::DOT_OFF);
  float* trs_all = reinterpret_cast<float*>(gbase + S::TRS_OFF);
  float2* ri_d_all = reinterpret_cast<float2*>(gbase + S::RID_OFF);
  int* ri_gj_all = reinterpret_cast<int*>(gbase + S::RIG_OFF);
  float* wv_s = reinterpret_cast<float*>(gbase + S::WV_OFF);
  [[maybe_unused]] float* xt_all = reinterpret_cast<float*>(gbase + S::XT_OFF);
  const uint32_t bar0 = base + S::BAR_OFF;
  auto w_full = [&](int s) { return bar0 + 8u * s; };
  auto w_empty = [&](int s) { return bar0 + 8u * (E3_NW + s); };
  // (slots 2 NW .. 3 NW - 1 are unused: the peer's relay arrives on the leader's w_full, whose count is 2)
  auto a_full = [&](int kc) { return bar0 + 8u * (3 * E3_NW + kc); };      // leader: A chunk kc of the current tile is in TMEM
  auto a_free = [&](int kc) { return bar0 + 8u * (3 * E3_NW + E3_NKC + kc); };  // the last third has consumed A chunk kc
  const uint32_t pq_full = bar0 + 8u * (3 * E3_NW + 2 * E3_NKC);
  const uint32_t pq_empty = pq_full + 8u;
  auto d_full = [&](int a) { return pq_full + 16u + 8u * a; };             // accumulator a holds a finished third
  auto d_free = [&](int a) { return pq_full + 32u + 8u * a; };             // leader: accumulator a has been read out
  const uint32_t e_full = pq_full + 48u;                                   // leader: messages staged + selector built
  const uint32_t e_done = pq_full + 56u;                                   // segment-sum MMAs complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + S::BAR_OFF + 8 * (3 * E3_NW + 2 * E3_NKC + 10));

  // warp index through a broadcast shuffle: tells the compiler that it is warp-uniform, so the role branches below are
  // uniform branches and the service warps' loop variables / descriptors can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // tile range of the pair: identical to k_tc_edge<.., kPair = true> (edge_tile_owner on the host mirrors it)
  const uint32_t crank = (uint32_t)__shfl_sync(0xffffffffu, (int)cluster_ctarank(), 0);  // (uniform for the compiler, see above)
  int t_begin, t_end, n_iter;
  {
    const int npairs = gridDim.x >> 1, pr = blockIdx.x >> 1;
    const int T0 = (int)(((long long)pr * p.n_tiles) / npairs), T1 = (int)(((long long)(pr + 1) * p.n_tiles) / npairs);
    n_iter = (T1 - T0 + 1) >> 1;
    t_begin = crank == 0 ? T0 : T0 + n_iter;
    t_end = crank == 0 ? T0 + n_iter : T1;
  }
  auto fetch_tile = [&](int it) -> EdgeTile {
    const int t = t_begin + it;
    const int tt = t < t_end ? t : max(t_end - 1, 0);
    const int4 a = p.tiles[2 * tt], b = p.tiles[2 * tt + 1];
    EdgeTile e{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (t >= t_end) {  // ghost: same molecule, no rows, no target atoms
      e.nrows = 0;
      e.ng = 0;
      e.fixa = e.fixb = -1;
    }
    return e;
  };
  constexpr int NARR = 2 * (EDGE_CT / 32);  // arrivals (one per compute warp of both CTAs) on the leader's barriers

  if (threadIdx.x == 0) {
    for (int s = 0; s < E3_NW; ++s) {
      mbar_init(w_full(s), crank == 0 ? 2 : 1);  // leader: own producer (expect_tx) + the peer's relay
      mbar_init(w_empty(s), 1);
    }
    for (int kc = 0; kc < E3_NKC; ++kc) {
      mbar_init(a_full(kc), NARR);
      mbar_init(a_free(kc), 1);
    }
    mbar_init(pq_full, 1);
    mbar_init(pq_empty, EDGE_CT / 32);
    for (int a = 0; a < 2; ++a) {
      mbar_init(d_full(a), 1);
      mbar_init(d_free(a), NARR);
    }
    mbar_init(e_full, NARR);
    mbar_init(e_done, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < HP; i += EDGE_THREADS) wv_s[i] = p.wv[i] * (1.0f / act_scale(kMode));
  // The distance-feature columns are read once per K chunk by every thread with a warp-uniform index.  From the constant
  // bank that is an indexed LDC, which measures ~8 issue cycles per warp instruction (tools/probe_alu.cu) -- as much as the
  // MUFU work of the chunk; from shared memory it is a broadcast LDS.128 (one wavefront).
  float* wc_s = reinterpret_cast<float*>(gbase + S::WCD_OFF);
  float* wd_s = wc_s + HP;
  uint32_t* wcdh_s = reinterpret_cast<uint32_t*>(gbase + S::WCD_OFF);
  if constexpr (kDistF32) {
    for (int i = threadIdx.x; i < HP; i += EDGE_THREADS) { wc_s[i] = p.wc[i]; wd_s[i] = p.wd[i]; }
  } else {
    for (int i = threadIdx.x; i < HP; i += EDGE_THREADS) wcdh_s[i] = p.wcd_h[i];
  }
  if (warp == 1) tmem_alloc_pair<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  // The kernel owns the SM (one CTA per SM, all 512 columns): the allocation starts at lane 0, column 0.  Relying on it
  // makes every TMEM address of the MMA issuer a compile-time constant (uniform registers, no per-MMA register moves).
  if (*tmem_slot != 0u) {
    if (threadIdx.x == 0) printf("mlcg: k_tc_edge3 expects the TMEM allocation at address 0, got 0x%x\n", *tmem_slot);
    __trap();
  }
  constexpr uint32_t tmem_base = 0u;
  auto arrive_leader = [&](uint32_t bar) { mbar_arrive_cluster(mapa(bar, 0)); };

  if (warp == 0) {
    // ===================== bulk-copy producer: P/Q rows per tile, W2 blocks per (third, K chunk) =====================
    // The whole warp runs the loop convergently.  W2 blocks: one elected lane per block (21 per tile; the loop is unrolled so
    // that the source offsets are constants and the slot / phase are running variables -- at ~290 cycles per block in the
    // full-rate thirds a scalar loop with index arithmetic cannot keep up).  P/Q rows: lane l issues row l, so the 12 (+ 39
    // on a molecule change) row copies of a tile cost a few hundred cycles of this warp instead of thousands.
    int prev_mol = -1;
    auto load_pq = [&](int j) {
      const EdgeTile tj = fetch_tile(j);
      const int node0 = p.node_off[tj.mol];
      if (j > 0) mbar_wait_lean_w(pq_empty, (uint32_t)((j - 1) & 1));
      const bool newmol = (tj.mol != prev_mol);
      if (lane == 0) mbar_arrive_expect_tx(pq_full, (uint32_t)((tj.ng + (newmol ? tj.n : 0)) * S::PQ_ROW));
      __syncwarp();
      const uint8_t* pqb = reinterpret_cast<const uint8_t*>(p.pq);
      if (lane < tj.ng)
        bulk_g2s(base + S::P_OFF + lane * S::PQ_PITCH, pqb + (size_t)(node0 + tj.i0 + lane) * (2 * S::PQ_ROW), S::PQ_ROW, pq_full);
      if (newmol)
        for (int j2 = lane; j2 < tj.n; j2 += 32)
          bulk_g2s(base + S::Q_OFF + j2 * S::PQ_PITCH, pqb + (size_t)(node0 + j2) * (2 * S::PQ_ROW) + S::PQ_ROW, S::PQ_ROW, pq_full);
      __syncwarp();
      prev_mol = tj.mol;
    };
    if (n_iter > 0) load_pq(0);
    uint32_t ws = 0, wph = 1;  // slot and the parity of "empty" to wait for (first pass: slots start empty)
    const uint8_t* wsrc = p.w2 + (size_t)(72 * crank) * CHUNK_BYTES;
    for (int it = 0; it < n_iter; ++it) {
      e3_for_blocks(
          [&](auto bc) {
          constexpr int blk = e3_blk(!kSeg, decltype(bc)::value), n3 = blk >> 3, kc = blk & 7;
          // the next tile's P/Q rows: requested once the second third is under way (the A generation of this tile is
          // complete by then, so pq_empty does not block the weight stream), needed one third later
          // (interleaved order with its deep ring: this warp runs up to 16 blocks ahead of the MMAs, so the request sits in the
          // last third -- earlier it would wait for pq_empty with the second third's last blocks still unrequested)
          if (n3 == (kSeg ? 1 : 2) && kc == (kSeg ? 5 : 2) && it + 1 < n_iter) load_pq(it + 1);
          if constexpr (kSeg) {
            // the previous tile's segment sum: its messages are staged and its selector is built once the compute warps are
            // through pass 1 of its last third, i.e. while they generate this tile's A chunks 2..; the weight blocks up to
            // here cover that stretch
            if (n3 == 0 && kc == 5 && it > 0 && crank == 0) {
              mbar_wait_lean_w(e_full, (uint32_t)((it - 1) & 1));
              tc_fence_after();
              edge3_issue_seg<kMode, S>(base, it - 1);
            }
          }
          mbar_wait_lean_w(w_empty(ws), wph);
          bulk_g2s_expect_e(base + S::W_OFF + ws * E3_WSLOT, wsrc + (size_t)kc * (HP * CHUNK_BYTES) + (size_t)(E3_NT * n3) * CHUNK_BYTES,
                            E3_WSLOT, w_full(ws));
          if (++ws == E3_NW) { ws = 0; wph ^= 1u; }
          },
          std::make_integer_sequence<int, 3 * E3_NKC>{});
    }
    if constexpr (kSeg) {
      if (crank == 0 && n_iter > 0) {  // the last tile's segment sum
        mbar_wait_lean_w(e_full, (uint32_t)((n_iter - 1) & 1));
        tc_fence_after();
        edge3_issue_seg<kMode, S>(base, n_iter - 1);
      }
    }
  } else if (warp == 1) {
    if (crank == 1) {
      // peer CTA: relay "my half of W slot s has landed" to the leader, which issues the joint MMAs (convergent warp, one
      // elected lane arrives)
      const uint32_t total = (uint32_t)n_iter * 3u * E3_NKC;
      uint32_t ws = 0, wph = 0;
      for (uint32_t wi = 0; wi < total; ++wi) {
        mbar_wait_lean_w(w_full(ws), wph);
        mbar_arrive_cluster_e(mapa(w_full(ws), 0));
        if (++ws == E3_NW) { ws = 0; wph ^= 1u; }
      }
    } else {
      // ===================== tcgen05.mma issuer (leader CTA) =====================
      edge3_issuer<kMode, kSeg, S>(n_iter);
    }
  } else {
    // ===================== compute warps: A generation, pass 1 per third, gate / segment-sum readout =====================
    const int ct = threadIdx.x - 64;
    const int cw = ct >> 5;
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int qq = cw >> 2;          // quarter of each K chunk (A generation) / column group of each third (pass 1)
    const int r = q * 32 + lane;     // tile row = TMEM lane
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    // pass 1: 18 units of 8 columns per third; quarters 0,1 own 5 units, quarters 2,3 own 4 (the four quarters of a lane
    // quarter share one SM sub-partition, so the MUFU pipe sees the same total either way)
    const int pu0 = qq < 2 ? 5 * qq : 10 + 4 * (qq - 2);
    const int pun = qq < 2 ? 5 : 4;
    long long* pacc = reinterpret_cast<long long*>(gbase + S::PROF_OFF);  // only thread ct == 0 touches it
    const bool profiling = kProf && (p.prof != nullptr) && (ct == 0);
    if (profiling)
      for (int k = 0; k < 16; ++k) pacc[k] = 0;
    long long pc0 = 0;
    auto tick = [&](int k) {
      if (profiling) {
        const long long c = clock64();
        pacc[k] += c - pc0;
        pc0 = c;
      }
    };

    auto tile_setup = [&](const EdgeTile ti, int buf) {
      const int i0 = ti.i0, n = ti.n, off0 = ti.off0, nrows = ti.nrows;
      const int nm1 = max(n - 1, 1);
      const int node0 = p.node_off[ti.mol];
      if (ct < TILE_M) {
        const int rr = ct;
        const bool rvalid = rr < nrows;
        const int g = rvalid ? (off0 + rr) / nm1 : 0;
        const int jj = rvalid ? off0 + rr - g * nm1 : 0;
        const int i = i0 + g;
        const int j = rvalid ? jj + (jj >= i ? 1 : 0) : 0;
        const float* xi = p.x_cur + (size_t)(node0 + i) * 3;
        const float* xj = p.x_cur + (size_t)(node0 + j) * 3;
        const float dx = xi[0] - xj[0], dy = xi[1] - xj[1], dz = xi[2] - xj[2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        const float* yi = p.x0 + (size_t)(node0 + i) * 3;
        const float* yj = p.x0 + (size_t)(node0 + j) * 3;
        const float ex = yi[0] - yj[0], ey = yi[1] - yj[1], ez = yi[2] - yj[2];
        ri_d_all[buf * TILE_M + rr] = make_float2(d2, ex * ex + ey * ey + ez * ez);
        ri_gj_all[buf * TILE_M + rr] = g | (j << 8) | (rvalid ? 0x10000 : 0);
        if constexpr (kEquiv) {
          const float inv = 1.0f / sqrtf(d2 + 1e-8f);
          float* tr = trs_all + buf * TILE_M * 3 + rr * 3;
          tr[0] = dx * inv; tr[1] = dy * inv; tr[2] = dz * inv;
        }
      } else if constexpr (kEquiv) {
        // the targets' coordinates, for the update at the end of the tile: loaded a tile ahead so that the coordinate sums do
        // not end in an exposed global-memory round trip
        const int k = ct - TILE_M;
        if (k < ti.ng * 3) xt_all[buf * 48 + k] = p.x_cur[(size_t)(node0 + i0) * 3 + k];
      }
    };
    auto pack_dist = [&](float d) {
      const float ds = (kMode == PREC_FP16) ? fminf(d * kDistScale, 60000.0f) : d;
      return pack_h2<kMode>(ds, ds);
    };
    // per-row inputs of the A generation of one tile
    struct RowIn {
      const uint8_t* P;
      const uint8_t* Q;
      float2 rd;
      uint32_t d2h, d02h;
    };
    auto row_inputs = [&](int buf) {
      const int info = ri_gj_all[buf * TILE_M + r];
      RowIn ri;
      ri.rd = ri_d_all[buf * TILE_M + r];
      ri.P = Ps + (info & 0xff) * S::PQ_PITCH + qq * ELEMS * 2;          // invalid rows read row 0
      ri.Q = Qs + ((info >> 8) & 0xff) * S::PQ_PITCH + qq * ELEMS * 2;
      ri.d2h = pack_dist(ri.rd.x);
      ri.d02h = pack_dist(ri.rd.y);
      return ri;
    };
    // The chunk stored last is published (wait::st, fence, arrive on a_full) when the next one is ready to be stored, or by an
    // explicit flush at the end of a run of chunks: the store latency is off the critical path of the generation.
    int a_pend = -1;
    auto agen_flush = [&]() {
      if (a_pend >= 0) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(a_full(a_pend));
        a_pend = -1;
      }
    };
    // A generation of tile j, K chunks kc0 .. kc1-1: SiLU(P_i + Q_j + d2*wc + d02*wd) -> TMEM columns 32*kc + 8*qq .. +7.
    // Software-pipelined over the chunks: the pre-activations of chunk kc+1 (shared-memory / constant-bank loads, packed adds,
    // fp32 distance terms: FMA-pipe work) are computed in the same straight-line region as the activation of chunk kc (MUFU
    // work), so that the scheduler interleaves the two.  Without this all four warps of a sub-partition run load -> FMA ->
    // MUFU -> store in lock step (they synchronise on every tile) and the two pipes take turns idling.
    auto agen_pre = [&](int kc, const RowIn& ri, uint32_t* h) {
      const int k0 = kc * EPC + qq * ELEMS;  // warp-uniform.  K columns >= 420 are padding: P, Q, wc, wd are zero there
#pragma unroll
      for (int e = 0; e < ELEMS; e += 8) {
        const uint4 pw = *reinterpret_cast<const uint4*>(ri.P + (kc * EPC + e) * 2);
        const uint4 qw = *reinterpret_cast<const uint4*>(ri.Q + (kc * EPC + e) * 2);
        const uint32_t pa[4] = {pw.x, pw.y, pw.z, pw.w}, qa[4] = {qw.x, qw.y, qw.z, qw.w};
        [[maybe_unused]] uint32_t wcp[4], wdp[4];
        [[maybe_unused]] float wcv[8], wdv[8];
        if constexpr (kDistF32) {
          const float4 a0 = *reinterpret_cast<const float4*>(wc_s + k0 + e), a1 = *reinterpret_cast<const float4*>(wc_s + k0 + e + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(wd_s + k0 + e), b1 = *reinterpret_cast<const float4*>(wd_s + k0 + e + 4);
          wcv[0] = a0.x; wcv[1] = a0.y; wcv[2] = a0.z; wcv[3] = a0.w; wcv[4] = a1.x; wcv[5] = a1.y; wcv[6] = a1.z; wcv[7] = a1.w;
          wdv[0] = b0.x; wdv[1] = b0.y; wdv[2] = b0.z; wdv[3] = b0.w; wdv[4] = b1.x; wdv[5] = b1.y; wdv[6] = b1.z; wdv[7] = b1.w;
        } else {
          const uint4 c0 = *reinterpret_cast<const uint4*>(wcdh_s + (k0 + e));
          const uint4 c1 = *reinterpret_cast<const uint4*>(wcdh_s + (k0 + e) + 4);
          wcp[0] = c0.x; wcp[1] = c0.y; wcp[2] = c1.x; wcp[3] = c1.y;
          wdp[0] = c0.z; wdp[1] = c0.w; wdp[2] = c1.z; wdp[3] = c1.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t h2 = hadd2<kMode>(pa[i], qa[i]);
          if constexpr (kDistF32) {
            // (an fp32 SiLU before the packing -- two issue cycles per pair fewer on paper, the packed 16-bit operations
            // issue at half rate -- measured 1.5 % slower and no more accurate: rejected)
            const float h_lo = fmaf(ri.rd.y, wdv[2 * i], fmaf(ri.rd.x, wcv[2 * i], h2_lo<kMode>(h2)));
            const float h_hi = fmaf(ri.rd.y, wdv[2 * i + 1], fmaf(ri.rd.x, wcv[2 * i + 1], h2_hi<kMode>(h2)));
            h2 = pack_h2<kMode>(h_lo, h_hi);
          } else {
            h2 = hfma2<kMode>(ri.d2h, wcp[i], h2);
            h2 = hfma2<kMode>(ri.d02h, wdp[i], h2);
          }
          h[(e >> 1) + i] = h2;
        }
      }
    };
    auto agen_post = [&](int j, int kc, const uint32_t* h, auto skip_pad) {
      const int k0 = kc * EPC + qq * ELEMS;
      uint32_t w[8];
#pragma unroll
      for (int e = 0; e < ELEMS; e += 8) {
        // last chunk only: K columns >= 424 are all padding, skip their MUFU work (the activation of 0 is 0)
        if (decltype(skip_pad)::value && k0 + e >= 424) {
          w[(e >> 1) + 0] = w[(e >> 1) + 1] = w[(e >> 1) + 2] = w[(e >> 1) + 3] = 0u;
          continue;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) w[(e >> 1) + i] = hsilu2<kMode>(h[(e >> 1) + i]);
      }
      // constant-1 column carrying b2 (K index 420 = element 4 of the run that starts at 416): low half of word 2
      if (k0 == BIAS_COL - 4) w[2] = (w[2] & 0xffff0000u) | h_act_one_bits<kMode>();
      if (j > 0) mbar_wait(a_free(kc), (uint32_t)((j - 1) & 1));  // the previous tile's last third is done with this chunk
      tc_fence_after();
      agen_flush();  // publish the previous chunk: its tcgen05.st has had a whole chunk of math to complete
      tmem_st8(trow + kc * 32 + qq * 8, w);
      a_pend = kc;
    };
    auto agen_run = [&](int j, int kc0, int kc1, const RowIn& ri) {
      uint32_t h[8];
      agen_pre(kc0, ri, h);
#pragma unroll 1
      for (int kc = kc0; kc + 1 < kc1; ++kc) {
        uint32_t hn[8];
        agen_pre(kc + 1, ri, hn);
        agen_post(j, kc, h, std::false_type{});
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = hn[i];
      }
      agen_post(j, kc1 - 1, h, std::true_type{});
      agen_flush();
    };

    if (n_iter > 0) {
      tile_setup(fetch_tile(0), 0);
      named_bar_sync(1, EDGE_CT);
      const RowIn ri0 = row_inputs(0);
      mbar_wait(pq_full, 0u);
      agen_run(0, 0, E3_NKC, ri0);
      __syncwarp();
      if (lane == 0) mbar_arrive(pq_empty);
    }
    for (int it = 0; it < n_iter; ++it) {
      const EdgeTile ti = fetch_tile(it);
      const int i0 = ti.i0, ng = ti.ng, n = ti.n;
      const int nm1 = max(n - 1, 1);
      const int node0 = p.node_off[ti.mol];
      const int buf = it & 1;
      const bool has_next = it + 1 < n_iter;
      auto glo = [&](int g) { return max(g * nm1 - ti.off0, 0); };
      auto ghi = [&](int g) { return min((g + 1) * nm1 - ti.off0, ti.nrows); };
      auto gfix = [&](int g) { return (g == 0 && ti.fixa >= 0) ? ti.fixa : (g == ng - 1 && ti.fixb >= 0) ? ti.fixb : -1; };
      float* carry_wr = reinterpret_cast<float*>(gbase + S::SEL_OFF + 4096) + (it & 1) * 464;
      const float* carry_rd = reinterpret_cast<const float*>(gbase + S::SEL_OFF + 4096) + ((it & 1) ^ 1) * 464;
      auto carried_in = [&](int g) { return g == 0 && ti.fixa == EDGE_CARRY; };
      auto carried_out = [&](int g) { return g == ng - 1 && ti.fixb == EDGE_CARRY; };
      const bool valid = (ri_gj_all[buf * TILE_M + r] & 0x10000) != 0;
      if (has_next) tile_setup(fetch_tile(it + 1), buf ^ 1);

      float dotp[4] = {0.f, 0.f, 0.f, 0.f};
      if (profiling) pc0 = clock64();
      // ---- pass 1 of one third: m = SiLU(D), partial dot with the gate / coordinate vector, messages staged (GCL) ----
      auto pass1 = [&](auto n3c) {
        constexpr int n3 = decltype(n3c)::value;
        const int G = 3 * it + n3, acc = G & 1;
        mbar_wait(d_full(acc), (uint32_t)((G >> 1) & 1));
        tc_fence_after();
        tick(n3 == 0 ? 0 : n3 == 1 ? 2 : 7);
        const uint32_t dcol = trow + E3_DCOL0 + acc * E3_NT;
        if constexpr (kEquiv) {
          // Equivariant variant (registers to spare: no message staging): all of this thread's columns of the third at once
          // (40 = 16 + 16 + 8 or 32 = 16 + 16), ONE wait, and the accumulator goes back to the tensor core BEFORE the
          // activation work.  (The same in the GCL variant spills: 0.87 instead of 0.71 ms.)
          float va[16], vc[16], vt[8];
          const int pcol = pu0 * 8;
          tmem_ld16(dcol + pcol, va);
          tmem_ld16(dcol + pcol + 16, vc);
          if (pun == 5) tmem_ld8(dcol + pcol + 32, vt);  // warp-uniform
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(d_free(acc));
          auto unit8 = [&](const float* v, int col) {
            const int ch0 = E3_NT * n3 + col;  // first of this unit's 8 output channels
            if (ch0 < 424) {  // channels >= 424 are padding: accumulator and coordinate-head weight are zero
              const float4 wa = *reinterpret_cast<const float4*>(wv_s + ch0);
              const float4 wb = *reinterpret_cast<const float4*>(wv_s + ch0 + 4);
              const float wv8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
              if constexpr (kMode == PREC_FP16) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  float t;
                  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(v[e] * (1.0f / act_scale(kMode))));
                  dotp[e & 3] = fmaf(fmaf(v[e], t, v[e]), wv8[e], dotp[e & 3]);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint32_t m2 = hsilu2<kMode>(pack_h2<kMode>(v[2 * e], v[2 * e + 1]));
                  dotp[(2 * e) & 3] = fmaf(h2_lo<kMode>(m2), wv8[2 * e], dotp[(2 * e) & 3]);
                  dotp[(2 * e + 1) & 3] = fmaf(h2_hi<kMode>(m2), wv8[2 * e + 1], dotp[(2 * e + 1) & 3]);
                }
              }
            }
          };
          unit8(va, pcol);
          unit8(va + 8, pcol + 8);
          unit8(vc, pcol + 16);
          unit8(vc + 8, pcol + 24);
          if (pun == 5) unit8(vt, pcol + 32);
        } else {
          float vb[2][8];
          tmem_ld8(dcol + pu0 * 8, vb[0]);
          tmem_wait_ld();
  #pragma unroll
          for (int i = 0; i < 5; ++i) {
            if (i < pun) {  // warp-uniform
              const int col = (pu0 + i) * 8;
              const int ch0 = E3_NT * n3 + col;  // first of this unit's 8 output channels
              float* v = vb[i & 1];
              if (i + 1 < pun) tmem_ld8(dcol + col + 8, vb[(i + 1) & 1]);
              if (ch0 < 424) {  // channels >= 424 are padding: accumulator, message and gate weight are zero
                const float4 wa = *reinterpret_cast<const float4*>(wv_s + ch0);
                const float4 wb = *reinterpret_cast<const float4*>(wv_s + ch0 + 4);
                const float wv8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                uint32_t mw[4];
                if constexpr (kMode == PREC_FP16) {
                  float mm[8];
  #pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    float t;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(v[e] * (1.0f / act_scale(kMode))));
                    mm[e] = fmaf(v[e], t, v[e]);
                    dotp[e & 3] = fmaf(mm[e], wv8[e], dotp[e & 3]);
                  }
  #pragma unroll
                  for (int e = 0; e < 4; ++e) mw[e] = pack_h2<kMode>(mm[2 * e], mm[2 * e + 1]);
                } else {
  #pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    mw[e] = hsilu2<kMode>(pack_h2<kMode>(v[2 * e], v[2 * e + 1]));
                    dotp[(2 * e) & 3] = fmaf(h2_lo<kMode>(mw[e]), wv8[2 * e], dotp[(2 * e) & 3]);
                    dotp[(2 * e + 1) & 3] = fmaf(h2_hi<kMode>(mw[e]), wv8[2 * e + 1], dotp[(2 * e + 1) & 3]);
                  }
                }
                if constexpr (kSeg) {
                  // natural channel order: chunk ch0/64 of the MN-major staging, 16-byte piece (ch0 % 64) / 8 of tile row r
                  *reinterpret_cast<uint4*>(gbase + S::STG_OFF + (ch0 >> 6) * A_CHUNK_BYTES + sw128_offset(r, (ch0 & 63) >> 3)) =
                      make_uint4(mw[0], mw[1], mw[2], mw[3]);
                }
              }
              if (i + 1 < pun) tmem_wait_ld();
            }
          }
          // this warp is done with the accumulator -- except the last third of a GCL tile, whose accumulator receives the
          // segment sums and is released after their readout
          if (!(kSeg && n3 == 2)) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader(d_free(acc));
          }
        }
        tick(n3 == 0 ? 1 : n3 == 1 ? 3 : 8);
      };
      pass1(std::integral_constant<int, 0>{});
      pass1(std::integral_constant<int, 1>{});
      named_bar_sync(1, EDGE_CT);  // the next tile's row metadata (tile_setup above) is visible
      RowIn rin{};
      if (has_next) {
        rin = row_inputs(buf ^ 1);
        mbar_wait(pq_full, (uint32_t)((it + 1) & 1));
        tick(4);
        agen_run(it + 1, 0, E3_EARLY, rin);
      }
      tick(5);
      pass1(std::integral_constant<int, 2>{});
      dots[qq * TILE_M + r] = (dotp[0] + dotp[1]) + (dotp[2] + dotp[3]);
      named_bar_sync(1, EDGE_CT);
      if constexpr (kEquiv) {
        // The next tile's remaining A chunks first, the coordinate sums after them: the sums are a short dependent chain in 9 of
        // the 16 warps; placed before the A generation, the other 7 warps ran ahead into it and the chain crawled behind them
        // (1.7 k cycles per tile), delaying A chunks that need all 16 warps.  dots / unit vectors / carry buffers are not
        // touched in between (next written after the end-of-tile barrier).  (One warp per target instead of eight threads per
        // (target, coordinate) measured 1 % slower.)
        if (has_next) {
          agen_run(it + 1, E3_EARLY, E3_NKC, rin);
          __syncwarp();
          if (lane == 0) mbar_arrive(pq_empty);
        }
        tick(9);
        // x_i += sum_j unit_ij * phi_ij / 100   (reference egnn.py:124-134).  Eight threads per (target, coordinate): each
        // sums every eighth row of the target's range (rows inside [glo, ghi) are valid by construction), then three
        // shuffle steps -- a fixed order, so the result is deterministic.
        const float* trs = trs_all + buf * TILE_M * 3;
        const int pc = ct >> 3, sub = ct & 7;
        const int gg = pc / 3, c = pc - gg * 3;
        float s = 0.f;
        if (gg < ng)
          for (int e = glo(gg) + sub; e < ghi(gg); e += 8) {
            const float phi = (dots[e] + dots[TILE_M + e]) + (dots[2 * TILE_M + e] + dots[3 * TILE_M + e]);
            s = fmaf(trs[e * 3 + c], phi, s);
          }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (gg < ng && sub == 0) {
          const int fx = gfix(gg);
          const size_t idx = (size_t)(node0 + i0 + gg) * 3 + c;
          if (carried_in(gg)) s = carry_rd[448 + c] + s;
          if (carried_out(gg)) carry_wr[448 + c] = s;
          else if (fx >= 0) atomicAdd(p.fix_dx + (size_t)fx * 4 + c, s);
          else p.x_next[idx] = xt_all[buf * 48 + gg * 3 + c] + s / 100.0f;
        }
        tick(10);
      } else {
        // e_ij = m_ij * sigmoid(w_a.m_ij + b_a); agg_i = sum_j e_ij / 100   (reference egnn.py:48-51, 59-64)
        const float full_dot = (dots[r] + dots[TILE_M + r]) + (dots[2 * TILE_M + r] + dots[3 * TILE_M + r]);
        const float gate = valid ? sigmoid_acc(full_dot + p.att_bias) : 0.f;
        float* gates = trs_all;  // [128]; the coordinate-message buffer is unused by GCL layers
        if (qq == 0) gates[r] = gate;
        named_bar_sync(1, EDGE_CT);  // all gates written; every thread's staging stores are issued
        if (ct >= 256) {
          // thread = (group g, 16-byte piece of 8 tile rows) of the K-major selector [16 groups x 128 rows]
          const int g = (ct - 256) >> 4, piece = ct & 15;
          const int lo_k = glo(g), hi_k = ghi(g);  // empty for g >= ng
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = piece * 8 + 2 * e;
            const float g0 = (k >= lo_k && k < hi_k) ? gates[k] : 0.f;
            const float g1 = (k + 1 >= lo_k && k + 1 < hi_k) ? gates[k + 1] : 0.f;
            w[e] = pack_h2<kMode>(g0, g1);
          }
          *reinterpret_cast<uint4*>(gbase + S::SEL_OFF + (piece >> 3) * 2048 + sw128_offset(g, piece & 7)) =
              make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async();  // staging + selector stores -> visible to the tensor core
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(e_full);
        tick(9);
        if (has_next) {
          agen_run(it + 1, E3_EARLY, E3_NKC, rin);
          __syncwarp();
          if (lane == 0) mbar_arrive(pq_empty);
        }
        tick(10);
        // D2 readout: lane r = channel 128*cb + r, columns = groups; quarter qq stores groups qq, qq+4, qq+8
        mbar_wait(e_done, (uint32_t)(it & 1));
        tc_fence_after();
        tick(11);
        const int aS = (3 * it + 2) & 1;
        const uint32_t d2col = trow + E3_DCOL0 + aS * E3_NT + crank * 16;
        float dval[4][3];
#pragma unroll
        for (int cb = 0; cb < 4; cb += 2) {
          float v[16], u[16];
          tmem_ld16(d2col + cb * 32, v);
          tmem_ld16(d2col + (cb + 1) * 32, u);
          tmem_wait_ld();
          dval[cb][0] = qq == 0 ? v[0] : qq == 1 ? v[1] : qq == 2 ? v[2] : v[3];
          dval[cb][1] = qq == 0 ? v[4] : qq == 1 ? v[5] : qq == 2 ? v[6] : v[7];
          dval[cb][2] = qq == 0 ? v[8] : qq == 1 ? v[9] : qq == 2 ? v[10] : v[11];
          dval[cb + 1][0] = qq == 0 ? u[0] : qq == 1 ? u[1] : qq == 2 ? u[2] : u[3];
          dval[cb + 1][1] = qq == 0 ? u[4] : qq == 1 ? u[5] : qq == 2 ? u[6] : u[7];
          dval[cb + 1][2] = qq == 0 ? u[8] : qq == 1 ? u[9] : qq == 2 ? u[10] : u[11];
        }
        // the accumulator is free: the next tile's second third may start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(d_free(aS));
        // channel ch = 128 cb + r of target node: operand chunk ch / 64 = 2 cb + r / 64, 16-byte piece (r / 8) % 8 (swizzled
        // by the node), element r % 8 -- everything but the chunk is independent of cb, so one base pointer per target
        const int r_piece = (r >> 3) & 7;
        const size_t r_off = (size_t)(r >> 6) * A_CHUNK_BYTES + (r & 7) * 2;
#pragma unroll
        for (int gi = 0; gi < 3; ++gi) {
          const int g = qq + 4 * gi;
          if (g < ng) {  // warp-uniform
            const int node = node0 + i0 + g;
            uint8_t* gdst = p.agg_op + (size_t)(node >> 7) * p.agg_chunks * A_CHUNK_BYTES + (node & 127) * 128 +
                            ((r_piece ^ (node & 7)) << 4) + r_off;
            const int fx = gfix(g);
            const bool cin = carried_in(g), cout = carried_out(g);
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
              const int ch = 128 * cb + r;
              if (ch < 424) {  // channels 420..423 are exact zeros; 424.. are never staged (padding)
                float val = dval[cb][gi];
                if (cin) val = carry_rd[ch] + val;
                if (cout) carry_wr[ch] = val;
                else if (fx >= 0) atomicAdd(p.fix_agg + (size_t)fx * HP + ch, val);
                else store_h<kMode>(gdst + (size_t)cb * 2 * A_CHUNK_BYTES, val * agg_out_scale(kMode));
              }
            }
          }
        }
      }
      tick(12);
      named_bar_sync(1, EDGE_CT);  // carry buffers, gates / coordinate messages and the staging are free for the next tile
      tick(13);
      if (profiling) pacc[6] += 1;
    }
    if (profiling)
      for (int k = 0; k < 14; ++k) p.prof[(size_t)blockIdx.x * 16 + k] = pacc[k];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's shared memory / TMEM through the joint MMAs
  if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace mlcg
