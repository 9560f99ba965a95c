// tcgen05 / TMEM kernels of the hot path (sm_100a only).
//
//  k_tc_gemm  : C[128-row tile, BN cols] = epilogue(A . W^T), persistent over the tile list.  A and W arrive as
//               pre-swizzled operand-format blocks through 1-D bulk async copies (UBLKCP) into a multi-stage mbarrier
//               ring; one thread issues tcgen05.mma with the fp32 accumulator in TMEM; eight epilogue warps read it back
//               with tcgen05.ld and move every 128-byte-per-row block through a swizzled staging block so that global
//               loads / stores are full lines.  Used for the per-node projections / node MLPs of the EGNN (reference
//               egnn.py:30-34,54-68) and for the AdjMatSeer linears (reference adj_mat_seer.py:53).
//  k_tc_edge  : the fused all-pairs edge MLP of one EGNN sub-layer (reference egnn.py:38-52 + 418-437 for GCL,
//               egnn.py:111-135 for EquivariantUpdate).  Edge tensors never touch HBM: the first-layer activation
//               SiLU(P_i + Q_j + d2*wc + d02*wd) is generated straight into an A-operand ring in TMEM, the 420x420
//               second layer runs on tcgen05 (TS mode, CTA pairs) with the accumulator in TMEM, and SiLU, attention
//               gate, masked neighbour sum (on the tensor core in bf16 mode) / coordinate update run in the epilogue.
#pragma once
#include "mlcg_common.cuh"

namespace mlcg {

// ---------------------------------------------------------------------------------------------------------------
// generic GEMM
// ---------------------------------------------------------------------------------------------------------------
enum GemmEpi { EPI_F32 = 0, EPI_SILU_OP = 1, EPI_RESID_OP = 2, EPI_BF16 = 3 };

struct GemmArgs {
  const uint8_t* a0;      // operand-format A, source 0
  const uint8_t* a1;      // operand-format A, source 1 (K chunks >= a0_chunks), may be null
  int a0_chunks;          // number of leading K chunks taken from a0
  int a0_per_tile;        // chunks per m-tile in a0
  int a1_per_tile;        // chunks per m-tile in a1
  int n_kc;               // total K chunks
  const uint8_t* w;       // packed weights [n_tile][n_kc][BN x 128 B]
  const float* bias;      // [n_tiles*BN]
  int m_rows;             // valid rows
  float* out_f32;         // EPI_F32: row-major fp32 output; EPI_BF16: the same pointer holds bf16 rows of ldo elements
  int ldo;
  int n_valid;            // EPI_F32: number of valid output columns
  const float* rowscale;  // EPI_F32: optional per-row multiplier of the bias (AdjMatSeer: rowsum of L)
  int relu;               // EPI_F32: apply ReLU
  uint8_t* out_op;        // EPI_SILU_OP / EPI_RESID_OP: operand-format output
  int out_op_chunks;
  float* resid;           // EPI_RESID_OP: fp32 residual stream, updated in place
  int ldr;
  float out_scale;        // fp16 mode: power-of-two scale of the 16-bit outputs (P/Q rows, operand-format rows); see
                          // mlcg_common.cuh "range management".  Ignored (1) in the other modes.
  int n_mtiles, n_ntiles; // tile grid (set by the launcher); the kernel is persistent and walks it with stride gridDim.x
  long long* prof;        // optional [grid][8] cycle counters (diagnostics): 0 producer waits for a free slot, 1 MMA issuer
                          // waits for operands, 2 MMA issuer waits for the epilogue, 3 epilogue waits for the accumulator,
                          // 4 epilogue proper, 5 tiles, 6 CTA lifetime
};

template <int BN>
struct GemmCfg {
  static constexpr int NSTAGE = (BN == 448) ? 3 : 4;
  static constexpr int NPER = (BN == 448) ? 224 : BN;  // N per tcgen05.mma (<= 256)
  static constexpr int NH = BN / NPER;
  static constexpr int STAGE_BYTES = A_CHUNK_BYTES + BN * CHUNK_BYTES;
  static constexpr int TMEM_COLS = (BN > 256) ? 512 : 256;
  static constexpr int BIAS_OFF = NSTAGE * STAGE_BYTES + 256;  // after the barriers: bias of the tile's BN columns, double-buffered
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * BN * 4;
};

// warp 0 = producer, warp 1 = MMA issuer, then 4 * GEMM_NSUB epilogue warps: warp w reads TMEM lanes 32*(w%4).. and every
// GEMM_NSUB-th 128-byte output block (the staged epilogue is latency bound per warp, so two warps share a lane quarter)
#ifndef MLCG_GEMM_NSUB
#define MLCG_GEMM_NSUB 2
#endif
constexpr int GEMM_NSUB = MLCG_GEMM_NSUB;
constexpr int GEMM_THREADS = 64 + 128 * GEMM_NSUB;

template <int kMode, int BN, int kEpi>
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_tc_gemm(const GemmArgs p) {
  using Cfg = GemmCfg<BN>;
  constexpr bool kFast = is16(kMode);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + Cfg::NSTAGE * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (Cfg::NSTAGE + s); };
  const uint32_t dfull_bar = bar0 + 8u * (2 * Cfg::NSTAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + Cfg::NSTAGE * Cfg::STAGE_BYTES + 8 * (2 * Cfg::NSTAGE + 2));

  const uint32_t epi_done = dfull_bar + 8u;   // all epilogue warps are done with the accumulator and the staging of a tile
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_mtiles * p.n_ntiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(dfull_bar, 1);
    mbar_init(epi_done, 4 * GEMM_NSUB);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Persistent: CTA b handles tiles b, b + gridDim.x, ... (tile = m-tile * n_ntiles + n-tile).  Chunk kc of a tile goes to
  // ring slot (kc + 1) % NSTAGE, so slot 0 -- whose memory doubles as the epilogue staging -- is the last one a tile
  // needs: while the epilogue of tile t runs, the producer already prefetches the first NSTAGE - 1 chunks of tile t + 1.
  // uses[s] counts how often slot s has been filled (mbarrier phase bookkeeping), kept in registers.
  if (warp == 0) {
    if (lane == 0) {
      uint32_t uses[Cfg::NSTAGE];
#pragma unroll
      for (int s = 0; s < Cfg::NSTAGE; ++s) uses[s] = 0;
      int tl = 0;
      long long w_slot = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
        const int mt = tile / p.n_ntiles, nt = tile - mt * p.n_ntiles;
        for (int kc = 0; kc < p.n_kc; ++kc) {
          const int sl = (kc + 1) % Cfg::NSTAGE;
          const long long c0 = p.prof ? clock64() : 0;
          if (tl > 0 && kc == Cfg::NSTAGE - 1) mbar_wait(epi_done, (uint32_t)((tl - 1) & 1));  // staging (slot 0) is free again
          uint32_t u = 0;
#pragma unroll
          for (int s = 0; s < Cfg::NSTAGE; ++s)
            if (s == sl) { u = uses[s]; uses[s]++; }
          mbar_wait(empty_bar(sl), (u & 1u) ^ 1u);
          if (p.prof) w_slot += clock64() - c0;
          mbar_arrive_expect_tx(full_bar(sl), Cfg::STAGE_BYTES);
          const uint8_t* asrc = (kc < p.a0_chunks)
                                    ? p.a0 + ((size_t)mt * p.a0_per_tile + kc) * A_CHUNK_BYTES
                                    : p.a1 + ((size_t)mt * p.a1_per_tile + (kc - p.a0_chunks)) * A_CHUNK_BYTES;
          const uint8_t* wsrc = p.w + ((size_t)nt * p.n_kc + kc) * (size_t)(BN * CHUNK_BYTES);
          const uint32_t sa = base + sl * Cfg::STAGE_BYTES;
          bulk_g2s(sa, asrc, A_CHUNK_BYTES, full_bar(sl));
          bulk_g2s(sa + A_CHUNK_BYTES, wsrc, BN * CHUNK_BYTES, full_bar(sl));
        }
      }
      if (p.prof) p.prof[(size_t)blockIdx.x * 8 + 0] = w_slot;
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(umma_fmt(kMode), TILE_M, Cfg::NPER);
      uint32_t uses[Cfg::NSTAGE];
#pragma unroll
      for (int s = 0; s < Cfg::NSTAGE; ++s) uses[s] = 0;
      int tl = 0;
      long long w_ops = 0, w_epi = 0;
      const long long t_start = p.prof ? clock64() : 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
        if (tl > 0) {  // the previous tile's accumulator has been read out
          const long long c0 = p.prof ? clock64() : 0;
          mbar_wait(epi_done, (uint32_t)((tl - 1) & 1));
          tc_fence_after();
          if (p.prof) w_epi += clock64() - c0;
        }
        for (int kc = 0; kc < p.n_kc; ++kc) {
          const int sl = (kc + 1) % Cfg::NSTAGE;
          uint32_t u = 0;
#pragma unroll
          for (int s = 0; s < Cfg::NSTAGE; ++s)
            if (s == sl) { u = uses[s]; uses[s]++; }
          const long long c1 = p.prof ? clock64() : 0;
          mbar_wait(full_bar(sl), u & 1u);
          tc_fence_after();
          if (p.prof) w_ops += clock64() - c1;
          const uint32_t sa = base + sl * Cfg::STAGE_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa);
#pragma unroll
          for (int nh = 0; nh < Cfg::NH; ++nh) {
            const uint64_t bdesc = umma_desc_sw128(sa + A_CHUNK_BYTES + nh * Cfg::NPER * CHUNK_BYTES);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma<kMode>(tmem_base + nh * Cfg::NPER, adesc + 2 * ks, bdesc + 2 * ks, idesc, (kc | ks) != 0);
          }
          umma_commit(empty_bar(sl));
        }
        umma_commit(dfull_bar);
      }
      if (p.prof) {
        p.prof[(size_t)blockIdx.x * 8 + 1] = w_ops;
        p.prof[(size_t)blockIdx.x * 8 + 2] = w_epi;
        p.prof[(size_t)blockIdx.x * 8 + 5] = tl;
        p.prof[(size_t)blockIdx.x * 8 + 6] = clock64() - t_start;
      }
    }
  } else {
   int tl = 0;
   long long w_acc = 0, t_epi = 0;
   for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
    const int mt = tile / p.n_ntiles, nt = tile - mt * p.n_ntiles;
    // bias of this tile's columns -> shared memory while the main loop runs (a global load inside the column loop would
    // expose its latency once per 32 columns).  Double-buffered by tile parity; the named barrier orders buffer reuse.
    float* bias_s = reinterpret_cast<float*>(gen_base + Cfg::BIAS_OFF) + (tl & 1) * BN;
    for (int i = threadIdx.x - 64; i < BN; i += GEMM_THREADS - 64) bias_s[i] = p.bias[nt * BN + i];
    named_bar_sync(1, GEMM_THREADS - 64);
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31; thread = one output row
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int grow = mt * TILE_M + row;
    const bool rvalid = grow < p.m_rows;
    const long long pc0 = p.prof ? clock64() : 0;
    mbar_wait(dfull_bar, (uint32_t)(tl & 1));
    tc_fence_after();
    const long long pc1 = p.prof ? clock64() : 0;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    if constexpr (kEpi == EPI_F32) {
      float rs = 1.0f;
      if (p.rowscale != nullptr && rvalid) rs = p.rowscale[grow];
      for (int c0 = ((warp - 2) >> 2) * 32; c0 < BN; c0 += 32 * GEMM_NSUB) {
        float v[32];
        tmem_ld32(trow + c0, v);
        tmem_wait_ld();
        const int gcol = nt * BN + c0;
        if (rvalid) {
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
          float* orow = p.out_f32 + (size_t)grow * p.ldo + gcol;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b = b4[e >> 2];
            float4 o;
            o.x = fmaf(rs, b.x, v[e + 0]);
            o.y = fmaf(rs, b.y, v[e + 1]);
            o.z = fmaf(rs, b.z, v[e + 2]);
            o.w = fmaf(rs, b.w, v[e + 3]);
            if (p.relu) {
              o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            }
            if (gcol + e + 3 < p.n_valid) {
              *reinterpret_cast<float4*>(orow + e) = o;
            } else {
              if (gcol + e + 0 < p.n_valid) orow[e + 0] = o.x;
              if (gcol + e + 1 < p.n_valid) orow[e + 1] = o.y;
              if (gcol + e + 2 < p.n_valid) orow[e + 2] = o.z;
            }
          }
        }
      }
    } else {
      // The accumulator is read with thread = row, but a warp-wide 16-byte store with thread = row touches 32 different
      // 128-byte lines.  Every 128-byte-per-row block therefore goes through a warp-private staging buffer (32 rows x
      // 128 B, 16-byte pieces XOR-swizzled by row & 7 = exactly the operand-format image of these rows) and is moved to /
      // from global memory with lane = (row 4i + lane/8, piece lane%8): 4 full 128-byte lines per instruction.  The
      // staging lives in the first ring stage, which is idle once the last MMA has completed.
      constexpr int EPC = epc(kMode);
      constexpr int UNIT = EPC / 32;             // 32-column accumulator groups per 128-byte output block
      constexpr int NBLK = BN / EPC;             // output blocks per row
      const int sub = (warp - 2) >> 2;           // this warp handles blocks sub, sub + GEMM_NSUB, ...
      const int n_my = ((NBLK - sub + GEMM_NSUB - 1) / GEMM_NSUB) * UNIT;  // its number of 32-column groups
      auto group_c0 = [&](int k) { return ((sub + GEMM_NSUB * (k / UNIT)) * UNIT + (k % UNIT)) * 32; };
      uint8_t* stg = gen_base + (sub * 4 + q) * 4096;
      const int crow = lane >> 3, cpiece = lane & 7;
      auto stg_at = [&](int r, int pc) { return stg + r * 128 + ((pc ^ (r & 7)) << 4); };
      const int row0 = mt * TILE_M + q * 32;  // first global row of this warp
      uint32_t ow[32];                        // packed words of the current 128-byte output block (thread's row)
      [[maybe_unused]] uint4 rnext[8];        // EPI_RESID_OP: prefetched residual block of the next column group
      if constexpr (kEpi == EPI_RESID_OP) {
        if (n_my > 0) {
          const uint4* nblk = reinterpret_cast<const uint4*>(p.resid + hres_index(row0, nt * BN + group_c0(0), 0));
#pragma unroll
          for (int i = 0; i < 8; ++i) rnext[i] = nblk[i * 32 + lane];
        }
      }
      for (int gk = 0; gk < n_my; ++gk) {
        const int c0 = group_c0(gk);
        float v[32];
        tmem_ld32(trow + c0, v);
        tmem_wait_ld();
        const int gcol = nt * BN + c0;
        const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
        if constexpr (kEpi == EPI_RESID_OP) {
          // residual block of the warp: 32 rows x 32 floats, contiguous 4 KB in the tiled layout; the block of the next
          // column group is already in flight (rnext), so the global-load latency is paid once per tile, not per group
          float* rblk = p.resid + hres_index(row0, gcol, 0);
#pragma unroll
          for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(stg_at(4 * i + crow, cpiece)) = rnext[i];
          if (gk + 1 < n_my) {
            const uint4* nblk = reinterpret_cast<const uint4*>(p.resid + hres_index(row0, nt * BN + group_c0(gk + 1), 0));
#pragma unroll
            for (int i = 0; i < 8; ++i) rnext[i] = nblk[i * 32 + lane];
          }
          __syncwarp();
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b = b4[e >> 2];
            float4 h = *reinterpret_cast<const float4*>(stg_at(lane, e >> 2));
            h.x += v[e + 0] + b.x;
            h.y += v[e + 1] + b.y;
            h.z += v[e + 2] + b.z;
            h.w += v[e + 3] + b.w;
            *reinterpret_cast<float4*>(stg_at(lane, e >> 2)) = h;
            v[e + 0] = h.x; v[e + 1] = h.y; v[e + 2] = h.z; v[e + 3] = h.w;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (row0 + 4 * i + crow < p.m_rows)
              reinterpret_cast<uint4*>(rblk)[i * 32 + lane] = *reinterpret_cast<const uint4*>(stg_at(4 * i + crow, cpiece));
          __syncwarp();
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b = b4[e >> 2];
            if constexpr (kEpi == EPI_SILU_OP) {
              v[e + 0] = silu<kFast>(v[e + 0] + b.x); v[e + 1] = silu<kFast>(v[e + 1] + b.y);
              v[e + 2] = silu<kFast>(v[e + 2] + b.z); v[e + 3] = silu<kFast>(v[e + 3] + b.w);
            } else {
              v[e + 0] += b.x; v[e + 1] += b.y; v[e + 2] += b.z; v[e + 3] += b.w;
            }
          }
        }
        if constexpr (kMode == PREC_FP16) {  // range management: exact power-of-two scale of the 16-bit copy
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] *= p.out_scale;
        }
        // pack into the 128-byte output block of this row
        const bool flush = (EPC == 32) || ((c0 & 32) != 0);
        if constexpr (EPC == 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) ow[j] = f32_to_tf32(v[j]);
        } else {
          if ((c0 & 32) == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) ow[j] = pack_h2<kMode>(v[2 * j], v[2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) ow[16 + j] = pack_h2<kMode>(v[2 * j], v[2 * j + 1]);
          }
        }
        if (flush) {
#pragma unroll
          for (int pc = 0; pc < 8; ++pc)
            *reinterpret_cast<uint4*>(stg_at(lane, pc)) = make_uint4(ow[4 * pc], ow[4 * pc + 1], ow[4 * pc + 2], ow[4 * pc + 3]);
          __syncwarp();
          const int cb0 = nt * BN + (c0 / EPC) * EPC;  // first channel of the block
          if constexpr (kEpi == EPI_BF16) {
            // row-major bf16 rows of ldo elements
            uint8_t* obase = reinterpret_cast<uint8_t*>(p.out_f32) + (size_t)cb0 * 2 + cpiece * 16;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + crow;
              if (row0 + r < p.m_rows)
                *reinterpret_cast<uint4*>(obase + (size_t)(row0 + r) * p.ldo * 2) = *reinterpret_cast<const uint4*>(stg_at(r, cpiece));
            }
          } else {
            // operand format: the staging image is the destination image (same swizzle): linear 4 KB copy
            uint8_t* dst = p.out_op + ((size_t)mt * p.out_op_chunks + cb0 / EPC) * A_CHUNK_BYTES + q * 4096;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (row0 + 4 * i + crow < p.m_rows)
                reinterpret_cast<uint4*>(dst)[i * 32 + lane] = reinterpret_cast<const uint4*>(stg)[i * 32 + lane];
          }
          __syncwarp();
        }
      }
    }
    // this warp is done with the tile's accumulator and staging
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(epi_done);
    if (p.prof) { w_acc += pc1 - pc0; t_epi += clock64() - pc1; }
   }
   if (p.prof && warp == 2 && lane == 0) {
     p.prof[(size_t)blockIdx.x * 8 + 3] = w_acc;
     p.prof[(size_t)blockIdx.x * 8 + 4] = t_epi;
   }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// fused edge kernel
// ---------------------------------------------------------------------------------------------------------------
// One CTA per SM, persistent over a contiguous range of tiles.  A tile = up to 128 consecutive rows of a molecule's
// target-major edge list (target i, then its neighbours j != i ascending), touching at most EDGE_MAXG targets; see
// EdgeTile below for how a target whose neighbour list is cut by a tile boundary is completed.
//
//   warp 0      bulk-copy producer: per tile the P rows of its targets + the molecule's Q rows (on a molecule change),
//               per K chunk the W2 block (pair mode: this CTA's 2 x 112 rows of it) into a 3-slot ring
//   warp 1      tcgen05.mma issuer, TS mode: A operand read from TMEM, B (W2) from shared memory, D in TMEM columns
//               0..447.  Pair mode (default): the leader CTA issues cta_group::2 MMAs with M = 256 for both CTAs' tiles,
//               the peer's warp 1 relays "my half of the W2 slot has landed".
//   warps 2-17  compute: 4 threads per tile row (= TMEM lane), each owning a quarter of the K / N range.
//     A generation  SiLU(P_i+Q_j+d2*wc+d02*wd) goes straight into a 2-stage A ring in TMEM (columns 448..511) with
//                   tcgen05.st -- the A operand never touches shared memory, which is the bandwidth-critical resource of
//                   this kernel (the SS-mode version spent ~200 KB of shared-memory traffic per K chunk and ran at the
//                   128 B/clk crossbar limit).  bf16 mode: packed bf16x2 arithmetic throughout.
//     pass 1        SiLU of the accumulator + dot with the gate / coordinate vector (thread-local per row); bf16 GCL also
//                   stages the messages in shared memory as the A operand of the segment-sum MMAs.
//     pass 2        GCL: gate, neighbour sum (bf16: on the tensor core with the gate folded into the selector; tf32: fp32
//                   shared-memory reader), output in operand format.  Equivariant layer: coordinate update.
//                   bf16: the first two A chunks of the NEXT tile are generated here, while the tensor core / the other
//                   warps finish this tile.
// The per-layer vectors wc, wd, wv live in the constant bank (kernel parameters): warp-uniform operands that cost no
// shared-memory bandwidth.
constexpr int EDGE_MAXG = 12;          // max target nodes (groups) per 128-row tile
constexpr int EDGE_MAXN = 39;          // max atoms per molecule (reference config.py MAX_N_NODES)
constexpr int EDGE_WSLOT = 224 * CHUNK_BYTES;  // 28,672 B: one (K chunk, N half) block of W2
constexpr int EDGE_NW = 3;             // W ring slots
constexpr int EDGE_NWMAX = EDGE_NW;
constexpr int EDGE_NA = 2;             // A ring stages (TMEM columns 448..479 and 480..511)
#ifndef MLCG_EDGE_EARLY
#define MLCG_EDGE_EARLY 2
#endif
constexpr int EDGE_EARLY = MLCG_EDGE_EARLY;  // K chunks of the next tile generated during the segment-sum MMAs (bf16 GCL), <= EDGE_NA
constexpr int EDGE_ACOL = 448;         // first TMEM column of the A ring
constexpr int EDGE_QPITCH = 452;       // floats; 1808 B rows -> conflict-free LDS.128 across consecutive j
constexpr int EDGE_CT = 512;           // compute threads (16 warps)
constexpr int EDGE_THREADS = 64 + EDGE_CT;  // warp 0 producer, warp 1 MMA, warps 2..17 compute

// Shared-memory layout.  bf16 mode stores the P/Q projections as bf16 (no accuracy cost: emulated and measured) which
// halves the A-generation loads and leaves room for two of the four segment-sum staging blocks (the other two borrow the
// W ring).  The second half of the selector area holds the carried partial sums of split targets.
template <int kMode>
struct EdgeSmemT {
  static constexpr bool BF = is16(kMode);
  static constexpr int PQ_ESIZE = BF ? 2 : 4;                       // bytes per P/Q element
  static constexpr int PQ_PITCH = BF ? 912 : EDGE_QPITCH * 4;       // row pitch in bytes (== 16 mod 128: conflict-free)
  static constexpr int PQ_ROW = HP * PQ_ESIZE;                      // bytes copied per row
  static constexpr int NSCR = BF ? 2 : 1;                           // 32 KB staging buffers
  static constexpr int W_OFF = 0;
  static constexpr int SCR_OFF = W_OFF + EDGE_NW * EDGE_WSLOT;
  static constexpr int SEL_OFF = SCR_OFF + NSCR * 2 * A_CHUNK_BYTES;   // 4 KB group selector S[16 x 128] (bf16)
  static constexpr int Q_OFF = SEL_OFF + 2 * 4096;   // selector is double-buffered (next tile's is built during the MMA tail)
  static constexpr int P_OFF = Q_OFF + ((EDGE_MAXN * PQ_PITCH + 127) / 128) * 128;
  static constexpr int DOT_OFF = P_OFF + ((EDGE_MAXG * PQ_PITCH + 127) / 128) * 128;  // [4][128] partial dots
  static constexpr int TRS_OFF = DOT_OFF + 4 * TILE_M * 4;      // [2][128][3] coordinate messages
  static constexpr int RID_OFF = TRS_OFF + 2 * TILE_M * 3 * 4;  // [2][128] float2 (d2, d0^2)
  static constexpr int RIG_OFF = RID_OFF + 2 * TILE_M * 8;      // [2][128] int  g | j<<8 | valid<<16
  static constexpr int WV_OFF = RIG_OFF + 2 * TILE_M * 4;       // [448] gate / coordinate-head vector
  static constexpr int PROF_OFF = WV_OFF + HP * 4;         // 16 x int64 phase counters (diagnostics)
  static constexpr int BAR_OFF = PROF_OFF + 128;
  static constexpr int TOTAL = BAR_OFF + 256;
  static constexpr int ALLOC = TOTAL + 1024;
  // bf16 GCL: staging of the tile's messages for the segment-sum MMAs, four 32 KB channel blocks (128 channels x 128 rows,
  // the last one half used).  Blocks 0,1 live in the scratch area, blocks 2,3 in the W ring, which is idle between the
  // tile's last main MMA and the next tile's first weight chunk.
  __host__ __device__ static constexpr int stage_off(int cb) { return cb < 2 ? SCR_OFF + cb * 2 * A_CHUNK_BYTES : W_OFF + (cb - 2) * 2 * A_CHUNK_BYTES; }
  static_assert(!BF || (2 * 2 * A_CHUNK_BYTES <= EDGE_NW * EDGE_WSLOT), "W ring too small for two staging blocks");
};

// One 128-row tile of the edge list of a molecule.  The n(n-1) real edges of a molecule are enumerated target-major
// (target i, then its n-1 neighbours in ascending order) and cut into ceil(n(n-1)/128) near-equal row ranges, so a tile
// may start and end in the middle of a target node's neighbour list.  Such a "split" target gets its neighbour sum from
// two tiles.  Normally both tiles are consecutive tiles of one CTA and the first partial sum is carried over in shared
// memory; at the ~150 boundaries between CTAs' tile ranges both tiles add their partial sum into a per-batch fp32 side
// buffer and k_edge_fixup turns it into the regular output afterwards.  Either way the result is first part + second part
// (fp32 addition is commutative, the side buffer starts at zero): it does not depend on the grid or on timing.
struct EdgeTile {
  int mol, off0, nrows, n;   // molecule; neighbours of the first target that precede this tile; rows; atoms
  int i0, ng, fixa, fixb;    // first target; targets touched; state of the first / last target (see below)
};
// fixa / fixb: EDGE_WHOLE = the target's neighbour list does not cross this end of the tile; EDGE_CARRY = it continues in
// the neighbouring tile and that tile is processed by the same CTA right before / after this one: the partial sum is
// handed over in shared memory; >= 0 = it continues in a tile of another CTA: side-buffer id for the atomic path.
constexpr int EDGE_WHOLE = -1, EDGE_CARRY = -2;
static_assert(sizeof(EdgeTile) == 32, "EdgeTile is fetched as two int4");

struct EdgeArgs {
  const int4* tiles;   // [n_tiles] EdgeTile
  int n_tiles;
  float* fix_agg;      // [n_fix][448] partial neighbour sums of split targets (GCL), zero between launches
  float* fix_dx;       // [n_fix][4]   partial coordinate updates of split targets (equivariant layer)
  const int* node_off; // [B+1] prefix sum of atom counts
  const void* pq;      // [nodes][896] (fp32, or bf16 in bf16 mode): P = W1a.h at 0..447, Q = W1b.h + b1 at 448..895
  const float* x_cur;  // [nodes][3] coordinates at block start
  const float* x0;     // [nodes][3] coordinates at EGNN input
  float* x_next;       // equivariant update output
  const uint8_t* w2;   // packed second-layer weights [n_kc][448 x 128 B] (bias folded into K column 420)
  int n_kc;
  float att_bias;
  uint8_t* agg_op;     // GCL: neighbour aggregate, operand format
  int agg_chunks;
  long long* prof;     // optional [grid][16] per-CTA phase cycle counters (diagnostics), or nullptr
  alignas(16) float wc[HP];  // first-layer column for d2 (W1[:,840]), zero padded     } constant bank
  alignas(16) float wd[HP];  // first-layer column for d0^2 (W1[:,841])                 }
  alignas(16) float wv[HP];  // attention vector (GCL) or coordinate head (equivariant) }
  alignas(16) uint32_t wcd_h[HP];  // bf16 mode: {wc[2k], wc[2k+1]} packed bf16x2 at [4*(k/2) + (k&1)],
                                   //            {wd[2k], wd[2k+1]} at [4*(k/2) + 2 + (k&1)]  (one 128-bit load = 4 channels of both)
};

// CTA-pair variants (cta_group::2): issued by the leader CTA only; M = 256 spans both CTAs' TMEM, B is split in N
// between the two CTAs' shared memories.
template <int kMode>
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (is16(kMode)) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_ss_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int kMode>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (is16(kMode)) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// kPair: the kernel is launched in clusters of 2 CTAs that share the W2 stream (each CTA holds half of every W2 block,
// tcgen05.mma.cta_group::2 with M = 256 covers both CTAs' tiles): halves the shared-memory traffic of the weight stream,
// which otherwise saturates the 128 B/clk crossbar on its own (64 B/clk of bulk-copy writes + 64 B/clk of operand reads).
// kDistF32 (bf16 mode only): evaluate the two distance terms of the first layer in fp32 instead of packed bf16x2.  About 8 %
// slower; only matters when d2 * wc dominates the pre-activation (worst teacher-forced step of the random-weight
// trajectory: eps rel-L2 6.5e-3 instead of 1.6e-2; typical inputs: no measurable difference).
// kProf: build the phase cycle counters in (mlcg_edge_phase_profile).  The production instantiations carry none of it: even
// predicated off, the counter code costs ~15 % of the issue slots of the A-generation loop.
template <int kMode, bool kEquiv, bool kPair, bool kDistF32 = false, bool kProf = false>
__global__ void __launch_bounds__(EDGE_THREADS, 1) k_tc_edge(const __grid_constant__ EdgeArgs p) {
  constexpr bool kFast = is16(kMode);
  constexpr int EPC = epc(kMode);
  constexpr int ELEMS = EPC / 4;  // K elements per thread per chunk (16 bf16 / fp16, 8 tf32) = 8 TMEM columns
  constexpr bool kSegMma = is16(kMode) && !kEquiv;  // neighbour sum on the tensor core (16-bit modes)
  constexpr bool kEarlyA = is16(kMode);  // first A chunks of tile t+1 are generated inside the epilogue of tile t
  constexpr float kDistScale = dist_scale(kMode);  // fp16: packed squared distances carry 2^-10 (inverse folded into wcd_h)
  using EdgeSmem = EdgeSmemT<kMode>;
  constexpr bool kPqBf16 = EdgeSmem::BF;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint8_t* Qs = gbase + EdgeSmem::Q_OFF;
  const uint8_t* Ps = gbase + EdgeSmem::P_OFF;
  float* dots = reinterpret_cast<float*>(gbase + EdgeSmem::DOT_OFF);
  float* trs_all = reinterpret_cast<float*>(gbase + EdgeSmem::TRS_OFF);
  float2* ri_d_all = reinterpret_cast<float2*>(gbase + EdgeSmem::RID_OFF);
  int* ri_gj_all = reinterpret_cast<int*>(gbase + EdgeSmem::RIG_OFF);
  float* wv_s = reinterpret_cast<float*>(gbase + EdgeSmem::WV_OFF);
  uint8_t* scratch = gbase + EdgeSmem::SCR_OFF;
  const uint32_t bar0 = base + EdgeSmem::BAR_OFF;
  auto w_full = [&](int s) { return bar0 + 8u * s; };
  auto w_empty = [&](int s) { return bar0 + 8u * (EDGE_NWMAX + s); };
  auto a_full = [&](int s) { return bar0 + 8u * (2 * EDGE_NWMAX + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 * EDGE_NWMAX + EDGE_NA + s); };
  const uint32_t pq_full = bar0 + 8u * (2 * EDGE_NWMAX + 2 * EDGE_NA);
  const uint32_t pq_empty = pq_full + 8u;
  const uint32_t d_full = pq_full + 16u;
  const uint32_t d_empty = pq_full + 24u;
  auto e_full = [&](int b) { return pq_full + 32u + 8u * b; };  // gated messages of a 128-channel block staged
  auto e_done = [&](int b) { return pq_full + 48u + 8u * b; };  // segment-sum MMA of that block complete
  auto w_peer = [&](int s) { return pq_full + 64u + 8u * s; };  // pair mode: the peer CTA's half of W slot s has landed
  const uint32_t d_half = e_done(1);  // accumulator columns 0..223 of the tile are final (one N half before d_full)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + EdgeSmem::BAR_OFF + 8 * (2 * EDGE_NWMAX + 2 * EDGE_NA + 8 + EDGE_NW));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Tile range.  Single-CTA mode: a contiguous range per CTA.  Pair mode: a contiguous range per pair, first half to
  // CTA 0 and second half to CTA 1; both CTAs run the same number of iterations (the MMAs are joint), CTA 1 pads with
  // an empty "ghost" tile when the range is odd.
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;
  int t_begin, t_end, n_iter;
  if constexpr (kPair) {
    const int npairs = gridDim.x >> 1, pr = blockIdx.x >> 1;
    const int T0 = (int)(((long long)pr * p.n_tiles) / npairs), T1 = (int)(((long long)(pr + 1) * p.n_tiles) / npairs);
    n_iter = (T1 - T0 + 1) >> 1;
    t_begin = crank == 0 ? T0 : T0 + n_iter;
    t_end = crank == 0 ? T0 + n_iter : T1;
  } else {
    t_begin = (int)(((long long)blockIdx.x * p.n_tiles) / gridDim.x);
    t_end = (int)(((long long)(blockIdx.x + 1) * p.n_tiles) / gridDim.x);
    n_iter = t_end - t_begin;
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
  constexpr int NARR = kPair ? 2 * (EDGE_CT / 32) : (EDGE_CT / 32);  // arrivals on the barriers the MMA issuer waits on

  if (threadIdx.x == 0) {
    for (int s = 0; s < EDGE_NWMAX; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
      mbar_init(w_peer(s), 1);
    }
    for (int s = 0; s < EDGE_NA; ++s) {
      mbar_init(a_full(s), NARR);
      mbar_init(a_empty(s), 1);
    }
    mbar_init(pq_full, 1);
    mbar_init(pq_empty, EDGE_CT / 32);
    mbar_init(d_full, 1);
    mbar_init(d_empty, NARR);
    for (int b = 0; b < 2; ++b) {
      mbar_init(e_full(b), NARR);
      mbar_init(e_done(b), 1);
    }
    fence_barrier_init();
  }
  // gate / coordinate-head vector; the messages it multiplies carry ACT (fp16 mode), undone here exactly
  for (int i = threadIdx.x; i < HP; i += EDGE_THREADS) wv_s[i] = p.wv[i] * (1.0f / act_scale(kMode));
  if (warp == 1) {
    if constexpr (kPair) tmem_alloc_pair<512>(smem_u32(tmem_slot));
    else tmem_alloc<512>(smem_u32(tmem_slot));
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // arrive on a barrier the MMA issuer (leader CTA) waits on
  auto arrive_leader = [&](uint32_t bar) {
    if constexpr (kPair) mbar_arrive_cluster(mapa(bar, 0));
    else mbar_arrive(bar);
  };

  if (warp == 0) {
    // ===================== bulk-copy producer =====================
    if (lane == 0) {
      int prev_mol = -1;
      uint32_t wi = 0;
      for (int it = 0; it < n_iter; ++it) {
        const EdgeTile ti = fetch_tile(it);
        const int mol = ti.mol, i0 = ti.i0, ng = ti.ng, n = ti.n;
        const int node0 = p.node_off[mol];
        mbar_wait(pq_empty, (uint32_t)((it & 1) ^ 1));
        const bool newmol = (mol != prev_mol);
        mbar_arrive_expect_tx(pq_full, (uint32_t)((ng + (newmol ? n : 0)) * EdgeSmem::PQ_ROW));
        const uint8_t* pqb = reinterpret_cast<const uint8_t*>(p.pq);
        for (int g = 0; g < ng; ++g)
          bulk_g2s(base + EdgeSmem::P_OFF + g * EdgeSmem::PQ_PITCH, pqb + (size_t)(node0 + i0 + g) * (2 * EdgeSmem::PQ_ROW),
                   EdgeSmem::PQ_ROW, pq_full);
        if (newmol)
          for (int j = 0; j < n; ++j)
            bulk_g2s(base + EdgeSmem::Q_OFF + j * EdgeSmem::PQ_PITCH,
                     pqb + (size_t)(node0 + j) * (2 * EdgeSmem::PQ_ROW) + EdgeSmem::PQ_ROW, EdgeSmem::PQ_ROW, pq_full);
        prev_mol = mol;
        if constexpr (kSegMma) {
          // the W ring doubles as staging for the previous tile's messages until its segment-sum MMAs have completed
          if (it > 0) mbar_wait(e_done(0), (uint32_t)((it - 1) & 1));
        }
        if constexpr (kPair) {
          // one ring slot per K chunk: this CTA's 112 rows of each N half (2 x 14,336 B)
          for (int kc = 0; kc < p.n_kc; ++kc, ++wi) {
            const int s = wi % EDGE_NW;
            mbar_wait(w_empty(s), ((wi / EDGE_NW) & 1) ^ 1u);
            mbar_arrive_expect_tx(w_full(s), EDGE_WSLOT);
            const uint8_t* src = p.w2 + (size_t)kc * 2 * EDGE_WSLOT + crank * (EDGE_WSLOT / 2);
            const uint32_t dst = base + EdgeSmem::W_OFF + s * EDGE_WSLOT;
            bulk_g2s(dst, src, EDGE_WSLOT / 2, w_full(s));
            bulk_g2s(dst + EDGE_WSLOT / 2, src + EDGE_WSLOT, EDGE_WSLOT / 2, w_full(s));
          }
        } else {
          for (int kc = 0; kc < p.n_kc; ++kc) {
            for (int nh = 0; nh < 2; ++nh, ++wi) {
              const int s = wi % EDGE_NW;
              const uint32_t ph = (wi / EDGE_NW) & 1;
              mbar_wait(w_empty(s), ph ^ 1u);
              mbar_arrive_expect_tx(w_full(s), EDGE_WSLOT);
              bulk_g2s(base + EdgeSmem::W_OFF + s * EDGE_WSLOT, p.w2 + ((size_t)kc * 2 + nh) * EDGE_WSLOT, EDGE_WSLOT,
                       w_full(s));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== tcgen05.mma issuer (A from TMEM, B from shared memory) =====================
    if (lane == 0 && kPair && crank == 1) {
      // peer CTA: relay "my half of W slot s has landed" to the leader, which issues the joint MMAs
      uint32_t wi = 0;
      for (int it = 0; it < n_iter; ++it)
        for (int kc = 0; kc < p.n_kc; ++kc, ++wi) {
          const int s = wi % EDGE_NW;
          mbar_wait(w_full(s), (wi / EDGE_NW) & 1);
          mbar_arrive_cluster(mapa(w_peer(s), 0));
        }
    } else if (lane == 0) {
      constexpr int MM = kPair ? 2 * TILE_M : TILE_M;
      constexpr uint32_t idesc = umma_idesc(umma_fmt(kMode), MM, 224);
      uint32_t wi = 0, ai = 0;
      for (int it = 0; it < n_iter; ++it) {
        if constexpr (kPair) mbar_wait(d_empty, (uint32_t)((it & 1) ^ 1));
        else mbar_wait(d_empty, (uint32_t)((it & 1) ^ 1));
        tc_fence_after();
        for (int kc = 0; kc < p.n_kc; ++kc, ++ai) {
          const int as = ai % EDGE_NA;
          if constexpr (kPair) mbar_wait(a_full(as), (ai / EDGE_NA) & 1);
          else mbar_wait(a_full(as), (ai / EDGE_NA) & 1);
          const uint32_t a_tmem = tmem_base + EDGE_ACOL + as * 32;
          if constexpr (kPair) {
            const int ws = wi % EDGE_NW;
            mbar_wait(w_full(ws), (wi / EDGE_NW) & 1);
            mbar_wait(w_peer(ws), (wi / EDGE_NW) & 1);
            tc_fence_after();
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) {
              const uint64_t bdesc = umma_desc_sw128(base + EdgeSmem::W_OFF + ws * EDGE_WSLOT + nh * (EDGE_WSLOT / 2));
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_ts_pair<kMode>(tmem_base + nh * 224, a_tmem + ks * 8, bdesc + 2 * ks, idesc, (kc | ks) != 0);
              // last chunk: the first N half is complete half a chunk before the second; the epilogue warps of columns
              // 0..223 start on it while the tensor core finishes columns 224..447
              if (nh == 0 && kc + 1 == p.n_kc) umma_commit_pair(d_half, 3);
            }
            umma_commit_pair(w_empty(ws), 3);
            ++wi;
            umma_commit_pair(a_empty(as), 3);
          } else {
            for (int nh = 0; nh < 2; ++nh, ++wi) {
              const int ws = wi % EDGE_NW;
              mbar_wait(w_full(ws), (wi / EDGE_NW) & 1);
              tc_fence_after();
              const uint64_t bdesc = umma_desc_sw128(base + EdgeSmem::W_OFF + ws * EDGE_WSLOT);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_ts<kMode>(tmem_base + nh * 224, a_tmem + ks * 8, bdesc + 2 * ks, idesc, (kc | ks) != 0);
              if (nh == 0 && kc + 1 == p.n_kc) umma_commit(d_half);
              umma_commit(w_empty(ws));
            }
            umma_commit(a_empty(as));
          }
        }
        if constexpr (kPair) umma_commit_pair(d_full, 3);
        else umma_commit(d_full);
        if constexpr (kSegMma) {
          // segment sum over neighbours on the tensor core: D2[128 channels x 16 groups] = E^T . S^T per 128-channel
          // block; E (gated messages, bf16) is staged row-major = MN-major A operand, S is the 0/1 group selector.
          // Pair mode: M = 256 covers both CTAs' messages, N = 32 = [S of CTA 0 ; S of CTA 1]; CTA r uses D2 columns
          // 16r .. 16r+15.
          constexpr uint32_t idesc2 = umma_idesc(umma_fmt(kMode), MM, kPair ? 32 : 16) | (1u << 15);  // A is MN-major
          constexpr int D2W = kPair ? 32 : 16;
          // all four channel blocks of the tile are staged (pass 1) and the gate-weighted selector is built: one batch
          mbar_wait(e_full(0), (uint32_t)(it & 1));
          tc_fence_after();
          for (int cb = 0; cb < 4; ++cb) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t adesc = umma_desc_mn_sw128(base + EdgeSmem::stage_off(cb) + ks * 2048, A_CHUNK_BYTES);
              const uint64_t bdesc = umma_desc_sw128(base + EdgeSmem::SEL_OFF + (ks >> 2) * 2048) + 2 * (ks & 3);
              // results go to D columns cb*D2W.. : D is free (every warp arrived on e_full only after its pass 1) and the
              // next tile's MMAs cannot start before every warp has finished this tile's readout
              if constexpr (kPair) umma_ss_bf16_pair(tmem_base + cb * D2W, adesc, bdesc, idesc2, ks != 0);
              else umma<kMode>(tmem_base + cb * D2W, adesc, bdesc, idesc2, ks != 0);
            }
          }
          if constexpr (kPair) umma_commit_pair(e_done(0), 3);
          else umma_commit(e_done(0));
        }
      }
    }
  } else {
    // ===================== compute warps: A generation + epilogue =====================
    const int ct = threadIdx.x - 64;       // 0..511
    const int cw = ct >> 5;                // 0..15
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int qq = cw >> 2;                // quarter of each K chunk (A generation) / of the output columns (epilogue)
    const int r = q * 32 + lane;           // tile row = TMEM lane
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t ai = 0;
    long long* pacc = reinterpret_cast<long long*>(gbase + EdgeSmem::PROF_OFF);  // only thread ct == 0 touches it
    const bool profiling = kProf && (p.prof != nullptr) && (ct == 0);
    if (profiling)
      for (int k = 0; k < 16; ++k) pacc[k] = 0;
    // Row metadata (d2, d0^2, group / neighbour ids, unit vectors) and the group selector of a tile are double-buffered:
    // they are computed for tile it+1 while tile it waits for its last MMAs.
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
      }
    };
    if (n_iter > 0) tile_setup(fetch_tile(0), 0);
    named_bar_sync(1, EDGE_CT);
    for (int it = 0; it < n_iter; ++it) {
      long long c0 = profiling ? clock64() : 0;
      const EdgeTile ti = fetch_tile(it);
      const int mol = ti.mol, i0 = ti.i0, ng = ti.ng, n = ti.n;
      const int nm1 = max(n - 1, 1);
      const int node0 = p.node_off[mol];
      const int buf = it & 1;
      // rows [glo(g), ghi(g)) of this tile belong to target i0+g; gfix(g) = split-target id if its list continues in a
      // neighbouring tile
      auto glo = [&](int g) { return max(g * nm1 - ti.off0, 0); };
      auto ghi = [&](int g) { return min((g + 1) * nm1 - ti.off0, ti.nrows); };
      auto gfix = [&](int g) { return (g == 0 && ti.fixa >= 0) ? ti.fixa : (g == ng - 1 && ti.fixb >= 0) ? ti.fixb : -1; };
      // carried partial sums: written for this tile's last target into buffer (it & 1), read for its first target from
      // buffer ((it - 1) & 1) -- the unused half of the selector area
      float* carry_wr = reinterpret_cast<float*>(gbase + EdgeSmem::SEL_OFF + 4096) + (it & 1) * 464;
      const float* carry_rd = reinterpret_cast<const float*>(gbase + EdgeSmem::SEL_OFF + 4096) + ((it & 1) ^ 1) * 464;
      auto carried_in = [&](int g) { return g == 0 && ti.fixa == EDGE_CARRY; };
      auto carried_out = [&](int g) { return g == ng - 1 && ti.fixb == EDGE_CARRY; };
      float* trs = trs_all + buf * TILE_M * 3;
      const float2* ri_d = ri_d_all + buf * TILE_M;
      const int* ri_gj = ri_gj_all + buf * TILE_M;
      const int info = ri_gj[r];
      const bool valid = (info & 0x10000) != 0;
      const float2 rd = ri_d[r];
      const uint8_t* Prow = Ps + (info & 0xff) * EdgeSmem::PQ_PITCH + qq * ELEMS * EdgeSmem::PQ_ESIZE;  // invalid rows read row 0
      const uint8_t* Qrow = Qs + ((info >> 8) & 0xff) * EdgeSmem::PQ_PITCH + qq * ELEMS * EdgeSmem::PQ_ESIZE;
      // 16-bit modes: packed distance features (fp16: scaled by 2^-10 and clamped below the fp16 maximum)
      auto pack_dist = [&](float d) {
        const float ds = (kMode == PREC_FP16) ? fminf(d * kDistScale, 60000.0f) : d;
        return pack_h2<kMode>(ds, ds);
      };
      const uint32_t d2h = pack_dist(rd.x), d02h = pack_dist(rd.y);
      if (!(kEarlyA && it > 0)) mbar_wait(pq_full, (uint32_t)(it & 1));  // (bf16: waited for during the previous tile)
      if (profiling) { long long c = clock64(); pacc[0] += c - c0; c0 = c; }  // row info + P/Q wait

      // One K chunk of the A operand for the tile whose row data is (Prow, Qrow, rd, d2h, d02h): SiLU(P_i + Q_j + d2*wc +
      // d02*wd) -> TMEM A ring stage ai % 2.
      auto agen_chunk = [&](int kc, const uint8_t* Prow, const uint8_t* Qrow, float2 rd, uint32_t d2h, uint32_t d02h) {
        const int as = ai % EDGE_NA;
        const int k0 = kc * EPC + qq * ELEMS;  // warp-uniform
        uint32_t w[8];
        if constexpr (kPqBf16) {
          // bf16 fast mode: the whole pre-activation is packed bf16x2 math -- P+Q (HADD2), the two distance terms
          // (HFMA2 with d2 / d0^2 rounded to bf16), tanh and h + h*tanh(h) (HFMA2); the result is directly the packed
          // A-operand word.  wc / wd come from the constant bank as packed bf16x2 (warp-uniform index).
#pragma unroll
          for (int e = 0; e < ELEMS; e += 8) {
            // K columns >= 424 are padding of the 420 (+ bias column) real ones: their weights are zero and so is the
            // activation (P = Q = 0 there); skip the loads and the MUFU work (warp-uniform: 1.5 of the last chunk's 4 quarters)
            if (k0 + e >= 424) {
              w[(e >> 1) + 0] = w[(e >> 1) + 1] = w[(e >> 1) + 2] = w[(e >> 1) + 3] = 0u;
              continue;
            }
            const uint4 pw = *reinterpret_cast<const uint4*>(Prow + (kc * EPC + e) * 2);
            const uint4 qw = *reinterpret_cast<const uint4*>(Qrow + (kc * EPC + e) * 2);
            const uint4 c0 = *reinterpret_cast<const uint4*>(&p.wcd_h[(k0 + e)]);      // channels k0+e .. +3: wc, wc, wd, wd
            const uint4 c1 = *reinterpret_cast<const uint4*>(&p.wcd_h[(k0 + e) + 4]);  // channels k0+e+4 .. +7
            const uint32_t pa[4] = {pw.x, pw.y, pw.z, pw.w}, qa[4] = {qw.x, qw.y, qw.z, qw.w};
            const uint32_t wcp[4] = {c0.x, c0.y, c1.x, c1.y}, wdp[4] = {c0.z, c0.w, c1.z, c1.w};
            [[maybe_unused]] float wcv[8], wdv[8];
            if constexpr (kDistF32) {  // fp32 columns of the two distance features: four 128-bit constant-bank loads
              const float4 a0 = *reinterpret_cast<const float4*>(&p.wc[k0 + e]), a1 = *reinterpret_cast<const float4*>(&p.wc[k0 + e + 4]);
              const float4 b0 = *reinterpret_cast<const float4*>(&p.wd[k0 + e]), b1 = *reinterpret_cast<const float4*>(&p.wd[k0 + e + 4]);
              wcv[0] = a0.x; wcv[1] = a0.y; wcv[2] = a0.z; wcv[3] = a0.w; wcv[4] = a1.x; wcv[5] = a1.y; wcv[6] = a1.z; wcv[7] = a1.w;
              wdv[0] = b0.x; wdv[1] = b0.y; wdv[2] = b0.z; wdv[3] = b0.w; wdv[4] = b1.x; wdv[5] = b1.y; wdv[6] = b1.z; wdv[7] = b1.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t h2 = hadd2<kMode>(pa[i], qa[i]);
              if constexpr (kDistF32) {
                // widen P + Q, add the two large, possibly cancelling distance terms in fp32, round once.  (Summing only the
                // distance terms in fp32 and adding them to P + Q in 16 bits saves one instruction per pair = 0.5 % of the
                // launch but costs 25 % in per-step error on the blown-up trajectories: measured, rejected.)
                const float h_lo = fmaf(rd.y, wdv[2 * i], fmaf(rd.x, wcv[2 * i], h2_lo<kMode>(h2)));
                const float h_hi = fmaf(rd.y, wdv[2 * i + 1], fmaf(rd.x, wcv[2 * i + 1], h2_hi<kMode>(h2)));
                h2 = pack_h2<kMode>(h_lo, h_hi);
              } else {
                h2 = hfma2<kMode>(d2h, wcp[i], h2);
                h2 = hfma2<kMode>(d02h, wdp[i], h2);
              }
              w[(e >> 1) + i] = hsilu2<kMode>(h2);
            }
          }
          // constant-1 column carrying b2 (K index 420 = element 4 of the run that starts at 416): low half of word 2
          if (k0 == BIAS_COL - 4) w[2] = (w[2] & 0xffff0000u) | h_act_one_bits<kMode>();
        } else {
          float wcv[ELEMS], wdv[ELEMS];  // 128-bit constant-bank loads (k0 is a multiple of 8)
#pragma unroll
          for (int e = 0; e < ELEMS; e += 4) {
            const float4 c4 = *reinterpret_cast<const float4*>(&p.wc[k0 + e]);
            const float4 d4 = *reinterpret_cast<const float4*>(&p.wd[k0 + e]);
            wcv[e] = c4.x; wcv[e + 1] = c4.y; wcv[e + 2] = c4.z; wcv[e + 3] = c4.w;
            wdv[e] = d4.x; wdv[e + 1] = d4.y; wdv[e + 2] = d4.z; wdv[e + 3] = d4.w;
          }
          float a[ELEMS];
#pragma unroll
          for (int e = 0; e < ELEMS; e += 4) {
            const float4 pv = *reinterpret_cast<const float4*>(Prow + (kc * EPC + e) * 4);
            const float4 qv = *reinterpret_cast<const float4*>(Qrow + (kc * EPC + e) * 4);
            a[e + 0] = silu_scaled<kFast>(fmaf(rd.y, wdv[e + 0], fmaf(rd.x, wcv[e + 0], pv.x + qv.x)));
            a[e + 1] = silu_scaled<kFast>(fmaf(rd.y, wdv[e + 1], fmaf(rd.x, wcv[e + 1], pv.y + qv.y)));
            a[e + 2] = silu_scaled<kFast>(fmaf(rd.y, wdv[e + 2], fmaf(rd.x, wcv[e + 2], pv.z + qv.z)));
            a[e + 3] = silu_scaled<kFast>(fmaf(rd.y, wdv[e + 3], fmaf(rd.x, wcv[e + 3], pv.w + qv.w)));
          }
          if (k0 == BIAS_COL - 4) a[4] = 1.0f;  // constant-1 column carrying b2
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = f32_to_tf32(a[i]);
        }
        long long cw0 = profiling ? clock64() : 0;
        mbar_wait(a_empty(as), ((ai / EDGE_NA) & 1) ^ 1u);
        if (profiling) pacc[5] += clock64() - cw0;  // A-ring back-pressure (MMA / weight stream slower than A generation)
        long long cw1 = profiling ? clock64() : 0;
        tc_fence_after();
        tmem_st8(trow + EDGE_ACOL + as * 32 + qq * 8, w);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(a_full(as));  // one arrival per warp
        if (profiling) pacc[11] += clock64() - cw1;  // A-operand hand-off (tcgen05.st + wait + publish)
        ++ai;
      };
      // First EDGE_EARLY chunks of tile it+1 (row data from the other metadata buffer: tile_setup ran during this tile's
      // MMA tail).  Called inside the epilogue of tile it, when the MUFU pipe is idle and the A ring is free.
      auto agen_early = [&]() {
        if (it + 1 < n_iter) {
          const int nb = buf ^ 1;
          const int info2 = ri_gj_all[nb * TILE_M + r];
          const float2 rd2 = ri_d_all[nb * TILE_M + r];
          const uint8_t* Prow2 = Ps + (info2 & 0xff) * EdgeSmem::PQ_PITCH + qq * ELEMS * EdgeSmem::PQ_ESIZE;
          const uint8_t* Qrow2 = Qs + ((info2 >> 8) & 0xff) * EdgeSmem::PQ_PITCH + qq * ELEMS * EdgeSmem::PQ_ESIZE;
          mbar_wait(pq_full, (uint32_t)((it + 1) & 1));
          for (int kc = 0; kc < min(EDGE_EARLY, p.n_kc); ++kc)
            agen_chunk(kc, Prow2, Qrow2, rd2, pack_dist(rd2.x), pack_dist(rd2.y));
        }
      };
      // ---- A generation ----  (bf16: chunks 0 and 1 were generated during the previous tile's epilogue)
      {
        const int kc_begin = (kEarlyA && it > 0) ? min(EDGE_EARLY, p.n_kc) : 0;
#pragma unroll 1
        for (int kc = kc_begin; kc < p.n_kc; ++kc) agen_chunk(kc, Prow, Qrow, rd, d2h, d02h);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(pq_empty);  // P/Q rows of this tile are no longer needed
      if (it + 1 < n_iter) tile_setup(fetch_tile(it + 1), buf ^ 1);  // overlaps this tile's last MMAs
      if (profiling) { long long c = clock64(); pacc[1] += c - c0; c0 = c; }  // A generation (incl. back-pressure)

      // ---- epilogue pass 1: m = SiLU(D) (written back to TMEM for GCL), partial dot with the gate / coord vector ----
      mbar_wait(qq < 2 ? d_half : d_full, (uint32_t)(it & 1));  // this thread's output columns are 112*qq .. 112*qq+111
      tc_fence_after();
      if (profiling) { long long c = clock64(); pacc[2] += c - c0; c0 = c; }  // MMA tail
      float dotp[4] = {0.f, 0.f, 0.f, 0.f};
      // thread (row r, quarter qq) owns output columns 112*qq .. 112*qq+111 in 7 runs of 16.  GCL bf16 mode keeps the
      // SiLU'd messages as packed bf16 in registers (ew) for pass 2; the other variants write them back to TMEM.
      // software-pipelined TMEM reads: the load of run ch+1 is in flight while run ch is processed
      float vbuf[2][16];
      tmem_ld16(trow + qq * 112, vbuf[0]);
      tmem_wait_ld();
#pragma unroll
      for (int ch = 0; ch < 7; ++ch) {
        const int col0 = qq * 112 + ch * 16;  // warp-uniform
        float* v = vbuf[ch & 1];
        if (ch + 1 < 7) tmem_ld16(trow + col0 + 16, vbuf[(ch + 1) & 1]);
        [[maybe_unused]] uint32_t mw[8];  // the run's 16 messages, packed bf16
        if constexpr (kFast) {
          // packed path: h (fp32, already halved) -> bf16x2, one MUFU.TANH per pair, m = h + h*tanh(h) as HFMA2; the dot
          // with the gate / coordinate vector accumulates the widened halves in fp32
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wv_s + col0 + e);
            uint32_t m2a, m2b;
            if constexpr (kMode == PREC_FP16) {
              // fp16 mode: the accumulator is ACT * h (fp32); SiLU in fp32 -- tanh of the unscaled argument, the message
              // stays ACT-scaled -- and ONE rounding when the pair is packed.  Same instruction count as the packed path.
              float mm[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float t;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(v[e + u] * (1.0f / act_scale(kMode))));
                mm[u] = fmaf(v[e + u], t, v[e + u]);
              }
              dotp[0] = fmaf(mm[0], w4.x, dotp[0]);
              dotp[1] = fmaf(mm[1], w4.y, dotp[1]);
              dotp[2] = fmaf(mm[2], w4.z, dotp[2]);
              dotp[3] = fmaf(mm[3], w4.w, dotp[3]);
              m2a = pack_h2<kMode>(mm[0], mm[1]);
              m2b = pack_h2<kMode>(mm[2], mm[3]);
            } else {
              m2a = hsilu2<kMode>(pack_h2<kMode>(v[e + 0], v[e + 1]));
              m2b = hsilu2<kMode>(pack_h2<kMode>(v[e + 2], v[e + 3]));
              dotp[0] = fmaf(h2_lo<kMode>(m2a), w4.x, dotp[0]);
              dotp[1] = fmaf(h2_hi<kMode>(m2a), w4.y, dotp[1]);
              dotp[2] = fmaf(h2_lo<kMode>(m2b), w4.z, dotp[2]);
              dotp[3] = fmaf(h2_hi<kMode>(m2b), w4.w, dotp[3]);
            }
            if constexpr (kSegMma) {
              mw[(e >> 1)] = m2a;
              mw[(e >> 1) + 1] = m2b;
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wv_s + col0 + e);
            v[e + 0] = silu_scaled<kFast>(v[e + 0]); v[e + 1] = silu_scaled<kFast>(v[e + 1]);
            v[e + 2] = silu_scaled<kFast>(v[e + 2]); v[e + 3] = silu_scaled<kFast>(v[e + 3]);
            dotp[0] = fmaf(v[e + 0], w4.x, dotp[0]); dotp[1] = fmaf(v[e + 1], w4.y, dotp[1]);
            dotp[2] = fmaf(v[e + 2], w4.z, dotp[2]); dotp[3] = fmaf(v[e + 3], w4.w, dotp[3]);
          }
        }
        if constexpr (kSegMma) {
          // Stage the (ungated) messages for the segment-sum MMAs -- the gate goes into the selector.  Block cb < 3 holds
          // runs 2cb, 2cb+1 of every quarter (32 channels x 4 quarters = 128 MMA rows, row m = 32*qq + c); block 3 holds
          // run 6 of every quarter (16 channels x 4 = 64 rows, m = 16*qq + c).  Layout per block: two 16 KB operand
          // chunks of [128 tile rows x 64 channels], i.e. MN-major A with M = channels, K = tile rows.
          // Blocks 2 and 3 live in the W ring.  The warps of columns 0..223 started on d_half, i.e. possibly while the last
          // chunk's MMAs for columns 224..447 were still reading their W slot: they must see d_full before the first store
          // into the ring (long complete by then; this only makes the ordering explicit).
          if (ch == 4 && qq < 2) mbar_wait(d_full, (uint32_t)(it & 1));
          uint8_t* dst;
          int piece;
          if (ch < 6) {
            dst = gbase + EdgeSmem::stage_off(ch >> 1) + (qq >> 1) * A_CHUNK_BYTES;
            piece = (qq & 1) * 4 + (ch & 1) * 2;
          } else {
            dst = gbase + EdgeSmem::stage_off(3);
            piece = qq * 2;
          }
          *reinterpret_cast<uint4*>(dst + sw128_offset(r, piece)) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
          *reinterpret_cast<uint4*>(dst + sw128_offset(r, piece + 1)) = make_uint4(mw[4], mw[5], mw[6], mw[7]);
        } else if constexpr (!kEquiv) {
          tmem_st16(trow + col0, v);
        }
        if (ch + 1 < 7) tmem_wait_ld();
      }
      dots[qq * TILE_M + r] = (dotp[0] + dotp[1]) + (dotp[2] + dotp[3]);
      if constexpr (!kEquiv && !kSegMma) tmem_wait_st();
      named_bar_sync(1, EDGE_CT);
      if (profiling) { long long c = clock64(); pacc[3] += c - c0; c0 = c; }  // pass 1
      const float full_dot = (dots[r] + dots[TILE_M + r]) + (dots[2 * TILE_M + r] + dots[3 * TILE_M + r]);
      if constexpr (kEquiv) {
        // x_i += sum_j unit_ij * phi_ij / 100   (reference egnn.py:124-134)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(d_empty);
        if (qq == 0) {
          const float phi = valid ? full_dot : 0.f;
          trs[r * 3 + 0] *= phi; trs[r * 3 + 1] *= phi; trs[r * 3 + 2] *= phi;
        }
        named_bar_sync(1, EDGE_CT);
        if (ct < ng * 3) {
          const int gg = ct / 3, c = ct - gg * 3;
          float s = 0.f;
          for (int e = glo(gg); e < ghi(gg); ++e) s += trs[e * 3 + c];
          const int fx = gfix(gg);
          const size_t idx = (size_t)(node0 + i0 + gg) * 3 + c;
          if (carried_in(gg)) s = carry_rd[448 + c] + s;
          if (carried_out(gg)) carry_wr[448 + c] = s;
          else if (fx >= 0) atomicAdd(p.fix_dx + (size_t)fx * 4 + c, s);
          else p.x_next[idx] = p.x_cur[idx] + s / 100.0f;
        }
        if constexpr (kEarlyA) agen_early();  // D was released above: the next tile's MMAs start as soon as chunk 0 is published
        named_bar_sync(1, EDGE_CT);
      } else {
        // e_ij = m_ij * sigmoid(w_a.m_ij + b_a); agg_i = sum_j e_ij / 100   (reference egnn.py:48-51, 59-64)
        const float gate = valid ? sigmoid_acc(full_dot + p.att_bias) : 0.f;
        if constexpr (kSegMma) {
          // The messages were staged in pass 1; the gate enters through the selector:
          //   agg[g][c] = sum_k S'[g][k] * m[k][c],   S'[g][k] = gate_k if tile row k belongs to target g, else 0.
          long long s0 = profiling ? clock64() : 0;
          float* gates = trs_all;  // [128]; the coordinate-message buffer is unused by GCL layers
          if (qq == 0) gates[r] = gate;
          // D2 readout: lane r = MMA row m, columns = groups; quarter qq stores groups g = qq, qq+4, qq+8.  First all four
          // blocks are pulled out of TMEM (3 values per block and thread), then D is released -- the next tile's MMAs start
          // while the global stores below are still in flight.
          float dval[4][3];
          auto readout_ld2 = [&](int cb) {  // blocks cb and cb+1: both TMEM loads in flight
            float v[16], u[16];
            tmem_ld16(trow + cb * (kPair ? 32 : 16) + (kPair ? crank * 16 : 0), v);
            tmem_ld16(trow + (cb + 1) * (kPair ? 32 : 16) + (kPair ? crank * 16 : 0), u);
            tmem_wait_ld();
            // select [qq + 4 gi] without dynamic register indexing
            dval[cb][0] = qq == 0 ? v[0] : qq == 1 ? v[1] : qq == 2 ? v[2] : v[3];
            dval[cb][1] = qq == 0 ? v[4] : qq == 1 ? v[5] : qq == 2 ? v[6] : v[7];
            dval[cb][2] = qq == 0 ? v[8] : qq == 1 ? v[9] : qq == 2 ? v[10] : v[11];
            dval[cb + 1][0] = qq == 0 ? u[0] : qq == 1 ? u[1] : qq == 2 ? u[2] : u[3];
            dval[cb + 1][1] = qq == 0 ? u[4] : qq == 1 ? u[5] : qq == 2 ? u[6] : u[7];
            dval[cb + 1][2] = qq == 0 ? u[8] : qq == 1 ? u[9] : qq == 2 ? u[10] : u[11];
          };
          auto readout_st = [&](int cb) {
            const int ch = (cb < 3) ? 112 * (r >> 5) + 32 * cb + (r & 31) : 112 * (r >> 4) + 96 + (r & 15);
            if (cb < 3 || r < 64) {
              uint8_t* cbase = p.agg_op + (size_t)(ch >> 6) * A_CHUNK_BYTES + (ch & 7) * 2;
              const int rd_piece = (ch >> 3) & 7;
#pragma unroll
              for (int gi = 0; gi < 3; ++gi) {
                const int g = qq + 4 * gi;
                if (g < ng) {
                  const int node = node0 + i0 + g;
                  uint8_t* dst = cbase + (size_t)(node >> 7) * p.agg_chunks * A_CHUNK_BYTES + (node & 127) * 128 +
                                 ((rd_piece ^ (node & 7)) << 4);
                  float val = dval[cb][gi];
                  const int fx = gfix(g);
                  if (carried_in(g)) val = carry_rd[ch] + val;
                  if (carried_out(g)) carry_wr[ch] = val;
                  else if (fx >= 0) atomicAdd(p.fix_agg + (size_t)fx * HP + ch, val);
                  else store_h<kMode>(dst, val * agg_out_scale(kMode));
                }
              }
            }
          };
          named_bar_sync(1, EDGE_CT);  // all gates written; every thread's pass-1 staging stores are issued
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
            *reinterpret_cast<uint4*>(gbase + EdgeSmem::SEL_OFF + (piece >> 3) * 2048 + sw128_offset(g, piece & 7)) =
                make_uint4(w[0], w[1], w[2], w[3]);
          }
          fence_proxy_async();  // staging + selector stores -> visible to the tensor core
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(e_full(0));
          if (profiling) { long long c = clock64(); pacc[9] += c - s0; s0 = c; }   // gate exchange + selector + publish
          // While the tensor core runs the segment-sum MMAs the MUFU pipe is idle and the A ring is free: generate the
          // first two K chunks of the NEXT tile now.  Its MMAs cannot start before this tile's readout (they wait for
          // d_empty, which is released after the readout), but then they find both ring stages full.
          agen_early();
          mbar_wait(e_done(0), (uint32_t)(it & 1));
          tc_fence_after();
          if (profiling) { long long c = clock64(); pacc[7] += c - s0; s0 = c; }   // segment-sum MMAs
          readout_ld2(0);
          readout_ld2(2);
          // D (which held the segment sums) is free: the next tile's MMAs may start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(d_empty);
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) readout_st(cb);
          if (profiling) { long long c = clock64(); pacc[10] += c - s0; s0 = c; }
        } else {
#pragma unroll 1
        for (int ch = 0; ch < 7; ++ch) {
          const int col0 = qq * 112 + ch * 16;
          float v[16];
          tmem_ld16(trow + col0, v);
          tmem_wait_ld();
          if (ch == 6) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader(d_empty);  // last TMEM read of this tile is complete
          }
          named_bar_sync(1, EDGE_CT);  // previous chunk's readers are done with the scratch
          // staging: two 16 KB halves of [128 rows x 32 fp32]; quarter qq lands in half qq>>1, columns (qq&1)*16..
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            float4 o;
            o.x = v[e + 0] * gate; o.y = v[e + 1] * gate; o.z = v[e + 2] * gate; o.w = v[e + 3] * gate;
            *reinterpret_cast<float4*>(scratch + (qq >> 1) * A_CHUNK_BYTES + sw128_offset(r, (qq & 1) * 4 + (e >> 2))) = o;
          }
          named_bar_sync(1, EDGE_CT);
          // segment sum over the neighbours of each target node (4 interleaved partial sums, fixed order)
          for (int pr = cw; pr < 2 * ng; pr += 16) {
            const int hh = pr & 1, gg = pr >> 1;
            const uint8_t* src = scratch + hh * A_CHUNK_BYTES + (lane & 3) * 4;
            const int r0 = glo(gg), cnt = ghi(gg) - r0;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int e = 0;
            for (; e + 3 < cnt; e += 4) {
              s0 += *reinterpret_cast<const float*>(src + sw128_offset(r0 + e + 0, lane >> 2));
              s1 += *reinterpret_cast<const float*>(src + sw128_offset(r0 + e + 1, lane >> 2));
              s2 += *reinterpret_cast<const float*>(src + sw128_offset(r0 + e + 2, lane >> 2));
              s3 += *reinterpret_cast<const float*>(src + sw128_offset(r0 + e + 3, lane >> 2));
            }
            for (; e < cnt; ++e) s0 += *reinterpret_cast<const float*>(src + sw128_offset(r0 + e, lane >> 2));
            const int col = (2 * hh + (lane >> 4)) * 112 + ch * 16 + (lane & 15);
            const int fx = gfix(gg);
            float tot = (s0 + s1) + (s2 + s3);
            if (carried_in(gg)) tot = carry_rd[col] + tot;
            if (carried_out(gg)) carry_wr[col] = tot;
            else if (fx >= 0) atomicAdd(p.fix_agg + (size_t)fx * HP + col, tot);
            else op_store1<kMode>(p.agg_op, p.agg_chunks, node0 + i0 + gg, col, tot / 100.0f);
          }
        }
        }
        if constexpr (!kSegMma) named_bar_sync(1, EDGE_CT);  // staging is free again (bf16 GCL: next written after later barriers)
      }
      if (profiling) { long long c = clock64(); pacc[4] += c - c0; pacc[6] += 1; }  // pass 2 / coordinate update
    }
    if (profiling)
      for (int k = 0; k < 12; ++k) p.prof[(size_t)blockIdx.x * 16 + k] = pacc[k];
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) {
    cluster_sync_all();  // the peer may still be reading this CTA's shared memory / TMEM through the joint MMAs
    if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
  } else {
    if (warp == 1) tmem_dealloc<512>(tmem_base);
  }
}

// Completes the split targets of an edge-kernel launch: their two partial sums have been added into the side buffer;
// write the regular output (aggregate in operand format, or the coordinate update) and re-zero the buffer.
template <int kMode, bool kEquiv>
__global__ void k_edge_fixup(const int* __restrict__ fix_node, int n_fix, float* fix_agg, float* fix_dx, uint8_t* agg_op,
                             int agg_chunks, const float* __restrict__ x_cur, float* x_next) {
  if constexpr (kEquiv) {
    // thread = (split target, coordinate)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = t >> 2, c = t & 3;
    if (f < n_fix && c < 3) {
      const size_t idx = (size_t)fix_node[f] * 3 + c;
      const float s = fix_dx[(size_t)f * 4 + c];
      fix_dx[(size_t)f * 4 + c] = 0.f;
      x_next[idx] = x_cur[idx] + s / 100.0f;
    }
  } else {
    // thread = (split target, 16-byte piece of the operand row): EPP channels -> one vector store
    constexpr int EPP = epp(kMode), NP = HP / EPP;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = t / NP, pc = t - f * NP;
    if (f >= n_fix) return;
    const int node = fix_node[f];
    float* src = fix_agg + (size_t)f * HP + pc * EPP;
    float v[EPP];
#pragma unroll
    for (int e = 0; e < EPP; e += 4) {
      const float4 x = *reinterpret_cast<const float4*>(src + e);
      *reinterpret_cast<float4*>(src + e) = make_float4(0.f, 0.f, 0.f, 0.f);
      v[e] = x.x; v[e + 1] = x.y; v[e + 2] = x.z; v[e + 3] = x.w;
    }
#pragma unroll
    for (int e = 0; e < EPP; ++e) v[e] = is16(kMode) ? v[e] * agg_out_scale(kMode) : v[e] / 100.0f;
    op_store<kMode, EPP>(agg_op, agg_chunks, node, pc * EPP, v);
  }
}

}  // namespace mlcg
