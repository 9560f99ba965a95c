// tcgen05 / TMEM kernels of the hot path (sm_100a only).
//
//  k_tc_gemm  : C[128-row tile, BN cols] = epilogue(A . W^T).  A and W arrive as pre-swizzled operand-format blocks
//               through 1-D bulk async copies (UBLKCP) into a multi-stage mbarrier ring; one thread issues
//               tcgen05.mma with the fp32 accumulator in TMEM; four epilogue warps read it back with tcgen05.ld.
//               Used for the per-node projections / node MLPs of the EGNN (reference egnn.py:30-34,54-68) and for
//               the AdjMatSeer linears (reference adj_mat_seer.py:53).
//  k_tc_edge  : the fused all-pairs edge MLP of one EGNN sub-layer (reference egnn.py:38-52 + 418-437 for GCL,
//               egnn.py:111-135 for EquivariantUpdate).  Edge tensors never touch HBM: the first-layer activation
//               SiLU(P_i + Q_j + d2*wc + d02*wd) is generated straight into the swizzled A-operand ring in shared
//               memory, the 420x420 second layer runs on tcgen05 with the accumulator in TMEM, and SiLU, attention
//               gate, masked neighbour sum / coordinate update run in the epilogue.
#pragma once
#include "mlcg_common.cuh"

namespace mlcg {

// ---------------------------------------------------------------------------------------------------------------
// generic GEMM
// ---------------------------------------------------------------------------------------------------------------
enum GemmEpi { EPI_F32 = 0, EPI_SILU_OP = 1, EPI_RESID_OP = 2 };

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
  float* out_f32;         // EPI_F32: row-major output
  int ldo;
  int n_valid;            // EPI_F32: number of valid output columns
  const float* rowscale;  // EPI_F32: optional per-row multiplier of the bias (AdjMatSeer: rowsum of L)
  int relu;               // EPI_F32: apply ReLU
  uint8_t* out_op;        // EPI_SILU_OP / EPI_RESID_OP: operand-format output
  int out_op_chunks;
  float* resid;           // EPI_RESID_OP: fp32 residual stream, updated in place
  int ldr;
};

template <int BN>
struct GemmCfg {
  static constexpr int NSTAGE = (BN == 448) ? 3 : 4;
  static constexpr int NPER = (BN == 448) ? 224 : BN;  // N per tcgen05.mma (<= 256)
  static constexpr int NH = BN / NPER;
  static constexpr int STAGE_BYTES = A_CHUNK_BYTES + BN * CHUNK_BYTES;
  static constexpr int TMEM_COLS = (BN > 256) ? 512 : 256;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int kMode, int BN, int kEpi>
__global__ void __launch_bounds__(192, 1) k_tc_gemm(const GemmArgs p) {
  using Cfg = GemmCfg<BN>;
  constexpr bool kFast = (kMode == PREC_BF16);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + Cfg::NSTAGE * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (Cfg::NSTAGE + s); };
  const uint32_t dfull_bar = bar0 + 8u * (2 * Cfg::NSTAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + Cfg::NSTAGE * Cfg::STAGE_BYTES + 8 * (2 * Cfg::NSTAGE + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, nt = blockIdx.y;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(dfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kc = 0; kc < p.n_kc; ++kc) {
        const int s = kc % Cfg::NSTAGE;
        const uint32_t ph = (kc / Cfg::NSTAGE) & 1;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
        const uint8_t* asrc = (kc < p.a0_chunks)
                                  ? p.a0 + ((size_t)mt * p.a0_per_tile + kc) * A_CHUNK_BYTES
                                  : p.a1 + ((size_t)mt * p.a1_per_tile + (kc - p.a0_chunks)) * A_CHUNK_BYTES;
        const uint8_t* wsrc = p.w + ((size_t)nt * p.n_kc + kc) * (size_t)(BN * CHUNK_BYTES);
        const uint32_t sa = base + s * Cfg::STAGE_BYTES;
        bulk_g2s(sa, asrc, A_CHUNK_BYTES, full_bar(s));
        bulk_g2s(sa + A_CHUNK_BYTES, wsrc, BN * CHUNK_BYTES, full_bar(s));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kMode == PREC_BF16 ? 1 : 2, TILE_M, Cfg::NPER);
      for (int kc = 0; kc < p.n_kc; ++kc) {
        const int s = kc % Cfg::NSTAGE;
        const uint32_t ph = (kc / Cfg::NSTAGE) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = base + s * Cfg::STAGE_BYTES;
        const uint64_t adesc = umma_desc_sw128(sa);
#pragma unroll
        for (int nh = 0; nh < Cfg::NH; ++nh) {
          const uint64_t bdesc = umma_desc_sw128(sa + A_CHUNK_BYTES + nh * Cfg::NPER * CHUNK_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma<kMode>(tmem_base + nh * Cfg::NPER, adesc + 2 * ks, bdesc + 2 * ks, idesc, (kc | ks) != 0);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(dfull_bar);
    }
  } else {
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31; thread = one output row
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int grow = mt * TILE_M + row;
    const bool rvalid = grow < p.m_rows;
    mbar_wait(dfull_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float rs = 1.0f;
    if (kEpi == EPI_F32 && p.rowscale != nullptr && rvalid) rs = p.rowscale[grow];
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(trow + c0, v);
      tmem_wait_ld();
      const int gcol = nt * BN + c0;
      if (rvalid) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + gcol);
        if constexpr (kEpi == EPI_F32) {
          float* orow = p.out_f32 + (size_t)grow * p.ldo + gcol;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b = __ldg(b4 + (e >> 2));
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
        } else if constexpr (kEpi == EPI_SILU_OP) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b = __ldg(b4 + (e >> 2));
            v[e + 0] = silu<kFast>(v[e + 0] + b.x);
            v[e + 1] = silu<kFast>(v[e + 1] + b.y);
            v[e + 2] = silu<kFast>(v[e + 2] + b.z);
            v[e + 3] = silu<kFast>(v[e + 3] + b.w);
          }
          op_store<kMode, 32>(p.out_op, p.out_op_chunks, grow, gcol, v);
        } else {
          float* rrow = p.resid + (size_t)grow * p.ldr + gcol;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b = __ldg(b4 + (e >> 2));
            float4 h = *reinterpret_cast<const float4*>(rrow + e);
            h.x += v[e + 0] + b.x;
            h.y += v[e + 1] + b.y;
            h.z += v[e + 2] + b.z;
            h.w += v[e + 3] + b.w;
            *reinterpret_cast<float4*>(rrow + e) = h;
            v[e + 0] = h.x; v[e + 1] = h.y; v[e + 2] = h.z; v[e + 3] = h.w;
          }
          op_store<kMode, 32>(p.out_op, p.out_op_chunks, grow, gcol, v);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// fused edge kernel
// ---------------------------------------------------------------------------------------------------------------
constexpr int EDGE_MAXG = 12;          // max target nodes (groups) per 128-row tile
constexpr int EDGE_MAXN = 39;          // max atoms per molecule (reference config.py MAX_N_NODES)
constexpr int EDGE_WSLOT = 224 * CHUNK_BYTES;  // 28,672 B: one (K chunk, N half) block of W2
constexpr int EDGE_NW = 3;             // W ring slots
constexpr int EDGE_NA = 2;             // A ring stages (also the epilogue's transposition scratch)
constexpr int EDGE_QPITCH = 452;       // floats; 1808 B rows -> conflict-free LDS.128 across consecutive j
constexpr int EDGE_THREADS = 320;      // warp 0 producer, warp 1 MMA, warps 2..9 compute

struct EdgeSmem {
  static constexpr int W_OFF = 0;
  static constexpr int A_OFF = W_OFF + EDGE_NW * EDGE_WSLOT;
  static constexpr int Q_OFF = A_OFF + EDGE_NA * A_CHUNK_BYTES;
  static constexpr int P_OFF = Q_OFF + ((EDGE_MAXN * EDGE_QPITCH * 4 + 127) / 128) * 128;
  static constexpr int VEC_OFF = P_OFF + EDGE_MAXG * HP * 4;   // wc, wd, wv
  static constexpr int DOT_OFF = VEC_OFF + 3 * HP * 4;          // [2][128] partial dots
  static constexpr int TRS_OFF = DOT_OFF + 2 * TILE_M * 4;      // [128][3] coordinate messages
  static constexpr int BAR_OFF = TRS_OFF + TILE_M * 3 * 4;
  static constexpr int TOTAL = BAR_OFF + 256;
  static constexpr int ALLOC = TOTAL + 1024;
};

struct EdgeArgs {
  const int4* tiles;   // {molecule, first target node i0, groups ng, atoms N}
  int n_tiles;
  const int* node_off; // [B+1] prefix sum of atom counts
  const float* pq;     // [nodes][896]: P = W1a.h at 0..447, Q = W1b.h + b1 at 448..895
  const float* x_cur;  // [nodes][3] coordinates at block start
  const float* x0;     // [nodes][3] coordinates at EGNN input
  float* x_next;       // equivariant update output
  const uint8_t* w2;   // packed second-layer weights [n_kc][448 x 128 B] (bias folded into K column 420)
  int n_kc;
  const float* wc;     // [448] first-layer column for d2 (W1[:,840])
  const float* wd;     // [448] first-layer column for d0^2 (W1[:,841])
  const float* wv;     // [448] attention vector (GCL) or coordinate head (equivariant update)
  float att_bias;
  uint8_t* agg_op;     // GCL: neighbour aggregate, operand format
  int agg_chunks;
};

template <int kMode, bool kEquiv>
__global__ void __launch_bounds__(EDGE_THREADS, 1) k_tc_edge(const EdgeArgs p) {
  constexpr bool kFast = (kMode == PREC_BF16);
  constexpr int EPC = epc(kMode), EPP = epp(kMode);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  float* Qs = reinterpret_cast<float*>(gbase + EdgeSmem::Q_OFF);
  float* Ps = reinterpret_cast<float*>(gbase + EdgeSmem::P_OFF);
  float* wc_s = reinterpret_cast<float*>(gbase + EdgeSmem::VEC_OFF);
  float* wd_s = wc_s + HP;
  float* wv_s = wd_s + HP;
  float* dots = reinterpret_cast<float*>(gbase + EdgeSmem::DOT_OFF);
  float* trs = reinterpret_cast<float*>(gbase + EdgeSmem::TRS_OFF);
  uint8_t* scratch = gbase + EdgeSmem::A_OFF;
  const uint32_t bar0 = base + EdgeSmem::BAR_OFF;
  auto w_full = [&](int s) { return bar0 + 8u * s; };
  auto w_empty = [&](int s) { return bar0 + 8u * (EDGE_NW + s); };
  auto a_full = [&](int s) { return bar0 + 8u * (2 * EDGE_NW + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 * EDGE_NW + EDGE_NA + s); };
  const uint32_t pq_full = bar0 + 8u * (2 * EDGE_NW + 2 * EDGE_NA);
  const uint32_t pq_empty = pq_full + 8u;
  const uint32_t d_full = pq_full + 16u;
  const uint32_t d_empty = pq_full + 24u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + EdgeSmem::BAR_OFF + 8 * (2 * EDGE_NW + 2 * EDGE_NA + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_begin = (int)(((long long)blockIdx.x * p.n_tiles) / gridDim.x);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * p.n_tiles) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < EDGE_NW; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    for (int s = 0; s < EDGE_NA; ++s) {
      mbar_init(a_full(s), 256);
      mbar_init(a_empty(s), 1);
    }
    mbar_init(pq_full, 1);
    mbar_init(pq_empty, 256);
    mbar_init(d_full, 1);
    mbar_init(d_empty, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  for (int i = threadIdx.x; i < HP; i += EDGE_THREADS) {
    wc_s[i] = p.wc[i];
    wd_s[i] = p.wd[i];
    wv_s[i] = p.wv[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== bulk-copy producer =====================
    if (lane == 0) {
      int prev_mol = -1;
      uint32_t wi = 0;
      for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
        const int4 ti = p.tiles[t];
        const int mol = ti.x, i0 = ti.y, ng = ti.z, n = ti.w;
        const int node0 = p.node_off[mol];
        mbar_wait(pq_empty, (uint32_t)((it & 1) ^ 1));
        const bool newmol = (mol != prev_mol);
        mbar_arrive_expect_tx(pq_full, (uint32_t)((ng + (newmol ? n : 0)) * HP * 4));
        for (int g = 0; g < ng; ++g)
          bulk_g2s(base + EdgeSmem::P_OFF + g * HP * 4, p.pq + (size_t)(node0 + i0 + g) * (2 * HP), HP * 4, pq_full);
        if (newmol)
          for (int j = 0; j < n; ++j)
            bulk_g2s(base + EdgeSmem::Q_OFF + j * EDGE_QPITCH * 4, p.pq + (size_t)(node0 + j) * (2 * HP) + HP, HP * 4,
                     pq_full);
        prev_mol = mol;
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
  } else if (warp == 1) {
    // ===================== tcgen05.mma issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kMode == PREC_BF16 ? 1 : 2, TILE_M, 224);
      uint32_t wi = 0, ai = 0;
      for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
        mbar_wait(d_empty, (uint32_t)((it & 1) ^ 1));
        tc_fence_after();
        for (int kc = 0; kc < p.n_kc; ++kc, ++ai) {
          const int as = ai % EDGE_NA;
          mbar_wait(a_full(as), (ai / EDGE_NA) & 1);
          const uint64_t adesc = umma_desc_sw128(base + EdgeSmem::A_OFF + as * A_CHUNK_BYTES);
          for (int nh = 0; nh < 2; ++nh, ++wi) {
            const int ws = wi % EDGE_NW;
            mbar_wait(w_full(ws), (wi / EDGE_NW) & 1);
            tc_fence_after();
            const uint64_t bdesc = umma_desc_sw128(base + EdgeSmem::W_OFF + ws * EDGE_WSLOT);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma<kMode>(tmem_base + nh * 224, adesc + 2 * ks, bdesc + 2 * ks, idesc, (kc | ks) != 0);
            umma_commit(w_empty(ws));
          }
          umma_commit(a_empty(as));
        }
        umma_commit(d_full);
      }
    }
  } else {
    // ===================== compute warps: A generation + epilogue =====================
    const int ct = threadIdx.x - 64;       // 0..255
    const int cw = ct >> 5;                // 0..7
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int hf = cw >> 2;                // which half of the K pieces / output columns
    const int r = q * 32 + lane;           // tile row = TMEM lane
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t ai = 0;
    for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
      const int4 ti = p.tiles[t];
      const int mol = ti.x, i0 = ti.y, ng = ti.z, n = ti.w;
      const int nm1 = max(n - 1, 1);
      const int nrows = ng * (n - 1);
      const int node0 = p.node_off[mol];
      const bool valid = r < nrows;
      const int g = valid ? r / nm1 : 0;
      const int jj = valid ? r - g * nm1 : 0;
      const int i = i0 + g;
      const int j = valid ? jj + (jj >= i ? 1 : 0) : 0;
      float d2, d02, ux, uy, uz;
      {
        const float* xi = p.x_cur + (size_t)(node0 + i) * 3;
        const float* xj = p.x_cur + (size_t)(node0 + j) * 3;
        const float dx = xi[0] - xj[0], dy = xi[1] - xj[1], dz = xi[2] - xj[2];
        d2 = dx * dx + dy * dy + dz * dz;
        const float inv = 1.0f / sqrtf(d2 + 1e-8f);
        ux = dx * inv; uy = dy * inv; uz = dz * inv;
        const float* yi = p.x0 + (size_t)(node0 + i) * 3;
        const float* yj = p.x0 + (size_t)(node0 + j) * 3;
        const float ex = yi[0] - yj[0], ey = yi[1] - yj[1], ez = yi[2] - yj[2];
        d02 = ex * ex + ey * ey + ez * ez;
      }
      mbar_wait(pq_full, (uint32_t)(it & 1));
      const float* Prow = Ps + g * HP;
      const float* Qrow = Qs + j * EDGE_QPITCH;

      // ---- A generation: SiLU(P_i + Q_j + d2*wc + d02*wd) -> swizzled operand chunks ----
      for (int kc = 0; kc < p.n_kc; ++kc, ++ai) {
        const int as = ai % EDGE_NA;
        mbar_wait(a_empty(as), ((ai / EDGE_NA) & 1) ^ 1u);
        uint8_t* stage = gbase + EdgeSmem::A_OFF + as * A_CHUNK_BYTES;
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const int piece = hf * 4 + pp;
          const int k0 = kc * EPC + piece * EPP;
          float a[EPP];
#pragma unroll
          for (int e = 0; e < EPP; e += 4) {
            const float4 pv = *reinterpret_cast<const float4*>(Prow + k0 + e);
            const float4 qv = *reinterpret_cast<const float4*>(Qrow + k0 + e);
            const float4 cv = *reinterpret_cast<const float4*>(wc_s + k0 + e);
            const float4 dv = *reinterpret_cast<const float4*>(wd_s + k0 + e);
            a[e + 0] = silu<kFast>(fmaf(d02, dv.x, fmaf(d2, cv.x, pv.x + qv.x)));
            a[e + 1] = silu<kFast>(fmaf(d02, dv.y, fmaf(d2, cv.y, pv.y + qv.y)));
            a[e + 2] = silu<kFast>(fmaf(d02, dv.z, fmaf(d2, cv.z, pv.z + qv.z)));
            a[e + 3] = silu<kFast>(fmaf(d02, dv.w, fmaf(d2, cv.w, pv.w + qv.w)));
          }
          constexpr int BE = BIAS_COL % EPP;
          if (k0 == BIAS_COL - BE) a[BE] = 1.0f;  // constant-1 column carrying b2
          uint4 w;
          if constexpr (kMode == PREC_BF16) {
            w.x = pack_bf16x2(a[0], a[1]); w.y = pack_bf16x2(a[2], a[3]);
            w.z = pack_bf16x2(a[4], a[5]); w.w = pack_bf16x2(a[6], a[7]);
          } else {
            w.x = f32_to_tf32(a[0]); w.y = f32_to_tf32(a[1]); w.z = f32_to_tf32(a[2]); w.w = f32_to_tf32(a[3]);
          }
          *reinterpret_cast<uint4*>(stage + sw128_offset(r, piece)) = w;
        }
        fence_proxy_async();
        mbar_arrive(a_full(as));
      }
      mbar_arrive(pq_empty);  // P/Q rows of this tile are no longer needed

      // ---- epilogue ----
      mbar_wait(d_full, (uint32_t)(it & 1));
      tc_fence_after();
      float dot = 0.f;
      for (int ch = 0; ch < 7; ++ch) {
        const int col0 = hf * 224 + ch * 32;
        float v[32];
        tmem_ld32(trow + col0, v);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 wv4 = *reinterpret_cast<const float4*>(wv_s + col0 + e);
          dot = fmaf(silu<kFast>(v[e + 0]), wv4.x, dot);
          dot = fmaf(silu<kFast>(v[e + 1]), wv4.y, dot);
          dot = fmaf(silu<kFast>(v[e + 2]), wv4.z, dot);
          dot = fmaf(silu<kFast>(v[e + 3]), wv4.w, dot);
        }
      }
      dots[hf * TILE_M + r] = dot;
      named_bar_sync(1, 256);
      const float full_dot = dots[r] + dots[TILE_M + r];
      if constexpr (kEquiv) {
        // x_i += sum_j unit_ij * phi_ij / 100   (reference egnn.py:124-134)
        tc_fence_before();
        mbar_arrive(d_empty);
        if (hf == 0) {
          const float phi = valid ? full_dot : 0.f;
          trs[r * 3 + 0] = ux * phi;
          trs[r * 3 + 1] = uy * phi;
          trs[r * 3 + 2] = uz * phi;
        }
        named_bar_sync(1, 256);
        if (ct < ng * 3) {
          const int gg = ct / 3, c = ct - gg * 3;
          float s = 0.f;
          for (int e = 0; e < n - 1; ++e) s += trs[(gg * nm1 + e) * 3 + c];
          const size_t idx = (size_t)(node0 + i0 + gg) * 3 + c;
          p.x_next[idx] = p.x_cur[idx] + s / 100.0f;
        }
        named_bar_sync(1, 256);
      } else {
        // e_ij = m_ij * sigmoid(w_a.m_ij + b_a); agg_i = sum_j e_ij / 100   (reference egnn.py:48-51, 59-64)
        const float gate = valid ? sigmoid_acc(full_dot + p.att_bias) : 0.f;
        for (int ch = 0; ch < 7; ++ch) {
          const int col0 = hf * 224 + ch * 32;
          float v[32];
          tmem_ld32(trow + col0, v);
          tmem_wait_ld();
          if (ch == 6) {
            tc_fence_before();
            mbar_arrive(d_empty);  // last TMEM read of this tile is complete
          }
          named_bar_sync(1, 256);  // previous chunk's readers are done with the scratch
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            float4 o;
            o.x = silu<kFast>(v[e + 0]) * gate;
            o.y = silu<kFast>(v[e + 1]) * gate;
            o.z = silu<kFast>(v[e + 2]) * gate;
            o.w = silu<kFast>(v[e + 3]) * gate;
            *reinterpret_cast<float4*>(scratch + hf * A_CHUNK_BYTES + sw128_offset(r, e >> 2)) = o;
          }
          named_bar_sync(1, 256);
          // segment sum over neighbours j in ascending order (the CPU reference's scatter_add order)
          for (int pr = cw; pr < 2 * ng; pr += 8) {
            const int hh = pr & 1, gg = pr >> 1;
            const uint8_t* src = scratch + hh * A_CHUNK_BYTES + (lane & 3) * 4;
            float s = 0.f;
            const int r0 = gg * nm1;
            for (int e = 0; e < n - 1; ++e)
              s += *reinterpret_cast<const float*>(src + sw128_offset(r0 + e, lane >> 2));
            op_store1<kMode>(p.agg_op, p.agg_chunks, node0 + i0 + gg, hh * 224 + ch * 32 + lane, s / 100.0f);
          }
        }
        named_bar_sync(1, 256);  // scratch (= A ring) is free again before the next tile's A generation
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace mlcg
