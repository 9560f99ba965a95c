// Non-tensor-core kernels of the hot path: diffusion-step math (HBM-bound), EGNN prepare / readout, weight
// packing into operand format, the exact-fp32 SIMT path (precision 0, used as the on-GPU fp32 reference and for
// debugging), and the AdjMatSeer helper kernels.
#pragma once
#include "mlcg_common.cuh"

namespace mlcg {

// =================================================================================================================
// Diffusion-step kernels: one warp per molecule, lanes own atoms (lane, lane+32).  N <= 64.
// =================================================================================================================
struct NoiseSrc {
  const float* raw;            // [B][N][11] raw N(0,1) draws injected by the host (parity mode) or nullptr
  unsigned long long seed;     // device Philox otherwise: keyed by (seed, global sample id, atom, draw index)
  unsigned long long draw;
  long long sample_offset;     // global id of sample 0 of this shard when `ids` is null (contiguous shard)
  const long long* ids;        // optional device array [B] of global sample ids (any sharding reproduces the unsharded run)
  const unsigned long long* ctl;  // optional device pointer to {seed, sample_offset}: overrides the two fields above so a
                                  // captured CUDA graph can be replayed with new seeds
};

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), written out so that the
// (key, counter) assignment is explicit and can be restated on the host (oracle/philox_oracle.py):
//   key     = {seed lo, seed hi}
//   counter = {3 * draw + k, atom, sample id lo, sample id hi},  k = 0, 1, 2  -> 12 uint32 per (sample, atom, draw)
// Every (sample, atom, draw) owns three counter values of its own, so no two draws share a random bit.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// Box-Muller on two uint32: u = (a + 0.5) / 2^32 in (0, 1], v = (b + 0.5) / 2^32; (r sin 2 pi v, r cos 2 pi v)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float v = fmaf((float)b, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float r = sqrtf(-2.0f * logf(u));
  float sn, cs;
  sincospif(2.0f * v, &sn, &cs);
  return make_float2(r * sn, r * cs);
}

// Raw draws for atom i of sample b (11 values: 3 position + 8 feature; reference order
// equivariant_diffusion.py:347-362).
__device__ __forceinline__ void raw_noise(const NoiseSrc& ns, int b, int i, int N, float* out) {
  if (ns.raw != nullptr) {
    const float* src = ns.raw + ((size_t)b * N + i) * ZC;
#pragma unroll
    for (int c = 0; c < ZC; ++c) out[c] = src[c];
  } else {
    const unsigned long long seed = ns.ctl ? ns.ctl[0] : ns.seed;
    const unsigned long long id = ns.ids ? (unsigned long long)ns.ids[b]
                                         : (unsigned long long)((ns.ctl ? (long long)ns.ctl[1] : ns.sample_offset) + b);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    float v[12];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)(3ull * ns.draw + k), (uint32_t)i, (uint32_t)id, (uint32_t)(id >> 32)), key);
      const float2 a = box_muller(r.x, r.y), c = box_muller(r.z, r.w);
      v[4 * k] = a.x; v[4 * k + 1] = a.y; v[4 * k + 2] = c.x; v[4 * k + 3] = c.y;
    }
#pragma unroll
    for (int c = 0; c < ZC; ++c) out[c] = v[c];
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Masked, centre-of-gravity-free combined noise (reference equivariant_diffusion.py:56-76, 341-363) for the two
// atoms owned by this lane.  eps[s][c], s = 0,1.
__device__ __forceinline__ void combined_noise(const NoiseSrc& ns, int b, int n, int N, int lane, float eps[2][ZC]) {
  float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    if (i < n) {
      raw_noise(ns, b, i, N, eps[s]);
      sx += eps[s][0]; sy += eps[s][1]; sz += eps[s][2];
    } else {
#pragma unroll
      for (int c = 0; c < ZC; ++c) eps[s][c] = 0.f;
    }
  }
  // reference: mean = sum / n ; x - mean * mask
  const float fn = (float)n;
  sx = __fdiv_rn(warp_sum(sx), fn); sy = __fdiv_rn(warp_sum(sy), fn); sz = __fdiv_rn(warp_sum(sz), fn);
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (lane + 32 * s < n) { eps[s][0] -= sx; eps[s][1] -= sy; eps[s][2] -= sz; }
  }
}

// z <- combined noise   (reference equivariant_diffusion.py:384)
__global__ void k_noise_init(float* __restrict__ z, const int* __restrict__ n_nodes, int B, int N, NoiseSrc ns) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  float eps[2][ZC];
  combined_noise(ns, b, n_nodes[b], N, lane, eps);
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    if (i < N) {
      float* dst = z + ((size_t)b * N + i) * ZC;
#pragma unroll
      for (int c = 0; c < ZC; ++c) dst[c] = eps[s][c];
    }
  }
}

// z_s = z_t/alpha_ts - c_eps*eps + c_sigma*noise, then centre-of-gravity removal of the position part
// (reference equivariant_diffusion.py:320-338).
__global__ void k_step(float* __restrict__ z, const float* __restrict__ eps_net, const int* __restrict__ n_nodes, int B,
                       int N, float alpha_ts, float c_eps, float c_sigma, NoiseSrc ns) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const int n = n_nodes[b];
  float eps[2][ZC], zn[2][ZC];
  combined_noise(ns, b, n, N, lane, eps);
  float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    if (i < n) {
      const float* zp = z + ((size_t)b * N + i) * ZC;
      const float* ep = eps_net + ((size_t)b * N + i) * ZC;
#pragma unroll
      for (int c = 0; c < ZC; ++c) {
        const float mu = __fsub_rn(__fdiv_rn(zp[c], alpha_ts), __fmul_rn(c_eps, ep[c]));
        zn[s][c] = __fadd_rn(mu, __fmul_rn(c_sigma, eps[s][c]));
      }
      sx += zn[s][0]; sy += zn[s][1]; sz += zn[s][2];
    } else {
#pragma unroll
      for (int c = 0; c < ZC; ++c) zn[s][c] = 0.f;
    }
  }
  const float fn = (float)n;
  sx = __fdiv_rn(warp_sum(sx), fn); sy = __fdiv_rn(warp_sum(sy), fn); sz = __fdiv_rn(warp_sum(sz), fn);
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    if (i < N) {
      float* zp = z + ((size_t)b * N + i) * ZC;
      if (i < n) { zn[s][0] -= sx; zn[s][1] -= sy; zn[s][2] -= sz; }
#pragma unroll
      for (int c = 0; c < ZC; ++c) zp[c] = zn[s][c];
    }
  }
}

// Fragment re-injection (reference equivariant_diffusion.py:473-493 and 79-105): forward-diffuse z_known to the
// current level, move its fixed-atom centre of mass onto the generated one, blend into z over the fixed atoms.
__global__ void k_reinject(float* __restrict__ z, const float* __restrict__ z_known, const float* __restrict__ fixed_mask,
                           const int* __restrict__ n_nodes, int B, int N, float alpha_s, float sigma_s, float blend,
                           NoiseSrc ns) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const int n = n_nodes[b];
  float eps[2][ZC], zk[2][ZC], zg[2][ZC], fm[2];
  combined_noise(ns, b, n, N, lane, eps);
  float cg[3] = {0.f, 0.f, 0.f}, ck[3] = {0.f, 0.f, 0.f}, cnt = 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    fm[s] = 0.f;
    if (i < N) {
      const size_t o = ((size_t)b * N + i) * ZC;
      fm[s] = fixed_mask[(size_t)b * N + i];
#pragma unroll
      for (int c = 0; c < ZC; ++c) {
        zg[s][c] = z[o + c];
        zk[s][c] = __fadd_rn(__fmul_rn(alpha_s, z_known[o + c]), __fmul_rn(sigma_s, eps[s][c]));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) { cg[c] += zg[s][c] * fm[s]; ck[c] += zk[s][c] * fm[s]; }
      cnt += fm[s];
    }
  }
  cnt = warp_sum(cnt);
  float shift[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) shift[c] = __fsub_rn(__fdiv_rn(warp_sum(cg[c]), cnt), __fdiv_rn(warp_sum(ck[c]), cnt));
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    if (i < N) {
      const size_t o = ((size_t)b * N + i) * ZC;
#pragma unroll
      for (int c = 0; c < ZC; ++c) {
        float k = zk[s][c];
        if (c < 3) k = __fadd_rn(k, __fmul_rn(shift[c], fm[s]));
        const float a = __fmul_rn(__fmul_rn(blend, k), fm[s]);
        const float bb = __fmul_rn(__fmul_rn(__fsub_rn(1.0f, blend), zg[s][c]), fm[s]);
        const float cc = __fmul_rn(zg[s][c], __fsub_rn(1.0f, fm[s]));
        z[o + c] = __fadd_rn(__fadd_rn(a, bb), cc);
      }
    }
  }
}

// z = alpha*z_known + sigma*noise  (merge_fragments' initial forward diffusion, reference :549-559)
__global__ void k_forward_diffuse(float* __restrict__ z, const float* __restrict__ z_known, const int* __restrict__ n_nodes,
                                  int B, int N, float alpha, float sigma, NoiseSrc ns) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  float eps[2][ZC];
  combined_noise(ns, b, n_nodes[b], N, lane, eps);
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    if (i < N) {
      const size_t o = ((size_t)b * N + i) * ZC;
#pragma unroll
      for (int c = 0; c < ZC; ++c) z[o + c] = __fadd_rn(__fmul_rn(alpha, z_known[o + c]), __fmul_rn(sigma, eps[s][c]));
    }
  }
}

// Final decode (reference equivariant_diffusion.py:261-285): x = (z0 - sigma0*eps)/alpha0 + sigma_x*noise;
// atom class = argmax over z0[:, 3:10] (7 channels), padded atoms -> -1.
__global__ void k_decode(const float* __restrict__ z0, const float* __restrict__ eps_net, const int* __restrict__ n_nodes,
                         int B, int N, float sigma0, float alpha0, float sigma_x, NoiseSrc ns, float* __restrict__ x_out,
                         int* __restrict__ cls_out, int* __restrict__ nonfinite) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const int n = n_nodes[b];
  float eps[2][ZC];
  combined_noise(ns, b, n, N, lane, eps);
  const float inv_alpha = __fdiv_rn(1.0f, alpha0);
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    if (i < N) {
      const size_t o = ((size_t)b * N + i) * ZC;
      float xo[3] = {0.f, 0.f, 0.f};
      int cls = -1;
      if (i < n) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float mu = __fmul_rn(inv_alpha, __fsub_rn(z0[o + c], __fmul_rn(sigma0, eps_net[o + c])));
          xo[c] = __fadd_rn(mu, __fmul_rn(sigma_x, eps[s][c]));
          // a diverged trajectory (or, in fp16 mode, one that left the fp16 range) must not pass silently
          if (nonfinite != nullptr && !isfinite(xo[c])) atomicOr(nonfinite, 1);
        }
        float best = z0[o + 3] * 9.0f;
        cls = 0;
#pragma unroll
        for (int c = 1; c < 7; ++c) {
          const float v = z0[o + 3 + c] * 9.0f;
          if (v > best) { best = v; cls = c; }
        }
      }
      x_out[((size_t)b * N + i) * 3 + 0] = xo[0];
      x_out[((size_t)b * N + i) * 3 + 1] = xo[1];
      x_out[((size_t)b * N + i) * 3 + 2] = xo[2];
      cls_out[(size_t)b * N + i] = cls;
    }
  }
}

// =================================================================================================================
// EGNN prepare (mask, split, time/context append, 12->420 embedding) and readout (420->8, velocity, COM removal)
// reference egnn.py:472-513, 315, 398-399
// =================================================================================================================
constexpr int PREP_NPB = 8;  // nodes per block of k_egnn_prepare
template <int kMode>
__global__ void __launch_bounds__(HP) k_egnn_prepare(const float* __restrict__ z, const float* __restrict__ t,
                                                      const float* __restrict__ ctx, const int* __restrict__ node_mol,
                                                      const int* __restrict__ node_off, int N, int M,
                                                      const float* __restrict__ w_emb, const float* __restrict__ b_emb,
                                                      float* __restrict__ h_res, int ldh, uint8_t* __restrict__ h_op,
                                                      int op_chunks, float* __restrict__ x0, float* __restrict__ x_cur) {
  const int c = threadIdx.x, nbase = blockIdx.x * PREP_NPB;
  __shared__ float hin[PREP_NPB][IN_NF];
  if (c < PREP_NPB * 15) {  // thread = (node of the block, one of its 12 inputs or 3 coordinates)
    const int nl = c / 15, k = c - nl * 15, node = nbase + nl;
    if (node < M) {
      const int b = node_mol[node], i = node - node_off[b];
      const float* zp = z + ((size_t)b * N + i) * ZC;
      if (k < 8) hin[nl][k] = zp[3 + k];
      else if (k == 8) hin[nl][8] = t[b];
      else if (k < 12) hin[nl][k] = ctx[b * 3 + (k - 9)];
      else {
        const float v = zp[k - 12];
        x0[(size_t)node * 3 + (k - 12)] = v;
        x_cur[(size_t)node * 3 + (k - 12)] = v;
      }
    }
  }
  __syncthreads();
  float w[IN_NF], bias = 0.f;  // this channel's embedding row stays in registers for the block's nodes
#pragma unroll
  for (int k = 0; k < IN_NF; ++k) w[k] = (c < HID) ? w_emb[c * IN_NF + k] : 0.f;
  if (c < HID) bias = b_emb[c];
  for (int nl = 0; nl < PREP_NPB; ++nl) {
    const int node = nbase + nl;
    if (node >= M) break;
    float acc = 0.f;
    if (c < HID) {
      acc = bias;
#pragma unroll
      for (int k = 0; k < IN_NF; ++k) acc = fmaf(w[k], hin[nl][k], acc);
    }
    h_res[hres_index(node, c, ldh)] = acc;
    if constexpr (kMode != PREC_FP32_SIMT) op_store1<kMode>(h_op, op_chunks, node, c, acc * op_scale(kMode));
  }
}

__global__ void __launch_bounds__(256) k_egnn_readout(const float* __restrict__ h_res, int ldh, const float* __restrict__ x_fin,
                                                       const float* __restrict__ x0, const int* __restrict__ n_nodes,
                                                       const int* __restrict__ node_off, int N, const float* __restrict__ w_out,
                                                       const float* __restrict__ b_out, float* __restrict__ eps) {
  const int b = blockIdx.x, n = n_nodes[b], node0 = node_off[b];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ float vel[64 * 3];
  __shared__ float mean[3];
  for (int idx = threadIdx.x; idx < N * 3; idx += blockDim.x) {
    const int i = idx / 3;
    vel[idx] = (i < n) ? x_fin[(size_t)node0 * 3 + idx] - x0[(size_t)node0 * 3 + idx] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += vel[i * 3 + threadIdx.x];
    mean[threadIdx.x] = s / (float)n;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * 3; idx += blockDim.x) {
    const int i = idx / 3, c = idx - i * 3;
    eps[((size_t)b * N + i) * ZC + c] = (i < n) ? vel[idx] - mean[c] : 0.f;
  }
  // class channels: only the first 8 of the 12 embedding_out rows reach eps (reference egnn.py:503-505).
  // One warp per atom: its h row is loaded once (14 independent loads per lane), the 8 output rows come from shared memory.
  __shared__ float wo[8 * HID];
  for (int idx = threadIdx.x; idx < 8 * HID; idx += blockDim.x) wo[idx] = w_out[idx];
  __syncthreads();
  const int nwarps = blockDim.x >> 5;
  for (int i = warp; i < N; i += nwarps) {
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
    if (i < n) {
      float hv[(HID + 31) / 32];
#pragma unroll
      for (int kk = 0; kk < (HID + 31) / 32; ++kk) {
        const int k = lane + 32 * kk;
        hv[kk] = (k < HID) ? h_res[hres_index(node0 + i, k, ldh)] : 0.f;
      }
#pragma unroll
      for (int o = 0; o < 8; ++o) {
#pragma unroll
        for (int kk = 0; kk < (HID + 31) / 32; ++kk) {
          const int k = lane + 32 * kk;
          if (k < HID) acc[o] = fmaf(wo[o * HID + k], hv[kk], acc[o]);  // same order as a per-lane loop k = lane, lane+32, ...
        }
        acc[o] = warp_sum(acc[o]) + b_out[o];
      }
    }
    if (lane < 8) {
      float v = 0.f;
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (lane == o) v = acc[o];
      eps[((size_t)b * N + i) * ZC + 3 + lane] = v;
    }
  }
}

// =================================================================================================================
// Weight packing into operand format (once, at load time)
// dst[n_tile][kc][BN rows x 128 B swizzled]; element (n, k) = W[n_src_off + n][ksrc(k)] or bias[n] at k == bias_k
// K is split in segments of seg_len elements; segment s reads source columns kofs[s] .. kofs[s]+kreal[s]-1.
// =================================================================================================================
struct PackArgs {
  const float* src; int ld; int n_real; int n_src_off;
  int bn; int n_kc;
  int seg_len; int kreal0; int kofs0; int kreal1; int kofs1;
  const float* bias; int bias_k;
  uint8_t* dst;
  float scale;  // multiplies every packed value (0.5 folds SiLU's half-argument into the weights; 0 is treated as 1)
  int split3;   // tf32 only: error-compensated split.  The K axis is tripled, W' = [W_hi | W_hi | W_lo] with W_hi = tf32(W),
                // W_lo = tf32(W - W_hi), to be multiplied with activations laid out as A' = [A_hi | A_lo | A_hi]:
                // A'.W'^T = A_hi.W_hi + A_lo.W_hi + A_hi.W_lo, i.e. the product to ~2^-21 instead of 2^-11 ("3xTF32")
};
template <int kMode>
__global__ void k_pack_weight(const PackArgs a) {
  constexpr int EPC = epc(kMode), EPP = epp(kMode);
  const int kc = blockIdx.x, nt = blockIdx.y;
  uint8_t* blk = a.dst + ((size_t)nt * a.n_kc + kc) * ((size_t)a.bn * CHUNK_BYTES);
  for (int idx = threadIdx.x; idx < a.bn * 8; idx += blockDim.x) {
    const int row = idx >> 3, piece = idx & 7;
    const int n = nt * a.bn + row;
    float v[8];
#pragma unroll
    for (int e = 0; e < EPP; ++e) {
      const int k = kc * EPC + piece * EPP + e;
      const int seg = k / a.seg_len, kk = k - seg * a.seg_len;
      const int kreal = (seg == 0 || a.split3) ? a.kreal0 : a.kreal1;
      const int kofs = (seg == 0 || a.split3) ? a.kofs0 : a.kofs1;
      float x = 0.f;
      if (n < a.n_real) {
        if ((seg < 2 || (a.split3 && seg < 3)) && kk < kreal) x = a.src[(size_t)(a.n_src_off + n) * a.ld + kofs + kk];
        if (a.bias != nullptr && k == a.bias_k) x = a.bias[a.n_src_off + n];
      }
      x *= (a.scale != 0.f ? a.scale : 1.0f);
      if (a.split3 && seg == 2) x -= __uint_as_float(f32_to_tf32(x));  // low part; rounded to tf32 below
      v[e] = x;
    }
    uint4 w;
    if constexpr (is16(kMode)) {
      w.x = pack_h2<kMode>(v[0], v[1]); w.y = pack_h2<kMode>(v[2], v[3]);
      w.z = pack_h2<kMode>(v[4], v[5]); w.w = pack_h2<kMode>(v[6], v[7]);
    } else {
      w.x = f32_to_tf32(v[0]); w.y = f32_to_tf32(v[1]); w.z = f32_to_tf32(v[2]); w.w = f32_to_tf32(v[3]);
    }
    *reinterpret_cast<uint4*>(blk + sw128_offset(row, piece)) = w;
  }
}

// copy a strided column / vector into a zero-padded fp32 vector of length n_pad
__global__ void k_pad_vector(const float* __restrict__ src, int stride, int n_real, float* __restrict__ dst, int n_pad,
                             float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = (i < n_real) ? src[(size_t)i * stride] * scale : 0.f;
}

// fp32 row-major [rows][ld] -> operand format (used by the AdjMatSeer L-multiply output and tests)
// split3 (tf32 only): the destination has 3 * n_chunks chunks per tile laid out [A_hi | A_lo | A_hi] (see PackArgs::split3)
template <int kMode>
__global__ void k_rowmajor_to_op(const float* __restrict__ src, int ld, int rows, int k_real, uint8_t* __restrict__ dst,
                                 int n_chunks, int split3 = 0) {
  constexpr int EPP = epp(kMode), EPC = epc(kMode);
  const int pieces_per_row = n_chunks * 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * pieces_per_row) return;
  const int row = (int)(idx / pieces_per_row), pc = (int)(idx % pieces_per_row);
  const int k0 = (pc >> 3) * EPC + (pc & 7) * EPP;
  float v[8];
#pragma unroll
  for (int e = 0; e < EPP; ++e) v[e] = (k0 + e < k_real) ? src[(size_t)row * ld + k0 + e] : 0.f;
  if (!split3) {
    op_store<kMode, EPP>(dst, n_chunks, row, k0, v);
  } else {
    float lo[8];
#pragma unroll
    for (int e = 0; e < EPP; ++e) lo[e] = v[e] - __uint_as_float(f32_to_tf32(v[e]));
    op_store<kMode, EPP>(dst, 3 * n_chunks, row, k0, v);
    op_store<kMode, EPP>(dst, 3 * n_chunks, row, k0 + n_chunks * EPC, lo);
    op_store<kMode, EPP>(dst, 3 * n_chunks, row, k0 + 2 * n_chunks * EPC, v);
  }
}

// =================================================================================================================
// Exact fp32 SIMT path (precision 0)
// =================================================================================================================
// C[M x N] = act( (acc ? C : 0) + A[M x K] . W[N x K]^T + rowscale*bias )   64x64 tiles, 256 threads, 4x4 / thread
enum { SG_BIAS = 1, SG_SILU = 2, SG_RELU = 4, SG_ACC = 8 };
__global__ void __launch_bounds__(256) k_simt_gemm(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                                    const float* __restrict__ bias, const float* __restrict__ rowscale,
                                                    float* __restrict__ C, int ldc, int M, int N, int K, int flags) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
      const int rr = idx >> 4, kk = idx & 15;
      const int gm = m0 + rr, gn = n0 + rr, gk = k0 + kk;
      As[kk][rr] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : 0.f;
      Ws[kk][rr] = (gn < N && gk < K) ? W[(size_t)gn * ldw + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = As[kk][ty * 4 + u]; w[u] = Ws[kk][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], w[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int gm = m0 + ty * 4 + u;
    if (gm >= M) continue;
    const float rs = rowscale ? rowscale[gm] : 1.0f;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int gn = n0 + tx * 4 + v;
      if (gn >= N) continue;
      float o = acc[u][v];
      if (flags & SG_BIAS) o += rs * bias[gn];
      if (flags & SG_ACC) o += C[(size_t)gm * ldc + gn];
      if (flags & SG_SILU) o = silu_ref(o);
      if (flags & SG_RELU) o = fmaxf(o, 0.f);
      C[(size_t)gm * ldc + gn] = o;
    }
  }
}

// A1[edge row][k] = SiLU(P_i[k] + Q_j[k] + d2*W1[k][840] + d02*W1[k][841]); one block per target node
__global__ void __launch_bounds__(HP) k_simt_build_a1(const float* __restrict__ pq, const float* __restrict__ x_cur,
                                                       const float* __restrict__ x0, const int* __restrict__ node_mol,
                                                       const int* __restrict__ node_off, const int* __restrict__ node_edge_off,
                                                       int node_begin, int edge_base, const float* __restrict__ w1,
                                                       float* __restrict__ a1) {
  const int node = node_begin + blockIdx.x, k = threadIdx.x;
  const int b = node_mol[node], node0 = node_off[b], n = node_off[b + 1] - node0, i = node - node0;
  const float pi = pq[(size_t)node * (2 * HP) + k];
  const float wc = (k < HID) ? w1[(size_t)k * (2 * HID + 2) + 2 * HID] : 0.f;
  const float wd = (k < HID) ? w1[(size_t)k * (2 * HID + 2) + 2 * HID + 1] : 0.f;
  const float xi0 = x_cur[(size_t)node * 3], xi1 = x_cur[(size_t)node * 3 + 1], xi2 = x_cur[(size_t)node * 3 + 2];
  const float yi0 = x0[(size_t)node * 3], yi1 = x0[(size_t)node * 3 + 1], yi2 = x0[(size_t)node * 3 + 2];
  int row = node_edge_off[node] - edge_base;
  for (int j = 0; j < n; ++j) {
    if (j == i) continue;
    const size_t nj = (size_t)(node0 + j);
    const float dx = xi0 - x_cur[nj * 3], dy = xi1 - x_cur[nj * 3 + 1], dz = xi2 - x_cur[nj * 3 + 2];
    const float ex = yi0 - x0[nj * 3], ey = yi1 - x0[nj * 3 + 1], ez = yi2 - x0[nj * 3 + 2];
    const float d2 = dx * dx + dy * dy + dz * dz, d02 = ex * ex + ey * ey + ez * ez;
    const float pre = pi + pq[nj * (2 * HP) + HP + k] + d2 * wc + d02 * wd;
    a1[(size_t)row * HP + k] = (k < HID) ? silu_ref(pre) : 0.f;
    ++row;
  }
}

// gate + aggregate (GCL) or coordinate update (equivariant) from M2 = SiLU(A1.W2^T + b2); one block per target node
template <bool kEquiv>
__global__ void __launch_bounds__(HP) k_simt_gate_agg(const float* __restrict__ m2, const float* __restrict__ wv, float att_bias,
                                                       const int* __restrict__ node_mol, const int* __restrict__ node_off,
                                                       const int* __restrict__ node_edge_off, int node_begin, int edge_base,
                                                       float* __restrict__ agg, int ldagg, const float* __restrict__ x_cur,
                                                       float* __restrict__ x_next) {
  const int node = node_begin + blockIdx.x, c = threadIdx.x;
  const int warp = c >> 5, lane = c & 31;
  const int b = node_mol[node], node0 = node_off[b], n = node_off[b + 1] - node0, i = node - node0;
  const int row0 = node_edge_off[node] - edge_base;
  __shared__ float gate[64];
  for (int e = warp; e < n - 1; e += HP / 32) {
    float acc = 0.f;
    for (int k = lane; k < HID; k += 32) acc = fmaf(m2[(size_t)(row0 + e) * HP + k], wv[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) gate[e] = kEquiv ? acc : 1.0f / (1.0f + expf(-(acc + att_bias)));
  }
  __syncthreads();
  if constexpr (kEquiv) {
    if (c < 3) {
      float s = 0.f;
      int e = 0;
      for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        const size_t nj = (size_t)(node0 + j);
        const float dx = x_cur[(size_t)node * 3] - x_cur[nj * 3], dy = x_cur[(size_t)node * 3 + 1] - x_cur[nj * 3 + 1],
                    dz = x_cur[(size_t)node * 3 + 2] - x_cur[nj * 3 + 2];
        const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz + 1e-8f);
        const float u = (c == 0 ? dx : (c == 1 ? dy : dz)) * inv;
        s += u * gate[e];
        ++e;
      }
      x_next[(size_t)node * 3 + c] = x_cur[(size_t)node * 3 + c] + s / 100.0f;
    }
  } else {
    float s = 0.f;
    for (int e = 0; e < n - 1; ++e) s += m2[(size_t)(row0 + e) * HP + c] * gate[e];
    agg[(size_t)node * ldagg + c] = (c < HID) ? s / 100.0f : 0.f;
  }
}

// =================================================================================================================
// AdjMatSeer helpers (reference adj_mat_seer.py:104-165, mol_utils.py:159-191, 210-211)
// =================================================================================================================
__constant__ int c_atomic_numbers[8] = {6, 7, 8, 9, 15, 16, 17, 35};
__constant__ float c_cov_radii[8] = {0.76f, 0.71f, 0.66f, 0.57f, 1.07f, 1.05f, 1.02f, 1.20f};

// Tensor layout of prepare_adj_mat_seer_input built on device from generated samples (declared connectivity rule:
// d <= 1.3*(Rcov_i+Rcov_j); parity unpinned against RDKit).  One block per molecule.
__global__ void k_seer_inputs(const float* __restrict__ x, const int* __restrict__ cls, const int* __restrict__ n_nodes, int N,
                              int* __restrict__ elements, float* __restrict__ dist, float* __restrict__ adj) {
  const int b = blockIdx.x, n = n_nodes[b];
  __shared__ float xs[SEER_D * 3];
  __shared__ int cs[SEER_D];
  for (int i = threadIdx.x; i < SEER_D; i += blockDim.x) {
    const bool real = i < n && i < N;
    cs[i] = real ? cls[(size_t)b * N + i] : -1;
    for (int c = 0; c < 3; ++c) xs[i * 3 + c] = real ? x[((size_t)b * N + i) * 3 + c] : 0.f;
    elements[(size_t)b * SEER_D + i] = real ? c_atomic_numbers[max(cs[i], 0) & 7] : 0;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < SEER_D * SEER_D; idx += blockDim.x) {
    const int i = idx / SEER_D, j = idx - i * SEER_D;
    float d = 0.f, a = 0.f;
    if (cs[i] >= 0 && cs[j] >= 0) {
      const float dx = xs[i * 3] - xs[j * 3], dy = xs[i * 3 + 1] - xs[j * 3 + 1], dz = xs[i * 3 + 2] - xs[j * 3 + 2];
      d = sqrtf(dx * dx + dy * dy + dz * dz);
      a = (d <= 1.3f * (c_cov_radii[cs[i] & 7] + c_cov_radii[cs[j] & 7])) ? 1.f : 0.f;
    }
    if (i == j) { d += 1.0f; a = 1.0f; }
    dist[(size_t)b * SEER_D * SEER_D + idx] = d;
    adj[(size_t)b * SEER_D * SEER_D + idx] = a;
  }
}

// L = D^-1/2 A D^-1/2 (GraphConv.l_norm, adj_mat_seer.py:32-41) and rowsum(L) (needed because the bias is added
// before the L-multiply: L.(XW^T + 1b^T) = (LX)W^T + rowsum(L) b^T).  One block per molecule.
__global__ void k_lnorm(const float* __restrict__ a, float* __restrict__ l, float* __restrict__ lrow) {
  const int b = blockIdx.x;
  __shared__ float inv[SEER_D];
  __shared__ float ls[SEER_D * SEER_D];
  const float* ab = a + (size_t)b * SEER_D * SEER_D;
  if (threadIdx.x < SEER_D) {
    float s = 0.f;
    for (int j = 0; j < SEER_D; ++j) s += ab[threadIdx.x * SEER_D + j];
    inv[threadIdx.x] = rsqrtf(fmaxf(s, 1e-12f));
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < SEER_D * SEER_D; idx += blockDim.x) {
    const int i = idx / SEER_D, j = idx - i * SEER_D;
    const float v = inv[i] * ab[idx] * inv[j];
    ls[idx] = v;
    l[(size_t)b * SEER_D * SEER_D + idx] = v;
  }
  __syncthreads();
  if (threadIdx.x < SEER_D) {
    float s = 0.f;
    for (int j = 0; j < SEER_D; ++j) s += ls[threadIdx.x * SEER_D + j];
    lrow[(size_t)b * SEER_D + threadIdx.x] = s;
  }
}

// X0[b, slot, :] = table[elements[b, slot], :]  (+ optional additive term), fp32 row-major [B*42][64]
__global__ void k_seer_embed(const int* __restrict__ elements, const float* __restrict__ table, const float* __restrict__ add,
                             float* __restrict__ out, int rows) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * SEER_E) return;
  const int row = idx / SEER_E, c = idx - row * SEER_E;
  float v = table[(size_t)elements[row] * SEER_E + c];
  if (add != nullptr) v += add[idx];
  out[idx] = v;
}

// Y[b] = L[b] . X[b]  (42x42 times 42xC), output either operand format (tensor-core modes) or fp32 row-major.
// grid (B, C/128), 128 threads: thread = one column, 42 accumulators.
template <int kMode>
__global__ void __launch_bounds__(128) k_lmul(const float* __restrict__ l, const float* __restrict__ x, int C, float* __restrict__ y_f32,
                                               uint8_t* __restrict__ y_op, int op_chunks, int split3 = 0) {
  const int b = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x;
  __shared__ float ls[SEER_D * SEER_D];
  for (int idx = threadIdx.x; idx < SEER_D * SEER_D; idx += 128) ls[idx] = l[(size_t)b * SEER_D * SEER_D + idx];
  __syncthreads();
  if (c >= C) return;
  float acc[SEER_D];
#pragma unroll
  for (int i = 0; i < SEER_D; ++i) acc[i] = 0.f;
  for (int s = 0; s < SEER_D; ++s) {
    const float xv = x[((size_t)b * SEER_D + s) * C + c];
#pragma unroll
    for (int i = 0; i < SEER_D; ++i) acc[i] = fmaf(ls[i * SEER_D + s], xv, acc[i]);
  }
#pragma unroll
  for (int i = 0; i < SEER_D; ++i) {
    const int row = b * SEER_D + i;
    if constexpr (kMode == PREC_FP32_SIMT) {
      y_f32[(size_t)row * C + c] = acc[i];
    } else if (!split3) {
      op_store1<kMode>(y_op, op_chunks, row, c, acc[i]);
    } else {  // [A_hi | A_lo | A_hi], op_chunks = chunks of ONE part
      const int kp = op_chunks * epc(kMode);
      const float lo = acc[i] - __uint_as_float(f32_to_tf32(acc[i]));
      op_store1<kMode>(y_op, 3 * op_chunks, row, c, acc[i]);
      op_store1<kMode>(y_op, 3 * op_chunks, row, c + kp, lo);
      op_store1<kMode>(y_op, 3 * op_chunks, row, c + 2 * kp, acc[i]);
    }
  }
}

// emb[row] = dm_resize(conv3_dm[row])  (2048 -> 1), one warp per row
__global__ void k_seer_bottleneck(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ emb, int rows) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc = 0.f;
  for (int k = lane; k < SEER_H; k += 32) acc = fmaf(x[(size_t)row * SEER_H + k], w[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) emb[row] = acc + bias[0];
}

// add[b, slot, e] = nodes_coord_fc(emb[b, :])[slot*64 + e]   (42 -> 2688; mixes all 42 slots incl. padding)
__global__ void k_seer_coord_fc(const float* __restrict__ emb, const float* __restrict__ w, const float* __restrict__ bias,
                                float* __restrict__ out, int B) {
  const int b = blockIdx.x;
  __shared__ float es[SEER_D];
  if (threadIdx.x < SEER_D) es[threadIdx.x] = emb[(size_t)b * SEER_D + threadIdx.x];
  __syncthreads();
  for (int o = threadIdx.x; o < SEER_D * SEER_E; o += blockDim.x) {
    float acc = bias[o];
#pragma unroll 6
    for (int k = 0; k < SEER_D; ++k) acc = fmaf(w[(size_t)o * SEER_D + k], es[k], acc);
    out[(size_t)b * SEER_D * SEER_E + o] = acc;
  }
}

// logits[b,i,j,:] = raw[b,i,j,:] + raw[b,j,i,:]; bond[b,i,j] = (i > j) ? argmax_k logits : 0
// raw: [B*42][ld] rows = (b,i), cols = j*5+k.   (reference adj_mat_seer.py:154-163, mol_utils.py:210-211)
__global__ void k_seer_symmetrise(const float* __restrict__ raw, int ld, float* __restrict__ logits, int8_t* __restrict__ bonds) {
  const int b = blockIdx.x;
  for (int idx = threadIdx.x; idx < SEER_D * SEER_D; idx += blockDim.x) {
    const int i = idx / SEER_D, j = idx - i * SEER_D;
    const float* pij = raw + ((size_t)b * SEER_D + i) * ld + j * SEER_NB;
    const float* pji = raw + ((size_t)b * SEER_D + j) * ld + i * SEER_NB;
    float best = 0.f;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < SEER_NB; ++k) {
      const float v = pji[k] + pij[k];  // torch.add(transpose, original)
      if (logits != nullptr) logits[(((size_t)b * SEER_D + i) * SEER_D + j) * SEER_NB + k] = v;
      if (k == 0 || v > best) { best = v; arg = k; }
    }
    if (bonds != nullptr) bonds[(size_t)b * SEER_D * SEER_D + idx] = (int8_t)((i > j) ? arg : 0);
  }
}

}  // namespace mlcg
