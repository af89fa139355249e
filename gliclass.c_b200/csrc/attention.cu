// K3 — fused DeBERTa disentangled attention for sm_100a (head dim 64).
//
//   ctx[b,i,h,:] = softmax_j( (Q_i.K_j + Q_i.posK[idx(i-j)] + K_j.posQ[idx(i-j)]) / sqrt(3d) + mask_j ) . V_j
//
// restating transformers' DisentangledSelfAttention (T:229-345: scores T:237-251, c2p T:313-324,
// p2c T:327-343, masked softmax T:256-259, context T:262-271) with the single index table
// idx(delta) = clamp(bucket(delta)+span, 0, 2*span-1).  ORT materialises the
// [B*h,S,S] score tensor and both [B*h,S,2*span] bias tensors per layer; here nothing but Q/K/V in
// and ctx out touches HBM.
//
// One CTA per (128-query tile, head, batch), looping over 64-key tiles t:
//   warp 0  TMA producer: Q once; per tile K_t,V_t (2-stage ring) and the <=192-row slices of
//           posK / posQ that the tile's relative positions can index (idx is monotone in i-j, so a
//           128x64 tile touches a contiguous slice of <=191 rows).
//   warp 1  tcgen05 issuer (M=128, fp32 accumulators in TMEM, 512 columns used):
//             C2P = Q . posK_slice^T   [128 x Npad]      cols 128..319
//             P2C'= posQ_slice . K_t^T [<=2 x 128 x 64]  cols 320..447   (rows = slice index)
//             S   = Q . K_t^T          [128 x 64]        cols 0..127 (double buffered)
//             PV  = P . V_t            [128 x 64]        cols 448..511
//   warps 2-9  softmax: 256 threads; thread = (row i, 32-key half g).  Stage C2P/P2C' TMEM->fp16
//           smem, gather both biases per score element, online softmax in registers (max shared
//           between the two halves through smem), P -> fp16 swizzled smem tile for the PV MMA,
//           O accumulated in registers with the usual rescale.
// Synchronisation is mbarrier based; every wait is bounded (ptx::mbar_wait traps on timeout).
#include <cuda_fp16.h>
#include <math_constants.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tma_desc.h"

namespace glc {
namespace {

constexpr int QT = 128;            // queries per CTA
constexpr int KT = 64;             // keys per tile
constexpr int D = 64;              // head dim
constexpr int WMAX = 192;          // max slice rows (191 needed)
constexpr int ATT_THREADS = 320;
constexpr int C2P_PITCH = 232;     // halves; 116 words = 20 mod 32 -> conflict-free 128-bit row stores, near conflict-free diagonal gathers
constexpr int P2C_PITCH = 66;      // halves; 33 words -> conflict-free
constexpr int LUT_MAX = 2048 + 128 + 64;

// shared memory map (bytes, from a 1024-aligned base)
constexpr int OFF_Q = 0;                          // 128 x 128 B
constexpr int OFF_K = OFF_Q + 16384;              // 2 x 8 KB
constexpr int OFF_V = OFF_K + 16384;              // 2 x 8 KB
constexpr int OFF_PK = OFF_V + 16384;             // 192 x 128 B
constexpr int OFF_PQ = OFF_PK + 24576;            // 256 x 128 B (two M=128 blocks)
constexpr int OFF_P = OFF_PQ + 32768;             // 128 x 128 B
constexpr int OFF_C2P = OFF_P + 16384;            // 128 x 464 B
constexpr int OFF_P2C = OFF_C2P + QT * C2P_PITCH * 2;      // 192 x 132 B
constexpr int OFF_XMAX = OFF_P2C + WMAX * P2C_PITCH * 2;   // 2 x 2 x 128 floats
constexpr int OFF_LUT = OFF_XMAX + 2048;          // uint16[LUT_MAX]
constexpr int OFF_MASK = OFF_LUT + ((LUT_MAX * 2 + 15) / 16) * 16;   // uint32[68]: key-validity words of this batch row
constexpr int OFF_BAR = OFF_MASK + 68 * 4;
constexpr int ATT_SMEM = OFF_BAR + 256 + 1024;

// TMEM columns
constexpr uint32_t TM_S = 0;       // 2 x 64
constexpr uint32_t TM_C2P = 128;   // 192
constexpr uint32_t TM_P2C = 320;   // 2 x 64
constexpr uint32_t TM_PV = 448;    // 64

struct AttnParams {
  const int32_t* rel_idx;    // [2*Spad-1], centre at Spad-1
  int rel_center;
  const uint32_t* mask_bits; // [B][ceil(S/32)]
  const int32_t* kv_len;     // [B]
  __half* ctx;        // [B*S, H]
  int B, S, heads, H;
  int buckets;               // position_buckets: idx(delta) = delta + buckets exactly for |delta| <= buckets/2
  float scale_log2;          // log2(e) / sqrt(3*d)
};

__device__ __forceinline__ void slice_bounds(const AttnParams& p, int q0, int k0, int& c0, int& w) {
  const int dmin = q0 - k0 - (KT - 1);
  const int dmax = q0 + (QT - 1) - k0;
  c0 = __ldg(p.rel_idx + p.rel_center + dmin);
  w = __ldg(p.rel_idx + p.rel_center + dmax) - c0 + 1;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fused_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_pk,
                       const __grid_constant__ CUtensorMap tm_pq, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (128B-swizzle atoms) by pointer arithmetic on the __shared__ symbol: an
  // integer round-trip would turn every later access into a generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;      // [2]
  uint64_t* kv_empty = bars + 3;     // [2]
  uint64_t* pos_full = bars + 5;
  uint64_t* pos_empty = bars + 6;
  uint64_t* mma1_full = bars + 7;    // bias + QK accumulators of tile t ready
  uint64_t* bias_free = bars + 8;    // softmax warps have drained the bias TMEM of tile t
  uint64_t* p_full = bars + 9;       // P tile written
  uint64_t* pv_full = bars + 10;     // PV accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int kvlen = p.kv_len[b];
  const int T = (kvlen + KT - 1) / KT;   // key tiles with at least one valid key

  if (q0 >= kvlen) {
    // padded queries only: their outputs are never read by valid rows
    for (int e = threadIdx.x; e < QT * 8; e += ATT_THREADS) {
      const int r = q0 + (e >> 3);
      if (r < p.S)
        *reinterpret_cast<uint4*>(p.ctx + ((int64_t)b * p.S + r) * p.H + head * D + (e & 7) * 8) = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tm_qkv);
    ptx::prefetch_tensormap(&tm_pk);
    ptx::prefetch_tensormap(&tm_pq);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
    ptx::mbar_init(pos_full, 1);
    ptx::mbar_init(pos_empty, 1);
    ptx::mbar_init(mma1_full, 1);
    ptx::mbar_init(bias_free, 8);
    ptx::mbar_init(p_full, 8);
    ptx::mbar_init(pv_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, QT * 128);
      ptx::tma_load_3d(smem + OFF_Q, &tm_qkv, q_full, head * D, q0, b);
      ptx::tma_load_3d(smem + OFF_Q + 8192, &tm_qkv, q_full, head * D, q0 + 64, b);
      for (int t = 0; t < T; ++t) {
        const int st = t & 1;
        const int k0 = t * KT;
        ptx::mbar_wait(&kv_empty[st], ((t >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&kv_full[st], 2 * KT * 128);
        ptx::tma_load_3d(smem + OFF_K + st * 8192, &tm_qkv, &kv_full[st], p.H + head * D, k0, b);
        ptx::tma_load_3d(smem + OFF_V + st * 8192, &tm_qkv, &kv_full[st], 2 * p.H + head * D, k0, b);
        int c0, w;
        slice_bounds(p, q0, k0, c0, w);
        const int nbox = (w + 63) >> 6;
        ptx::mbar_wait(pos_empty, (t & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(pos_full, (uint32_t)(2 * nbox * 64 * 128));
        for (int x = 0; x < nbox; ++x) {
          ptx::tma_load_3d(smem + OFF_PK + x * 8192, &tm_pk, pos_full, 0, c0 + x * 64, head);
          ptx::tma_load_3d(smem + OFF_PQ + x * 8192, &tm_pq, pos_full, 0, c0 + x * 64, head);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t sQ = ptx::smem_u32(smem + OFF_Q);
      const uint32_t sP = ptx::smem_u32(smem + OFF_P);
      const uint32_t sPK = ptx::smem_u32(smem + OFF_PK);
      const uint32_t sPQ = ptx::smem_u32(smem + OFF_PQ);
      constexpr uint32_t idesc_n64 = ptx::idesc_f16(128, 64);
      constexpr uint32_t idesc_pv = ptx::idesc_f16(128, 64, 0, 1);   // B (=V) is MN-major
      ptx::mbar_wait(q_full, 0);
      for (int t = 0; t <= T; ++t) {
        if (t < T) {
          const int st = t & 1;
          const int k0 = t * KT;
          int c0, w;
          slice_bounds(p, q0, k0, c0, w);
          const uint32_t npad = (uint32_t)((w + 15) & ~15);
          const int nblk = (w > 128) ? 2 : 1;
          const uint32_t sK = ptx::smem_u32(smem + OFF_K + st * 8192);
          ptx::mbar_wait(&kv_full[st], (t >> 1) & 1);
          ptx::mbar_wait(pos_full, t & 1);
          if (t > 0) ptx::mbar_wait(bias_free, (t - 1) & 1);
          ptx::tc_fence_after();
          const uint32_t idesc_c2p = ptx::idesc_f16(128, npad);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(tmem + TM_C2P, ptx::smem_desc_sw128(sQ + k * 32), ptx::smem_desc_sw128(sPK + k * 32), idesc_c2p,
                            (uint32_t)(k != 0));
          for (int blk = 0; blk < nblk; ++blk) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::mma_f16_ss(tmem + TM_P2C + blk * 64, ptx::smem_desc_sw128(sPQ + blk * 16384 + k * 32),
                              ptx::smem_desc_sw128(sK + k * 32), idesc_n64, (uint32_t)(k != 0));
          }
          ptx::mma_commit(pos_empty);   // pos slices consumed once the MMAs above retire
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(tmem + TM_S + (uint32_t)(st * 64), ptx::smem_desc_sw128(sQ + k * 32),
                            ptx::smem_desc_sw128(sK + k * 32), idesc_n64, (uint32_t)(k != 0));
          ptx::mma_commit(mma1_full);
        }
        if (t > 0) {
          const int tp = t - 1;
          const int st = tp & 1;
          const uint32_t sV = ptx::smem_u32(smem + OFF_V + st * 8192);
          ptx::mbar_wait(p_full, tp & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(tmem + TM_PV, ptx::smem_desc_sw128(sP + k * 32), ptx::smem_desc_sw128(sV + k * 2048), idesc_pv,
                            (uint32_t)(k != 0));
          ptx::mma_commit(&kv_empty[st]);
          ptx::mma_commit(pv_full);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int sw = warp - 2;          // 0..7
    const int g = sw >> 2;            // key half of the tile: keys [32g, 32g+32)
    const int qd = warp & 3;          // TMEM lane quarter
    const int i = qd * 32 + lane;     // row in the query tile
    const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
    __half* c2p_s = reinterpret_cast<__half*>(smem + OFF_C2P);
    __half* p2c_s = reinterpret_cast<__half*>(smem + OFF_P2C);
    float* xmax = reinterpret_cast<float*>(smem + OFF_XMAX);
    uint16_t* lut = reinterpret_cast<uint16_t*>(smem + OFF_LUT);
    const int words = (p.S + 31) >> 5;

    // relative-position LUT for every delta this CTA can see: delta in [q0 - (T*64-1), q0+127]
    const int d_lo = q0 - (T * KT - 1);
    const int n_lut = T * KT + QT - 1;
    for (int e = threadIdx.x - 64; e < n_lut; e += 256) lut[e] = (uint16_t)__ldg(p.rel_idx + p.rel_center + d_lo + e);
    uint32_t* kmask = reinterpret_cast<uint32_t*>(smem + OFF_MASK);
    if (threadIdx.x - 64 < 68) kmask[threadIdx.x - 64] = (threadIdx.x - 64 < words) ? __ldg(p.mask_bits + (int64_t)b * words + (threadIdx.x - 64)) : 0u;
    ptx::named_bar_sync(1, 256);

    float m_run = -CUDART_INF_F, l_run = 0.f, alpha_prev = 1.f;
    float o[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) o[k] = 0.f;
    const float sc = p.scale_log2;

    for (int t = 0; t < T; ++t) {
      const int k0 = t * KT;
      const int dmin = q0 - k0 - (KT - 1), dmax = q0 + (QT - 1) - k0;
      const int c0 = lut[dmin - d_lo];
      const int w = lut[dmax - d_lo] - c0 + 1;
      const int npad = (w + 15) & ~15;
      const int nblk = (w > 128) ? 2 : 1;
      const bool linear = (dmin >= -(p.buckets >> 1)) && (dmax <= (p.buckets >> 1));   // idx = delta + buckets

      ptx::mbar_wait(mma1_full, t & 1);
      ptx::tc_fence_after();

      // ---- stage C2P (row i, 32-column chunks ch = g, g+2, g+4) as fp16
      for (int ch = g; ch * 32 < npad; ch += 2) {
        uint32_t r[32];
        ptx::tmem_ld_x32(t_lane + TM_C2P + (uint32_t)(ch * 32), r);
        ptx::tmem_ld_wait();
        uint4* dst = reinterpret_cast<uint4*>(c2p_s + i * C2P_PITCH + ch * 32);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 o4;
          o4.x = ptx::pack_f16(__uint_as_float(r[8 * v + 0]), __uint_as_float(r[8 * v + 1]));
          o4.y = ptx::pack_f16(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3]));
          o4.z = ptx::pack_f16(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5]));
          o4.w = ptx::pack_f16(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7]));
          dst[v] = o4;
        }
      }
      // ---- stage P2C' (slice row c = blk*128 + i, keys [32g, 32g+32)) as fp16
      for (int blk = 0; blk < nblk; ++blk) {
        if (blk == 1 && qd >= 2) break;   // rows 192..255 are never indexed (warp-uniform)
        uint32_t r[32];
        ptx::tmem_ld_x32(t_lane + TM_P2C + (uint32_t)(blk * 64 + g * 32), r);
        ptx::tmem_ld_wait();
        uint32_t* dst = reinterpret_cast<uint32_t*>(p2c_s + (blk * 128 + i) * P2C_PITCH + g * 32);
#pragma unroll
        for (int v = 0; v < 16; ++v) dst[v] = ptx::pack_f16(__uint_as_float(r[2 * v]), __uint_as_float(r[2 * v + 1]));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bias_free);
      ptx::named_bar_sync(1, 256);   // staged biases visible to all softmax threads

      // ---- scores for (row i, keys k0+32g .. +32): S + c2p[i][idx] + p2c[idx][j]
      float s[32];
      {
        uint32_t r[32];
        ptx::tmem_ld_x32(t_lane + TM_S + (uint32_t)((t & 1) * 64 + g * 32), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) s[jj] = __uint_as_float(r[jj]);
      }
      const int kb = k0 + g * 32;
      if (linear) {
        // every delta of this tile is in the linear bucket region: slice row = delta - dmin, so both
        // gathers are affine in jj (immediate offsets, no LUT)
        const int rel0 = i + (KT - 1) - g * 32;                               // slice row for jj = 0
        const __half* c2p_i = c2p_s + i * C2P_PITCH + rel0;
        const __half* p2c_i = p2c_s + rel0 * P2C_PITCH + g * 32;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
          s[jj] += __half2float(__hadd(c2p_i[-jj], p2c_i[-jj * (P2C_PITCH - 1)]));
      } else {
        const uint16_t* lut_i = lut + (q0 + i - kb - d_lo);                   // indexed by -jj
        const __half* c2p_i = c2p_s + i * C2P_PITCH - c0;
        const __half* p2c_g = p2c_s + g * 32 - c0 * P2C_PITCH;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int c = lut_i[-jj];
          s[jj] += __half2float(__hadd(c2p_i[c], p2c_g[c * P2C_PITCH + jj]));
        }
      }
      const uint32_t kbits = kmask[kb >> 5];
      if (kbits != 0xffffffffu) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
          if (!((kbits >> jj) & 1u)) s[jj] = -CUDART_INF_F;
      }
      float mloc = s[0];
#pragma unroll
      for (int jj = 1; jj < 32; ++jj) mloc = fmaxf(mloc, s[jj]);
      // ---- row max shared between the two key halves
      xmax[((t & 1) * 2 + g) * 128 + i] = mloc;
      ptx::named_bar_sync(2, 256);
      const float mo = xmax[((t & 1) * 2 + (g ^ 1)) * 128 + i];
      const float m_new = fmaxf(m_run, fmaxf(mloc, mo));
      const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
      const float alpha = ptx::ex2((m_run - m_use) * sc);
      const float neg_ms = -m_use * sc;
      float psum = 0.f;
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        const float e = ptx::ex2(fmaf(s[jj], sc, neg_ms));
        s[jj] = e;
        psum += e;
      }
      l_run = l_run * alpha + psum;
      m_run = m_new;

      // ---- fold in PV of the previous tile (also guarantees the P buffer is free again)
      if (t > 0) {
        ptx::mbar_wait(pv_full, (t - 1) & 1);
        ptx::tc_fence_after();
        uint32_t r[32];
        ptx::tmem_ld_x32(t_lane + TM_PV + (uint32_t)(g * 32), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) o[k] = fmaf(o[k], alpha_prev, __uint_as_float(r[k]));
      }
      alpha_prev = alpha;

      // ---- P tile: row i, 16-byte chunks 4g..4g+3, 128-byte swizzle
      {
        uint8_t* prow = smem + OFF_P + (i >> 3) * 1024 + (i & 7) * 128;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 o4;
          o4.x = ptx::pack_f16(s[8 * v + 0], s[8 * v + 1]);
          o4.y = ptx::pack_f16(s[8 * v + 2], s[8 * v + 3]);
          o4.z = ptx::pack_f16(s[8 * v + 4], s[8 * v + 5]);
          o4.w = ptx::pack_f16(s[8 * v + 6], s[8 * v + 7]);
          *reinterpret_cast<uint4*>(prow + (((4 * g + v) ^ (i & 7)) << 4)) = o4;
        }
      }
      ptx::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }

    // ---- last PV, normalise, write ctx
    ptx::mbar_wait(pv_full, (T - 1) & 1);
    ptx::tc_fence_after();
    {
      uint32_t r[32];
      ptx::tmem_ld_x32(t_lane + TM_PV + (uint32_t)(g * 32), r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) o[k] = fmaf(o[k], alpha_prev, __uint_as_float(r[k]));
    }
    xmax[g * 128 + i] = l_run;   // reuse the exchange buffer for the row sums
    ptx::named_bar_sync(2, 256);
    const float l_tot = l_run + xmax[(g ^ 1) * 128 + i];
    const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
    const int row = q0 + i;
    if (row < p.S) {
      __half* dst = p.ctx + ((int64_t)b * p.S + row) * p.H + head * D + g * 32;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint4 o4;
        o4.x = ptx::pack_f16(o[8 * v + 0] * inv, o[8 * v + 1] * inv);
        o4.y = ptx::pack_f16(o[8 * v + 2] * inv, o[8 * v + 3] * inv);
        o4.z = ptx::pack_f16(o[8 * v + 4] * inv, o[8 * v + 5] * inv);
        o4.w = ptx::pack_f16(o[8 * v + 6] * inv, o[8 * v + 7] * inv);
        reinterpret_cast<uint4*>(dst)[v] = o4;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------
// slow restatement (CUDA cores, fp32 math on the same fp16 inputs): one warp per (b, h, i)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attention_naive_kernel(const __half* __restrict__ qkv, const __half* __restrict__ pos_k,
                       const __half* __restrict__ pos_q, const int32_t* __restrict__ rel_idx, int rel_center,
                       const uint32_t* __restrict__ mask_bits, __half* __restrict__ ctx, int B, int S, int heads,
                       int ld_pos, float inv_scale) {
  const int H = heads * D;
  const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= B * heads * S) return;
  const int i = gw % S;
  const int h = (gw / S) % heads;
  const int b = gw / (S * heads);
  const int words = (S + 31) >> 5;
  const __half* qrow = qkv + ((int64_t)b * S + i) * 3 * H + h * D;
  const float q0 = __half2float(qrow[lane]), q1 = __half2float(qrow[lane + 32]);
  float m = -CUDART_INF_F, l = 0.f, a0 = 0.f, a1 = 0.f;
  for (int j = 0; j < S; ++j) {
    if (!((mask_bits[(int64_t)b * words + (j >> 5)] >> (j & 31)) & 1u)) continue;
    const __half* krow = qkv + ((int64_t)b * S + j) * 3 * H + H + h * D;
    const __half* vrow = krow + H;
    const int idx = rel_idx[rel_center + i - j];
    const __half* pk = pos_k + (int64_t)idx * ld_pos + h * D;
    const __half* pq = pos_q + (int64_t)idx * ld_pos + h * D;
    const float k0 = __half2float(krow[lane]), k1 = __half2float(krow[lane + 32]);
    float s = q0 * k0 + q1 * k1 + q0 * __half2float(pk[lane]) + q1 * __half2float(pk[lane + 32]) +
              k0 * __half2float(pq[lane]) + k1 * __half2float(pq[lane + 32]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s *= inv_scale;
    const float mn = fmaxf(m, s);
    const float al = __expf(m - mn), pe = __expf(s - mn);
    l = l * al + pe;
    a0 = a0 * al + pe * __half2float(vrow[lane]);
    a1 = a1 * al + pe * __half2float(vrow[lane + 32]);
    m = mn;
  }
  const float inv = l > 0.f ? 1.f / l : 0.f;
  __half* dst = ctx + ((int64_t)b * S + i) * H + h * D;
  dst[lane] = __float2half_rn(a0 * inv);
  dst[lane + 32] = __float2half_rn(a1 * inv);
}

}  // namespace

cudaError_t attention_fused(const void* qkv, const void* pos_k, const void* pos_q, int64_t ld_pos, const int32_t* rel_idx,
                            const uint32_t* mask_bits, const int32_t* kv_len, void* ctx, int B, int S, int heads,
                            int buckets, int num_sms, cudaStream_t stream) {
  (void)num_sms;
  if (B <= 0 || S <= 0) return cudaSuccess;
  if (S > 2048) return cudaErrorInvalidValue;   // reference MAX_LENGTH (include/configs.h:5)
  const int H = heads * D;
  const int R = 2 * buckets;
  const int Spad = ((S + QT - 1) / QT) * QT;
  // qkv viewed as [B][S][3H]; box 64 cols x 64 rows
  uint64_t dq[3] = {(uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
  uint64_t sq[2] = {(uint64_t)(3 * H) * 2, (uint64_t)S * 3 * H * 2};
  uint32_t bq[3] = {64, 64, 1};
  // pos tables are [R][ld_pos] row-major (head h = columns h*64..): dims (d, r, head)
  uint64_t dp[3] = {64, (uint64_t)R, (uint64_t)heads};
  uint64_t sp[2] = {(uint64_t)ld_pos * 2, 128};
  uint32_t bp[3] = {64, 64, 1};
  CUtensorMap tm_qkv = make_tmap_16b(qkv, 3, dq, sq, bq);
  CUtensorMap tm_pk = make_tmap_16b(pos_k, 3, dp, sp, bp);
  CUtensorMap tm_pq = make_tmap_16b(pos_q, 3, dp, sp, bp);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attention_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  AttnParams p;
  p.rel_idx = rel_idx;
  p.rel_center = Spad - 1;
  p.mask_bits = mask_bits;
  p.kv_len = kv_len;
  p.ctx = (__half*)ctx;
  p.B = B; p.S = S; p.heads = heads; p.H = H;
  p.buckets = buckets;
  p.scale_log2 = 1.4426950408889634f / sqrtf(3.0f * D);
  dim3 grid((S + QT - 1) / QT, heads, B);
  attention_fused_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm_qkv, tm_pk, tm_pq, p);
  return cudaGetLastError();
}

cudaError_t attention_naive(const void* qkv, const void* pos_k, const void* pos_q, int64_t ld_pos, const int32_t* rel_idx,
                            const uint32_t* mask_bits, void* ctx, int B, int S, int heads, int buckets,
                            cudaStream_t stream) {
  if (B <= 0 || S <= 0) return cudaSuccess;
  const int Spad = ((S + QT - 1) / QT) * QT;
  const int rows = B * heads * S;
  attention_naive_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(
      (const __half*)qkv, (const __half*)pos_k, (const __half*)pos_q, rel_idx, Spad - 1, mask_bits,
      (__half*)ctx, B, S, heads, (int)ld_pos, 1.0f / sqrtf(3.0f * D));
  return cudaGetLastError();
}

}  // namespace glc
