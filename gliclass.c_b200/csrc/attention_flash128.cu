// Plain flash attention for the decoder backbones (Qwen2 / Llama style; reference Readme.md:91-94, BASELINE.json
// configs[4]): head dim 128, grouped-query heads, BIDIRECTIONAL with a key-padding mask (GLiClass uses decoder
// checkpoints as encoders), no bias terms:
//
//   ctx[b,i,h,:] = softmax_j( Q_i . K_j / sqrt(128) + mask_j ) . V_j        K, V of kv head h / (heads / kv_heads)
//
// (transformers modeling_qwen2.py: repeat_kv Q:150-158, eager_attention_forward Q:160-185, with the causal mask of Q:378-393
// replaced by the key-padding mask.)  Q and K arrive already rotated (rope_inplace).
//
// Same skeleton as the DeBERTa kernel (attention_persist.cu, in its one-CTA-per-item form) without the position-bias machinery: one CTA per (128-query
// tile, q head, batch row); 64-key tiles; S = Q.K_t^T double-buffered in TMEM (fp32), P as fp16 pairs in TMEM, O (128
// fp32 columns) resident in TMEM with a sticky row maximum; one softmax thread per query row of a tile; three softmax
// warpgroups rotate over the key tiles and chain the maximum (tile t -> group t % 3).
//   warp 0  TMA producer: Q (4 boxes), then K_t | V_t into a 3-stage ring (2 + 2 boxes of 64 x 64 halves per tile)
//   warp 1  tcgen05 issuer S: 8 k-steps (A = Q from TMEM, B = the two 64-column atoms of K_t)
//   warp 2  tcgen05 issuer PV: O[:, 0:64] (+)= P.V_t[:, 0:64], O[:, 64:128] (+)= P.V_t[:, 64:128] (V atoms are MN-major)
//   warps 4..15  softmax
// Synchronisation is mbarrier based; every wait is bounded (ptx::mbar_wait traps on timeout).
#include <cuda_fp16.h>
#include <math_constants.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"
#include "tma_desc.h"

namespace glc {
namespace {

constexpr int QT = 128;
constexpr int KT = 64;
constexpr int D = 128;
constexpr int NWG = 3;
constexpr int FTHREADS = 128 + 128 * NWG;
constexpr int UTIL_REGS = 40;
constexpr int SMX_REGS = 152;
constexpr int KVSTAGES = 5;   // a stage is refilled only after the PV MMA of its tile retired, and a TMA round trip is ~2 tiles long:
                              // with 3 stages the softmax groups slept on s_full for 31 % of the ncu samples

constexpr int OFF_Q = 0;                              // [2 atoms][128 rows][128 B]
constexpr int OFF_K = OFF_Q + 32768;                  // KVSTAGES x [2 atoms][64 rows][128 B]
constexpr int OFF_V = OFF_K + KVSTAGES * 16384;       // KVSTAGES x [2 d-halves][64 keys][128 B]
constexpr int OFF_MROW = OFF_V + KVSTAGES * 16384;    // float[128]
constexpr int OFF_LSUM = OFF_MROW + QT * 4;           // float[NWG][2][128]
constexpr int OFF_BAR = OFF_LSUM + NWG * 2 * QT * 4;
constexpr int NUM_BARS = 48;
constexpr int FLASH_SMEM = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
static_assert(OFF_BAR % 8 == 0, "barrier alignment");
static_assert(FLASH_SMEM <= 227 * 1024, "smem budget");

constexpr uint32_t TM_S = 0;      // 2 x 64
constexpr uint32_t TM_Q = 128;    // 64: Q tile as fp16 pairs (128 halves per row)
constexpr uint32_t TM_P = 192;    // 32
constexpr uint32_t TM_O = 256;    // 128

struct FlashParams {
  const uint32_t* mask_bits;
  const int32_t* kv_len;
  __half* ctx;               // [B*S, heads*128]
  int B, S, heads, kv_heads;
  float scale_log2;          // log2(e) / sqrt(128)
  // packed (varlen) layout, same conventions as attention_persist.cu: text b owns rows [text_row[b], text_row[b+1]) (a
  // multiple of 128), grid.x runs over the 128-row query tiles, tile_info[t] = (query tile within its text) << 24 | text
  const int32_t* tile_info;
  const int32_t* text_row;
};

__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float r = x - (t - 12582912.0f);
  float p = fmaf(0.05517090f, r, 0.24260953f);
  p = fmaf(p, r, 0.69326097f);
  p = fmaf(p, r, 0.99992818f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <uint32_t N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <uint32_t N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(FTHREADS, 1)
attention_flash128_kernel(const __grid_constant__ CUtensorMap tm_qkv, const FlashParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* qt_full = bars + 1;      // Q copied to TMEM (4 warps)
  uint64_t* kv_full = bars + 2;      // [KVSTAGES]
  uint64_t* kv_empty = kv_full + KVSTAGES;   // [KVSTAGES] the PV MMA of the tile retired
  uint64_t* s_free = kv_empty + KVSTAGES;    // [2] S buffer drained (4 warps)
  uint64_t* p_full = s_free + 2;     // P written (4 warps)
  uint64_t* l_bar = p_full + 1;      // partial sums published (4 * NWG warps)
  uint64_t* s_full = l_bar + 1;      // [NWG] S of tile t ready (index t % NWG)
  uint64_t* pv_full = s_full + NWG;  // [NWG]
  uint64_t* m_bar = pv_full + NWG;   // [NWG][4]
  static_assert(2 + 2 * KVSTAGES + 4 + 2 * NWG + 4 * NWG <= NUM_BARS, "barrier slots");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  float* mrow = reinterpret_cast<float*>(smem + OFF_MROW);
  float* lsum = reinterpret_cast<float*>(smem + OFF_LSUM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  int q0, b, grow, rows, trow, tb, mword;   // query offset in the text, text, first row of the text in ctx / qkv, rows it owns,
  if (p.tile_info != nullptr) {              // TMA (row, batch) coordinates of its position 0, its first mask word
    const int info = __ldg(p.tile_info + blockIdx.x);
    b = info & 0xffffff;
    q0 = (info >> 24) * QT;
    grow = __ldg(p.text_row + b);
    rows = __ldg(p.text_row + b + 1) - grow;
    trow = grow;
    tb = 0;
    mword = grow >> 5;
  } else {
    q0 = blockIdx.x * QT;
    b = blockIdx.z;
    grow = b * p.S;
    rows = p.S;
    trow = 0;
    tb = b;
    mword = b * ((p.S + 31) >> 5);
  }
  const int kvh = head / (p.heads / p.kv_heads);
  const int H = p.heads * D;
  const int kvlen = p.kv_len[b];
  const int T = (kvlen + KT - 1) / KT;

  if (q0 >= kvlen) {
    // padded queries only: never read by valid rows
    for (int e = threadIdx.x; e < QT * 16; e += FTHREADS) {
      const int r = q0 + (e >> 4);
      if (r < rows) *reinterpret_cast<uint4*>(p.ctx + ((int64_t)grow + r) * H + head * D + (e & 15) * 8) = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tm_qkv);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(qt_full, 4);
    for (int s = 0; s < KVSTAGES; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
    ptx::mbar_init(&s_free[0], 4);
    ptx::mbar_init(&s_free[1], 4);
    ptx::mbar_init(p_full, 4);
    ptx::mbar_init(l_bar, 4 * NWG);
    for (int k = 0; k < NWG; ++k) {
      ptx::mbar_init(&s_full[k], 1);
      ptx::mbar_init(&pv_full[k], 1);
      for (int q = 0; q < 4; ++q) ptx::mbar_init(&m_bar[k * 4 + q], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<UTIL_REGS>();
    if (warp == 0) {
      if (lane == 0) {
        const int qcol = head * D, kcol = (p.heads + kvh) * D, vcol = (p.heads + p.kv_heads + kvh) * D;
        ptx::mbar_arrive_expect_tx(q_full, QT * D * 2);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int r = 0; r < 2; ++r)
            ptx::tma_load_3d(smem + OFF_Q + a * 16384 + r * 8192, &tm_qkv, q_full, qcol + a * 64, trow + q0 + r * 64, tb);
        for (int t = 0; t < T; ++t) {
          const int st = t % KVSTAGES;
          ptx::mbar_wait(&kv_empty[st], (uint32_t)(((t / KVSTAGES) & 1) ^ 1));
          ptx::mbar_arrive_expect_tx(&kv_full[st], 2 * KT * D * 2);
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            ptx::tma_load_3d(smem + OFF_K + st * 16384 + a * 8192, &tm_qkv, &kv_full[st], kcol + a * 64, trow + t * KT, tb);
            ptx::tma_load_3d(smem + OFF_V + st * 16384 + a * 8192, &tm_qkv, &kv_full[st], vcol + a * 64, trow + t * KT, tb);
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ---------------------------------------------------------------- S = Q . K_t^T
      constexpr uint32_t idesc_s = ptx::idesc_f16(128, 64);
      ptx::mbar_wait(qt_full, 0);
      ptx::tc_fence_after();
      for (int t = 0; t < T; ++t) {
        const int st = t % KVSTAGES, sb = t & 1;
        const uint64_t dK0 = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_K + st * 16384));
        const uint64_t dK1 = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_K + st * 16384 + 8192));
        ptx::mbar_wait(&kv_full[st], (uint32_t)((t / KVSTAGES) & 1));
        if (t >= 2) ptx::mbar_wait(&s_free[sb], (uint32_t)(((t >> 1) - 1) & 1));   // the previous user of this S buffer drained it
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t dS = tmem + TM_S + (uint32_t)(64 * sb);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::mma_f16_ts(dS, tmem + TM_Q + 8 * k, dK0 + 2 * k, idesc_s, (uint32_t)(k != 0));
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::mma_f16_ts(dS, tmem + TM_Q + 32 + 8 * k, dK1 + 2 * k, idesc_s, 1u);
          ptx::mma_commit(&s_full[t % NWG]);
        }
        __syncwarp();
      }
    } else if (warp == 2) {
      // ---------------------------------------------------------------- O (+)= P . V_t, two 64-column halves of d
      constexpr uint32_t idesc_pv = ptx::idesc_f16(128, 64, 0, 1);   // B (= V) is MN-major
      for (int t = 0; t < T; ++t) {
        const int st = t % KVSTAGES;
        const uint64_t dV0 = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_V + st * 16384));
        const uint64_t dV1 = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_V + st * 16384 + 8192));
        ptx::mbar_wait(&kv_full[st], (uint32_t)((t / KVSTAGES) & 1));
        ptx::mbar_wait(p_full, (uint32_t)(t & 1));
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ts(tmem + TM_O, tmem + TM_P + 8 * k, dV0 + 128 * k, idesc_pv, (uint32_t)(k != 0 || t > 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ts(tmem + TM_O + 64, tmem + TM_P + 8 * k, dV1 + 128 * k, idesc_pv, (uint32_t)(k != 0 || t > 0));
          ptx::mma_commit(&kv_empty[st]);
          ptx::mma_commit(&pv_full[t % NWG]);
        }
        __syncwarp();
      }
    }
  } else {
    setmaxnreg_inc<SMX_REGS>();
    const int wg = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int i = qd * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
    const float sc = p.scale_log2;
    const int words = (rows + 31) >> 5;

    if (wg == 0) {
      // Q tile -> TMEM: this thread's row, 2 atoms x 8 16-byte chunks of the swizzled 128-byte rows
      ptx::mbar_wait(q_full, 0);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const uint8_t* qrow = smem + OFF_Q + a * 16384 + (i >> 3) * 1024 + (i & 7) * 128;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint4 lo = *reinterpret_cast<const uint4*>(qrow + (((2 * u) ^ (i & 7)) << 4));
          const uint4 hi = *reinterpret_cast<const uint4*>(qrow + (((2 * u + 1) ^ (i & 7)) << 4));
          const uint32_t qr[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
          ptx::tmem_st_x8(t_lane + TM_Q + (uint32_t)(32 * a + 8 * u), qr);
        }
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(qt_full);
    }

    float l_part = 0.f, m_seen = -CUDART_INF_F;
    for (int t = wg; t < T; t += NWG) {
      const int k0 = t * KT;
      const uint32_t par = (uint32_t)((t / NWG) & 1);
      const uint32_t kb0 = __ldg(p.mask_bits + mword + (k0 >> 5));
      const uint32_t kb1 = ((k0 >> 5) + 1 < words) ? __ldg(p.mask_bits + mword + (k0 >> 5) + 1) : 0u;
      float s[KT];
      ptx::mbar_wait(&s_full[t % NWG], par);
      ptx::tc_fence_after();
      {
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_x32(t_lane + TM_S + (uint32_t)(64 * (t & 1)), r0);
        ptx::tmem_ld_x32(t_lane + TM_S + (uint32_t)(64 * (t & 1) + 32), r1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s_free[t & 1]);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          s[jj] = __uint_as_float(r0[jj]);
          s[32 + jj] = __uint_as_float(r1[jj]);
        }
      }
      if ((kb0 & kb1) != 0xffffffffu) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          if (!((kb0 >> jj) & 1u)) s[jj] = -CUDART_INF_F;
          if (!((kb1 >> jj) & 1u)) s[32 + jj] = -CUDART_INF_F;
        }
      }
      float mx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v = s[16 * c];
#pragma unroll
        for (int k = 1; k + 1 < 16; k += 2) v = fmaxf(fmaxf(v, s[16 * c + k]), s[16 * c + k + 1]);
        mx[c] = fmaxf(v, s[16 * c + 15]);
      }
      const float mloc = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      float m_prev = -CUDART_INF_F;
      if (t > 0) {
        const int pg = (t - 1) % NWG;
        ptx::mbar_wait(&m_bar[pg * 4 + qd], (uint32_t)(((t - 1) / NWG) & 1));
        m_prev = mrow[i];
      }
      const bool raise = (mloc - m_prev) * sc > 8.0f;
      const float m_cur = raise ? mloc : m_prev;
      mrow[i] = m_cur;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&m_bar[wg * 4 + qd]);
      const float alpha = raise ? ptx::ex2((m_prev - m_cur) * sc) : 1.0f;
      const float neg_ms = (m_cur == -CUDART_INF_F) ? 0.f : -m_cur * sc;
      l_part *= (m_seen == m_cur) ? 1.0f : ((m_seen == -CUDART_INF_F) ? 0.f : ptx::ex2((m_seen - m_cur) * sc));
      m_seen = m_cur;
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
#pragma unroll
      for (int jj = 0; jj < KT; jj += 2) {
        const float x0 = fmaf(s[jj], sc, neg_ms), x1 = fmaf(s[jj + 1], sc, neg_ms);
        const float e0 = ptx::ex2(x0);
        const float e1 = ((jj & 3) == 2) ? exp2_poly(x1) : ptx::ex2(x1);
        ps[jj & 3] += e0;
        ps[(jj + 1) & 3] += e1;
        pk[jj >> 1] = ptx::pack_f16(e0, e1);
      }
      l_part += (ps[0] + ps[1]) + (ps[2] + ps[3]);
      if (t > 0) {
        ptx::mbar_wait(&pv_full[(t - 1) % NWG], (uint32_t)(((t - 1) / NWG) & 1));
        if (__any_sync(0xffffffffu, raise)) {
          ptx::tc_fence_after();
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t r[32];
            ptx::tmem_ld_x32(t_lane + TM_O + (uint32_t)(32 * q4), r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) r[jj] = __float_as_uint(__uint_as_float(r[jj]) * alpha);
            ptx::tmem_st_x32(t_lane + TM_O + (uint32_t)(32 * q4), r);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t pr[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) pr[v] = pk[8 * u + v];
        ptx::tmem_st_x8(t_lane + TM_P + (uint32_t)(8 * u), pr);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }

    const int last_wg = (T - 1) % NWG;
    lsum[(wg * 2 + 0) * QT + i] = l_part;
    lsum[(wg * 2 + 1) * QT + i] = m_seen;
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(l_bar);
    if (wg == last_wg) {
      ptx::mbar_wait(l_bar, 0);
      float l_tot = 0.f;
#pragma unroll
      for (int k = 0; k < NWG; ++k) {
        const float lk = lsum[(k * 2 + 0) * QT + i], mk = lsum[(k * 2 + 1) * QT + i];
        l_tot += (mk == m_seen) ? lk : ((mk == -CUDART_INF_F) ? 0.f : lk * ptx::ex2((mk - m_seen) * sc));
      }
      const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
      ptx::mbar_wait(&pv_full[(T - 1) % NWG], (uint32_t)(((T - 1) / NWG) & 1));
      ptx::tc_fence_after();
      const int row = q0 + i;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint32_t r[32];
        ptx::tmem_ld_x32(t_lane + TM_O + (uint32_t)(32 * q4), r);
        ptx::tmem_ld_wait();
        if (row < rows) {
          __half* dst = p.ctx + ((int64_t)grow + row) * H + head * D + 32 * q4;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 o4;
            o4.x = ptx::pack_f16(__uint_as_float(r[8 * v + 0]) * inv, __uint_as_float(r[8 * v + 1]) * inv);
            o4.y = ptx::pack_f16(__uint_as_float(r[8 * v + 2]) * inv, __uint_as_float(r[8 * v + 3]) * inv);
            o4.z = ptx::pack_f16(__uint_as_float(r[8 * v + 4]) * inv, __uint_as_float(r[8 * v + 5]) * inv);
            o4.w = ptx::pack_f16(__uint_as_float(r[8 * v + 6]) * inv, __uint_as_float(r[8 * v + 7]) * inv);
            reinterpret_cast<uint4*>(dst)[v] = o4;
          }
        }
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

}  // namespace

static cudaError_t launch_flash128(const void* qkv, FlashParams p, uint64_t tm_rows, uint64_t tm_batch, dim3 grid, cudaStream_t stream) {
  const int W = (p.heads + 2 * p.kv_heads) * D;
  uint64_t dq[3] = {(uint64_t)W, tm_rows, tm_batch};
  uint64_t sq[2] = {(uint64_t)W * 2, tm_rows * W * 2};
  uint32_t bq[3] = {64, 64, 1};
  CUtensorMap tm_qkv = make_tmap_16b(qkv, 3, dq, sq, bq);
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)D);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attention_flash128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FLASH_SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  attention_flash128_kernel<<<grid, FTHREADS, FLASH_SMEM, stream>>>(tm_qkv, p);
  return cudaGetLastError();
}

cudaError_t attention_flash128(const void* qkv, const uint32_t* mask_bits, const int32_t* kv_len, void* ctx, int B, int S,
                               int heads, int kv_heads, cudaStream_t stream) {
  if (B <= 0 || S <= 0) return cudaSuccess;
  if (heads <= 0 || kv_heads <= 0 || heads % kv_heads) return cudaErrorInvalidValue;
  FlashParams p{};
  p.mask_bits = mask_bits;
  p.kv_len = kv_len;
  p.ctx = (__half*)ctx;
  p.B = B; p.S = S; p.heads = heads; p.kv_heads = kv_heads;
  return launch_flash128(qkv, p, (uint64_t)S, (uint64_t)B, dim3((S + QT - 1) / QT, heads, B), stream);
}

cudaError_t attention_flash128_packed(const void* qkv, const uint32_t* row_bits, const int32_t* kv_len, const int32_t* text_row,
                                      const int32_t* tile_info, void* ctx, int B, int rows, int n_tiles, int heads, int kv_heads,
                                      cudaStream_t stream) {
  if (B <= 0 || rows <= 0 || n_tiles <= 0) return cudaSuccess;
  if (heads <= 0 || kv_heads <= 0 || heads % kv_heads || (rows % QT) || n_tiles != rows / QT || B >= (1 << 24)) return cudaErrorInvalidValue;
  FlashParams p{};
  p.mask_bits = row_bits;
  p.kv_len = kv_len;
  p.ctx = (__half*)ctx;
  p.B = B; p.S = rows; p.heads = heads; p.kv_heads = kv_heads;
  p.tile_info = tile_info;
  p.text_row = text_row;
  return launch_flash128(qkv, p, (uint64_t)rows, 1, dim3(n_tiles, heads, 1), stream);
}

}  // namespace glc
