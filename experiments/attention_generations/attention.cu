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

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"
#include "tma_desc.h"

namespace glc {
namespace {

constexpr int QT = 128;            // queries per CTA
constexpr int KT = 64;             // keys per tile
constexpr int D = 64;              // head dim
constexpr int G = 4;               // key groups per tile: softmax thread = (query row, 16-key group)
constexpr int E = KT / G;          // scores per softmax thread per tile
constexpr int SM_WARPS = 4 * G;    // 4 TMEM lane quarters x G groups
constexpr int SM_THREADS = 32 * SM_WARPS;
constexpr int WMAX = 192;          // max slice rows (191 needed)
constexpr int ATT_THREADS = 64 + SM_THREADS;
constexpr int C2P_PITCH = 120;     // halves; a 32-row warp quarter touches <= 7 16-column chunks of the slice (band-packed);
                                   // 60 words = 28 mod 32 -> conflict-free 128-bit row stores
constexpr int P2C_PITCH = 66;      // halves; 33 words -> conflict-free
constexpr int LUT_MAX = 2048 + 128 + 64;
constexpr int TMAX = 2048 / KT;    // key tiles at the reference MAX_LENGTH

// shared memory map (bytes, from a 1024-aligned base)
constexpr int OFF_Q = 0;                          // 128 x 128 B
constexpr int OFF_K = OFF_Q + 16384;              // 2 x 8 KB
constexpr int OFF_V = OFF_K + 16384;              // 2 x 8 KB
constexpr int POS_BYTES = WMAX * 128;             // one 192-row slice
constexpr int OFF_PK = OFF_V + 16384;             // 2 x 192 x 128 B
constexpr int OFF_PQ = OFF_PK + 2 * POS_BYTES;    // 2 x 192 x 128 B; the second M=128 block of a slice reads 64 rows past
                                                  // its end (next buffer / P tile): never-indexed accumulator rows
constexpr int OFF_P = OFF_PQ + 2 * POS_BYTES;     // 128 x 128 B
constexpr int OFF_C2P = OFF_P + 16384;            // 128 x 240 B
constexpr int OFF_P2C = OFF_C2P + QT * C2P_PITCH * 2;      // 192 x 132 B
constexpr int OFF_XMAX = OFF_P2C + WMAX * P2C_PITCH * 2;   // G x 128 floats
constexpr int OFF_LUT = OFF_XMAX + G * QT * 4;    // uint16[LUT_MAX]
constexpr int OFF_MASK = OFF_LUT + ((LUT_MAX * 2 + 15) / 16) * 16;   // uint32[68]: key-validity words of this batch row
constexpr int OFF_TILE = OFF_MASK + 68 * 4;       // uint32[TMAX]: per key tile (slice start c0 | slice rows w << 16)
constexpr int OFF_BAR = OFF_TILE + TMAX * 4;
constexpr int ATT_SMEM = OFF_BAR + 256 + 1024;
static_assert(ATT_SMEM <= 227 * 1024, "attention smem budget");

// TMEM columns
constexpr uint32_t TM_S = 0;       // 64
constexpr uint32_t TM_Q = 64;      // 32: the Q tile as fp16 pairs (A operand of the S and C2P MMAs)
constexpr uint32_t TM_P = 96;      // 32: the P tile as fp16 pairs (A operand of the PV MMA)
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
  int flags;                 // TRACE instantiation only: developer what-if switches (results become wrong)
  long long* trace;          // TRACE instantiation only: [2 roles][TMAX][8] clock64 stamps of CTA (1,0,0)
};

#define GLC_TRACE(role, tile, slot)                                                               \
  do {                                                                                            \
    if (TRACE && p.trace && (threadIdx.x & 31) == 0 && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0)                         \
      p.trace[((role) * TMAX + (tile)) * 8 + (slot)] = clock64();                                 \
  } while (0)

// 16 fp32 accumulator columns -> 16 halves -> two 128-bit shared stores
__device__ __forceinline__ void pack16_store(const uint32_t (&r)[16], uint4* dst) {
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    uint4 o4;
    o4.x = ptx::pack_f16(__uint_as_float(r[8 * v + 0]), __uint_as_float(r[8 * v + 1]));
    o4.y = ptx::pack_f16(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3]));
    o4.z = ptx::pack_f16(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5]));
    o4.w = ptx::pack_f16(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7]));
    dst[v] = o4;
  }
}

template <bool TRACE>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fused_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_pk,
                       const __grid_constant__ CUtensorMap tm_pq, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (128B-swizzle atoms) by pointer arithmetic on the __shared__ symbol: an
  // integer round-trip would turn every later access into a generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  // two operand rings, each two deep, freed at different moments of a tile:
  //   A = K_t + posK/posQ slices (read by the S / C2P / P2C' MMAs)   B = V_t (read by the PV MMA)
  uint64_t* a_full = bars + 1;       // [2]
  uint64_t* a_empty = bars + 3;      // [2]
  uint64_t* b_full = bars + 5;       // [2]
  uint64_t* b_empty = bars + 7;      // [2]
  uint64_t* mma1_full = bars + 9;    // bias + QK accumulators of tile t ready
  uint64_t* bias_free = bars + 10;   // softmax warps have drained the bias TMEM of tile t
  uint64_t* p_full = bars + 11;      // P tile written
  uint64_t* pv_full = bars + 12;     // PV accumulator ready
  uint64_t* qt_full = bars + 13;     // Q tile copied into TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  uint16_t* lut = reinterpret_cast<uint16_t*>(smem + OFF_LUT);
  uint32_t* kmask = reinterpret_cast<uint32_t*>(smem + OFF_MASK);
  uint32_t* tile_cw = reinterpret_cast<uint32_t*>(smem + OFF_TILE);

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
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
    ptx::mbar_init(mma1_full, 1);
    ptx::mbar_init(bias_free, SM_WARPS);
    ptx::mbar_init(p_full, SM_WARPS);
    ptx::mbar_init(pv_full, 1);
    ptx::mbar_init(qt_full, SM_WARPS);
    ptx::fence_barrier_init();
    // Q and the first V tile do not depend on the tables built below: start them now
    ptx::mbar_arrive_expect_tx(q_full, QT * 128);
    ptx::tma_load_3d(smem + OFF_Q, &tm_qkv, q_full, head * D, q0, b);
    ptx::tma_load_3d(smem + OFF_Q + 8192, &tm_qkv, q_full, head * D, q0 + 64, b);
    ptx::mbar_arrive_expect_tx(&b_full[0], KT * 128);
    ptx::tma_load_3d(smem + OFF_V, &tm_qkv, &b_full[0], 2 * p.H + head * D, 0, b);
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);

  // relative-position LUT for every delta this CTA can see: delta in [q0 - (T*64-1), q0+127];
  // per key tile the slice [c0, c0+w) of the position tables its deltas index; key-validity words
  const int d_lo = q0 - (T * KT - 1);
  {
    const int n_lut = T * KT + QT - 1;
    const int32_t* rel = p.rel_idx + p.rel_center + d_lo;
    for (int e = threadIdx.x; e < n_lut; e += ATT_THREADS) lut[e] = (uint16_t)__ldg(rel + e);
    if (threadIdx.x < T) {
      const int k0 = threadIdx.x * KT;
      const int c0 = __ldg(p.rel_idx + p.rel_center + q0 - k0 - (KT - 1));
      const int c1 = __ldg(p.rel_idx + p.rel_center + q0 + (QT - 1) - k0);
      tile_cw[threadIdx.x] = (uint32_t)c0 | ((uint32_t)(c1 - c0 + 1) << 16);
    }
    const int words = (p.S + 31) >> 5;
    const int e = threadIdx.x - 64;
    if (e >= 0 && e < 68) kmask[e] = (e < words) ? __ldg(p.mask_bits + (int64_t)b * words + e) : 0u;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // ring A: K_x and the position slices of tile x; free once the bias MMAs of tile x-2 retired
      auto load_a = [&](int x) {
        const uint32_t cw = tile_cw[x];
        const int c0 = (int)(cw & 0xffffu), w = (int)(cw >> 16);
        const int nbox = (TRACE && (p.flags & 16)) ? 1 : ((w + 63) >> 6);
        const int st = x & 1;
        ptx::mbar_wait(&a_empty[st], ((x >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&a_full[st], (uint32_t)((1 + 2 * nbox) * 64 * 128));
        ptx::tma_load_3d(smem + OFF_K + st * 8192, &tm_qkv, &a_full[st], p.H + head * D, x * KT, b);
        for (int bx = 0; bx < nbox; ++bx) {
          ptx::tma_load_3d(smem + OFF_PK + st * POS_BYTES + bx * 8192, &tm_pk, &a_full[st], 0, c0 + bx * 64, head);
          ptx::tma_load_3d(smem + OFF_PQ + st * POS_BYTES + bx * 8192, &tm_pq, &a_full[st], 0, c0 + bx * 64, head);
        }
      };
      // ring B: V_x; free once the PV MMA of tile x-2 retired
      auto load_b = [&](int x) {
        const int st = x & 1;
        ptx::mbar_wait(&b_empty[st], ((x >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&b_full[st], KT * 128);
        ptx::tma_load_3d(smem + OFF_V + st * 8192, &tm_qkv, &b_full[st], 2 * p.H + head * D, x * KT, b);
      };
      // issue order = order in which the buffers come free: bias(t) retires before PV(t-1)
      load_a(0);
      if (T > 1) load_a(1);
      for (int t = 0; t < T; ++t) {
        if (t + 2 < T) load_a(t + 2);
        if (t + 1 < T) load_b(t + 1);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // All 32 lanes walk the (warp-uniform) loop and wait on the barriers; one elected lane issues the
    // tcgen05 instructions.  Issuing from inside `if (lane == 0)` makes the compiler wrap every UTCHMMA
    // in an R2UR + ELECT retry loop (~100 cycles per instruction).
    constexpr uint32_t idesc_n64 = ptx::idesc_f16(128, 64);
    constexpr uint32_t idesc_pv = ptx::idesc_f16(128, 64, 0, 1);   // B (=V) is MN-major
    ptx::mbar_wait(qt_full, 0);
    ptx::tc_fence_after();
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        const int st = t & 1;
        const uint32_t cw = tile_cw[t];
        const int w = (int)(cw >> 16);
        const uint32_t npad = (uint32_t)((w + 15) & ~15);
        const uint64_t dK = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_K + st * 8192));
        const uint64_t dPK = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_PK + st * POS_BYTES));
        const uint64_t dPQ = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_PQ + st * POS_BYTES));
        GLC_TRACE(1, t, 0);
        ptx::mbar_wait(&a_full[st], (t >> 1) & 1);
        ptx::tc_fence_after();
        GLC_TRACE(1, t, 1);
        GLC_TRACE(1, t, 2);
        if (t > 0) ptx::mbar_wait(bias_free, (t - 1) & 1);   // S, C2P and P2C' accumulators drained
        ptx::tc_fence_after();
        GLC_TRACE(1, t, 3);
        if (ptx::elect_one()) {
          const uint32_t idesc_c2p = ptx::idesc_f16(128, npad);
          // A = Q from TMEM: 16 halves along K = 8 columns per step
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ts(tmem + TM_S, tmem + TM_Q + 8 * k, dK + 2 * k, idesc_n64, (uint32_t)(k != 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ts(tmem + TM_C2P, tmem + TM_Q + 8 * k, dPK + 2 * k, idesc_c2p, (uint32_t)(k != 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(tmem + TM_P2C, dPQ + 2 * k, dK + 2 * k, idesc_n64, (uint32_t)(k != 0));
          if (w > 128) {
#pragma unroll
            for (int k = 0; k < 4; ++k)   // second M=128 block of the slice: +16384 bytes = +1024 in the descriptor
              ptx::mma_f16_ss(tmem + TM_P2C + 64, dPQ + 1024 + 2 * k, dK + 2 * k, idesc_n64, (uint32_t)(k != 0));
          }
          ptx::mma_commit(&a_empty[st]);   // K_t and the slices are consumed once the MMAs above retire
          ptx::mma_commit(mma1_full);
        }
        __syncwarp();
        GLC_TRACE(1, t, 4);
      }
      if (t > 0) {
        const int tp = t - 1;
        const int st = tp & 1;
        const uint64_t dV = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_V + st * 8192));
        ptx::mbar_wait(&b_full[st], (tp >> 1) & 1);
        ptx::mbar_wait(p_full, tp & 1);
        ptx::tc_fence_after();
        GLC_TRACE(1, tp, 5);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)   // V is MN-major: 16 keys further = +2048 bytes = +128 in the descriptor
            ptx::mma_f16_ts(tmem + TM_PV, tmem + TM_P + 8 * k, dV + 128 * k, idesc_pv, (uint32_t)(k != 0));
          ptx::mma_commit(&b_empty[st]);
          ptx::mma_commit(pv_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int sw = warp - 2;          // 0..SM_WARPS-1
    const int g = sw >> 2;            // key group of the tile: keys [E*g, E*g+E)
    const int qd = warp & 3;          // TMEM lane quarter
    const int i = qd * 32 + lane;     // row in the query tile
    const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
    __half* c2p_s = reinterpret_cast<__half*>(smem + OFF_C2P);
    __half* p2c_s = reinterpret_cast<__half*>(smem + OFF_P2C);
    float* xmax = reinterpret_cast<float*>(smem + OFF_XMAX);
    const uint32_t s_c2p = ptx::smem_u32(c2p_s), s_p2c = ptx::smem_u32(p2c_s), s_lut = ptx::smem_u32(lut);

    // ---- Q tile -> TMEM once (row i, halves [16g, 16g+16) = 16-byte chunks 2g, 2g+1 of the swizzled row)
    ptx::mbar_wait(q_full, 0);
    {
      const uint8_t* qrow = smem + OFF_Q + (i >> 3) * 1024 + (i & 7) * 128;
      const uint4 lo = *reinterpret_cast<const uint4*>(qrow + (((2 * g) ^ (i & 7)) << 4));
      const uint4 hi = *reinterpret_cast<const uint4*>(qrow + (((2 * g + 1) ^ (i & 7)) << 4));
      const uint32_t qr[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      ptx::tmem_st_x8(t_lane + TM_Q + (uint32_t)(8 * g), qr);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(qt_full);
    }

    float m_run = -CUDART_INF_F, l_run = 0.f, alpha_prev = 1.f;
    float o[E];
#pragma unroll
    for (int k = 0; k < E; ++k) o[k] = 0.f;
    const float sc = p.scale_log2;
    const int half_b = p.buckets >> 1;

    for (int t = 0; t < T; ++t) {
      const int k0 = t * KT;
      const int dmin = q0 - k0 - (KT - 1), dmax = q0 + (QT - 1) - k0;
      const uint32_t cw = tile_cw[t];
      const int c0 = (int)(cw & 0xffffu), w = (int)(cw >> 16);
      const bool linear = (dmin >= -half_b) && (dmax <= half_b);   // idx = delta + buckets

      if (sw == 0 && lane == 0) GLC_TRACE(0, t, 0);
      ptx::mbar_wait(mma1_full, t & 1);
      ptx::tc_fence_after();
      if (sw == 0 && lane == 0) GLC_TRACE(0, t, 1);
      // ---- S for (row i, keys k0+E*g .. +E) first: the single S accumulator is released together with
      //      the bias accumulators
      float s[E];
      {
        uint32_t r[16];
        ptx::tmem_ld_x16(t_lane + TM_S + (uint32_t)(g * E), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < E; ++jj) s[jj] = __uint_as_float(r[jj]);
      }

      // ---- stage C2P as fp16, band-packed: the 32 rows of this warp quarter only index slice columns
      //      [lo, hi] (<= 95 wide), i.e. 16-column chunks u_lo..u_hi (<= 7); group g takes u_lo+g, u_lo+g+G
      const int dq = q0 + qd * 32 - k0 - d_lo;   // LUT position of (first row of the quarter, key k0)
      const int u_lo = ((int)lut[dq - (KT - 1)] - c0) >> 4;
      const int u_hi = ((int)lut[dq + 31] - c0) >> 4;
      {
        uint4* dst = reinterpret_cast<uint4*>(c2p_s + i * C2P_PITCH + g * 16);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int ch = u_lo + g + u * G;
          if (ch <= u_hi && !(TRACE && (p.flags & 2))) {   // warp-uniform
            uint32_t ra[16];
            ptx::tmem_ld_x16(t_lane + TM_C2P + (uint32_t)(ch * 16), ra);
            ptx::tmem_ld_wait();
            pack16_store(ra, dst + 2 * G * u);
          }
        }
      }
      // ---- stage P2C' (slice row c = blk*128 + i, keys [E*g, E*g+E)) as fp16
      {
        uint32_t ra[16], rb[16];
        const bool two = (w > 128) && (qd < 2);   // rows 192..255 are never indexed (warp-uniform)
        if (!(TRACE && (p.flags & 4)))
        ptx::tmem_ld_x16(t_lane + TM_P2C + (uint32_t)(g * E), ra);
        if (two) ptx::tmem_ld_x16(t_lane + TM_P2C + (uint32_t)(64 + g * E), rb);
        ptx::tmem_ld_wait();
        uint32_t* dst = reinterpret_cast<uint32_t*>(p2c_s + i * P2C_PITCH + g * E);
        if (!(TRACE && (p.flags & 4)))
#pragma unroll
        for (int v = 0; v < 8; ++v) dst[v] = ptx::pack_f16(__uint_as_float(ra[2 * v]), __uint_as_float(ra[2 * v + 1]));
        if (two) {
          uint32_t* dst2 = dst + 128 * P2C_PITCH / 2;
#pragma unroll
          for (int v = 0; v < 8; ++v) dst2[v] = ptx::pack_f16(__uint_as_float(rb[2 * v]), __uint_as_float(rb[2 * v + 1]));
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bias_free);
      if (sw == 0 && lane == 0) GLC_TRACE(0, t, 2);
      ptx::named_bar_sync(1, SM_THREADS);   // staged biases visible to all softmax threads
      if (sw == 0 && lane == 0) GLC_TRACE(0, t, 3);

      // ---- scores: S + c2p[i][idx] + p2c[idx][j]
      const int kb = k0 + g * E;
      if (TRACE && (p.flags & 1)) {
      } else if (linear) {
        // every delta of this tile is in the linear bucket region: slice row = delta - dmin, so both
        // gathers are affine in jj (immediate offsets, no LUT)
        const int rel0 = i + (KT - 1) - g * E;                               // slice row for jj = 0
        const uint32_t a1 = s_c2p + 2u * (uint32_t)(i * C2P_PITCH + rel0 - u_lo * 16);
        const uint32_t a2 = s_p2c + 2u * (uint32_t)(rel0 * P2C_PITCH + g * E);
#pragma unroll
        for (int jj = 0; jj < E; ++jj)
          s[jj] += __half2float(__hadd(ptx::lds_f16(a1 - 2u * jj), ptx::lds_f16(a2 - 2u * jj * (P2C_PITCH - 1))));
      } else {
        // byte addresses by hand (one multiply-add per gather): the compiler otherwise spends six
        // integer instructions per element on index -> address conversion
        const uint32_t al = s_lut + 2u * (uint32_t)(q0 + i - kb - d_lo);                 // LUT entry of jj = 0
        uint32_t a1 = s_c2p + 2u * (uint32_t)(i * C2P_PITCH - c0 - u_lo * 16);           // + 2 c
        uint32_t a2 = s_p2c + 2u * (uint32_t)(g * E - c0 * P2C_PITCH);                   // + 132 c + 2 jj
        asm volatile("" : "+r"(a1), "+r"(a2));   // opaque: keep the folded bases, one IMAD per gather
#pragma unroll
        for (int jj = 0; jj < E; ++jj) {
          const uint32_t c = ptx::lds_u16(al - 2u * jj);
          s[jj] += __half2float(__hadd(ptx::lds_f16(a1 + 2u * c), ptx::lds_f16(a2 + (2u * P2C_PITCH) * c + 2u * jj)));
        }
      }
      const uint32_t kbits = (kmask[kb >> 5] >> (kb & 31)) & ((1u << E) - 1u);
      if (kbits != ((1u << E) - 1u)) {
#pragma unroll
        for (int jj = 0; jj < E; ++jj)
          if (!((kbits >> jj) & 1u)) s[jj] = -CUDART_INF_F;
      }
      float mloc = s[0];
#pragma unroll
      for (int jj = 1; jj < E; ++jj) mloc = fmaxf(mloc, s[jj]);
      // ---- row max shared between the key groups (the barrier of the next tile's staging orders
      //      these reads before the next writes)
      xmax[g * QT + i] = mloc;
      if (sw == 0 && lane == 0) GLC_TRACE(0, t, 4);
      ptx::named_bar_sync(2 + qd, 32 * G);   // only the G warps of this lane quarter share rows
      if (sw == 0 && lane == 0) GLC_TRACE(0, t, 5);
      float m_new = m_run;
#pragma unroll
      for (int gg = 0; gg < G; ++gg) m_new = fmaxf(m_new, xmax[gg * QT + i]);
      const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
      const float alpha = ptx::ex2((m_run - m_use) * sc);
      const float neg_ms = -m_use * sc;
      float psum = 0.f;
#pragma unroll
      for (int jj = 0; jj < E; ++jj) {
        const float e = (TRACE && (p.flags & 8)) ? fmaf(s[jj], sc, neg_ms) : ptx::ex2(fmaf(s[jj], sc, neg_ms));
        s[jj] = e;
        psum += e;
      }
      l_run = l_run * alpha + psum;
      m_run = m_new;

      // ---- fold in PV of the previous tile (also guarantees the P buffer is free again)
      if (sw == 0 && lane == 0) GLC_TRACE(0, t, 6);
      if (t > 0) {
        ptx::mbar_wait(pv_full, (t - 1) & 1);
        ptx::tc_fence_after();
        uint32_t r[16];
        ptx::tmem_ld_x16(t_lane + TM_PV + (uint32_t)(g * E), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < E; ++k) o[k] = fmaf(o[k], alpha_prev, __uint_as_float(r[k]));
      }
      alpha_prev = alpha;

      // ---- P tile -> TMEM: row i, fp16 pairs at columns 8g .. 8g+7 (PV of tile t-1 has retired: pv_full)
      {
        uint32_t pr[E / 2];
#pragma unroll
        for (int v = 0; v < E / 2; ++v) pr[v] = ptx::pack_f16(s[2 * v], s[2 * v + 1]);
        ptx::tmem_st_x8(t_lane + TM_P + (uint32_t)(8 * g), pr);
        ptx::tmem_st_wait();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
      if (sw == 0 && lane == 0) GLC_TRACE(0, t, 7);
    }

    // ---- last PV, normalise, write ctx
    ptx::mbar_wait(pv_full, (T - 1) & 1);
    ptx::tc_fence_after();
    {
      uint32_t r[16];
      ptx::tmem_ld_x16(t_lane + TM_PV + (uint32_t)(g * E), r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < E; ++k) o[k] = fmaf(o[k], alpha_prev, __uint_as_float(r[k]));
    }
    ptx::named_bar_sync(2 + qd, 32 * G);  // the quarter is past its last row-max read
    xmax[g * QT + i] = l_run;             // reuse the exchange buffer for the row sums
    ptx::named_bar_sync(2 + qd, 32 * G);
    float l_tot = 0.f;
#pragma unroll
    for (int gg = 0; gg < G; ++gg) l_tot += xmax[gg * QT + i];
    const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
    const int row = q0 + i;
    if (row < p.S) {
      __half* dst = p.ctx + ((int64_t)b * p.S + row) * p.H + head * D + g * E;
#pragma unroll
      for (int v = 0; v < E / 8; ++v) {
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
    cudaError_t e = cudaFuncSetAttribute(attention_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
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
  p.trace = nullptr;
  p.flags = 0;
  dim3 grid((S + QT - 1) / QT, heads, B);
  // developer aid: GLC_ATTN_TRACE=<file> dumps per-tile clock64 stamps of CTA (1,0,0) (synchronous)
  if (const char* fl = getenv("GLC_ATTN_FLAGS")) p.flags = atoi(fl);
  if (getenv("GLC_ATTN_FLAGS") && !getenv("GLC_ATTN_TRACE")) {
    attention_fused_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm_qkv, tm_pk, tm_pq, p);
    return cudaGetLastError();
  }
  if (const char* tf = getenv("GLC_ATTN_TRACE")) {
    const size_t n = 2 * TMAX * 8;
    if (cudaMalloc(&p.trace, n * sizeof(long long)) != cudaSuccess) return cudaGetLastError();
    cudaMemsetAsync(p.trace, 0, n * sizeof(long long), stream);
    attention_fused_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm_qkv, tm_pk, tm_pq, p);
    cudaError_t e = cudaStreamSynchronize(stream);
    std::vector<long long> h(n);
    cudaMemcpy(h.data(), p.trace, n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    if (FILE* f = fopen(tf, "w")) {
      for (int role = 0; role < 2; ++role)
        for (int t = 0; t < TMAX; ++t) {
          if (!h[(role * TMAX + t) * 8 + 1]) continue;
          fprintf(f, "%s t=%d", role ? "mma" : "smx", t);
          for (int k = 0; k < 8; ++k) fprintf(f, " %lld", h[(role * TMAX + t) * 8 + k] - h[1]);
          fprintf(f, "\n");
        }
      fclose(f);
    }
    return e;
  }
  attention_fused_kernel<false><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm_qkv, tm_pk, tm_pq, p);
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
