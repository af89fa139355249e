// K3 (superseded generation, kept for reference; not part of libgliclass_b200.so) — fused DeBERTa disentangled attention for sm_100a, head dim 64:
//
//   ctx[b,i,h,:] = softmax_j( (Q_i.K_j + Q_i.posK[idx(i-j)] + K_j.posQ[idx(i-j)]) / sqrt(3d) + mask_j ) . V_j
//
// (transformers DisentangledSelfAttention, T:229-345: scores T:237-251, c2p T:313-324, p2c T:327-343, masked
// softmax T:256-259, context T:262-271; idx(delta) = clamp(bucket(delta)+span, 0, 2*span-1))
// Replaces the attention sub-graph ORT executes inside the reference's session Run (src/model.c:173-182).
//
// Tensor-core side (unchanged from attention_shift.cu): the position tables are expanded at load to one row per DELTA
//     EK [rho]   = posK[idx(2047 - rho)]          EQr[sigma] = posQ[idx(sigma - 2047)]
// so that inside a (128-query, 64-key) tile (a = query row, b = key column) both biases are Toeplitz,
//     c2p[a,b] = C[a, 127 - a + b]     C = Q_tile . EK[rho0 .. rho0+191]^T          rho0   = 1920 - q0 + k0
//     p2c[a,b] = G[a + 63 - b, b]      G = EQr[sigma0 .. sigma0+191] . K_tile^T     sigma0 = 1984 + q0 - k0
// and the skews are undone in registers: c2p by a 5-stage barrel shifter on the packed fp16 window of C, p2c by a lane
// rotation (the tensor core computes G three times, with the A window starting at slice rows 0 / 32 / 64).
//
// What is new here is the softmax side.  ncu on attention_shift (profiles/r1e_kernels_ncu.md) shows 61 % of the issue
// slots empty, no pipe above 30 %: each of its 8 softmax warps runs a ~480-instruction dependent chain per tile at one
// instruction per ~7 cycles, two warps per scheduler, the two warps of a row in lock step (row-maximum exchange through
// shared memory + a named barrier, then both in the MUFU-bound exponentials at the same time).  Here
//   * a softmax thread owns a WHOLE row of a tile (64 keys: two 32-key halves through the same register-skew code), so
//     there is no exchange, no named barrier, one straight-line code path for every softmax warp;
//   * NWG = 3 warpgroups (12 softmax warps, three per scheduler) ROTATE over the key tiles: tile t belongs to group
//     t mod 3.  The accumulators S | C | G stay single-buffered in TMEM (512 columns, as before): the tensor core
//     computes tile t+1 as soon as group t has drained tile t into registers, which happens in the first third of its work;
//   * O stays in TMEM and accumulates over all tiles with a STICKY row maximum (raised only when the row maximum grew by
//     more than 2^8).  The maximum is chained from tile to tile: group t reads m(t-1) published by group t-1 (shared
//     memory + one mbarrier per (group, lane quarter)), decides m(t), publishes it, then runs its exponentials — while
//     group t+1 is still in its skew stages.  The (rare) rescale of O is done by the group that raises the maximum;
//   * each group keeps a partial row sum relative to the maximum it last saw; the group of the last tile merges them.
//
//   * the C window SLIDES: consecutive key tiles of a query tile use slices shifted by 64 table rows, so C lives in a ring
//     of three 64-column blocks and the tensor core computes only the new block per tile (N = 64 instead of 192; the
//     first tile computes all three).  A group drains S and both halves of its C window into registers in one go and
//     frees the accumulators before it starts shifting, which is what paces the tile pipeline (drain -> S|C MMA -> drain).
//
// Warps: 0 TMA producer (K + position slices, 3 stages) | 1 tcgen05 issuer for S, C, G | 2 tcgen05 issuer for P.V |
// 3 TMA producer (Q, V) | 4..15 softmax (group =
// (warp-4)/4, TMEM lane quarter = warp % 4).  setmaxnreg moves registers from the utility warpgroup to the softmax ones.
// Synchronisation is mbarrier based; every wait is bounded (ptx::mbar_wait traps on timeout).  Barriers that a softmax
// group waits on are indexed by t mod 3: a group can run two tiles ahead of a slow one, and a single phase bit would alias.
#include <cuda_fp16.h>
#include <math_constants.h>

#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "kernels.h"
#include "kernels_exp.h"
#include "ptx.cuh"
#include "tma_desc.h"

namespace glc {
namespace {

constexpr int QT = 128;            // queries per CTA
constexpr int KT = 64;             // keys per tile
constexpr int D = 64;              // head dim
constexpr int NWG = 3;             // softmax warpgroups rotating over the key tiles
constexpr int SLICE = 192;         // table rows per tile (191 deltas + the never-consumed last row of copy 64)
constexpr int EXP_CENTER = 2047;
constexpr int EXP_ROWS = 4096;
constexpr int ROWS_THREADS = 128 + 128 * NWG;
constexpr int UTIL_REGS = 40;      // setmaxnreg of the utility warpgroup ...
constexpr int SMX_REGS = 152;      // ... and of the softmax warpgroups: 128 * 40 + 384 * 152 = 63488 <= 65536 = 512 * 128

// shared memory map (bytes, from a 1024-aligned base)
constexpr int OFF_Q = 0;                           // 128 x 128 B
constexpr int ASTAGES = 3;                         // ring A (K_t + position slices) depth: a TMA round trip is ~2 tiles long
constexpr int OFF_K = OFF_Q + 16384;               // ASTAGES x 8 KB
constexpr int OFF_V = OFF_K + ASTAGES * 8192;      // 2 x 8 KB
constexpr int POS_BYTES = SLICE * 128;
constexpr int OFF_EK = OFF_V + 16384;              // ASTAGES x 192 x 128 B (tile 0 uses all 192 rows, later tiles 64)
constexpr int OFF_EQ = OFF_EK + ASTAGES * POS_BYTES;     // ASTAGES x 192 x 128 B
constexpr int OFF_MROW = OFF_EQ + ASTAGES * POS_BYTES;   // float[128]: sticky row maximum, chained from tile to tile
constexpr int OFF_LSUM = OFF_MROW + QT * 4;        // float[NWG][2][128]: (partial row sum, maximum it is relative to) per group
constexpr int OFF_MASK = OFF_LSUM + NWG * 2 * QT * 4;  // uint32[68]: key-validity words of this batch row
constexpr int HM_PITCH = 80;                       // bytes per lane row of the p2c copy-select masks (16 words + pad: conflict-free LDS.128)
constexpr int OFF_HM = OFF_MASK + 68 * 4;          // uint32[32][20]
constexpr int OFF_BAR = OFF_HM + 32 * HM_PITCH;
constexpr int NUM_BARS = 48;
constexpr int ROWS_SMEM = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
static_assert(OFF_BAR % 8 == 0, "barrier alignment");
static_assert(ROWS_SMEM <= 227 * 1024, "attention smem budget");

// TMEM columns
constexpr uint32_t TM_S = 0;       // 64
constexpr uint32_t TM_Q = 64;      // 32: Q tile as fp16 pairs
constexpr uint32_t TM_P = 96;      // 32: P tile as fp16 pairs
constexpr uint32_t TM_C = 128;     // 192 (fp16 accumulators, one per 32-bit column)
constexpr uint32_t TM_G32 = 320;   // 64: window rows 32..159, keys 0..63
constexpr uint32_t TM_G64 = 384;   // 32: window rows 64..191, keys 0..31
constexpr uint32_t TM_G0 = 416;    // 32: window rows 0..127,  keys 32..63
constexpr uint32_t TM_PV = 448;    // 64

struct RowsParams {
  const uint32_t* mask_bits; // [B][ceil(S/32)]
  const int32_t* kv_len;     // [B]
  __half* ctx;               // [B*S, H]
  int B, S, heads, H;
  float scale_log2;          // log2(e) / sqrt(3*d)
  long long* trace;          // TRACE instantiation only: [16 warps][TMAX][8] clock64 stamps of CTA (1,0,0)
};

constexpr int TMAX = 2048 / KT;
#define ROWS_TRACE(tile, slot)                                                                              \
  do {                                                                                                      \
    if (TRACE && p.trace && (threadIdx.x & 31) == 0 && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0) \
      p.trace[((threadIdx.x >> 5) * TMAX + (tile)) * 8 + (slot)] = clock64();                               \
  } while (0)

__device__ __forceinline__ uint32_t sel(bool p, uint32_t a, uint32_t b) { return p ? a : b; }

// 2^x on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial on [-0.5, 0.5], max relative error 7.5e-5 — well
// below the fp16 rounding of P): the exponential stage is bound by the MUFU unit (8 cycles per warp instruction) while
// the FMA pipe idles, so every 4th score of a thread takes this route instead.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float r = x - (t - 12582912.0f);    // [-0.5, 0.5]
  float p = fmaf(0.05517090f, r, 0.24260953f);
  p = fmaf(p, r, 0.69326097f);
  p = fmaf(p, r, 0.99992818f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <uint32_t N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <uint32_t N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <bool TRACE>
__global__ void __launch_bounds__(ROWS_THREADS, 1)
attention_rows_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_ek,
                      const __grid_constant__ CUtensorMap tm_eq, const RowsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* qt_full = bars + 1;      // Q tile copied into TMEM (4 warps of group 0)
  uint64_t* a_full = bars + 40;      // [ASTAGES]  K_t + slices landed
  uint64_t* a_empty = bars + 43;     // [ASTAGES]  ... consumed by the S / C / G MMAs
  uint64_t* b_full = bars + 6;       // [2]  V_t landed
  uint64_t* b_empty = bars + 8;      // [2]  ... consumed by the PV MMA
  uint64_t* sc_free = bars + 10;     // S and C accumulators drained (4 warps of the tile's group)
  uint64_t* g_free = bars + 11;      // G copies drained
  uint64_t* p_full = bars + 12;      // P tile written (4 warps)
  uint64_t* l_bar = bars + 14;       // partial row sums published (4 * NWG warps)
  uint64_t* sc_full = bars + 16;     // [NWG] S and C accumulators of tile t ready      (index t % NWG)
  uint64_t* g_full = bars + 20;      // [NWG] G copies of tile t ready
  uint64_t* pv_full = bars + 24;     // [NWG] O += P.V of tile t complete
  uint64_t* m_bar = bars + 28;       // [NWG][4] row maxima of the group's latest tile published, per lane quarter
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  uint32_t* kmask = reinterpret_cast<uint32_t*>(smem + OFF_MASK);
  float* mrow = reinterpret_cast<float*>(smem + OFF_MROW);
  float* lsum = reinterpret_cast<float*>(smem + OFF_LSUM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int kvlen = p.kv_len[b];
  const int T = (kvlen + KT - 1) / KT;   // key tiles with at least one valid key

  if (q0 >= kvlen) {
    // padded queries only: their outputs are never read by valid rows
    for (int e = threadIdx.x; e < QT * 8; e += ROWS_THREADS) {
      const int r = q0 + (e >> 3);
      if (r < p.S)
        *reinterpret_cast<uint4*>(p.ctx + ((int64_t)b * p.S + r) * p.H + head * D + (e & 7) * 8) = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tm_qkv);
    ptx::prefetch_tensormap(&tm_ek);
    ptx::prefetch_tensormap(&tm_eq);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(qt_full, 4);
    for (int s = 0; s < ASTAGES; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
    ptx::mbar_init(sc_free, 4);
    ptx::mbar_init(g_free, 4);
    ptx::mbar_init(p_full, 4);
    ptx::mbar_init(l_bar, 4 * NWG);
    for (int k = 0; k < NWG; ++k) {
      ptx::mbar_init(&sc_full[k], 1);
      ptx::mbar_init(&g_full[k], 1);
      ptx::mbar_init(&pv_full[k], 1);
      for (int q = 0; q < 4; ++q) ptx::mbar_init(&m_bar[k * 4 + q], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  {
    const int words = (p.S + 31) >> 5;
    const int e = threadIdx.x - 128;
    if (e >= 0 && e < 68) kmask[e] = (e < words) ? __ldg(p.mask_bits + (int64_t)b * words + e) : 0u;
    // p2c copy-select masks: lane L SUPPLIES the upper G copy for key column jj (of a 32-key half) iff jj < 31 - L;
    // word m of lane L covers the packed pair (2m, 2m+1)
    for (int k = threadIdx.x; k < 32 * 16; k += ROWS_THREADS) {
      const int L = k >> 4, m = k & 15;
      const uint32_t lo16 = (2 * m < 31 - L) ? 0xffffu : 0u;
      const uint32_t hi16 = (2 * m + 1 < 31 - L) ? 0xffff0000u : 0u;
      *reinterpret_cast<uint32_t*>(smem + OFF_HM + L * HM_PITCH + m * 4) = lo16 | hi16;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<UTIL_REGS>();
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer A: K_x and the position slices of tile x
      //      (ring of ASTAGES stages; a stage is free once the bias MMAs of tile x - ASTAGES retired)
      if (lane == 0) {
        for (int x = 0; x < T; ++x) {
          const int k0 = x * KT;
          const int rho0 = EXP_CENTER - (QT - 1) - q0 + k0;
          const int sig0 = EXP_CENTER - (KT - 1) + q0 - k0;
          const int st = x % ASTAGES;
          ptx::mbar_wait(&a_empty[st], (uint32_t)(((x / ASTAGES) & 1) ^ 1));
          // the C window slides by 64 table rows per key tile: only tile 0 needs the whole 192-row EK slice
          ptx::mbar_arrive_expect_tx(&a_full[st], (uint32_t)(KT * 128 + POS_BYTES + (x == 0 ? POS_BYTES : 8192)));
          ptx::tma_load_3d(smem + OFF_K + st * 8192, &tm_qkv, &a_full[st], p.H + head * D, k0, b);
          if (x == 0) {
#pragma unroll
            for (int bx = 0; bx < SLICE / 64; ++bx)
              ptx::tma_load_3d(smem + OFF_EK + st * POS_BYTES + bx * 8192, &tm_ek, &a_full[st], 0, rho0 + bx * 64, head);
          } else {
            ptx::tma_load_3d(smem + OFF_EK + st * POS_BYTES, &tm_ek, &a_full[st], 0, rho0 + 128, head);
          }
#pragma unroll
          for (int bx = 0; bx < SLICE / 64; ++bx)
            ptx::tma_load_3d(smem + OFF_EQ + st * POS_BYTES + bx * 8192, &tm_eq, &a_full[st], 0, sig0 + bx * 64, head);
        }
      }
      __syncwarp();
    } else if (warp == 3) {
      // ------------------------------------------------------------------ TMA producer B: Q once, then V_x (two stages; a stage
      //      is free once the PV MMA of tile x - 2 retired).  Its own warp: the V ring must never hold back the A ring.
      if (lane == 0) {
        ptx::mbar_arrive_expect_tx(q_full, QT * 128);
        ptx::tma_load_3d(smem + OFF_Q, &tm_qkv, q_full, head * D, q0, b);
        ptx::tma_load_3d(smem + OFF_Q + 8192, &tm_qkv, q_full, head * D, q0 + 64, b);
        for (int x = 0; x < T; ++x) {
          const int st = x & 1;
          ptx::mbar_wait(&b_empty[st], (uint32_t)(((x >> 1) & 1) ^ 1));
          ptx::mbar_arrive_expect_tx(&b_full[st], KT * 128);
          ptx::tma_load_3d(smem + OFF_V + st * 8192, &tm_qkv, &b_full[st], 2 * p.H + head * D, x * KT, b);
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer A: S | C and the G copies of every tile
      constexpr uint32_t idesc_s = ptx::idesc_f16(128, 64);
      constexpr uint32_t idesc_c = ptx::idesc_f16(128, SLICE, 0, 0, ptx::FMT_F16, ptx::FMT_F16, 0u);   // fp16 accumulators
      constexpr uint32_t idesc_c64 = ptx::idesc_f16(128, 64, 0, 0, ptx::FMT_F16, ptx::FMT_F16, 0u);
      constexpr uint32_t idesc_g64 = ptx::idesc_f16(128, 64, 0, 0, ptx::FMT_F16, ptx::FMT_F16, 0u);
      constexpr uint32_t idesc_g32 = ptx::idesc_f16(128, 32, 0, 0, ptx::FMT_F16, ptx::FMT_F16, 0u);
      ptx::mbar_wait(qt_full, 0);
      ptx::tc_fence_after();
      for (int t = 0; t < T; ++t) {
        const int st = t % ASTAGES;
        const uint64_t dK = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_K + st * 8192));
        const uint64_t dEK = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_EK + st * POS_BYTES));
        const uint64_t dEQ = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_EQ + st * POS_BYTES));
        ptx::mbar_wait(&a_full[st], (uint32_t)((t / ASTAGES) & 1));
        ROWS_TRACE(t, 0);
        if (t > 0) ptx::mbar_wait(sc_free, (t - 1) & 1);   // S and C accumulators drained by the group of tile t-1
        ptx::tc_fence_after();
        ROWS_TRACE(t, 1);
        if (ptx::elect_one()) {
          // S = Q.K_t^T and C = Q.EK_slice^T (A = Q from TMEM): 16 halves along K = 8 TMEM columns per step
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ts(tmem + TM_S, tmem + TM_Q + 8 * k, dK + 2 * k, idesc_s, (uint32_t)(k != 0));
          // C ring: block j of tile t sits in slot (t + j) % 3; tile 0 fills all three, tile t > 0 only its new block j = 2
          if (t == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::mma_f16_ts(tmem + TM_C, tmem + TM_Q + 8 * k, dEK + 2 * k, idesc_c, (uint32_t)(k != 0));
          } else {
            const uint32_t dC = tmem + TM_C + (uint32_t)(64 * ((t + 2) % 3));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::mma_f16_ts(dC, tmem + TM_Q + 8 * k, dEK + 2 * k, idesc_c64, (uint32_t)(k != 0));
          }
          ptx::mma_commit(&sc_full[t % NWG]);
        }
        __syncwarp();
        if (t > 0) ptx::mbar_wait(g_free, (t - 1) & 1);
        ptx::tc_fence_after();
        ROWS_TRACE(t, 2);
        if (ptx::elect_one()) {
          // the three row-shifted copies of G = EQr_slice . K_t^T; 32 table / key rows = 256 in a descriptor
#pragma unroll
          for (int k = 0; k < 4; ++k)   // rows 32..159 x keys 0..63
            ptx::mma_f16_ss(tmem + TM_G32, dEQ + 256 + 2 * k, dK + 2 * k, idesc_g64, (uint32_t)(k != 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // rows 64..191 x keys 0..31
            ptx::mma_f16_ss(tmem + TM_G64, dEQ + 512 + 2 * k, dK + 2 * k, idesc_g32, (uint32_t)(k != 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // rows 0..127 x keys 32..63
            ptx::mma_f16_ss(tmem + TM_G0, dEQ + 2 * k, dK + 256 + 2 * k, idesc_g32, (uint32_t)(k != 0));
          ptx::mma_commit(&a_empty[st]);
          ptx::mma_commit(&g_full[t % NWG]);
        }
        __syncwarp();
      }
    } else if (warp == 2) {
      // ------------------------------------------------------------------ MMA issuer B: O (+)= P . V_t
      constexpr uint32_t idesc_pv = ptx::idesc_f16(128, 64, 0, 1);   // B (= V) is MN-major
      for (int t = 0; t < T; ++t) {
        const int st = t & 1;
        const uint64_t dV = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_V + st * 8192));
        ptx::mbar_wait(&b_full[st], (t >> 1) & 1);
        ptx::mbar_wait(p_full, t & 1);
        ptx::tc_fence_after();
        ROWS_TRACE(t, 0);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)   // V is MN-major: 16 keys further = +2048 bytes = +128 in the descriptor
            ptx::mma_f16_ts(tmem + TM_PV, tmem + TM_P + 8 * k, dV + 128 * k, idesc_pv, (uint32_t)(k != 0 || t > 0));
          ptx::mma_commit(&b_empty[st]);
          ptx::mma_commit(&pv_full[t % NWG]);
        }
        __syncwarp();
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax warpgroups
    setmaxnreg_inc<SMX_REGS>();
    const int wg = (warp - 4) >> 2;   // tiles wg, wg + NWG, ...
    const int qd = warp & 3;          // TMEM lane quarter
    const int i = qd * 32 + lane;     // row in the query tile
    const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
    const float sc = p.scale_log2;

    if (wg == 0) {
      // ---- Q tile -> TMEM once: this thread's row, 8 16-byte chunks of the swizzled 128-byte row
      ptx::mbar_wait(q_full, 0);
      const uint8_t* qrow = smem + OFF_Q + (i >> 3) * 1024 + (i & 7) * 128;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint4 lo = *reinterpret_cast<const uint4*>(qrow + (((2 * u) ^ (i & 7)) << 4));
        const uint4 hi = *reinterpret_cast<const uint4*>(qrow + (((2 * u + 1) ^ (i & 7)) << 4));
        const uint32_t qr[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        ptx::tmem_st_x8(t_lane + TM_Q + (uint32_t)(8 * u), qr);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(qt_full);
    }

    // tile-independent shift controls (both 32-key halves of a tile use the same ones: their first key is 0 mod 32)
    const int sh = 31 - lane;                  // c2p: element shift inside the window
    const bool sh16 = sh & 16, sh8 = sh & 8, sh4 = sh & 4, sh2 = sh & 2;
    const uint32_t prmt_sel = (sh & 1) ? 0x5432u : 0x3210u;
    const int rot0 = lane + 31;                // p2c: source lane of key column jj (of a half) is (rot0 - jj) & 31
    const uint4* hm_row = reinterpret_cast<const uint4*>(smem + OFF_HM + lane * HM_PITCH);   // this lane's copy-select masks

    float l_part = 0.f;                  // this group's partial row sum, relative to m_seen
    float m_seen = -CUDART_INF_F;        // sticky maximum this group used for its latest tile

    for (int t = wg; t < T; t += NWG) {
      const int k0 = t * KT;
      const int slot = t % NWG;          // == wg
      const uint32_t par = (uint32_t)((t / NWG) & 1);
      float s[KT];

      // ---- drain S and this warp's two 64-column windows of C in one go, then hand the accumulators back: the next
      //      tile's S | C MMA waits for exactly this (it is what paces the tile pipeline)
      ROWS_TRACE(t, 0);
      ptx::mbar_wait(&sc_full[slot], par);
      ptx::tc_fence_after();
      ROWS_TRACE(t, 1);
      uint32_t w0[32], w1[32];
      {
        // window of (warp quarter qd, half hf) starts at window column 96 - 32 qd + 32 hf; window block j is ring slot
        // (t + j) % 3, i.e. window column c is ring column (64 (t % 3) + c) % 192: four 32-column pieces
        const int base = 64 * (t % 3) + 96 - 32 * qd;
        uint32_t r0[32], r1[32], c0[16], c1[16], c2[16];
        ptx::tmem_ld_x32(t_lane + TM_S, r0);
        ptx::tmem_ld_x32(t_lane + TM_S + 32, r1);
        ptx::tmem_ld_x16_pack16(t_lane + TM_C + (uint32_t)(base % 192), c0);
        ptx::tmem_ld_x16_pack16(t_lane + TM_C + (uint32_t)((base + 32) % 192), c1);
        ptx::tmem_ld_x16_pack16(t_lane + TM_C + (uint32_t)((base + 64) % 192), c2);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(sc_free);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          s[jj] = __uint_as_float(r0[jj]);
          s[32 + jj] = __uint_as_float(r1[jj]);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          w0[k] = c0[k];
          w0[16 + k] = c1[k];
          w1[k] = c1[k];
          w1[16 + k] = c2[k];
        }
      }
      ROWS_TRACE(t, 2);
      auto c2p_half = [&](auto half_tag, uint32_t (&w)[32]) {
        constexpr int HF = decltype(half_tag)::value;
        // shift the packed window left by sh elements
#pragma unroll
        for (int k = 0; k < 32 - 8; ++k) w[k] = sel(sh16, w[k + 8], w[k]);
#pragma unroll
        for (int k = 0; k < 32 - 12; ++k) w[k] = sel(sh8, w[k + 4], w[k]);
#pragma unroll
        for (int k = 0; k < 32 - 14; ++k) w[k] = sel(sh4, w[k + 2], w[k]);
#pragma unroll
        for (int k = 0; k < 32 - 15; ++k) w[k] = sel(sh2, w[k + 1], w[k]);
#pragma unroll
        for (int m = 0; m < 16; ++m)
          ptx::add_f16x2_to_f32(s[32 * HF + 2 * m], s[32 * HF + 2 * m + 1], __byte_perm(w[m], w[m + 1], prmt_sel));
      };
      c2p_half(std::integral_constant<int, 0>{}, w0);
      c2p_half(std::integral_constant<int, 1>{}, w1);

      // ---- p2c, both halves: lane rotation by 31 - (b mod 32), the source lane picks the copy
      ptx::mbar_wait(&g_full[slot], par);
      ptx::tc_fence_after();
      ROWS_TRACE(t, 3);
      auto p2c_half = [&](auto half_tag) {
        constexpr int HF = decltype(half_tag)::value;
        // keys 0..31: copies G32 (lower) / G64 (upper); keys 32..63: copies G0 (lower) / G32 (upper)
        const uint32_t a_lo = t_lane + (HF == 0 ? TM_G32 : TM_G0);
        const uint32_t a_hi = t_lane + (HF == 0 ? TM_G64 : TM_G32 + 32);
        uint32_t lo[16], hi[16];
        ptx::tmem_ld_x16_pack16(a_lo, lo);
        ptx::tmem_ld_x16_pack16(a_hi, hi);
        ptx::tmem_ld_wait();
        if (HF == 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(g_free);
        }
#pragma unroll
        for (int m4 = 0; m4 < 4; ++m4) {
          const uint4 hq = hm_row[m4];
          const uint32_t hmv[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
          for (int mm = 0; mm < 4; ++mm) {
            const int m = 4 * m4 + mm;
            const uint32_t v = (hi[m] & hmv[mm]) | (lo[m] & ~hmv[mm]);
            const uint32_t x0 = __shfl_sync(0xffffffffu, v, rot0 - 2 * m);        // low half = key 2m
            const uint32_t x1 = __shfl_sync(0xffffffffu, v, rot0 - 2 * m - 1);    // high half = key 2m + 1
            ptx::add_f16_lo_to_f32(s[32 * HF + 2 * m], x0);
            ptx::add_f16_hi_to_f32(s[32 * HF + 2 * m + 1], x1);
          }
        }
      };
      p2c_half(std::integral_constant<int, 0>{});
      p2c_half(std::integral_constant<int, 1>{});

      // ---- key validity, local row maximum (four independent chains of 3-input maxima)
      {
        const uint32_t kb0 = kmask[k0 >> 5], kb1 = kmask[(k0 >> 5) + 1];
        if ((kb0 & kb1) != 0xffffffffu) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            if (!((kb0 >> jj) & 1u)) s[jj] = -CUDART_INF_F;
            if (!((kb1 >> jj) & 1u)) s[32 + jj] = -CUDART_INF_F;
          }
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

      // ---- chain the sticky maximum: m(t) from m(t-1) of the previous tile's group and this tile's row maximum
      ROWS_TRACE(t, 4);
      float m_prev = -CUDART_INF_F;
      if (t > 0) {
        const int pg = (t - 1) % NWG;
        ptx::mbar_wait(&m_bar[pg * 4 + qd], (uint32_t)(((t - 1) / NWG) & 1));
        m_prev = mrow[i];
      }
      // raised only when the row maximum grew by more than 2^8 (P <= 2^8 keeps fp16 headroom); false when both are -inf
      // (NaN), true for the first finite maximum
      const bool raise = (mloc - m_prev) * sc > 8.0f;
      const float m_cur = raise ? mloc : m_prev;
      mrow[i] = m_cur;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&m_bar[wg * 4 + qd]);
      ROWS_TRACE(t, 5);
      const float alpha = raise ? ptx::ex2((m_prev - m_cur) * sc) : 1.0f;   // m_prev = -inf: 0 (O holds nothing yet)
      const float neg_ms = (m_cur == -CUDART_INF_F) ? 0.f : -m_cur * sc;
      // this group's partial sum follows the maximum it last saw
      l_part *= (m_seen == m_cur) ? 1.0f : ((m_seen == -CUDART_INF_F) ? 0.f : ptx::ex2((m_seen - m_cur) * sc));
      m_seen = m_cur;

      // ---- exponentials (every 4th on the FMA pipe), row sum in four chains, P as fp16 pairs
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

      // ---- P buffer free again and O stable once the PV of tile t-1 has completed; rare rescale of O by the raising group
      ROWS_TRACE(t, 6);
      if (t > 0) {
        ptx::mbar_wait(&pv_full[(t - 1) % NWG], (uint32_t)(((t - 1) / NWG) & 1));
        if (__any_sync(0xffffffffu, raise)) {
          ptx::tc_fence_after();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t r[32];
            ptx::tmem_ld_x32(t_lane + TM_PV + (uint32_t)(32 * hf), r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) r[jj] = __float_as_uint(__uint_as_float(r[jj]) * alpha);
            ptx::tmem_st_x32(t_lane + TM_PV + (uint32_t)(32 * hf), r);
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
      ROWS_TRACE(t, 7);
    }

    // ---- final merge: every group publishes (partial row sum, the maximum it is relative to); the group of the last tile
    //      — whose m_seen is the final maximum — rescales and adds them, normalises O and writes the context rows
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
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t r[32];
        ptx::tmem_ld_x32(t_lane + TM_PV + (uint32_t)(32 * hf), r);
        ptx::tmem_ld_wait();
        if (row < p.S) {
          __half* dst = p.ctx + ((int64_t)b * p.S + row) * p.H + head * D + 32 * hf;
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

cudaError_t attention_rows(const void* qkv, const void* exp_k, const void* exp_qr, int64_t ld_exp,
                           const uint32_t* mask_bits, const int32_t* kv_len, void* ctx, int B, int S, int heads,
                           cudaStream_t stream) {
  if (B <= 0 || S <= 0) return cudaSuccess;
  if (S > 2048) return cudaErrorInvalidValue;   // reference MAX_LENGTH (include/configs.h:5)
  const int H = heads * D;
  // qkv viewed as [B][S][3H]; box 64 cols x 64 rows
  uint64_t dq[3] = {(uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
  uint64_t sq[2] = {(uint64_t)(3 * H) * 2, (uint64_t)S * 3 * H * 2};
  uint32_t bq[3] = {64, 64, 1};
  // expanded tables are [EXP_ROWS][ld_exp] row-major (head h = columns h*64..): dims (d, row, head)
  uint64_t dp[3] = {64, (uint64_t)EXP_ROWS, (uint64_t)heads};
  uint64_t sp[2] = {(uint64_t)ld_exp * 2, 128};
  uint32_t bp[3] = {64, 64, 1};
  CUtensorMap tm_qkv = make_tmap_16b(qkv, 3, dq, sq, bq);
  CUtensorMap tm_ek = make_tmap_16b(exp_k, 3, dp, sp, bp);
  CUtensorMap tm_eq = make_tmap_16b(exp_qr, 3, dp, sp, bp);
  RowsParams p;
  p.mask_bits = mask_bits;
  p.kv_len = kv_len;
  p.ctx = (__half*)ctx;
  p.B = B; p.S = S; p.heads = heads; p.H = H;
  p.scale_log2 = 1.4426950408889634f / sqrtf(3.0f * D);
  p.trace = nullptr;
  dim3 grid((S + QT - 1) / QT, heads, B);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attention_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS_SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  // developer aid: GLC_ATTN_TRACE=<file> dumps per-tile clock64 stamps of every warp of CTA (1,0,0) (synchronous)
  if (const char* tf = getenv("GLC_ATTN_TRACE")) {
    const size_t n = (size_t)16 * TMAX * 8;
    if (cudaMalloc(&p.trace, n * sizeof(long long)) != cudaSuccess) return cudaGetLastError();
    cudaMemsetAsync(p.trace, 0, n * sizeof(long long), stream);
    attention_rows_kernel<true><<<grid, ROWS_THREADS, ROWS_SMEM, stream>>>(tm_qkv, tm_ek, tm_eq, p);
    cudaError_t e = cudaStreamSynchronize(stream);
    std::vector<long long> h(n);
    cudaMemcpy(h.data(), p.trace, n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    if (FILE* f = fopen(tf, "w")) {
      long long t0 = 0;
      for (size_t k = 0; k < n; ++k) if (h[k] && (!t0 || h[k] < t0)) t0 = h[k];
      for (int w = 0; w < 16; ++w)
        for (int t = 0; t < TMAX; ++t) {
          bool any = false;
          for (int k = 0; k < 8; ++k) any |= h[((size_t)w * TMAX + t) * 8 + k] != 0;
          if (!any) continue;
          fprintf(f, "warp %2d t=%2d", w, t);
          for (int k = 0; k < 8; ++k) fprintf(f, " %7lld", h[((size_t)w * TMAX + t) * 8 + k] ? h[((size_t)w * TMAX + t) * 8 + k] - t0 : -1);
          fprintf(f, "\n");
        }
      fclose(f);
    }
    return e;
  }
  attention_rows_kernel<false><<<grid, ROWS_THREADS, ROWS_SMEM, stream>>>(tm_qkv, tm_ek, tm_eq, p);
  return cudaGetLastError();
}

}  // namespace glc
