// K3 (production) — fused DeBERTa disentangled attention for sm_100a, head dim 64, PERSISTENT form:
//
//   ctx[b,i,h,:] = softmax_j( (Q_i.K_j + Q_i.posK[idx(i-j)] + K_j.posQ[idx(i-j)]) / sqrt(3d) + mask_j ) . V_j
//
// (transformers DisentangledSelfAttention, T:229-345: scores T:237-251, c2p T:313-324, p2c T:327-343, masked
// softmax T:256-259, context T:262-271; idx(delta) = clamp(bucket(delta)+span, 0, 2*span-1))
// Replaces the attention sub-graph ORT executes inside the reference's session Run (src/model.c:173-182).
//
// Tensor-core side: the position tables are expanded at load to one row per DELTA
//     EK [rho]   = posK[idx(2047 - rho)]          EQr[sigma] = posQ[idx(sigma - 2047)]
// so that inside a (128-query, 64-key) tile (a = query row, b = key column) both biases are Toeplitz,
//     c2p[a,b] = C[a, 127 - a + b]     C = Q_tile . EK[rho0 .. rho0+191]^T          rho0   = 1920 - q0 + k0
//     p2c[a,b] = G[a + 63 - b, b]      G = EQr[sigma0 .. sigma0+191] . K_tile^T     sigma0 = 1984 + q0 - k0
// and the skews are undone in registers, where the accumulators already are: c2p by a 5-stage barrel shifter on the packed
// fp16 window of C (tcgen05.ld.pack::16b), p2c by a lane rotation — the tensor core computes G three times, with the A
// window starting at slice rows 0 / 32 / 64, so that row a + s of the skew sits in lane (a + s) mod 32 of one of two copies:
// one select (precomputed lane masks) + one SHFL per score.  C and G use fp16 accumulators.
// Softmax side:
//   * a softmax thread owns a WHOLE row of a tile (64 keys: two 32-key halves through the same register-skew code): no
//     cross-warp exchange, one straight-line code path for every softmax warp;
//   * NWG = 3 warpgroups (12 softmax warps, three per scheduler) ROTATE over the key tiles: tile g belongs to group g mod 3.
//     S | C | G stay single-buffered in TMEM (all 512 columns are used): the tensor core computes the next tile as soon as
//     the group of the current one has drained it into registers;
//   * O stays in TMEM and accumulates over all tiles with a STICKY row maximum (raised only when the row maximum grew by
//     more than 2^8).  The maximum is chained from tile to tile through shared memory + one mbarrier per (group, lane
//     quarter); the (rare) rescale of O is done by the group that raises the maximum; each group keeps a partial row sum
//     relative to the maximum it last saw, merged by the group of the item's last tile in tile-residue order;
//   * the C window SLIDES: consecutive key tiles use slices shifted by 64 table rows, so C lives in a ring of three
//     64-column blocks and the tensor core computes only the new block per tile (the first tile computes all three).
//
// What the per-tile clock traces of the one-CTA-per-item form of this kernel showed (experiments/attention_generations/
// attention_rows.cu, profiles/r2_attention_trace.md): a CTA spends ~3000 cycles in
// its prologue (barrier init, TMEM allocation, first TMA round trip) and ~1500 in its tail for 8 tiles of work, and the
// tile period (1750 cycles) is set by the L2 -> SM fabric: 48 KB per tile, of which 32 KB are position-table slices that
// every CTA of the same (head, query tile) fetches again.  Hence this kernel:
//   * one CTA per SM for the whole launch; work items (head, query tile, batch row) are dealt out in contiguous chunks of
//     the list ordered with the batch row fastest, so a CTA stays on one (head, query tile) for all but one switch;
//   * for S <= 512 both expanded table windows of that (head, query tile) — (Tmax + 2) * 64 rows each, 80 KB + 80 KB —
//     stay RESIDENT in shared memory; a tile then costs 16 KB of L2 traffic (K_t, V_t).  Longer sequences stream the
//     slices per tile (MODE 0);
//   * tiles of consecutive items form one stream g = 0, 1, 2, ...: group g % 3 owns tile g, the S | C | G accumulators,
//     the P buffer and the operand rings are handed over by the same barriers across item boundaries, so the pipeline never
//     drains.  Per item only the Q tile (copied to TMEM by the group that owns the item's first tile), the O accumulator
//     (released by the group of the last tile once it has read it) and the row-maximum / row-sum slots change hands.
//
// Warps: 0 TMA producer A (tables, K) | 1 tcgen05 issuer for S, C, G | 2 tcgen05 issuer for P.V | 3 TMA producer B (Q, V) |
// 4..15 softmax (group = (warp-4)/4, TMEM lane quarter = warp % 4).  Every wait is bounded (ptx::mbar_wait traps).
// Barriers a softmax group waits on are arrays indexed by (tile % 3) or (item % 4): a group can be up to two tiles /
// three items ahead of the slowest one, and a single phase bit would alias.
#include <cuda_fp16.h>
#include <math_constants.h>

#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"
#include "tma_desc.h"

namespace glc {
namespace {

constexpr int QT = 128;            // queries per item
constexpr int KT = 64;             // keys per tile
constexpr int D = 64;              // head dim
constexpr int NWG = 3;             // softmax warpgroups rotating over the key tiles
constexpr int SLICE = 192;         // table rows per tile (191 deltas + the never-consumed last row of copy 64)
constexpr int EXP_CENTER = 2047;
constexpr int EXP_ROWS = 4096;
constexpr int PTHREADS = 128 + 128 * NWG;
constexpr int UTIL_REGS = 40;      // setmaxnreg of the utility warpgroup ...
constexpr int SMX_REGS = 152;      // ... and of the softmax warpgroups: 128 * 40 + 384 * 152 = 63488 <= 65536 = 512 * 128
constexpr int TMAX_RES = 8;        // MODE 2 (resident tables): S <= 512
constexpr int RES_ROWS = (TMAX_RES + 2) * 64;   // 640 table rows = 80 KB per table
constexpr int ISLOTS = 4;          // item-level slots (row maxima, partial sums, item barriers)
constexpr int HM_PITCH = 80;       // bytes per lane row of the p2c copy-select masks (16 words + pad: conflict-free LDS.128)

// shared memory map (bytes, from a 1024-aligned base); MODE 2: tables resident, K ring only; MODE 0: K + slices per stage
template <int MODE>
struct Smem {
  static constexpr int ASTAGES = (MODE == 2) ? 2 : 3;
  static constexpr int POS_BYTES = SLICE * 128;
  static constexpr int OFF_Q = 0;                                            // 128 x 128 B
  static constexpr int OFF_K = OFF_Q + 16384;                                // ASTAGES x 8 KB
  static constexpr int OFF_V = OFF_K + ASTAGES * 8192;                       // 2 x 8 KB
  static constexpr int OFF_EK = OFF_V + 16384;                               // MODE 2: RES_ROWS x 128 B; MODE 0: ASTAGES x 24 KB
  static constexpr int EK_BYTES = (MODE == 2) ? RES_ROWS * 128 : ASTAGES * POS_BYTES;
  static constexpr int OFF_EQ = OFF_EK + EK_BYTES;
  static constexpr int OFF_MROW = OFF_EQ + EK_BYTES;                         // float[ISLOTS][128]
  static constexpr int OFF_LSUM = OFF_MROW + ISLOTS * QT * 4;                // float[ISLOTS][NWG][2][128]
  static constexpr int OFF_HM = OFF_LSUM + ISLOTS * NWG * 2 * QT * 4;        // uint32[32][20]
  static constexpr int OFF_BAR = OFF_HM + 32 * HM_PITCH;
  static constexpr int NUM_BARS = 64;
  static constexpr int BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
  static_assert(OFF_BAR % 8 == 0, "barrier alignment");
  static_assert(BYTES <= 227 * 1024, "attention smem budget");
};

// TMEM columns
constexpr uint32_t TM_S = 0;       // 64
constexpr uint32_t TM_Q = 64;      // 32: Q tile as fp16 pairs
constexpr uint32_t TM_P = 96;      // 32: P tile as fp16 pairs
constexpr uint32_t TM_C = 128;     // 192 (fp16 accumulators, one per 32-bit column): ring of three 64-column blocks
constexpr uint32_t TM_G32 = 320;   // 64: window rows 32..159, keys 0..63
constexpr uint32_t TM_G64 = 384;   // 32: window rows 64..191, keys 0..31
constexpr uint32_t TM_G0 = 416;    // 32: window rows 0..127,  keys 32..63
constexpr uint32_t TM_PV = 448;    // 64

struct PersistParams {
  const uint32_t* mask_bits; // [B][ceil(S/32)]; packed layout: one bit per packed row (text b starts at word text_row[b] / 32)
  const int32_t* kv_len;     // [B]
  __half* ctx;               // [B*S, H]; packed layout: [packed rows, H]
  int B, S, heads, H;        // packed layout: S = the longest text's row count (sizes the resident position tables)
  int nq;                    // query tiles per row = ceil(S / 128)
  int n_items;               // heads * nq * B, ordered (head, query tile, batch row) with the batch row fastest
  float scale_log2;          // log2(e) / sqrt(3*d)
  // Packed (varlen) layout, engine.cu run_host: text b occupies rows [text_row[b], text_row[b+1]) of qkv / ctx, a multiple
  // of 128 rows, so no 128-row query tile or 64-row key tile ever spans two texts.  tile_info[t] = (query tile index in
  // its text) << 24 | text, for the n_tiles query tiles ordered (query tile index, text): items are (head, t) with t
  // fastest, which keeps the (head, query tile) runs the resident position tables are reused over.  nullptr = [B,S] layout.
  const int32_t* tile_info;
  const int32_t* text_row;
  int n_tiles;
  const void* qkv_base;      // host-side only (tensor-map construction)
};

struct Item {
  int head, q0, b, T;        // T = key tiles with at least one valid key (0: padded query tile, zero-filled)
  int trow, tb;              // TMA coordinates of the text's position 0: (row, batch) = (0, b), or (text_row[b], 0) when packed
  int grow;                  // row of ctx / qkv holding the text's position 0
  int rows;                  // rows the text owns (S, or its packed row count)
  int mword;                 // first mask word of the text
};

__device__ __forceinline__ Item decode_item(const PersistParams& p, int idx) {
  Item it;
  if (p.tile_info != nullptr) {
    const int info = __ldg(p.tile_info + idx % p.n_tiles);
    it.head = idx / p.n_tiles;
    it.b = info & 0xffffff;
    it.q0 = (info >> 24) * QT;
    it.grow = __ldg(p.text_row + it.b);
    it.rows = __ldg(p.text_row + it.b + 1) - it.grow;
    it.trow = it.grow;
    it.tb = 0;
    it.mword = it.grow >> 5;
  } else {
    it.b = idx % p.B;
    const int r = idx / p.B;
    it.q0 = (r % p.nq) * QT;
    it.head = r / p.nq;
    it.grow = it.b * p.S;
    it.rows = p.S;
    it.trow = 0;
    it.tb = it.b;
    it.mword = it.b * ((p.S + 31) >> 5);
  }
  const int kvlen = __ldg(p.kv_len + it.b);
  it.T = (it.q0 < kvlen) ? (kvlen + KT - 1) / KT : 0;
  return it;
}

__device__ __forceinline__ uint32_t sel(bool p, uint32_t a, uint32_t b) { return p ? a : b; }

// 2^x on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial on [-0.5, 0.5], max relative error 7.5e-5 — well
// below the fp16 rounding of P): every 4th score of a thread takes this route, the others the MUFU unit.
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

template <int MODE, bool POLY>
__global__ void __launch_bounds__(PTHREADS, 1)
attention_persist_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_ek,
                         const __grid_constant__ CUtensorMap tm_eq, const PersistParams p) {
  using SM = Smem<MODE>;
  constexpr int AST = SM::ASTAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
  // tile-stream barriers (g = running tile count of this CTA)
  uint64_t* a_full = bars + 0;       // [AST]  K_g (+ slices in MODE 0) landed
  uint64_t* a_empty = bars + 3;      // [AST]  ... consumed by the S / C / G MMAs of tile g
  uint64_t* b_full = bars + 6;       // [2]    V_g landed
  uint64_t* b_empty = bars + 8;      // [2]    ... consumed by the PV MMA
  uint64_t* sc_free = bars + 10;     // S and C accumulators drained (4 warps of the tile's group)
  uint64_t* g_free = bars + 11;      // G copies drained
  uint64_t* p_full = bars + 12;      // P tile written (4 warps)
  uint64_t* sc_full = bars + 13;     // [NWG] S and C accumulators of tile g ready      (index g % NWG)
  uint64_t* g_full = bars + 16;      // [NWG] G copies of tile g ready
  uint64_t* pv_full = bars + 19;     // [NWG] O (+)= P.V of tile g complete
  uint64_t* m_bar = bars + 22;       // [NWG][4] row maxima of the group's latest tile published, per lane quarter
  // item-level barriers (iq = running count of non-empty items of this CTA)
  uint64_t* q_full = bars + 34;      // [ISLOTS] Q tile of item iq landed in shared memory           (index iq % ISLOTS)
  uint64_t* tmq_free = bars + 38;    // [ISLOTS] every S / C MMA of item iq has read the Q operand in TMEM
  uint64_t* l_bar = bars + 42;       // [ISLOTS] partial row sums of item iq published (4 * NWG warps)
  uint64_t* q_empty = bars + 46;     // Q tile copied to TMEM: shared-memory Q may be reloaded (4 warps)
  uint64_t* tmq_full = bars + 47;    // Q operand of item iq is in TMEM (4 warps)
  uint64_t* o_free = bars + 48;      // O of item iq read out by the group of its last tile (4 warps)
  uint64_t* tab_full = bars + 49;    // MODE 2: resident tables of the current (head, query tile) landed
  uint64_t* tab_free = bars + 50;    // MODE 2: every MMA that read the previous resident tables has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + SM::NUM_BARS);
  float* mrow = reinterpret_cast<float*>(smem + SM::OFF_MROW);
  float* lsum = reinterpret_cast<float*>(smem + SM::OFF_LSUM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int item_lo = (int)(((long long)blockIdx.x * p.n_items) / gridDim.x);
  const int item_hi = (int)(((long long)(blockIdx.x + 1) * p.n_items) / gridDim.x);
  const int Tmax = (p.S + KT - 1) / KT;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tm_qkv);
    ptx::prefetch_tensormap(&tm_ek);
    ptx::prefetch_tensormap(&tm_eq);
    for (int s = 0; s < AST; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
    ptx::mbar_init(sc_free, 4);
    ptx::mbar_init(g_free, 4);
    ptx::mbar_init(p_full, 4);
    for (int k = 0; k < NWG; ++k) {
      ptx::mbar_init(&sc_full[k], 1);
      ptx::mbar_init(&g_full[k], 1);
      ptx::mbar_init(&pv_full[k], 1);
      for (int q = 0; q < 4; ++q) ptx::mbar_init(&m_bar[k * 4 + q], 1);
    }
    for (int k = 0; k < ISLOTS; ++k) {
      ptx::mbar_init(&q_full[k], 1);
      ptx::mbar_init(&tmq_free[k], 1);
      ptx::mbar_init(&l_bar[k], 4 * NWG);
    }
    ptx::mbar_init(q_empty, 4);
    ptx::mbar_init(tmq_full, 4);
    ptx::mbar_init(o_free, 4);
    ptx::mbar_init(tab_full, 1);
    ptx::mbar_init(tab_free, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  // p2c copy-select masks: lane L SUPPLIES the upper G copy for key column jj (of a 32-key half) iff jj < 31 - L;
  // word m of lane L covers the packed pair (2m, 2m+1)
  for (int k = threadIdx.x; k < 32 * 16; k += PTHREADS) {
    const int L = k >> 4, m = k & 15;
    const uint32_t lo16 = (2 * m < 31 - L) ? 0xffffu : 0u;
    const uint32_t hi16 = (2 * m + 1 < 31 - L) ? 0xffff0000u : 0u;
    *reinterpret_cast<uint32_t*>(smem + SM::OFF_HM + L * HM_PITCH + m * 4) = lo16 | hi16;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ptx::pdl_wait();                 // everything above overlapped the previous kernel's tail (launch.h)
  ptx::pdl_launch_dependents();

  if (warp < 4) {
    setmaxnreg_dec<UTIL_REGS>();
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer A: resident tables (MODE 2), K_g
      //      (+ the position slices of the tile in MODE 0)
      if (lane == 0) {
        int g = 0, cur_pair = -1, n_sw = 0;
        for (int idx = item_lo; idx < item_hi; ++idx) {
          const Item it = decode_item(p, idx);
          if (it.T == 0) continue;
          if (MODE == 2) {
            const int pair = it.head * 32 + it.q0 / QT;   // (head, query tile): what the resident tables depend on
            if (pair != cur_pair) {
              // the tables of another (head, query tile): every MMA that read the old ones must have completed
              if (n_sw > 0) ptx::mbar_wait(tab_free, (uint32_t)((n_sw - 1) & 1));
              cur_pair = pair;
              const int rho_base = EXP_CENTER - (QT - 1) - it.q0;                     // EK local row r <-> rho_base + r
              const int sig_base = EXP_CENTER - (KT - 1) + it.q0 - KT * (Tmax - 1);   // EQ local row r <-> sig_base + r
              const int nblk = Tmax + 2;
              ptx::mbar_arrive_expect_tx(tab_full, (uint32_t)(2 * nblk * 8192));
              for (int bx = 0; bx < nblk; ++bx) {
                ptx::tma_load_3d(smem + SM::OFF_EK + bx * 8192, &tm_ek, tab_full, 0, rho_base + bx * 64, it.head);
                ptx::tma_load_3d(smem + SM::OFF_EQ + bx * 8192, &tm_eq, tab_full, 0, sig_base + bx * 64, it.head);
              }
              ++n_sw;
            }
          }
          for (int t = 0; t < it.T; ++t, ++g) {
            const int st = g % AST;
            ptx::mbar_wait(&a_empty[st], (uint32_t)(((g / AST) & 1) ^ 1));
            if (MODE == 2) {
              ptx::mbar_arrive_expect_tx(&a_full[st], KT * 128);
              ptx::tma_load_3d(smem + SM::OFF_K + st * 8192, &tm_qkv, &a_full[st], p.H + it.head * D, it.trow + t * KT, it.tb);
            } else {
              const int k0 = t * KT;
              const int rho0 = EXP_CENTER - (QT - 1) - it.q0 + k0;
              const int sig0 = EXP_CENTER - (KT - 1) + it.q0 - k0;
              // the C window slides by 64 table rows per key tile: only the first tile needs the whole 192-row EK slice
              ptx::mbar_arrive_expect_tx(&a_full[st], (uint32_t)(KT * 128 + SM::POS_BYTES + (t == 0 ? SM::POS_BYTES : 8192)));
              ptx::tma_load_3d(smem + SM::OFF_K + st * 8192, &tm_qkv, &a_full[st], p.H + it.head * D, it.trow + k0, it.tb);
              if (t == 0) {
#pragma unroll
                for (int bx = 0; bx < SLICE / 64; ++bx)
                  ptx::tma_load_3d(smem + SM::OFF_EK + st * SM::POS_BYTES + bx * 8192, &tm_ek, &a_full[st], 0, rho0 + bx * 64, it.head);
              } else {
                ptx::tma_load_3d(smem + SM::OFF_EK + st * SM::POS_BYTES, &tm_ek, &a_full[st], 0, rho0 + 128, it.head);
              }
#pragma unroll
              for (int bx = 0; bx < SLICE / 64; ++bx)
                ptx::tma_load_3d(smem + SM::OFF_EQ + st * SM::POS_BYTES + bx * 8192, &tm_eq, &a_full[st], 0, sig0 + bx * 64, it.head);
            }
          }
        }
      }
      __syncwarp();
    } else if (warp == 3) {
      // ------------------------------------------------------------------ TMA producer B: Q of every item, V_g (two stages)
      if (lane == 0) {
        int g = 0, iq = 0;
        for (int idx = item_lo; idx < item_hi; ++idx) {
          const Item it = decode_item(p, idx);
          if (it.T == 0) continue;
          if (iq > 0) ptx::mbar_wait(q_empty, (uint32_t)((iq - 1) & 1));   // the previous Q tile has been copied to TMEM
          ptx::mbar_arrive_expect_tx(&q_full[iq % ISLOTS], QT * 128);
          ptx::tma_load_3d(smem + SM::OFF_Q, &tm_qkv, &q_full[iq % ISLOTS], it.head * D, it.trow + it.q0, it.tb);
          ptx::tma_load_3d(smem + SM::OFF_Q + 8192, &tm_qkv, &q_full[iq % ISLOTS], it.head * D, it.trow + it.q0 + 64, it.tb);
          for (int t = 0; t < it.T; ++t, ++g) {
            const int st = g & 1;
            ptx::mbar_wait(&b_empty[st], (uint32_t)(((g >> 1) & 1) ^ 1));
            ptx::mbar_arrive_expect_tx(&b_full[st], KT * 128);
            ptx::tma_load_3d(smem + SM::OFF_V + st * 8192, &tm_qkv, &b_full[st], 2 * p.H + it.head * D, it.trow + t * KT, it.tb);
          }
          ++iq;
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
      int g = 0, iq = 0, cur_pair = -1, n_sw = 0;
      for (int idx = item_lo; idx < item_hi; ++idx) {
        const Item it = decode_item(p, idx);
        if (it.T == 0) continue;
        if (MODE == 2) {
          const int pair = it.head * 32 + it.q0 / QT;   // (head, query tile): what the resident tables depend on
          if (pair != cur_pair) {
            cur_pair = pair;
            // every MMA issued so far read the old tables: tell the producer when they have all completed
            if (n_sw > 0 && ptx::elect_one()) ptx::mma_commit(tab_free);
            __syncwarp();
            ptx::mbar_wait(tab_full, (uint32_t)(n_sw & 1));
            ++n_sw;
          }
        }
        ptx::mbar_wait(tmq_full, (uint32_t)(iq & 1));   // this item's Q operand is in TMEM
        ptx::tc_fence_after();
        for (int t = 0; t < it.T; ++t, ++g) {
          const int st = g % AST;
          const uint64_t dK = ptx::smem_desc_sw128(ptx::smem_u32(smem + SM::OFF_K + st * 8192));
          // EK: tile t's window starts at local row 64 t (MODE 2) / at row 0 of the stage (MODE 0; later tiles hold only
          // the new block there).  EQ: window of tile t starts at local row 64 (Tmax - 1 - t) (MODE 2) / row 0 (MODE 0).
          const uint64_t dEKw = ptx::smem_desc_sw128(ptx::smem_u32(smem + SM::OFF_EK + (MODE == 2 ? 64 * t * 128 : st * SM::POS_BYTES)));
          const uint64_t dEKn = (MODE == 2) ? dEKw + (uint64_t)((128 * 128) >> 4) : dEKw;   // the new block: window rows 128..191
          const uint64_t dEQ = ptx::smem_desc_sw128(ptx::smem_u32(smem + SM::OFF_EQ + (MODE == 2 ? 64 * (Tmax - 1 - t) * 128 : st * SM::POS_BYTES)));
          ptx::mbar_wait(&a_full[st], (uint32_t)((g / AST) & 1));
          if (g > 0) ptx::mbar_wait(sc_free, (uint32_t)((g - 1) & 1));   // S and C accumulators drained by the group of tile g-1
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            // S = Q.K_t^T and C = Q.EK^T (A = Q from TMEM): 16 halves along K = 8 TMEM columns per step
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::mma_f16_ts(tmem + TM_S, tmem + TM_Q + 8 * k, dK + 2 * k, idesc_s, (uint32_t)(k != 0));
            // C ring: block j of tile t sits in slot (t + j) % 3; the first tile fills all three, tile t > 0 only block j = 2
            if (t == 0) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::mma_f16_ts(tmem + TM_C, tmem + TM_Q + 8 * k, dEKw + 2 * k, idesc_c, (uint32_t)(k != 0));
            } else {
              const uint32_t dC = tmem + TM_C + (uint32_t)(64 * ((t + 2) % 3));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::mma_f16_ts(dC, tmem + TM_Q + 8 * k, dEKn + 2 * k, idesc_c64, (uint32_t)(k != 0));
            }
            ptx::mma_commit(&sc_full[g % NWG]);
            if (t == it.T - 1) ptx::mma_commit(&tmq_free[iq % ISLOTS]);   // the Q operand may be replaced
          }
          __syncwarp();
          if (g > 0) ptx::mbar_wait(g_free, (uint32_t)((g - 1) & 1));
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            // the three row-shifted copies of G = EQr_window . K_t^T; 32 table / key rows = 256 in a descriptor
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
            ptx::mma_commit(&g_full[g % NWG]);
          }
          __syncwarp();
        }
        ++iq;
      }
    } else if (warp == 2) {
      // ------------------------------------------------------------------ MMA issuer B: O (+)= P . V_g
      constexpr uint32_t idesc_pv = ptx::idesc_f16(128, 64, 0, 1);   // B (= V) is MN-major
      int g = 0, iq = 0;
      for (int idx = item_lo; idx < item_hi; ++idx) {
        const Item it = decode_item(p, idx);
        if (it.T == 0) continue;
        for (int t = 0; t < it.T; ++t, ++g) {
          const int st = g & 1;
          const uint64_t dV = ptx::smem_desc_sw128(ptx::smem_u32(smem + SM::OFF_V + st * 8192));
          ptx::mbar_wait(&b_full[st], (uint32_t)((g >> 1) & 1));
          ptx::mbar_wait(p_full, (uint32_t)(g & 1));
          if (t == 0 && iq > 0) ptx::mbar_wait(o_free, (uint32_t)((iq - 1) & 1));   // the previous item's O has been read out
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)   // V is MN-major: 16 keys further = +2048 bytes = +128 in the descriptor
              ptx::mma_f16_ts(tmem + TM_PV, tmem + TM_P + 8 * k, dV + 128 * k, idesc_pv, (uint32_t)(k != 0 || t > 0));
            ptx::mma_commit(&b_empty[st]);
            ptx::mma_commit(&pv_full[g % NWG]);
          }
          __syncwarp();
        }
        ++iq;
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax warpgroups
    setmaxnreg_inc<SMX_REGS>();
    const int wg = (warp - 4) >> 2;   // owns tiles g with g % NWG == wg
    const int qd = warp & 3;          // TMEM lane quarter
    const int i = qd * 32 + lane;     // row in the query tile
    const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
    const float sc = p.scale_log2;

    // tile-independent shift controls (both 32-key halves of a tile use the same ones: their first key is 0 mod 32)
    const int sh = 31 - lane;                  // c2p: element shift inside the window
    const bool sh16 = sh & 16, sh8 = sh & 8, sh4 = sh & 4, sh2 = sh & 2;
    const uint32_t prmt_sel = (sh & 1) ? 0x5432u : 0x3210u;
    const int rot0 = lane + 31;                // p2c: source lane of key column jj (of a half) is (rot0 - jj) & 31
    const uint4* hm_row = reinterpret_cast<const uint4*>(smem + SM::OFF_HM + lane * HM_PITCH);   // this lane's copy-select masks

    int g0 = 0, iq = 0;
    for (int idx = item_lo; idx < item_hi; ++idx) {
      const Item it = decode_item(p, idx);
      if (it.T == 0) {
        // padded queries only: their outputs are never read by valid rows
        for (int e = threadIdx.x - 128; e < QT * 8; e += 128 * NWG) {
          const int r = it.q0 + (e >> 3);
          if (r < it.rows)
            *reinterpret_cast<uint4*>(p.ctx + ((int64_t)it.grow + r) * p.H + it.head * D + (e & 7) * 8) = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      const int islot = iq % ISLOTS;
      const uint32_t ipar = (uint32_t)((iq / ISLOTS) & 1);
      float* mrow_i = mrow + islot * QT;
      float* lsum_i = lsum + islot * (NWG * 2 * QT);

      if (wg == g0 % NWG) {
        // ---- this group owns the item's first tile: Q tile -> TMEM (its row, 8 16-byte chunks of the swizzled 128-byte row)
        if (iq > 0) {
          const int ps = (iq - 1) % ISLOTS;
          ptx::mbar_wait(&tmq_free[ps], (uint32_t)(((iq - 1) / ISLOTS) & 1));   // the previous item's S | C MMAs have read TM_Q
        }
        ptx::mbar_wait(&q_full[islot], ipar);
        ptx::tc_fence_after();
        const uint8_t* qrow = smem + SM::OFF_Q + (i >> 3) * 1024 + (i & 7) * 128;
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
        if (lane == 0) {
          ptx::mbar_arrive(tmq_full);
          ptx::mbar_arrive(q_empty);
        }
      }

      float l_part = 0.f;                  // this group's partial row sum of the item, relative to m_seen
      float m_seen = -CUDART_INF_F;        // sticky maximum this group used for its latest tile of the item
      const int t_first = (wg - g0 % NWG + NWG) % NWG;
      for (int t = t_first; t < it.T; t += NWG) {
        const int g = g0 + t;
        const int slot = g % NWG;          // == wg
        const uint32_t par = (uint32_t)((g / NWG) & 1);
        const int k0 = t * KT;
        float s[KT];
        // key-validity words of this tile (warp-uniform broadcast loads; consumed after the skew stages)
        const int words = (it.rows + 31) >> 5;
        const uint32_t kb0 = __ldg(p.mask_bits + it.mword + (k0 >> 5));
        const uint32_t kb1 = ((k0 >> 5) + 1 < words) ? __ldg(p.mask_bits + it.mword + (k0 >> 5) + 1) : 0u;

        // ---- drain S and this warp's two 64-column windows of C in one go, then hand the accumulators back: the next
        //      tile's S | C MMA waits for exactly this
        ptx::mbar_wait(&sc_full[slot], par);
        ptx::tc_fence_after();
        uint32_t w0[32], w1[32];
        {
          // window of (warp quarter qd, half hf) starts at window column 96 - 32 qd + 32 hf; window block j is ring slot
          // (t + j) % 3, i.e. window column c is ring column (64 (t % 3) + c) % 192: three 32-column pieces
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
        // both halves are fetched and reduced to the selected copy FIRST (16 registers per half), so the G accumulators go
        // back to the tensor core after two load round trips instead of after half of the rotations (228.5 -> 222.5 us).
        // Tried on top and rejected: issuing S|C one tile ahead of G (340 us: G(g) is delayed, its K stage is released
        // late and the TMA latency of K(g+2) is exposed — shared memory has no room for a third K stage), and issuing
        // S|C / G in readiness order with non-blocking mbarrier probes (291 us: all three groups run concurrently and the
        // group that holds G finishes later)
        uint32_t gv[2][16];
        auto p2c_fetch = [&](auto half_tag) {
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
              gv[HF][m] = (hi[m] & hmv[mm]) | (lo[m] & ~hmv[mm]);
            }
          }
        };
        p2c_fetch(std::integral_constant<int, 0>{});
        p2c_fetch(std::integral_constant<int, 1>{});
#pragma unroll
        for (int HF = 0; HF < 2; ++HF) {
#pragma unroll
          for (int m = 0; m < 16; ++m) {
            const uint32_t v = gv[HF][m];
            const uint32_t x0 = __shfl_sync(0xffffffffu, v, rot0 - 2 * m);        // low half = key 2m
            const uint32_t x1 = __shfl_sync(0xffffffffu, v, rot0 - 2 * m - 1);    // high half = key 2m + 1
            ptx::add_f16_lo_to_f32(s[32 * HF + 2 * m], x0);
            ptx::add_f16_hi_to_f32(s[32 * HF + 2 * m + 1], x1);
          }
        }

        // ---- key validity, local row maximum (four independent chains of 3-input maxima)
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

        // ---- chain the sticky maximum inside the item: m(t) from m(t-1) of the previous tile's group and this row maximum
        float m_prev = -CUDART_INF_F;
        if (t > 0) {
          const int pg = (g - 1) % NWG;
          ptx::mbar_wait(&m_bar[pg * 4 + qd], (uint32_t)(((g - 1) / NWG) & 1));
          m_prev = mrow_i[i];
        }
        // raised only when the row maximum grew by more than 2^8 (P <= 2^8 keeps fp16 headroom); false when both are -inf
        // (NaN), true for the first finite maximum
        const bool raise = (mloc - m_prev) * sc > 8.0f;
        const float m_cur = raise ? mloc : m_prev;
        mrow_i[i] = m_cur;
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&m_bar[wg * 4 + qd]);
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
          const float e1 = (POLY && (jj & 3) == 2) ? exp2_poly(x1) : ptx::ex2(x1);
          ps[jj & 3] += e0;
          ps[(jj + 1) & 3] += e1;
          pk[jj >> 1] = ptx::pack_f16(e0, e1);
        }
        l_part += (ps[0] + ps[1]) + (ps[2] + ps[3]);

        // ---- the P buffer is free once the PV of tile g-1 (possibly of the previous item) has completed; inside the item
        //      that also makes O stable for the rare rescale by the raising group
        if (g > 0) {
          ptx::mbar_wait(&pv_full[(g - 1) % NWG], (uint32_t)(((g - 1) / NWG) & 1));
          if (t > 0 && __any_sync(0xffffffffu, raise)) {
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
      }

      // ---- item merge: every group publishes (partial row sum, the maximum it is relative to); the group of the last tile
      //      — whose m_seen is the final maximum — rescales and adds them, normalises O and writes the context rows
      const int last_wg = (g0 + it.T - 1) % NWG;
      lsum_i[(wg * 2 + 0) * QT + i] = l_part;
      lsum_i[(wg * 2 + 1) * QT + i] = m_seen;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&l_bar[islot]);
      if (wg == last_wg) {
        ptx::mbar_wait(&l_bar[islot], ipar);
        // summed in the order of the tile residues t % 3 = 0, 1, 2 (group (g0 + r) % 3), NOT in group order: a row's
        // result must not depend on where its item falls in the CTA's tile stream (batch-composition invariance, bit for bit)
        float l_tot = 0.f;
#pragma unroll
        for (int r = 0; r < NWG; ++r) {
          const int k = (g0 + r) % NWG;
          const float lk = lsum_i[(k * 2 + 0) * QT + i], mk = lsum_i[(k * 2 + 1) * QT + i];
          l_tot += (mk == m_seen) ? lk : ((mk == -CUDART_INF_F) ? 0.f : lk * ptx::ex2((mk - m_seen) * sc));
        }
        const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
        const int gl = g0 + it.T - 1;
        ptx::mbar_wait(&pv_full[gl % NWG], (uint32_t)((gl / NWG) & 1));
        ptx::tc_fence_after();
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_x32(t_lane + TM_PV, r0);
        ptx::tmem_ld_x32(t_lane + TM_PV + 32, r1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(o_free);   // the next item's first PV may overwrite O
        const int row = it.q0 + i;
        if (row < it.rows) {
          __half* dst = p.ctx + ((int64_t)it.grow + row) * p.H + it.head * D;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 o4;
            o4.x = ptx::pack_f16(__uint_as_float(r0[8 * v + 0]) * inv, __uint_as_float(r0[8 * v + 1]) * inv);
            o4.y = ptx::pack_f16(__uint_as_float(r0[8 * v + 2]) * inv, __uint_as_float(r0[8 * v + 3]) * inv);
            o4.z = ptx::pack_f16(__uint_as_float(r0[8 * v + 4]) * inv, __uint_as_float(r0[8 * v + 5]) * inv);
            o4.w = ptx::pack_f16(__uint_as_float(r0[8 * v + 6]) * inv, __uint_as_float(r0[8 * v + 7]) * inv);
            reinterpret_cast<uint4*>(dst)[v] = o4;
          }
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 o4;
            o4.x = ptx::pack_f16(__uint_as_float(r1[8 * v + 0]) * inv, __uint_as_float(r1[8 * v + 1]) * inv);
            o4.y = ptx::pack_f16(__uint_as_float(r1[8 * v + 2]) * inv, __uint_as_float(r1[8 * v + 3]) * inv);
            o4.z = ptx::pack_f16(__uint_as_float(r1[8 * v + 4]) * inv, __uint_as_float(r1[8 * v + 5]) * inv);
            o4.w = ptx::pack_f16(__uint_as_float(r1[8 * v + 6]) * inv, __uint_as_float(r1[8 * v + 7]) * inv);
            reinterpret_cast<uint4*>(dst + 32)[v] = o4;
          }
        }
      }
      g0 += it.T;
      ++iq;
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

static cudaError_t launch_persist(const void* exp_k, const void* exp_qr, int64_t ld_exp, PersistParams p,
                                  uint64_t tm_rows, uint64_t tm_batch, int num_sms, cudaStream_t stream) {
  const int H = p.H, S = p.S;
  // qkv viewed as [batch][rows][3H]; box 64 cols x 64 rows
  uint64_t dq[3] = {(uint64_t)(3 * H), tm_rows, tm_batch};
  uint64_t sq[2] = {(uint64_t)(3 * H) * 2, tm_rows * 3 * H * 2};
  uint32_t bq[3] = {64, 64, 1};
  // expanded tables are [EXP_ROWS][ld_exp] row-major (head h = columns h*64..): dims (d, row, head)
  uint64_t dp[3] = {64, (uint64_t)EXP_ROWS, (uint64_t)p.heads};
  uint64_t sp[2] = {(uint64_t)ld_exp * 2, 128};
  uint32_t bp[3] = {64, 64, 1};
  CUtensorMap tm_qkv = make_tmap_16b(p.qkv_base, 3, dq, sq, bq);
  CUtensorMap tm_ek = make_tmap_16b(exp_k, 3, dp, sp, bp);
  CUtensorMap tm_eq = make_tmap_16b(exp_qr, 3, dp, sp, bp);
  p.scale_log2 = 1.4426950408889634f / sqrtf(3.0f * D);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attention_persist_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<2>::BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_persist_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<2>::BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_persist_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<0>::BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_persist_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<0>::BYTES);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  const int grid = p.n_items < num_sms ? p.n_items : num_sms;
  static const int force_mode = [] { const char* e = getenv("GLC_ATTN_MODE"); return e ? atoi(e) : -1; }();   // developer switch
  // Every 4th exponential goes through the FMA-pipe polynomial (exp2_poly) by default.  GLC_ATTN_POLY=0 puts them all on the
  // MUFU unit: 112 fewer instructions per tile, 218.8 instead of 229.7 us per launch in isolation, but no change of the
  // power-capped step time (8.43 ms either way), and the logits move by their noise level — which on the deep random
  // fixtures sits AT the 2e-2 bar (S=1024 / 100 labels: 1.58e-2 with the polynomial, 2.04e-2 without; C2: 1.74e-2 / 1.67e-2;
  // profiles/r2h_precision_probe.txt).  The default is the variant the whole parity suite was validated with.
  static const bool poly = [] { const char* e = getenv("GLC_ATTN_POLY"); return !(e && e[0] == '0'); }();   // developer switch
  if (force_mode != 0 && (S + KT - 1) / KT <= TMAX_RES) {
    if (!poly) return launch_pdl(attention_persist_kernel<2, false>, dim3(grid), dim3(PTHREADS), Smem<2>::BYTES, stream, tm_qkv, tm_ek, tm_eq, p);
    return launch_pdl(attention_persist_kernel<2, true>, dim3(grid), dim3(PTHREADS), Smem<2>::BYTES, stream, tm_qkv, tm_ek, tm_eq, p);
  }
  // (both table modes use the same exponential: a text's logits must not depend on the padded length of its launch)
  if (!poly) return launch_pdl(attention_persist_kernel<0, false>, dim3(grid), dim3(PTHREADS), Smem<0>::BYTES, stream, tm_qkv, tm_ek, tm_eq, p);
  return launch_pdl(attention_persist_kernel<0, true>, dim3(grid), dim3(PTHREADS), Smem<0>::BYTES, stream, tm_qkv, tm_ek, tm_eq, p);
}

cudaError_t attention_persist(const void* qkv, const void* exp_k, const void* exp_qr, int64_t ld_exp,
                              const uint32_t* mask_bits, const int32_t* kv_len, void* ctx, int B, int S, int heads,
                              int num_sms, cudaStream_t stream) {
  if (B <= 0 || S <= 0) return cudaSuccess;
  if (S > 2048) return cudaErrorInvalidValue;   // reference MAX_LENGTH (include/configs.h:5)
  PersistParams p{};
  p.qkv_base = qkv;
  p.mask_bits = mask_bits;
  p.kv_len = kv_len;
  p.ctx = (__half*)ctx;
  p.B = B; p.S = S; p.heads = heads; p.H = heads * D;
  p.nq = (S + QT - 1) / QT;
  p.n_items = heads * p.nq * B;
  return launch_persist(exp_k, exp_qr, ld_exp, p, (uint64_t)S, (uint64_t)B, num_sms, stream);
}

cudaError_t attention_persist_packed(const void* qkv, const void* exp_k, const void* exp_qr, int64_t ld_exp,
                                     const uint32_t* row_bits, const int32_t* kv_len, const int32_t* text_row,
                                     const int32_t* tile_info, void* ctx, int B, int rows, int max_text_rows, int n_tiles,
                                     int heads, int num_sms, cudaStream_t stream) {
  if (B <= 0 || rows <= 0 || n_tiles <= 0) return cudaSuccess;
  if (max_text_rows > 2048 || (rows % QT) || (max_text_rows % QT) || n_tiles != rows / QT || B >= (1 << 24)) return cudaErrorInvalidValue;
  PersistParams p{};
  p.qkv_base = qkv;
  p.mask_bits = row_bits;
  p.kv_len = kv_len;
  p.ctx = (__half*)ctx;
  p.B = B; p.S = max_text_rows; p.heads = heads; p.H = heads * D;
  p.nq = max_text_rows / QT;
  p.tile_info = tile_info;
  p.text_row = text_row;
  p.n_tiles = n_tiles;
  p.n_items = heads * n_tiles;
  return launch_persist(exp_k, exp_qr, ld_exp, p, (uint64_t)rows, 1, num_sms, stream);
}

}  // namespace glc
