// K3 (second generation) — fused DeBERTa disentangled attention for sm_100a, head dim 64, with the
// relative-position biases added by the tensor core instead of gathered by the softmax threads.
//
//   ctx[b,i,h,:] = softmax_j( (Q_i.K_j + Q_i.posK[idx(i-j)] + K_j.posQ[idx(i-j)]) / sqrt(3d) + mask_j ) . V_j
//
// (transformers DisentangledSelfAttention, T:229-345; idx(delta) = clamp(bucket(delta)+span, 0, 2*span-1),
// SURVEY.md App. A.6.)  The first-generation kernel (attention.cu) stages per-tile slices of
// Q.posK^T / posQ.K^T in shared memory and lets every softmax thread gather two fp16 values per score:
// the in-kernel trace shows those gathers (bank-conflicted 2-byte LDS) are 45 % of a tile's time.
//
// Here the position tables are expanded once at load to one row per DELTA (expand_pos_table:
// E[rho] = pos[idx(2047 - rho)], rho in [0,4096)), which makes both biases Toeplitz in (i, j):
//     c2p[i,j] = QE[i, c]   QE = Q_tile . EK_slice^T      (128 x 192)
//     p2c[i,j] = EK[c, j]   EK = EQ_slice . K_tile^T      (192 x 64)        c = 127 - i + j (tile-local)
// for the 192-row slice rho0 .. rho0+191 of a (128-query, 64-key) tile.  The threads that drain those
// accumulators from TMEM write each value exactly once, already skewed, into two operand tiles
//     Y1[i][j] = c2p (K-major A operand)          X[j][i] = p2c (MN-major A operand)
// (16 two-byte stores each per thread per tile, at addresses that do not depend on the tile, so they are
// computed once) and the tensor core adds them:   S = Q.K^T + Y1.I + X^T.I   (I = 64x64 identity).
// The softmax threads read finished scores from TMEM; there is no LUT, no gather, no bias arithmetic.
// Consecutive key tiles use slices shifted by 64 rows, so the slices live in rings of 64-row blocks and
// each tile fetches only one new block per table (L2 -> SM traffic per tile: K, V and two 8 KB blocks).
//
// One CTA (576 threads) per (128-query tile, head, batch row); 64-key tiles t = 0..T-1:
//   warp 0   TMA producer (K ring x3, V ring x2, position-block rings x4 (+ mirror slot, see OFF_PQ))
//   warp 1   tcgen05 issuer:  A(t): QE, EK     B(t): S = QK^T + Y1.I + X^T.I     C(t): PV = P.V_t
//            issue order  A(0) | A(1) B(0) | A(t+2) B(t+1) C(t) | ... ; Q and P are TMEM A-operands; Y1 / X are
//            double buffered so that stage(t+1) never waits for B(t)
//   warps 2-17  thread = (query row i / TMEM lane, 16-key group g): stage(t+1) -> softmax(t) -> P
// TMEM (512 columns): S 64 | Q 32 | P 32 | QE 192 | EK 2x64 | PV 64.
// Synchronisation is mbarrier based; every wait is bounded (ptx::mbar_wait traps on timeout).
#include <cuda_fp16.h>
#include <math_constants.h>

#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "kernels.h"
#include "model_weights.h"
#include "ptx.cuh"
#include "tma_desc.h"

namespace glc {
namespace {

constexpr int QT = 128;            // queries per CTA
constexpr int KT = 64;             // keys per tile
constexpr int D = 64;              // head dim
constexpr int G = 4;               // key groups per tile: softmax thread = (query row, 16-key group)
constexpr int E = KT / G;          // scores per softmax thread per tile
constexpr int SM_WARPS = 4 * G;
constexpr int ATT_THREADS = 64 + 32 * SM_WARPS;
constexpr int EXP_CENTER = 2047;   // expanded tables: row rho = EXP_CENTER - delta
constexpr int EXP_ROWS = 4096;

// shared memory map (bytes, from a 1024-aligned base).  The two Y1/X buffer pairs come first: store addresses into
// pair 0 are kept as 16-bit absolute shared addresses, pair 1 is the same plus an immediate.
constexpr int OFF_Y1 = 0;                    // 128 x 128 B   c2p tile, K-major A operand           (pair b: + b * PAIR_BYTES)
constexpr int OFF_X = 16384;                 // 2 x 64 x 128 B   p2c tile, MN-major A operand (two 64-query atoms)
constexpr int PAIR_BYTES = 32768;
constexpr int OFF_I = 65536;                 // 64 x 128 B    identity, K-major B operand
constexpr int OFF_K = OFF_I + 8192;          // 3 x 8 KB
constexpr int OFF_V = OFF_K + 3 * 8192;      // 2 x 8 KB
constexpr int OFF_PK = OFF_V + 2 * 8192;     // 4 x 8 KB: 64-row blocks of the expanded posK table (B operands of QE)
constexpr int OFF_PQ = OFF_PK + 4 * 8192;    // 5 x 8 KB: blocks of the expanded posQ table (A operands of EK, M = 128 rows =
                                             // two consecutive blocks); slot 4 mirrors slot 0 so (3,0) is contiguous too
constexpr int OFF_Q = OFF_PQ + 5 * 8192;     // 128 x 128 B
constexpr int OFF_XMAX = OFF_Q + 16384;      // 2 x G x 128 floats (row-max exchange, double buffered by tile parity)
constexpr int OFF_MASK = OFF_XMAX + 2 * G * QT * 4;   // uint32[72]: key-validity words of this batch row
constexpr int OFF_BAR = OFF_MASK + 72 * 4;
constexpr int ATT_SMEM = OFF_BAR + 256 + 1024;
static_assert(ATT_SMEM <= 227 * 1024, "attention smem budget");
static_assert(OFF_BAR % 8 == 0, "barrier alignment");

// TMEM columns
constexpr uint32_t TM_S = 0;       // 64
constexpr uint32_t TM_Q = 64;      // 32: Q tile as fp16 pairs (A operand of the S and QE MMAs)
constexpr uint32_t TM_P = 96;      // 32: P tile as fp16 pairs (A operand of the PV MMA)
constexpr uint32_t TM_QE = 128;    // 192
constexpr uint32_t TM_EK = 320;    // 2 x 64 (slice rows 0..127 | 128..191 in lanes 0..63)
constexpr uint32_t TM_PV = 448;    // 64

struct AttnParams {
  const uint32_t* mask_bits; // [B][ceil(S/32)]
  const int32_t* kv_len;     // [B]
  __half* ctx;               // [B*S, H]
  int B, S, heads, H;
  float scale_log2;          // log2(e) / sqrt(3*d)
  int flags;                 // TRACE instantiation only: developer what-if switches (results become wrong)
  long long* trace;          // TRACE instantiation only: [2 roles][TMAX][8] clock64 stamps of CTA (1,0,0)
};

constexpr int TMAX = 2048 / KT;
#define GLC_TRACE(role, tile, slot)                                                                                   \
  do {                                                                                                                \
    if (TRACE && p.trace && (threadIdx.x & 31) == 0 && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0)         \
      p.trace[((role) * TMAX + (tile)) * 8 + (slot)] = clock64();                                                     \
  } while (0)
#define GLC_FLAG(bit) (TRACE && (p.flags & (bit)))

__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// two 2-byte shared stores: low half of `vals` to the address in the low half of `addrs`, high to high
template <int OFFSET>
__device__ __forceinline__ void sts_u16_pair(uint32_t addrs, uint32_t vals) {
  asm volatile(
      "{\n\t.reg .b16 vl, vh, al, ah;\n\t.reg .b32 a0, a1;\n\t"
      "mov.b32 {vl, vh}, %1;\n\tmov.b32 {al, ah}, %0;\n\t"
      "cvt.u32.u16 a0, al;\n\tcvt.u32.u16 a1, ah;\n\t"
      "st.shared.b16 [a0+%2], vl;\n\tst.shared.b16 [a1+%2], vh;\n\t}" ::"r"(addrs),
      "r"(vals), "n"(OFFSET)
      : "memory");
}

template <bool TRACE>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_toeplitz_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_ek,
                          const __grid_constant__ CUtensorMap tm_eq, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* qt_full = bars + 1;      // Q tile copied into TMEM
  uint64_t* k_full = bars + 2;       // [3]
  uint64_t* k_empty = bars + 5;      // [3]  freed by B(t)
  uint64_t* v_full = bars + 8;       // [2]
  uint64_t* v_empty = bars + 10;     // [2]  freed by C(t)
  uint64_t* blk_full = bars + 12;    // [4]  position block beta -> slot beta & 3
  uint64_t* blk_empty = bars + 16;   // [4]  freed by A(beta) (last tile that reads block beta)
  uint64_t* bias_full = bars + 20;   // QE / EK accumulators of tile t ready
  uint64_t* stage_done = bars + 21;  // Y1 / X (pair t & 1) of tile t written, QE / EK drained
  uint64_t* s_full = bars + 22;      // finished scores of tile t ready (also: Y1 / X pair t & 1 free again)
  uint64_t* s_free = bars + 23;      // softmax warps hold S of tile t in registers
  uint64_t* p_full = bars + 24;      // P tile written
  uint64_t* pv_full = bars + 25;     // PV accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  uint32_t* kmask = reinterpret_cast<uint32_t*>(smem + OFF_MASK);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int kvlen = p.kv_len[b];
  const int T = (kvlen + KT - 1) / KT;   // key tiles with at least one valid key
  const int NBLK = T + 2;                // position blocks 0 .. T+1 (tile t reads blocks t, t+1, t+2)
  const int rho_base = EXP_CENTER - (q0 + QT - 1);   // first row of block 0 (delta = q0 + 127)

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
    if (ptx::smem_u32(smem) + PAIR_BYTES > 65536u) __trap();   // 16-bit store addresses (see sts_u16_pair)
    ptx::prefetch_tensormap(&tm_qkv);
    ptx::prefetch_tensormap(&tm_ek);
    ptx::prefetch_tensormap(&tm_eq);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(qt_full, SM_WARPS);
    for (int s = 0; s < 3; ++s) { ptx::mbar_init(&k_full[s], 1); ptx::mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&v_full[s], 1); ptx::mbar_init(&v_empty[s], 1); }
    for (int s = 0; s < 4; ++s) { ptx::mbar_init(&blk_full[s], 1); ptx::mbar_init(&blk_empty[s], 1); }
    ptx::mbar_init(bias_full, 1);
    ptx::mbar_init(stage_done, SM_WARPS);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_free, SM_WARPS);
    ptx::mbar_init(p_full, SM_WARPS);
    ptx::mbar_init(pv_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);

  // identity operand (64 x 64 fp16, K-major, 128-byte swizzle) and the key-validity words
  {
    uint4* iz = reinterpret_cast<uint4*>(smem + OFF_I);
    for (int e = threadIdx.x; e < 8192 / 16; e += ATT_THREADS) {
      // 16-byte chunk e: row n = e >> 3, physical chunk e & 7 holds logical chunk (e & 7) ^ (n & 7)
      const int n = e >> 3, lc = (e & 7) ^ (n & 7);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (lc == (n >> 3)) {
        const uint32_t one = (n & 1) ? 0x3C000000u : 0x00003C00u;   // fp16 1.0 at element n & 7 of the chunk
        const int w = (n & 7) >> 1;
        v.x = w == 0 ? one : 0u; v.y = w == 1 ? one : 0u; v.z = w == 2 ? one : 0u; v.w = w == 3 ? one : 0u;
      }
      iz[e] = v;
    }
    const int words = (p.S + 31) >> 5;
    const int e = threadIdx.x - 64;
    if (e >= 0 && e < 72) kmask[e] = (e < words) ? __ldg(p.mask_bits + (int64_t)b * words + e) : 0u;
  }
  ptx::fence_proxy_async();   // the identity is read by the tensor core (async proxy)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      auto load_k = [&](int x) {
        const int st = x % 3;
        ptx::mbar_wait(&k_empty[st], ((x / 3) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&k_full[st], KT * 128);
        ptx::tma_load_3d(smem + OFF_K + st * 8192, &tm_qkv, &k_full[st], p.H + head * D, x * KT, b);
      };
      auto load_v = [&](int x) {
        const int st = x & 1;
        ptx::mbar_wait(&v_empty[st], ((x >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&v_full[st], KT * 128);
        ptx::tma_load_3d(smem + OFF_V + st * 8192, &tm_qkv, &v_full[st], 2 * p.H + head * D, x * KT, b);
      };
      auto load_blk = [&](int beta) {
        const int s = beta & 3;
        const int rho = rho_base + 64 * beta;
        ptx::mbar_wait(&blk_empty[s], ((beta >> 2) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&blk_full[s], (uint32_t)((s == 0 ? 3 : 2) * 8192));
        ptx::tma_load_3d(smem + OFF_PK + s * 8192, &tm_ek, &blk_full[s], 0, rho, head);
        ptx::tma_load_3d(smem + OFF_PQ + s * 8192, &tm_eq, &blk_full[s], 0, rho, head);
        if (s == 0) ptx::tma_load_3d(smem + OFF_PQ + 4 * 8192, &tm_eq, &blk_full[s], 0, rho, head);
      };
      ptx::mbar_arrive_expect_tx(q_full, QT * 128);
      ptx::tma_load_3d(smem + OFF_Q, &tm_qkv, q_full, head * D, q0, b);
      ptx::tma_load_3d(smem + OFF_Q + 8192, &tm_qkv, q_full, head * D, q0 + 64, b);
      load_k(0);
      load_blk(0); load_blk(1); load_blk(2);
      if (T > 1) load_k(1);
      if (NBLK > 3) load_blk(3);
      load_v(0);
      if (T > 2) load_k(2);
      if (T > 1) load_v(1);
      // then in the order the tensor core frees buffers:  A(0) A(1) B(0) | A(t+2) B(t+1) C(t) | ...
      if (NBLK > 4) load_blk(4);          // after A(0)
      if (NBLK > 5) load_blk(5);          // after A(1)
      if (T > 3) load_k(3);               // after B(0)
      for (int t = 0; t < T; ++t) {
        if (t + 6 < NBLK) load_blk(t + 6);   // after A(t+2)
        if (t + 4 < T) load_k(t + 4);        // after B(t+1)
        if (t + 2 < T) load_v(t + 2);        // after C(t)
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // All 32 lanes walk the warp-uniform loop and wait on the barriers; one elected lane issues.
    constexpr uint32_t idesc_n64 = ptx::idesc_f16(128, 64);
    constexpr uint32_t idesc_xadd = ptx::idesc_f16(128, 64, 1, 0);   // A (= X) is MN-major
    constexpr uint32_t idesc_pv = ptx::idesc_f16(128, 64, 0, 1);     // B (= V) is MN-major
    const uint64_t dY1 = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_Y1));
    const uint64_t dX = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_X), 1024, 8192);   // LBO: next 64-query atom
    constexpr uint64_t PAIR_DESC = PAIR_BYTES >> 4;                                        // pair 1: + 32 KB
    const uint64_t dI = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_I));
    const uint32_t sK = ptx::smem_u32(smem + OFF_K), sV = ptx::smem_u32(smem + OFF_V);
    const uint32_t sPK = ptx::smem_u32(smem + OFF_PK), sPQ = ptx::smem_u32(smem + OFF_PQ);

    // A(x): bias accumulators of tile x
    auto issue_A = [&](int x) {
      ptx::mbar_wait(&k_full[x % 3], (x / 3) & 1);
      if (x == 0) {
        ptx::mbar_wait(&blk_full[0], 0);
        ptx::mbar_wait(&blk_full[1], 0);
      }
      ptx::mbar_wait(&blk_full[(x + 2) & 3], ((x + 2) >> 2) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t dK = ptx::smem_desc_sw128(sK + (x % 3) * 8192);
#pragma unroll
        for (int m = 0; m < (GLC_FLAG(16) ? 0 : 3); ++m) {
          const uint64_t dPK = ptx::smem_desc_sw128(sPK + ((x + m) & 3) * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // A = Q from TMEM: 16 halves along K = 8 columns per step
            ptx::mma_f16_ts(tmem + TM_QE + 64 * m, tmem + TM_Q + 8 * k, dPK + 2 * k, idesc_n64, (uint32_t)(k != 0));
        }
        const uint64_t dPQ0 = ptx::smem_desc_sw128(sPQ + (x & 3) * 8192);         // slice rows 0..127
        const uint64_t dPQ1 = ptx::smem_desc_sw128(sPQ + ((x + 2) & 3) * 8192);   // slice rows 128..191 (+64 unused)
        if (!GLC_FLAG(8)) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(tmem + TM_EK, dPQ0 + 2 * k, dK + 2 * k, idesc_n64, (uint32_t)(k != 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(tmem + TM_EK + 64, dPQ1 + 2 * k, dK + 2 * k, idesc_n64, (uint32_t)(k != 0));
        }
        ptx::mma_commit(bias_full);
        ptx::mma_commit(&blk_empty[x & 3]);
      }
      __syncwarp();
    };
    // C(x): PV of tile x
    auto issue_C = [&](int x) {
      ptx::mbar_wait(&v_full[x & 1], (x >> 1) & 1);
      ptx::mbar_wait(p_full, x & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t dV = ptx::smem_desc_sw128(sV + (x & 1) * 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // V is MN-major: 16 keys further = +2048 bytes = +128 in the descriptor
          ptx::mma_f16_ts(tmem + TM_PV, tmem + TM_P + 8 * k, dV + 128 * k, idesc_pv, (uint32_t)(k != 0));
        ptx::mma_commit(&v_empty[x & 1]);
        ptx::mma_commit(pv_full);
      }
      __syncwarp();
    };

    // B(t): S = Q.K_t^T + Y1.I + X^T.I
    auto issue_B = [&](int t) {
      GLC_TRACE(1, t, 1);
      if (t > 0) ptx::mbar_wait(s_free, (t - 1) & 1);
      ptx::tc_fence_after();
      GLC_TRACE(1, t, 2);
      if (ptx::elect_one()) {
        const uint64_t dK = ptx::smem_desc_sw128(sK + (t % 3) * 8192);
        const uint64_t pair = (uint64_t)(t & 1) * PAIR_DESC;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::mma_f16_ts(tmem + TM_S, tmem + TM_Q + 8 * k, dK + 2 * k, idesc_n64, (uint32_t)(k != 0));
        if (!GLC_FLAG(4)) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(tmem + TM_S, dY1 + pair + 2 * k, dI + 2 * k, idesc_n64, 1u);
        }
        if (!GLC_FLAG(2)) {
#pragma unroll
          for (int k = 0; k < 4; ++k)   // X is MN-major: 16 keys further = +2048 bytes
            ptx::mma_f16_ss(tmem + TM_S, dX + pair + 128 * k, dI + 2 * k, idesc_xadd, 1u);
        }
        ptx::mma_commit(s_full);
        ptx::mma_commit(&k_empty[t % 3]);
      }
      __syncwarp();
      GLC_TRACE(1, t, 3);
    };

    // issue order  A(0) | A(1) B(0) | A(t+2) B(t+1) C(t) ...: A(x+1) only needs the bias accumulators drained
    // (stage_done(x)), so it runs ahead of B(x) and the staging warps never wait for it behind a B
    ptx::mbar_wait(qt_full, 0);
    ptx::tc_fence_after();
    issue_A(0);
    ptx::mbar_wait(stage_done, 0);
    ptx::tc_fence_after();
    if (T > 1) issue_A(1);
    issue_B(0);
    for (int t = 0; t < T; ++t) {
      if (t + 1 < T) {
        GLC_TRACE(1, t + 1, 0);
        ptx::mbar_wait(stage_done, (t + 1) & 1);
        ptx::tc_fence_after();
        if (t + 2 < T) issue_A(t + 2);
        GLC_TRACE(1, t + 1, 4);
        issue_B(t + 1);
      }
      issue_C(t);
      GLC_TRACE(1, t, 5);
    }
  } else {
    // ------------------------------------------------------------------ staging + softmax warps
    const int sw = warp - 2;          // 0..SM_WARPS-1
    const int g = sw >> 2;            // key group of the tile: keys [E*g, E*g+E)
    const int qd = warp & 3;          // TMEM lane quarter
    const int i = qd * 32 + lane;     // row in the query tile == TMEM lane
    const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
    float* xmax = reinterpret_cast<float*>(smem + OFF_XMAX);

    // ---- tile-invariant store addresses (16-bit absolute shared addresses, two per register)
    // c2p: this thread drains QE columns cq0 + {8g+e, 64+8g+e, 32+8g+e}, e < 8 (cq0 = 96 - 32 qd is the first
    //      column rows 32qd..32qd+31 can index).  Row i uses exactly one of columns k, k+64 (k < 32) and all
    //      of 32..63: key j = column + lane - 31 (mod 64).
    // p2c: this thread drains EK row c = TMEM lane i (block 0) / 128 + i (block 1, lanes < 64), keys 16g..16g+15;
    //      for lanes < 64 exactly one of the two blocks maps into the tile: query = (127 - c + j) & 127.
    uint32_t ya[8], xa[8];
    {
      const uint32_t y_row = ptx::smem_u32(smem + OFF_Y1) + (uint32_t)i * 128u;
      const uint32_t y_sw = (uint32_t)(i & 7) << 3;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t j_sel = (uint32_t)(8 * g + e + lane - 31) & 63u;
        const uint32_t j_mid = (uint32_t)(32 + 8 * g + e + lane - 31);
        ya[e] = (y_row + 2u * (j_sel ^ y_sw)) | ((y_row + 2u * (j_mid ^ y_sw)) << 16);
      }
      const uint32_t x_base = ptx::smem_u32(smem + OFF_X);
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        uint32_t a2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t j = (uint32_t)(E * g + 2 * v + u);
          const uint32_t qi = (uint32_t)(127 - i + (int)j) & 127u;
          a2[u] = x_base + ((qi >> 6) << 13) + j * 128u + 2u * ((qi & 63u) ^ ((j & 7u) << 3));
        }
        xa[v] = a2[0] | (a2[1] << 16);
      }
    }
    const int c2p_thr = 31 - 8 * g - lane;   // selected column e comes from the low window iff e >= c2p_thr
    const int p2c_thr = i - E * g;            // (lanes < 64) key jj comes from block 0 iff jj <= p2c_thr
    const uint32_t cq0 = (uint32_t)(96 - 32 * qd);

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

    // drain QE / EK of tile x into the Y1 / X operand tiles
    auto stage = [&](int x, auto pair_tag) {
      constexpr int PO = decltype(pair_tag)::value * PAIR_BYTES;
      ptx::mbar_wait(bias_full, x & 1);
      ptx::tc_fence_after();
      if (sw == 0 && x > 0) GLC_TRACE(0, x - 1, 2);
      if (GLC_FLAG(1)) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(stage_done);
        return;
      }
      {
        uint32_t lo[8], hi[8], mid[8];
        tmem_ld_x8(t_lane + TM_QE + cq0 + (uint32_t)(8 * g), lo);
        tmem_ld_x8(t_lane + TM_QE + cq0 + (uint32_t)(64 + 8 * g), hi);
        tmem_ld_x8(t_lane + TM_QE + cq0 + (uint32_t)(32 + 8 * g), mid);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t sel = (e >= c2p_thr) ? lo[e] : hi[e];
          sts_u16_pair<PO>(ya[e], ptx::pack_f16(__uint_as_float(sel), __uint_as_float(mid[e])));
        }
      }
      {
        uint32_t a[16];
        ptx::tmem_ld_x16(t_lane + TM_EK + (uint32_t)(E * g), a);
        if (qd < 2) {   // warp-uniform
          uint32_t c[16];
          ptx::tmem_ld_x16(t_lane + TM_EK + (uint32_t)(64 + E * g), c);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) a[jj] = (jj <= p2c_thr) ? a[jj] : c[jj];
        } else {
          ptx::tmem_ld_wait();
        }
#pragma unroll
        for (int v = 0; v < 8; ++v)
          sts_u16_pair<PO>(xa[v], ptx::pack_f16(__uint_as_float(a[2 * v]), __uint_as_float(a[2 * v + 1])));
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();   // generic-proxy stores -> tensor-core (async proxy) reads
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(stage_done);
    };

    float m_run = -CUDART_INF_F, l_run = 0.f, alpha_prev = 1.f;
    float o[E];
#pragma unroll
    for (int k = 0; k < E; ++k) o[k] = 0.f;
    const float sc = p.scale_log2;

    using Pair0 = std::integral_constant<int, 0>;
    using Pair1 = std::integral_constant<int, 1>;
    stage(0, Pair0{});
    for (int t = 0; t < T; ++t) {
      if (sw == 0) GLC_TRACE(0, t, 0);
      // pair (t+1) & 1 was last read by B(t-1), whose completion (s_full(t-1)) this thread has already observed
      if (t + 1 < T) {
        if (t & 1) stage(t + 1, Pair0{}); else stage(t + 1, Pair1{});
      }
      if (sw == 0) GLC_TRACE(0, t, 1);
      ptx::mbar_wait(s_full, t & 1);   // scores of tile t finished
      if (sw == 0) GLC_TRACE(0, t, 3);
      ptx::tc_fence_after();
      float s[E];
      {
        uint32_t r[16];
        ptx::tmem_ld_x16(t_lane + TM_S + (uint32_t)(g * E), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < E; ++jj) s[jj] = __uint_as_float(r[jj]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s_free);

      const int kb = t * KT + g * E;
      const uint32_t kbits = (kmask[kb >> 5] >> (kb & 31)) & ((1u << E) - 1u);
      if (kbits != ((1u << E) - 1u)) {
#pragma unroll
        for (int jj = 0; jj < E; ++jj)
          if (!((kbits >> jj) & 1u)) s[jj] = -CUDART_INF_F;
      }
      float mloc = s[0];
#pragma unroll
      for (int jj = 1; jj < E; ++jj) mloc = fmaxf(mloc, s[jj]);
      // ---- row max shared between the key groups of a lane quarter (buffers alternate by tile parity: the
      //      quarter barrier of tile t+1 orders the reads of tile t before the writes of tile t+2)
      float* xm = xmax + (t & 1) * (G * QT);
      xm[g * QT + i] = mloc;
      if (sw == 0) GLC_TRACE(0, t, 4);
      ptx::named_bar_sync(2 + qd, 32 * G);
      if (sw == 0) GLC_TRACE(0, t, 5);
      float m_new = m_run;
#pragma unroll
      for (int gg = 0; gg < G; ++gg) m_new = fmaxf(m_new, xm[gg * QT + i]);
      const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
      const float alpha = ptx::ex2((m_run - m_use) * sc);
      const float neg_ms = -m_use * sc;
      float psum = 0.f;
#pragma unroll
      for (int jj = 0; jj < E; ++jj) {
        const float e = ptx::ex2(fmaf(s[jj], sc, neg_ms));
        s[jj] = e;
        psum += e;
      }
      l_run = l_run * alpha + psum;
      m_run = m_new;

      // ---- fold in PV of the previous tile (also guarantees the P buffer is free again)
      if (sw == 0) GLC_TRACE(0, t, 6);
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

      // ---- P tile -> TMEM: row i, fp16 pairs at columns 8g .. 8g+7
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
      if (sw == 0) GLC_TRACE(0, t, 7);
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
    float* xs = xmax + (T & 1) * (G * QT);   // the buffer tile T-1 did not use
    ptx::named_bar_sync(2 + qd, 32 * G);     // ... and the quarter is past its reads of tile T-2
    xs[g * QT + i] = l_run;
    ptx::named_bar_sync(2 + qd, 32 * G);
    float l_tot = 0.f;
#pragma unroll
    for (int gg = 0; gg < G; ++gg) l_tot += xs[gg * QT + i];
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

// dst[rho][0:cols) = src[idx[rho]][0:cols) for rho < rows-1 (idx < 0: zero row), 16 bytes per thread
__global__ void expand_rows_kernel(const __half* __restrict__ src, int64_t ld_src, const int32_t* __restrict__ idx,
                                   __half* __restrict__ dst, int64_t ld_dst, int rows, int cols) {
  const int chunks = cols >> 3;
  const int64_t n = (int64_t)rows * chunks;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / chunks), c = (int)(e % chunks);
    const int s = idx[r];
    uint4 v = make_uint4(0, 0, 0, 0);
    if (s >= 0) v = *reinterpret_cast<const uint4*>(src + (int64_t)s * ld_src + c * 8);
    *reinterpret_cast<uint4*>(dst + (int64_t)r * ld_dst + c * 8) = v;
  }
}

}  // namespace

int expanded_pos_rows() { return EXP_ROWS; }

void expanded_pos_index(int buckets, int max_pos, int32_t* out /* [EXP_ROWS] */) {
  // row rho holds delta = EXP_CENTER - rho; the last row (delta = -2048) is never indexed -> zero row
  std::vector<int32_t> rel((size_t)2 * (EXP_CENTER + 1) - 1);
  rel_index_table(EXP_CENTER + 1, buckets, max_pos, rel.data());   // rel[delta + EXP_CENTER]
  for (int rho = 0; rho < EXP_ROWS - 1; ++rho) out[rho] = rel[(size_t)(EXP_CENTER - rho) + EXP_CENTER];
  out[EXP_ROWS - 1] = -1;
}

cudaError_t expand_pos_table(const void* pos_f16, int64_t ld_src, const int32_t* d_exp_index, void* out_f16, int64_t ld_dst,
                             int cols, cudaStream_t stream) {
  if (cols % 8 != 0 || ld_src % 8 != 0 || ld_dst % 8 != 0) return cudaErrorInvalidValue;
  expand_rows_kernel<<<592, 256, 0, stream>>>((const __half*)pos_f16, ld_src, d_exp_index, (__half*)out_f16, ld_dst,
                                              EXP_ROWS, cols);
  return cudaGetLastError();
}

cudaError_t attention_toeplitz(const void* qkv, const void* exp_k, const void* exp_q, int64_t ld_exp,
                               const uint32_t* mask_bits, const int32_t* kv_len, void* ctx, int B, int S, int heads,
                               cudaStream_t stream) {
  if (B <= 0 || S <= 0) return cudaSuccess;
  if (S > 2048) return cudaErrorInvalidValue;   // reference MAX_LENGTH (include/configs.h:5)
  const int H = heads * D;
  // qkv viewed as [B][S][3H]; box 64 cols x 64 rows
  uint64_t dq[3] = {(uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
  uint64_t sq[2] = {(uint64_t)(3 * H) * 2, (uint64_t)S * 3 * H * 2};
  uint32_t bq[3] = {64, 64, 1};
  // expanded tables are [EXP_ROWS][ld_exp] row-major (head h = columns h*64..): dims (d, rho, head)
  uint64_t dp[3] = {64, (uint64_t)EXP_ROWS, (uint64_t)heads};
  uint64_t sp[2] = {(uint64_t)ld_exp * 2, 128};
  uint32_t bp[3] = {64, 64, 1};
  CUtensorMap tm_qkv = make_tmap_16b(qkv, 3, dq, sq, bq);
  CUtensorMap tm_ek = make_tmap_16b(exp_k, 3, dp, sp, bp);
  CUtensorMap tm_eq = make_tmap_16b(exp_q, 3, dp, sp, bp);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attention_toeplitz_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_toeplitz_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  AttnParams p;
  p.mask_bits = mask_bits;
  p.kv_len = kv_len;
  p.ctx = (__half*)ctx;
  p.B = B; p.S = S; p.heads = heads; p.H = H;
  p.scale_log2 = 1.4426950408889634f / sqrtf(3.0f * D);
  p.flags = 0;
  p.trace = nullptr;
  dim3 grid((S + QT - 1) / QT, heads, B);
  // developer aids: GLC_ATTN_FLAGS=<bits> what-if switches; GLC_ATTN_TRACE=<file> dumps per-tile clock64 stamps of
  // CTA (1,0,0) (synchronous).  Both select the TRACE instantiation; production launches never pay for them.
  const char* fl = getenv("GLC_ATTN_FLAGS");
  const char* tf = getenv("GLC_ATTN_TRACE");
  if (fl) p.flags = atoi(fl);
  if (tf) {
    const size_t n = 2 * TMAX * 8;
    if (cudaMalloc(&p.trace, n * sizeof(long long)) != cudaSuccess) return cudaGetLastError();
    cudaMemsetAsync(p.trace, 0, n * sizeof(long long), stream);
    attention_toeplitz_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm_qkv, tm_ek, tm_eq, p);
    cudaError_t e = cudaStreamSynchronize(stream);
    std::vector<long long> h(n);
    cudaMemcpy(h.data(), p.trace, n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    if (FILE* f = fopen(tf, "w")) {
      const long long t0 = h[1];
      for (int role = 0; role < 2; ++role)
        for (int t = 0; t < TMAX; ++t) {
          if (!h[(role * TMAX + t) * 8 + 1]) continue;
          fprintf(f, "%s t=%d", role ? "mma" : "smx", t);
          for (int k = 0; k < 8; ++k) fprintf(f, " %lld", h[(role * TMAX + t) * 8 + k] ? h[(role * TMAX + t) * 8 + k] - t0 : -1);
          fprintf(f, "\n");
        }
      fclose(f);
    }
    return e;
  }
  if (fl) {
    attention_toeplitz_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm_qkv, tm_ek, tm_eq, p);
    return cudaGetLastError();
  }
  attention_toeplitz_kernel<false><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm_qkv, tm_ek, tm_eq, p);
  return cudaGetLastError();
}

}  // namespace glc
