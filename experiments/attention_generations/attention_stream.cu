// K3 (fourth generation) — fused DeBERTa disentangled attention for sm_100a, head dim 64: the register-skew
// bias scheme of attention_shift.cu with the softmax reorganised as two INDEPENDENT key-half streams whose
// output accumulators stay in tensor memory.
//
//   ctx[b,i,h,:] = softmax_j( (Q_i.K_j + Q_i.posK[idx(i-j)] + K_j.posQ[idx(i-j)]) / sqrt(3d) + mask_j ) . V_j
//
// (transformers DisentangledSelfAttention, T:229-345: scores T:237-251, c2p T:313-324, p2c T:327-343, masked
// softmax T:256-259, context T:262-271; idx(delta) = clamp(bucket(delta)+span, 0, 2*span-1))
//
// Biases (see attention_shift.cu for the derivation): with the position tables expanded to one row per delta,
//     c2p[a,b] = C[a, 127 - a + b]     C = Q_tile . EK[rho0 .. rho0+191]^T          rho0   = 1920 - q0 + k0
//     p2c[a,b] = G[a + 63 - b, b]      G = EQr[sigma0 .. sigma0+191] . K_tile^T     sigma0 = 1984 + q0 - k0
// c2p is un-skewed by a 5-stage barrel shifter on the packed accumulator window of each thread, p2c by one lane
// rotation per key column over three row-shifted copies of G.
//
// What attention_shift.cu's in-kernel trace showed (2500 cycles per 128x64 tile): each softmax warp runs
// TMEM drain (64 B/clk per lane quarter) -> ALU-pipe skew -> MUFU exponentials -> PV fold strictly in sequence,
// the two warps of a scheduler in lock-step because they exchange the row maximum every tile.  Here
//   * stream x (x = 0, 1) owns keys [32x, 32x+32) of EVERY key tile and keeps its own running maximum m_x, sum l_x
//     and its own output accumulator O_x: no per-tile exchange between the two warps of a row, so one warp's
//     drain overlaps the other's arithmetic;
//   * O_x is accumulated by the tensor core in TMEM over all key tiles (FlashAttention-4 style): P is scaled with a
//     sticky maximum that is only raised when the row maximum grew by more than 2^8; the (rare) rescale is a warp-
//     local tcgen05.ld / multiply / tcgen05.st of the warp's own 32 TMEM lanes.  No per-tile PV read-out or fold;
//   * the two streams are merged once at the end through shared memory.
// TMEM (512 columns): S 64 | C 192 | G copies 64+32+32 | O_0 64 | O_1 64; Q and P are shared-memory A operands.
//
// One CTA (320 threads) per (128-query tile, head, batch row), looping over 64-key tiles t:
//   warp 0   TMA producer: ring A (x2) = K_t + the two 192-row slices, ring B (x2) = V_t
//   warp 1   tcgen05 issuer, event driven (polls which of  M1(t) | PV_0(t) | PV_1(t)  is ready)
//   warps 2-9  softmax: thread = (query row / TMEM lane, stream x)
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
constexpr int E = 32;              // keys per stream per tile
constexpr int NW = (E + 32) / 2;   // packed words of a thread's C window (its E keys x the 32 lane shifts of its warp)
constexpr int ATT_THREADS = 64 + 256;
constexpr int SLICE = 192;         // table rows per tile
constexpr int EXP_CENTER = 2047;
constexpr int EXP_ROWS = 4096;
constexpr int TMAX = 2048 / KT;
constexpr float RESCALE_LOG2 = 8.0f;   // raise the sticky maximum only when 2^(new - old) would exceed 2^8

// shared memory map (bytes, from a 1024-aligned base)
constexpr int OFF_Q = 0;                           // 128 x 128 B
constexpr int OFF_V = OFF_Q + 16384;               // 2 x 8 KB
constexpr int POS_BYTES = SLICE * 128;
constexpr int EKK_BYTES = POS_BYTES + KT * 128;    // EK slice (192 rows) immediately followed by the K tile (64 rows): one N = 256 B operand
constexpr int OFF_EK = OFF_V + 16384;              // 2 x 256 x 128 B   (reused as the 2 x 32 x 128 float merge buffer at the end)
constexpr int OFF_EQ = OFF_EK + 2 * EKK_BYTES;     // 2 x 192 x 128 B
constexpr int OFF_P = OFF_EQ + 2 * POS_BYTES;      // 128 x 128 B: P tile, K-major; stream x owns 16-byte chunks 4x..4x+3 of each row
constexpr int OFF_ML = OFF_P + 16384;              // 2 streams x 128 rows x (m, l)
constexpr int OFF_MASK = OFF_ML + 2 * QT * 8;      // uint32[68]: key-validity words of this batch row
constexpr int OFF_BAR = OFF_MASK + 68 * 4;
constexpr int ATT_SMEM = OFF_BAR + 256 + 1024;
static_assert(OFF_BAR % 8 == 0, "barrier alignment");
static_assert(ATT_SMEM <= 227 * 1024, "attention smem budget");
static_assert(2 * 32 * QT * 4 <= 2 * EKK_BYTES, "merge buffer fits in the slice rings");

// TMEM columns
constexpr uint32_t TM_C = 0;       // 192   one N = 256 MMA writes C | S
constexpr uint32_t TM_S = 192;     // 64
constexpr uint32_t TM_G32 = 256;   // 64: window rows 32..159, keys 0..63
constexpr uint32_t TM_G64 = 320;   // 32: window rows 64..191, keys 0..31
constexpr uint32_t TM_G0 = 352;    // 32: window rows 0..127,  keys 32..63
constexpr uint32_t TM_O = 384;     // 2 x 64

struct StreamParams {
  long long* trace;          // TRACE instantiation only: [2 roles][TMAX][8] clock64 stamps of CTA (1,0,0)
  const uint32_t* mask_bits; // [B][ceil(S/32)]
  const int32_t* kv_len;     // [B]
  __half* ctx;               // [B*S, H]
  int B, S, heads, H;
  float scale_log2;          // log2(e) / sqrt(3*d)
  int stagger;               // developer switch (GLC_ATTN_STAGGER=0 disables the half-period offset of stream 1)
};

#define GLC_TRACE(role, tile, slot)                                                                              \
  do {                                                                                                           \
    if (TRACE && p.trace && (threadIdx.x & 31) == 0 && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0)    \
      p.trace[((role) * TMAX + (tile)) * 8 + (slot)] = clock64();                                                \
  } while (0)

__device__ __forceinline__ uint32_t sel(bool p, uint32_t a, uint32_t b) { return p ? a : b; }

template <bool TRACE>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_stream_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_ek,
                        const __grid_constant__ CUtensorMap tm_eq, const StreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* a_full = bars + 1;       // [2]  K_t + slices landed
  uint64_t* a_empty = bars + 3;      // [2]  ... consumed by the S / C / G MMAs
  uint64_t* b_full = bars + 5;       // [2]  V_t landed
  uint64_t* b_empty = bars + 7;      // [2]  ... consumed by both PV MMAs
  uint64_t* cs_full = bars + 9;      // C | S accumulators of tile t ready
  uint64_t* g_full = bars + 10;      // G copies of tile t ready
  uint64_t* cs_free = bars + 11;     // ... drained by the 8 softmax warps
  uint64_t* g_free = bars + 12;
  uint64_t* p_full = bars + 13;      // [2 streams]  P_x written (and O_x rescaled if needed)
  uint64_t* pv_full = bars + 15;     // [2 streams]  O_x += P_x . V_x retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  uint32_t* kmask = reinterpret_cast<uint32_t*>(smem + OFF_MASK);

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
    ptx::prefetch_tensormap(&tm_ek);
    ptx::prefetch_tensormap(&tm_eq);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&a_full[s], 1);
      ptx::mbar_init(&a_empty[s], 1);
      ptx::mbar_init(&b_full[s], 1);
      ptx::mbar_init(&b_empty[s], 1);
      ptx::mbar_init(&p_full[s], 4);
      ptx::mbar_init(&pv_full[s], 1);
    }
    ptx::mbar_init(cs_full, 1);
    ptx::mbar_init(g_full, 1);
    ptx::mbar_init(cs_free, 8);
    ptx::mbar_init(g_free, 8);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  {
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
      ptx::mbar_arrive_expect_tx(q_full, QT * 128);
      ptx::tma_load_3d(smem + OFF_Q, &tm_qkv, q_full, head * D, q0, b);
      ptx::tma_load_3d(smem + OFF_Q + 8192, &tm_qkv, q_full, head * D, q0 + 64, b);
      // ring A: K_x and the position slices of tile x; free once the M1 MMAs of tile x-2 retired
      auto load_a = [&](int x) {
        const int k0 = x * KT;
        const int rho0 = EXP_CENTER - (QT - 1) - q0 + k0;
        const int sig0 = EXP_CENTER - (KT - 1) + q0 - k0;
        const int st = x & 1;
        ptx::mbar_wait(&a_empty[st], ((x >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&a_full[st], (uint32_t)(KT * 128 + 2 * POS_BYTES));
        ptx::tma_load_3d(smem + OFF_EK + st * EKK_BYTES + POS_BYTES, &tm_qkv, &a_full[st], p.H + head * D, k0, b);
#pragma unroll
        for (int bx = 0; bx < SLICE / 64; ++bx) {
          ptx::tma_load_3d(smem + OFF_EK + st * EKK_BYTES + bx * 8192, &tm_ek, &a_full[st], 0, rho0 + bx * 64, head);
          ptx::tma_load_3d(smem + OFF_EQ + st * POS_BYTES + bx * 8192, &tm_eq, &a_full[st], 0, sig0 + bx * 64, head);
        }
      };
      // ring B: V_x; free once both PV MMAs of tile x-2 retired
      auto load_b = [&](int x) {
        const int st = x & 1;
        ptx::mbar_wait(&b_empty[st], ((x >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&b_full[st], KT * 128);
        ptx::tma_load_3d(smem + OFF_V + st * 8192, &tm_qkv, &b_full[st], 2 * p.H + head * D, x * KT, b);
      };
      load_a(0);
      load_b(0);
      if (T > 1) load_a(1);
      for (int t = 0; t < T; ++t) {
        if (t + 2 < T) load_a(t + 2);
        if (t + 1 < T) load_b(t + 1);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: event driven, warp-uniform
    // All 32 lanes evaluate the (warp-uniform) readiness tests with the non-blocking mbarrier probe; one elected
    // lane issues.  Tasks per tile:  CS(t) = C | S   ->   GG(t) = G copies   ->   PV_0(t), PV_1(t);  the softmax warps
    // drain C, S while GG(t) executes and G while CS(t+1) executes, so the tensor pipe never waits for a drain.
    constexpr uint32_t idesc_n32 = ptx::idesc_f16(128, 32);
    constexpr uint32_t idesc_n64 = ptx::idesc_f16(128, 64);
    constexpr uint32_t idesc_cs = ptx::idesc_f16(128, 256);
    constexpr uint32_t idesc_pv = ptx::idesc_f16(128, 64, 0, 1);   // B (=V) is MN-major
    const uint64_t dQ = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_Q));
    const uint64_t dP = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_P));
    auto ready = [&](uint64_t* bar, uint32_t parity) { return __all_sync(0xffffffffu, ptx::mbar_test_wait(bar, parity)) != 0; };
    ptx::mbar_wait(q_full, 0);
    int tc = 0, tg = 0;  // next tile of the CS / GG tasks
    int tp[2] = {0, 0};  // next PV tile per stream
    uint32_t idle = 0;
    while (tp[0] < T || tp[1] < T) {
      bool progress = false;
      if (tc < T && ready(&a_full[tc & 1], (tc >> 1) & 1) && (tc == 0 || ready(cs_free, (tc - 1) & 1))) {
        const uint64_t dEKK = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_EK + (tc & 1) * EKK_BYTES));
        ptx::tc_fence_after();
        GLC_TRACE(1, tc, 0);
        if (ptx::elect_one()) {
          // descriptors advance 32 B (= 2) per 16-element k step
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem + TM_C, dQ + 2 * k, dEKK + 2 * k, idesc_cs, (uint32_t)(k != 0));
          ptx::mma_commit(cs_full);
        }
        __syncwarp();
        GLC_TRACE(1, tc, 1);
        ++tc;
        progress = true;
      }
      if (tg < tc && (tg == 0 || ready(g_free, (tg - 1) & 1))) {
        const int st = tg & 1;
        const uint64_t dK = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_EK + st * EKK_BYTES + POS_BYTES));
        const uint64_t dEQ = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_EQ + st * POS_BYTES));
        ptx::tc_fence_after();
        GLC_TRACE(1, tg, 2);
        if (ptx::elect_one()) {
          // 32 table/key rows = 4096 B = 256 in a descriptor
#pragma unroll
          for (int k = 0; k < 4; ++k)   // rows 32..159 x keys 0..63
            ptx::mma_f16_ss(tmem + TM_G32, dEQ + 256 + 2 * k, dK + 2 * k, idesc_n64, (uint32_t)(k != 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // rows 64..191 x keys 0..31
            ptx::mma_f16_ss(tmem + TM_G64, dEQ + 512 + 2 * k, dK + 2 * k, idesc_n32, (uint32_t)(k != 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // rows 0..127 x keys 32..63
            ptx::mma_f16_ss(tmem + TM_G0, dEQ + 2 * k, dK + 256 + 2 * k, idesc_n32, (uint32_t)(k != 0));
          ptx::mma_commit(g_full);
          ptx::mma_commit(&a_empty[st]);
        }
        __syncwarp();
        GLC_TRACE(1, tg, 3);
        ++tg;
        progress = true;
      }
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int t = tp[x];
        if (t < tg && ready(&p_full[x], t & 1) && ready(&b_full[t & 1], (t >> 1) & 1)) {
          const int st = t & 1;
          const uint64_t dV = ptx::smem_desc_sw128(ptx::smem_u32(smem + OFF_V + st * 8192));
          ptx::tc_fence_after();
          GLC_TRACE(1, t, 4 + x);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 2 * x; k < 2 * x + 2; ++k)   // V is MN-major: 16 keys further = +2048 bytes = +128 in the descriptor
              ptx::mma_f16_ss(tmem + TM_O + 64 * x, dP + 2 * k, dV + 128 * k, idesc_pv, (uint32_t)(t > 0 || k != 2 * x));
            ptx::mma_commit(&pv_full[x]);
            if (tp[x ^ 1] > t) ptx::mma_commit(&b_empty[st]);   // the other stream's PV of this tile is already in flight
          }
          __syncwarp();
          tp[x] = t + 1;
          progress = true;
        }
      }
      if (!progress && ++idle > (1u << 28)) __trap();   // protocol bug: surface as an error, never hang
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int sw = warp - 2;          // 0..7
    const int x = sw >> 2;            // stream: keys [32x, 32x+32) of every tile
    const int qd = warp & 3;          // TMEM lane quarter
    const int i = qd * 32 + lane;     // row in the query tile
    const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);

    // tile-independent addresses and shift controls
    const uint32_t a_s = t_lane + TM_S + (uint32_t)(32 * x);
    const uint32_t a_c = t_lane + TM_C + (uint32_t)(96 - 32 * qd + 32 * x);   // window of this warp: its 32 keys x 32 lane shifts
    const uint32_t a_lo = t_lane + (x == 0 ? TM_G32 : TM_G0);                  // copy 32 s0
    const uint32_t a_hi = t_lane + (x == 0 ? TM_G64 : TM_G32 + 32);            // copy 32 s0 + 32
    const uint32_t a_o = t_lane + TM_O + (uint32_t)(64 * x);
    const int sh = 31 - lane;                  // c2p: element shift inside the window
    const bool sh16 = sh & 16, sh8 = sh & 8, sh4 = sh & 4, sh2 = sh & 2;
    const uint32_t prmt_sel = (sh & 1) ? 0x5432u : 0x3210u;
    const int rot0 = lane + 31;                // p2c: source lane of column jj is (rot0 - jj) & 31
    const int thr0 = lane - 31;                // ... and this lane supplies the upper copy iff thr0 + jj < 0
    uint8_t* prow = smem + OFF_P + (i >> 3) * 1024 + (i & 7) * 128;

    float m_run = -CUDART_INF_F;   // sticky maximum (raw score units) the stream's P and O are scaled with
    float l_run = 0.f;
    const float sc = p.scale_log2;
    const bool stagger = p.stagger != 0;

    for (int t = 0; t < T; ++t) {
      const int k0 = t * KT;
      if (sw == 0) GLC_TRACE(0, t, 0);
      ptx::mbar_wait(cs_full, t & 1);
      // stagger: stream 1 starts its tile only when the G copies have landed too (about half a tile period after C | S),
      // so that its TMEM drain / ALU skew / MUFU phases fall on different units than stream 0's at any time
      if (stagger && x == 1) ptx::mbar_wait(g_full, t & 1);
      ptx::tc_fence_after();
      if (sw == 0) GLC_TRACE(0, t, 1);

      float s[E];
      uint32_t w[NW];
      {
        uint32_t r[32];
        ptx::tmem_ld_x32(a_s, r);
        uint32_t c0[32], c1[32];
        ptx::tmem_ld_x32(a_c, c0);
        ptx::tmem_ld_x32(a_c + 32, c1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(cs_free);
#pragma unroll
        for (int jj = 0; jj < E; ++jj) s[jj] = __uint_as_float(r[jj]);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          w[k] = ptx::pack_f16(__uint_as_float(c0[2 * k]), __uint_as_float(c0[2 * k + 1]));
          w[16 + k] = ptx::pack_f16(__uint_as_float(c1[2 * k]), __uint_as_float(c1[2 * k + 1]));
        }
      }
      // ---- c2p: shift the packed window left by sh elements
#pragma unroll
      for (int k = 0; k < NW - 8; ++k) w[k] = sel(sh16, w[k + 8], w[k]);
#pragma unroll
      for (int k = 0; k < NW - 12; ++k) w[k] = sel(sh8, w[k + 4], w[k]);
#pragma unroll
      for (int k = 0; k < NW - 14; ++k) w[k] = sel(sh4, w[k + 2], w[k]);
#pragma unroll
      for (int k = 0; k < NW - 15; ++k) w[k] = sel(sh2, w[k + 1], w[k]);
#pragma unroll
      for (int m = 0; m < E / 2; ++m) {
        const uint32_t xw = __byte_perm(w[m], w[m + 1], prmt_sel);
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&xw));
        s[2 * m] += f.x;
        s[2 * m + 1] += f.y;
      }
      if (sw == 0) GLC_TRACE(0, t, 2);

      // ---- p2c: lane rotation by s1 = 31 - jj, source lane picks the copy
      ptx::mbar_wait(g_full, t & 1);
      ptx::tc_fence_after();
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t lo[16], hi[16];
        ptx::tmem_ld_x16(a_lo + 16 * u, lo);
        ptx::tmem_ld_x16(a_hi + 16 * u, hi);
        ptx::tmem_ld_wait();
        if (u == 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(g_free);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int jj = 16 * u + k;
          const uint32_t v = sel(thr0 + jj < 0, hi[k], lo[k]);
          s[jj] += __uint_as_float(__shfl_sync(0xffffffffu, v, rot0 - jj));
        }
      }

      const int kb = k0 + 32 * x;
      const uint32_t kbits = kmask[kb >> 5];
      if (kbits != 0xffffffffu) {
#pragma unroll
        for (int jj = 0; jj < E; ++jj)
          if (!((kbits >> jj) & 1u)) s[jj] = -CUDART_INF_F;
      }
      float mloc = s[0];
#pragma unroll
      for (int jj = 1; jj < E; ++jj) mloc = fmaxf(mloc, s[jj]);
      if (sw == 0) GLC_TRACE(0, t, 3);

      // ---- sticky maximum: raise it (and rescale l and the O rows of this warp) only when needed
      const bool raise = (mloc - m_run) * sc > RESCALE_LOG2;   // also true for the first finite maximum (m_run = -inf)
      if (t > 0) {
        ptx::mbar_wait(&pv_full[x], (t - 1) & 1);   // P buffer free again, O_x stable
        if (__any_sync(0xffffffffu, raise)) {
          ptx::tc_fence_after();
          const float alpha = raise ? ptx::ex2((m_run - mloc) * sc) : 1.0f;   // m_run = -inf: 0 (O and l hold exact zeros)
          l_run *= alpha;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint32_t r[32];
            ptx::tmem_ld_x32(a_o + 32 * u, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * alpha);
            ptx::tmem_st_x32(a_o + 32 * u, r);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
        }
      }
      if (raise) m_run = mloc;
      if (sw == 0) GLC_TRACE(0, t, 4);

      const float neg_ms = (m_run == -CUDART_INF_F) ? 0.f : -m_run * sc;
      float psum = 0.f;
#pragma unroll
      for (int jj = 0; jj < E; ++jj) {
        const float e = ptx::ex2(fmaf(s[jj], sc, neg_ms));
        s[jj] = e;
        psum += e;
      }
      l_run += psum;
      if (sw == 0) GLC_TRACE(0, t, 5);

      // ---- P_x -> shared memory (K-major, 128-byte swizzle): row i, 16-byte chunks 4x..4x+3
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 o4;
        o4.x = ptx::pack_f16(s[8 * u + 0], s[8 * u + 1]);
        o4.y = ptx::pack_f16(s[8 * u + 2], s[8 * u + 3]);
        o4.z = ptx::pack_f16(s[8 * u + 4], s[8 * u + 5]);
        o4.w = ptx::pack_f16(s[8 * u + 6], s[8 * u + 7]);
        *reinterpret_cast<uint4*>(prow + (((4 * x + u) ^ (i & 7)) << 4)) = o4;
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[x]);
      if (sw == 0) GLC_TRACE(0, t, 6);
    }

    // ---- last PV retired: merge the two streams (each finishes 32 of the 64 output dims)
    ptx::mbar_wait(&pv_full[x], (T - 1) & 1);
    ptx::tc_fence_after();
    float* mbuf = reinterpret_cast<float*>(smem + OFF_EK);   // [stream][32 dims][128 rows]
    float2* ml = reinterpret_cast<float2*>(smem + OFF_ML);   // [stream][128 rows]
    float mine[32], theirs[32];
    {
      uint32_t r[32];
      ptx::tmem_ld_x32(a_o + 32 * (x ^ 1), r);   // the half the partner finishes
      ptx::tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) theirs[k] = __uint_as_float(r[k]);
      ptx::tmem_ld_x32(a_o + 32 * x, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) mine[k] = __uint_as_float(r[k]);
    }
    // both warps of the lane quarter are past their last pv_full: every MMA of the CTA has retired (tcgen05 commits
    // are in issue order) and every TMA load was consumed, so the slice rings can be reused as the merge buffer
    ptx::named_bar_sync(2 + qd, 64);
#pragma unroll
    for (int k = 0; k < 32; ++k) mbuf[(x * 32 + k) * QT + i] = theirs[k];
    ml[x * QT + i] = make_float2(m_run, l_run);
    ptx::named_bar_sync(2 + qd, 64);
    const float2 other = ml[(x ^ 1) * QT + i];
    const float m_all = fmaxf(m_run, other.x);
    float w_me = 0.f, w_ot = 0.f;
    if (m_all != -CUDART_INF_F) {
      w_me = ptx::ex2((m_run - m_all) * sc);
      w_ot = ptx::ex2((other.x - m_all) * sc);
    }
    const float l_tot = w_me * l_run + w_ot * other.y;
    const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
    w_me *= inv;
    w_ot *= inv;
    const int row = q0 + i;
    float o[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) o[k] = w_me * mine[k] + w_ot * mbuf[((x ^ 1) * 32 + k) * QT + i];
    if (row < p.S) {
      __half* dst = p.ctx + ((int64_t)b * p.S + row) * p.H + head * D + 32 * x;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint4 o4;
        o4.x = ptx::pack_f16(o[8 * v + 0], o[8 * v + 1]);
        o4.y = ptx::pack_f16(o[8 * v + 2], o[8 * v + 3]);
        o4.z = ptx::pack_f16(o[8 * v + 4], o[8 * v + 5]);
        o4.w = ptx::pack_f16(o[8 * v + 6], o[8 * v + 7]);
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

}  // namespace

cudaError_t attention_stream(const void* qkv, const void* exp_k, const void* exp_qr, int64_t ld_exp,
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
  StreamParams p;
  p.trace = nullptr;
  p.mask_bits = mask_bits;
  p.kv_len = kv_len;
  p.ctx = (__half*)ctx;
  p.B = B; p.S = S; p.heads = heads; p.H = H;
  p.scale_log2 = 1.4426950408889634f / sqrtf(3.0f * D);
  static const int stagger = [] { const char* e = getenv("GLC_ATTN_STAGGER"); return (e && e[0] == '0') ? 0 : 1; }();
  p.stagger = stagger;
  dim3 grid((S + QT - 1) / QT, heads, B);
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attention_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  // developer aid: GLC_ATTN_TRACE=<file> dumps per-tile clock64 stamps of CTA (1,0,0) (synchronous)
  if (const char* tf = getenv("GLC_ATTN_TRACE")) {
    const size_t n = 2 * TMAX * 8;
    if (cudaMalloc(&p.trace, n * sizeof(long long)) != cudaSuccess) return cudaGetLastError();
    cudaMemsetAsync(p.trace, 0, n * sizeof(long long), stream);
    attention_stream_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm_qkv, tm_ek, tm_eq, p);
    cudaError_t e = cudaStreamSynchronize(stream);
    std::vector<long long> h(n);
    cudaMemcpy(h.data(), p.trace, n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    if (FILE* f = fopen(tf, "w")) {
      long long t0 = 0;
      for (size_t k = 0; k < n; ++k) if (h[k] && (!t0 || h[k] < t0)) t0 = h[k];
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
  attention_stream_kernel<false><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm_qkv, tm_ek, tm_eq, p);
  return cudaGetLastError();
}

}  // namespace glc
