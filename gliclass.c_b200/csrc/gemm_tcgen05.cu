// K2 — dense projection GEMM for sm_100a:  C[M,N] = act(A[M,K] · W[N,K]^T + bias[N])
// A fp16, W fp16, fp32 accumulation, C fp16 (or fp32).  (tcgen05 kind::f16 needs A and B in the
// same 16-bit format: an fp16 x bf16 descriptor raises an illegal-instruction fault on B200.)
//
// Replaces the MatMul+Add(+Erf-GELU) groups ORT executes for query/key/value_proj,
// attention.output.dense, intermediate.dense and output.dense (arithmetic
// T:231-233, T:49-53, T:393-396, T:408-412 of transformers' modeling_deberta_v2.py), and the
// GLiClass projector Linears.
//
// Design (one CTA per SM, persistent over 128 x BN output tiles):
//   warp 0   : TMA producer   — cp.async.bulk.tensor 2D, 128B swizzle, BK = 64 per stage
//   warp 1   : MMA issuer     — tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16 x4 per stage,
//                               fp32 accumulators in TMEM, double buffered (2 x BN columns)
//   warps 2.. : epilogue      — EPI groups of 4 warps (one warp per TMEM lane quarter); group g takes
//                               the 32-column chunks c = g, g+EPI, ...: tcgen05.ld 32x32b -> +bias ->
//                               (erf-GELU) -> fp16/fp32 -> global.  EPI = 4 for the GELU variant (the
//                               epilogue, not the MMA, bounds FFN1 otherwise), 2 for the rest.
// Pipelines: smem ring full/empty (TMA <-> MMA) and TMEM full/empty (MMA <-> epilogue), all mbarrier.
// Ragged M/N/K are handled by TMA out-of-bounds zero fill on loads and predication on stores.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"
#include "tma_desc.h"

namespace glc {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
template <int ACT>
struct EpiCfg {
  static constexpr int GROUPS = (ACT == 1 || ACT == 3) ? 4 : 2;   // epilogue warp groups (4 warps each)
  static constexpr int THREADS = 64 + 128 * GROUPS;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;   // +1024: manual alignment slack
  static constexpr uint32_t TMEM_COLS = 2 * BN;
};

// erf-GELU: x * Phi(x) = 0.5 x (1 + erf(x / sqrt 2)), with erf(z) = tanh(z (a0 + a1 z^2 + a2 z^4)):
// atanh(erf(z)) is a smooth odd function, and a least-squares fit of its odd quintic (weighted by the
// GELU error 0.5 x derr) gives max |GELU error| 2.9e-5 for all x -- far below the fp16 rounding of the
// stored result -- with ONE MUFU op (tanh.approx, rel. error 2^-11) and 7 FP32 ops per element instead of
// two MUFU + ~15 (the erfc form with rcp and ex2 made the FFN1 epilogue, not the MMA, pace the tile).
// In x: p = x (b0 + b1 x^2 + b2 x^4); x^2 is clamped at 49 where tanh has long saturated (b2 < 0).
__device__ __forceinline__ float gelu_erf(float x) {
  const float x2 = fminf(x * x, 49.0f);
  float q = fmaf(-3.5785683e-4f, x2, 3.7043383e-2f);
  q = fmaf(q, x2, 7.9746913e-1f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(q * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// The same fit evaluated on two values in packed fp16 (HFMA2 / MUFU.TANH.F16x2): half the instructions per element.  Only
// for the e4m3-output epilogue, whose 3-bit result mantissa hides fp16's 2^-11 (the FP8 FFN1 tile is paced by its
// epilogue, not by the e4m3 MMA, which runs at twice the fp16 rate).
__device__ __forceinline__ __half2 gelu_erf_h2(__half2 x) {
  const __half2 x2 = __hmin2(__hmul2(x, x), __float2half2_rn(49.0f));
  __half2 q = __hfma2(__float2half2_rn(-3.5785683e-4f), x2, __float2half2_rn(3.7043383e-2f));
  q = __hfma2(q, x2, __float2half2_rn(7.9746913e-1f));
  const __half2 p = __hmul2(q, x);
  uint32_t t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&p)));
  const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
  return __hfma2(hx, *reinterpret_cast<const __half2*>(&t), hx);
}

// SwiGLU (ACT = 3; Qwen2 / Llama MLP, transformers modeling_qwen2.py:46-48): the weight rows are interleaved in blocks
// of 32 — rows [64 j, 64 j + 32) = gate_proj rows [32 j, 32 j + 32), rows [64 j + 32, 64 j + 64) = the matching up_proj
// rows — so an epilogue warp holds a gate chunk and its up chunk and writes silu(gate) * up: the output has N / 2 columns.
// silu(x) = x sigmoid(x) = 0.5 x (1 + tanh(0.5 x)): one MUFU op.
__device__ __forceinline__ float silu_mul(float g, float u) {
  float t;
  const float hg = 0.5f * g;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hg));
  return fmaf(hg, t, hg) * u;
}

// Residual operand of the fused "+ r" epilogue: 32 fp16 of this lane's row (4 x 128-bit; the second half of every 32-byte
// sector a load touches is the next load's data, so L1 serves it).  Issued BEFORE the TMEM load they are added to.
struct ResidRow {
  uint4 q[4];
  __device__ __forceinline__ void load(const __half* __restrict__ resid, int64_t ldr, int row, int n, int M, int N) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q[j] = make_uint4(0u, 0u, 0u, 0u);
      if (row < M && n + 8 * j + 8 <= N) q[j] = *reinterpret_cast<const uint4*>(resid + (int64_t)row * ldr + n + 8 * j);
    }
  }
  __device__ __forceinline__ void add_to(float (&v)[32]) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2* h = reinterpret_cast<const __half2*>(&q[j]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h[k]);
        v[8 * j + 2 * k] += f.x;
        v[8 * j + 2 * k + 1] += f.y;
      }
    }
  }
};

// Epilogue of one accumulator tile for one warp: TMEM lanes of this warp's quarter (t_row), the
// 32-column chunks c = grp, grp+GROUPS, ...: tcgen05.ld -> +bias -> (erf-GELU) -> fp16/fp32 -> global.
template <int BN, int ACT, bool OUT_F32, bool RESID>
__device__ __forceinline__ void epilogue_tile(uint32_t t_row, int grp, int n0, int row, const float* __restrict__ bias,
                                              void* __restrict__ Cout, int64_t ldc, int M, int N,
                                              const __half* __restrict__ resid, int64_t ldr) {
#pragma unroll 1
  for (int c = grp; c < BN / 32; c += EpiCfg<ACT>::GROUPS) {
  const int n = n0 + c * 32;
  ResidRow rr;
  if (RESID && n < N) rr.load(resid, ldr, row, n, M, N);
  uint32_t r[32];
  ptx::tmem_ld_x32(t_row + (uint32_t)(c * 32), r);
  ptx::tmem_ld_wait();
  if (n >= N) continue;   // warp-uniform
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (bias != nullptr) {
    if (n + 32 <= N) {
      const float4* b4 = reinterpret_cast<const float4*>(bias + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = __ldg(b4 + j);
        v[4 * j + 0] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (n + j < N) v[j] += __ldg(bias + n + j);
    }
  }
  if (ACT == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  }
  if (ACT == 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (RESID) rr.add_to(v);
  if (row < M) {
    if (OUT_F32) {
      float* dst = reinterpret_cast<float*>(Cout) + (int64_t)row * ldc + n;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (n + 4 * j + 4 <= N) reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      __half* dst = reinterpret_cast<__half*>(Cout) + (int64_t)row * ldc + n;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + 8 * j + 8 <= N) {
          uint4 o;
          o.x = ptx::pack_f16(v[8 * j + 0], v[8 * j + 1]);
          o.y = ptx::pack_f16(v[8 * j + 2], v[8 * j + 3]);
          o.z = ptx::pack_f16(v[8 * j + 4], v[8 * j + 5]);
          o.w = ptx::pack_f16(v[8 * j + 6], v[8 * j + 7]);
          reinterpret_cast<uint4*>(dst)[j] = o;
        }
      }
    }
  }
}
}

// SwiGLU epilogue, direct stores (small problems / fp32 output): chunk pairs (gate, up) -> 32 output columns
template <int BN, bool OUT_F32>
__device__ __forceinline__ void epilogue_tile_swiglu(uint32_t t_row, int grp, int n0, int row, const float* __restrict__ bias,
                                                     void* __restrict__ Cout, int64_t ldc, int M, int N) {
#pragma unroll 1
  for (int pc = grp; pc < BN / 64; pc += EpiCfg<3>::GROUPS) {
    const int n = n0 + pc * 64;          // gate columns [n, n+32), up columns [n+32, n+64) of the interleaved GEMM
    uint32_t rg[32], ru[32];
    ptx::tmem_ld_x32(t_row + (uint32_t)(pc * 64), rg);
    ptx::tmem_ld_x32(t_row + (uint32_t)(pc * 64 + 32), ru);
    ptx::tmem_ld_wait();
    if (n + 64 > N) continue;            // warp-uniform (N is a multiple of 64 for this activation)
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float g = __uint_as_float(rg[j]), u = __uint_as_float(ru[j]);
      if (bias != nullptr) { g += __ldg(bias + n + j); u += __ldg(bias + n + 32 + j); }
      v[j] = silu_mul(g, u);
    }
    if (row < M) {
      const int no = n / 2;
      if (OUT_F32) {
        float* dst = reinterpret_cast<float*>(Cout) + (int64_t)row * ldc + no;
#pragma unroll
        for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        __half* dst = reinterpret_cast<__half*>(Cout) + (int64_t)row * ldc + no;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          o.x = ptx::pack_f16(v[8 * j + 0], v[8 * j + 1]);
          o.y = ptx::pack_f16(v[8 * j + 2], v[8 * j + 3]);
          o.z = ptx::pack_f16(v[8 * j + 4], v[8 * j + 5]);
          o.w = ptx::pack_f16(v[8 * j + 6], v[8 * j + 7]);
          reinterpret_cast<uint4*>(dst)[j] = o;
        }
      }
    }
  }
}

// SwiGLU epilogue through shared memory + TMA store (tm_c describes the [M, N/2] output)
template <int BN, int NBUF>
__device__ __forceinline__ void epilogue_tile_tma_swiglu(uint32_t t_row, int grp, int n0, int row0, const float* __restrict__ bias,
                                                         const CUtensorMap* tm_c, uint8_t* stage, int& buf, int lane, int N) {
#pragma unroll 1
  for (int pc = grp; pc < BN / 64; pc += EpiCfg<3>::GROUPS) {
    const int n = n0 + pc * 64;
    uint32_t rg[32], ru[32];
    ptx::tmem_ld_x32(t_row + (uint32_t)(pc * 64), rg);
    ptx::tmem_ld_x32(t_row + (uint32_t)(pc * 64 + 32), ru);
    ptx::tmem_ld_wait();
    if (n + 64 > N) continue;   // warp-uniform
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float g = __uint_as_float(rg[j]), u = __uint_as_float(ru[j]);
      if (bias != nullptr) { g += __ldg(bias + n + j); u += __ldg(bias + n + 32 + j); }
      v[j] = silu_mul(g, u);
    }
    if (lane == 0) ptx::bulk_wait_group_read<NBUF - 1>();
    __syncwarp();
    uint8_t* sb = stage + buf * 2048;
    uint8_t* srow = sb + lane * 64;
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 o;
      o.x = ptx::pack_f16(v[8 * j + 0], v[8 * j + 1]);
      o.y = ptx::pack_f16(v[8 * j + 2], v[8 * j + 3]);
      o.z = ptx::pack_f16(v[8 * j + 4], v[8 * j + 5]);
      o.w = ptx::pack_f16(v[8 * j + 6], v[8 * j + 7]);
      *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = o;
    }
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_2d(tm_c, sb, n / 2, row0);
      ptx::bulk_commit_group();
    }
    buf = (buf + 1 == NBUF) ? 0 : buf + 1;
  }
}

// FP8 (e4m3) operands: the accumulator is sum(a8 * w8); the real product is that times the activation scale of the row
// (per-row dynamic scale from the LN kernel, or one constant for a statically scaled activation) times the weight
// scale of the output channel (per-row-of-W scale from the load-time quantiser).  out_mult: the static multiplier an
// e4m3 OUTPUT is stored with (its consumer passes 1 / out_mult as `k`).
struct Fp8Scales {
  const float* row;   // [M] or nullptr
  const float* col;   // [N]
  float k;            // constant factor (1 when `row` carries everything)
  float out_mult;
  // ACT = 4 (rotary embedding fused into the QKV projection of the decoder stack, head dim 128): columns < rope_cols are
  // rotated in pairs (p, p + 64) of their head by the angle of the row's position; cs = (cos, sin) [positions][64],
  // position = rope_tile_pos[row / 128] + row % 128 (packed layout) or row % rope_S
  const float2* rope_cs;
  const int32_t* rope_tile_pos;
  int rope_S, rope_cols;
};

template <bool SCALED>
__device__ __forceinline__ void apply_scales(float (&v)[32], const Fp8Scales& sc, int row, int n, int M, int N) {
  if (!SCALED) return;
  const float rs = (sc.row != nullptr && row < M) ? __ldg(sc.row + row) * sc.k : sc.k;
  if (n + 32 <= N) {
    const float4* c4 = reinterpret_cast<const float4*>(sc.col + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 cc = __ldg(c4 + j);
      v[4 * j + 0] *= rs * cc.x; v[4 * j + 1] *= rs * cc.y; v[4 * j + 2] *= rs * cc.z; v[4 * j + 3] *= rs * cc.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) if (n + j < N) v[j] *= rs * __ldg(sc.col + n + j);
  }
}

// e4m3 OUTPUT epilogue (FFN1 of the FP8 path: scale -> +bias -> erf-GELU -> * out_mult -> saturating e4m3), 32 rows x 32
// bytes per chunk through an unswizzled staging buffer and one bulk tensor store (tm_c: uint8 [M, N], box 32 x 32).
template <int BN, int ACT, int NBUF>
__device__ __forceinline__ void epilogue_tile_tma_f8out(uint32_t t_row, int grp, int n0, int row0, const float* __restrict__ bias,
                                                        const CUtensorMap* tm_c, uint8_t* stage, int& buf, int lane, int M, int N,
                                                        const Fp8Scales& sc) {
#pragma unroll 1
  for (int c = grp; c < BN / 32; c += EpiCfg<ACT>::GROUPS) {
    const int n = n0 + c * 32;
    uint32_t r[32];
    ptx::tmem_ld_x32(t_row + (uint32_t)(c * 32), r);
    ptx::tmem_ld_wait();
    if (n >= N) continue;   // warp-uniform
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (n + 32 <= N) {
      // scale and bias in one FFMA per element (this tile is paced by its epilogue: ncu shows 64 % of the issue slots busy
      // against 46 % tensor-pipe activity)
      const int row = row0 + lane;
      const float rs = (sc.row != nullptr && row < M) ? __ldg(sc.row + row) * sc.k : sc.k;
      const float4* c4 = reinterpret_cast<const float4*>(sc.col + n);
      const float4* b4 = reinterpret_cast<const float4*>(bias + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 cc = __ldg(c4 + j);
        const float4 bb = bias != nullptr ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * j + 0] = fmaf(v[4 * j + 0], rs * cc.x, bb.x);
        v[4 * j + 1] = fmaf(v[4 * j + 1], rs * cc.y, bb.y);
        v[4 * j + 2] = fmaf(v[4 * j + 2], rs * cc.z, bb.z);
        v[4 * j + 3] = fmaf(v[4 * j + 3], rs * cc.w, bb.w);
      }
    } else {
      apply_scales<true>(v, sc, row0 + lane, n, M, N);
      if (bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (n + j < N) v[j] += __ldg(bias + n + j);
      }
    }
    // activation and the output multiplier in packed fp16, then f16x2 -> e4m3x2 (saturating)
    const __half2 om2 = __float2half2_rn(sc.out_mult);
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t h01 = ptx::pack_f16(v[4 * j + 0], v[4 * j + 1]), h23 = ptx::pack_f16(v[4 * j + 2], v[4 * j + 3]);
      __half2 a = *reinterpret_cast<__half2*>(&h01), b = *reinterpret_cast<__half2*>(&h23);
      if (ACT == 1) { a = gelu_erf_h2(a); b = gelu_erf_h2(b); }
      if (ACT == 2) { a = __hmax2(a, __float2half2_rn(0.f)); b = __hmax2(b, __float2half2_rn(0.f)); }
      a = __hmul2(a, om2);
      b = __hmul2(b, om2);
      pk[j] = ptx::pack_e4m3x2_h2(*reinterpret_cast<uint32_t*>(&a)) | (ptx::pack_e4m3x2_h2(*reinterpret_cast<uint32_t*>(&b)) << 16);
    }
    if (lane == 0) ptx::bulk_wait_group_read<NBUF - 1>();
    __syncwarp();
    uint8_t* sb = stage + buf * 2048;
    uint8_t* srow = sb + lane * 32;
    *reinterpret_cast<uint4*>(srow) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(srow + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_2d(tm_c, sb, n, row0);
      ptx::bulk_commit_group();
    }
    buf = (buf + 1 == NBUF) ? 0 : buf + 1;
  }
}

// SwiGLU with e4m3 operands AND e4m3 output (the decoder MLP of the FP8 tier): gate / up chunks scaled by the row scale and
// their own channel scales, silu(gate) * up * out_mult in packed fp16, saturating e4m3; tm_c: uint8 [M, N/2], box 32 x 32.
__device__ __forceinline__ __half2 silu_mul_h2(__half2 g, __half2 u) {
  const __half2 hg = __hmul2(g, __float2half2_rn(0.5f));
  uint32_t t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&hg)));
  return __hmul2(__hfma2(hg, *reinterpret_cast<const __half2*>(&t), hg), u);
}

template <int BN, int NBUF>
__device__ __forceinline__ void epilogue_tile_tma_swiglu_f8(uint32_t t_row, int grp, int n0, int row0, const float* __restrict__ bias,
                                                            const CUtensorMap* tm_c, uint8_t* stage, int& buf, int lane, int M, int N,
                                                            const Fp8Scales& sc) {
#pragma unroll 1
  for (int pc = grp; pc < BN / 64; pc += EpiCfg<3>::GROUPS) {
    const int n = n0 + pc * 64;
    uint32_t rg[32], ru[32];
    ptx::tmem_ld_x32(t_row + (uint32_t)(pc * 64), rg);
    ptx::tmem_ld_x32(t_row + (uint32_t)(pc * 64 + 32), ru);
    ptx::tmem_ld_wait();
    if (n + 64 > N) continue;   // warp-uniform (N is a multiple of 64 for this activation)
    const int row = row0 + lane;
    const float rs = (sc.row != nullptr && row < M) ? __ldg(sc.row + row) * sc.k : sc.k;
    const __half2 om2 = __float2half2_rn(sc.out_mult);
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 cg = __ldg(reinterpret_cast<const float4*>(sc.col + n) + j);
      const float4 cu = __ldg(reinterpret_cast<const float4*>(sc.col + n + 32) + j);
      float g[4] = {__uint_as_float(rg[4 * j]) * rs * cg.x, __uint_as_float(rg[4 * j + 1]) * rs * cg.y,
                    __uint_as_float(rg[4 * j + 2]) * rs * cg.z, __uint_as_float(rg[4 * j + 3]) * rs * cg.w};
      float u[4] = {__uint_as_float(ru[4 * j]) * rs * cu.x, __uint_as_float(ru[4 * j + 1]) * rs * cu.y,
                    __uint_as_float(ru[4 * j + 2]) * rs * cu.z, __uint_as_float(ru[4 * j + 3]) * rs * cu.w};
      if (bias != nullptr) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { g[k] += __ldg(bias + n + 4 * j + k); u[k] += __ldg(bias + n + 32 + 4 * j + k); }
      }
      uint32_t g01 = ptx::pack_f16(g[0], g[1]), g23 = ptx::pack_f16(g[2], g[3]);
      uint32_t u01 = ptx::pack_f16(u[0], u[1]), u23 = ptx::pack_f16(u[2], u[3]);
      __half2 a = __hmul2(silu_mul_h2(*reinterpret_cast<__half2*>(&g01), *reinterpret_cast<__half2*>(&u01)), om2);
      __half2 b = __hmul2(silu_mul_h2(*reinterpret_cast<__half2*>(&g23), *reinterpret_cast<__half2*>(&u23)), om2);
      pk[j] = ptx::pack_e4m3x2_h2(*reinterpret_cast<uint32_t*>(&a)) | (ptx::pack_e4m3x2_h2(*reinterpret_cast<uint32_t*>(&b)) << 16);
    }
    if (lane == 0) ptx::bulk_wait_group_read<NBUF - 1>();
    __syncwarp();
    uint8_t* sb = stage + buf * 2048;
    uint8_t* srow = sb + lane * 32;
    *reinterpret_cast<uint4*>(srow) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(srow + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_2d(tm_c, sb, n / 2, row0);
      ptx::bulk_commit_group();
    }
    buf = (buf + 1 == NBUF) ? 0 : buf + 1;
  }
}

// QKV projection of the decoder stack with the rotary embedding applied to the fp32 accumulators (transformers
// modeling_qwen2.py Q:60-80 rotate_half / apply_rotary_pos_emb): a head is 128 columns = four 32-column chunks, the pair
// (p, p + 64) sits in chunks c and c + 2 of the same row, and with two epilogue groups one warp owns both.  One rounding to
// fp16 instead of two (GEMM output, then the rotation), and no separate pass over the [M, (q + kv heads) * 128] slab.
template <int BN, int NBUF>
__device__ __forceinline__ void epilogue_tile_tma_rope(uint32_t t_row, int grp, int n0, int row0, const float* __restrict__ bias,
                                                       const CUtensorMap* tm_c, uint8_t* stage, int& buf, int lane, int M, int N,
                                                       const Fp8Scales& sc) {
  static_assert(EpiCfg<4>::GROUPS == 2 && BN % 128 == 0, "chunk pairing of the rotary epilogue");
  const int row = row0 + lane;
  int pos = 0;
  if (row < M) pos = sc.rope_tile_pos ? __ldg(sc.rope_tile_pos + (row >> 7)) + (row & 127) : row % sc.rope_S;
#pragma unroll 1
  for (int hd = 0; hd < BN / 128; ++hd) {
    const int c_lo = 4 * hd + grp, c_hi = c_lo + 2;
    const int n_lo = n0 + 32 * c_lo, n_hi = n0 + 32 * c_hi;
    uint32_t r_lo[32], r_hi[32];
    ptx::tmem_ld_x32(t_row + (uint32_t)(32 * c_lo), r_lo);
    ptx::tmem_ld_x32(t_row + (uint32_t)(32 * c_hi), r_hi);
    ptx::tmem_ld_wait();
    if (n_lo >= N) continue;   // warp-uniform (N is a multiple of 128 here)
    float lo[32], hi[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      lo[j] = __uint_as_float(r_lo[j]);
      hi[j] = __uint_as_float(r_hi[j]);
    }
    if (bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(bias + n_lo) + j), b = __ldg(reinterpret_cast<const float4*>(bias + n_hi) + j);
        lo[4 * j] += a.x; lo[4 * j + 1] += a.y; lo[4 * j + 2] += a.z; lo[4 * j + 3] += a.w;
        hi[4 * j] += b.x; hi[4 * j + 1] += b.y; hi[4 * j + 2] += b.z; hi[4 * j + 3] += b.w;
      }
    }
    if (n_lo < sc.rope_cols) {   // warp-uniform: a Q or K head (V heads pass through)
      const float4* t4 = reinterpret_cast<const float4*>(sc.rope_cs + (int64_t)pos * 64 + 32 * grp);   // two (cos, sin) per float4
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 t = __ldg(t4 + j);
        const float a0 = lo[2 * j], b0 = hi[2 * j], a1 = lo[2 * j + 1], b1 = hi[2 * j + 1];
        lo[2 * j] = a0 * t.x - b0 * t.y;          // x1' = x1 cos - x2 sin
        hi[2 * j] = b0 * t.x + a0 * t.y;          // x2' = x2 cos + x1 sin
        lo[2 * j + 1] = a1 * t.z - b1 * t.w;
        hi[2 * j + 1] = b1 * t.z + a1 * t.w;
      }
    }
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      const float* v = part == 0 ? lo : hi;
      if (lane == 0) ptx::bulk_wait_group_read<NBUF - 1>();
      __syncwarp();
      uint8_t* sb = stage + buf * 2048;
      uint8_t* srow = sb + lane * 64;
      const int sw = (lane >> 1) & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = ptx::pack_f16(v[8 * j + 0], v[8 * j + 1]);
        o.y = ptx::pack_f16(v[8 * j + 2], v[8 * j + 3]);
        o.z = ptx::pack_f16(v[8 * j + 4], v[8 * j + 5]);
        o.w = ptx::pack_f16(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = o;
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_2d(tm_c, sb, part == 0 ? n_lo : n_hi, row0);
        ptx::bulk_commit_group();
      }
      buf = (buf + 1 == NBUF) ? 0 : buf + 1;
    }
  }
}

// fp16 epilogue through shared memory + TMA store.  A warp's direct stores put 32 different rows in
// every STG (32 L1 line visits per instruction: the epilogue then outlasts a K=768 tile's MMAs); here
// the 32 x 32 chunk is written once to a 64B-swizzled staging buffer and stored by one bulk tensor copy.
//   stage: this warp's NBUF x 2 KB staging buffers (1024-byte aligned); row0: first tile row of the warp
template <int BN, int ACT, int NBUF, bool RESID, bool SCALED = false>
__device__ __forceinline__ void epilogue_tile_tma(uint32_t t_row, int grp, int n0, int row0, const float* __restrict__ bias,
                                                  const CUtensorMap* tm_c, uint8_t* stage, int& buf, int lane, int M, int N,
                                                  const __half* __restrict__ resid, int64_t ldr, const Fp8Scales& sc = Fp8Scales{}) {
#pragma unroll 1
  for (int c = grp; c < BN / 32; c += EpiCfg<ACT>::GROUPS) {
    const int n = n0 + c * 32;
    ResidRow rr;
    if (RESID && n < N) rr.load(resid, ldr, row0 + lane, n, M, N);
    uint32_t r[32];
    ptx::tmem_ld_x32(t_row + (uint32_t)(c * 32), r);
    ptx::tmem_ld_wait();
    if (n >= N) continue;   // warp-uniform
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    apply_scales<SCALED>(v, sc, row0 + lane, n, M, N);
    if (bias != nullptr) {
      if (n + 32 <= N) {
        const float4* b4 = reinterpret_cast<const float4*>(bias + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(b4 + j);
          v[4 * j + 0] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (n + j < N) v[j] += __ldg(bias + n + j);
      }
    }
    if (ACT == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    }
    if (ACT == 2) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (RESID) rr.add_to(v);
    // the bulk store that last read this staging buffer must have finished reading it
    if (lane == 0) ptx::bulk_wait_group_read<NBUF - 1>();
    __syncwarp();
    uint8_t* sb = stage + buf * 2048;
    // row = lane, 64 bytes per row; 64B swizzle: 16-byte chunk j lands at j ^ ((row >> 1) & 3)
    uint8_t* srow = sb + lane * 64;
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 o;
      o.x = ptx::pack_f16(v[8 * j + 0], v[8 * j + 1]);
      o.y = ptx::pack_f16(v[8 * j + 2], v[8 * j + 3]);
      o.z = ptx::pack_f16(v[8 * j + 4], v[8 * j + 5]);
      o.w = ptx::pack_f16(v[8 * j + 6], v[8 * j + 7]);
      *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = o;
    }
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_2d(tm_c, sb, n, row0);   // rows >= M and columns >= N are clipped by the tensor map
      ptx::bulk_commit_group();
    }
    buf = (buf + 1 == NBUF) ? 0 : buf + 1;
  }
}

template <int BN, int ACT, bool OUT_F32, bool RESID>
__global__ void __launch_bounds__(EpiCfg<ACT>::THREADS, 1)
gemm_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                         const float* __restrict__ bias, void* __restrict__ Cout, int64_t ldc, int M, int N, int K,
                         const __half* __restrict__ resid, int64_t ldr) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (128B-swizzle atoms) by pointer arithmetic on the __shared__ symbol: an
  // integer round-trip would turn every later access into a generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_w);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], 4 * EpiCfg<ACT>::GROUPS);   // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();                 // everything above overlapped the previous kernel's tail (launch.h)
  ptx::pdl_launch_dependents();

  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_k = (K + BK - 1) / BK;
  const int tiles = num_m * num_n;

  // Producer and MMA warps run their loops warp-uniformly (all lanes wait on the barriers) and elect
  // one lane per issue: tcgen05 / TMA instructions issued from inside `if (lane == 0)` get wrapped by
  // the compiler in an R2UR + ELECT retry loop that costs ~100 cycles per instruction.
  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * BM;
      const int n0 = (tile % num_n) * BN;
      for (int kb = 0; kb < num_k; ++kb) {
        ptx::mbar_wait(&empty[s], ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          ptx::tma_load_2d(sa, &tm_a, &full[s], kb * BK, m0);
          ptx::tma_load_2d(sa + Cfg::A_BYTES, &tm_w, &full[s], kb * BK, n0);
        }
        __syncwarp();
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = ptx::idesc_f16(BM, BN);
    const uint64_t desc0 = ptx::smem_desc_sw128(ptx::smem_u32(smem));
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tempty[acc], acc_ph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_k; ++kb) {
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          // descriptor address field is in 16-byte units: stage s starts s*STAGE_BYTES/16 further
          const uint64_t da = desc0 + (uint64_t)(s * (Cfg::STAGE_BYTES >> 4));
          const uint64_t db = da + (Cfg::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            ptx::mma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (uint32_t)((kb | k) != 0));
          ptx::mma_commit(&empty[s]);   // frees the smem stage once these MMAs have read it
          if (kb == num_k - 1) ptx::mma_commit(&tfull[acc]);   // accumulator complete
        }
        __syncwarp();
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1;
    }
  } else {
    const int q = warp & 3;   // TMEM lane quarter this warp may read
    const int grp = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * BM;
      const int n0 = (tile % num_n) * BN;
      const int row = m0 + q * 32 + lane;
      ptx::mbar_wait(&tfull[acc], acc_ph);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
      if (ACT == 3)
        epilogue_tile_swiglu<BN, OUT_F32>(t_row, grp, n0, row, bias, Cout, ldc, M, N);
      else
        epilogue_tile<BN, ACT, OUT_F32, RESID>(t_row, grp, n0, row, bias, Cout, ldc, M, N, resid, ldr);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant: 256 x 256 output tile per cluster of two CTAs (tcgen05 cta_group::2).
// Each CTA stages its own 128 rows of A and ITS HALF (128 N-rows) of the W tile, so a k-block costs
// 32 KB of L2->SM traffic per SM instead of 48 KB and the smem ring is six deep instead of four:
// the single-CTA kernel keeps the tensor pipe only 45-60 % busy because operand delivery, not the
// MMA, paces it.  The leader CTA (cluster rank 0) issues the MMAs for both; every CTA runs its own
// TMA producer and its own epilogue on its 128 TMEM lanes.
//   full[s]   (leader's copy)  <- transaction bytes of both CTAs' loads of stage s
//   empty[s]  (both copies)    <- multicast tcgen05.commit: the stage may be refilled
//   tfull[a]  (both copies)    <- multicast tcgen05.commit: accumulator a complete
//   tempty[a] (leader's copy)  <- epilogue warps of both CTAs
// ---------------------------------------------------------------------------------------------
template <int BN_>
struct Gemm2Cfg {
  static constexpr int BN = BN_;            // cluster tile N (256, or 192 where that quantises better); per CTA B half = BN/2 rows
  static constexpr int STAGES = 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = 32768;   // TMA-store staging: 2 KB buffers, split over the epilogue warps
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 + 1024;
  static constexpr uint32_t TMEM_COLS = 512;   // two BN-column accumulators
};

// FP8 = true: A and W are e4m3 bytes (tensor maps over uint8, 128 elements per 128-byte stage row, kind::f8f6f4); the
// epilogue multiplies by the scales in `sc`.  OUT_F8: the output is e4m3 too (see epilogue_tile_tma_f8out).
template <int BN_, int ACT, bool OUT_F32, bool RESID, bool FP8 = false, bool OUT_F8 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EpiCfg<ACT>::THREADS, 1)
gemm_f16_2cta_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                     const __grid_constant__ CUtensorMap tm_c, const float* __restrict__ bias, void* __restrict__ Cout,
                     int64_t ldc, int M, int N, int K, const __half* __restrict__ resid, int64_t ldr, const Fp8Scales sc) {
  constexpr int BKE = FP8 ? 2 * BK : BK;   // elements per 128-byte stage row
  using Cfg = Gemm2Cfg<BN_>;
  constexpr int BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* epi_stage = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + Cfg::EPI_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = (rank == 0);

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_w);
    if (!OUT_F32) ptx::prefetch_tensormap(&tm_c);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], 2 * 4 * EpiCfg<ACT>::GROUPS);   // epilogue warps of both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync();   // barrier inits and the TMEM allocation of both CTAs are visible
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();                 // everything above overlapped the previous kernel's tail (launch.h)
  ptx::pdl_launch_dependents();

  const int num_m = (M + 2 * BM - 1) / (2 * BM);
  const int num_n = (N + BN - 1) / BN;
  const int num_k = (K + BKE - 1) / BKE;
  const int tiles = num_m * num_n;
  const int cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int tile = cl; tile < tiles; tile += ncl) {
      const int m0 = (tile / num_n) * (2 * BM) + (int)rank * BM;
      const int n0 = (tile % num_n) * BN + (int)rank * (BN / 2);
      for (int kb = 0; kb < num_k; ++kb) {
        ptx::mbar_wait(&empty[s], ph ^ 1);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_arrive_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
          const uint32_t fb = ptx::smem_u32(&full[s]) & ptx::PEER_BIT_MASK;   // the leader's barrier
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          ptx::tma_load_2d_2cta(sa, &tm_a, fb, kb * BKE, m0);
          ptx::tma_load_2d_2cta(sa + Cfg::A_BYTES, &tm_w, fb, kb * BKE, n0);
        }
        __syncwarp();
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = ptx::idesc_f16(2 * BM, BN);
      const uint64_t desc0 = ptx::smem_desc_sw128(ptx::smem_u32(smem));
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (int tile = cl; tile < tiles; tile += ncl) {
        ptx::mbar_wait(&tempty[acc], acc_ph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_k; ++kb) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint64_t da = desc0 + (uint64_t)(s * (Cfg::STAGE_BYTES >> 4));
            const uint64_t db = da + (Cfg::A_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if (FP8) ptx::mma_e4m3_ss_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (uint32_t)((kb | k) != 0));
              else ptx::mma_f16_ss_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (uint32_t)((kb | k) != 0));
            }
            ptx::mma_commit_2cta_mc(&empty[s], 3);
            if (kb == num_k - 1) ptx::mma_commit_2cta_mc(&tfull[acc], 3);
          }
          __syncwarp();
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    constexpr int NBUF = Cfg::EPI_BYTES / (4 * EpiCfg<ACT>::GROUPS) / 2048;   // staging buffers per epilogue warp
    uint8_t* my_stage = epi_stage + (warp - 2) * (NBUF * 2048);
    int sbuf = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = cl; tile < tiles; tile += ncl) {
      const int m0 = (tile / num_n) * (2 * BM) + (int)rank * BM;
      const int n0 = (tile % num_n) * BN;
      const int row = m0 + q * 32 + lane;
      ptx::mbar_wait(&tfull[acc], acc_ph);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
      if constexpr (ACT == 4)
        epilogue_tile_tma_rope<BN, NBUF>(t_row, grp, n0, m0 + q * 32, bias, &tm_c, my_stage, sbuf, lane, M, N, sc);
      else if (OUT_F8 && ACT == 3)
        epilogue_tile_tma_swiglu_f8<BN, NBUF>(t_row, grp, n0, m0 + q * 32, bias, &tm_c, my_stage, sbuf, lane, M, N, sc);
      else if (OUT_F8)
        epilogue_tile_tma_f8out<BN, ACT, NBUF>(t_row, grp, n0, m0 + q * 32, bias, &tm_c, my_stage, sbuf, lane, M, N, sc);
      else if (ACT == 3 && OUT_F32)
        epilogue_tile_swiglu<BN, true>(t_row, grp, n0, row, bias, Cout, ldc, M, N);
      else if (ACT == 3)
        epilogue_tile_tma_swiglu<BN, NBUF>(t_row, grp, n0, m0 + q * 32, bias, &tm_c, my_stage, sbuf, lane, N);
      else if (OUT_F32)
        epilogue_tile<BN, ACT, OUT_F32, RESID>(t_row, grp, n0, row, bias, Cout, ldc, M, N, resid, ldr);
      else
        epilogue_tile_tma<BN, ACT, NBUF, RESID, FP8>(t_row, grp, n0, m0 + q * 32, bias, &tm_c, my_stage, sbuf, lane, M, N, resid, ldr, sc);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::smem_u32(&tempty[acc]) & ptx::PEER_BIT_MASK);
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1;
    }
    if (!OUT_F32 && lane == 0) ptx::bulk_wait_group_read<0>();   // staging smem must outlive its readers
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();   // nobody leaves (or frees TMEM) while the peer may still signal or read
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN_, int ACT, bool OUT_F32>
cudaError_t launch_gemm_2cta(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc,
                             int M, int N, int K, int num_sms, cudaStream_t stream, const void* resid, int64_t ldr,
                             const Fp8Scales& extra = Fp8Scales{}) {
  using Cfg = Gemm2Cfg<BN_>;
  uint64_t da[2] = {(uint64_t)K, (uint64_t)M};
  uint64_t sa[1] = {(uint64_t)lda * 2};
  uint32_t ba[2] = {BK, BM};
  uint64_t dw[2] = {(uint64_t)K, (uint64_t)N};
  uint64_t sw[1] = {(uint64_t)ldw * 2};
  uint32_t bw[2] = {BK, (uint32_t)(Cfg::BN / 2)};
  CUtensorMap tm_a = make_tmap_16b(A, 2, da, sa, ba);
  CUtensorMap tm_w = make_tmap_16b(W, 2, dw, sw, bw);
  CUtensorMap tm_c = tm_a;   // unused by the fp32-output instantiation
  if (!OUT_F32) {
    uint64_t dc[2] = {(uint64_t)(ACT == 3 ? N / 2 : N), (uint64_t)M};
    uint64_t sc[1] = {(uint64_t)ldc * 2};
    uint32_t bc[2] = {32, 32};
    tm_c = make_tmap_16b(C, 2, dc, sc, bc, CU_TENSOR_MAP_SWIZZLE_64B);
  }
  auto kern = resid ? gemm_f16_2cta_kernel<BN_, ACT, OUT_F32, true> : gemm_f16_2cta_kernel<BN_, ACT, OUT_F32, false>;
  static bool attr_set[64][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63][resid ? 1 : 0]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63][resid ? 1 : 0] = true;
  }
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + Cfg::BN - 1) / Cfg::BN);
  const int max_cl = num_sms / 2;
  const int ncl = tiles < max_cl ? tiles : max_cl;
  return launch_pdl(kern, dim3(2 * ncl), dim3(EpiCfg<ACT>::THREADS), Cfg::SMEM_BYTES, stream, tm_a, tm_w, tm_c, bias, C, ldc, M, N, K,
                    (const __half*)resid, ldr, extra);
}

// e4m3 x e4m3 -> fp16 (OUT_F8 = false) or e4m3 (OUT_F8 = true), always on CTA pairs
template <int BN_, int ACT, bool OUT_F8>
cudaError_t launch_gemm_e4m3(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc,
                             int M, int N, int K, int num_sms, cudaStream_t stream, const Fp8Scales& sc) {
  using Cfg = Gemm2Cfg<BN_>;
  uint64_t da[2] = {(uint64_t)K, (uint64_t)M};
  uint64_t sa[1] = {(uint64_t)lda};
  uint32_t ba[2] = {2 * BK, BM};
  uint64_t dw[2] = {(uint64_t)K, (uint64_t)N};
  uint64_t sw[1] = {(uint64_t)ldw};
  uint32_t bw[2] = {2 * BK, (uint32_t)(Cfg::BN / 2)};
  CUtensorMap tm_a = make_tmap_8b(A, 2, da, sa, ba, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap tm_w = make_tmap_8b(W, 2, dw, sw, bw, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap tm_c;
  uint64_t dc[2] = {(uint64_t)(ACT == 3 ? N / 2 : N), (uint64_t)M};
  uint32_t bc[2] = {32, 32};
  if (OUT_F8) {
    uint64_t sc1[1] = {(uint64_t)ldc};
    tm_c = make_tmap_8b(C, 2, dc, sc1, bc, CU_TENSOR_MAP_SWIZZLE_NONE);
  } else {
    uint64_t sc1[1] = {(uint64_t)ldc * 2};
    tm_c = make_tmap_16b(C, 2, dc, sc1, bc, CU_TENSOR_MAP_SWIZZLE_64B);
  }
  auto kern = gemm_f16_2cta_kernel<BN_, ACT, false, false, true, OUT_F8>;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + Cfg::BN - 1) / Cfg::BN);
  const int max_cl = num_sms / 2;
  const int ncl = tiles < max_cl ? tiles : max_cl;
  return launch_pdl(kern, dim3(2 * ncl), dim3(EpiCfg<ACT>::THREADS), Cfg::SMEM_BYTES, stream, tm_a, tm_w, tm_c, bias, C, ldc, M, N, K,
                    (const __half*)nullptr, (int64_t)0, sc);
}

template <int BN, int ACT, bool OUT_F32>
cudaError_t launch_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc,
                        int M, int N, int K, int num_sms, cudaStream_t stream, const void* resid, int64_t ldr) {
  using Cfg = GemmCfg<BN>;
  uint64_t da[2] = {(uint64_t)K, (uint64_t)M};
  uint64_t sa[1] = {(uint64_t)lda * 2};
  uint32_t ba[2] = {BK, BM};
  uint64_t dw[2] = {(uint64_t)K, (uint64_t)N};
  uint64_t sw[1] = {(uint64_t)ldw * 2};
  uint32_t bw[2] = {BK, BN};
  CUtensorMap tm_a = make_tmap_16b(A, 2, da, sa, ba);
  CUtensorMap tm_w = make_tmap_16b(W, 2, dw, sw, bw);
  auto kern = resid ? gemm_f16_tcgen05_kernel<BN, ACT, OUT_F32, true> : gemm_f16_tcgen05_kernel<BN, ACT, OUT_F32, false>;
  static bool attr_set[64][2] = {};   // per template instantiation, per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63][resid ? 1 : 0]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63][resid ? 1 : 0] = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  return launch_pdl(kern, dim3(grid), dim3(EpiCfg<ACT>::THREADS), Cfg::SMEM_BYTES, stream, tm_a, tm_w, bias, C, ldc, M, N, K,
                    (const __half*)resid, ldr);
}

}  // namespace

cudaError_t gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M,
                     int N, int K, int act, bool out_f32, int num_sms, cudaStream_t stream) {
  return gemm_f16_resid(A, lda, W, ldw, bias, nullptr, 0, C, ldc, M, N, K, act, out_f32, num_sms, stream);
}

cudaError_t gemm_f16_resid(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* resid,
                           int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int act, bool out_f32, int num_sms,
                           cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaErrorInvalidValue;
  if (resid != nullptr && ((ldr % 8) || (reinterpret_cast<uintptr_t>(resid) & 15))) return cudaErrorInvalidValue;
  if ((lda % 8) || (ldw % 8) || (K % 8) || (N % 8) || (ldc % (out_f32 ? 4 : 8))) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C)) & 15)
    return cudaErrorInvalidValue;
  if (act < 0 || act > 3) return cudaErrorInvalidValue;
  if (act == 3) {
    // SwiGLU: gate / up rows interleaved in blocks of 32 (see silu_mul); output [M, N/2]
    if (N % 64 != 0 || resid != nullptr || out_f32) return cudaErrorInvalidValue;
    const int t2 = ((M + 255) / 256) * ((N + 255) / 256);
    if (t2 >= num_sms / 2 && N >= 256) return launch_gemm_2cta<256, 3, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, nullptr, 0);
    return launch_gemm<128, 3, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, nullptr, 0);
  }
  if (out_f32) {
    if (act == 0) return launch_gemm<128, 0, true>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
    if (act == 1) return launch_gemm<128, 1, true>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
    return launch_gemm<128, 2, true>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
  }
  // Tile configuration by a wave model: every candidate keeps 128 output rows per SM, so its time is (rounds of tiles over
  // the SMs / clusters) x (tile columns); single-CTA tiles pay ~10 % for fetching whole W tiles per SM (see the CTA-pair
  // kernel's header), the 192-wide pair tile 2 % for re-reading A more often.  Examples on 148 SMs: M = 32768, N = 768 ->
  // 256 x 192 pairs in 7 rounds (6 rounds of 256 cost 14 % more); M = 4096 (one batch-8 request), N = 768 -> 64 pair
  // tiles of 256 x 192 in ONE round instead of two rounds of 128 x 128 (-25 % on the K = 3072 FFN2).
  static const bool no_pairs = getenv("GLC_GEMM_NO_PAIRS") != nullptr;
  static const bool no_192 = getenv("GLC_GEMM_NO_192") != nullptr;
  const int ncl = num_sms / 2, mt1 = (M + BM - 1) / BM, mt2 = (M + 2 * BM - 1) / (2 * BM);
  auto rounds = [](int tiles, int slots) { return (int64_t)((tiles + slots - 1) / slots); };
  const int64_t c128 = rounds(mt1 * ((N + 127) / 128), num_sms) * 128 * 110;
  const int64_t c256 = (N % 256 == 0) ? rounds(mt1 * (N / 256), num_sms) * 256 * 110 : INT64_MAX;
  const int64_t p256 = (!no_pairs && N >= 256 && ncl > 0) ? rounds(mt2 * ((N + 255) / 256), ncl) * 256 * 100 : INT64_MAX;
  const int64_t p192 = (!no_pairs && !no_192 && act == 0 && N % 192 == 0 && ncl > 0) ? rounds(mt2 * (N / 192), ncl) * 192 * 102 : INT64_MAX;
  const int64_t best = std::min(std::min(c128, c256), std::min(p256, p192));
  if (best == p256) {
    if (act == 0) return launch_gemm_2cta<256, 0, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
    if (act == 1) return launch_gemm_2cta<256, 1, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
    return launch_gemm_2cta<256, 2, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
  }
  if (best == p192) return launch_gemm_2cta<192, 0, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
  if (best == c256) {
    if (act == 0) return launch_gemm<256, 0, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
    if (act == 1) return launch_gemm<256, 1, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
    return launch_gemm<256, 2, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
  }
  if (act == 0) return launch_gemm<128, 0, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
  if (act == 1) return launch_gemm<128, 1, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
  return launch_gemm<128, 2, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, resid, ldr);
}

cudaError_t gemm_f16_rope(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N,
                          int K, const void* rope_cs_f32x2, int rope_S, const int32_t* rope_tile_pos, int rope_cols, int num_sms,
                          cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || rope_cs_f32x2 == nullptr || rope_S <= 0) return cudaErrorInvalidValue;
  if ((lda % 8) || (ldw % 8) || (K % 8) || (N % 128) || (rope_cols % 128) || rope_cols > N || (ldc % 8) || num_sms < 2) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C)) & 15) return cudaErrorInvalidValue;
  Fp8Scales ex{};
  ex.rope_cs = (const float2*)rope_cs_f32x2;
  ex.rope_tile_pos = rope_tile_pos;
  ex.rope_S = rope_S;
  ex.rope_cols = rope_cols;
  return launch_gemm_2cta<256, 4, false>(A, lda, W, ldw, bias, C, ldc, M, N, K, num_sms, stream, nullptr, 0, ex);
}

cudaError_t gemm_e4m3(const void* A8, int64_t lda, const void* W8, int64_t ldw, const float* a_scale, float a_const,
                      const float* w_scale, const float* bias, void* C, int64_t ldc, int M, int N, int K, int act, bool out_e4m3,
                      float out_mult, int num_sms, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || w_scale == nullptr) return cudaErrorInvalidValue;
  if ((lda % 16) || (ldw % 16) || (K % 16) || (N % 16) || (ldc % (out_e4m3 ? 16 : 8))) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(A8) | reinterpret_cast<uintptr_t>(W8) | reinterpret_cast<uintptr_t>(C)) & 15)
    return cudaErrorInvalidValue;
  if (num_sms < 2) return cudaErrorInvalidValue;
  const Fp8Scales sc{a_scale, w_scale, a_const, out_mult};
  if (out_e4m3) {
    if (act == 3) {   // SwiGLU: gate / up rows interleaved in blocks of 32, output [M, N/2]
      if (N % 64 != 0) return cudaErrorInvalidValue;
      return launch_gemm_e4m3<256, 3, true>(A8, lda, W8, ldw, bias, C, ldc, M, N, K, num_sms, stream, sc);
    }
    if (act == 1) return launch_gemm_e4m3<256, 1, true>(A8, lda, W8, ldw, bias, C, ldc, M, N, K, num_sms, stream, sc);
    if (act == 0) return launch_gemm_e4m3<256, 0, true>(A8, lda, W8, ldw, bias, C, ldc, M, N, K, num_sms, stream, sc);
    return cudaErrorInvalidValue;
  }
  if (act != 0) return cudaErrorInvalidValue;
  // same wave model as the fp16 pairs: 192-wide tiles when they quantise into clearly fewer column-rounds
  const int ncl = num_sms / 2, mt2 = (M + 2 * BM - 1) / (2 * BM);
  const int64_t p256 = (int64_t)((mt2 * ((N + 255) / 256) + ncl - 1) / ncl) * 256 * 100;
  const int64_t p192 = (N % 192 == 0) ? (int64_t)((mt2 * (N / 192) + ncl - 1) / ncl) * 192 * 102 : INT64_MAX;
  if (p192 < p256) return launch_gemm_e4m3<192, 0, false>(A8, lda, W8, ldw, bias, C, ldc, M, N, K, num_sms, stream, sc);
  return launch_gemm_e4m3<256, 0, false>(A8, lda, W8, ldw, bias, C, ldc, M, N, K, num_sms, stream, sc);
}

}  // namespace glc
