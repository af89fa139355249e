// HBM-bound kernels of the hot path: K1 embedding gather + LayerNorm + mask, K4 residual +
// LayerNorm, attention-mask bit packing, K5a <<LABEL>> pooling and K5b dot scorer (+sigmoid /
// threshold).  One warp per row, 128-bit loads/stores, fp32 statistics via warp shuffles.
//
// Arithmetic restated from transformers' modeling_deberta_v2.py (T:) and SURVEY.md App. B:
//   K1  T:520-564   e = LN(word_emb[ids]) * mask           (no position / token-type term in v3)
//   K4  T:49-53, T:408-412   y = LN(dense_out + residual)  (eps 1e-7, biased variance)
//   K5  App. B      cls[b,c] = h[b,pos_c]; pooled = h[b,0]; logit = <t_b, k_bc>
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace glc {
namespace {

constexpr int ROWS_PER_BLOCK = 8;   // 8 warps / block, one row per warp
constexpr int MAXC = 8;             // up to 8 chunks of 8 elements per lane -> H <= 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// read-once 128-bit load that does not linger in L1
__device__ __forceinline__ uint4 ld_stream(const __half* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __half2* p = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// Saturation probe for fp16-stored pre-LN sums: the GEMM epilogue packs with cvt.rn.satfinite, so a value beyond the
// fp16 range arrives here as +-65504 instead of inf.  Any |x + r| >= 6e4 raises the engine's overflow flag (checked by
// the host after the forward: the Run fails loudly instead of returning logits computed from clamped activations).
constexpr float SAT_LIMIT = 6.0e4f;
template <int NC>
__device__ __forceinline__ void sat_probe(const float (&v)[NC][8], int lane, int H, int* __restrict__ flag) {
  if (flag == nullptr) return;
  float amax = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if ((lane + 32 * c) * 8 < H) {
#pragma unroll
      for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(v[c][i]));
    }
  if (!(amax < SAT_LIMIT)) *flag = 1;   // also catches NaN
}

// gamma / beta of this lane's chunks, fetched up front so that their (L2) latency overlaps the row loads
template <int NC>
struct LnParams {
  float g[NC][8], b[NC][8];
  __device__ __forceinline__ void load(int lane, int H, const float* __restrict__ gamma, const float* __restrict__ beta) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int e0 = (lane + 32 * c) * 8;
      if (e0 < H) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + e0));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + e0 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + e0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + e0 + 4));
        g[c][0] = g0.x; g[c][1] = g0.y; g[c][2] = g0.z; g[c][3] = g0.w;
        g[c][4] = g1.x; g[c][5] = g1.y; g[c][6] = g1.z; g[c][7] = g1.w;
        b[c][0] = b0.x; b[c][1] = b0.y; b[c][2] = b0.z; b[c][3] = b0.w;
        b[c][4] = b1.x; b[c][5] = b1.y; b[c][6] = b1.z; b[c][7] = b1.w;
      }
    }
  }
};

// LN over a row held as NC chunks of 8 floats per lane; writes fp16, scaled by `post`.  With out8 != nullptr the row is
// ALSO written as e4m3 under its own dynamic scale (y * 448 / amax_row; *scale_out = amax_row / 448): the A operand of the
// opt-in FP8 FFN1 GEMM (gemm_e4m3), produced here so no separate quantiser pass re-reads the activations.
template <int NC>
__device__ __forceinline__ void ln_store(float (&v)[NC][8], int lane, int H, const LnParams<NC>& gb, float eps, float post,
                                         __half* __restrict__ out, uint8_t* __restrict__ out8 = nullptr,
                                         float* __restrict__ scale_out = nullptr) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if ((lane + 32 * c) * 8 < H) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[c][i];
    }
  const float mean = warp_sum(s) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if ((lane + 32 * c) * 8 < H) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[c][i] - mean;
        q += d * d;
      }
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)H + eps);
  float amax = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int e0 = (lane + 32 * c) * 8;
    if (e0 < H) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[c][i] = ((v[c][i] - mean) * rstd * gb.g[c][i] + gb.b[c][i]) * post;
        amax = fmaxf(amax, fabsf(v[c][i]));
      }
      uint4 o;
      o.x = ptx::pack_f16(v[c][0], v[c][1]);
      o.y = ptx::pack_f16(v[c][2], v[c][3]);
      o.z = ptx::pack_f16(v[c][4], v[c][5]);
      o.w = ptx::pack_f16(v[c][6], v[c][7]);
      *reinterpret_cast<uint4*>(out + e0) = o;
    }
  }
  if (out8 != nullptr) {   // warp-uniform
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float inv = amax > 0.f ? 448.0f / amax : 1.0f;
    if (lane == 0) *scale_out = amax > 0.f ? amax / 448.0f : 1.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int e0 = (lane + 32 * c) * 8;
      if (e0 < H) {
        uint2 o;
        o.x = ptx::pack_e4m3x4(v[c][0] * inv, v[c][1] * inv, v[c][2] * inv, v[c][3] * inv);
        o.y = ptx::pack_e4m3x4(v[c][4] * inv, v[c][5] * inv, v[c][6] * inv, v[c][7] * inv);
        *reinterpret_cast<uint2*>(out8 + e0) = o;
      }
    }
  }
}

// ln_store for rows of exactly NC * 256 elements (every lane owns NC full chunks: no guards) — the hot case (H = 768 /
// 1024).  Same arithmetic and summation ORDER per lane as ln_store would not be needed for parity, but batch-composition
// invariance needs one order everywhere, so this is the only LN the bulk kernel uses for such rows.  Four independent
// partial sums per statistic (the single chain of ln_store is 24 dependent FADDs), d = x - mean kept in place and reused
// by the normalisation: 120 FP32 instructions per lane and row instead of 168.
template <int NC>
__device__ __forceinline__ void ln_store_full(float (&v)[NC][8], int lane, const LnParams<NC>& gb, float eps,
                                              __half* __restrict__ out, uint8_t* __restrict__ out8, float* __restrict__ scale_out) {
  constexpr float inv_h = 1.0f / (float)(NC * 256);
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) s4[i & 3] += v[c][i];
  const float mean = warp_sum((s4[0] + s4[1]) + (s4[2] + s4[3])) * inv_h;
  float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[c][i] -= mean;
      q4[i & 3] = fmaf(v[c][i], v[c][i], q4[i & 3]);
    }
  const float rstd = rsqrtf(warp_sum((q4[0] + q4[1]) + (q4[2] + q4[3])) * inv_h + eps);
  float amax = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[c][i] = fmaf(v[c][i] * rstd, gb.g[c][i], gb.b[c][i]);
    uint4 o;
    o.x = ptx::pack_f16(v[c][0], v[c][1]);
    o.y = ptx::pack_f16(v[c][2], v[c][3]);
    o.z = ptx::pack_f16(v[c][4], v[c][5]);
    o.w = ptx::pack_f16(v[c][6], v[c][7]);
    *reinterpret_cast<uint4*>(out + (lane + 32 * c) * 8) = o;
  }
  if (out8 != nullptr) {   // warp-uniform
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(v[c][i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float inv = amax > 0.f ? 448.0f / amax : 1.0f;
    if (lane == 0) *scale_out = amax > 0.f ? amax / 448.0f : 1.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      uint2 o;
      o.x = ptx::pack_e4m3x4(v[c][0] * inv, v[c][1] * inv, v[c][2] * inv, v[c][3] * inv);
      o.y = ptx::pack_e4m3x4(v[c][4] * inv, v[c][5] * inv, v[c][6] * inv, v[c][7] * inv);
      *reinterpret_cast<uint2*>(out8 + (lane + 32 * c) * 8) = o;
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
embed_ln_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ mask,
                const __half* __restrict__ emb, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, __half* __restrict__ y, int M, int H, int vocab) {
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  if (row >= M) return;
  int64_t id = ids[row];
  if (id < 0 || id >= vocab) id = 0;   // ORT's Gather would fail; clamp to [PAD] instead of reading out of bounds
  const float post = mask[row] != 0 ? 1.0f : 0.0f;
  const __half* src = emb + id * (int64_t)H;
  uint4 raw[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int e0 = (lane + 32 * c) * 8;
    if (e0 < H) raw[c] = __ldg(reinterpret_cast<const uint4*>(src + e0));
  }
  LnParams<NC> gb;
  gb.load(lane, H, gamma, beta);
  float v[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if ((lane + 32 * c) * 8 < H) unpack8(raw[c], v[c]);
  ln_store<NC>(v, lane, H, gb, eps, post, y + (int64_t)row * H);
}

// Each warp walks rows (row = global warp id, += total warps): gamma / beta stay in registers and the
// loads of the next row are issued before the statistics of the current one (two rows in flight).
template <int NC>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
residual_ln_kernel(const __half* __restrict__ x, const __half* __restrict__ r,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   __half* __restrict__ y, int M, int H, int* __restrict__ flag, uint8_t* __restrict__ y8,
                   float* __restrict__ y8_scale) {
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * ROWS_PER_BLOCK;
  int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  if (row >= M) return;
  uint4 xa[NC], ra[NC];
  auto fetch = [&](int rw) {
    const __half* xs = x + (int64_t)rw * H;
    const __half* rs = r + (int64_t)rw * H;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int e0 = (lane + 32 * c) * 8;
      if (e0 < H) {
        xa[c] = ld_stream(xs + e0);
        if (r) ra[c] = ld_stream(rs + e0);
      }
    }
  };
  fetch(row);
  LnParams<NC> gb;
  gb.load(lane, H, gamma, beta);
  while (true) {
    float v[NC][8];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if ((lane + 32 * c) * 8 < H) {
        unpack8(xa[c], v[c]);
        if (r) {
          float t[8];
          unpack8(ra[c], t);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[c][i] += t[i];
        }
      }
    }
    const int cur = row;
    row += stride;
    const bool more = row < M;   // warp-uniform
    if (more) fetch(row);
    sat_probe<NC>(v, lane, H, flag);
    ln_store<NC>(v, lane, H, gb, eps, 1.0f, y + (int64_t)cur * H, y8 ? y8 + (int64_t)cur * H : nullptr, y8_scale + cur);
    if (!more) break;
  }
}

// K4 with the pre-LN sum kept in fp32 (GLC_PRELN_F32=1, the robust mode for checkpoints whose dense outputs leave the
// fp16 range): x fp32 [M,H] straight from the GEMM's fp32 epilogue, r fp16.  One row per warp, no prefetch ring — this
// mode trades speed for range.
template <int NC>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
residual_ln_xf32_kernel(const float* __restrict__ x, const __half* __restrict__ r, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, __half* __restrict__ y, int M, int H) {
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * ROWS_PER_BLOCK;
  LnParams<NC> gb;
  gb.load(lane, H, gamma, beta);
  for (int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5); row < M; row += stride) {
    const float* xs = x + (int64_t)row * H;
    float v[NC][8];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int e0 = (lane + 32 * c) * 8;
      if (e0 < H) {
        const float4 a = *reinterpret_cast<const float4*>(xs + e0);
        const float4 b = *reinterpret_cast<const float4*>(xs + e0 + 4);
        v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w;
        v[c][4] = b.x; v[c][5] = b.y; v[c][6] = b.z; v[c][7] = b.w;
        if (r) {
          float t[8];
          unpack8(ld_stream(r + (int64_t)row * H + e0), t);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[c][i] += t[i];
        }
      }
    }
    ln_store<NC>(v, lane, H, gb, eps, 1.0f, y + (int64_t)row * H);
  }
}

// K4, bulk-copy variant: every warp owns a ring of LN_STAGES shared-memory stages, each holding one row of x and the
// matching row of r, filled by cp.async.bulk (one elected lane, completion on the stage's mbarrier).  LN_STAGES rows
// per warp are in flight (2 blocks x 8 warps x 4 stages x 3 KB = 192 KB per SM at H = 768) without holding them in
// registers, so the loads of a warp never wait for its arithmetic.  No block-level synchronisation: each warp runs its
// own pipeline over rows  warp_global, warp_global + total_warps, ...
constexpr int LN_STAGES = 4;   // at most; rows wider than 1.7 KB get fewer stages (ring <= 110 KB per block, two blocks per SM)

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ptx::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

// FAST: H == NC * 256 and a residual operand (every layer LN of the base / large stacks): unguarded chunk loops and
// ln_store_full.  The generic instantiation handles ragged widths and the no-residual form.
template <int NC, bool FAST>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
residual_ln_bulk_kernel(const __half* __restrict__ x, const __half* __restrict__ r, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, __half* __restrict__ y, int M, int H, int stages,
                        int* __restrict__ flag, uint8_t* __restrict__ y8, float* __restrict__ y8_scale) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)H * 2u;
  const bool has_r = r != nullptr;   // without a residual operand (the GEMM epilogue already added it) a stage is one row
  const uint32_t st_bytes = has_r ? 2 * row_bytes : row_bytes;
  uint8_t* ring = ln_smem + (size_t)warp * stages * st_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + (size_t)ROWS_PER_BLOCK * stages * st_bytes) + warp * LN_STAGES;
  const int nwarps = gridDim.x * ROWS_PER_BLOCK;
  const int row0 = blockIdx.x * ROWS_PER_BLOCK + warp;
  if (row0 >= M) return;
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) ptx::mbar_init(&bars[s], 1);
    ptx::fence_barrier_init();
  }
  __syncwarp();
  ptx::pdl_wait();                 // barrier set-up overlapped the previous kernel's tail (launch.h)
  ptx::pdl_launch_dependents();
  auto issue = [&](int rw, int s) {
    if (lane == 0) {
      uint8_t* dst = ring + (size_t)s * st_bytes;
      ptx::mbar_arrive_expect_tx(&bars[s], st_bytes);
      bulk_load(dst, x + (int64_t)rw * H, row_bytes, &bars[s]);
      if (has_r) bulk_load(dst + row_bytes, r + (int64_t)rw * H, row_bytes, &bars[s]);
    }
  };
  for (int s = 0; s < stages; ++s) {
    const int rw = row0 + s * nwarps;
    if (rw < M) issue(rw, s);
  }
  LnParams<NC> gb;
  gb.load(lane, H, gamma, beta);
  int s = 0;
  uint32_t parity = 0;
  for (int rw = row0; rw < M; rw += nwarps) {
    ptx::mbar_wait(&bars[s], parity);
    const uint8_t* xs = ring + (size_t)s * st_bytes;
    float v[NC][8];
    if (FAST) {
      uint4 xa[NC], ra[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        xa[c] = *reinterpret_cast<const uint4*>(xs + (lane + 32 * c) * 16);
        ra[c] = *reinterpret_cast<const uint4*>(xs + row_bytes + (lane + 32 * c) * 16);
      }
      __syncwarp();   // every lane has read the stage: refill it
      const int nxt = rw + stages * nwarps;
      if (nxt < M) issue(nxt, s);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float t[8];
        unpack8(xa[c], v[c]);
        unpack8(ra[c], t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[c][i] += t[i];
      }
      sat_probe<NC>(v, lane, NC * 256, flag);
      ln_store_full<NC>(v, lane, gb, eps, y + (int64_t)rw * H, y8 ? y8 + (int64_t)rw * H : nullptr, y8_scale + rw);
    } else {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int e0 = (lane + 32 * c) * 8;
      if (e0 < H) {
        const uint4 xa = *reinterpret_cast<const uint4*>(xs + e0 * 2);
        unpack8(xa, v[c]);
        if (has_r) {
          const uint4 ra = *reinterpret_cast<const uint4*>(xs + row_bytes + e0 * 2);
          float t[8];
          unpack8(ra, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[c][i] += t[i];
        }
      }
    }
    __syncwarp();   // every lane has read the stage: refill it
    const int nxt = rw + stages * nwarps;
    if (nxt < M) issue(nxt, s);
    sat_probe<NC>(v, lane, H, flag);
    ln_store<NC>(v, lane, H, gb, eps, 1.0f, y + (int64_t)rw * H, y8 ? y8 + (int64_t)rw * H : nullptr, y8_scale + rw);
    }
    if (++s == stages) { s = 0; parity ^= 1u; }
  }
}

template <int NC>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
ln_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
              __half* __restrict__ y, int M, int H) {
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xs = x + (int64_t)row * H;
  float v[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int e0 = (lane + 32 * c) * 8;
    if (e0 < H) {
      const float4 a = *reinterpret_cast<const float4*>(xs + e0);
      const float4 b = *reinterpret_cast<const float4*>(xs + e0 + 4);
      v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w;
      v[c][4] = b.x; v[c][5] = b.y; v[c][6] = b.z; v[c][7] = b.w;
    }
  }
  LnParams<NC> gb;
  gb.load(lane, H, gamma, beta);
  ln_store<NC>(v, lane, H, gb, eps, 1.0f, y + (int64_t)row * H);
}

// one warp per batch row: 32 mask words at a time via ballot
__global__ void mask_prep_kernel(const int64_t* __restrict__ mask, uint32_t* __restrict__ bits,
                                 int32_t* __restrict__ kv_len, int B, int S) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  if (b >= B) return;
  const int words = (S + 31) / 32;
  int last = 0;
  for (int w = 0; w < words; ++w) {
    const int j = w * 32 + lane;
    const bool v = (j < S) && (mask[(int64_t)b * S + j] != 0);
    const uint32_t m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) bits[(int64_t)b * words + w] = m;
    if (m) last = w * 32 + (32 - __clz(m));
  }
  if (lane == 0) kv_len[b] = last;
}

// one warp per batch row scans for <<LABEL>> tokens; then every lane copies rows with 128-bit loads
__global__ void __launch_bounds__(128)
head_gather_kernel(const __half* __restrict__ h, const int64_t* __restrict__ ids, const int64_t* __restrict__ mask,
                   int64_t class_token, int pool_mode, __half* __restrict__ pooled, __half* __restrict__ cls, int B, int S_in,
                   int H, int C, int pos_offset, const int32_t* __restrict__ text_row) {
  extern __shared__ int pos_s[];   // [C] per block (one batch row per block)
  const int b = blockIdx.x;
  // packed (varlen) layout: the text's rows are [text_row[b], text_row[b+1]) of the flat ids / mask / h arrays
  const int64_t base = text_row ? (int64_t)text_row[b] : (int64_t)b * S_in;
  const int S = text_row ? text_row[b + 1] - text_row[b] : S_in;
  ids += base;
  if (mask) mask += base;
  h += base * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  if (warp == 0) {
    int count = 0;
    for (int j0 = 0; j0 < S; j0 += 32) {
      const int j = j0 + lane;
      const bool hit = (j < S) && (ids[j] == class_token);
      const uint32_t m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int c = count + __popc(m & ((1u << lane) - 1u));
        if (c < C) pos_s[c] = (j + pos_offset < S) ? j + pos_offset : S - 1;   // embed_class_token=false reads the next token
      }
      count += __popc(m);
    }
    for (int c = count + lane; c < C; c += 32) pos_s[c] = -1;
  }
  __syncthreads();
  const int vec = H / 8;
  if (pool_mode >= 2) {
    // masked mean / masked max over the sequence (gliclass poolings): thread = 8 columns
    for (int i = threadIdx.x; i < vec; i += blockDim.x) {
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = (pool_mode == 2) ? 0.f : -3.4028234663852886e38f;
      int cnt = 0;
      for (int j = 0; j < S; ++j) {
        if (mask[j] == 0) continue;   // block-uniform
        ++cnt;
        const uint4 u = *reinterpret_cast<const uint4*>(h + (int64_t)j * H + i * 8);
        const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(hp[e]);
          if (pool_mode == 2) { acc[2 * e] += f.x; acc[2 * e + 1] += f.y; }
          else { acc[2 * e] = fmaxf(acc[2 * e], f.x); acc[2 * e + 1] = fmaxf(acc[2 * e + 1], f.y); }
        }
      }
      const float inv = (pool_mode == 2) ? 1.0f / (float)cnt : 1.0f;
      uint4 o;
      __half2* op = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) op[e] = __floats2half2_rn(acc[2 * e] * inv, acc[2 * e + 1] * inv);
      *reinterpret_cast<uint4*>(pooled + (int64_t)b * H + i * 8) = o;
    }
  }
  // row 0 of the output block is the pooled (first / last token) row, rows 1..C the class rows
  for (int r = (pool_mode >= 2 ? warp + 1 : warp); r <= C; r += (blockDim.x >> 5)) {
    const int p = (r == 0) ? (pool_mode == 1 ? S - 1 : 0) : pos_s[r - 1];
    __half* dst = (r == 0) ? pooled + (int64_t)b * H : cls + ((int64_t)b * C + (r - 1)) * H;
    const __half* src = h + (int64_t)(p < 0 ? 0 : p) * H;
    for (int i = lane; i < vec; i += 32) {
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (p >= 0) u = *reinterpret_cast<const uint4*>(src + i * 8);
      *reinterpret_cast<uint4*>(dst + i * 8) = u;
    }
  }
}

// one warp per (b,c): fp32 dot over Hh, then sigmoid / strict threshold (postprocessor.c:14-16,93-95)
__global__ void __launch_bounds__(256)
head_score_kernel(const float* __restrict__ t, const float* __restrict__ k, float* __restrict__ logits,
                  float* __restrict__ probs, uint8_t* __restrict__ decisions, float threshold, int B, int C, int Hh) {
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (idx >= B * C) return;
  const int b = idx / C;
  const float4* tv = reinterpret_cast<const float4*>(t + (int64_t)b * Hh);
  const float4* kv = reinterpret_cast<const float4*>(k + (int64_t)idx * Hh);
  float s = 0.f;
  for (int i = lane; i < Hh / 4; i += 32) {
    const float4 a = tv[i], c = kv[i];
    s = fmaf(a.x, c.x, s); s = fmaf(a.y, c.y, s); s = fmaf(a.z, c.z, s); s = fmaf(a.w, c.w, s);
  }
  s = warp_sum(s);
  if (lane == 0) {
    logits[idx] = s;
    const float p = 1.0f / (1.0f + expf(-s));
    if (probs) probs[idx] = p;
    if (decisions) decisions[idx] = p > threshold ? 1 : 0;
  }
}


// generalised scorer tail: one warp per (b,c) row.
//   logits[row] = scale * <t[b*t_stride ..], k[row,:]> * (normalize ? 1 / ((|t|+eps)(|k|+eps)) : 1) + bias
// t_stride = 0 turns it into the last Linear(K -> 1) of the MLP / weighted-dot scorers (t = weight row).
__global__ void __launch_bounds__(256)
head_score_ex_kernel(const float* __restrict__ t, int64_t t_stride, const float* __restrict__ k, float* __restrict__ logits,
                     float* __restrict__ probs, uint8_t* __restrict__ decisions, float threshold, int B, int C, int K,
                     int normalize, float eps, float scale, float bias) {
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  if (idx >= B * C) return;
  const int b = idx / C;
  const float4* tv = reinterpret_cast<const float4*>(t + (int64_t)b * t_stride);
  const float4* kv = reinterpret_cast<const float4*>(k + (int64_t)idx * K);
  float s = 0.f, tt = 0.f, kk = 0.f;
  for (int i = lane; i < K / 4; i += 32) {
    const float4 a = tv[i], c = kv[i];
    s = fmaf(a.x, c.x, s); s = fmaf(a.y, c.y, s); s = fmaf(a.z, c.z, s); s = fmaf(a.w, c.w, s);
    if (normalize) {
      tt = fmaf(a.x, a.x, tt); tt = fmaf(a.y, a.y, tt); tt = fmaf(a.z, a.z, tt); tt = fmaf(a.w, a.w, tt);
      kk = fmaf(c.x, c.x, kk); kk = fmaf(c.y, c.y, kk); kk = fmaf(c.z, c.z, kk); kk = fmaf(c.w, c.w, kk);
    }
  }
  s = warp_sum(s);
  if (normalize) {
    tt = warp_sum(tt);
    kk = warp_sum(kk);
    s = s / ((sqrtf(tt) + eps) * (sqrtf(kk) + eps));
  }
  if (lane == 0) {
    s = fmaf(s, scale, bias);
    logits[idx] = s;
    const float p = 1.0f / (1.0f + expf(-s));
    if (probs) probs[idx] = p;
    if (decisions) decisions[idx] = p > threshold ? 1 : 0;
  }
}

// fp32 rows -> fp16 rows with optional L2 normalisation x / (|x| + eps); one warp per row.
// Destination row r goes to dst + r * ld_dst + col0 (so the same kernel fills halves of a concat buffer);
// src row = r / rep (rep = C broadcasts the text row of batch b to its C class rows).
__global__ void __launch_bounds__(256)
head_rows16_kernel(const float* __restrict__ src, int K, int rep, __half* __restrict__ dst, int64_t ld_dst, int col0, int rows,
                   int normalize, float eps) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float4* sv = reinterpret_cast<const float4*>(src + (int64_t)(r / rep) * K);
  float inv = 1.0f;
  if (normalize) {
    float q = 0.f;
    for (int i = lane; i < K / 4; i += 32) {
      const float4 a = sv[i];
      q = fmaf(a.x, a.x, q); q = fmaf(a.y, a.y, q); q = fmaf(a.z, a.z, q); q = fmaf(a.w, a.w, q);
    }
    inv = 1.0f / (sqrtf(warp_sum(q)) + eps);
  }
  __half2* d = reinterpret_cast<__half2*>(dst + (int64_t)r * ld_dst + col0);
  for (int i = lane; i < K / 4; i += 32) {
    const float4 a = sv[i];
    d[2 * i] = __floats2half2_rn(a.x * inv, a.y * inv);
    d[2 * i + 1] = __floats2half2_rn(a.z * inv, a.w * inv);
  }
}

// weighted-dot scorer combine: pt [B,2Hh], pl [B*C,2Hh] fp32 with (d, half) interleaved ->
// cat[row] = [pt0_b | pl0_row | pt1_b * pl1_row]  fp16 [B*C, 3Hh]
__global__ void __launch_bounds__(256)
head_wdot_combine_kernel(const float* __restrict__ pt, const float* __restrict__ pl, __half* __restrict__ cat, int B, int C,
                         int Hh) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= B * C) return;
  const float2* tv = reinterpret_cast<const float2*>(pt + (int64_t)(r / C) * 2 * Hh);
  const float2* lv = reinterpret_cast<const float2*>(pl + (int64_t)r * 2 * Hh);
  __half* o = cat + (int64_t)r * 3 * Hh;
  for (int d = lane; d < Hh; d += 32) {
    const float2 a = tv[d], c = lv[d];
    o[d] = __float2half_rn(a.x);
    o[Hh + d] = __float2half_rn(c.x);
    o[2 * Hh + d] = __float2half_rn(a.y * c.y);
  }
}

// Reference behaviour at PADDED query rows (only observable through pooling_strategy='last' on a right-padded
// batch): the traced graph masks scores with the outer product of the attention mask (T:603-610, T:256), so a
// padded query row has every score at finfo.min, its softmax is exactly uniform over ALL S keys and its context
// is the plain mean of V over the S rows of the batch.  One block per batch row; no-op for unpadded rows.
__global__ void __launch_bounds__(256)
pad_rows_mean_v_kernel(const __half* __restrict__ qkv, const int64_t* __restrict__ mask, __half* __restrict__ ctx, int S, int H) {
  const int b = blockIdx.x;
  __shared__ int any_pad;
  if (threadIdx.x == 0) any_pad = 0;
  __syncthreads();
  int pad = 0;
  for (int j = threadIdx.x; j < S; j += blockDim.x) pad |= (mask[(int64_t)b * S + j] == 0);
  if (pad) any_pad = 1;
  __syncthreads();
  if (!any_pad) return;
  for (int c = threadIdx.x; c < H / 2; c += blockDim.x) {
    float ax = 0.f, ay = 0.f;
    for (int j = 0; j < S; ++j) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(qkv + ((int64_t)b * S + j) * 3 * H + 2 * H + 2 * c));
      ax += f.x; ay += f.y;
    }
    const __half2 m = __floats2half2_rn(ax / (float)S, ay / (float)S);
    for (int j = 0; j < S; ++j)
      if (mask[(int64_t)b * S + j] == 0) *reinterpret_cast<__half2*>(ctx + ((int64_t)b * S + j) * H + 2 * c) = m;
  }
}

inline int ln_grid_cap() {
  static int cap = 0;
  if (!cap) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const char* e = getenv("GLC_LN_BLOCKS_PER_SM");
    cap = sms * (e ? atoi(e) : 4);
  }
  return cap;
}

template <typename F>
cudaError_t dispatch_nc(int H, F&& f) {
  if (H % 8 != 0 || H > 256 * MAXC) return cudaErrorInvalidValue;
  const int nc = (H + 255) / 256;
  switch (nc) {
    case 1: return f(std::integral_constant<int, 1>{});
    case 2: return f(std::integral_constant<int, 2>{});
    case 3: return f(std::integral_constant<int, 3>{});
    case 4: return f(std::integral_constant<int, 4>{});
    case 5: case 6: return f(std::integral_constant<int, 6>{});
    default: return f(std::integral_constant<int, 8>{});
  }
}

}  // namespace

cudaError_t embed_ln(const int64_t* ids, const int64_t* mask, const void* emb, const float* gamma, const float* beta,
                     float eps, void* y, int M, int H, int vocab, cudaStream_t stream) {
  if (M <= 0) return cudaSuccess;
  return dispatch_nc(H, [&](auto nc) {
    constexpr int NC = decltype(nc)::value;
    return launch_pdl(embed_ln_kernel<NC>, dim3((M + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), dim3(ROWS_PER_BLOCK * 32), 0, stream, ids,
                      mask, (const __half*)emb, gamma, beta, eps, (__half*)y, M, H, vocab);
  });
}

cudaError_t residual_ln(const void* x, const void* r, const float* gamma, const float* beta, float eps, void* y, int M,
                        int H, cudaStream_t stream, int* overflow_flag, bool x_is_f32, void* y8, float* y8_scale) {
  if (M <= 0) return cudaSuccess;
  if ((y8 != nullptr) != (y8_scale != nullptr) || (y8 != nullptr && x_is_f32)) return cudaErrorInvalidValue;
  return dispatch_nc(H, [&](auto nc) {
    constexpr int NC = decltype(nc)::value;
    if (x_is_f32) {
      int blocks = (M + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
      const int cap = ln_grid_cap();
      if (blocks > cap) blocks = cap;
      residual_ln_xf32_kernel<NC><<<blocks, ROWS_PER_BLOCK * 32, 0, stream>>>((const float*)x, (const __half*)r, gamma, beta, eps,
                                                                              (__half*)y, M, H);
      return cudaGetLastError();
    }
    // bulk-copy ring variant (needs the residual operand and 16-byte rows); GLC_LN_BULK=0 keeps the register-prefetch kernel
    static const bool bulk_on = [] { const char* e = getenv("GLC_LN_BULK"); return !(e && e[0] == '0'); }();
    const size_t stage_bytes = (size_t)ROWS_PER_BLOCK * (r ? 2 : 1) * (size_t)H * 2;   // one stage of all 8 warps
    int stages = (int)((110 * 1024) / stage_bytes);
    if (stages > LN_STAGES) stages = LN_STAGES;
    const size_t ring_bytes = stage_bytes * stages + ROWS_PER_BLOCK * LN_STAGES * 8;
    if (bulk_on && (H * 2) % 16 == 0 && stages >= 2) {
      static bool attr_set[64][9] = {};
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      if (!attr_set[dev & 63][NC]) {
        cudaError_t e = cudaFuncSetAttribute(residual_ln_bulk_kernel<NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(residual_ln_bulk_kernel<NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63][NC] = true;
      }
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      int blocks = (M + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
      const int per_sm = (int)((224 * 1024) / (ring_bytes + 1024));
      if (blocks > sms * per_sm) blocks = sms * per_sm;
      if (r != nullptr && H == NC * 256)
        return launch_pdl(residual_ln_bulk_kernel<NC, true>, dim3(blocks), dim3(ROWS_PER_BLOCK * 32), ring_bytes, stream, (const __half*)x,
                          (const __half*)r, gamma, beta, eps, (__half*)y, M, H, stages, overflow_flag, (uint8_t*)y8, y8_scale);
      return launch_pdl(residual_ln_bulk_kernel<NC, false>, dim3(blocks), dim3(ROWS_PER_BLOCK * 32), ring_bytes, stream, (const __half*)x,
                        (const __half*)r, gamma, beta, eps, (__half*)y, M, H, stages, overflow_flag, (uint8_t*)y8, y8_scale);
    }
    int blocks = (M + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
    const int cap = ln_grid_cap();   // a few resident blocks per SM, each warp walking several rows
    if (blocks > cap) blocks = cap;
    return launch_pdl(residual_ln_kernel<NC>, dim3(blocks), dim3(ROWS_PER_BLOCK * 32), 0, stream, (const __half*)x, (const __half*)r,
                      gamma, beta, eps, (__half*)y, M, H, overflow_flag, (uint8_t*)y8, y8_scale);
  });
}

// one warp per row: amax, then e4m3 under the row's scale (tests and the load-time weight quantiser use it)
__global__ void __launch_bounds__(256)
quantize_rows_e4m3_kernel(const __half* __restrict__ x, int64_t ldx, uint8_t* __restrict__ q, int64_t ldq, float* __restrict__ scale,
                          int M, int K) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const __half* xr = x + (int64_t)row * ldx;
  float amax = 0.f;
  for (int k = lane; k < K; k += 32) amax = fmaxf(amax, fabsf(__half2float(xr[k])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  const float inv = amax > 0.f ? 448.0f / amax : 1.0f;
  if (lane == 0) scale[row] = amax > 0.f ? amax / 448.0f : 1.0f;
  uint8_t* qr = q + (int64_t)row * ldq;
  for (int k = 2 * lane; k < K; k += 64) {
    const float a = __half2float(xr[k]) * inv, b = (k + 1 < K) ? __half2float(xr[k + 1]) * inv : 0.f;
    const uint32_t p = ptx::pack_e4m3x2(a, b);
    qr[k] = (uint8_t)(p & 0xff);
    if (k + 1 < K) qr[k + 1] = (uint8_t)(p >> 8);
  }
}

cudaError_t quantize_rows_e4m3(const void* x_f16, int64_t ldx, void* q8, int64_t ldq, float* scale, int M, int K, cudaStream_t stream) {
  if (M <= 0 || K <= 0) return cudaSuccess;
  quantize_rows_e4m3_kernel<<<(M + 7) / 8, 256, 0, stream>>>((const __half*)x_f16, ldx, (uint8_t*)q8, ldq, scale, M, K);
  return cudaGetLastError();
}

cudaError_t ln_f32_to_f16(const float* x, const float* gamma, const float* beta, float eps, void* y, int M, int H,
                           cudaStream_t stream) {
  if (M <= 0) return cudaSuccess;
  return dispatch_nc(H, [&](auto nc) {
    constexpr int NC = decltype(nc)::value;
    ln_f32_kernel<NC><<<(M + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK, ROWS_PER_BLOCK * 32, 0, stream>>>(
        x, gamma, beta, eps, (__half*)y, M, H);
    return cudaGetLastError();
  });
}

cudaError_t mask_prep(const int64_t* mask, uint32_t* bits, int32_t* kv_len, int B, int S, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  return launch_pdl(mask_prep_kernel, dim3((B + 3) / 4), dim3(128), 0, stream, mask, bits, kv_len, B, S);
}

cudaError_t head_gather(const void* h, const int64_t* ids, int64_t class_token, void* pooled, void* cls, int B, int S,
                        int H, int C, cudaStream_t stream) {
  return head_gather_pool(h, ids, nullptr, class_token, 0, pooled, cls, B, S, H, C, stream, 0);
}

cudaError_t head_gather_pool(const void* h, const int64_t* ids, const int64_t* mask, int64_t class_token, int pool_mode,
                             void* pooled, void* cls, int B, int S, int H, int C, cudaStream_t stream, int class_pos_offset,
                             const int32_t* text_row) {
  if (B <= 0) return cudaSuccess;
  if (H % 8 != 0 || pool_mode < 0 || pool_mode > 3 || (pool_mode >= 2 && !mask) || class_pos_offset < 0) return cudaErrorInvalidValue;
  return launch_pdl(head_gather_kernel, dim3(B), dim3(128), (size_t)(C > 0 ? C : 1) * sizeof(int), stream, (const __half*)h, ids, mask,
                    class_token, pool_mode, (__half*)pooled, (__half*)cls, B, S, H, C, class_pos_offset, text_row);
}

cudaError_t pad_rows_mean_v(const void* qkv, const int64_t* mask, void* ctx, int B, int S, int H, cudaStream_t stream) {
  if (B <= 0 || S <= 0) return cudaSuccess;
  if (H % 2) return cudaErrorInvalidValue;
  pad_rows_mean_v_kernel<<<B, 256, 0, stream>>>((const __half*)qkv, mask, (__half*)ctx, S, H);
  return cudaGetLastError();
}

cudaError_t head_score_ex(const float* t, int64_t t_stride, const float* k, float* logits, float* probs, uint8_t* decisions,
                          float threshold, int B, int C, int K, bool normalize, float eps, float scale, float bias,
                          cudaStream_t stream) {
  if (B * C <= 0) return cudaSuccess;
  if (K % 4 != 0 || (t_stride % 4) != 0) return cudaErrorInvalidValue;
  return launch_pdl(head_score_ex_kernel, dim3((B * C + 7) / 8), dim3(256), 0, stream, t, t_stride, k, logits, probs, decisions,
                    threshold, B, C, K, normalize ? 1 : 0, eps, scale, bias);
}

cudaError_t head_rows16(const float* src, int K, int rep, void* dst_f16, int64_t ld_dst, int col0, int rows, bool normalize,
                        float eps, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (K % 4 != 0 || rep < 1 || (ld_dst % 2) || (col0 % 2)) return cudaErrorInvalidValue;
  head_rows16_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(src, K, rep, (__half*)dst_f16, ld_dst, col0, rows, normalize ? 1 : 0,
                                                         eps);
  return cudaGetLastError();
}

cudaError_t head_wdot_combine(const float* pt, const float* pl, void* cat_f16, int B, int C, int Hh, cudaStream_t stream) {
  if (B * C <= 0) return cudaSuccess;
  head_wdot_combine_kernel<<<(B * C + 7) / 8, 256, 0, stream>>>(pt, pl, (__half*)cat_f16, B, C, Hh);
  return cudaGetLastError();
}

cudaError_t head_score(const float* t, const float* k, float* logits, float* probs, uint8_t* decisions, float threshold,
                       int B, int C, int Hh, cudaStream_t stream) {
  if (B * C <= 0) return cudaSuccess;
  if (Hh % 4 != 0) return cudaErrorInvalidValue;
  head_score_kernel<<<(B * C + 7) / 8, 256, 0, stream>>>(t, k, logits, probs, decisions, threshold, B, C, Hh);
  return cudaGetLastError();
}

}  // namespace glc
