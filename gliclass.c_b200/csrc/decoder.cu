// HBM-bound kernels of the decoder-backbone path (Qwen2 / Llama style stack behind the GLiClass head; reference
// Readme.md:91-94 lists gliclass-qwen-1.5B / gliclass-llama-1.3B, BASELINE.json configs[4]).  Arithmetic restated from
// transformers' modeling_qwen2.py (Q:):
//   embedding gather              Q:365-366   h = embed_tokens[input_ids]
//   RMSNorm                       Q:258-264   y = w * x * rsqrt(mean(x^2) + eps)   (statistics in fp32)
//   pre-norm residual stream      Q:290-310   h = h + attn(norm1(h));  h = h + mlp(norm2(h))
//   rotary embedding              Q:102-147   q' = q cos + rotate_half(q) sin  (positions 0..S-1, pairs (p, p + d/2))
// The residual stream h stays fp32 in HBM (decoder checkpoints carry outlier channels far beyond the fp16 range of a
// pre-norm stream); the normalised activations that feed the tensor cores are fp16.  One warp per row, 128-bit accesses.
#include <cuda_fp16.h>

#include <type_traits>

#include "kernels.h"
#include "ptx.cuh"

namespace glc {
namespace {

constexpr int ROWS_PER_BLOCK = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// h[m,:] = emb[ids[m],:]  (fp16 table -> fp32 stream)
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
embed_rows_f32_kernel(const int64_t* __restrict__ ids, const __half* __restrict__ emb, float* __restrict__ h, int M, int H, int vocab) {
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  int64_t id = ids[row];
  if (id < 0 || id >= vocab) id = 0;
  const __half* src = emb + id * (int64_t)H;
  float* dst = h + (int64_t)row * H;
  for (int e0 = lane * 8; e0 < H; e0 += 256) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + e0));
    const __half2* p2 = reinterpret_cast<const __half2*>(&u);
    float4 a, b;
    float2 t = __half22float2(p2[0]); a.x = t.x; a.y = t.y;
    t = __half22float2(p2[1]); a.z = t.x; a.w = t.y;
    t = __half22float2(p2[2]); b.x = t.x; b.y = t.y;
    t = __half22float2(p2[3]); b.z = t.x; b.w = t.y;
    *reinterpret_cast<float4*>(dst + e0) = a;
    *reinterpret_cast<float4*>(dst + e0 + 4) = b;
  }
}

// h += delta (fp16, optional); y = rmsnorm(h) * g  (fp16).  NC chunks of 8 elements per lane.
template <int NC>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
add_rmsnorm_kernel(float* __restrict__ h, const __half* __restrict__ delta, const float* __restrict__ g, float eps,
                   __half* __restrict__ y, int M, int H, uint8_t* __restrict__ y8, float* __restrict__ y8_scale) {
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * ROWS_PER_BLOCK;
  float gw[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int e0 = (lane + 32 * c) * 8;
    if (e0 < H) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(g + e0)), b = __ldg(reinterpret_cast<const float4*>(g + e0 + 4));
      gw[c][0] = a.x; gw[c][1] = a.y; gw[c][2] = a.z; gw[c][3] = a.w;
      gw[c][4] = b.x; gw[c][5] = b.y; gw[c][6] = b.z; gw[c][7] = b.w;
    }
  }
  for (int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5); row < M; row += stride) {
    float* hr = h + (int64_t)row * H;
    float v[NC][8];
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int e0 = (lane + 32 * c) * 8;
      if (e0 < H) {
        const float4 a = *reinterpret_cast<const float4*>(hr + e0), b = *reinterpret_cast<const float4*>(hr + e0 + 4);
        v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w;
        v[c][4] = b.x; v[c][5] = b.y; v[c][6] = b.z; v[c][7] = b.w;
        if (delta) {
          const uint4 u = *reinterpret_cast<const uint4*>(delta + (int64_t)row * H + e0);
          const __half2* p2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 t = __half22float2(p2[k]);
            v[c][2 * k] += t.x;
            v[c][2 * k + 1] += t.y;
          }
          *reinterpret_cast<float4*>(hr + e0) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
          *reinterpret_cast<float4*>(hr + e0 + 4) = make_float4(v[c][4], v[c][5], v[c][6], v[c][7]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) q = fmaf(v[c][k], v[c][k], q);
      }
    }
    const float r = rsqrtf(warp_sum(q) / (float)H + eps);
    float amax = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int e0 = (lane + 32 * c) * 8;
      if (e0 < H) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[c][k] = v[c][k] * r * gw[c][k];
          amax = fmaxf(amax, fabsf(v[c][k]));
        }
        uint4 o;
        o.x = ptx::pack_f16(v[c][0], v[c][1]);
        o.y = ptx::pack_f16(v[c][2], v[c][3]);
        o.z = ptx::pack_f16(v[c][4], v[c][5]);
        o.w = ptx::pack_f16(v[c][6], v[c][7]);
        *reinterpret_cast<uint4*>(y + (int64_t)row * H + e0) = o;
      }
    }
    if (y8 != nullptr) {   // warp-uniform: the row again as e4m3 under its own scale (operand of the FP8 gate|up GEMM)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      const float inv = amax > 0.f ? 448.0f / amax : 1.0f;
      if (lane == 0) y8_scale[row] = amax > 0.f ? amax / 448.0f : 1.0f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int e0 = (lane + 32 * c) * 8;
        if (e0 < H) {
          uint2 o;
          o.x = ptx::pack_e4m3x4(v[c][0] * inv, v[c][1] * inv, v[c][2] * inv, v[c][3] * inv);
          o.y = ptx::pack_e4m3x4(v[c][4] * inv, v[c][5] * inv, v[c][6] * inv, v[c][7] * inv);
          *reinterpret_cast<uint2*>(y8 + (int64_t)row * H + e0) = o;
        }
      }
    }
  }
}

// rotary embedding in place on the Q and K parts of qkv [M, ld] (fp16): head j occupies columns [j*d, (j+1)*d), j <
// n_rot_heads (q heads then kv heads); pair (p, p + d/2) of a head at position s = m % S is rotated by s * inv_freq[p].
// cs: float2 [S, d/2] = (cos, sin).  One thread per (row, head, 8 consecutive p).
__global__ void __launch_bounds__(256)
rope_kernel(__half* __restrict__ qkv, int64_t ld, const float2* __restrict__ cs, int M, int S, int n_rot_heads, int d,
            const int32_t* __restrict__ tile_pos) {
  const int half = d >> 1, chunks = half >> 3;
  const int64_t total = (int64_t)M * n_rot_heads * chunks;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % chunks);
    const int hd = (int)((e / chunks) % n_rot_heads);
    const int64_t m = e / ((int64_t)chunks * n_rot_heads);
    // packed (varlen) layout: tile_pos[m / 128] = position of the first row of that 128-row tile within its text
    const int s = tile_pos ? __ldg(tile_pos + (m >> 7)) + (int)(m & 127) : (int)(m % S);
    __half* base = qkv + m * ld + (int64_t)hd * d + c * 8;
    uint4 lo = *reinterpret_cast<const uint4*>(base), hi = *reinterpret_cast<const uint4*>(base + half);
    __half2* l2 = reinterpret_cast<__half2*>(&lo);
    __half2* h2 = reinterpret_cast<__half2*>(&hi);
    const float2* t = cs + (int64_t)s * half + c * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = __half22float2(l2[k]), b = __half22float2(h2[k]);
      const float2 t0 = __ldg(t + 2 * k), t1 = __ldg(t + 2 * k + 1);
      // x1' = x1 cos - x2 sin ; x2' = x2 cos + x1 sin
      l2[k] = __floats2half2_rn(a.x * t0.x - b.x * t0.y, a.y * t1.x - b.y * t1.y);
      h2[k] = __floats2half2_rn(b.x * t0.x + a.x * t0.y, b.y * t1.x + a.y * t1.y);
    }
    *reinterpret_cast<uint4*>(base) = lo;
    *reinterpret_cast<uint4*>(base + half) = hi;
  }
}

__global__ void rope_table_kernel(const float* __restrict__ inv_freq, float2* __restrict__ cs, int S, int half) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S * half) return;
  const int s = e / half, p = e % half;
  const float a = (float)s * inv_freq[p];   // fp32 product, as Q:108-112 computes it (fp32 matmul of inv_freq and positions)
  cs[e] = make_float2(cosf(a), sinf(a));
}

inline int grid_cap() {
  static int cap = 0;
  if (!cap) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cap = sms * 8;
  }
  return cap;
}

}  // namespace

cudaError_t embed_rows_f32(const int64_t* ids, const void* emb_f16, float* h, int M, int H, int vocab, cudaStream_t stream) {
  if (M <= 0) return cudaSuccess;
  if (H % 8) return cudaErrorInvalidValue;
  embed_rows_f32_kernel<<<(M + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK, ROWS_PER_BLOCK * 32, 0, stream>>>(ids, (const __half*)emb_f16, h, M, H, vocab);
  return cudaGetLastError();
}

cudaError_t add_rmsnorm(float* h, const void* delta_f16, const float* g, float eps, void* y_f16, int M, int H,
                        cudaStream_t stream, void* y8, float* y8_scale) {
  if ((y8 != nullptr) != (y8_scale != nullptr)) return cudaErrorInvalidValue;
  if (M <= 0) return cudaSuccess;
  if (H % 8 != 0 || H > 256 * 8) return cudaErrorInvalidValue;
  int blocks = (M + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
  if (blocks > grid_cap()) blocks = grid_cap();
  const int nc = (H + 255) / 256;
  auto launch = [&](auto tag) {
    constexpr int NC = decltype(tag)::value;
    add_rmsnorm_kernel<NC><<<blocks, ROWS_PER_BLOCK * 32, 0, stream>>>(h, (const __half*)delta_f16, g, eps, (__half*)y_f16, M, H,
                                                                       (uint8_t*)y8, y8_scale);
    return cudaGetLastError();
  };
  switch (nc) {
    case 1: return launch(std::integral_constant<int, 1>{});
    case 2: return launch(std::integral_constant<int, 2>{});
    case 3: return launch(std::integral_constant<int, 3>{});
    case 4: return launch(std::integral_constant<int, 4>{});
    case 5: case 6: return launch(std::integral_constant<int, 6>{});
    default: return launch(std::integral_constant<int, 8>{});
  }
}

cudaError_t rope_table(const float* inv_freq, void* cs_f32x2, int S, int head_dim, cudaStream_t stream) {
  const int half = head_dim / 2, n = S * half;
  if (n <= 0) return cudaSuccess;
  rope_table_kernel<<<(n + 255) / 256, 256, 0, stream>>>(inv_freq, (float2*)cs_f32x2, S, half);
  return cudaGetLastError();
}

cudaError_t rope_inplace(void* qkv_f16, int64_t ld, const void* cs_f32x2, int M, int S, int n_rot_heads, int head_dim, cudaStream_t stream,
                         const int32_t* tile_pos) {
  if (M <= 0) return cudaSuccess;
  if (head_dim % 16 != 0 || ld % 8 != 0) return cudaErrorInvalidValue;
  const int64_t total = (int64_t)M * n_rot_heads * (head_dim / 16);
  int64_t blocks = (total + 255) / 256;
  if (blocks > grid_cap() * 4) blocks = grid_cap() * 4;
  rope_kernel<<<(int)blocks, 256, 0, stream>>>((__half*)qkv_f16, ld, (const float2*)cs_f32x2, M, S, n_rot_heads, head_dim, tile_pos);
  return cudaGetLastError();
}

}  // namespace glc
