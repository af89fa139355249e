// Helpers around the K3 attention kernels:
//   * the position tables expanded to one row per relative distance (load time, input independent),
//   * a slow CUDA-core restatement of the attention op (tests only: the on-GPU reference that localises a
//     bug to a (batch, head, row) without a host round trip).
// Arithmetic: transformers DisentangledSelfAttention, T:229-345; idx(delta) = clamp(bucket(delta) + span, 0, 2 span - 1)
// (SURVEY.md App. A.6).
#include <cuda_fp16.h>
#include <math_constants.h>

#include <vector>

#include "kernels.h"
#include "model_weights.h"

namespace glc {
namespace {

constexpr int D = 64;
constexpr int QT = 128;
constexpr int EXP_CENTER = 2047;   // expanded tables: row rho = EXP_CENTER - delta
constexpr int EXP_ROWS = 4096;

// one warp per (b, h, i): fp32 math on the same fp16 inputs
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

void expanded_pos_index_rev(int buckets, int max_pos, int32_t* out /* [EXP_ROWS] */) {
  // row sigma holds delta = sigma - EXP_CENTER; the last row (delta = +2048) is never consumed -> zero row
  std::vector<int32_t> rel((size_t)2 * (EXP_CENTER + 1) - 1);
  rel_index_table(EXP_CENTER + 1, buckets, max_pos, rel.data());   // rel[delta + EXP_CENTER]
  for (int s = 0; s < EXP_ROWS - 1; ++s) out[s] = rel[(size_t)s];
  out[EXP_ROWS - 1] = -1;
}

cudaError_t expand_pos_table(const void* pos_f16, int64_t ld_src, const int32_t* d_exp_index, void* out_f16, int64_t ld_dst,
                             int cols, cudaStream_t stream) {
  if (cols % 8 != 0 || ld_src % 8 != 0 || ld_dst % 8 != 0) return cudaErrorInvalidValue;
  expand_rows_kernel<<<592, 256, 0, stream>>>((const __half*)pos_f16, ld_src, d_exp_index, (__half*)out_f16, ld_dst,
                                              EXP_ROWS, cols);
  return cudaGetLastError();
}

cudaError_t attention_naive(const void* qkv, const void* pos_k, const void* pos_q, int64_t ld_pos, const int32_t* rel_idx,
                            const uint32_t* mask_bits, void* ctx, int B, int S, int heads, int buckets,
                            cudaStream_t stream) {
  (void)buckets;
  if (B <= 0 || S <= 0) return cudaSuccess;
  const int Spad = ((S + QT - 1) / QT) * QT;
  const int rows = B * heads * S;
  attention_naive_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(
      (const __half*)qkv, (const __half*)pos_k, (const __half*)pos_q, rel_idx, Spad - 1, mask_bits,
      (__half*)ctx, B, S, heads, (int)ld_pos, 1.0f / sqrtf(3.0f * D));
  return cudaGetLastError();
}

}  // namespace glc
