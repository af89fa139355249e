// Launchers of the superseded K3 generations kept for reference (not part of libgliclass_b200.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace glc {
cudaError_t attention_fused(const void* qkv, const void* pos_k, const void* pos_q, int64_t ld_pos, const int32_t* rel_idx,
                            const uint32_t* mask_bits, const int32_t* kv_len, void* ctx, int B, int S, int heads,
                            int buckets, int num_sms, cudaStream_t stream);
cudaError_t attention_toeplitz(const void* qkv, const void* exp_k, const void* exp_q, int64_t ld_exp,
                               const uint32_t* mask_bits, const int32_t* kv_len, void* ctx, int B, int S, int heads,
                               cudaStream_t stream);
cudaError_t attention_stream(const void* qkv, const void* exp_k, const void* exp_qr, int64_t ld_exp,
                             const uint32_t* mask_bits, const int32_t* kv_len, void* ctx, int B, int S, int heads,
                             cudaStream_t stream);
// round-1 production kernel: two warps share a query row and exchange the row maximum
cudaError_t attention_shift(const void* qkv, const void* exp_k, const void* exp_qr, int64_t ld_exp,
                            const uint32_t* mask_bits, const int32_t* kv_len, void* ctx, int B, int S, int heads,
                            cudaStream_t stream);
// round-2 intermediate: row-owner softmax, one CTA per (query tile, head, text); attention_persist.cu is its persistent form
cudaError_t attention_rows(const void* qkv, const void* exp_k, const void* exp_qr, int64_t ld_exp,
                           const uint32_t* mask_bits, const int32_t* kv_len, void* ctx, int B, int S, int heads,
                           cudaStream_t stream);
}  // namespace glc
