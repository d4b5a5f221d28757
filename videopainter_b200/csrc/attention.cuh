// Parameter block shared by attention.cu (kernel) and capi.cu (C ABI).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vp {

struct AttnParams {
  int batch, heads;
  int seq_q;                 // query rows per (batch, head)
  int kv_len0, kv_len1;      // keys in segment 0 / segment 1 (0 = absent)
  float scale_log2;          // softmax scale * log2(e)
  __nv_bfloat16* out;        // [batch, seq_q, ldo]; head h occupies columns [h*64, h*64+64)
  int ldo;
  float out_scale;           // out = (accumulate ? out : 0) + out_scale * softmax(QKᵀ)V   (prev-window blend AP:2176-2189)
  int accumulate;
  // peer mode (Ulysses over NVLink peer memory): when peer_out[0] != null, query row q belongs to rank q / peer_rows and is
  // stored into that rank's buffer [source rank][peer_rows][ldo] at source slot peer_src; `out` is ignored
  __nv_bfloat16* peer_out[8];
  int peer_rows;
  int peer_src;
  float one;                 // 1.0f, set by the launcher (a multiplier the compiler cannot fold: keeps packed adds on FFMA2)
};

int launch_attention(const void* q, const void* k0, const void* v0, const void* k1, const void* v1, const AttnParams& p,
                     cudaStream_t st);

}  // namespace vp
