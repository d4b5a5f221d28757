// Parameter block shared by gemm.cu (kernel) and capi.cu (C ABI).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace vp {

enum GemmEpilogue : int {
  EPI_BIAS = 0,    // out = (acc + bias) * alpha
  EPI_GELU = 1,    // out = gelu_tanh(acc + bias)                        (ATT:1200 + ACT:83)
  EPI_RESID = 2,   // out = res + gate * (acc + bias) [+ inject]          (T3D:169-170, 181-182, 596-609)
  EPI_QKV = 3,     // bias -> per-head LayerNorm(64) -> RoPE -> [B,H,S,64] (AP:2132-2154, 2255-2281)
};

struct GemmParams {
  int M, N, K;
  int group_m;               // rasterisation: m-tiles per group (L2 reuse of A)
  // A may be split along K into chunks of a_k_chunk columns that live a_chunk_stride elements apart (the attention output
  // gathered from the Ulysses peers arrives as [peer][row][heads_per_peer * 64]); a_k_chunk = K means one plain matrix
  int a_k_chunk;
  long long a_chunk_stride;
  // logical row m -> batch b = m / rows_per_batch, token s = m % rows_per_batch
  int rows_per_batch;
  // ---- plain / gelu / residual output: row (b * out_batch_rows + out_row_offset + s), skipped if s + off < 0
  __nv_bfloat16* out;
  long long out_batch_rows;
  int out_row_offset;
  int ldo;
  const __nv_bfloat16* bias;      // [N] or null
  float alpha;
  // ---- residual epilogue
  const __nv_bfloat16* res;       // row (b * res_batch_rows + res_row_offset + s), leading dim ldr
  long long res_batch_rows;
  int res_row_offset;
  int ldr;
  const float* gate;              // gate[b * gate_batch_stride + (s < text_len ? gate_text_off : gate_video_off) + n]; null -> 1
  long long gate_batch_stride;
  int gate_video_off, gate_text_off;
  int text_len;
  const __nv_bfloat16* inject;    // [B, Sv, ldi] added to video rows (s >= text_len) where inject_mask == 0
  long long inject_batch_stride;  // elements
  int ldi;
  const uint8_t* inject_mask;     // [B, Sv] (1 = inside the region to synthesise -> no add) or null
  int video_len;
  // ---- QKV epilogue
  int d_model;                    // H * 64
  int heads;
  int qkv_first;                  // 0: columns are [Q|K|V]; 1: columns are [K|V]
  // head h is written to q_out + (h / heads_per_dest) * dest_stride + ((b * heads_per_dest + h % heads_per_dest) * S + s) * 64:
  // heads_per_dest = heads gives the plain [B, H, S, 64]; smaller values lay the heads out per Ulysses destination rank
  int heads_per_dest;
  long long dest_stride;          // elements
  // peer mode (Ulysses over NVLink peer memory): when peer_base[0] != null the outputs of head h go straight into the
  // attention layout [slot][heads_per_dest][peer_seq][64] of destination rank h / heads_per_dest, at token row
  // peer_row_off + s:  dst = peer_base[dest] + (x_out - local_base) + ((h % heads_per_dest) * peer_seq + peer_row_off + s) * 64
  // (x_out then only selects the slot inside the local copy of that buffer); dest_stride is ignored
  __nv_bfloat16* peer_base[8];
  const __nv_bfloat16* local_base;
  int peer_seq;
  int peer_row_off;
  __nv_bfloat16* q_out;           // [B, H, S, 64]
  __nv_bfloat16* k_out;
  __nv_bfloat16* v_out;
  __nv_bfloat16* k2_out;          // masked copy (resample processor) or null
  __nv_bfloat16* v2_out;
  const uint8_t* mask2;           // [M] 1 = keep
  const float* row_scale;         // [M] multiplies (acc + bias) before the norm, or null
  const __nv_bfloat16* nq_w; const __nv_bfloat16* nq_b;
  const __nv_bfloat16* nk_w; const __nv_bfloat16* nk_b;
  float qk_eps;
  const float* rope_cos;          // [Sv, 64] fp32 or null
  const float* rope_sin;
  const float* rope_cs;           // [Sv, 32][cos, sin] fp32: compact form of pair-repeated tables (preferred when given)
};

int launch_gemm(int epi, const void* A, long long lda, const void* W, long long ldw, const GemmParams& p, cudaStream_t st);

}  // namespace vp
