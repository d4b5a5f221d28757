/* vp_b200 — C ABI of the B200-native denoising hot path of VideoPainter (CogVideoX-5B-I2V backbone + context-encoder
 * branch).  This header is the drop-in boundary: every entry point replaces one piece of the reference's eager-PyTorch
 * path and is what a binding from the reference's Python (ctypes / torch custom-op) calls.  INTEGRATION.md shows the
 * reference-side stubs.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated otherwise; bf16 tensors are passed as `const void*`
 *   - `stream` is a cudaStream_t (the caller's current stream); nothing here allocates, synchronises or blocks the host
 *   - return value: 0 = ok, <0 = error (VP_ERR_*); vp_last_error() gives the text, vp_last_cuda_error() the cudaError_t
 *   - there is no CPU path and no alternate backend: unsupported shapes are errors
 *
 * Reference files (relative to /root/reference/diffusers/src/diffusers/models):
 *   T3D = transformers/cogvideox_transformer_3d.py   BR = branch_cogvideox.py   AP = attention_processor.py
 *   NRM = normalization.py   EMB = embeddings.py   ATT = attention.py   ACT = activations.py
 */
#ifndef VP_B200_H
#define VP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VP_OK 0
#define VP_ERR_BAD_SHAPE (-1)
#define VP_ERR_BAD_ALIGN (-2)
#define VP_ERR_UNSUPPORTED (-3)
#define VP_ERR_CUDA (-4)
#define VP_ERR_DRIVER (-5)

int vp_version(void);
const char* vp_last_error(void);       /* host string, valid until the next failing call on this thread */
int vp_last_cuda_error(void);

/* Timesteps / get_timestep_embedding (EMB:27-78, 777-793): out[b] = [cos(t w) | sin(t w)] (flip) in fp32.
 * Exactly one of t_i64 / t_f32 is non-null. */
int vp_time_sinusoid(const int64_t* t_i64, const float* t_f32, float* out, int batch, int dim, int flip_sin_to_cos,
                     float freq_shift, void* stream);

/* out[b, n] = sum_k act(in[b, k]) W[n, k] + bias[n]; act = SiLU when act_silu != 0.  fp32 activations, bf16 weights.
 * Replaces TimestepEmbedding.linear_1/linear_2 (EMB:762-774), CogVideoXLayerNormZero.linear(silu(temb)) (NRM:376) and
 * AdaLayerNorm.linear(silu(temb)) (NRM:73).  batch <= 8. */
int vp_gemv(const float* in, const void* weight, const void* bias, float* out, int batch, int n, int k, int act_silu,
            void* stream);

/* CogVideoXLayerNormZero modulation (NRM:377-378) over the joint (text | video) sequence:
 *   y[b, s] = LN(x[b, s]; gamma, beta, eps) * (1 + mod[b, scale_off(s)]) + mod[b, shift_off(s)]
 * where the text expert offsets are used for s < text_len.  x rows are read at (b * x_batch_rows + x_row_offset + s);
 * y is compact [batch * rows_per_batch, dim].  mod may be null (plain affine LayerNorm). */
int vp_ln_modulate(const void* x, long long x_batch_rows, int x_row_offset, void* y, int batch, int rows_per_batch, int dim,
                   const void* gamma, const void* beta, float eps, const float* mod, long long mod_batch_stride,
                   int shift_video_off, int scale_video_off, int shift_text_off, int scale_text_off, int text_len,
                   void* stream);

/* Final head normalisation (T3D:613-624): y = LN(LN(x; g1, b1); g2, b2) * (1 + mod[b, scale_off]) + mod[b, shift_off]
 * (norm_final, then AdaLayerNorm with chunk order shift, scale: NRM:78). */
int vp_ln_final(const void* x, long long x_batch_rows, int x_row_offset, void* y, int batch, int rows_per_batch, int dim,
                const void* gamma1, const void* beta1, const void* gamma2, const void* beta2, float eps, const float* mod,
                long long mod_batch_stride, int shift_off, int scale_off, void* stream);

/* tcgen05 GEMMs  C[M, N] = A[M, K] W[N, K]^T  (nn.Linear layout), bf16 in, fp32 accumulate, fused epilogues.
 * Logical row m maps to batch b = m / rows_per_batch and token s = m % rows_per_batch; the output row is
 * (b * out_batch_rows + out_row_offset + s) and rows with s + out_row_offset < 0 are dropped.  N % 64 == 0, K % 8 == 0. */

/* out = (A W^T + bias) * alpha            — text_proj (EMB:408), proj_out (T3D:624), branch_blocks (BR:416-421) */
int vp_gemm_bias(const void* a, long long lda, const void* w, long long ldw, const void* bias, void* out, int ldo, int m,
                 int n, int k, int rows_per_batch, long long out_batch_rows, int out_row_offset, float alpha, void* stream);

/* out = gelu_tanh(A W^T + bias)            — FeedForward.net[0] (ATT:1200, ACT:83) */
int vp_gemm_gelu(const void* a, long long lda, const void* w, long long ldw, const void* bias, void* out, int ldo, int m,
                 int n, int k, void* stream);

/* out = res + gate[b, expert(s)] * (A W^T + bias) [+ inject[b, s - text_len] where inject_mask == 0]
 * — attention out-proj + gated residual (AP:2202, T3D:169-170), FFN-2 + gated residual + branch injection
 * (ATT:1201, T3D:181-182, 596-609), and patch-embed conv + positional table (EMB:410-451; gate = null).
 * gate is fp32: gate[b * gate_batch_stride + (s < text_len ? gate_text_off : gate_video_off) + n].
 * a_k_chunk / a_chunk_stride: A may arrive split along K in chunks of a_k_chunk columns that lie a_chunk_stride
 * elements apart (the Ulysses all-to-all delivers the attention output as [peer][row][heads_per_peer * 64]);
 * a_k_chunk = 0 (or k) means a plain [m, k] matrix. */
int vp_gemm_gate_residual(const void* a, long long lda, const void* w, long long ldw, const void* bias, void* out, int ldo,
                          int m, int n, int k, int rows_per_batch, long long out_batch_rows, int out_row_offset,
                          const void* res, int ldr, long long res_batch_rows, int res_row_offset, const float* gate,
                          long long gate_batch_stride, int gate_video_off, int gate_text_off, int text_len,
                          const void* inject, long long inject_batch_stride, int ldi, const uint8_t* inject_mask,
                          int video_len, int a_k_chunk, long long a_chunk_stride, void* stream);

/* Fused to_q/to_k/to_v + QK LayerNorm(64) + 3D RoPE (AP:2132-2154), written head-major [B, H, S, 64].
 * w is [Wq; Wk; Wv] (qkv_first = 0) or [Wk; Wv] (qkv_first = 1, previous-window keys AP:2157-2172, 2247-2252).
 * row_scale[m] (nullable) multiplies the projection before the norm (prev_resample_mask * prev_clip_weight).
 * k2_out/v2_out (nullable) receive the masked copy of the ID-resample processor: K2 = RoPE(norm_k(k * mask2)),
 * V2 = v * mask2 (AP:2255-2281). rope tables are fp32 [video_len, 64] (nullable); rope_cs (nullable) is the compact form
 * [video_len, 32][cos, sin] of tables that repeat every value twice, as get_1d_rotary_pos_embed builds them (EMB:641-642):
 * when given it is used instead (one 256-byte fetch per token row and tile instead of 512 bytes per row and head).
 * heads_per_dest / dest_stride: head h of token s goes to
 *   x_out + (h / heads_per_dest) * dest_stride + ((b * heads_per_dest + h % heads_per_dest) * batch_rows + s) * 64,
 * i.e. heads_per_dest = heads (dest_stride ignored) is the plain [B, H, S, 64]; heads_per_dest = heads / P writes each
 * Ulysses destination rank's heads into its own contiguous send block (no pack pass before the all-to-all). */
int vp_gemm_qkv(const void* a, long long lda, const void* w, long long ldw, const void* bias, int m, int k, int batch_rows,
                int heads, int qkv_first, void* q_out, void* k_out, void* v_out, void* k2_out, void* v2_out,
                const uint8_t* mask2, const float* row_scale, const void* norm_q_w, const void* norm_q_b,
                const void* norm_k_w, const void* norm_k_b, float qk_eps, const float* rope_cos, const float* rope_sin,
                const float* rope_cs, int text_len, int heads_per_dest, long long dest_stride, void* stream);

/* Step end, one pass over the latent of ONE sample (n elements, bf16 latents as in the pipeline):
 *   model_output = uncond + guidance * (text - uncond)                     PIPE:981, 995-997 (fp32)
 *   CogVideoXDPMScheduler.step, prediction_type = "v_prediction"           DPM:386-436 (first / second order)
 *   latents = bf16(prev_sample); replace_gt re-noise and blend with mask    PIPE:1014-1034 (gt == null: skipped)
 * Coefficients are the scheduler's per-step scalars (DPM:306-328, 420-422, 451-463), computed by the host in float64 as the
 * reference does; the *_bf ones must already be rounded to bf16 (they multiply bf16 tensors in the reference, which casts
 * the 0-dim coefficient to the tensor's dtype first).  mask is [frames, 1, hw] bf16, broadcast over `chan` channels.
 * Results are bit-identical to the reference's sequence of eager ops. */
int vp_step_end(const void* noise_pred, float guidance, const void* sample, const float* old_pred, const void* noise,
                float c_sqrt_alpha_bf, float c_sqrt_beta, float c_m0_bf, float c_m1, float c_m2, float c_m3, float c_mn_bf,
                int second_order, float* pred_out, float* prev_out, void* latents_out, const void* gt, const void* noise0,
                const void* mask, int chan, long long hw, float sa_bf, float sb_bf, int renoise, int mask_background, long long n,
                void* stream);

/* ---- Ulysses over NVLink peer memory: the all-to-all is fused into the producing kernels' epilogues ------------------
 * Every rank of the sequence-parallel group owns a q/k/v buffer [slot][heads/peers][seq_total][64] and an attention-output
 * buffer [peers][seq_total/peers][ldo]; `peer_*` are HOST arrays of DEVICE pointers to all ranks' buffers (own rank included,
 * peer memory mapped through CUDA IPC), indexed by rank.  vp_gemm_qkv_peer writes head h of its rows straight into rank
 * h / (heads/peers)'s buffer at token rows [row_offset, row_offset + m) (q_out .. v2_out select the slot: they point into
 * the LOCAL buffer local_base); vp_attention_peer stores query row r into rank r / (seq_q/peers)'s output buffer at source
 * slot my_rank.  vp_peer_barrier orders the two: all earlier peer stores of every rank are visible to every rank after it
 * (system-scope release / acquire on per-rank flag words; epoch must increase by one per call, identically on all ranks;
 * epoch == 0: the kernel counts its own calls in word 9 of the rank's flag buffer instead, so that the launch does not
 * depend on the call history and can be replayed from a CUDA graph — one flag buffer must stick to one of the two modes). */
int vp_gemm_qkv_peer(const void* a, long long lda, const void* w, long long ldw, const void* bias, int m, int k, int heads,
                     int qkv_first, void* q_out, void* k_out, void* v_out, void* k2_out, void* v2_out, const uint8_t* mask2,
                     const float* row_scale, const void* norm_q_w, const void* norm_q_b, const void* norm_k_w,
                     const void* norm_k_b, float qk_eps, const float* rope_cos, const float* rope_sin, const float* rope_cs,
                     int text_len, void* const* peer_base, int peers, const void* local_base, int seq_total, int row_offset,
                     void* stream);
int vp_attention_peer(const void* q, const void* k0, const void* v0, int kv_len0, const void* k1, const void* v1, int kv_len1,
                      void* const* peer_out, int peers, int my_rank, int ldo, int heads, int seq_q, float softmax_scale,
                      float out_scale, void* stream);
int vp_peer_barrier(void* const* peer_flags, int peers, int my_rank, unsigned int epoch, void* stream);
/* The barrier's wait is bounded (default 20 000 ms, or the environment variable VP_B200_PEER_TIMEOUT_MS at first use; 0 = wait
 * for ever): a rank whose peers do not show up in time adds one to 32-bit word 8 of ITS flag buffer (the buffers are 64
 * bytes: words 0..7 = epochs written by the peers, word 8 = time-out count, word 9 = call count of the epoch == 0 mode) and lets the stream continue — no trap, no
 * sticky CUDA error; the host reads the word whenever it likes.  Kernel-replay profilers (ncu) stall single ranks for
 * longer than any sensible bound: profile peer mode with the timeout set to 0 and `--replay-mode application`. */
int vp_peer_set_timeout_ms(long long ms);
/* Alternative to vp_attention_peer: chunk d (bytes_per_peer bytes) of the local buffer `src` is copied into slot my_rank of
 * rank d's buffer peer_dst[d] ([peers][bytes_per_peer]) by one kernel of 16-byte peer stores (whole lines per warp). */
int vp_peer_scatter(const void* src, void* const* peer_dst, int peers, int my_rank, long long bytes_per_peer, void* stream);
/* Peer-visible device memory (host calls, synchronous): vp_peer_alloc = cudaMalloc + zero fill + CUDA IPC export (64-byte
 * handle, to be sent to the other ranks of the node by any means); vp_peer_open maps a peer's handle into the calling
 * process's current device with peer access enabled; vp_peer_close / vp_peer_free undo them. */
int vp_peer_alloc(long long bytes, void** ptr, unsigned char* handle64);
int vp_peer_open(const unsigned char* handle64, void** ptr);
int vp_peer_close(void* ptr);
int vp_peer_free(void* ptr);

/* Ulysses receive side (NCCL path): src [peers][slots][heads_local][rows_per_peer][64] (what the all-to-all delivers when every peer
 * sent its vp_gemm_qkv destination block) -> dst[slot] [heads_local][peers * rows_per_peer][64], the layout vp_attention
 * reads.  slots <= 5 (q, k, v, k2, v2); dst pointers beyond `slots` are ignored. */
int vp_a2a_unpack_heads(const void* src, void* dst0, void* dst1, void* dst2, void* dst3, void* dst4, int slots, int peers,
                        int heads_local, int rows_per_peer, void* stream);

/* softmax(Q K^T * scale) V over one or two K/V segments, d_head = 64, non-causal, unmasked (AP:2192-2197, 2285-2290).
 * q [B, H, seq_q, 64], k0/v0/k1/v1 [B, H, kv_len, 64]; out [B, seq_q, ldo] with head h at columns [64h, 64h + 64).
 * out = (accumulate ? out : 0) + out_scale * attention   (previous-window blend, AP:2176-2189). */
int vp_attention(const void* q, const void* k0, const void* v0, int kv_len0, const void* k1, const void* v1, int kv_len1,
                 void* out, int ldo, int batch, int heads, int seq_q, float softmax_scale, float out_scale, int accumulate,
                 void* stream);

/* CogVideoXPatchEmbed.proj input gather (EMB:404-414): A[(b f y x), c*4 + dy*2 + dx] from [BF, C0(+C1), H, W];
 * the two sources are concatenated along channels (BR:359); columns >= 4 (C0 + C1) are zero. */
int vp_patchify(const void* src0, int c0, const void* src1, int c1, int bf, int h, int w, void* out, int kpad, void* stream);

/* masks -> (avg_pool2d(mask, 2) > 0) per patch (EMB:417-426); mask is bf16 [BF, 1, H, W], out uint8 [BF * H/2 * W/2] */
int vp_mask_pool(const void* mask, int bf, int h, int w, uint8_t* out, void* stream);

/* unpatchify (T3D:630-632): proj [(b f y x), C*4] -> out [BF, C, H, W] */
int vp_unpatchify(const void* proj, int bf, int c, int h, int w, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VP_B200_H */
