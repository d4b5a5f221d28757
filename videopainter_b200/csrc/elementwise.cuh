// Parameter blocks / launchers of the token-wise kernels (elementwise.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vp {

struct LnModParams {
  long long rows;                 // output rows (B * rows_per_batch)
  int rows_per_batch;
  int D;
  const __nv_bfloat16* x;         // input row (b * x_batch_rows + x_row_offset + s)
  long long x_batch_rows;
  int x_row_offset;
  __nv_bfloat16* y;               // compact [rows, D]
  const __nv_bfloat16* gamma; const __nv_bfloat16* beta;
  const __nv_bfloat16* gamma2; const __nv_bfloat16* beta2;   // non-null: y = LN2(LN1(x)) modulated (final head)
  float eps;
  const float* mod;               // fp32 modulation table, [B, mod_batch_stride]; null = affine LayerNorm only
  long long mod_batch_stride;
  int shift_video_off, scale_video_off, shift_text_off, scale_text_off;
  int text_len;                   // rows s < text_len use the text expert
};

struct StepEndParams {
  long long n;                               // latent elements of one sample
  const __nv_bfloat16* noise_pred;           // [2, n]: uncond, text
  float guidance;
  const __nv_bfloat16* sample;               // [n]
  const float* old_pred;                     // [n], read when second_order
  const __nv_bfloat16* noise;                // [n] the noise the scheduler draws for the branch taken
  float c_sqrt_alpha_bf, c_sqrt_beta, c_m0_bf, c_m1, c_m2, c_m3, c_mn_bf;   // *_bf: already rounded to bf16
  int second_order;
  float* pred_out;                           // [n] pred_original_sample (carried to the next step)
  float* prev_out;                           // [n] fp32 prev_sample, or null
  __nv_bfloat16* latents_out;                // [n]
  const __nv_bfloat16* gt;                   // [n] or null (no replace_gt)
  const __nv_bfloat16* noise0;               // [n]
  const __nv_bfloat16* mask;                 // [frames, 1, hw] broadcast over `chan` channels
  int chan;
  long long hw;
  float sa_bf, sb_bf;
  int renoise, mask_background;
};
int launch_step_end(const StepEndParams& p, cudaStream_t st);
int launch_ln_modulate(const LnModParams& p, cudaStream_t st);
int launch_gemv(const float* in, const void* W, const void* bias, float* out, int B, int N, int K, int act_silu, cudaStream_t st);
int launch_timestep_sinusoid(const long long* t_i64, const float* t_f32, float* out, int B, int dim, int flip, float shift,
                             cudaStream_t st);
int launch_patchify(const void* src0, int C0, const void* src1, int C1, int BF, int H, int W, void* out, int Kpad, cudaStream_t st);
int launch_mask_pool(const void* mask, int BF, int H, int W, uint8_t* out, cudaStream_t st);
int launch_a2a_unpack_heads(const void* src, void* const* dst, int slots, int peers, int heads_local, int rows_per_peer,
                            cudaStream_t st);
int launch_peer_scatter(const void* src, void* const* peer_dst, int peers, int my_rank, long long bytes_per_peer, cudaStream_t st);
int set_peer_timeout_ms(long long ms);
int launch_peer_barrier(uint32_t* const* peer_flags, int peers, int my_rank, uint32_t epoch, cudaStream_t st);
int launch_unpatchify(const void* proj, int BF, int C, int H, int W, void* out, cudaStream_t st);

}  // namespace vp
