// extern "C" surface declared in include/vp_b200.h: argument validation + launches.  No torch types, no allocation.
#include "../../include/vp_b200.h"

#include "attention.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "host_util.cuh"

using namespace vp;

extern "C" {

int vp_version(void) { return 100; }
const char* vp_last_error(void) { return err_state().msg; }
int vp_last_cuda_error(void) { return err_state().cuda_error; }

int vp_time_sinusoid(const int64_t* t_i64, const float* t_f32, float* out, int batch, int dim, int flip_sin_to_cos,
                     float freq_shift, void* stream) {
  return launch_timestep_sinusoid(reinterpret_cast<const long long*>(t_i64), t_f32, out, batch, dim, flip_sin_to_cos, freq_shift,
                                  (cudaStream_t)stream);
}

int vp_gemv(const float* in, const void* weight, const void* bias, float* out, int batch, int n, int k, int act_silu,
            void* stream) {
  VP_REQUIRE(in && weight && out, VP_ERR_BAD_SHAPE, "gemv: null pointer");
  return launch_gemv(in, weight, bias, out, batch, n, k, act_silu, (cudaStream_t)stream);
}

int vp_ln_modulate(const void* x, long long x_batch_rows, int x_row_offset, void* y, int batch, int rows_per_batch, int dim,
                   const void* gamma, const void* beta, float eps, const float* mod, long long mod_batch_stride,
                   int shift_video_off, int scale_video_off, int shift_text_off, int scale_text_off, int text_len,
                   void* stream) {
  VP_REQUIRE(x && y && gamma && beta, VP_ERR_BAD_SHAPE, "ln_modulate: null pointer");
  VP_REQUIRE(batch > 0 && rows_per_batch > 0, VP_ERR_BAD_SHAPE, "ln_modulate: empty input");
  LnModParams p{};
  p.rows = (long long)batch * rows_per_batch;
  p.rows_per_batch = rows_per_batch;
  p.D = dim;
  p.x = (const __nv_bfloat16*)x; p.x_batch_rows = x_batch_rows; p.x_row_offset = x_row_offset;
  p.y = (__nv_bfloat16*)y;
  p.gamma = (const __nv_bfloat16*)gamma; p.beta = (const __nv_bfloat16*)beta;
  p.gamma2 = nullptr; p.beta2 = nullptr;
  p.eps = eps;
  p.mod = mod; p.mod_batch_stride = mod_batch_stride;
  p.shift_video_off = shift_video_off; p.scale_video_off = scale_video_off;
  p.shift_text_off = shift_text_off; p.scale_text_off = scale_text_off;
  p.text_len = text_len;
  return launch_ln_modulate(p, (cudaStream_t)stream);
}

int vp_ln_final(const void* x, long long x_batch_rows, int x_row_offset, void* y, int batch, int rows_per_batch, int dim,
                const void* gamma1, const void* beta1, const void* gamma2, const void* beta2, float eps, const float* mod,
                long long mod_batch_stride, int shift_off, int scale_off, void* stream) {
  VP_REQUIRE(x && y && gamma1 && beta1 && gamma2 && beta2 && mod, VP_ERR_BAD_SHAPE, "ln_final: null pointer");
  VP_REQUIRE(batch > 0 && rows_per_batch > 0, VP_ERR_BAD_SHAPE, "ln_final: empty input");
  LnModParams p{};
  p.rows = (long long)batch * rows_per_batch;
  p.rows_per_batch = rows_per_batch;
  p.D = dim;
  p.x = (const __nv_bfloat16*)x; p.x_batch_rows = x_batch_rows; p.x_row_offset = x_row_offset;
  p.y = (__nv_bfloat16*)y;
  p.gamma = (const __nv_bfloat16*)gamma1; p.beta = (const __nv_bfloat16*)beta1;
  p.gamma2 = (const __nv_bfloat16*)gamma2; p.beta2 = (const __nv_bfloat16*)beta2;
  p.eps = eps;
  p.mod = mod; p.mod_batch_stride = mod_batch_stride;
  p.shift_video_off = shift_off; p.scale_video_off = scale_off;
  p.shift_text_off = shift_off; p.scale_text_off = scale_off;
  p.text_len = 0;
  return launch_ln_modulate(p, (cudaStream_t)stream);
}

static GemmParams base_params(int m, int n, int k, int rows_per_batch, void* out, int ldo, long long out_batch_rows,
                              int out_row_offset, const void* bias) {
  GemmParams p{};
  p.M = m; p.N = n; p.K = k;
  p.group_m = 16;
  p.rows_per_batch = rows_per_batch;
  p.out = (__nv_bfloat16*)out;
  p.out_batch_rows = out_batch_rows;
  p.out_row_offset = out_row_offset;
  p.ldo = ldo;
  p.bias = (const __nv_bfloat16*)bias;
  p.alpha = 1.0f;
  return p;
}

int vp_gemm_bias(const void* a, long long lda, const void* w, long long ldw, const void* bias, void* out, int ldo, int m,
                 int n, int k, int rows_per_batch, long long out_batch_rows, int out_row_offset, float alpha, void* stream) {
  VP_REQUIRE(a && w && out, VP_ERR_BAD_SHAPE, "gemm_bias: null pointer");
  VP_REQUIRE(ldo % 8 == 0, VP_ERR_BAD_ALIGN, "gemm_bias: ldo must be a multiple of 8");
  GemmParams p = base_params(m, n, k, rows_per_batch, out, ldo, out_batch_rows, out_row_offset, bias);
  p.alpha = alpha;
  return launch_gemm(EPI_BIAS, a, lda, w, ldw, p, (cudaStream_t)stream);
}

int vp_gemm_gelu(const void* a, long long lda, const void* w, long long ldw, const void* bias, void* out, int ldo, int m,
                 int n, int k, void* stream) {
  VP_REQUIRE(a && w && out, VP_ERR_BAD_SHAPE, "gemm_gelu: null pointer");
  VP_REQUIRE(ldo % 8 == 0, VP_ERR_BAD_ALIGN, "gemm_gelu: ldo must be a multiple of 8");
  GemmParams p = base_params(m, n, k, m, out, ldo, 0, 0, bias);
  p.group_m = 32;
  return launch_gemm(EPI_GELU, a, lda, w, ldw, p, (cudaStream_t)stream);
}

int vp_gemm_gate_residual(const void* a, long long lda, const void* w, long long ldw, const void* bias, void* out, int ldo,
                          int m, int n, int k, int rows_per_batch, long long out_batch_rows, int out_row_offset,
                          const void* res, int ldr, long long res_batch_rows, int res_row_offset, const float* gate,
                          long long gate_batch_stride, int gate_video_off, int gate_text_off, int text_len,
                          const void* inject, long long inject_batch_stride, int ldi, const uint8_t* inject_mask,
                          int video_len, int a_k_chunk, long long a_chunk_stride, void* stream) {
  VP_REQUIRE(a && w && out && res, VP_ERR_BAD_SHAPE, "gemm_gate_residual: null pointer");
  VP_REQUIRE(ldo % 8 == 0 && ldr % 8 == 0 && (inject == nullptr || ldi % 8 == 0), VP_ERR_BAD_ALIGN,
             "gemm_gate_residual: leading dims must be multiples of 8");
  VP_REQUIRE(gate == nullptr || (gate_batch_stride % 4 == 0 && gate_video_off % 4 == 0 && gate_text_off % 4 == 0),
             VP_ERR_BAD_ALIGN, "gemm_gate_residual: gate offsets must be multiples of 4");
  GemmParams p = base_params(m, n, k, rows_per_batch, out, ldo, out_batch_rows, out_row_offset, bias);
  p.res = (const __nv_bfloat16*)res; p.ldr = ldr; p.res_batch_rows = res_batch_rows; p.res_row_offset = res_row_offset;
  p.gate = gate; p.gate_batch_stride = gate_batch_stride; p.gate_video_off = gate_video_off; p.gate_text_off = gate_text_off;
  p.text_len = text_len;
  p.inject = (const __nv_bfloat16*)inject; p.inject_batch_stride = inject_batch_stride; p.ldi = ldi;
  p.inject_mask = inject_mask; p.video_len = video_len;
  p.a_k_chunk = a_k_chunk; p.a_chunk_stride = a_chunk_stride;
  // rasterisation: the A panel of one m-group (group_m * 128 rows * K) has to stay in L2 next to W while the group walks over
  // all n-tiles; with K = 12288 (FFN-2) 16 m-tiles are 50 MB and A was fetched 3-4 times from DRAM (ncu: 3.56 GB per launch).
  // CTA-pair kernel, stand-alone at M = 35552: group 2 / 4 / 8 / 16 / 32 = 1519 / 1512 / 1477 / 1500 / 1483 TFLOP/s
  if ((long long)k * 2 * 128 * 16 > (24ll << 20)) p.group_m = 4;
  return launch_gemm(EPI_RESID, a, lda, w, ldw, p, (cudaStream_t)stream);
}

struct PeerDest {
  void* const* base = nullptr;    // host array of device pointers, one per destination rank
  int peers = 0;
  const void* local_base = nullptr;
  int seq = 0, row_off = 0;
};

static int gemm_qkv_impl(const void* a, long long lda, const void* w, long long ldw, const void* bias, int m, int k, int batch_rows,
                         int heads, int qkv_first, void* q_out, void* k_out, void* v_out, void* k2_out, void* v2_out,
                         const uint8_t* mask2, const float* row_scale, const void* norm_q_w, const void* norm_q_b,
                         const void* norm_k_w, const void* norm_k_b, float qk_eps, const float* rope_cos, const float* rope_sin,
                         const float* rope_cs, int text_len, int heads_per_dest, long long dest_stride, const PeerDest& peer,
                         void* stream) {
  VP_REQUIRE(a && w && bias && k_out && v_out && norm_k_w && norm_k_b, VP_ERR_BAD_SHAPE, "gemm_qkv: null pointer");
  VP_REQUIRE(heads_per_dest > 0 && heads % heads_per_dest == 0 && dest_stride % 8 == 0, VP_ERR_BAD_SHAPE,
             "gemm_qkv: heads_per_dest must divide heads");
  VP_REQUIRE(qkv_first == 0 || qkv_first == 1, VP_ERR_BAD_SHAPE, "gemm_qkv: qkv_first must be 0 or 1");
  VP_REQUIRE(qkv_first == 1 || (q_out && norm_q_w && norm_q_b), VP_ERR_BAD_SHAPE, "gemm_qkv: q outputs missing");
  VP_REQUIRE((k2_out == nullptr) == (v2_out == nullptr) && (k2_out == nullptr || mask2 != nullptr), VP_ERR_BAD_SHAPE,
             "gemm_qkv: masked copy needs k2, v2 and mask2");
  VP_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr), VP_ERR_BAD_SHAPE, "gemm_qkv: rope tables");
  VP_REQUIRE(heads > 0 && batch_rows > 0 && m % batch_rows == 0, VP_ERR_BAD_SHAPE, "gemm_qkv: m must be batch * batch_rows");
  const int d_model = heads * 64;
  const int n = (3 - qkv_first) * d_model;
  GemmParams p = base_params(m, n, k, batch_rows, nullptr, 0, 0, 0, bias);
  p.d_model = d_model; p.heads = heads; p.qkv_first = qkv_first;
  p.q_out = (__nv_bfloat16*)q_out; p.k_out = (__nv_bfloat16*)k_out; p.v_out = (__nv_bfloat16*)v_out;
  p.k2_out = (__nv_bfloat16*)k2_out; p.v2_out = (__nv_bfloat16*)v2_out;
  p.mask2 = mask2; p.row_scale = row_scale;
  p.nq_w = (const __nv_bfloat16*)norm_q_w; p.nq_b = (const __nv_bfloat16*)norm_q_b;
  p.nk_w = (const __nv_bfloat16*)norm_k_w; p.nk_b = (const __nv_bfloat16*)norm_k_b;
  p.qk_eps = qk_eps;
  p.rope_cos = rope_cos; p.rope_sin = rope_sin; p.rope_cs = rope_cs;
  p.text_len = text_len;
  p.heads_per_dest = heads_per_dest; p.dest_stride = dest_stride;
  p.group_m = 32;                 // measured: 1.413 ms against 1.436 (16) and 1.458 (8) at M = 35552
  if (peer.base) {
    VP_REQUIRE(peer.peers >= 1 && peer.peers <= 8 && peer.peers * heads_per_dest == heads && peer.local_base && peer.seq > 0 &&
                   m == batch_rows, VP_ERR_BAD_SHAPE, "gemm_qkv_peer: one sample per rank, peers * heads_per_dest == heads");
    for (int i = 0; i < peer.peers; ++i) {
      VP_REQUIRE(peer.base[i] != nullptr, VP_ERR_BAD_SHAPE, "gemm_qkv_peer: null peer pointer");
      p.peer_base[i] = (__nv_bfloat16*)peer.base[i];
    }
    p.local_base = (const __nv_bfloat16*)peer.local_base;
    p.peer_seq = peer.seq; p.peer_row_off = peer.row_off;
  }
  return launch_gemm(EPI_QKV, a, lda, w, ldw, p, (cudaStream_t)stream);
}

int vp_gemm_qkv(const void* a, long long lda, const void* w, long long ldw, const void* bias, int m, int k, int batch_rows,
                int heads, int qkv_first, void* q_out, void* k_out, void* v_out, void* k2_out, void* v2_out,
                const uint8_t* mask2, const float* row_scale, const void* norm_q_w, const void* norm_q_b,
                const void* norm_k_w, const void* norm_k_b, float qk_eps, const float* rope_cos, const float* rope_sin,
                const float* rope_cs, int text_len, int heads_per_dest, long long dest_stride, void* stream) {
  return gemm_qkv_impl(a, lda, w, ldw, bias, m, k, batch_rows, heads, qkv_first, q_out, k_out, v_out, k2_out, v2_out, mask2,
                       row_scale, norm_q_w, norm_q_b, norm_k_w, norm_k_b, qk_eps, rope_cos, rope_sin, rope_cs, text_len,
                       heads_per_dest, dest_stride, PeerDest{}, stream);
}

int vp_gemm_qkv_peer(const void* a, long long lda, const void* w, long long ldw, const void* bias, int m, int k, int heads,
                     int qkv_first, void* q_out, void* k_out, void* v_out, void* k2_out, void* v2_out, const uint8_t* mask2,
                     const float* row_scale, const void* norm_q_w, const void* norm_q_b, const void* norm_k_w,
                     const void* norm_k_b, float qk_eps, const float* rope_cos, const float* rope_sin, const float* rope_cs,
                     int text_len, void* const* peer_base, int peers, const void* local_base, int seq_total, int row_offset,
                     void* stream) {
  VP_REQUIRE(peer_base && peers >= 1 && heads % peers == 0, VP_ERR_BAD_SHAPE, "gemm_qkv_peer: peers must divide heads");
  PeerDest pd;
  pd.base = peer_base; pd.peers = peers; pd.local_base = local_base; pd.seq = seq_total; pd.row_off = row_offset;
  return gemm_qkv_impl(a, lda, w, ldw, bias, m, k, m, heads, qkv_first, q_out, k_out, v_out, k2_out, v2_out, mask2, row_scale,
                       norm_q_w, norm_q_b, norm_k_w, norm_k_b, qk_eps, rope_cos, rope_sin, rope_cs, text_len, heads / peers, 0, pd,
                       stream);
}

int vp_attention(const void* q, const void* k0, const void* v0, int kv_len0, const void* k1, const void* v1, int kv_len1,
                 void* out, int ldo, int batch, int heads, int seq_q, float softmax_scale, float out_scale, int accumulate,
                 void* stream) {
  VP_REQUIRE(q && k0 && v0 && out, VP_ERR_BAD_SHAPE, "attention: null pointer");
  AttnParams p{};
  p.batch = batch; p.heads = heads; p.seq_q = seq_q;
  p.kv_len0 = kv_len0; p.kv_len1 = kv_len1;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.out = (__nv_bfloat16*)out; p.ldo = ldo;
  p.out_scale = out_scale; p.accumulate = accumulate;
  return launch_attention(q, k0, v0, k1, v1, p, (cudaStream_t)stream);
}

int vp_attention_peer(const void* q, const void* k0, const void* v0, int kv_len0, const void* k1, const void* v1, int kv_len1,
                      void* const* peer_out, int peers, int my_rank, int ldo, int heads, int seq_q, float softmax_scale,
                      float out_scale, void* stream) {
  VP_REQUIRE(q && k0 && v0 && peer_out, VP_ERR_BAD_SHAPE, "attention_peer: null pointer");
  VP_REQUIRE(peers >= 1 && peers <= 8 && seq_q % peers == 0 && my_rank >= 0 && my_rank < peers, VP_ERR_BAD_SHAPE,
             "attention_peer: the query rows must divide over 1..8 peers");
  AttnParams p{};
  p.batch = 1; p.heads = heads; p.seq_q = seq_q;
  p.kv_len0 = kv_len0; p.kv_len1 = kv_len1;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.out = (__nv_bfloat16*)peer_out[my_rank]; p.ldo = ldo;
  p.out_scale = out_scale; p.accumulate = 0;
  for (int i = 0; i < peers; ++i) {
    VP_REQUIRE(peer_out[i] != nullptr, VP_ERR_BAD_SHAPE, "attention_peer: null peer pointer");
    p.peer_out[i] = (__nv_bfloat16*)peer_out[i];
  }
  p.peer_rows = seq_q / peers; p.peer_src = my_rank;
  return launch_attention(q, k0, v0, k1, v1, p, (cudaStream_t)stream);
}

int vp_step_end(const void* noise_pred, float guidance, const void* sample, const float* old_pred, const void* noise,
                float c_sqrt_alpha_bf, float c_sqrt_beta, float c_m0_bf, float c_m1, float c_m2, float c_m3, float c_mn_bf,
                int second_order, float* pred_out, float* prev_out, void* latents_out, const void* gt, const void* noise0,
                const void* mask, int chan, long long hw, float sa_bf, float sb_bf, int renoise, int mask_background, long long n,
                void* stream) {
  VP_REQUIRE(noise_pred && sample && noise && pred_out && latents_out && n > 0, VP_ERR_BAD_SHAPE, "step_end: null pointer");
  VP_REQUIRE(!second_order || old_pred, VP_ERR_BAD_SHAPE, "step_end: the second-order update needs old_pred");
  VP_REQUIRE(gt == nullptr || (mask && chan > 0 && hw > 0 && n % ((long long)chan * hw) == 0 && (!renoise || noise0)), VP_ERR_BAD_SHAPE,
             "step_end: replace_gt needs gt, mask [frames, 1, hw] and n = frames * chan * hw");
  StepEndParams p{};
  p.n = n; p.noise_pred = (const __nv_bfloat16*)noise_pred; p.guidance = guidance;
  p.sample = (const __nv_bfloat16*)sample; p.old_pred = old_pred; p.noise = (const __nv_bfloat16*)noise;
  p.c_sqrt_alpha_bf = c_sqrt_alpha_bf; p.c_sqrt_beta = c_sqrt_beta; p.c_m0_bf = c_m0_bf; p.c_m1 = c_m1; p.c_m2 = c_m2; p.c_m3 = c_m3;
  p.c_mn_bf = c_mn_bf; p.second_order = second_order;
  p.pred_out = pred_out; p.prev_out = prev_out; p.latents_out = (__nv_bfloat16*)latents_out;
  p.gt = (const __nv_bfloat16*)gt; p.noise0 = (const __nv_bfloat16*)noise0; p.mask = (const __nv_bfloat16*)mask;
  p.chan = chan; p.hw = hw; p.sa_bf = sa_bf; p.sb_bf = sb_bf; p.renoise = renoise; p.mask_background = mask_background;
  return launch_step_end(p, (cudaStream_t)stream);
}

int vp_peer_scatter(const void* src, void* const* peer_dst, int peers, int my_rank, long long bytes_per_peer, void* stream) {
  VP_REQUIRE(src && peer_dst && peers >= 1 && peers <= 8 && my_rank >= 0 && my_rank < peers && bytes_per_peer > 0, VP_ERR_BAD_SHAPE,
             "peer_scatter: bad arguments");
  VP_REQUIRE(bytes_per_peer % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0, VP_ERR_BAD_ALIGN,
             "peer_scatter: 16-byte granularity");
  for (int d = 0; d < peers; ++d) VP_REQUIRE(peer_dst[d] != nullptr, VP_ERR_BAD_SHAPE, "peer_scatter: null peer pointer");
  return launch_peer_scatter(src, peer_dst, peers, my_rank, bytes_per_peer, (cudaStream_t)stream);
}

int vp_peer_alloc(long long bytes, void** ptr, unsigned char* handle64) {
  VP_REQUIRE(bytes > 0 && ptr && handle64, VP_ERR_BAD_SHAPE, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  VP_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
  VP_CHECK_CUDA(cudaMemset(p, 0, (size_t)bytes));
  VP_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  VP_CHECK_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  *ptr = p;
  return VP_OK;
}

int vp_peer_open(const unsigned char* handle64, void** ptr) {
  VP_REQUIRE(handle64 && ptr, VP_ERR_BAD_SHAPE, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  VP_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));   // maps into the CURRENT device's context
  *ptr = p;
  return VP_OK;
}

int vp_peer_close(void* ptr) {
  VP_REQUIRE(ptr, VP_ERR_BAD_SHAPE, "peer_close: null pointer");
  VP_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return VP_OK;
}

int vp_peer_free(void* ptr) {
  VP_REQUIRE(ptr, VP_ERR_BAD_SHAPE, "peer_free: null pointer");
  VP_CHECK_CUDA(cudaFree(ptr));
  return VP_OK;
}

int vp_peer_set_timeout_ms(long long ms) { return set_peer_timeout_ms(ms); }

int vp_peer_barrier(void* const* peer_flags, int peers, int my_rank, unsigned int epoch, void* stream) {
  VP_REQUIRE(peer_flags && peers >= 1 && peers <= 8 && my_rank >= 0 && my_rank < peers, VP_ERR_BAD_SHAPE, "peer_barrier: bad arguments");
  uint32_t* f[8];
  for (int i = 0; i < peers; ++i) {
    VP_REQUIRE(peer_flags[i] != nullptr, VP_ERR_BAD_SHAPE, "peer_barrier: null flag pointer");
    f[i] = (uint32_t*)peer_flags[i];
  }
  return launch_peer_barrier(f, peers, my_rank, epoch, (cudaStream_t)stream);
}

int vp_a2a_unpack_heads(const void* src, void* dst0, void* dst1, void* dst2, void* dst3, void* dst4, int slots, int peers,
                        int heads_local, int rows_per_peer, void* stream) {
  VP_REQUIRE(src && dst0 && slots >= 1 && slots <= 5 && peers >= 1 && heads_local >= 1 && rows_per_peer >= 1, VP_ERR_BAD_SHAPE,
             "a2a_unpack_heads: bad arguments");
  void* dst[5] = {dst0, dst1, dst2, dst3, dst4};
  for (int i = 0; i < slots; ++i) VP_REQUIRE(dst[i] != nullptr, VP_ERR_BAD_SHAPE, "a2a_unpack_heads: null destination");
  return launch_a2a_unpack_heads(src, dst, slots, peers, heads_local, rows_per_peer, (cudaStream_t)stream);
}

int vp_patchify(const void* src0, int c0, const void* src1, int c1, int bf, int h, int w, void* out, int kpad, void* stream) {
  VP_REQUIRE(src0 && out && (c1 == 0 || src1), VP_ERR_BAD_SHAPE, "patchify: null pointer");
  return launch_patchify(src0, c0, src1, c1, bf, h, w, out, kpad, (cudaStream_t)stream);
}

int vp_mask_pool(const void* mask, int bf, int h, int w, uint8_t* out, void* stream) {
  VP_REQUIRE(mask && out, VP_ERR_BAD_SHAPE, "mask_pool: null pointer");
  return launch_mask_pool(mask, bf, h, w, out, (cudaStream_t)stream);
}

int vp_unpatchify(const void* proj, int bf, int c, int h, int w, void* out, void* stream) {
  VP_REQUIRE(proj && out, VP_ERR_BAD_SHAPE, "unpatchify: null pointer");
  return launch_unpatchify(proj, bf, c, h, w, out, (cudaStream_t)stream);
}

}  // extern "C"
