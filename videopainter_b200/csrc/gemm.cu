// tcgen05 GEMM for the token-wise linears of CogVideoXBlock:  C[M,N] = A[M,K] · W[N,K]ᵀ  (bf16 in, fp32 accumulate)
// with the follow-up elementwise work fused into the epilogue (see GemmEpilogue in gemm.cuh).
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 0   TMA producer   : A tile 128x64 and W tile 256x64 per stage, 4-stage mbarrier ring
//   warp 1   MMA issuer     : one thread issues tcgen05.mma 128x256x16, fp32 accumulators in TMEM
//   warp 2   TMEM allocator : 512 columns = two 128x256 accumulator stages
//   warp 4-7 epilogue       : tcgen05.ld (thread == row), fused epilogue, 16-byte global stores;
//                             overlaps with the MMAs of the next tile through the second TMEM stage
#include <stdlib.h>
#include <string.h>

#include "gemm.cuh"
#include "host_util.cuh"

namespace vp {

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int BAR_BYTES = 256;
constexpr int EPI_STAGE_BYTES = 32 * 128;                              // one head (64 bf16) of 32 rows, per epilogue warp
constexpr int SMEM_EPI = STAGES * STAGE_BYTES + BAR_BYTES;
constexpr int SMEM_BYTES = SMEM_EPI + 4 * EPI_STAGE_BYTES + 1024;     // +1024: manual alignment slack
constexpr int NUM_THREADS = 256;
constexpr uint32_t TMEM_COLS = 512;

__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int n_tiles, int group_m, int& m_blk, int& n_blk) {
  const int per_group = group_m * n_tiles;
  const int g = tile / per_group;
  const int first_m = g * group_m;
  const int gsz = min(group_m, m_tiles - first_m);
  const int r = tile - g * per_group;
  m_blk = first_m + r % gsz;
  n_blk = r / gsz;
}

__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + i);
    f[i * 8 + 0] = bf16_lo(u.x); f[i * 8 + 1] = bf16_hi(u.x);
    f[i * 8 + 2] = bf16_lo(u.y); f[i * 8 + 3] = bf16_hi(u.y);
    f[i * 8 + 4] = bf16_lo(u.z); f[i * 8 + 5] = bf16_hi(u.z);
    f[i * 8 + 6] = bf16_lo(u.w); f[i * 8 + 7] = bf16_hi(u.w);
  }
}

#ifndef VP_GEMM_STREAM_STORES
#define VP_GEMM_STREAM_STORES 0      // 1: epilogue stores carry the .cs (streaming) hint: outputs do not displace A / W in L2
#endif
__device__ __forceinline__ void st_global_v4(unsigned long long addr, const uint4& v) {
#if VP_GEMM_STREAM_STORES
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
  *reinterpret_cast<uint4*>(addr) = v;
#endif
}

// ---- QKV epilogue on one head (64 accumulator columns) of one row ----------------------------------------------
// Each thread owns one 128-byte output row (one head of one token).  Storing it directly would make every warp-wide
// store touch 32 different lines with 16 bytes each — harmless behind the local L2, but over NVLink (peer mode) every
// such piece travels as its own small packet.  So the warp transposes through 4 KB of XOR-swizzled shared memory
// (conflict-free both ways) and then writes whole 128-byte lines: 8 lanes per row, 4 rows per store instruction.
__device__ __forceinline__ void warp_store_rows128(uint8_t* stage, int lane, const uint4 (&c)[8], __nv_bfloat16* dst) {
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(stage + lane * 128 + ((i ^ (lane & 7)) << 4)) = c[i];
  __syncwarp();
  const unsigned long long d = reinterpret_cast<unsigned long long>(dst);     // 0 = this row is not stored
  const int cc = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3);
    const uint4 v = *reinterpret_cast<const uint4*>(stage + r * 128 + ((cc ^ (r & 7)) << 4));
    const unsigned long long dr = __shfl_sync(0xffffffffu, d, r);
    if (dr) st_global_v4(dr + (cc << 4), v);
  }
  __syncwarp();
}

// The same for 64-byte row pieces (32 bf16 columns per thread): 4 lanes per row, 8 rows per store instruction.
__device__ __forceinline__ void warp_store_rows64(uint8_t* stage, int lane, const uint4 (&c)[4], __nv_bfloat16* dst) {
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stage + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = c[i];
  __syncwarp();
  const unsigned long long d = reinterpret_cast<unsigned long long>(dst);     // 0 = this row is not stored
  const int cc = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    const uint4 v = *reinterpret_cast<const uint4*>(stage + r * 64 + ((cc ^ ((r >> 1) & 3)) << 4));
    const unsigned long long dr = __shfl_sync(0xffffffffu, d, r);
    if (dr) st_global_v4(dr + (cc << 4), v);
  }
  __syncwarp();
}

// The reverse: every thread wants 64 bytes (32 bf16) of ITS row (src == nullptr: zeros).  The warp fetches whole 64-byte row
// segments (4 lanes per row) and redistributes through the same swizzled staging buffer.
__device__ __forceinline__ void warp_load_rows64(uint8_t* stage, int lane, const __nv_bfloat16* src, float* f) {
  const unsigned long long s = reinterpret_cast<unsigned long long>(src);
  const int cc = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    const unsigned long long sr = __shfl_sync(0xffffffffu, s, r);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (sr) v = __ldg(reinterpret_cast<const uint4*>(sr + (cc << 4)));
    *reinterpret_cast<uint4*>(stage + r * 64 + ((cc ^ ((r >> 1) & 3)) << 4)) = v;
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = *reinterpret_cast<const uint4*>(stage + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4));
    f[i * 8 + 0] = bf16_lo(u.x); f[i * 8 + 1] = bf16_hi(u.x);
    f[i * 8 + 2] = bf16_lo(u.y); f[i * 8 + 3] = bf16_hi(u.y);
    f[i * 8 + 4] = bf16_lo(u.z); f[i * 8 + 5] = bf16_hi(u.z);
    f[i * 8 + 6] = bf16_lo(u.w); f[i * 8 + 7] = bf16_hi(u.w);
  }
  __syncwarp();
}

__device__ __forceinline__ void pack_bf16x32(const float* f, uint4* u) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    u[i].x = pack_bf16(f[i * 8 + 0], f[i * 8 + 1]);
    u[i].y = pack_bf16(f[i * 8 + 2], f[i * 8 + 3]);
    u[i].z = pack_bf16(f[i * 8 + 4], f[i * 8 + 5]);
    u[i].w = pack_bf16(f[i * 8 + 6], f[i * 8 + 7]);
  }
}

__device__ __forceinline__ void load_bf16x16(const __nv_bfloat16* p, float* f) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + i);
    f[i * 8 + 0] = bf16_lo(u.x); f[i * 8 + 1] = bf16_hi(u.x);
    f[i * 8 + 2] = bf16_lo(u.y); f[i * 8 + 3] = bf16_hi(u.y);
    f[i * 8 + 4] = bf16_lo(u.z); f[i * 8 + 5] = bf16_hi(u.z);
    f[i * 8 + 6] = bf16_lo(u.w); f[i * 8 + 7] = bf16_hi(u.w);
  }
}

// LayerNorm(64) of (x * pre) with affine (w, bia), optional interleaved-pair RoPE, packed to bf16 in `out` (8 x 16 bytes).
// RoPE (EMB:683-692: out = x*cos + rot(x)*sin, rot(x)[2i] = -x[2i+1], rot(x)[2i+1] = x[2i]) comes either from `cs` — this
// row's 32 (cos, sin) pairs, already in registers, valid when the tables repeat every value twice as the reference builds
// them (EMB:641-642) — or from the general per-element tables cosr / sinr.
__device__ __forceinline__ void ln64_rope_pack(const float* x, float pre, const __nv_bfloat16* w, const __nv_bfloat16* bia,
                                               float eps, const float* cs, const float* cosr, const float* sinr, uint4 (&out)[8]) {
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;             // four partial sums: no 64-long dependency chain
#pragma unroll
  for (int j = 0; j < 64; j += 4) { m0 += x[j]; m1 += x[j + 1]; m2 += x[j + 2]; m3 += x[j + 3]; }
  const float mean = ((m0 + m1) + (m2 + m3)) * pre * (1.0f / 64.0f);
  float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
  for (int j = 0; j < 64; j += 4) {
    const float d0 = fmaf(x[j], pre, -mean), d1 = fmaf(x[j + 1], pre, -mean), d2 = fmaf(x[j + 2], pre, -mean), d3 = fmaf(x[j + 3], pre, -mean);
    v0 = fmaf(d0, d0, v0); v1 = fmaf(d1, d1, v1); v2 = fmaf(d2, d2, v2); v3 = fmaf(d3, d3, v3);
  }
  const float rstd = rsqrtf(((v0 + v1) + (v2 + v3)) * (1.0f / 64.0f) + eps);
  const float a = pre * rstd, nb = -mean * rstd;
#pragma unroll
  for (int c = 0; c < 4; ++c) {   // 16 columns at a time keeps the live set small
    float g[16], be[16], y[16];
    load_bf16x16(w + c * 16, g);
    load_bf16x16(bia + c * 16, be);
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = fmaf(fmaf(x[c * 16 + j], a, nb), g[j], be[j]);
    if (cs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float co = cs[(c * 8 + j) * 2], si = cs[(c * 8 + j) * 2 + 1];
        const float a0 = y[2 * j], a1 = y[2 * j + 1];
        y[2 * j] = fmaf(a0, co, -a1 * si);
        y[2 * j + 1] = fmaf(a1, co, a0 * si);
      }
    } else if (cosr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 cv = __ldg(reinterpret_cast<const float4*>(cosr + c * 16) + j);
        float4 sv = __ldg(reinterpret_cast<const float4*>(sinr + c * 16) + j);
        const float a0 = y[j * 4 + 0], a1 = y[j * 4 + 1], a2 = y[j * 4 + 2], a3 = y[j * 4 + 3];
        y[j * 4 + 0] = a0 * cv.x - a1 * sv.x;
        y[j * 4 + 1] = a1 * cv.y + a0 * sv.y;
        y[j * 4 + 2] = a2 * cv.z - a3 * sv.z;
        y[j * 4 + 3] = a3 * cv.w + a2 * sv.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      out[c * 2 + i].x = pack_bf16(y[i * 8 + 0], y[i * 8 + 1]);
      out[c * 2 + i].y = pack_bf16(y[i * 8 + 2], y[i * 8 + 3]);
      out[c * 2 + i].z = pack_bf16(y[i * 8 + 4], y[i * 8 + 5]);
      out[c * 2 + i].w = pack_bf16(y[i * 8 + 6], y[i * 8 + 7]);
    }
  }
}

// Runs warp-wide (all 32 lanes, also those whose row is out of range: row_ok = false stores nothing).
__device__ __forceinline__ void epilogue_qkv_head(const GemmParams& p, const uint32_t* acc, bool row_ok, int m, int b, int s, int n0,
                                                  uint8_t* stage, int lane, const float* cs) {
  float x[64];
  load_bf16x32(p.bias + n0, x);
  load_bf16x32(p.bias + n0 + 32, x + 32);
#pragma unroll
  for (int j = 0; j < 64; ++j) x[j] += __uint_as_float(acc[j]);
  const int which = n0 / p.d_model + p.qkv_first;      // 0 = Q, 1 = K, 2 = V
  const int head = (n0 % p.d_model) >> 6;
  const float rs = (p.row_scale && row_ok) ? p.row_scale[m] : 1.0f;
  const int dest = head / p.heads_per_dest, hl = head - dest * p.heads_per_dest;
  long long off = dest * p.dest_stride + (((long long)b * p.heads_per_dest + hl) * p.rows_per_batch + s) * 64;
  if (p.peer_base[0]) {           // store into the destination rank's memory (P2P): offset relative to the local twin buffer
    off = (p.peer_base[dest] - p.local_base) + ((long long)hl * p.peer_seq + p.peer_row_off + s) * 64;
  }
  uint4 out[8];
  if (which == 2) {
    float y[32];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int j = 0; j < 32; ++j) y[j] = x[c * 32 + j] * rs;
      pack_bf16x32(y, &out[c * 4]);
    }
    warp_store_rows128(stage, lane, out, row_ok ? p.v_out + off : nullptr);
    if (p.v2_out) {
      const float mk = (row_ok && p.mask2[m]) ? 1.0f : 0.0f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = x[c * 32 + j] * mk;
        pack_bf16x32(y, &out[c * 4]);
      }
      warp_store_rows128(stage, lane, out, row_ok ? p.v2_out + off : nullptr);
    }
    return;
  }
  const bool rope = cs == nullptr && p.rope_cos != nullptr && s >= p.text_len && row_ok;
  const float* cosr = rope ? p.rope_cos + (long long)(s - p.text_len) * 64 : nullptr;
  const float* sinr = rope ? p.rope_sin + (long long)(s - p.text_len) * 64 : nullptr;
  if (which == 0) {
    ln64_rope_pack(x, rs, p.nq_w, p.nq_b, p.qk_eps, cs, cosr, sinr, out);
    warp_store_rows128(stage, lane, out, row_ok ? p.q_out + off : nullptr);
  } else {
    ln64_rope_pack(x, rs, p.nk_w, p.nk_b, p.qk_eps, cs, cosr, sinr, out);
    warp_store_rows128(stage, lane, out, row_ok ? p.k_out + off : nullptr);
    if (p.k2_out) {
      const float mk = (row_ok && p.mask2[m]) ? 1.0f : 0.0f;    // masked-out keys become RoPE(norm_k.bias): AP:2255, 2272, 2281
      ln64_rope_pack(x, mk, p.nk_w, p.nk_b, p.qk_eps, cs, cosr, sinr, out);
      warp_store_rows128(stage, lane, out, row_ok ? p.k2_out + off : nullptr);
    }
  }
}

// ---- plain / gelu / residual epilogue on 32 accumulator columns of one row -------------------------------------
// Runs warp-wide (row_ok = false lanes compute on zeros and store nothing); the 64-byte row pieces leave through the
// swizzled shared-memory transpose so that every store instruction writes whole 64-byte row segments.
template <int EPI>
__device__ __forceinline__ void epilogue_cols32(const GemmParams& p, const uint32_t* acc, bool row_ok, int b, int s, int n0,
                                                uint8_t* stage, int lane) {
  float v[32];
  if (p.bias) {
    load_bf16x32(p.bias + n0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(acc[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  }
  if (EPI == EPI_BIAS) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
  } else if (EPI == EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
  } else if (EPI == EPI_RESID) {
    if (p.gate && row_ok) {
      const float* g = p.gate + (long long)b * p.gate_batch_stride + (s < p.text_len ? p.gate_text_off : p.gate_video_off) + n0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 gv = __ldg(reinterpret_cast<const float4*>(g) + j);
        v[j * 4 + 0] *= gv.x; v[j * 4 + 1] *= gv.y; v[j * 4 + 2] *= gv.z; v[j * 4 + 3] *= gv.w;
      }
    }
    float r[32];
    warp_load_rows64(stage, lane, row_ok ? p.res + ((long long)b * p.res_batch_rows + p.res_row_offset + s) * p.ldr + n0 : nullptr, r);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += r[j];
    if (p.inject) {                                        // warp-uniform; rows that take no injection fetch zeros
      const int sv = s - p.text_len;
      const bool take = row_ok && s >= p.text_len && (!p.inject_mask || p.inject_mask[(long long)b * p.video_len + sv] == 0);
      warp_load_rows64(stage, lane, take ? p.inject + (long long)b * p.inject_batch_stride + (long long)sv * p.ldi + n0 : nullptr, r);
      if (take) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += r[j];
      }
    }
  }
  uint4 u[4];
  pack_bf16x32(v, u);
  warp_store_rows64(stage, lane, u, row_ok ? p.out + ((long long)b * p.out_batch_rows + p.out_row_offset + s) * p.ldo + n0 : nullptr);
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The producer and MMA loops run warp-wide (uniform control flow) and only the asynchronous instructions themselves
  // are issued by one elected lane: descriptor / coordinate arithmetic then stays on the uniform datapath, which is what
  // keeps the tcgen05.mma issue rate at the tensor pipe's own rate (a single divergent thread costs ~100 clk per MMA).
  if (warp == 0) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, m_tiles, n_tiles, p.group_m, m_blk, n_blk);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
          const int k0 = kb * BK, chunk = k0 / p.a_k_chunk;
          tma_load_3d(sa, &tmap_a, &full[stage], k0 - chunk * p.a_k_chunk, m_blk * BM, chunk, kEvictNormal);
          tma_load_2d(sb, &tmap_b, &full[stage], kb * BK, n_blk * BN, kEvictLast);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = make_desc_sw128(sa, 1024, 0);
          const uint64_t bdesc = make_desc_sw128(sa + A_BYTES, 1024, 0);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes per 16-element K step inside the 128-byte swizzle atom (descriptor address unit = 16 B)
            mma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          tc_commit(&empty[stage]);   // frees the smem stage once these MMAs have read it
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) tc_commit(&tfull[acc]);       // accumulator complete -> epilogue
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue ----------------------------------------------------
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int m_blk, n_blk;
      tile_coords(tile, m_tiles, n_tiles, p.group_m, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int m = m_blk * BM + row_in_tile;
      int b = 0, s = 0;
      bool row_ok = m < p.M;
      if (row_ok) {
        b = m / p.rows_per_batch;
        s = m - b * p.rows_per_batch;
        if (EPI != EPI_QKV && s + p.out_row_offset < 0) row_ok = false;
      }
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      if (EPI == EPI_QKV) {
        // the four heads of this tile share the token row: its RoPE pairs are fetched once (compact table, 256 B per row)
        float cs[64];
        const bool use_cs = p.rope_cs != nullptr && (n_blk * BN) / p.d_model + p.qkv_first != 2 && row_ok && s >= p.text_len;
        if (use_cs) {
          const float4* src = reinterpret_cast<const float4*>(p.rope_cs + (long long)(s - p.text_len) * 64);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 t4 = __ldg(src + i);
            cs[i * 4 + 0] = t4.x; cs[i * 4 + 1] = t4.y; cs[i * 4 + 2] = t4.z; cs[i * 4 + 3] = t4.w;
          }
        }
#pragma unroll 1
        for (int hc = 0; hc < BN / 64; ++hc) {
          const int n0 = n_blk * BN + hc * 64;
          if (n0 >= p.N) break;
          uint32_t r[64];
          tmem_ld_x32(taddr + hc * 64, r);
          tmem_ld_x32(taddr + hc * 64 + 32, r + 32);
          tmem_wait_ld();
          epilogue_qkv_head(p, r, row_ok, m, b, s, n0, smem + SMEM_EPI + (warp - 4) * EPI_STAGE_BYTES, lane, use_cs ? cs : nullptr);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= p.N) break;
          uint32_t r[32];
          tmem_ld_x32(taddr + c * 32, r);
          tmem_wait_ld();
          epilogue_cols32<EPI>(p, r, row_ok, b, s, n0, smem + SMEM_EPI + (warp - 4) * EPI_STAGE_BYTES, lane);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair kernel (the default; VP_B200_GEMM=single selects the one-CTA kernel above): two CTAs of a cluster (one TPC)
// compute a 256 x 256 tile with
// tcgen05.mma.cta_group::2.  Each CTA loads ITS 128 rows of A and ITS 128-row half of the W tile (32 KB per stage instead of
// 48 KB: one third less TMA / L2 traffic and shared-memory fill per CTA, six stages instead of four); the tensor cores of
// both SMs read the two halves of W from both shared memories.  Only the leader (cluster rank 0) issues MMAs; its `full`
// barriers collect the bytes of both CTAs' loads (cp.async.bulk.tensor ... cta_group::2 signals the leader's barrier), the
// commits are multicast to the `empty` / `tfull` barriers of both CTAs, and the epilogue warps of both CTAs (each owns its
// 128 accumulator rows in its own TMEM) arrive on the leader's `tempty`.  Same K order per accumulator as the one-CTA kernel:
// results are bit-identical.  Measured (B200, production shapes, stand-alone): QKV 1431 -> 1544, out-projection 1510 -> 1606,
// FFN-1 1456 -> 1552 (cuBLAS 1545), FFN-2 1389 -> 1397 TFLOP/s; inside the power-capped step +8 % on every GEMM, step -2.2 %.
// ------------------------------------------------------------------------------------------------------------------
constexpr int STAGES2 = 6;
constexpr int BH_BYTES = (BN / 2) * BK * 2;                          // this CTA's half of the W tile
constexpr int STAGE2_BYTES = A_BYTES + BH_BYTES;
constexpr int SMEM2_EPI = STAGES2 * STAGE2_BYTES + BAR_BYTES;
constexpr int SMEM2_BYTES = SMEM2_EPI + 4 * EPI_STAGE_BYTES + 1024;
static_assert((3 * STAGES2 + 4) * 8 + 8 <= BAR_BYTES, "barrier block too small");

__device__ __forceinline__ uint32_t cluster_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {      // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {      // shared::cta address -> shared::cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                                 int c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {   // the same warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on the barrier at this shared-memory offset in the CTAs of `mask` once every MMA issued so far has completed
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mma_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES2 * STAGE2_BYTES);     // used in the leader only
  uint64_t* empty = full + STAGES2;
  uint64_t* tfull = empty + STAGES2;
  uint64_t* tempty = tfull + 2;                                                     // used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_cta_rank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM);                               // 256-row tiles
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + BK - 1) / BK;
  const int group_m = p.group_m > 1 ? p.group_m / 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 256);                    // the epilogue threads of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();                                // barriers of both CTAs exist before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (both CTAs) ------------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m_blk, n_blk;
      tile_coords(tile, m_tiles, n_tiles, group_m, m_blk, n_blk);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * STAGE2_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * STAGE2_BYTES);
          const uint32_t bar = map_to_cta(smem_u32(&full[stage]), 0);
          const int k0 = kb * BK, chunk = k0 / p.a_k_chunk;
          tma_load_3d_pair(sa, &tmap_a, bar, k0 - chunk * p.a_k_chunk, (m_blk * 2 + (int)rank) * BM, chunk, kEvictNormal);
          tma_load_2d_pair(sb, &tmap_b, bar, kb * BK, n_blk * BN + (int)rank * (BN / 2), kEvictLast);
        }
        __syncwarp();
        if (++stage == STAGES2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA) -------------------------------------
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * STAGE2_BYTES);
            const uint64_t adesc = make_desc_sw128(sa, 1024, 0);
            const uint64_t bdesc = make_desc_sw128(sa + A_BYTES, 1024, 0);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) mma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            tc_commit_pair(&empty[stage], 3);        // both CTAs may refill the stage
          }
          __syncwarp();
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit_pair(&tfull[acc], 3);     // accumulators complete in both CTAs' TMEM
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue (both CTAs: own 128 rows) --------------------------
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      int m_blk, n_blk;
      tile_coords(tile, m_tiles, n_tiles, group_m, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int m = (m_blk * 2 + (int)rank) * BM + row_in_tile;
      int b = 0, s = 0;
      bool row_ok = m < p.M;
      if (row_ok) {
        b = m / p.rows_per_batch;
        s = m - b * p.rows_per_batch;
        if (EPI != EPI_QKV && s + p.out_row_offset < 0) row_ok = false;
      }
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      if (EPI == EPI_QKV) {
        float cs[64];
        const bool use_cs = p.rope_cs != nullptr && (n_blk * BN) / p.d_model + p.qkv_first != 2 && row_ok && s >= p.text_len;
        if (use_cs) {
          const float4* src = reinterpret_cast<const float4*>(p.rope_cs + (long long)(s - p.text_len) * 64);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 t4 = __ldg(src + i);
            cs[i * 4 + 0] = t4.x; cs[i * 4 + 1] = t4.y; cs[i * 4 + 2] = t4.z; cs[i * 4 + 3] = t4.w;
          }
        }
#pragma unroll 1
        for (int hc = 0; hc < BN / 64; ++hc) {
          const int n0 = n_blk * BN + hc * 64;
          if (n0 >= p.N) break;
          uint32_t r[64];
          tmem_ld_x32(taddr + hc * 64, r);
          tmem_ld_x32(taddr + hc * 64 + 32, r + 32);
          tmem_wait_ld();
          epilogue_qkv_head(p, r, row_ok, m, b, s, n0, smem + SMEM2_EPI + (warp - 4) * EPI_STAGE_BYTES, lane, use_cs ? cs : nullptr);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= p.N) break;
          uint32_t r[32];
          tmem_ld_x32(taddr + c * 32, r);
          tmem_wait_ld();
          epilogue_cols32<EPI>(p, r, row_ok, b, s, n0, smem + SMEM2_EPI + (warp - 4) * EPI_STAGE_BYTES, lane);
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(map_to_cta(smem_u32(&tempty[acc]), 0));
    }
  }

  tc_fence_before();
  cluster_sync_all();                                // nobody leaves while its shared memory / TMEM is still a target
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

template <int EPI>
int launch_pair_impl(const CUtensorMap& ta, const CUtensorMap& tb_half, const GemmParams& p, cudaStream_t st) {
  const int rc_cfg = configure_once(reinterpret_cast<const void*>(gemm_pair_kernel<EPI>), SMEM2_BYTES);
  if (rc_cfg) return rc_cfg;
  const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM), n_tiles = (p.N + BN - 1) / BN;
  const int sms = sm_count();
  if (sms <= 0) return fail(VP_ERR_CUDA, "no CUDA device");
  const int pairs = m_tiles * n_tiles < sms / 2 ? m_tiles * n_tiles : sms / 2;
  gemm_pair_kernel<EPI><<<2 * pairs, NUM_THREADS, SMEM2_BYTES, st>>>(ta, tb_half, p);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

template <int EPI>
int launch_impl(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  const int rc_cfg = configure_once(reinterpret_cast<const void*>(gemm_bf16_kernel<EPI>), SMEM_BYTES);
  if (rc_cfg) return rc_cfg;
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  const int sms = sm_count();
  if (sms <= 0) return fail(VP_ERR_CUDA, "no CUDA device");
  const int grid = m_tiles * n_tiles < sms ? m_tiles * n_tiles : sms;
  gemm_bf16_kernel<EPI><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ta, tb, p);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

}  // namespace

int launch_gemm(int epi, const void* A, long long lda, const void* W, long long ldw, const GemmParams& p, cudaStream_t st) {
  VP_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, VP_ERR_BAD_SHAPE, "gemm: empty problem");
  VP_REQUIRE(p.N % 64 == 0, VP_ERR_BAD_SHAPE, "gemm: N must be a multiple of 64");
  VP_REQUIRE(p.K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, VP_ERR_BAD_ALIGN, "gemm: K / leading dims must be multiples of 8");
  VP_REQUIRE(p.rows_per_batch > 0, VP_ERR_BAD_SHAPE, "gemm: rows_per_batch");
  GemmParams q = p;
  if (q.a_k_chunk <= 0 || q.a_k_chunk >= q.K) {
    q.a_k_chunk = q.K;
    q.a_chunk_stride = (long long)q.K;       // unused: a single chunk
  }
  VP_REQUIRE(q.K % q.a_k_chunk == 0 && (q.a_k_chunk == q.K || q.a_k_chunk % BK == 0) && q.a_chunk_stride % 8 == 0, VP_ERR_BAD_SHAPE,
             "gemm: K chunks of A must be multiples of 64 columns that divide K");
  CUtensorMap ta, tb;
  {
    uint64_t dims[3] = {(uint64_t)q.a_k_chunk, (uint64_t)p.M, (uint64_t)(q.K / q.a_k_chunk)};
    uint64_t str[2] = {(uint64_t)lda * 2, (uint64_t)q.a_chunk_stride * 2};
    uint32_t box[3] = {BK, BM, 1};
    int rc = make_tmap_bf16(&ta, A, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {BK, BN};
    int rc = make_tmap_bf16(&tb, W, 2, dims, str, box);
    if (rc) return rc;
  }
  if (q.group_m <= 0) q.group_m = 16;
  {
    static const int forced = []() { const char* e = getenv("VP_GEMM_GROUP_M"); return e ? atoi(e) : 0; }();   // tuning aid
    if (forced > 0) q.group_m = forced;
  }
  if (q.heads_per_dest <= 0) q.heads_per_dest = q.heads > 0 ? q.heads : 1;
  static const bool use_pair = []() { const char* e = getenv("VP_B200_GEMM"); return !(e && !strcmp(e, "single")); }();   // default
  if (use_pair) {
    CUtensorMap tbh;
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {BK, BN / 2};
    int rc = make_tmap_bf16(&tbh, W, 2, dims, str, box);
    if (rc) return rc;
    switch (epi) {
      case EPI_BIAS: return launch_pair_impl<EPI_BIAS>(ta, tbh, q, st);
      case EPI_GELU: return launch_pair_impl<EPI_GELU>(ta, tbh, q, st);
      case EPI_RESID: return launch_pair_impl<EPI_RESID>(ta, tbh, q, st);
      case EPI_QKV: return launch_pair_impl<EPI_QKV>(ta, tbh, q, st);
      default: return fail(VP_ERR_UNSUPPORTED, "gemm: unknown epilogue");
    }
  }
  switch (epi) {
    case EPI_BIAS: return launch_impl<EPI_BIAS>(ta, tb, q, st);
    case EPI_GELU: return launch_impl<EPI_GELU>(ta, tb, q, st);
    case EPI_RESID: return launch_impl<EPI_RESID>(ta, tb, q, st);
    case EPI_QKV: return launch_impl<EPI_QKV>(ta, tb, q, st);
    default: return fail(VP_ERR_UNSUPPORTED, "gemm: unknown epilogue");
  }
}

}  // namespace vp
