// HBM-bound token-wise kernels of the denoising step: expert adaLN-zero modulation (NRM:373-379), the small
// conditioning GEMVs (timestep MLP EMB:762-774, adaLN tables NRM:376, norm_out NRM:73), patch gathering for the
// 2x2/stride-2 patch-embed convolution (EMB:408-414), mask pooling (EMB:417-426), the final double LayerNorm
// (T3D:613-624) and unpatchify (T3D:630-632).
#include "elementwise.cuh"
#include <string.h>

#include <stdlib.h>

#include <atomic>
#include "host_util.cuh"

namespace vp {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
  u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm + expert modulation:  y = LN(x; gamma, beta, eps) * (1 + scale[b, e]) + shift[b, e],  e = text if s < text_len
// else video.  HBM-bound (read + write of [rows, D] bf16), so the kernel is organised around memory traffic:
//  * thread t owns columns [8t, 8t + 8) of EVERY row its CTA touches, hence the per-column coefficients
//      A = gamma (1 + scale),  C = beta (1 + scale) + shift      (fp32, one pair per (batch, expert))
//    sit in 16 registers (the earlier warp-per-row version re-read 36 KB of parameters per 6 KB row through L1 and was
//    LSU-bound at 47 % of the HBM roofline);
//  * persistent CTAs walk units of LN_RB consecutive rows of one (batch, expert) segment; a unit is one contiguous
//    block of x, fetched by a single bulk async copy (cp.async.bulk -> mbarrier) into a shared-memory ring LN_STAGES
//    deep, so ~100 KB per CTA are in flight without costing registers (HBM latency under load is ~2 us);
//  * row statistics are fp32, two-pass on registers, reduced warp-shuffle -> shared memory -> one warp per row;
//    results leave with 16-byte stores straight from registers.
// ------------------------------------------------------------------------------------------------------------------
#ifndef VP_LN_RB
#define VP_LN_RB 4
#define VP_LN_CTAS 2
#endif
constexpr int LN_RB = VP_LN_RB;          // rows per unit
constexpr int LN_CTAS = VP_LN_CTAS;      // resident CTAs per SM the register / shared-memory budget is set for
constexpr int LN_MAX_WARPS = 16;         // D <= 4096
constexpr int LN_MAX_STAGES = 6;
constexpr int LN_SMEM_BUDGET = 108 * 1024;   // per CTA (two CTAs share the 227 KB of an SM)

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// v[r] = this thread's partial of row r  ->  stat[r] = sum over the CTA.  `between` runs after the first barrier (every
// thread has passed the code before the call).  Hazards: `red` of one statistic is rewritten only after two further
// barriers; `stat` is rewritten by the one-warp-per-row step of the same statistic in the next unit, which follows that
// unit's first barrier, which every thread reaches only after its last read of `stat` in this unit.
template <typename F>
__device__ __forceinline__ void ln_block_stat(float (&v)[LN_RB], float* red, float* stat, int warp, int lane, int nwarps, F between) {
  if (LN_RB == 4) {
    // four rows reduced together in 6 shuffles instead of 20: lanes trade rows pairwise (16, 8), then fold (4, 2, 1);
    // lanes [8r, 8r + 8) end up with the warp's sum of row r
    const bool hi16 = lane & 16, hi8 = lane & 8;
    float a0 = hi16 ? v[2] : v[0], a1 = hi16 ? v[3] : v[1];
    const float s0 = hi16 ? v[0] : v[2], s1 = hi16 ? v[1] : v[3];
    a0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    a1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    float k = hi8 ? a1 : a0;
    k += __shfl_xor_sync(0xffffffffu, hi8 ? a0 : a1, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    if ((lane & 7) == 0) red[(lane >> 3) * LN_MAX_WARPS + warp] = k;
  } else {
#pragma unroll
    for (int r = 0; r < LN_RB; ++r) v[r] = warp_sum(v[r]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < LN_RB; ++r) red[r * LN_MAX_WARPS + warp] = v[r];
    }
  }
  __syncthreads();
  between();
  for (int r = warp; r < LN_RB; r += nwarps) {
    float t = lane < nwarps ? red[r * LN_MAX_WARPS + lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) stat[r] = t;
  }
  __syncthreads();
}

template <int WARPS>   // upper bound of the block size: sets the register budget (LN_CTAS CTAs per SM)
__global__ void __launch_bounds__(WARPS * 32, LN_CTAS) ln_modulate_kernel(const LnModParams p, int units_text, int units_video,
                                                                          int stages) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  __shared__ float red[2][LN_RB * LN_MAX_WARPS];
  __shared__ float stat[2][LN_RB];
  __shared__ uint64_t full[LN_MAX_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int col = threadIdx.x * 8;
  const bool active = col < p.D;
  const int batch = (int)(p.rows / p.rows_per_batch);
  const int units_per_batch = units_text + units_video;
  const int total_units = batch * units_per_batch;
  const float invD = 1.0f / (float)p.D;
  const uint32_t row_bytes = (uint32_t)p.D * 2u, stage_bytes = row_bytes * LN_RB;

  // Each CTA walks a contiguous range of units; (batch, index within the batch) cursors advance without divisions.
  const int u_begin = (int)((long long)blockIdx.x * total_units / gridDim.x);
  const int u_end = (int)((long long)(blockIdx.x + 1) * total_units / gridDim.x);
  auto rows_of = [&](int k, int& text, int& s0, int& n) {       // unit k of a batch -> expert, first row, row count
    text = k < units_text ? 1 : 0;
    s0 = text ? k * LN_RB : p.text_len + (k - units_text) * LN_RB;
    n = min(LN_RB, (text ? p.text_len : p.rows_per_batch) - s0);
  };
  auto fetch = [&](int b, int k, int stage) {                   // one thread: the unit's rows are contiguous in x
    int text, s0, n;
    rows_of(k, text, s0, n);
    const __nv_bfloat16* x = p.x + ((long long)b * p.x_batch_rows + p.x_row_offset + s0) * p.D;
    mbar_arrive_expect_tx(&full[stage], (uint32_t)n * row_bytes);
    bulk_load(ln_smem + (size_t)stage * stage_bytes, x, (uint32_t)n * row_bytes, &full[stage]);
  };

  int cb = u_begin / units_per_batch, ck = u_begin - cb * units_per_batch;        // cursor of the unit being processed
  int fb = cb, fk = ck, fu = u_begin;                                              // cursor of the next unit to fetch
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages && fu < u_end; ++i, ++fu) {
      fetch(fb, fk, i);
      if (++fk == units_per_batch) { fk = 0; ++fb; }
    }
  }

  float A[8], C[8];
  int cur_b = -1, cur_text = -1;
  int stage = 0;
  uint32_t phase = 0;
  for (int u = u_begin; u < u_end; ++u) {
    const int b = cb;
    int text, s0, n;
    rows_of(ck, text, s0, n);
    if (++ck == units_per_batch) { ck = 0; ++cb; }
    if (active && (b != cur_b || text != cur_text)) {           // new (batch, expert): rebuild the coefficient registers
      cur_b = b; cur_text = text;
      float g[8], be[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.gamma + col)), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.beta + col)), be);
      if (p.mod) {
        const float* sh = p.mod + (long long)b * p.mod_batch_stride + (text ? p.shift_text_off : p.shift_video_off) + col;
        const float* sc = p.mod + (long long)b * p.mod_batch_stride + (text ? p.scale_text_off : p.scale_video_off) + col;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float one_plus = 1.f + __ldg(sc + e);
          A[e] = g[e] * one_plus;
          C[e] = fmaf(be[e], one_plus, __ldg(sh + e));
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) { A[e] = g[e]; C[e] = be[e]; }
      }
    }
    __nv_bfloat16* y = p.y + ((long long)b * p.rows_per_batch + s0) * p.D + col;

    while (!mbar_try_wait(&full[stage], phase)) {}               // a bulk copy always completes: no watchdog needed here
    float v[LN_RB][8];                                             // unpacked once; three passes read them
    float st[LN_RB];
    const uint8_t* src = ln_smem + (size_t)stage * stage_bytes + (size_t)col * 2;
#pragma unroll
    for (int r = 0; r < LN_RB; ++r) {
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      if (active && r < n) raw = *reinterpret_cast<const uint4*>(src + (size_t)r * row_bytes);
      unpack8(raw, v[r]);
      st[r] = ((v[r][0] + v[r][1]) + (v[r][2] + v[r][3])) + ((v[r][4] + v[r][5]) + (v[r][6] + v[r][7]));
    }
    // after the first barrier of this reduction every thread holds its part of the stage in registers: refill the stage
    ln_block_stat(st, red[0], stat[0], warp, lane, nwarps, [&]() {
      if (threadIdx.x == 0 && fu < u_end) {
        fetch(fb, fk, stage);
        ++fu;
        if (++fk == units_per_batch) { fk = 0; ++fb; }
      }
    });
#pragma unroll
    for (int r = 0; r < LN_RB; ++r) {
      const float mean = stat[0][r] * invD;
      float q = 0.f;
      if (active) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = v[r][e] - mean; q = fmaf(d, d, q); }
      }
      st[r] = q;
    }
    ln_block_stat(st, red[1], stat[1], warp, lane, nwarps, []() {});
#pragma unroll
    for (int r = 0; r < LN_RB; ++r) {
      if (active && r < n) {
        const float rstd = rsqrtf(stat[1][r] * invD + p.eps);
        const float nmr = -stat[0][r] * invD * rstd;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(fmaf(v[r][e], rstd, nmr), A[e], C[e]);
        *reinterpret_cast<uint4*>(y + (long long)r * p.D) = pack8(o);
      }
    }
    if (++stage == stages) { stage = 0; phase ^= 1; }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Second organisation of the same operation (the default; VP_B200_LN=cols selects the column-owner kernel above): a WARP
// owns LNW_ROWS whole rows, so the two row statistics are warp shuffles — no block-wide reduction and no __syncthreads in
// the row loop (the column-owner kernel spends ~195 instructions per row and thread around its four barriers per unit and
// is issue-bound at 66 % of the HBM rate; here a row costs ~700 warp instructions instead of ~2300).
//   * every warp has its own ring of LNW stages in shared memory, filled by bulk async copies (one contiguous copy per
//     unit of LNW_ROWS rows) that one lane issues — the next unit streams in while the current one is processed, without
//     costing registers (holding the rows in registers was tried: ptxas keeps the unpacked copy of every pass alive, 252
//     registers for two rows);
//   * the three passes (sum; centred squares; output) read the packed rows from shared memory, 16 bytes per lane;
//   * the per-column coefficients A = gamma (1 + scale), C = beta (1 + scale) + shift of the CTA's (batch, expert) segment
//     are built once per CTA in shared memory (fp32, laid out so that a lane's two float4 per 16-byte column group are
//     conflict-free) and read once per LNW_ROWS rows — the round-1 warp-per-row kernel re-read 36 KB of parameters through
//     L1 per 6 KB row.  CTAs are assigned to (batch, expert) segments in proportion to their rows.
// Measured at [2 x 17 776, 3072] (B200, L2 flushed, a plain device copy of the same bytes: 73.7 us = 5.93 TB/s):
//   column-owner kernel 88.1 us;  this kernel with 8 warps x 2 rows x 2 stages 98.3 us (ncu: 2 warps per scheduler, one
//   instruction per 5.2 clk and warp: latency-bound),  16 warps x 2 rows x 1 stage 90.1 us,  16 warps x 1 row x 2 stages
//   81.9 us = 5.33 TB/s (the default);  inside the power-capped step 8.21 ms for 440 launches vs 8.78 ms (4.68 vs 4.38 TB/s).
// Shared-memory traffic per row at D = 3072: 6 KB written by the copy, 18 KB read by the passes, 24 KB / LNW_ROWS of
// coefficients.
// ------------------------------------------------------------------------------------------------------------------
#ifndef VP_LNW_ROWS
#define VP_LNW_ROWS 1
#define VP_LNW_WARPS 16
#endif
constexpr int LNW_ROWS = VP_LNW_ROWS;     // rows per unit (one bulk copy, one read of the coefficients)
constexpr int LNW_WARPS = VP_LNW_WARPS;
constexpr int LNW_MAX_STAGES = 3;
constexpr int LNW_MAX_SEGS = 16;
constexpr int LNW_SMEM_BUDGET = 222 * 1024;
struct LnSegs {
  int n;
  int cta0[LNW_MAX_SEGS + 1];      // first CTA of segment i; cta0[n] = grid
  int batch[LNW_MAX_SEGS], text[LNW_MAX_SEGS], row0[LNW_MAX_SEGS], nrows[LNW_MAX_SEGS];
};

__device__ __forceinline__ uint64_t lnw_pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void lnw_unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;\n" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t lnw_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t lnw_add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;\n" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// one 32-bit word of two bf16 -> (low element, high element) as an fp32 pair
__device__ __forceinline__ uint64_t lnw_pair(uint32_t w) { return lnw_pack2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }

template <int CH>   // 16-byte column groups per lane: D <= CH * 256
__global__ void __launch_bounds__(LNW_WARPS * 32, 1) ln_rows_kernel(const LnModParams p, const LnSegs segs, int stages) {
  extern __shared__ __align__(128) uint8_t lnw_smem[];               // A | C (fp32, permuted) | ring [warp][stage][rows][D] bf16
  __shared__ uint64_t full[LNW_WARPS][LNW_MAX_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = p.D, nchunk = D >> 3, halfD = D >> 1;
  const uint32_t row_bytes = (uint32_t)D * 2u, stage_bytes = row_bytes * LNW_ROWS;
  int seg = 0;
  while (seg + 1 < segs.n && (int)blockIdx.x >= segs.cta0[seg + 1]) ++seg;
  const int b = segs.batch[seg], text = segs.text[seg], seg_row0 = segs.row0[seg], nrows = segs.nrows[seg];
  const int cta_in_seg = blockIdx.x - segs.cta0[seg], ctas_in_seg = segs.cta0[seg + 1] - segs.cta0[seg];
  float* tA = reinterpret_cast<float*>(lnw_smem);
  float* tC = tA + D;
  uint8_t* ring = lnw_smem + (size_t)2 * D * sizeof(float) + (size_t)warp * stages * stage_bytes;

  const int units = (nrows + LNW_ROWS - 1) / LNW_ROWS;
  const int u_first = cta_in_seg * LNW_WARPS + warp, u_step = ctas_in_seg * LNW_WARPS;
  const __nv_bfloat16* xseg = p.x + ((long long)b * p.x_batch_rows + p.x_row_offset + seg_row0) * D;
  auto fetch = [&](int u, int stage) {                               // one lane: the unit's rows are contiguous in x
    const int n = min(LNW_ROWS, nrows - u * LNW_ROWS);
    mbar_arrive_expect_tx(&full[warp][stage], (uint32_t)n * row_bytes);
    bulk_load(ring + (size_t)stage * stage_bytes, xseg + (long long)u * LNW_ROWS * D, (uint32_t)n * row_bytes, &full[warp][stage]);
  };
  if (lane == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&full[warp][i], 1);
    fence_barrier_init();
    for (int i = 0; i < stages; ++i)
      if (u_first + i * u_step < units) fetch(u_first + i * u_step, i);
  }

  // element k = 8c + 4h + e of a coefficient table lives at h * D/2 + 4c + e: lane c reads float4 c of each half
  for (int c = threadIdx.x; c < nchunk; c += LNW_WARPS * 32) {
    float g[8], be[8], A[8], C[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(p.gamma) + c), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(p.beta) + c), be);
    if (p.mod) {
      const float* sh = p.mod + (long long)b * p.mod_batch_stride + (text ? p.shift_text_off : p.shift_video_off) + c * 8;
      const float* sc = p.mod + (long long)b * p.mod_batch_stride + (text ? p.scale_text_off : p.scale_video_off) + c * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float one_plus = 1.f + __ldg(sc + e);
        A[e] = g[e] * one_plus;
        C[e] = fmaf(be[e], one_plus, __ldg(sh + e));
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) { A[e] = g[e]; C[e] = be[e]; }
    }
    *reinterpret_cast<float4*>(tA + c * 4) = make_float4(A[0], A[1], A[2], A[3]);
    *reinterpret_cast<float4*>(tA + halfD + c * 4) = make_float4(A[4], A[5], A[6], A[7]);
    *reinterpret_cast<float4*>(tC + c * 4) = make_float4(C[0], C[1], C[2], C[3]);
    *reinterpret_cast<float4*>(tC + halfD + c * 4) = make_float4(C[4], C[5], C[6], C[7]);
  }
  __syncthreads();

  const float invD = 1.0f / (float)D;
  int stage = 0;
  uint32_t phase = 0;
  for (int u = u_first; u < units; u += u_step) {
    const int r0 = u * LNW_ROWS;
    const int n = min(LNW_ROWS, nrows - r0);
    __nv_bfloat16* y = p.y + ((long long)b * p.rows_per_batch + seg_row0 + r0) * D;
    const uint8_t* src = ring + (size_t)stage * stage_bytes + (size_t)lane * 16;
    while (!mbar_try_wait(&full[warp][stage], phase)) {}            // a bulk copy always completes
    float mean[LNW_ROWS], rstd[LNW_ROWS];
#pragma unroll
    for (int r = 0; r < LNW_ROWS; ++r) {                              // pass 1: sums
      uint64_t a0 = lnw_pack2(0.f, 0.f), a1 = a0;
      if (r < n) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          if (lane + i * 32 < nchunk) {
            const uint4 w = *reinterpret_cast<const uint4*>(src + (size_t)r * row_bytes + i * 512);
            a0 = lnw_add2(a0, lnw_add2(lnw_pair(w.x), lnw_pair(w.y)));
            a1 = lnw_add2(a1, lnw_add2(lnw_pair(w.z), lnw_pair(w.w)));
          }
        }
      }
      float lo, hi;
      lnw_unpack2(lnw_add2(a0, a1), lo, hi);
      mean[r] = warp_sum(lo + hi) * invD;
    }
#pragma unroll
    for (int r = 0; r < LNW_ROWS; ++r) {                              // pass 2: centred squares
      const uint64_t nm2 = lnw_pack2(-mean[r], -mean[r]);
      uint64_t q0 = lnw_pack2(0.f, 0.f), q1 = q0;
      if (r < n) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          if (lane + i * 32 < nchunk) {
            const uint4 w = *reinterpret_cast<const uint4*>(src + (size_t)r * row_bytes + i * 512);
            const uint64_t d0 = lnw_add2(lnw_pair(w.x), nm2), d1 = lnw_add2(lnw_pair(w.y), nm2);
            const uint64_t d2 = lnw_add2(lnw_pair(w.z), nm2), d3 = lnw_add2(lnw_pair(w.w), nm2);
            q0 = lnw_fma2(d0, d0, q0); q1 = lnw_fma2(d1, d1, q1);
            q0 = lnw_fma2(d2, d2, q0); q1 = lnw_fma2(d3, d3, q1);
          }
        }
      }
      float lo, hi;
      lnw_unpack2(lnw_add2(q0, q1), lo, hi);
      rstd[r] = rsqrtf(warp_sum(lo + hi) * invD + p.eps);
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {                                    // pass 3: normalise, modulate, store
      const int c = lane + i * 32;
      if (c < nchunk) {
        const float4 A0 = *reinterpret_cast<const float4*>(tA + c * 4), A1 = *reinterpret_cast<const float4*>(tA + halfD + c * 4);
        const float4 C0 = *reinterpret_cast<const float4*>(tC + c * 4), C1 = *reinterpret_cast<const float4*>(tC + halfD + c * 4);
        const uint64_t Ax = lnw_pack2(A0.x, A0.y), Ay = lnw_pack2(A0.z, A0.w), Az = lnw_pack2(A1.x, A1.y), Aw = lnw_pack2(A1.z, A1.w);
        const uint64_t Cx = lnw_pack2(C0.x, C0.y), Cy = lnw_pack2(C0.z, C0.w), Cz = lnw_pack2(C1.x, C1.y), Cw = lnw_pack2(C1.z, C1.w);
#pragma unroll
        for (int r = 0; r < LNW_ROWS; ++r) {
          if (r < n) {
            const uint4 w = *reinterpret_cast<const uint4*>(src + (size_t)r * row_bytes + i * 512);
            const uint64_t rs2 = lnw_pack2(rstd[r], rstd[r]);
            const float nmr = -mean[r] * rstd[r];
            const uint64_t nmr2 = lnw_pack2(nmr, nmr);
            float o0, o1, o2, o3, o4, o5, o6, o7;
            lnw_unpack2(lnw_fma2(lnw_fma2(lnw_pair(w.x), rs2, nmr2), Ax, Cx), o0, o1);
            lnw_unpack2(lnw_fma2(lnw_fma2(lnw_pair(w.y), rs2, nmr2), Ay, Cy), o2, o3);
            lnw_unpack2(lnw_fma2(lnw_fma2(lnw_pair(w.z), rs2, nmr2), Az, Cz), o4, o5);
            lnw_unpack2(lnw_fma2(lnw_fma2(lnw_pair(w.w), rs2, nmr2), Aw, Cw), o6, o7);
            uint4 out;
            out.x = pack_bf16(o0, o1); out.y = pack_bf16(o2, o3); out.z = pack_bf16(o4, o5); out.w = pack_bf16(o6, o7);
            *reinterpret_cast<uint4*>(y + (long long)r * D + c * 8) = out;
          }
        }
      }
    }
    __syncwarp();                                                     // every lane has read its part of the stage
    if (lane == 0) {
      const int nxt = u + stages * u_step;
      if (nxt < units) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads above before the async write below
        fetch(nxt, stage);
      }
    }
    if (++stage == stages) { stage = 0; phase ^= 1; }
  }
}

// Final head normalisation: y = LN2(LN1(x)) * (1 + scale[b]) + shift[b] on the video rows only (T3D:613-624).
template <int CH>
__global__ void __launch_bounds__(256) ln_double_kernel(const LnModParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= p.rows) return;
  const int b = (int)(row / p.rows_per_batch);
  const int s = (int)(row - (long long)b * p.rows_per_batch);
  const __nv_bfloat16* x = p.x + ((long long)b * p.x_batch_rows + p.x_row_offset + s) * p.D;
  __nv_bfloat16* y = p.y + row * p.D;
  const int nchunk = p.D >> 3;
  const float invD = 1.0f / (float)p.D;

  float v[CH][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + i * 32;
    if (c < nchunk) {
      unpack8(ldg_nc_v4(x + c * 8), v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[i][e];
    }
  }
  float mean = warp_sum(sum) * invD;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + i * 32;
    if (c < nchunk) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mean; sq += d * d; }
    }
  }
  float rstd = rsqrtf(warp_sum(sq) * invD + p.eps);
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + i * 32;
    if (c < nchunk) {
      float g[8], be[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.gamma) + c), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.beta) + c), be);
#pragma unroll
      for (int e = 0; e < 8; ++e) { v[i][e] = (v[i][e] - mean) * rstd * g[e] + be[e]; sum += v[i][e]; }
    }
  }
  mean = warp_sum(sum) * invD;
  sq = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + i * 32;
    if (c < nchunk) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mean; sq += d * d; }
    }
  }
  rstd = rsqrtf(warp_sum(sq) * invD + p.eps);
  const float* shift = p.mod + (long long)b * p.mod_batch_stride + p.shift_video_off;
  const float* scale = p.mod + (long long)b * p.mod_batch_stride + p.scale_video_off;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + i * 32;
    if (c < nchunk) {
      float g[8], be[8], o[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.gamma2) + c), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.beta2) + c), be);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        o[e] = ((v[i][e] - mean) * rstd * g[e] + be[e]) * (1.f + __ldg(scale + c * 8 + e)) + __ldg(shift + c * 8 + e);
      *reinterpret_cast<uint4*>(y + c * 8) = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// out[b, n] = sum_k act(in[b, k]) * W[n, k] + bias[n]   (fp32 in/out, bf16 weights).  Weight-bandwidth bound: the adaLN
// tables of one forward stream 1.6 GB of weights through it (44 blocks x 2 x [18432, 512] bf16, i.e. 1 KiB rows).
//   * the activations (batch <= 8 rows of K floats) go through the optional SiLU ONCE per CTA into shared memory (the
//     first version evaluated 1024 expf per KiB of weights: MUFU-bound at 28 % of the HBM rate), stored so that the two
//     float4 a lane needs per 16-byte weight load are conflict-free: element k = 8g + 4c + e lives at c * K/2 + 4g + e;
//   * EIGHT lanes per output row, four rows per warp: one warp-wide 16-byte load covers 128 contiguous bytes of four rows,
//     the cross-lane reduction is three shuffles per (row, batch) — with a whole warp per 1 KiB row it was five, and the
//     reduce phase (no loads in flight) took as long as the loads;
//   * eight independent 16-byte loads per lane are issued before the first is used (4 KiB in flight per warp);
//   * persistent CTAs, grid-stride over the row groups.
// ------------------------------------------------------------------------------------------------------------------
constexpr int GEMV_MAX_B = 8;
constexpr int GEMV_WARPS = 8;
constexpr int GEMV_ROWS_PER_WARP = 4;
constexpr int GEMV_LOADS = 8;                                         // 16-byte weight loads in flight per lane
template <int BT>
__global__ void __launch_bounds__(GEMV_WARPS * 32, 4) gemv_kernel(const float* __restrict__ in, const __nv_bfloat16* __restrict__ W,
                                                                  const __nv_bfloat16* __restrict__ bias, float* __restrict__ out,
                                                                  int B, int N, int K, int act_silu) {
  extern __shared__ float gemv_x[];                                  // [B][K] activations after the optional SiLU, permuted
  const int halfK = K >> 1;
  for (int i = threadIdx.x; i < B * K; i += GEMV_WARPS * 32) {
    const int b = i / K, k = i - b * K;
    const float v = in[i];
    gemv_x[b * K + ((k >> 2) & 1) * halfK + (k >> 3) * 4 + (k & 3)] = act_silu ? silu(v) : v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane & 7, rg = lane >> 3;                          // lane `sub` of the eight that share row `rg` of the warp
  const int groups = K >> 3;                                         // 16-byte groups per row
  for (long long n0 = (long long)(blockIdx.x * GEMV_WARPS + warp) * GEMV_ROWS_PER_WARP; n0 < N;
       n0 += (long long)gridDim.x * GEMV_WARPS * GEMV_ROWS_PER_WARP) {
    const long long n = n0 + rg < N ? n0 + rg : N - 1;               // rows past the end re-read the last row (not stored)
    const __nv_bfloat16* wrow = W + n * K;
    float acc[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = 0.f;
    for (int g0 = 0; g0 < groups; g0 += 8 * GEMV_LOADS) {
      uint4 w[GEMV_LOADS];
#pragma unroll
      for (int j = 0; j < GEMV_LOADS; ++j) {
        const int g = g0 + j * 8 + sub;
        w[j] = g < groups ? ldg_nc_v4(wrow + g * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int j = 0; j < GEMV_LOADS; ++j) {
        const int g = g0 + j * 8 + sub;
        if (g < groups) {
          float wf[8];
          unpack8(w[j], wf);
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            if (b < B) {
              const float4 a0 = *reinterpret_cast<const float4*>(gemv_x + b * K + g * 4);
              const float4 a1 = *reinterpret_cast<const float4*>(gemv_x + b * K + halfK + g * 4);
              acc[b] = fmaf(a0.x, wf[0], acc[b]); acc[b] = fmaf(a0.y, wf[1], acc[b]);
              acc[b] = fmaf(a0.z, wf[2], acc[b]); acc[b] = fmaf(a0.w, wf[3], acc[b]);
              acc[b] = fmaf(a1.x, wf[4], acc[b]); acc[b] = fmaf(a1.y, wf[5], acc[b]);
              acc[b] = fmaf(a1.z, wf[6], acc[b]); acc[b] = fmaf(a1.w, wf[7], acc[b]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      float v = acc[b];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      if (b < B && sub == 0 && n0 + rg < N) out[(long long)b * N + n] = v + (bias ? __bfloat162float(bias[n]) : 0.f);
    }
  }
}

// sinusoidal timestep features, EMB:56-73 with flip_sin_to_cos: out[b] = [cos(t w_i) | sin(t w_i)]
__global__ void timestep_sinusoid_kernel(const long long* __restrict__ t_i64, const float* __restrict__ t_f32,
                                         float* __restrict__ out, int B, int dim, int flip_sin_to_cos, float freq_shift) {
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * half) return;
  const int b = idx / half, i = idx - b * half;
  const float t = t_i64 ? (float)t_i64[b] : t_f32[b];
  const float w = expf(-9.210340371976184f * (float)i / ((float)half - freq_shift));   // ln(10000)
  const float a = t * w;
  const float sv = sinf(a), cv = cosf(a);
  float* o = out + (long long)b * dim;
  if (flip_sin_to_cos) { o[i] = cv; o[half + i] = sv; }
  else { o[i] = sv; o[half + i] = cv; }
}

// ------------------------------------------------------------------------------------------------------------------
// patch gather for Conv2d(C -> D, k=2, s=2): A[(b, f, py, px), c*4 + dy*2 + dx] = src[b, f, c, 2py+dy, 2px+dx]
// (channels may come from two tensors concatenated along C: BR:359).  Columns >= 4C are zero padding.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patchify_kernel(const __nv_bfloat16* __restrict__ src0, int C0,
                                                       const __nv_bfloat16* __restrict__ src1, int C1, int BF, int H, int W,
                                                       __nv_bfloat16* __restrict__ out, int Kpad) {
  const int ph = H / 2, pw = W / 2;
  const long long total = (long long)BF * ph * pw * (Kpad / 2);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int kp = (int)(idx % (Kpad / 2));            // pair index: (c, dy) with dx = 0, 1
    long long r = idx / (Kpad / 2);
    const int px = (int)(r % pw); r /= pw;
    const int py = (int)(r % ph);
    const long long bf = r / ph;
    const int c = kp >> 1, dy = kp & 1;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
    if (c < C0 + C1) {
      const __nv_bfloat16* s = c < C0 ? src0 + ((bf * C0 + c) * H + (2 * py + dy)) * (long long)W + 2 * px
                                      : src1 + ((bf * C1 + (c - C0)) * H + (2 * py + dy)) * (long long)W + 2 * px;
      v = *reinterpret_cast<const __nv_bfloat162*>(s);
    }
    *reinterpret_cast<__nv_bfloat162*>(out + ((bf * ph + py) * pw + px) * (long long)Kpad + kp * 2) = v;
  }
}

// mask [B*F, 1, H, W] (any float dtype given as bf16) -> uint8 [B*F*ph*pw] = (avg_pool2d(mask, 2) > 0)   EMB:417-426
__global__ void mask_pool_kernel(const __nv_bfloat16* __restrict__ mask, int BF, int H, int W, uint8_t* __restrict__ out) {
  const int ph = H / 2, pw = W / 2;
  const long long total = (long long)BF * ph * pw;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int px = (int)(idx % pw);
  const int py = (int)((idx / pw) % ph);
  const long long bf = idx / ((long long)pw * ph);
  const __nv_bfloat16* m = mask + (bf * H + 2 * py) * (long long)W + 2 * px;
  const float s = __bfloat162float(m[0]) + __bfloat162float(m[1]) + __bfloat162float(m[W]) + __bfloat162float(m[W + 1]);
  out[idx] = (s * 0.25f) > 0.0f ? 1 : 0;
}

// unpatchify T3D:630-632: proj[(b, f, py, px), c*4 + dy*2 + dx] -> out[b, f, c, 2py+dy, 2px+dx]
__global__ void unpatchify_kernel(const __nv_bfloat16* __restrict__ proj, int BF, int C, int H, int W,
                                  __nv_bfloat16* __restrict__ out) {
  const int ph = H / 2, pw = W / 2;
  const long long total = (long long)BF * C * H * pw;       // one thread per horizontal pixel pair
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int px = (int)(idx % pw);
  long long r = idx / pw;
  const int y = (int)(r % H); r /= H;
  const int c = (int)(r % C);
  const long long bf = r / C;
  const int py = y >> 1, dy = y & 1;
  const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(proj + ((bf * ph + py) * pw + px) * (long long)(C * 4) + c * 4 + dy * 2);
  *reinterpret_cast<__nv_bfloat162*>(out + ((bf * C + c) * H + y) * (long long)W + 2 * px) = v;
}

// Ulysses receive side: [peer][slot][head][row][64] -> per slot [head][peer * rows + row][64]; one 16-byte vector per
// thread, both sides read / written in whole 128-byte head rows.
struct UnpackDst { __nv_bfloat16* p[5]; };
__global__ void __launch_bounds__(256) a2a_unpack_heads_kernel(const uint4* __restrict__ src, UnpackDst dst, int slots, int peers,
                                                               int heads, int rows) {
  const long long total = (long long)peers * slots * heads * rows * 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx & 7);
    long long r = idx >> 3;
    const int row = (int)(r % rows); r /= rows;
    const int h = (int)(r % heads); r /= heads;
    const int slot = (int)(r % slots);
    const int peer = (int)(r / slots);
    uint4* d = reinterpret_cast<uint4*>(dst.p[slot]) + (((long long)h * peers + peer) * rows + row) * 8 + v;
    *d = ldg_nc_v4(src + idx);
  }
}

// Cross-GPU barrier over peer memory: every rank runs one of these on its own GPU.  Thread i tells rank i "rank my_rank has
// reached epoch" (release store into rank i's flag array, after a system-scope fence that orders this GPU's earlier peer
// stores), then waits until rank i has said the same to us.  Bounded (timeout_ns, 0 = never): a rank that runs out of time
// counts the event in word kPeerErrWord of its own flag buffer and carries on — the host decides what to do; nothing traps.
struct PeerFlags { uint32_t* p[8]; };
constexpr int kPeerErrWord = 8;
constexpr int kPeerCountWord = 9;     // epoch == 0: the epoch is this rank's own barrier count, kept in its flag buffer
__global__ void peer_barrier_kernel(PeerFlags flags, int peers, int my_rank, uint32_t epoch, unsigned long long timeout_ns) {
  const int i = threadIdx.x;
  if (epoch == 0) {
    // Device-side epoch: nothing in the launch depends on how many barriers ran before, so the launch can be replayed from
    // a CUDA graph.  Every rank runs the same sequence of barriers, hence the counts agree; launches of one stream are
    // serialised, so the read-modify-write needs no atomics.
    volatile uint32_t* count = flags.p[my_rank] + kPeerCountWord;
    epoch = *count + 1;
    __syncwarp();
    if (i == 0) *count = epoch;
  }
  if (i >= peers) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(flags.p[i] + my_rank), "r"(epoch) : "memory");
  const uint32_t* mine = flags.p[my_rank] + i;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    if ((++spins & 0xfff) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (timeout_ns != 0 && now - t0 > timeout_ns) {
        atomicAdd(flags.p[my_rank] + kPeerErrWord, 1u);
        break;
      }
    }
  }
  __threadfence_system();
}

// Ulysses output exchange: chunk d of the local buffer -> slot my_rank of rank d's buffer; 16-byte vectors, every warp
// writes 512 contiguous bytes (whole lines on the NVLink side).  One launch instead of `peers` copy-engine transfers, whose
// fixed cost (~20 us each) dominates at these sizes.
struct PeerPtrs { uint4* p[8]; };
__global__ void __launch_bounds__(256) peer_scatter_kernel(const uint4* __restrict__ src, PeerPtrs dst, int peers, int my_rank,
                                                           long long vec_per_peer) {
  const long long total = vec_per_peer * peers;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx / vec_per_peer);
    const long long off = idx - (long long)d * vec_per_peer;
    dst.p[d][(long long)my_rank * vec_per_peer + off] = ldg_nc_v4(src + idx);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Step end: classifier-free-guidance combine (PIPE:995-997), CogVideoXDPMScheduler.step for v-prediction
// (DPM:386-436) and the replace_gt re-noise / blend (PIPE:1017-1034) in one pass over the latent.  The arithmetic mirrors
// the reference op by op, including its dtype promotions (a 0-dim coefficient times a bf16 tensor is a bf16 product of the
// bf16-rounded coefficient), so results are bit-identical; __f*_rn keeps the compiler from contracting into FMAs.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__global__ void __launch_bounds__(256) step_end_kernel(const StepEndParams p) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    const float u = __bfloat162float(p.noise_pred[i]), c = __bfloat162float(p.noise_pred[p.n + i]);
    const float mo = __fadd_rn(u, __fmul_rn(p.guidance, __fsub_rn(c, u)));
    const float x = __bfloat162float(p.sample[i]);
    const float pred = __fsub_rn(bf16r(__fmul_rn(x, p.c_sqrt_alpha_bf)), __fmul_rn(mo, p.c_sqrt_beta));
    float den = pred;
    if (p.second_order) den = __fsub_rn(__fmul_rn(pred, p.c_m2), __fmul_rn(p.old_pred[i], p.c_m3));
    const float prev = __fadd_rn(__fsub_rn(bf16r(__fmul_rn(x, p.c_m0_bf)), __fmul_rn(den, p.c_m1)),
                                 bf16r(__fmul_rn(__bfloat162float(p.noise[i]), p.c_mn_bf)));
    p.pred_out[i] = pred;
    if (p.prev_out) p.prev_out[i] = prev;
    float lat = bf16r(prev);
    if (p.gt) {
      float proper = __bfloat162float(p.gt[i]);
      if (p.renoise)
        proper = bf16r(__fadd_rn(bf16r(__fmul_rn(proper, p.sa_bf)), bf16r(__fmul_rn(__bfloat162float(p.noise0[i]), p.sb_bf))));
      const long long f = i / ((long long)p.chan * p.hw);
      const float m = __bfloat162float(p.mask[f * p.hw + i % p.hw]);
      const float om = bf16r(__fsub_rn(1.0f, m));
      lat = p.mask_background ? bf16r(__fadd_rn(bf16r(__fmul_rn(m, proper)), bf16r(__fmul_rn(om, lat))))
                              : bf16r(__fadd_rn(bf16r(__fmul_rn(om, proper)), bf16r(__fmul_rn(m, lat))));
    }
    p.latents_out[i] = __float2bfloat16_rn(lat);
  }
}

}  // namespace

int launch_step_end(const StepEndParams& p, cudaStream_t st) {
  long long blocks = (p.n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  step_end_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

int launch_peer_scatter(const void* src, void* const* peer_dst, int peers, int my_rank, long long bytes_per_peer, cudaStream_t st) {
  PeerPtrs d{};
  for (int i = 0; i < peers; ++i) d.p[i] = (uint4*)peer_dst[i];
  const long long vec = bytes_per_peer / 16;
  long long blocks = (vec * peers + 255) / 256;
  const int sms = sm_count();
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  peer_scatter_kernel<<<(unsigned)blocks, 256, 0, st>>>((const uint4*)src, d, peers, my_rank, vec);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

static std::atomic<long long>& peer_timeout_ms() {
  static std::atomic<long long> v{[]() -> long long {
    const char* e = getenv("VP_B200_PEER_TIMEOUT_MS");
    return e ? atoll(e) : 20000ll;
  }()};
  return v;
}

int set_peer_timeout_ms(long long ms) {
  if (ms < 0) return fail(VP_ERR_BAD_SHAPE, "peer_set_timeout_ms: negative timeout");
  peer_timeout_ms().store(ms);
  return VP_OK;
}

int launch_peer_barrier(uint32_t* const* peer_flags, int peers, int my_rank, uint32_t epoch, cudaStream_t st) {
  PeerFlags f{};
  for (int i = 0; i < peers; ++i) f.p[i] = peer_flags[i];
  const unsigned long long timeout_ns = (unsigned long long)peer_timeout_ms().load() * 1000000ull;
  peer_barrier_kernel<<<1, 32, 0, st>>>(f, peers, my_rank, epoch, timeout_ns);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

int launch_a2a_unpack_heads(const void* src, void* const* dst, int slots, int peers, int heads_local, int rows_per_peer,
                            cudaStream_t st) {
  UnpackDst d{};
  for (int i = 0; i < slots; ++i) d.p[i] = (__nv_bfloat16*)dst[i];
  const long long total = (long long)peers * slots * heads_local * rows_per_peer * 8;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  a2a_unpack_heads_kernel<<<(unsigned)blocks, 256, 0, st>>>((const uint4*)src, d, slots, peers, heads_local, rows_per_peer);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

int launch_ln_modulate(const LnModParams& p, cudaStream_t st) {
  VP_REQUIRE(p.rows > 0 && p.D > 0 && p.D % 8 == 0, VP_ERR_BAD_SHAPE, "ln_modulate: D must be a positive multiple of 8");
  VP_REQUIRE(p.D <= 4096, VP_ERR_UNSUPPORTED, "ln_modulate: D > 4096 not supported");
  VP_REQUIRE(p.rows_per_batch > 0 && p.rows % p.rows_per_batch == 0, VP_ERR_BAD_SHAPE, "ln_modulate: rows must be batch * rows_per_batch");
  const bool dbl = p.gamma2 != nullptr;
  if (dbl) {
    VP_REQUIRE(p.mod != nullptr, VP_ERR_BAD_SHAPE, "ln_double: modulation table required");
    const int ch = (p.D / 8 + 31) / 32;
    const unsigned grid = (unsigned)((p.rows + 7) / 8);
#define VP_LN_CASE(N)                                      \
  if (ch <= N) {                                           \
    ln_double_kernel<N><<<grid, 256, 0, st>>>(p);          \
    VP_CHECK_CUDA(cudaGetLastError());                     \
    return VP_OK;                                          \
  }
    VP_LN_CASE(1)
    VP_LN_CASE(2)
    VP_LN_CASE(4)
    VP_LN_CASE(8)
    VP_LN_CASE(12)
    VP_LN_CASE(16)
#undef VP_LN_CASE
    return fail(VP_ERR_UNSUPPORTED, "ln_double: unsupported width");
  }
  LnModParams q = p;
  if (q.mod == nullptr || q.text_len < 0) q.text_len = 0;
  if (q.text_len > q.rows_per_batch) q.text_len = q.rows_per_batch;
  VP_REQUIRE((reinterpret_cast<uintptr_t>(p.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0, VP_ERR_BAD_ALIGN,
             "ln_modulate: x and y must be 16-byte aligned");
  static const bool cols = []() { const char* e = getenv("VP_B200_LN"); return e && !strcmp(e, "cols"); }();
  const int batch_n = (int)(p.rows / p.rows_per_batch);
  if (!cols && 2 * batch_n <= LNW_MAX_SEGS) {
    const int sms_w = sm_count();
    if (sms_w <= 0) return fail(VP_ERR_CUDA, "no CUDA device");
    // CTAs per (batch, expert) segment in proportion to its rows, at least one, at most one per LNW_WARPS units
    LnSegs sg{};
    int seg_rows[LNW_MAX_SEGS];
    for (int bi = 0; bi < batch_n; ++bi) {
      if (q.text_len > 0) {
        sg.batch[sg.n] = bi; sg.text[sg.n] = 1; sg.row0[sg.n] = 0; sg.nrows[sg.n] = q.text_len; seg_rows[sg.n++] = q.text_len;
      }
      if (q.rows_per_batch > q.text_len) {
        sg.batch[sg.n] = bi; sg.text[sg.n] = 0; sg.row0[sg.n] = q.text_len; sg.nrows[sg.n] = q.rows_per_batch - q.text_len;
        seg_rows[sg.n++] = q.rows_per_batch - q.text_len;
      }
    }
    const long long budget = (long long)sms_w;                       // one CTA of LNW_WARPS warps per SM
    int cta = 0;
    for (int i = 0; i < sg.n; ++i) {
      const int units = (seg_rows[i] + LNW_ROWS - 1) / LNW_ROWS;
      long long want = (budget * seg_rows[i] + p.rows / 2) / p.rows;
      const long long cap = (units + LNW_WARPS - 1) / LNW_WARPS;
      if (want > cap) want = cap;
      if (want < 1) want = 1;
      sg.cta0[i] = cta;
      cta += (int)want;
    }
    sg.cta0[sg.n] = cta;
    const size_t tab_bytes = (size_t)2 * p.D * sizeof(float);
    const size_t stage_all = (size_t)LNW_WARPS * LNW_ROWS * p.D * 2;  // one stage of every warp
    int stages_w = (int)((LNW_SMEM_BUDGET - tab_bytes) / stage_all);
    if (stages_w > LNW_MAX_STAGES) stages_w = LNW_MAX_STAGES;
    const size_t smem_w = tab_bytes + (size_t)stages_w * stage_all;
    const int ch = (p.D / 8 + 31) / 32;
#define VP_LNW_CASE(N)                                                                                   \
  if (ch <= N && stages_w >= 1) {                                                                        \
    const int rc_w = configure_once(reinterpret_cast<const void*>(ln_rows_kernel<N>), LNW_SMEM_BUDGET);  \
    if (rc_w) return rc_w;                                                                               \
    ln_rows_kernel<N><<<(unsigned)cta, LNW_WARPS * 32, smem_w, st>>>(q, sg, stages_w);                   \
    VP_CHECK_CUDA(cudaGetLastError());                                                                   \
    return VP_OK;                                                                                        \
  }
    VP_LNW_CASE(1)
    VP_LNW_CASE(2)
    VP_LNW_CASE(4)
    VP_LNW_CASE(8)
    VP_LNW_CASE(12)
    VP_LNW_CASE(16)
#undef VP_LNW_CASE
  }
  const int threads = ((p.D / 8 + 31) / 32) * 32;
  const int units_text = (q.text_len + LN_RB - 1) / LN_RB;
  const int units_video = (q.rows_per_batch - q.text_len + LN_RB - 1) / LN_RB;
  const long long total = (p.rows / p.rows_per_batch) * (long long)(units_text + units_video);
  const int sms = sm_count();
  if (sms <= 0) return fail(VP_ERR_CUDA, "no CUDA device");
  VP_REQUIRE((reinterpret_cast<uintptr_t>(p.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0, VP_ERR_BAD_ALIGN,
             "ln_modulate: x and y must be 16-byte aligned");
  long long grid = (long long)sms * LN_CTAS;           // resident CTAs per SM (launch bounds), persistent
  if (grid > total) grid = total;
  const int stage_bytes = LN_RB * p.D * 2;
  int stages = LN_SMEM_BUDGET / stage_bytes;
  if (stages > LN_MAX_STAGES) stages = LN_MAX_STAGES;
  if (stages < 1) stages = 1;
  const size_t smem = (size_t)stages * stage_bytes;
  int rc_cfg;
  if ((rc_cfg = configure_once(reinterpret_cast<const void*>(ln_modulate_kernel<12>), LN_SMEM_BUDGET))) return rc_cfg;
  if ((rc_cfg = configure_once(reinterpret_cast<const void*>(ln_modulate_kernel<LN_MAX_WARPS>), LN_SMEM_BUDGET))) return rc_cfg;
  if (threads <= 12 * 32) ln_modulate_kernel<12><<<(unsigned)grid, threads, smem, st>>>(q, units_text, units_video, stages);
  else ln_modulate_kernel<LN_MAX_WARPS><<<(unsigned)grid, threads, smem, st>>>(q, units_text, units_video, stages);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

int launch_gemv(const float* in, const void* W, const void* bias, float* out, int B, int N, int K, int act_silu,
                cudaStream_t st) {
  VP_REQUIRE(B > 0 && B <= GEMV_MAX_B, VP_ERR_UNSUPPORTED, "gemv: batch must be in [1, 8]");
  VP_REQUIRE(N > 0 && K > 0 && K % 8 == 0, VP_ERR_BAD_SHAPE, "gemv: K must be a multiple of 8");
  const size_t smem = (size_t)B * K * sizeof(float);
  VP_REQUIRE(smem <= 200 * 1024, VP_ERR_UNSUPPORTED, "gemv: batch x K activations do not fit in shared memory");
  const void* fn = B == 1   ? reinterpret_cast<const void*>(gemv_kernel<1>)
                   : B == 2 ? reinterpret_cast<const void*>(gemv_kernel<2>)
                   : B <= 4 ? reinterpret_cast<const void*>(gemv_kernel<4>)
                            : reinterpret_cast<const void*>(gemv_kernel<8>);
  if (smem > 48 * 1024) {
    const int rc = configure_once(fn, 200 * 1024);
    if (rc) return rc;
  }
  const int rows_per_cta = GEMV_WARPS * GEMV_ROWS_PER_WARP;
  const int sms = sm_count();
  int grid = (N + rows_per_cta - 1) / rows_per_cta;
  if (sms > 0 && grid > sms * 4) grid = sms * 4;                       // persistent: 4 CTAs of 8 warps per SM
  const __nv_bfloat16* w = (const __nv_bfloat16*)W;
  const __nv_bfloat16* bs = (const __nv_bfloat16*)bias;
  const unsigned threads = GEMV_WARPS * 32;
  if (B == 1) gemv_kernel<1><<<grid, threads, smem, st>>>(in, w, bs, out, B, N, K, act_silu);
  else if (B == 2) gemv_kernel<2><<<grid, threads, smem, st>>>(in, w, bs, out, B, N, K, act_silu);
  else if (B <= 4) gemv_kernel<4><<<grid, threads, smem, st>>>(in, w, bs, out, B, N, K, act_silu);
  else gemv_kernel<8><<<grid, threads, smem, st>>>(in, w, bs, out, B, N, K, act_silu);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

int launch_timestep_sinusoid(const long long* t_i64, const float* t_f32, float* out, int B, int dim, int flip, float shift,
                             cudaStream_t st) {
  VP_REQUIRE(B > 0 && dim > 0 && dim % 2 == 0, VP_ERR_BAD_SHAPE, "timestep: dim must be even");
  VP_REQUIRE((t_i64 != nullptr) != (t_f32 != nullptr), VP_ERR_BAD_SHAPE, "timestep: exactly one timestep pointer");
  const int total = B * dim / 2;
  timestep_sinusoid_kernel<<<(total + 255) / 256, 256, 0, st>>>(t_i64, t_f32, out, B, dim, flip, shift);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

int launch_patchify(const void* src0, int C0, const void* src1, int C1, int BF, int H, int W, void* out, int Kpad,
                    cudaStream_t st) {
  VP_REQUIRE(BF > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, VP_ERR_BAD_SHAPE, "patchify: H and W must be even");
  VP_REQUIRE(Kpad % 8 == 0 && Kpad >= 4 * (C0 + C1), VP_ERR_BAD_SHAPE, "patchify: Kpad too small or not a multiple of 8");
  const long long total = (long long)BF * (H / 2) * (W / 2) * (Kpad / 2);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  patchify_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)src0, C0, (const __nv_bfloat16*)src1, C1, BF, H, W,
                                                   (__nv_bfloat16*)out, Kpad);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

int launch_mask_pool(const void* mask, int BF, int H, int W, uint8_t* out, cudaStream_t st) {
  VP_REQUIRE(BF > 0 && H % 2 == 0 && W % 2 == 0, VP_ERR_BAD_SHAPE, "mask_pool: H and W must be even");
  const long long total = (long long)BF * (H / 2) * (W / 2);
  mask_pool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)mask, BF, H, W, out);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

int launch_unpatchify(const void* proj, int BF, int C, int H, int W, void* out, cudaStream_t st) {
  VP_REQUIRE(BF > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, VP_ERR_BAD_SHAPE, "unpatchify: H and W must be even");
  const long long total = (long long)BF * C * H * (W / 2);
  unpatchify_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)proj, BF, C, H, W,
                                                                     (__nv_bfloat16*)out);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

}  // namespace vp
