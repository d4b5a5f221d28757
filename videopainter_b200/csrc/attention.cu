// Flash attention for the joint text+video sequence of CogVideoX (AP:2192-2197: non-causal, no mask, d_head = 64,
// scale 1/8), tcgen05 + TMEM + TMA.  This is the second of four designs measured in round 2 and the one that won, stand-alone
// and inside the power-capped step (profiles/r2_attention_design_notes.md); the first (round 1) and the fourth are kept
// selectable for A/B runs on one box: VP_B200_ATTN=v1|v4 (attention_v1.cu, attention_v4.cu).
//
// d_head = 64 attention on B200 is bound by the exponentials, not by the tensor pipe: a 128 x 128 score tile needs 742 clk
// of tcgen05.mma here (614 with 128-key Q K^T tiles) but 1024 clk of MUFU.EX2 per SM (profiles/r1_pipe_throughput_b200.txt,
// r1_tcgen05_mma_issue_b200.txt).  The design goal is therefore that the MUFU pipe of every SM sub-partition is kept fed:
//
//   * one CTA = FOUR independent query tiles ("streams") of 128 rows of one (batch, head); 16 softmax warps, warp w serves
//     stream w / 4 and TMEM lane quadrant w % 4, so every sub-partition hosts one warp of each stream and the four streams
//     are at different phases of their tile loop (one loads S while the others exponentiate);
//   * thread == query row over a 64-key tile: the row maximum and the row sum need no cross-thread exchange, no shared
//     memory and no named barrier (the first design split a row over two warps);
//   * TMEM (512 columns) = 4 x (S 64 | O 64).  P (bf16) is written over the first 32 columns of its own S: every thread
//     has its whole S row in registers before it stores P, and the next Q K^T of the stream is issued behind the P V that
//     reads P (tcgen05.mma of one thread execute in order), so the softmax warps never wait for the P V MMA;
//   * two MMA-issuing warps, each serving two streams in turn (P V_s(j), then Q K_s(j+1)^T): tcgen05.mma issue blocks while
//     the pipe's short queue is full, so a single issuer left the pipe idle whenever it probed the next stream's barrier
//     (traced: ~80 clk per stream and tile; 10.7 ms with one issuer, 8.6 with two);
//   * K/V tiles (64 keys, 8 KiB each) stream through a TMA ring shared by the four streams: 512 query rows per K/V byte
//     fetched from L2 (the first design: 256);
//   * lazy rescaling: the running maximum is only refreshed when it grows by more than 2^8, so O is rarely touched;
//   * optionally a share of the exponentials is evaluated by a polynomial on the FMA pipe (VP_ATTN_POLY_PER8).
//
// Up to two K/V segments are attended in one softmax (ID-resample processor: AP:2283-2284); `out_scale` / `accumulate`
// implement the previous-window blend (AP:2176-2189).
#include "attention.cuh"
#include "host_util.cuh"

#include <mutex>
#include <stdlib.h>

namespace vp {

int launch_attention_v1(const void* q, const void* k0, const void* v0, const void* k1, const void* v1, const AttnParams& p,
                        cudaStream_t st);
int launch_attention_v4(const void* q, const void* k0, const void* v0, const void* k1, const void* v1, const AttnParams& p,
                        cudaStream_t st);

namespace {

constexpr int BQ = 128;          // query rows per stream
constexpr int NS = 4;            // streams (query tiles) per CTA
constexpr int BKV = 64;          // keys per tile
constexpr int DH = 64;           // head dim
#ifndef VP_ATTN_KV_STAGES
#define VP_ATTN_KV_STAGES 6
#endif
constexpr int ST = VP_ATTN_KV_STAGES;
constexpr int Q_BYTES = BQ * DH * 2;        // 16 KiB
constexpr int KV_BYTES = BKV * DH * 2;      // 8 KiB
constexpr int SMEM_Q = 0;
constexpr int SMEM_K = NS * Q_BYTES;
constexpr int SMEM_V = SMEM_K + ST * KV_BYTES;
constexpr int SMEM_BAR = SMEM_V + ST * KV_BYTES;
constexpr int SMEM_BYTES = SMEM_BAR + 512 + 1024;
constexpr int NUM_SOFTMAX_WARPS = 4 * NS;
constexpr int NUM_THREADS = (NUM_SOFTMAX_WARPS + 4) * 32;   // 16 softmax + TMA + MMA + 2 idle (setmaxnreg: whole warpgroups)
#ifndef VP_ATTN_SOFTMAX_REGS
#define VP_ATTN_SOFTMAX_REGS 104
#define VP_ATTN_OTHER_REGS 64
#endif
constexpr int SOFTMAX_REGS = VP_ATTN_SOFTMAX_REGS, OTHER_REGS = VP_ATTN_OTHER_REGS;
static_assert(512 * SOFTMAX_REGS + 128 * OTHER_REGS <= 640 * 96, "register pool of the CTA exceeded");
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_STREAM = 128, COL_S = 0, COL_O = 64;   // P aliases S columns [0, 32)
#ifndef VP_ATTN_POLY_PER8
#define VP_ATTN_POLY_PER8 0                  // of every 8 element pairs, this many take the polynomial exp2 (0..8)
#endif
#ifndef VP_ATTN_RESCALE_LOG2
#define VP_ATTN_RESCALE_LOG2 8.0f
#endif
constexpr float RESCALE_THRESHOLD = VP_ATTN_RESCALE_LOG2;   // log2 units
#ifndef VP_ATTN_CHUNK_PAIRS
#define VP_ATTN_CHUNK_PAIRS 8                // element pairs per P store (4, 8, 16: tcgen05.st x4 / x8 / x16)
#endif
#ifndef VP_ATTN_MMA_WARPS
#define VP_ATTN_MMA_WARPS 2                  // MMA-issuing warps (1 or 2): warp m serves streams m, m + NMMA, ...
#endif
constexpr int NMMA = VP_ATTN_MMA_WARPS;
static_assert(NMMA == 1 || NMMA == 2, "one or two MMA-issuing warps");
#ifndef VP_ATTN_OPTIMISTIC
#define VP_ATTN_OPTIMISTIC 0                 // 1: exponentials against the maximum in use, the tile's maximum verified after
                                             //    (bit-identical; measured 2 % SLOWER: 8.79 vs 8.59 - 8.75 ms, in-step 10.49 vs 10.33)
#endif
#ifndef VP_ATTN_TRACE
#define VP_ATTN_TRACE 0                      // 1: CTA (0, 0) records clock64() at the phase boundaries of its first tiles
#endif

#if VP_ATTN_TRACE
constexpr int TRACE_TILES = 48, TRACE_EV = 8, TRACE_WARPS = 18;
__device__ unsigned long long g_trace[TRACE_WARPS * TRACE_TILES * TRACE_EV];
#define VP_TRACE(warp_, tile_, ev_)                                                                          \
  do {                                                                                                       \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (tile_) < TRACE_TILES)               \
      g_trace[((warp_) * TRACE_TILES + (tile_)) * TRACE_EV + (ev_)] = clock64();                              \
  } while (0)
#else
#define VP_TRACE(warp_, tile_, ev_) do {} while (0)
#endif

// ---- packed fp32x2 arithmetic (one issue slot for two lanes of FMA-pipe work) and 3-input max -------------------------
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;\n" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;\n" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// exp2 of two values on the FMA / ALU pipes (no MUFU): Cody-Waite split with the 1.5*2^23 rounding trick and a
// degree-3 minimax polynomial on [-0.5, 0.5] (max relative error 7.5e-5, far below the bf16 rounding of P).
__device__ __forceinline__ void exp2_poly2(uint64_t y2, uint64_t one2, float& e0, float& e1) {
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);
  const uint64_t nmagic = pack2(-12582912.0f, -12582912.0f);
  const uint64_t c3 = pack2(0.055171459913253784f, 0.055171459913253784f);
  const uint64_t c2 = pack2(0.2426108568906784f, 0.2426108568906784f);
  const uint64_t c1 = pack2(0.6932609677314758f, 0.6932609677314758f);
  const uint64_t c0 = pack2(0.9999281167984009f, 0.9999281167984009f);
  float ya, yb;
  unpack2(y2, ya, yb);
  y2 = pack2(fmaxf(ya, -126.0f), fmaxf(yb, -126.0f));       // 2^y underflows below; keeps the exponent add in range
  const uint64_t t2 = fma2(y2, one2, magic);                 // low mantissa bits = round(y)
  const uint64_t fl2 = fma2(t2, one2, nmagic);               // round(y) as a float
  float fa, fb, la, lb;
  unpack2(fl2, la, lb);
  unpack2(y2, fa, fb);
  const uint64_t f2 = pack2(fa - la, fb - lb);               // y - round(y) in [-0.5, 0.5]
  uint64_t p2 = fma2(f2, c3, c2);
  p2 = fma2(p2, f2, c1);
  p2 = fma2(p2, f2, c0);
  float ta, tb, pa, pb;
  unpack2(t2, ta, tb);
  unpack2(p2, pa, pb);
  e0 = __uint_as_float(__float_as_uint(pa) + (__float_as_uint(ta) << 23));   // (magic bits << 23) == 0 mod 2^32
  e1 = __uint_as_float(__float_as_uint(pb) + (__float_as_uint(tb) << 23));
}

struct Bars {
  uint64_t q_full;
  uint64_t k_full[ST], k_empty[ST];
  uint64_t v_full[ST], v_empty[ST];
  uint64_t s_full[NS], p_full[NS], o_done[NS];
  uint32_t tmem_slot;
};
static_assert(sizeof(Bars) <= 512, "barrier block too large");

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k0,
                const __grid_constant__ CUtensorMap tmap_v0, const __grid_constant__ CUtensorMap tmap_k1,
                const __grid_constant__ CUtensorMap tmap_v1, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars* bars = reinterpret_cast<Bars*>(smem + SMEM_BAR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * (NS * BQ);
  const int n_t0 = (p.kv_len0 + BKV - 1) / BKV;
  const int n_t1 = (p.kv_len1 + BKV - 1) / BKV;
  const int n_tiles = n_t0 + n_t1;

  if (warp == NUM_SOFTMAX_WARPS && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k0);
    tma_prefetch_desc(&tmap_v0);
    if (n_t1 > 0) {
      tma_prefetch_desc(&tmap_k1);
      tma_prefetch_desc(&tmap_v1);
    }
  }
  if (warp == NUM_SOFTMAX_WARPS + 1 && lane == 0) {
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], NMMA);               // every MMA-issuing warp commits once per stage
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], NMMA);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->p_full[s], BQ);
      mbar_init(&bars->o_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == NUM_SOFTMAX_WARPS) tmem_alloc(&bars->tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp < NUM_SOFTMAX_WARPS) {
    // ================================================ softmax ===================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(SOFTMAX_REGS));
    const int s = warp >> 2;                                  // stream (query tile)
    const int quad = warp & 3;                                // TMEM lane quadrant (hardware: warp id % 4)
    const int row = quad * 32 + lane;
    const uint32_t a_s_full = smem_u32(&bars->s_full[s]);
    const uint32_t a_p_full = smem_u32(&bars->p_full[s]);
    const uint32_t a_o_done = smem_u32(&bars->o_done[s]);
    const uint32_t tS = tmem_base + s * COL_STREAM + COL_S + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t tO = tS + (COL_O - COL_S);
    const float c = p.scale_log2;
    const uint64_t c2v = pack2(c, c);
    const uint64_t one2 = pack2(p.one, p.one);                // 1.0 the compiler cannot see: keeps x * 1 + y an FFMA2
    float m_used = -INFINITY;     // maximum the exponents are currently referenced to (raw score units)
    float row_sum = 0.f;
    // tiles whose tail keys do not exist (last tile of each segment), and how many of their 64 columns are real
    const int rag0 = n_t0 - 1, rag1 = n_t1 > 0 ? n_tiles - 1 : -1;
    const int val0 = p.kv_len0 - (n_t0 - 1) * BKV;
    const int val1 = p.kv_len1 - (n_t1 - 1) * BKV;

#pragma unroll 2
    for (int j = 0; j < n_tiles; ++j) {
      VP_TRACE(warp, j, 0);
      mbar_wait_a(a_s_full, j & 1);
      tc_fence_after();
      VP_TRACE(warp, j, 1);
      uint32_t sr[64];
      tmem_ld_x32(tS + 0, sr + 0);
      tmem_ld_x32(tS + 32, sr + 32);
      tmem_wait_ld_dep64(sr);
      VP_TRACE(warp, j, 2);

#if VP_ATTN_OPTIMISTIC
      // Experiment kept for the record (off by default, see above): taking the row-maximum scan off the softmax warps'
      // critical path does not help — the scan is not what the S round trip waits for.
      // Optimistic reference maximum: the exponentials of tile j are issued at once against the maximum in use (m_used)
      // while the tile's own maximum is reduced in their shadow (the scan costs issue slots, not MUFU slots); only when the
      // maximum turns out to have grown by more than the threshold is the tile redone (S columns [32, 64) are still in TMEM,
      // [0, 32) are kept in registers) after O and the row sum have been rescaled.  Tile 0 scans first.  Results are
      // bit-identical to scanning first: the same reference maximum is used for every tile either way.
      auto mask_ragged = [&](int lo) {
        if (j == rag0 || j == rag1) {                         // ragged last tile of a segment
          const int v = (j == rag0) ? val0 : val1;
          if (v < BKV) {
#pragma unroll
            for (int i = lo; i < 64; ++i)
              if (i >= v) sr[i] = 0xff800000u;                // -inf
          }
        }
      };
      mask_ragged(0);
      if (j == 0) {
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
          mx0 = max3(mx0, __uint_as_float(sr[i + 0]), __uint_as_float(sr[i + 1]));
          mx1 = max3(mx1, __uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3]));
          mx2 = max3(mx2, __uint_as_float(sr[i + 4]), __uint_as_float(sr[i + 5]));
          mx3 = max3(mx3, __uint_as_float(sr[i + 6]), __uint_as_float(sr[i + 7]));
        }
        m_used = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      }
      VP_TRACE(warp, j, 3);
      uint64_t acc0, acc1;
#pragma unroll 1
      for (int pass = 0;; ++pass) {
        const float neg_mc = -m_used * c;
        const uint64_t nmc2 = pack2(neg_mc, neg_mc);
        acc0 = pack2(0.f, 0.f);
        acc1 = pack2(0.f, 0.f);
        float mx0 = -INFINITY, mx1 = -INFINITY;
        constexpr int CP = VP_ATTN_CHUNK_PAIRS;
#pragma unroll
        for (int ch = 0; ch < 32 / CP; ++ch) {
          uint32_t pk[CP];
          uint64_t y2[CP];
#pragma unroll
          for (int pr = 0; pr < CP; ++pr)
            y2[pr] = fma2(pack2(__uint_as_float(sr[(ch * CP + pr) * 2]), __uint_as_float(sr[(ch * CP + pr) * 2 + 1])), c2v, nmc2);
#pragma unroll
          for (int pr = 0; pr < CP; ++pr) {
            float e0, e1;
            if ((pr & 7) < VP_ATTN_POLY_PER8) {
              exp2_poly2(y2[pr], one2, e0, e1);
            } else {
              float y0, y1;
              unpack2(y2[pr], y0, y1);
              e0 = fast_exp2(y0);
              e1 = fast_exp2(y1);
            }
            pk[pr] = pack_bf16(e0, e1);
            if (pr & 1) acc1 = fma2(pack2(e0, e1), one2, acc1);
            else acc0 = fma2(pack2(e0, e1), one2, acc0);
            const int i = (ch * CP + pr) * 2;
            if (pr & 1) mx1 = max3(mx1, __uint_as_float(sr[i]), __uint_as_float(sr[i + 1]));
            else mx0 = max3(mx0, __uint_as_float(sr[i]), __uint_as_float(sr[i + 1]));
          }
          if (CP == 4) tmem_st_x4(tS + ch * CP, pk);
          else if (CP == 8) tmem_st_x8(tS + ch * CP, pk);
          else tmem_st_x16(tS + ch * CP, pk);
        }
        const float tile_max = fmaxf(mx0, mx1);
        const bool need = (tile_max - m_used) * c > RESCALE_THRESHOLD;
        if (!__any_sync(0xffffffffu, need)) break;                     // always taken in the second pass
        float factor = 1.0f;
        if (need) {
          factor = fast_exp2((m_used - tile_max) * c);
          m_used = tile_max;
          row_sum *= factor;
        }
        mbar_wait_a(a_o_done, (j - 1) & 1);                            // j >= 1 here: P_{j-1} V_{j-1} has landed in O
        tc_fence_after();
        tmem_wait_st();                                                // the first pass' P stores, before S is read again
#pragma unroll 1
        for (int c8 = 0; c8 < DH; c8 += 8) {
          uint32_t o[8];
          tmem_ld_x8(tO + c8, o);
          tmem_wait_ld();
          asm volatile("" : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]));
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
          tmem_st_x8(tO + c8, o);
        }
        tmem_ld_x32(tS + 32, sr + 32);                                 // columns [32, 64) of S were not overwritten by P
        tmem_wait_ld_dep32(sr + 32);
        mask_ragged(32);
      }
      {
        float a0, a1;
        unpack2(fma2(acc0, one2, acc1), a0, a1);
        row_sum += a0 + a1;
      }
#else
      if (j == rag0 || j == rag1) {                           // ragged last tile of a segment
        const int v = (j == rag0) ? val0 : val1;
        if (v < BKV) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= v) sr[i] = 0xff800000u;                  // -inf
        }
      }

      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        mx0 = max3(mx0, __uint_as_float(sr[i + 0]), __uint_as_float(sr[i + 1]));
        mx1 = max3(mx1, __uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3]));
        mx2 = max3(mx2, __uint_as_float(sr[i + 4]), __uint_as_float(sr[i + 5]));
        mx3 = max3(mx3, __uint_as_float(sr[i + 6]), __uint_as_float(sr[i + 7]));
      }
      const float tile_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

      const bool need = (tile_max - m_used) * c > RESCALE_THRESHOLD;   // true at j == 0 (m_used = -inf)
      if (__any_sync(0xffffffffu, need)) {                             // tcgen05.ld / st are warp-collective
        float factor = 1.0f;
        if (need) {
          factor = fast_exp2((m_used - tile_max) * c);                 // exp2(-inf) = 0 at j == 0
          m_used = tile_max;
          row_sum *= factor;
        }
        if (j > 0) {
          mbar_wait_a(a_o_done, (j - 1) & 1);                          // P_{j-1} V_{j-1} has landed in O
          tc_fence_after();
          // rare path: eight columns at a time, so that the 64 live score registers are not spilled around it
#pragma unroll 1
          for (int c8 = 0; c8 < DH; c8 += 8) {
            uint32_t o[8];
            tmem_ld_x8(tO + c8, o);
            tmem_wait_ld();
            asm volatile("" : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]));
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_x8(tO + c8, o);
          }
        }
      }
      VP_TRACE(warp, j, 3);
      const float neg_mc = -m_used * c;
      const uint64_t nmc2 = pack2(neg_mc, neg_mc);
      uint64_t acc0 = pack2(0.f, 0.f), acc1 = pack2(0.f, 0.f);
      // CP pairs at a time: scale, exponentiate, accumulate the row sum, pack to bf16 and store the packed words over the
      // (already consumed) S columns right away — few live registers, MUFU / FMA / ALU work of neighbouring chunks overlaps
      constexpr int CP = VP_ATTN_CHUNK_PAIRS;
#pragma unroll
      for (int ch = 0; ch < 32 / CP; ++ch) {
        uint32_t pk[CP];
        uint64_t y2[CP];
#pragma unroll
        for (int pr = 0; pr < CP; ++pr)
          y2[pr] = fma2(pack2(__uint_as_float(sr[(ch * CP + pr) * 2]), __uint_as_float(sr[(ch * CP + pr) * 2 + 1])), c2v, nmc2);
#pragma unroll
        for (int pr = 0; pr < CP; ++pr) {
          float e0, e1;
          if ((pr & 7) < VP_ATTN_POLY_PER8) {
            exp2_poly2(y2[pr], one2, e0, e1);
          } else {
            float y0, y1;
            unpack2(y2[pr], y0, y1);
#if defined(VP_ATTN_DEBUG_NOEXP)
            e0 = y0; e1 = y1;                                      // timing experiment only (wrong results)
#else
            e0 = fast_exp2(y0);
            e1 = fast_exp2(y1);
#endif
          }
          pk[pr] = pack_bf16(e0, e1);
          if (pr & 1) acc1 = fma2(pack2(e0, e1), one2, acc1);
          else acc0 = fma2(pack2(e0, e1), one2, acc0);
        }
        if (CP == 4) tmem_st_x4(tS + ch * CP, pk);
        else if (CP == 8) tmem_st_x8(tS + ch * CP, pk);
        else tmem_st_x16(tS + ch * CP, pk);
      }
      {
        float a0, a1;
        unpack2(fma2(acc0, one2, acc1), a0, a1);
        row_sum += a0 + a1;
      }
#endif
      VP_TRACE(warp, j, 4);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive_a(a_p_full);
      VP_TRACE(warp, j, 5);
    }

    // -------- epilogue: O / l -> bf16 -> out[b, q, h*64 ...] --------
    mbar_wait_a(a_o_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const int q_row = q0 + s * BQ + row;
    const bool row_ok = q_row < p.seq_q;
    const float inv = p.out_scale / row_sum;
    const int b = bh / p.heads, h = bh - b * p.heads;
    __nv_bfloat16* dst = p.out + ((long long)b * p.seq_q + q_row) * p.ldo + h * DH;
    if (p.peer_rows > 0 && row_ok) {                             // P2P store into the rank that owns this token row
      const int dest = q_row / p.peer_rows;
      dst = p.peer_out[dest] + ((long long)p.peer_src * p.peer_rows + (q_row - dest * p.peer_rows)) * p.ldo + h * DH;
    }
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      uint32_t o[32];
      tmem_ld_x32(tO + hlf * 32, o);
      tmem_wait_ld_dep32(o);
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[i * 8 + e]) * inv;
          __nv_bfloat16* d8 = dst + hlf * 32 + i * 8;
          if (p.accumulate) {                                    // read-modify-write in fp32 (previous-window blend)
            const uint4 old = *reinterpret_cast<const uint4*>(d8);
            f[0] += bf16_lo(old.x); f[1] += bf16_hi(old.x); f[2] += bf16_lo(old.y); f[3] += bf16_hi(old.y);
            f[4] += bf16_lo(old.z); f[5] += bf16_hi(old.z); f[6] += bf16_lo(old.w); f[7] += bf16_hi(old.w);
          }
          uint4 u;
          u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
          u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
          *reinterpret_cast<uint4*>(d8) = u;
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(OTHER_REGS));
    // Producer and MMA loops run warp-wide; only the asynchronous instructions are issued by one elected lane, so that
    // descriptor arithmetic stays on the uniform datapath (a divergent single thread costs ~100 clk per tcgen05.mma).
    if (warp == NUM_SOFTMAX_WARPS) {
      // ================================================ TMA producer ============================================
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->q_full, NS * Q_BYTES);
#pragma unroll
        for (int s = 0; s < NS; ++s)
          tma_load_3d(smem + SMEM_Q + s * Q_BYTES, &tmap_q, &bars->q_full, 0, q0 + s * BQ, bh, kEvictFirst);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const bool seg1 = j >= n_t0;
        const int kv0 = (seg1 ? j - n_t0 : j) * BKV;
        mbar_wait(&bars->k_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->k_full[stage], KV_BYTES);
          tma_load_3d(smem + SMEM_K + stage * KV_BYTES, seg1 ? &tmap_k1 : &tmap_k0, &bars->k_full[stage], 0, kv0, bh,
                      kEvictLast);
        }
        __syncwarp();
        mbar_wait(&bars->v_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->v_full[stage], KV_BYTES);
          tma_load_3d(smem + SMEM_V + stage * KV_BYTES, seg1 ? &tmap_v1 : &tmap_v0, &bars->v_full[stage], 0, kv0, bh,
                      kEvictLast);
        }
        __syncwarp();
        if (++stage == ST) { stage = 0; phase ^= 1; }
      }
    } else if (warp <= NUM_SOFTMAX_WARPS + NMMA) {
      // ================================================ MMA issuers =============================================
      // Per stream s and key tile j:  S_s = Q_s K_j^T (SS-MMA 128x64x64)  ->  softmax writes P_s over S_s  ->
      // O_s += P_s V_j (TS-MMA 128x64x64, P from TMEM, V as MN-major B)  ->  S_s = Q_s K_{j+1}^T behind it, in order.
      // tcgen05.mma issue blocks while the pipe's short queue is full, so a single issuer leaves the pipe idle whenever it
      // polls the next stream's P barrier (measured with VP_ATTN_TRACE: ~80 clk per stream and tile); with two issuers (streams
      // {0, 2} and {1, 3}) one of them always has MMAs queued.  Ordering is only needed inside a stream, i.e. inside a warp.
      const int mw = warp - (NUM_SOFTMAX_WARPS + 1);
      const int s_last = mw + NS - NMMA;
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 0, 1);     // B = V, MN-major
      const uint32_t sq = smem_u32(smem + SMEM_Q);
      const uint32_t sk = smem_u32(smem + SMEM_K);
      const uint32_t sv = smem_u32(smem + SMEM_V);

      auto issue_qk = [&](int s, int stage) {
        if (elect_one()) {
          const uint64_t adesc = make_desc_sw128(sq + s * Q_BYTES, 1024, 0);
          const uint64_t bdesc = make_desc_sw128(sk + stage * KV_BYTES, 1024, 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            mma_ss(tmem_base + s * COL_STREAM + COL_S, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
          tc_commit(&bars->s_full[s]);
          if (s == s_last) tc_commit(&bars->k_empty[stage]);
        }
        __syncwarp();
      };

      mbar_wait(&bars->q_full, 0);
      mbar_wait(&bars->k_full[0], 0);
      tc_fence_after();
      for (int s = mw; s < NS; s += NMMA) issue_qk(s, 0);

      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t par = j & 1;
        const int nstage = (stage + 1 == ST) ? 0 : stage + 1;
        const uint32_t nphase = (stage + 1 == ST) ? (phase ^ 1) : phase;
        const bool more = j + 1 < n_tiles;
        for (int s = mw; s < NS; s += NMMA) {
          VP_TRACE(16 + mw, j, (s / NMMA) * 2);
          mbar_wait(&bars->p_full[s], par);                    // P_s(j) is in TMEM
          if (s == mw) {
            mbar_wait(&bars->v_full[stage], phase);
            if (more) mbar_wait(&bars->k_full[nstage], nphase);
          }
          tc_fence_after();
          VP_TRACE(16 + mw, j, (s / NMMA) * 2 + 1);
          if (elect_one()) {
            // V tile [64 keys][64 d] as MN-major B: 8-key groups are 1024 B apart, 16 keys per MMA = 2048 B
            const uint64_t vdesc = make_desc_sw128(sv + stage * KV_BYTES, 1024, 1024);
            const uint32_t tP = tmem_base + s * COL_STREAM + COL_S;
            const uint32_t tOs = tmem_base + s * COL_STREAM + COL_O;
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k) mma_ts(tOs, tP + k * 8, vdesc + (uint64_t)(128 * k), idesc_pv, (j | k) != 0);
            tc_commit(&bars->o_done[s]);
            if (s == s_last) tc_commit(&bars->v_empty[stage]);
          }
          __syncwarp();
          if (more) issue_qk(s, nstage);
        }
        stage = nstage;
        phase = nphase;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NUM_SOFTMAX_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int make_map3(CUtensorMap* map, const void* ptr, long long bh, long long len, int box_rows) {
  uint64_t dims[3] = {(uint64_t)DH, (uint64_t)len, (uint64_t)bh};
  uint64_t str[2] = {(uint64_t)DH * 2, (uint64_t)len * DH * 2};
  uint32_t box[3] = {DH, (uint32_t)box_rows, 1};
  return make_tmap_bf16(map, ptr, 3, dims, str, box);
}

int which_design() {
  static const int v = []() {
    const char* e = getenv("VP_B200_ATTN");
    if (e && e[0] == 'v' && e[1] == '1') return 1;
    if (e && e[0] == 'v' && e[1] == '4') return 4;
    return 0;
  }();
  return v;
}

}  // namespace

#if VP_ATTN_TRACE
extern "C" int vp_debug_attn_trace(unsigned long long* host_out, int n) {
  const int total = TRACE_WARPS * TRACE_TILES * TRACE_EV;
  if (n > total) n = total;
  return cudaMemcpyFromSymbol(host_out, g_trace, sizeof(unsigned long long) * n) == cudaSuccess ? n : -1;
}
#endif

int launch_attention(const void* q, const void* k0, const void* v0, const void* k1, const void* v1, const AttnParams& p_in,
                     cudaStream_t st) {
  // development aid: other designs of this kernel, for A/B measurements on one box
  if (which_design() == 1) return launch_attention_v1(q, k0, v0, k1, v1, p_in, st);
  if (which_design() == 4) return launch_attention_v4(q, k0, v0, k1, v1, p_in, st);
  AttnParams p = p_in;
  p.one = 1.0f;
  VP_REQUIRE(p.batch > 0 && p.heads > 0 && p.seq_q > 0 && p.kv_len0 > 0 && p.kv_len1 >= 0, VP_ERR_BAD_SHAPE,
             "attention: bad shape");
  VP_REQUIRE(p.ldo % 8 == 0, VP_ERR_BAD_ALIGN, "attention: output leading dim must be a multiple of 8");
  VP_REQUIRE(p.kv_len1 == 0 || (k1 && v1), VP_ERR_BAD_SHAPE, "attention: second K/V segment missing");
  VP_REQUIRE(p.peer_out[0] == nullptr || (p.batch == 1 && p.peer_rows > 0 && !p.accumulate), VP_ERR_UNSUPPORTED,
             "attention: peer output needs batch 1 and no accumulation");
  if (p.peer_out[0] == nullptr) p.peer_rows = 0;
  int rc = configure_once(reinterpret_cast<const void*>(attn_fwd_kernel), SMEM_BYTES);
  if (rc) return rc;
  const long long bh = (long long)p.batch * p.heads;
  CUtensorMap mq, mk0, mv0, mk1, mv1;
  if ((rc = make_map3(&mq, q, bh, p.seq_q, BQ))) return rc;
  if ((rc = make_map3(&mk0, k0, bh, p.kv_len0, BKV))) return rc;
  if ((rc = make_map3(&mv0, v0, bh, p.kv_len0, BKV))) return rc;
  if (p.kv_len1 > 0) {
    if ((rc = make_map3(&mk1, k1, bh, p.kv_len1, BKV))) return rc;
    if ((rc = make_map3(&mv1, v1, bh, p.kv_len1, BKV))) return rc;
  } else {
    mk1 = mk0;
    mv1 = mv0;
  }
  dim3 grid((p.seq_q + NS * BQ - 1) / (NS * BQ), (unsigned)bh);
  attn_fwd_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mq, mk0, mv0, mk1, mv1, p);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

}  // namespace vp
