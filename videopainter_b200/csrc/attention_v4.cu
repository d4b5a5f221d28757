// FOURTH DESIGN of the attention kernel (round 2), kept selectable with VP_B200_ATTN=v4 for A/B measurements on one box; the
// product kernel is attention.cu (the second design).  On one box, S = 17 776: this design 8.87 - 8.99 ms stand-alone and
// 726 ms per denoise step, the product kernel 8.63 - 8.66 ms and 718 ms (profiles/r2_attention_design_notes.md).
// Flash attention for the joint text+video sequence of CogVideoX (AP:2192-2197: non-causal, no mask, d_head = 64,
// scale 1/8), tcgen05 + TMEM + TMA.
//
// d_head = 64 attention on B200 is bound by the exponentials, not by the tensor pipe: per SM a 128 x 128 score tile needs
// 1024 clk of MUFU.EX2 against 600-750 clk of tcgen05.mma (profiles/r1_pipe_throughput_b200.txt, r1_tcgen05_mma_issue_b200.txt).
// What the measurements of this round showed (profiles/r2_attention_design_notes.md):
//   - one warp cannot keep its sub-partition's MUFU busy (in-order issue; 2 warps per sub-partition reach ~75 %), four can;
//   - any hand-off that makes the softmax warps wait for an MMA round trip (P ready -> P V -> next Q K^T -> S ready) costs
//     ~1000 clk per key tile and leaves the MUFU idle for that long.
// Hence:
//   * one CTA = FOUR independent query tiles ("streams") of 128 rows of one (batch, head); 16 softmax warps, warp w serves
//     stream w / 4 and TMEM lane quadrant w % 4 — every sub-partition hosts one warp of each stream;
//   * thread == query row over a 64-key tile: row maximum and row sum need no cross-thread exchange;
//   * the scores leave TMEM as soon as they are computed: each thread loads its S row into registers and releases the S
//     columns at once (s_free), so the MMA warp issues the stream's NEXT Q K^T while the exponentials of this tile are still
//     running — S(j+1) is complete long before the softmax asks for it;
//   * P (bf16) goes to shared memory (128B-swizzled K-major tile, the layout TMA gives Q), not over S in TMEM, which is what
//     decouples the next Q K^T from this tile's P V; O += P V is an SS-MMA (48 clk per 128x64x16 step instead of 45); the P
//     tiles are double-buffered, so the softmax does not wait for the P V of the previous tile either;
//   * four service warps, one per SM sub-partition, share the MMA issue evenly (two issue Q K^T, two issue P V, for two
//     streams each; two of them also feed the K and V rings by TMA);
//   * TMEM (512 columns) = 4 x (S 64 | O 64); K/V tiles (64 keys) stream through a TMA ring shared by the four streams;
//   * lazy rescaling: the running maximum is only refreshed when it grows by more than 2^8, so O is rarely touched;
//   * optionally a share of the exponentials is evaluated by a polynomial on the FMA pipe (VP_ATTN_POLY_PER8).
//
// Up to two K/V segments are attended in one softmax (ID-resample processor: AP:2283-2284); `out_scale` / `accumulate`
// implement the previous-window blend (AP:2176-2189).
#include "attention.cuh"
#include "host_util.cuh"

#include <stdlib.h>

namespace vp {


namespace {

constexpr int BQ = 128;          // query rows per stream
constexpr int NS = 4;            // streams (query tiles) per CTA
constexpr int BKV = 64;          // keys per tile
constexpr int DH = 64;           // head dim
#undef VP_ATTN_KV_STAGES
#define VP_ATTN_KV_STAGES 2                  // Q 64 KiB + P 128 KiB leave room for two 8 KiB K and V stages (each tile is needed
                                             // by the first stream a whole round after the last stream released its stage)
constexpr int ST = VP_ATTN_KV_STAGES;
constexpr int Q_BYTES = BQ * DH * 2;        // 16 KiB
constexpr int P_BYTES = BQ * BKV * 2;       // 16 KiB: one stream's P tile, same layout as a Q tile
constexpr int KV_BYTES = BKV * DH * 2;      // 8 KiB
constexpr int SMEM_Q = 0;
constexpr int SMEM_P = SMEM_Q + NS * Q_BYTES;
constexpr int SMEM_K = SMEM_P + 2 * NS * P_BYTES;   // P tiles are double-buffered by key-tile parity
constexpr int SMEM_V = SMEM_K + ST * KV_BYTES;
constexpr int SMEM_BAR = SMEM_V + ST * KV_BYTES;
constexpr int SMEM_BYTES = SMEM_BAR + 512 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
constexpr int NUM_SOFTMAX_WARPS = 4 * NS;
constexpr int NUM_THREADS = (NUM_SOFTMAX_WARPS + 4) * 32;   // 16 softmax + 4 service warps (TMA producers / MMA issuers)
#undef VP_ATTN_SOFTMAX_REGS
#undef VP_ATTN_OTHER_REGS
#define VP_ATTN_SOFTMAX_REGS 104
#define VP_ATTN_OTHER_REGS 64
constexpr int SOFTMAX_REGS = VP_ATTN_SOFTMAX_REGS, OTHER_REGS = VP_ATTN_OTHER_REGS;
static_assert(512 * SOFTMAX_REGS + 128 * OTHER_REGS <= 640 * 96, "register pool of the CTA exceeded");
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_STREAM = 128, COL_S = 0, COL_O = 64;
#undef VP_ATTN_POLY_PER8
#define VP_ATTN_POLY_PER8 1                  // of every 8 element pairs, this many take the polynomial exp2 (0..8)
#ifndef VP_ATTN_RESCALE_LOG2
#define VP_ATTN_RESCALE_LOG2 8.0f
#endif
constexpr float RESCALE_THRESHOLD = VP_ATTN_RESCALE_LOG2;   // log2 units
#ifndef VP_ATTN_SVC_SLEEP_NS
#define VP_ATTN_SVC_SLEEP_NS 0               // >0: service warps sleep this long between barrier probes (nanosleep is ~1 us coarse:
                                             // it delays the next Q K^T by a third of a tile and was measured no faster)
#endif
#if VP_ATTN_SVC_SLEEP_NS > 0
#define VP_SVC_WAIT(bar, parity) mbar_wait_relaxed(bar, parity, VP_ATTN_SVC_SLEEP_NS)
#else
#define VP_SVC_WAIT(bar, parity) mbar_wait(bar, parity)
#endif
#ifndef VP_ATTN_GROUPS
#define VP_ATTN_GROUPS 1                     // exponential groups per 64-key tile separated by scheduling fences (1, 2, 4, 8);
                                             // measured 8.38 / 9.07 / 9.45 / 9.54 ms for 1 / 2 / 4 / 8: the fences cost more than they hide
#endif
#ifndef VP_ATTN_BALANCED
#define VP_ATTN_BALANCED 1                   // 1: each service warp issues the MMAs of two streams; 0: one warp issues every Q K^T,
#endif                                       // another every P V, the TMA producers only produce
#undef VP_ATTN_TRACE
#define VP_ATTN_TRACE 0                      // the trace facility belongs to the product kernel (attention.cu)

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
// True, but neither the compiler nor ptxas can know (`bits` comes from a kernel parameter and is laundered through a
// volatile asm, so two uses are not recognised as the same condition): a branch on it splits a basic block, i.e. it is an
// instruction-scheduling fence.
__device__ __forceinline__ bool opaque_true(uint32_t bits) {
  uint32_t t;
  asm volatile("mov.u32 %0, %1;\n" : "=r"(t) : "r"(bits));
  return t != 0;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
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
  uint64_t s_full[NS];      // Q K^T of the stream's next tile has landed in its S columns              (tcgen05.commit)
  uint64_t s_free[NS];      // every thread of the stream holds its S row in registers                  (128 arrivals)
  uint64_t p_full[NS][2];   // the stream's P tile of an even / odd key tile is in shared memory        (128 arrivals)
  uint64_t o_done[NS][2];   // the stream's P V of an even / odd tile has completed: O is up to date, that P buffer is free
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
      mbar_init(&bars->k_empty[i], VP_ATTN_BALANCED ? 2 : 1);   // one tcgen05.commit from each Q K^T issuer
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], VP_ATTN_BALANCED ? 2 : 1);   // one from each P V issuer
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->s_free[s], BQ);
      mbar_init(&bars->p_full[s][0], BQ);
      mbar_init(&bars->p_full[s][1], BQ);
      mbar_init(&bars->o_done[s][0], 1);
      mbar_init(&bars->o_done[s][1], 1);
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
    const uint32_t a_s_free = smem_u32(&bars->s_free[s]);
    const uint32_t a_p_full = smem_u32(&bars->p_full[s][0]);   // [0] even tiles, [1] (+8 bytes) odd tiles
    const uint32_t a_o_done = smem_u32(&bars->o_done[s][0]);   // [0] even tiles, [1] (+8 bytes) odd tiles
    const uint32_t tS = tmem_base + s * COL_STREAM + COL_S + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t tO = tS + (COL_O - COL_S);
    // this thread's row of the stream's P tile: 128 bytes, 16-byte chunk c stored at c ^ (row & 7) (SWIZZLE_128B, K-major)
    const uint32_t p_row = smem_u32(smem + SMEM_P) + s * 2 * P_BYTES + row * 128;   // buffer of odd tiles: + P_BYTES
    const uint32_t p_xor = static_cast<uint32_t>(row & 7) << 4;
    const float c = p.scale_log2;
    const uint64_t c2v = pack2(c, c);
    const uint64_t one2 = pack2(p.one, p.one);                // 1.0 the compiler cannot see: keeps x * 1 + y an FFMA2
    const uint32_t one_bits = __float_as_uint(p.one);
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
      tc_fence_before();
      mbar_arrive_a(a_s_free);                                // the stream's next Q K^T may overwrite S now
      VP_TRACE(warp, j, 2);

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

      // The P buffer of this tile's parity is free once P V of tile j - 2 has completed (issued two tiles ago: no stall).
      // One barrier per parity and every phase waited for, in order: parity waits can only tell two consecutive phases apart.
      if (j > 1) mbar_wait_a(a_o_done + (j & 1) * 8, ((j >> 1) - 1) & 1);
      const bool need = (tile_max - m_used) * c > RESCALE_THRESHOLD;   // true at j == 0 (m_used = -inf)
      if (__any_sync(0xffffffffu, need)) {                             // tcgen05.ld / st are warp-collective
        float factor = 1.0f;
        if (need) {
          factor = fast_exp2((m_used - tile_max) * c);                 // exp2(-inf) = 0 at j == 0
          m_used = tile_max;
          row_sum *= factor;
        }
        if (j > 0) {
          mbar_wait_a(a_o_done + ((j - 1) & 1) * 8, ((j - 1) >> 1) & 1);   // P V of the previous tile has landed in O
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
          tmem_wait_st();
        }
      }
      VP_TRACE(warp, j, 3);
      const float neg_mc = -m_used * c;
      const uint64_t nmc2 = pack2(neg_mc, neg_mc);
      uint64_t acc0 = pack2(0.f, 0.f), acc1 = pack2(0.f, 0.f);
      // The 64 scores become 64 probabilities IN PLACE (sr), in groups of G = 64 / NG keys; `produce(g)` scales and exponentiates
      // group g, `consume(g)` adds it to the row sum, packs it to bf16 and stores its 16-byte chunks to the P tile.
      // Left alone (NG = 1) ptxas puts the consumer of an exponential pair one pair behind its MUFU.EX2 (`MUFU, MUFU,
      // FFMA2(prev), F2FP(prev), MUFU, MUFU, F2FP(this pair) ...`) and re-derives that schedule from any source order inside a
      // basic block.  NG > 1 separates the groups by branches it cannot remove (opaque_true), so that block k holds produce(k)
      // and consume(k - 1), whose operands were issued a whole group earlier — an experiment on whether the short
      // producer-consumer distance is what keeps the MUFU pipe at ~80 %.  It is not: every NG > 1 measured slower.
      auto produce = [&](int g, int G) {
#pragma unroll
        for (int i = g * G; i < (g + 1) * G; i += 2) {
          const uint64_t y2 = fma2(pack2(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])), c2v, nmc2);
          float e0, e1;
          if (((i >> 1) & 7) < VP_ATTN_POLY_PER8) {
            exp2_poly2(y2, one2, e0, e1);
          } else {
            float y0, y1;
            unpack2(y2, y0, y1);
#if defined(VP_ATTN_DEBUG_NOEXP)
            e0 = y0; e1 = y1;                                      // timing experiment only (wrong results)
#else
            e0 = fast_exp2(y0);
            e1 = fast_exp2(y1);
#endif
          }
          sr[i] = __float_as_uint(e0);
          sr[i + 1] = __float_as_uint(e1);
        }
      };
      const uint32_t p_dst = p_row + (j & 1) * P_BYTES;
      auto consume = [&](int g, int G) {
#pragma unroll
        for (int i = g * G; i < (g + 1) * G; i += 8) {             // 8 keys = one 16-byte chunk of the P row
          uint32_t pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float e0 = __uint_as_float(sr[i + 2 * q]), e1 = __uint_as_float(sr[i + 2 * q + 1]);
            pk[q] = pack_bf16(e0, e1);
            if (q & 1) acc1 = fma2(pack2(e0, e1), one2, acc1);
            else acc0 = fma2(pack2(e0, e1), one2, acc0);
          }
          sts_v4(p_dst + ((static_cast<uint32_t>(i >> 3) << 4) ^ p_xor), pk[0], pk[1], pk[2], pk[3]);
        }
      };
      constexpr int NG = VP_ATTN_GROUPS;                           // 1 = one block (the compiler's own interleaving)
      constexpr int G = 64 / NG;
      if (NG == 1) {
        produce(0, 64);
        consume(0, 64);
      } else {
        produce(0, G);
#pragma unroll
        for (int g = 1; g < NG; ++g) {
          if (opaque_true(one_bits)) {
            produce(g, G);
            consume(g - 1, G);
          }
        }
        if (opaque_true(one_bits)) consume(NG - 1, G);
      }
      {
        float a0, a1;
        unpack2(fma2(acc0, one2, acc1), a0, a1);
        row_sum += a0 + a1;
      }
      VP_TRACE(warp, j, 4);
      fence_proxy_async_smem();                               // generic-proxy stores -> visible to tcgen05.mma (async proxy)
      tc_fence_before();                                      // orders the rescale path's tcgen05.st before the P V MMA
      // (one barrier per tile parity, like o_done: a stream may finish P(j + 1) before the issuer has looked at P(j) of a
      //  slower stream, and a parity wait cannot tell phases j and j + 2 of one barrier apart)
      mbar_arrive_a(a_p_full + (j & 1) * 8);
      VP_TRACE(warp, j, 5);
    }

    // -------- epilogue: O / l -> bf16 -> out[b, q, h*64 ...] --------
    mbar_wait_a(a_o_done + ((n_tiles - 1) & 1) * 8, ((n_tiles - 1) >> 1) & 1);
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
    // Four service warps, one per SM sub-partition.  Issuing a tcgen05.mma costs its sub-partition issue time (measured: the
    // softmax warps sharing a sub-partition with a warp that issues 16 MMAs per round run 17 % slower than the others), so the
    // 32 MMAs of a round are spread evenly: 8 per sub-partition.
    //   warp 16: Q and K tiles by TMA, Q K^T of streams 0, 1        warp 17: Q K^T of streams 2, 3
    //   warp 18: P V of streams 0, 1                                 warp 19: V tiles by TMA, P V of streams 2, 3
    const int svc = warp - NUM_SOFTMAX_WARPS;
    // streams whose Q K^T (svc 0, 1) / P V (svc 2, 3) this warp issues: NI of them from s_lo on
#if VP_ATTN_BALANCED
    constexpr int NI = 2;
    const int s_lo = (svc & 1) * 2;
    const bool issuer = true;
#else
    constexpr int NI = NS;                                           // warp 17 issues every Q K^T, warp 18 every P V
    const int s_lo = 0;
    const bool issuer = svc == 1 || svc == 2;
#endif
    if (svc <= 1) {
      // ================================================ Q K^T issuers (+ K producer) ============================
      // S_s(j) = Q_s K_j^T (SS-MMA 128x64x64) as soon as the stream's threads have taken S_s(j-1) into registers.
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);
      const uint32_t sq = smem_u32(smem + SMEM_Q);
      const uint32_t sk = smem_u32(smem + SMEM_K);
      auto load_k = [&](int j) {                                       // K tile j into stage j % ST (warp 16 only)
        const int stage = j % ST;
        const bool seg1 = j >= n_t0;
        const int kv0 = (seg1 ? j - n_t0 : j) * BKV;
        VP_SVC_WAIT(&bars->k_empty[stage], ((j / ST) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->k_full[stage], KV_BYTES);
          tma_load_3d(smem + SMEM_K + stage * KV_BYTES, seg1 ? &tmap_k1 : &tmap_k0, &bars->k_full[stage], 0, kv0, bh, kEvictLast);
        }
        __syncwarp();
      };
      if (svc == 0) {
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->q_full, NS * Q_BYTES);
#pragma unroll
          for (int s = 0; s < NS; ++s)
            tma_load_3d(smem + SMEM_Q + s * Q_BYTES, &tmap_q, &bars->q_full, 0, q0 + s * BQ, bh, kEvictFirst);
        }
        __syncwarp();
        for (int j = 0; j < ST - 1 && j < n_tiles; ++j) load_k(j);
      }
      VP_SVC_WAIT(&bars->q_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        // K tile j + ST - 1 goes into the stage that tile j - 1 used: free once all four Q K^T of tile j - 1 have completed
        if (svc == 0 && j + ST - 1 < n_tiles) load_k(j + ST - 1);
        if (!issuer) continue;
        VP_SVC_WAIT(&bars->k_full[stage], phase);
#pragma unroll
        for (int si = 0; si < NI; ++si) {
          const int s = s_lo + si;
          if (j > 0) VP_SVC_WAIT(&bars->s_free[s], (j - 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = make_desc_sw128(sq + s * Q_BYTES, 1024, 0);
            const uint64_t bdesc = make_desc_sw128(sk + stage * KV_BYTES, 1024, 0);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
              mma_ss(tmem_base + s * COL_STREAM + COL_S, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
            tc_commit(&bars->s_full[s]);
            if (si == NI - 1) tc_commit(&bars->k_empty[stage]);        // one arrival per tile from each Q K^T issuer
          }
          __syncwarp();
          VP_TRACE(16 + svc, j, s & 3);
        }
        if (++stage == ST) { stage = 0; phase ^= 1; }
      }
    } else {
      // ================================================ P V issuers (+ V producer) ==============================
      // O_s += P_s(j) V_j (SS-MMA 128x64x64, P from shared memory, V as MN-major B) once the stream's P tile is complete.
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 0, 1);     // B = V, MN-major
      const uint32_t sp = smem_u32(smem + SMEM_P);
      const uint32_t sv = smem_u32(smem + SMEM_V);
      auto load_v = [&](int j) {                                       // V tile j into stage j % ST (warp 19 only)
        const int stage = j % ST;
        const bool seg1 = j >= n_t0;
        const int kv0 = (seg1 ? j - n_t0 : j) * BKV;
        VP_SVC_WAIT(&bars->v_empty[stage], ((j / ST) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->v_full[stage], KV_BYTES);
          tma_load_3d(smem + SMEM_V + stage * KV_BYTES, seg1 ? &tmap_v1 : &tmap_v0, &bars->v_full[stage], 0, kv0, bh, kEvictLast);
        }
        __syncwarp();
      };
      if (svc == 3)
        for (int j = 0; j < ST - 1 && j < n_tiles; ++j) load_v(j);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        if (svc == 3 && j + ST - 1 < n_tiles) load_v(j + ST - 1);
        if (!issuer) continue;
        VP_SVC_WAIT(&bars->v_full[stage], phase);
#pragma unroll
        for (int si = 0; si < NI; ++si) {
          const int s = s_lo + si;
          VP_SVC_WAIT(&bars->p_full[s][j & 1], (j >> 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            // V tile [64 keys][64 d] as MN-major B: 8-key groups are 1024 B apart, 16 keys per MMA = 2048 B
            const uint64_t pdesc = make_desc_sw128(sp + (s * 2 + (j & 1)) * P_BYTES, 1024, 0);
            const uint64_t vdesc = make_desc_sw128(sv + stage * KV_BYTES, 1024, 1024);
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k)
              mma_ss(tmem_base + s * COL_STREAM + COL_O, pdesc + 2 * k, vdesc + (uint64_t)(128 * k), idesc_pv, (j | k) != 0);
            tc_commit(&bars->o_done[s][j & 1]);
            if (si == NI - 1) tc_commit(&bars->v_empty[stage]);        // one arrival per tile from each P V issuer
          }
          __syncwarp();
          VP_TRACE(16 + svc, j, s & 3);
        }
        if (++stage == ST) { stage = 0; phase ^= 1; }
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

}  // namespace

int launch_attention_v4(const void* q, const void* k0, const void* v0, const void* k1, const void* v1, const AttnParams& p_in,
                        cudaStream_t st) {
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
