// FIRST DESIGN of the attention kernel (round 1), kept selectable with VP_B200_ATTN=v1 for A/B measurements on one box; the
// product kernel is attention.cu.
// Flash attention for the joint text+video sequence of CogVideoX (AP:2192-2197: non-causal, no mask, d_head = 64,
// scale 1/8), tcgen05 + TMEM + TMA.  One CTA owns 256 query rows of one (batch, head); K/V stream through a
// 3-stage TMA ring in 128-key tiles; up to two K/V segments are attended in one softmax (the ID-resample
// processor concatenates a second, masked K/V copy: AP:2283-2284).
//
//   warps 0-15  softmax: warp w serves query tile (w>>2)&1, TMEM lane quadrant w&3 and key columns [64*(w>>3), +64) of
//               the 128-key tile — two warps share each 32-row slab so that four softmax warps per SM sub-partition
//               keep the MUFU, FMA and ALU pipes busy at the same time (one warp alone issues in order and cannot).
//               thread == query row; S is read from TMEM, P is written back to TMEM as bf16; the pair exchanges its
//               partial row maxima through shared memory.
//   warp 16     TMEM allocator (512 columns: per tile S 128 | P 64 | O 64) and TMA producer (Q once, then K_j / V_j)
//   warp 17/18  MMA issuers, one per query tile (S_t = Q_t K_jᵀ : SS-MMA 128x128x64;  O_t += P_t V_j : TS-MMA
//               128x64x128 with P from TMEM and V as MN-major B)
// setmaxnreg moves registers from the three service warps (40) to the softmax warps (112).
// The running maximum is only refreshed when it grows by more than 2^8 (lazy rescale), so the O accumulator in
// TMEM is rarely touched by the softmax warps.  A tunable share of the exponentials is evaluated with a polynomial
// on the FMA pipe instead of MUFU.EX2 (d_head = 64 attention is exponent-bound on B200).
#include "attention.cuh"
#include "host_util.cuh"

namespace vp {

namespace {

constexpr int BQ = 128;          // query rows per softmax warpgroup
constexpr int BKV = 128;         // keys per tile
constexpr int DH = 64;           // head dim
#ifndef VP_ATTN_KV_STAGES
#define VP_ATTN_KV_STAGES 3
#endif
constexpr int KV_STAGES = VP_ATTN_KV_STAGES;
constexpr int TILE_BYTES = BKV * DH * 2;   // 16 KiB (Q tile has the same size)
constexpr int SMEM_Q = 0;
constexpr int SMEM_K = 2 * TILE_BYTES;
constexpr int SMEM_V = SMEM_K + KV_STAGES * TILE_BYTES;
constexpr int SMEM_BAR = SMEM_V + KV_STAGES * TILE_BYTES;
constexpr int SMEM_XCH = SMEM_BAR + 256;                   // float [2 buffers][2 tiles][2 halves][128 rows]
constexpr int SMEM_BYTES = SMEM_XCH + 2 * 2 * 2 * 128 * 4 + 1024;
constexpr int NUM_SOFTMAX_WARPS = 16;
constexpr int NUM_THREADS = (NUM_SOFTMAX_WARPS + 4) * 32;   // warp 19 idles: setmaxnreg needs whole warpgroups
#ifndef VP_ATTN_SOFTMAX_REGS
#define VP_ATTN_SOFTMAX_REGS 104
#define VP_ATTN_OTHER_REGS 56
#endif
// setmaxnreg only redistributes the CTA's launch allocation (640 threads * 96): 512 * softmax + 128 * other <= 61440
constexpr int SOFTMAX_REGS = VP_ATTN_SOFTMAX_REGS, OTHER_REGS = VP_ATTN_OTHER_REGS;
static_assert(512 * SOFTMAX_REGS + 128 * OTHER_REGS <= 640 * 96, "register pool of the CTA exceeded");
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192, COL_TILE = 256;
#ifndef VP_ATTN_POLY_PER16
#define VP_ATTN_POLY_PER16 0                 // of every 8 element pairs, this many take the polynomial exp2 (0..8)
#endif
// (Tried and removed: software pipelining over key tiles — refilling the registers of every finished 16-column chunk with
//  S(j+1) and scanning its maximum inside the exponentials of tile j.  With ONE S buffer per query tile in TMEM, Q K_{j+2}ᵀ can
//  then only start when tile j is three quarters done and lands on the critical path: 13.1 ms instead of 8.8 ms.  It needs
//  a second S buffer, i.e. 64-key tiles or a smaller O — a TMEM-budget redesign.)
#ifndef VP_ATTN_LEAN_WAIT
#define VP_ATTN_LEAN_WAIT 0                  // 1: softmax warps spin on try_wait without the watchdog (smaller loop body)
#endif
#ifndef VP_ATTN_UNROLL2
#define VP_ATTN_UNROLL2 1                    // 1: key-tile loop unrolled by two (barrier parities become constants)
#endif
#ifndef VP_ATTN_FULLMAX
#define VP_ATTN_FULLMAX 0                    // 1: every softmax warp also reads its partner's 64 columns for the row maximum
#endif                                       //    (twice the TMEM read traffic, no shared-memory exchange / pair barrier)
#ifndef VP_ATTN_RESCALE_LOG2
#define VP_ATTN_RESCALE_LOG2 8.0f
#endif
constexpr float RESCALE_THRESHOLD = VP_ATTN_RESCALE_LOG2;   // log2 units
#ifndef VP_ATTN_LATE_ODONE
#define VP_ATTN_LATE_ODONE 0                 // 1: wait for P_{j-1} V_{j-1} only before the first P store of tile j
#endif
#ifndef VP_ATTN_SKEW_CLK
#define VP_ATTN_SKEW_CLK 0                   // query tile 1 starts its softmax this many clocks late (de-phases the two tiles)
#endif
#ifndef VP_ATTN_CHUNK_PAIRS
#define VP_ATTN_CHUNK_PAIRS 8                // element pairs per P hand-off chunk (4, 8 or 16: tcgen05.st x4 / x8 / x16)
#endif
// (whole chunks on a four-pairs-at-a-time staged polynomial were measured too: slower than none at every share)
// (ex2.approx.ftz.bf16x2 was measured too: it lowers to two MUFU.EX2.BF16 operations, 16 clk per pair — no gain.)


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
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;\n" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// named barriers 1..8: one per pair of softmax warps that share a 32-row slab
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(n) : "memory"); }

// exp2 of two values on the FMA / ALU pipes (no MUFU): Cody-Waite split with the 1.5*2^23 rounding trick and a
// degree-3 minimax polynomial on [-0.5, 0.5] (max relative error 7.5e-5, far below the bf16 rounding of P).  All the
// packed additions are written as FFMA2 (x * 1 + y): measured 2.3 clk per warp instruction against 3.4 for FADD2.
__device__ __forceinline__ void exp2_poly2(uint64_t y2, float& e0, float& e1) {
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);
  const uint64_t nmagic = pack2(-12582912.0f, -12582912.0f);
  const uint64_t one = pack2(1.0f, 1.0f), mone = pack2(-1.0f, -1.0f);
  const uint64_t c3 = pack2(0.055171459913253784f, 0.055171459913253784f);
  const uint64_t c2 = pack2(0.2426108568906784f, 0.2426108568906784f);
  const uint64_t c1 = pack2(0.6932609677314758f, 0.6932609677314758f);
  const uint64_t c0 = pack2(0.9999281167984009f, 0.9999281167984009f);
  float ya, yb;
  unpack2(y2, ya, yb);
  y2 = pack2(fmaxf(ya, -126.0f), fmaxf(yb, -126.0f));       // 2^y underflows below; keeps the exponent add in range
  const uint64_t t2 = fma2(y2, one, magic);                  // low mantissa bits = round(y)
  const uint64_t fl2 = fma2(t2, one, nmagic);                // round(y) as a float
  const uint64_t f2 = fma2(fl2, mone, y2);                   // y - round(y) in [-0.5, 0.5]
  uint64_t p2 = fma2(f2, c3, c2);
  p2 = fma2(p2, f2, c1);
  p2 = fma2(p2, f2, c0);
  float ta, tb, pa, pb;
  unpack2(t2, ta, tb);
  unpack2(p2, pa, pb);
  e0 = __uint_as_float(__float_as_uint(pa) + (__float_as_uint(ta) << 23));   // (magic bits << 23) == 0 mod 2^32
  e1 = __uint_as_float(__float_as_uint(pb) + (__float_as_uint(tb) << 23));
}

__device__ __forceinline__ void softmax_wait(uint32_t bar, uint32_t parity) {
#if VP_ATTN_LEAN_WAIT
  while (!mbar_try_wait_a(bar, parity)) {}
#else
  mbar_wait_a(bar, parity);
#endif
}

struct Bars {
  uint64_t q_full;
  uint64_t k_full[KV_STAGES], k_empty[KV_STAGES];
  uint64_t v_full[KV_STAGES], v_empty[KV_STAGES];
  uint64_t s_full[2], s_free[2], p_full[2], o_done[2];
  uint32_t tmem_slot;
};

// PEER = true: the output rows are stored into the owning ranks' buffers (Ulysses over NVLink peer memory).  Two
// instantiations on purpose: the main loop's code generation is sensitive to everything that shares its registers.
template <bool PEER>
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
  const int q0 = blockIdx.x * (2 * BQ);
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
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], 2);            // both MMA warps commit
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bars->s_full[t], 1);
      mbar_init(&bars->s_free[t], 256);
      mbar_init(&bars->p_full[t], 256);
      mbar_init(&bars->o_done[t], 1);
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
    const int t = (warp >> 2) & 1;                            // query tile
    const int quad = warp & 3;                                // TMEM lane quadrant (hardware: warp id % 4)
    const int half = warp >> 3;                               // key columns [64*half, 64*half + 64) of each tile
    const int row = quad * 32 + lane;
    const int pair_bar = 1 + t * 4 + quad;                    // named barrier shared with the partner warp (w ^ 8)
    // shared-window addresses, computed once (the loop below only adds constants)
    // shared-window addresses (the loop below only adds constants).  Forcing them to stay in registers (volatile moves)
    // was measured 5 % slower than letting the compiler rematerialise them.
    const uint32_t a_s_full = smem_u32(&bars->s_full[t]), a_s_free = smem_u32(&bars->s_free[t]);
    const uint32_t a_p_full = smem_u32(&bars->p_full[t]), a_o_done = smem_u32(&bars->o_done[t]);
    const uint32_t a_xw = smem_u32(smem + SMEM_XCH) + ((t * 2 + half) * 128 + row) * 4;          // [buf][tile][half][row]
    const uint32_t a_xr = smem_u32(smem + SMEM_XCH) + ((t * 2 + (half ^ 1)) * 128 + row) * 4;
    constexpr uint32_t XBUF = 2 * 2 * 128 * 4;                // bytes between the two exchange buffers
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + t * COL_TILE + COL_S + half * 64 + lane_base;
    const uint32_t tP = tmem_base + t * COL_TILE + COL_P + half * 32 + lane_base;
    const uint32_t tO = tmem_base + t * COL_TILE + COL_O + half * 32 + lane_base;
    const float c = p.scale_log2;
    const uint64_t c2v = pack2(c, c), one2 = pack2(1.0f, 1.0f);
    float m_used = -INFINITY;     // maximum the exponents are currently referenced to (raw score units)
    float row_sum = 0.f;          // partial: this warp's 64 columns only
    // tiles whose tail keys do not exist (last tile of each segment), and how many of this warp's 64 columns are real
    const int rag0 = n_t0 - 1, rag1 = n_t1 > 0 ? n_tiles - 1 : -1;
    const int val0 = p.kv_len0 - (n_t0 - 1) * BKV - half * 64;
    const int val1 = p.kv_len1 - (n_t1 - 1) * BKV - half * 64;

    if (VP_ATTN_SKEW_CLK > 0 && t == 1) {
      const long long t0 = clock64();
      while (clock64() - t0 < VP_ATTN_SKEW_CLK) {}
    }
#if VP_ATTN_UNROLL2
#pragma unroll 2
#endif
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t par = j & 1;
      softmax_wait(a_s_full, par);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld_x32(tS + 0, sr + 0);
      tmem_ld_x32(tS + 32, sr + 32);
#if VP_ATTN_FULLMAX
      float omax;
      {
        const uint32_t tSo = tS + (half ? -64 : 64);             // the partner's columns of the same rows
        int vo = BKV;                                            // how many of them are real keys
        if (j == rag0) vo = val0 + half * 64 - (half ? 0 : 64);
        else if (j == rag1) vo = val1 + half * 64 - (half ? 0 : 64);
        uint32_t ot[32];
        float m0 = -INFINITY, m1 = -INFINITY;
        tmem_ld_x32(tSo, ot);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          m0 = max3(m0, i + 0 < vo ? __uint_as_float(ot[i + 0]) : -INFINITY, i + 1 < vo ? __uint_as_float(ot[i + 1]) : -INFINITY);
          m1 = max3(m1, i + 2 < vo ? __uint_as_float(ot[i + 2]) : -INFINITY, i + 3 < vo ? __uint_as_float(ot[i + 3]) : -INFINITY);
        }
        tmem_ld_x32(tSo + 32, ot);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          m0 = max3(m0, i + 32 < vo ? __uint_as_float(ot[i + 0]) : -INFINITY, i + 33 < vo ? __uint_as_float(ot[i + 1]) : -INFINITY);
          m1 = max3(m1, i + 34 < vo ? __uint_as_float(ot[i + 2]) : -INFINITY, i + 35 < vo ? __uint_as_float(ot[i + 3]) : -INFINITY);
        }
        omax = fmaxf(m0, m1);
      }
#else
      tmem_wait_ld();
#endif
      tc_fence_before();
      mbar_arrive_a(a_s_free);                                // S_t may be overwritten by the next QKᵀ

      if (j == rag0 || j == rag1) {                           // ragged last tile of a segment
        const int v = (j == rag0) ? val0 : val1;
        if (v < 64) {
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
      float tile_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
#if defined(VP_ATTN_DEBUG_NOMAX)
      tile_max = 0.f;                                            // timing experiment only (wrong results)
#elif VP_ATTN_FULLMAX
      tile_max = fmaxf(tile_max, omax);
#else
      // combine with the partner's half of the row (double-buffered by tile parity, one barrier per tile)
      sts_f32(a_xw + par * XBUF, tile_max);
      named_bar_sync(pair_bar, 64);
      tile_max = fmaxf(tile_max, lds_f32(a_xr + par * XBUF));
#endif

      bool waited_o = false;
      const bool need = (tile_max - m_used) * c > RESCALE_THRESHOLD;   // true at j == 0 (m_used = -inf)
      if (__any_sync(0xffffffffu, need)) {                             // same decision in both warps of the pair
        const float m_new = fmaxf(m_used, tile_max);
        const float factor = fast_exp2((m_used - m_new) * c);          // exp2(-inf) = 0 at j == 0
        m_used = m_new;
        row_sum *= factor;
        if (j > 0) {
          softmax_wait(a_o_done, (j - 1) & 1);                          // P_{j-1} V_{j-1} finished
          tc_fence_after();
          waited_o = true;
          uint32_t o[32];                                              // this warp rescales its 32 columns of O
          tmem_ld_x32(tO, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
          tmem_st_x32(tO, o);
        }
      }
#if !VP_ATTN_LATE_ODONE
      if (j > 0 && !waited_o) {                                        // P region still read by P_{j-1} V_{j-1}
        softmax_wait(a_o_done, (j - 1) & 1);
        tc_fence_after();
      }
#endif
      const float neg_mc = -m_used * c;
      const uint64_t nmc2 = pack2(neg_mc, neg_mc);
      uint64_t acc0 = pack2(0.f, 0.f), acc1 = pack2(0.f, 0.f);
      // 16 columns at a time: scale, exponentiate, accumulate the row sum, pack to bf16 and hand the 8 packed words to
      // TMEM right away (few live registers, and the MUFU / FMA / ALU work of neighbouring chunks overlaps)
      constexpr int CP = VP_ATTN_CHUNK_PAIRS;                       // pairs per chunk
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
          if ((pr & 7) < VP_ATTN_POLY_PER16) {
            exp2_poly2(y2[pr], e0, e1);
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
          if (pr & 1) acc1 = fma2(pack2(e0, e1), one2, acc1);      // FFMA2 issues faster than FADD2 on sm_100
          else acc0 = fma2(pack2(e0, e1), one2, acc0);
        }
        if (VP_ATTN_LATE_ODONE && ch == 0 && j > 0 && !waited_o) {   // P region still read by P_{j-1} V_{j-1}: wait as late
          mbar_wait_a(a_o_done, (j - 1) & 1);                        // as possible (the first chunk is already computed)
          tc_fence_after();
        }
        if (CP == 4) tmem_st_x4(tP + ch * CP, pk);
        else if (CP == 8) tmem_st_x8(tP + ch * CP, pk);
        else tmem_st_x16(tP + ch * CP, pk);
      }
      {
        float a0, a1;
        unpack2(fma2(acc0, one2, acc1), a0, a1);
        row_sum += a0 + a1;
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive_a(a_p_full);
    }

    // -------- epilogue: O / l -> bf16 -> out[b, q, h*64 + 32*half ...] --------
    // exchange buffer of parity n_tiles & 1 was last used by tile n_tiles - 2 (or never): free
    const uint32_t xb = (n_tiles & 1) * XBUF;
    sts_f32(a_xw + xb, row_sum);
    named_bar_sync(pair_bar, 64);
    const float total = row_sum + lds_f32(a_xr + xb);
    mbar_wait_a(a_o_done, (n_tiles - 1) & 1);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld_x32(tO, o);
    tmem_wait_ld();
    const int q_row = q0 + t * BQ + row;
    const bool row_ok = q_row < p.seq_q;
    const float inv = p.out_scale / total;
    const int b = bh / p.heads, h = bh - b * p.heads;
    __nv_bfloat16* dst = p.out + ((long long)b * p.seq_q + q_row) * p.ldo + h * DH + half * 32;
    if (PEER && row_ok) {                                        // P2P store into the rank that owns this token row
      const int dest = q_row / p.peer_rows;
      dst = p.peer_out[dest] + ((long long)p.peer_src * p.peer_rows + (q_row - dest * p.peer_rows)) * p.ldo + h * DH + half * 32;
    }
    if (!PEER) {                                                 // local output: direct 16-byte stores (L2 merges the lines)
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[i * 8 + e]) * inv;
          if (p.accumulate) {                                    // read-modify-write in fp32 (previous-window blend)
            const uint4 old = *reinterpret_cast<const uint4*>(dst + i * 8);
            f[0] += bf16_lo(old.x); f[1] += bf16_hi(old.x); f[2] += bf16_lo(old.y); f[3] += bf16_hi(old.y);
            f[4] += bf16_lo(old.z); f[5] += bf16_hi(old.z); f[6] += bf16_lo(old.w); f[7] += bf16_hi(old.w);
          }
          uint4 u;
          u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
          u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
          *reinterpret_cast<uint4*>(dst + i * 8) = u;
        }
      }
    } else {
      // Peer output.  Each thread holds 64 bytes of its row.  The warp transposes through 2 KB of swizzled shared memory (the Q tile of
      // this query tile: every MMA that read it has completed) so that 4 lanes write one row's 64 bytes together —
      // 8 rows x 64 B per store instruction instead of 32 rows x 16 B, which matters for the NVLink peer stores.
      uint8_t* stage = smem + SMEM_Q + t * TILE_BYTES + (half * 4 + quad) * 2048;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16(__uint_as_float(o[i * 8 + 0]) * inv, __uint_as_float(o[i * 8 + 1]) * inv);
        u.y = pack_bf16(__uint_as_float(o[i * 8 + 2]) * inv, __uint_as_float(o[i * 8 + 3]) * inv);
        u.z = pack_bf16(__uint_as_float(o[i * 8 + 4]) * inv, __uint_as_float(o[i * 8 + 5]) * inv);
        u.w = pack_bf16(__uint_as_float(o[i * 8 + 6]) * inv, __uint_as_float(o[i * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(stage + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = u;
      }
      __syncwarp();
      const unsigned long long d = row_ok ? reinterpret_cast<unsigned long long>(dst) : 0ull;
      const int cc = lane & 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2);
        const uint4 v = *reinterpret_cast<const uint4*>(stage + r * 64 + ((cc ^ ((r >> 1) & 3)) << 4));
        const unsigned long long dr = __shfl_sync(0xffffffffu, d, r);
        if (dr) *reinterpret_cast<uint4*>(dr + (cc << 4)) = v;
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(OTHER_REGS));
    // Producer and MMA loops run warp-wide; only the asynchronous instructions are issued by one elected lane, so that
    // descriptor arithmetic stays on the uniform datapath (a divergent single thread costs ~100 clk per tcgen05.mma).
    if (warp == NUM_SOFTMAX_WARPS) {
      // ================================================ TMA producer ============================================
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->q_full, 2 * TILE_BYTES);
        tma_load_3d(smem + SMEM_Q, &tmap_q, &bars->q_full, 0, q0, bh, kEvictFirst);
        tma_load_3d(smem + SMEM_Q + TILE_BYTES, &tmap_q, &bars->q_full, 0, q0 + BQ, bh, kEvictFirst);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const bool seg1 = j >= n_t0;
        const int kv0 = (seg1 ? j - n_t0 : j) * BKV;
        mbar_wait(&bars->k_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->k_full[stage], TILE_BYTES);
          tma_load_3d(smem + SMEM_K + stage * TILE_BYTES, seg1 ? &tmap_k1 : &tmap_k0, &bars->k_full[stage], 0, kv0, bh,
                      kEvictLast);
        }
        __syncwarp();
        mbar_wait(&bars->v_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->v_full[stage], TILE_BYTES);
          tma_load_3d(smem + SMEM_V + stage * TILE_BYTES, seg1 ? &tmap_v1 : &tmap_v0, &bars->v_full[stage], 0, kv0, bh,
                      kEvictLast);
        }
        __syncwarp();
        if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
      }
    } else if (warp <= NUM_SOFTMAX_WARPS + 2) {
      // ================================================ MMA issuers =============================================
      // one issuing warp per query tile: the two tiles' S/P hand-shakes never block each other (the tensor pipe itself
      // interleaves the two instruction streams); K/V stages are released when both warps have committed.
      const int t = (warp == NUM_SOFTMAX_WARPS + 1) ? 0 : 1;
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 0, 1);     // B = V, MN-major
      const uint32_t sq = smem_u32(smem + SMEM_Q) + t * TILE_BYTES;
      const uint32_t sk = smem_u32(smem + SMEM_K);
      const uint32_t sv = smem_u32(smem + SMEM_V);
      const uint32_t tmS = tmem_base + t * COL_TILE + COL_S;
      const uint32_t tmP = tmem_base + t * COL_TILE + COL_P;
      const uint32_t tmO = tmem_base + t * COL_TILE + COL_O;

      auto issue_qk = [&](int stage) {
        if (elect_one()) {
          const uint64_t adesc = make_desc_sw128(sq, 1024, 0);
          const uint64_t bdesc = make_desc_sw128(sk + stage * TILE_BYTES, 1024, 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) mma_ss(tmS, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
          tc_commit(&bars->s_full[t]);
          tc_commit(&bars->k_empty[stage]);
        }
        __syncwarp();
      };

      mbar_wait(&bars->q_full, 0);
      mbar_wait(&bars->k_full[0], 0);
      tc_fence_after();
      issue_qk(0);

      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t par = j & 1;
        if (j + 1 < n_tiles) {
          const int nstage = (stage + 1 == KV_STAGES) ? 0 : stage + 1;
          const uint32_t nphase = (stage + 1 == KV_STAGES) ? (phase ^ 1) : phase;
          mbar_wait(&bars->k_full[nstage], nphase);
          mbar_wait(&bars->s_free[t], par);                    // the softmax warps hold S_t(j) in registers
          tc_fence_after();
          issue_qk(nstage);
        }
        mbar_wait(&bars->v_full[stage], phase);
        mbar_wait(&bars->p_full[t], par);                      // P_t(j) is in TMEM
        tc_fence_after();
        if (elect_one()) {
          // V tile [128 keys][64 d] as MN-major B: 8-key groups are 1024 B apart, 16 keys per MMA = 2048 B
          const uint64_t vdesc = make_desc_sw128(sv + stage * TILE_BYTES, 1024, 1024);
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) mma_ts(tmO, tmP + k * 8, vdesc + (uint64_t)(128 * k), idesc_pv, (j | k) != 0);
          tc_commit(&bars->o_done[t]);
          tc_commit(&bars->v_empty[stage]);
        }
        __syncwarp();
        if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
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

int make_map3(CUtensorMap* map, const void* ptr, long long bh, long long len) {
  uint64_t dims[3] = {(uint64_t)DH, (uint64_t)len, (uint64_t)bh};
  uint64_t str[2] = {(uint64_t)DH * 2, (uint64_t)len * DH * 2};
  uint32_t box[3] = {DH, BKV, 1};
  return make_tmap_bf16(map, ptr, 3, dims, str, box);
}

}  // namespace

int launch_attention_v1(const void* q, const void* k0, const void* v0, const void* k1, const void* v1, const AttnParams& p,
                        cudaStream_t st) {
  VP_REQUIRE(p.batch > 0 && p.heads > 0 && p.seq_q > 0 && p.kv_len0 > 0 && p.kv_len1 >= 0, VP_ERR_BAD_SHAPE,
             "attention: bad shape");
  VP_REQUIRE(p.ldo % 8 == 0, VP_ERR_BAD_ALIGN, "attention: output leading dim must be a multiple of 8");
  VP_REQUIRE(p.kv_len1 == 0 || (k1 && v1), VP_ERR_BAD_SHAPE, "attention: second K/V segment missing");
  VP_REQUIRE(p.peer_out[0] == nullptr || (p.batch == 1 && p.peer_rows > 0 && !p.accumulate), VP_ERR_UNSUPPORTED,
             "attention: peer output needs batch 1 and no accumulation");
  int rc0;
  if ((rc0 = configure_once(reinterpret_cast<const void*>(attn_fwd_kernel<false>), SMEM_BYTES))) return rc0;
  if ((rc0 = configure_once(reinterpret_cast<const void*>(attn_fwd_kernel<true>), SMEM_BYTES))) return rc0;
  const long long bh = (long long)p.batch * p.heads;
  CUtensorMap mq, mk0, mv0, mk1, mv1;
  int rc;
  if ((rc = make_map3(&mq, q, bh, p.seq_q))) return rc;
  if ((rc = make_map3(&mk0, k0, bh, p.kv_len0))) return rc;
  if ((rc = make_map3(&mv0, v0, bh, p.kv_len0))) return rc;
  if (p.kv_len1 > 0) {
    if ((rc = make_map3(&mk1, k1, bh, p.kv_len1))) return rc;
    if ((rc = make_map3(&mv1, v1, bh, p.kv_len1))) return rc;
  } else {
    mk1 = mk0;
    mv1 = mv0;
  }
  dim3 grid((p.seq_q + 2 * BQ - 1) / (2 * BQ), (unsigned)bh);
  if (p.peer_out[0]) attn_fwd_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mq, mk0, mv0, mk1, mv1, p);
  else attn_fwd_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mq, mk0, mv0, mk1, mv1, p);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

}  // namespace vp
