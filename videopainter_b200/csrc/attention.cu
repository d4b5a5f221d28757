// Flash attention for the joint text+video sequence of CogVideoX (AP:2192-2197: non-causal, no mask, d_head = 64,
// scale 1/8), tcgen05 + TMEM + TMA.  One CTA owns 256 query rows of one (batch, head); K/V stream through a
// 3-stage TMA ring in 128-key tiles; up to two K/V segments are attended in one softmax (the ID-resample
// processor concatenates a second, masked K/V copy: AP:2283-2284).
//
//   warps 0-3   softmax of query tile 0 (thread == query row; S read from TMEM, P written back to TMEM as bf16)
//   warps 4-7   softmax of query tile 1
//   warp  8     TMA producer (Q once, then K_j / V_j)
//   warp  9     MMA issuer   (S_t = Q_t K_jᵀ : SS-MMA 128x128x64;  O_t += P_t V_j : TS-MMA 128x64x128, V MN-major)
//   warp 10     TMEM allocator (512 columns: per tile S 128 | P 64 | O 64)
// The running maximum is only refreshed when it grows by more than 2^8 (lazy rescale), so the O accumulator in
// TMEM is rarely touched by the softmax warps.
#include "attention.cuh"
#include "host_util.cuh"

namespace vp {

namespace {

constexpr int BQ = 128;          // query rows per softmax warpgroup
constexpr int BKV = 128;         // keys per tile
constexpr int DH = 64;           // head dim
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = BKV * DH * 2;   // 16 KiB (Q tile has the same size)
constexpr int SMEM_Q = 0;
constexpr int SMEM_K = 2 * TILE_BYTES;
constexpr int SMEM_V = SMEM_K + KV_STAGES * TILE_BYTES;
constexpr int SMEM_BAR = SMEM_V + KV_STAGES * TILE_BYTES;
constexpr int SMEM_BYTES = SMEM_BAR + 256 + 1024;
constexpr int NUM_THREADS = 384;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192, COL_TILE = 256;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units

struct Bars {
  uint64_t q_full;
  uint64_t k_full[KV_STAGES], k_empty[KV_STAGES];
  uint64_t v_full[KV_STAGES], v_empty[KV_STAGES];
  uint64_t s_full[2], s_free[2], p_full[2], o_done[2];
  uint32_t tmem_slot;
};

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

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k0);
    tma_prefetch_desc(&tmap_v0);
    if (n_t1 > 0) {
      tma_prefetch_desc(&tmap_k1);
      tma_prefetch_desc(&tmap_v1);
    }
  }
  if (warp == 9 && lane == 0) {
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bars->s_full[t], 1);
      mbar_init(&bars->s_free[t], 128);
      mbar_init(&bars->p_full[t], 128);
      mbar_init(&bars->o_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc(&bars->tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp < 8) {
    // ================================================ softmax ===================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;\n");
    const int t = warp >> 2;                                  // query tile of this warpgroup
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + t * COL_TILE + COL_S + lane_base;
    const uint32_t tP = tmem_base + t * COL_TILE + COL_P + lane_base;
    const uint32_t tO = tmem_base + t * COL_TILE + COL_O + lane_base;
    const float c = p.scale_log2;
    float m_used = -INFINITY;     // maximum the exponents are currently referenced to (raw score units)
    float row_sum = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t par = j & 1;
      mbar_wait(&bars->s_full[t], par);
      tc_fence_after();
      uint32_t sr[128];
      tmem_ld_x32(tS + 0, sr + 0);
      tmem_ld_x32(tS + 32, sr + 32);
      tmem_ld_x32(tS + 64, sr + 64);
      tmem_ld_x32(tS + 96, sr + 96);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&bars->s_free[t]);                          // S_t may be overwritten by the next QKᵀ

      // ragged last tile of a segment: keys beyond the segment end do not exist
      int valid = BKV;
      if (j == n_t0 - 1) valid = p.kv_len0 - (n_t0 - 1) * BKV;
      else if (j == n_tiles - 1 && n_t1 > 0) valid = p.kv_len1 - (n_t1 - 1) * BKV;
      if (valid < BKV) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= valid) sr[i] = 0xff800000u;                // -inf
      }

      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(sr[i + 0]));
        mx1 = fmaxf(mx1, __uint_as_float(sr[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(sr[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(sr[i + 3]));
      }
      const float tile_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

      bool waited_o = false;
      const bool need = (tile_max - m_used) * c > RESCALE_THRESHOLD;   // true at j == 0 (m_used = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_used, tile_max);
        const float factor = fast_exp2((m_used - m_new) * c);          // exp2(-inf) = 0 at j == 0
        m_used = m_new;
        row_sum *= factor;
        if (j > 0) {
          mbar_wait(&bars->o_done[t], (j - 1) & 1);                    // P_{j-1} V_{j-1} finished
          tc_fence_after();
          waited_o = true;
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {                             // 32 columns at a time: S is still live
            uint32_t o[32];
            tmem_ld_x32(tO + cc * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_x32(tO + cc * 32, o);
          }
        }
      }

      const float neg_mc = -m_used * c;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        const float e0 = fast_exp2(fmaf(__uint_as_float(sr[i + 0]), c, neg_mc));
        const float e1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), c, neg_mc));
        const float e2 = fast_exp2(fmaf(__uint_as_float(sr[i + 2]), c, neg_mc));
        const float e3 = fast_exp2(fmaf(__uint_as_float(sr[i + 3]), c, neg_mc));
        s0 += e0; s1 += e1; s2 += e2; s3 += e3;
        pk[i / 2 + 0] = pack_bf16(e0, e1);
        pk[i / 2 + 1] = pack_bf16(e2, e3);
      }
      row_sum += (s0 + s1) + (s2 + s3);

      if (j > 0 && !waited_o) {                                        // P region still read by P_{j-1} V_{j-1}
        mbar_wait(&bars->o_done[t], (j - 1) & 1);
        tc_fence_after();
      }
      tmem_st_x32(tP, pk);
      tmem_st_x32(tP + 32, pk + 32);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bars->p_full[t]);
    }

    // -------- epilogue: O / l -> bf16 -> out[b, q, h*64 ...] --------
    mbar_wait(&bars->o_done[t], (n_tiles - 1) & 1);
    tc_fence_after();
    uint32_t o[64];
    tmem_ld_x32(tO, o);
    tmem_ld_x32(tO + 32, o + 32);
    tmem_wait_ld();
    const int q_row = q0 + t * BQ + row;
    if (q_row < p.seq_q) {
      const float inv = p.out_scale / row_sum;
      const int b = bh / p.heads, h = bh - b * p.heads;
      __nv_bfloat16* dst = p.out + ((long long)b * p.seq_q + q_row) * p.ldo + h * DH;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[i * 8 + e]) * inv;
        if (p.accumulate) {
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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
    if (warp == 8) {
      // ================================================ TMA producer ============================================
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars->q_full, 2 * TILE_BYTES);
        tma_load_3d(smem + SMEM_Q, &tmap_q, &bars->q_full, 0, q0, bh, kEvictFirst);
        tma_load_3d(smem + SMEM_Q + TILE_BYTES, &tmap_q, &bars->q_full, 0, q0 + BQ, bh, kEvictFirst);
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_tiles; ++j) {
          const bool seg1 = j >= n_t0;
          const int kv0 = (seg1 ? j - n_t0 : j) * BKV;
          mbar_wait(&bars->k_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bars->k_full[stage], TILE_BYTES);
          tma_load_3d(smem + SMEM_K + stage * TILE_BYTES, seg1 ? &tmap_k1 : &tmap_k0, &bars->k_full[stage], 0, kv0, bh,
                      kEvictLast);
          mbar_wait(&bars->v_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bars->v_full[stage], TILE_BYTES);
          tma_load_3d(smem + SMEM_V + stage * TILE_BYTES, seg1 ? &tmap_v1 : &tmap_v0, &bars->v_full[stage], 0, kv0, bh,
                      kEvictLast);
          if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 9) {
      // ================================================ MMA issuer ==============================================
      if (lane == 0) {
        constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 0, 1);     // B = V, MN-major
        const uint32_t sq = smem_u32(smem + SMEM_Q);
        const uint32_t sk = smem_u32(smem + SMEM_K);
        const uint32_t sv = smem_u32(smem + SMEM_V);

        auto issue_qk = [&](int t, int stage) {
          const uint64_t adesc = make_desc_sw128(sq + t * TILE_BYTES, 1024, 0);
          const uint64_t bdesc = make_desc_sw128(sk + stage * TILE_BYTES, 1024, 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            mma_ss(tmem_base + t * COL_TILE + COL_S, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
          tc_commit(&bars->s_full[t]);
        };

        mbar_wait(&bars->q_full, 0);
        mbar_wait(&bars->k_full[0], 0);
        tc_fence_after();
        issue_qk(0, 0);
        issue_qk(1, 0);
        tc_commit(&bars->k_empty[0]);

        for (int j = 0; j < n_tiles; ++j) {
          const int stage = j % KV_STAGES;
          const uint32_t phase = (j / KV_STAGES) & 1;
          const uint32_t par = j & 1;
          if (j + 1 < n_tiles) {
            const int nstage = (j + 1) % KV_STAGES;
            const uint32_t nphase = ((j + 1) / KV_STAGES) & 1;
            mbar_wait(&bars->k_full[nstage], nphase);
            for (int t = 0; t < 2; ++t) {
              mbar_wait(&bars->s_free[t], par);                // softmax t holds S_t(j) in registers
              tc_fence_after();
              issue_qk(t, nstage);
            }
            tc_commit(&bars->k_empty[nstage]);
          }
          mbar_wait(&bars->v_full[stage], phase);
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&bars->p_full[t], par);                  // P_t(j) is in TMEM
            tc_fence_after();
            // V tile [128 keys][64 d] as MN-major B: 8-key groups are 1024 B apart, 16 keys per MMA = 2048 B
            const uint64_t vdesc = make_desc_sw128(sv + stage * TILE_BYTES, 1024, 1024);
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k)
              mma_ts(tmem_base + t * COL_TILE + COL_O, tmem_base + t * COL_TILE + COL_P + k * 8, vdesc + (uint64_t)(128 * k),
                     idesc_pv, (j | k) != 0);
            tc_commit(&bars->o_done[t]);
          }
          tc_commit(&bars->v_empty[stage]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
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

int launch_attention(const void* q, const void* k0, const void* v0, const void* k1, const void* v1, const AttnParams& p,
                     cudaStream_t st) {
  VP_REQUIRE(p.batch > 0 && p.heads > 0 && p.seq_q > 0 && p.kv_len0 > 0 && p.kv_len1 >= 0, VP_ERR_BAD_SHAPE,
             "attention: bad shape");
  VP_REQUIRE(p.ldo % 8 == 0, VP_ERR_BAD_ALIGN, "attention: output leading dim must be a multiple of 8");
  VP_REQUIRE(p.kv_len1 == 0 || (k1 && v1), VP_ERR_BAD_SHAPE, "attention: second K/V segment missing");
  static bool configured = false;
  if (!configured) {
    VP_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
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
  attn_fwd_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mq, mk0, mv0, mk1, mv1, p);
  VP_CHECK_CUDA(cudaGetLastError());
  return VP_OK;
}

}  // namespace vp
