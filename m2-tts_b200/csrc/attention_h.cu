// attention_h.cu — flash attention on tcgen05 with a 16-BIT split: every operand x is carried as two fp16 numbers,
// x = hi + lo (hi = fp16(x), lo = fp16(x - hi): 22 significant bits, the same as the TF32 split of attention_tc.cu),
// and a product is a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with fp32 accumulation in TMEM. Against the TF32 split this
// halves the operand bytes (K = 16 per UMMA instead of 8) and therefore the UMMA count — the tensor pipe here is
// bound by UMMA instructions of ~45-64 cycles each, not by math (mma_bench.cu). Accuracy on the reference's value
// ranges is that of the TF32 split (tools: the CPU emulation in DESIGN.md §4.1; parity tests <= 1e-4); fp16 needs
// |x| < 65504, which the producer (QKV GEMM epilogue) enforces by saturating — LayerNorm'd activations times
// Xavier-scale weights are O(1-10).
//
// Same organisation as attention_ws_kernel (attention_tc.cu): one CTA = two 128-query tiles of one (utterance, head)
// sharing every K/V tile; warp 0 TMA loader, warp 1 UMMA issuer (warp-collective), warps 4-7 / 8-11 softmax
// warpgroups; lazy rescaling with O accumulating in TMEM. With the tensor work halved the softmax warpgroups are
// the bottleneck, so the SCORES ARE DOUBLE-BUFFERED in TMEM: QK(t+2) is issued right behind PV(t), a warpgroup
// finds S(t+1) waiting when it finishes tile t, and never idles on the tensor pipe. (Double buffering needs the 64-column
// score tile, so Q K^T keeps its three terms as separate N = 64 UMMAs; P V folds V_hi|V_lo into N = 2 hd.)
// Operands: qkvh[6][B][nh][hd][Lp] fp16 = {Q_hi,Q_lo,K_hi,K_lo,V_hi,V_lo}, positions contiguous, Lp % 8 == 0, Q
// pre-multiplied by scale*log2(e). A TMA box = 64 positions (128 B) x hd rows: for Q and K the MN-major 128B-swizzle
// operand (K dim = d), for V the K-major 128B-swizzle operand (rows = d, K dim = keys) — layouts verified with
// umma_probe_f16.cu. P is written back to TMEM as packed fp16 pairs (low half = even key).
#include "attention_tc.cuh"
#include <cuda_fp16.h>

namespace m2 {

constexpr int AH_THREADS = 128 + 512;      // loader, two issuers, idle | four softmax warpgroups (two per query tile)
constexpr uint32_t AH_TMEM_COLS = 512;
// TMEM per query tile (tile stride 256 columns): two score buffers of 64 columns (tile t uses buffer t & 1; P(t) is
// packed over it: P_hi in columns 0:32, P_lo in 32:64 of the buffer) and O in 2 hd columns from 128.
constexpr uint32_t AH_COL_S = 0, AH_COL_PLO = 32, AH_COL_O = 128;
constexpr uint32_t AH_COL_TILE = 256;
constexpr int AH_STAGES = 4;        // K / V ring depth

template <int HD>
struct AhSmem {
  static constexpr uint32_t box = HD * 128;                  // 64 positions x hd rows of fp16
  static constexpr uint32_t q_bytes = 4 * box;               // one query tile: hi (2 boxes) + lo (2 boxes)
  static constexpr uint32_t kv_bytes = 2 * box;              // one K (or V) tile: hi box + lo box
  static constexpr uint32_t off_k = 2 * q_bytes;
  static constexpr uint32_t off_v = off_k + AH_STAGES * kv_bytes;
  static constexpr uint32_t off_bar = off_v + AH_STAGES * kv_bytes;
  static constexpr uint32_t off_exch = off_bar + 512;          // float [parity 2][tile 2][half 2][128 rows]: row maxima / final row sums
  static constexpr uint32_t total = off_exch + 4096 + 1024 /*align slack*/;
};

__host__ __device__ constexpr uint32_t ah_idesc(int M, int N, int mn_major) {   // kind::f16: fp16 x fp16 -> fp32
  return (1u << 4) | ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ss_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

template <int HD>
__global__ void __launch_bounds__(AH_THREADS, 1)
attention_h_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ ctx, const int64_t* __restrict__ lengths,
                   int B, int L, int nh, float* __restrict__ ctx_lo, __half* __restrict__ ctx_h, long long* __restrict__ prof, int dbg_skip,
                   int32_t* __restrict__ status) {
  static_assert(HD % 16 == 0 && HD >= 16 && HD <= 64, "16-bit split attention: head_dim in {16,32,48,64}");
  constexpr uint32_t BOX = AhSmem<HD>::box;
  constexpr int KSTEPS_D = HD / 16;
  constexpr uint32_t IDESC_QK1 = ah_idesc(TC_BQ, TC_BK, 1);       // Q_* x K_*
  constexpr uint32_t IDESC_PV2 = ah_idesc(TC_BQ, 2 * HD, 0);      // P_hi x [V_hi | V_lo]
  constexpr uint32_t IDESC_PV1 = ah_idesc(TC_BQ, HD, 0);          // P_lo x V_hi

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = sbase;                                   // [tile][hi|lo][2 boxes][HD rows][128 B]
  const uint32_t sK = sbase + AhSmem<HD>::off_k;               // [stage][hi|lo][HD rows][128 B]
  const uint32_t sV = sbase + AhSmem<HD>::off_v;
  const uint32_t sBar = sbase + AhSmem<HD>::off_bar;
  // barriers: q_full[2] s_full[2 tiles][2 buffers] p_ready[2] pv_done[2] | k_full[S] k_empty[S] v_full[S] v_empty[S]
  const uint32_t bar_qf = sBar, bar_sf = sBar + 16, bar_pr = sBar + 48, bar_pv = sBar + 64;
  const uint32_t bar_kf = sBar + 80, bar_ke = bar_kf + 8 * AH_STAGES, bar_vf = bar_ke + 8 * AH_STAGES, bar_ve = bar_vf + 8 * AH_STAGES;
  const uint32_t tmem_slot = bar_ve + 8 * AH_STAGES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // 1-D grid, longest first: all CTAs with two query tiles, then (odd tile count) the single-tile CTAs, which take about
  // 0.6 of the time and fill the last wave
  const int n_tiles_q = (L + TC_BQ - 1) / TC_BQ, npf = n_tiles_q >> 1, n_long = npf * nh * B;
  int qx, bh;
  if ((int)blockIdx.x < n_long) { bh = (int)blockIdx.x / npf; qx = (int)blockIdx.x % npf; }
  else { bh = (int)blockIdx.x - n_long; qx = npf; }
  const int q0 = qx * (2 * TC_BQ), head = bh % nh, b = bh / nh;
  const int ntq = q0 + TC_BQ < L ? 2 : 1;

  int Leff = L;
  bool all_masked = false;
  if (lengths != nullptr) {
    const long long len = lengths[b];
    if (len <= 0) all_masked = true;
    else if (len < L) Leff = (int)len;
  }
  const int nkt = (Leff + TC_BK - 1) / TC_BK;
  const int plane = B * nh * HD;
  const int row_q = (b * nh + head) * HD;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_qf + 8 * i, 1); mbar_init(bar_sf + 16 * i, 1); mbar_init(bar_sf + 16 * i + 8, 1);
      mbar_init(bar_pr + 8 * i, 8); mbar_init(bar_pv + 8 * i, 1);
    }
    for (int i = 0; i < AH_STAGES; ++i) {
      mbar_init(bar_kf + 8 * i, 1); mbar_init(bar_ke + 8 * i, (uint32_t)ntq); mbar_init(bar_vf + 8 * i, 1); mbar_init(bar_ve + 8 * i, (uint32_t)ntq);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(AH_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ===== loader =====
      for (int x = 0; x < ntq; ++x) {
        mbar_expect_tx(bar_qf + 8 * x, AhSmem<HD>::q_bytes);
        for (int h = 0; h < 2; ++h)
          for (int j = 0; j < 2; ++j)
            tma_load_2d(sQ + (uint32_t)x * AhSmem<HD>::q_bytes + (h * 2 + j) * BOX, &tmap, q0 + x * TC_BQ + j * 64, h * plane + row_q,
                        bar_qf + 8 * x);
      }
      for (int t = 0; t < nkt; ++t) {
        const int st = t % AH_STAGES;
        const uint32_t par_prev = (uint32_t)(((t / AH_STAGES) - 1) & 1);
        if (t >= AH_STAGES) mbar_wait(bar_ke + 8 * st, par_prev);
        mbar_expect_tx(bar_kf + 8 * st, AhSmem<HD>::kv_bytes);
        for (int h = 0; h < 2; ++h)
          tma_load_2d(sK + (uint32_t)st * AhSmem<HD>::kv_bytes + h * BOX, &tmap, t * TC_BK, (2 + h) * plane + row_q, bar_kf + 8 * st);
        if (t >= AH_STAGES) mbar_wait(bar_ve + 8 * st, par_prev);
        mbar_expect_tx(bar_vf + 8 * st, AhSmem<HD>::kv_bytes);
        for (int h = 0; h < 2; ++h)
          tma_load_2d(sV + (uint32_t)st * AhSmem<HD>::kv_bytes + h * BOX, &tmap, t * TC_BK, (4 + h) * plane + row_q, bar_vf + 8 * st);
      }
    }
  } else if ((warp == 1 || warp == 2) && warp - 1 < ntq) {
    // ===== UMMA issuer of query tile x (whole warp, one elected lane issues) =====
    const int x = warp - 1;
    const bool pr_on = prof != nullptr && blockIdx.x == 0;
    auto issue_qk = [&](int x, int st, int buf) {
      if (dbg_skip & 1) return;      // bring-up timing experiment (M2TTS_ATT_DBG): results invalid
      const uint32_t q = sQ + (uint32_t)x * AhSmem<HD>::q_bytes, k = sK + (uint32_t)st * AhSmem<HD>::kv_bytes;
      const uint32_t d = tmem_base + (uint32_t)x * AH_COL_TILE + AH_COL_S + (uint32_t)buf * 64u;
      // MN-major, 128B swizzle, 16-bit: LBO = next 64 positions (next box), SBO = next 8 d-rows (1024 B);
      // one k-step = 16 d-rows = 2048 B. Terms: hi*hi, hi*lo, lo*hi.
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t qa = q + (term == 2 ? 2 * BOX : 0u), kb = k + (term == 1 ? BOX : 0u);
#pragma unroll
        for (int ks = 0; ks < KSTEPS_D; ++ks)
          umma_f16_ss_w(d, umma_desc(qa + ks * 2048u, BOX, 1024u, 2u), umma_desc(kb + ks * 2048u, BOX, 1024u, 2u), IDESC_QK1,
                        (term | ks) ? 1u : 0u);
      }
    };
    auto issue_pv = [&](int x, int st, int buf, uint32_t accumulate) {
      if (dbg_skip & 2) return;
      const uint32_t v = sV + (uint32_t)st * AhSmem<HD>::kv_bytes;
      const uint32_t tb = tmem_base + (uint32_t)x * AH_COL_TILE;
      const uint32_t pb = tb + AH_COL_S + (uint32_t)buf * 64u;
      // V^T box: rows = d (V_hi rows then V_lo rows), keys contiguous; one k-step = 16 keys = 32 B = 8 TMEM columns of P
#pragma unroll
      for (int ks = 0; ks < TC_BK / 16; ++ks)
        umma_f16_ts_w(tb + AH_COL_O, pb + ks * 8, umma_desc(v + ks * 32u, 16u, 1024u, 2u), IDESC_PV2, ks ? 1u : accumulate);
#pragma unroll
      for (int ks = 0; ks < TC_BK / 16; ++ks)
        umma_f16_ts_w(tb + AH_COL_O, pb + AH_COL_PLO + ks * 8, umma_desc(v + ks * 32u, 16u, 1024u, 2u), IDESC_PV1, 1u);
    };
    // One issuer warp per query tile (warps 1 and 2 sit on different schedulers): an UMMA costs its issuer ~12 instructions
    // (descriptor moves into uniform registers, elect), ~50 cycles next to two busy softmax warps, and a single issuer
    // for both tiles was the critical path (tools/attn_prof.py: 1870 of 2355 cycles per key tile spent issuing).
    // A K/V stage is free when BOTH issuers' UMMAs on it have completed (k_empty / v_empty count = tiles).
    // prologue: the scores of key tiles 0 and 1
    for (int t = 0; t < 2 && t < nkt; ++t) {
      mbar_wait(bar_kf + 8 * t, 0);
      if (t == 0) mbar_wait(bar_qf + 8 * x, 0);
      tc_fence_after();
      issue_qk(x, t, t);
      tc_commit_w(bar_sf + 16 * x + 8 * t);
      tc_commit_w(bar_ke + 8 * t);
    }
    for (int t = 0; t < nkt; ++t) {
      const int st = t % AH_STAGES, s2 = (t + 2) % AH_STAGES, buf = t & 1;
      const uint32_t par = (uint32_t)(t & 1);
      if (pr_on && lane == 0 && t >= 8 && t < 40) prof[384 + (t - 8) * 8 + 3 * x] = clock64();
      mbar_wait(bar_pr + 8 * x, par);                         // P(x,t) is in TMEM (and O has been rescaled if needed)
      if (pr_on && lane == 0 && t >= 8 && t < 40) prof[384 + (t - 8) * 8 + 3 * x + 1] = clock64();
      mbar_wait(bar_vf + 8 * st, (uint32_t)((t / AH_STAGES) & 1));
      tc_fence_after();
      issue_pv(x, st, buf, t > 0 ? 1u : 0u);
      tc_commit_w(bar_pv + 8 * x);
      tc_commit_w(bar_ve + 8 * st);
      if (t + 2 < nkt) {                                      // score buffer `buf` is free again once PV(x,t) has read P
        mbar_wait(bar_kf + 8 * s2, (uint32_t)(((t + 2) / AH_STAGES) & 1));
        tc_fence_after();
        issue_qk(x, s2, buf);
        tc_commit_w(bar_sf + 16 * x + 8 * buf);
        tc_commit_w(bar_ke + 8 * s2);
      }
      if (pr_on && lane == 0 && t >= 8 && t < 40) prof[384 + (t - 8) * 8 + 3 * x + 2] = clock64();
    }
  } else if (warp >= 4 && (((warp - 4) >> 2) & 1) < ntq) {
    // ===== softmax warpgroups: query tile x has TWO of them (warps 4-7 / 12-15 for tile A, 8-11 / 16-19 for tile B); thread =
    // query row = TMEM lane, warpgroup `half` owns 32 of the 64 score columns of every key tile. Four softmax warps per
    // scheduler instead of two: the loop is bound by MUFU (quarter rate) and by the latency of its own dependent
    // instructions, not by issue slots (tools/attn_prof.py: 2054 cycles per key tile even with the UMMAs switched off).
    // The two warps that share a row exchange their partial row maxima through shared memory at a 64-thread named barrier,
    // which also orders "both have loaded their scores" before either packs P over the score columns.
    const int x = ((warp - 4) >> 2) & 1, half = (warp - 4) >> 3;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)x * AH_COL_TILE;
    float* exch = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)) + AhSmem<HD>::off_exch);
    const int pair_bar = 1 + x * 4 + (warp & 3);          // named barrier of the two warps sharing these 32 rows
    float m_ref = -INFINITY, l_run = 0.f;
    const bool pw = prof != nullptr && blockIdx.x == 0 && x == 0 && half == 0 && row == 0;
    for (int t = 0; t < nkt; ++t) {
      const bool pt = pw && t >= 8 && t < 40;
      long long* pp = prof + (pt ? (t - 8) * 8 : 0);
      const uint32_t t_s = t_lane + AH_COL_S + (uint32_t)(t & 1) * 64u;     // this tile's score buffer (P goes over it)
      if (pt) pp[0] = clock64();
      mbar_wait(bar_sf + 16 * x + 8 * (t & 1), (uint32_t)((t >> 1) & 1));
      if (pt) pp[1] = clock64();
      __syncwarp();
      tc_fence_after();
      uint32_t sr[32];
      tmem_ld16(t_s + half * 32, sr);
      tmem_ld16(t_s + half * 32 + 16, sr + 16);
      tmem_wait_ld();
      if (pt) pp[2] = clock64();
      const int kbase = t * TC_BK + half * 32;
      if (all_masked || kbase + 32 > Leff) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = __uint_as_float(sr[j]);
          if (all_masked) v = (kbase + j < L) ? 0.f : -INFINITY;
          else if (kbase + j >= Leff) v = -INFINITY;
          sr[j] = __float_as_uint(v);
        }
      }
      float mx = __uint_as_float(sr[0]);
#pragma unroll
      for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(sr[j]));
      {  // row maximum over all 64 columns
        float* e = exch + ((t & 1) * 4 + x * 2) * 128;
        e[half * 128 + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        mx = fmaxf(mx, e[(half ^ 1) * 128 + row]);
      }
      if (t == 0) {
        m_ref = mx;
      }
      // The pv_done barrier completes one phase per key tile. We wait for phase t-1 in EVERY tile (a parity wait is
      // only exact while the waiter is at most one phase behind): early when O has to be rescaled, otherwise at the
      // end of the tile, when PV(t-1) has long landed and the wait costs nothing.
      bool pv_waited = t == 0;
      if (t > 0 && __any_sync(0xffffffffu, mx > m_ref + 8.0f)) {
        // lazy rescale (both warps of the pair take the same decision: they see the same maxima): PV(t-1) has landed and
        // PV(t) waits for our arrival below, so O is quiescent; this half rescales its hd of the 2 hd accumulator columns
        mbar_wait(bar_pv + 8 * x, (uint32_t)((t - 1) & 1));
        pv_waited = true;
        tc_fence_after();
        const float m_new = fmaxf(m_ref, mx);
        const float alpha = ws_ex2(m_ref - m_new);
        m_ref = m_new;
        l_run *= alpha;
#pragma unroll
        for (int c = 0; c < HD; c += 16) {
          uint32_t orr[16];
          tmem_ld16(t_lane + AH_COL_O + half * HD + c, orr);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) orr[j] = __float_as_uint(__uint_as_float(orr[j]) * alpha);
          tmem_st16(t_lane + AH_COL_O + half * HD + c, orr);
        }
      }
      if (pt) pp[3] = clock64();
      float rs = 0.f;
      {        // 32 keys -> 16 packed columns of P_hi and of P_lo
        uint32_t ph[16], pl[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p0 = ws_ex2(__uint_as_float(sr[2 * j]) - m_ref);
          const float p1 = ws_ex2(__uint_as_float(sr[2 * j + 1]) - m_ref);
          rs += p0 + p1;
          const __half2 h = __floats2half2_rn(p0, p1);
          const float2 hf = __half22float2(h);
          const __half2 lo = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
          ph[j] = *reinterpret_cast<const uint32_t*>(&h);
          pl[j] = *reinterpret_cast<const uint32_t*>(&lo);
        }
        tmem_st16(t_s + half * 16, ph);                      // P overwrites the scores in place
        tmem_st16(t_s + AH_COL_PLO + half * 16, pl);
      }
      l_run += rs;
      if (pt) pp[4] = clock64();
      tmem_wait_st();
      if (pt) pp[5] = clock64();
      if (!pv_waited) mbar_wait(bar_pv + 8 * x, (uint32_t)((t - 1) & 1));
      if (pt) pp[6] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_pr + 8 * x) : "memory");
    }
    // total row sum: the second warpgroup of the tile publishes its part and is done
    {
      float* e = exch + ((nkt & 1) * 4 + x * 2) * 128;
      if (half == 1) e[128 + row] = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      if (half == 0) l_run += e[128 + row];
    }
    if (half == 0) {
    mbar_wait(bar_pv + 8 * x, (uint32_t)((nkt - 1) & 1));     // the last PV has landed: O is complete
    __syncwarp();
    tc_fence_after();
    float o[HD];
#pragma unroll
    for (int c = 0; c < HD; c += 16) {
      uint32_t orr[16], or2[16];
      tmem_ld16(t_lane + AH_COL_O + c, orr);
      tmem_ld16(t_lane + AH_COL_O + HD + c, or2);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) o[c + j] = __uint_as_float(orr[j]) + __uint_as_float(or2[j]);
    }
    const int qi = q0 + x * TC_BQ + row;
    if (qi < L) {
      const float inv = 1.0f / l_run;
      float* dst = ctx + ((long long)b * L + qi) * (nh * HD) + head * HD;
      if (ctx_h != nullptr) {   // fp16 hi/lo planes [2][B*L][nh*HD] for the 16-bit split out_proj (lin_h.cu)
        __half* dh = ctx_h + ((long long)b * L + qi) * (nh * HD) + head * HD;
        __half* dl = dh + (long long)B * L * (nh * HD);
        bool bad = false;
#pragma unroll
        for (int c = 0; c < HD; c += 8) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) h_split2(o[c + 2 * e] * inv, o[c + 2 * e + 1] * inv, hi[e], lo[e], bad);
          *reinterpret_cast<uint4*>(dh + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dl + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        h_flag(bad, status);
      } else if (ctx_lo == nullptr) {
#pragma unroll
        for (int c = 0; c < HD; c += 4)
          *reinterpret_cast<float4*>(dst + c) = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
      } else {   // TF32 hi/lo planes for the tensor-core out_proj
        float* dlo = ctx_lo + ((long long)b * L + qi) * (nh * HD) + head * HD;
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
          float h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float v = o[c + e] * inv; h[e] = __uint_as_float(tf32_hi(v)); l[e] = __uint_as_float(tf32_hi(v - h[e])); }
          *reinterpret_cast<float4*>(dst + c) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(dlo + c) = make_float4(l[0], l[1], l[2], l[3]);
        }
      }
    }
    }      // half == 0
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(AH_TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFnH)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnH ah_encode_fn() {
  static EncodeTiledFnH fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFnH)p;
  }
  return fn;
}

extern long long* g_ws_prof;      // attention_tc.cu (m2tts_attention_set_prof)

template <int HD>
static int launch_ah_hd(const CUtensorMap& tmap, float* ctx, const int64_t* lengths, int B, int L, int nh, cudaStream_t s, float* ctx_lo,
                        __half* ctx_h, int32_t* status) {
  const size_t smem = AhSmem<HD>::total;
  static int dbg_skip = -1;
  if (dbg_skip < 0) dbg_skip = tools_env_int("M2TTS_ATT_DBG", 0);
  M2_CUDA_OK(allow_smem(attention_h_kernel<HD>, smem));
  dim3 grid((unsigned)(ceil_div(L, 2 * TC_BQ) * nh * B), 1, 1);
  M2_LAUNCH(M2TTS_STAGE_ATTENTION, attention_h_kernel<HD>, grid, AH_THREADS, smem, s, tmap, ctx, lengths, B, L, nh, ctx_lo, ctx_h,
            tools_env_int("M2TTS_LIN_PROF_STAGE", -1) >= 0 ? nullptr : g_ws_prof, dbg_skip, status);      // the buffer belongs to tools/lin_prof.py then
  return M2TTS_OK;
}

// qkvh: [6][B][nh][hd][Lp] fp16 (Q_hi,Q_lo,K_hi,K_lo,V_hi,V_lo), Lp % 8 == 0, Q pre-scaled by scale*log2e.
int launch_attention_h(const void* qkvh, float* ctx, const int64_t* lengths, int B, int L, int Lp, int nh, int hd,
                       cudaStream_t s, float* ctx_lo, void* ctx_half_planes, int32_t* status) {
  __half* ctx_h = reinterpret_cast<__half*>(ctx_half_planes);
  M2_REQUIRE(qkvh && (ctx || ctx_half_planes), M2TTS_E_NULLPTR, "attention_h: null pointer");
  M2_REQUIRE(attention_tc_supported(hd), M2TTS_E_UNSUPPORTED, "attention_h: head_dim %d unsupported", hd);
  M2_REQUIRE(B > 0 && L > 0 && nh > 0 && B <= 65535 && nh <= 65535 && (Lp & 7) == 0 && Lp >= L, M2TTS_E_BADSHAPE,
             "attention_h: B=%d L=%d Lp=%d nh=%d", B, L, Lp, nh);
  M2_REQUIRE((((uintptr_t)qkvh) & 15) == 0 && ((nh * hd) & 3) == 0, M2TTS_E_BADSHAPE, "attention_h: misaligned operands");
  EncodeTiledFnH enc = ah_encode_fn();
  M2_REQUIRE(enc != nullptr, M2TTS_E_CUDA, "attention_h: cuTensorMapEncodeTiled unavailable");
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)L, (cuuint64_t)6 * B * nh * hd};
  const cuuint64_t strides[1] = {(cuuint64_t)Lp * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)hd};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(qkvh), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "attention_h: cuTensorMapEncodeTiled failed (%d)", (int)r);
  switch (hd) {
    case 16: return launch_ah_hd<16>(tmap, ctx, lengths, B, L, nh, s, ctx_lo, ctx_h, status);
    case 32: return launch_ah_hd<32>(tmap, ctx, lengths, B, L, nh, s, ctx_lo, ctx_h, status);
    case 48: return launch_ah_hd<48>(tmap, ctx, lengths, B, L, nh, s, ctx_lo, ctx_h, status);
    default: return launch_ah_hd<64>(tmap, ctx, lengths, B, L, nh, s, ctx_lo, ctx_h, status);
  }
}

}  // namespace m2
