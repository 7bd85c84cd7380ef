// attention_h.cu — flash attention on tcgen05 with a 16-BIT split: every operand x is carried as two fp16 numbers,
// x = hi + lo (hi = fp16(x), lo = fp16(x - hi): 22 significant bits, the same as the TF32 split of attention_tc.cu),
// and a product is a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with fp32 accumulation in TMEM. Against the TF32 split this
// halves the operand bytes (K = 16 per UMMA instead of 8) and therefore the UMMA count. Accuracy on the reference's value
// ranges is that of the TF32 split (parity tests <= 1e-4); fp16 needs |x| <= 65504: the producers (QKV GEMM epilogue, this
// kernel's own output planes) check their values and raise M2TTS_ST_FP16_RANGE (include/m2tts_b200.h).
//
// A work item = two 128-query tiles of one (utterance, head) sharing every 64-key K/V tile (the odd last tile of an utterance
// is an item of its own); the kernel is PERSISTENT, one CTA per SM walking its list of items (attention_hp_kernel). Warp 0 TMA
// loader, warps 1 / 2 the UMMA issuers of query-tile slot A / B (warp-collective issue out of uniform registers: the warp index
// is a shuffle broadcast, so the role branches are uniform branches for ptxas), warps 4-19 four softmax warpgroups, two per
// slot; lazy rescaling with O accumulating in TMEM. The SCORES ARE DOUBLE-BUFFERED in TMEM: QK(t+2) is issued right behind PV(t).
//   * softmax warpgroup w of a query tile owns the key tiles t = w (mod 2) and score buffer w; a thread = one query row over
//     all 64 keys of the tile (the first version split a tile's columns between the two warpgroups, exchanged row maxima
//     through shared memory at a named barrier and made the two query tiles take turns in the exponential phase: 57 % of its
//     executed instructions were spin loops, barrier traffic and moves — profiles/README.md);
//   * the per-row reference maximum m_ref lives in shared memory and its per-tile decision is handed from warpgroup to
//     warpgroup through a one-directional named barrier (details at the kernel);
//   * 9 instructions per PAIR of scores: packed fp32 pairs (FADD2) for s - m, the row sum and p - p_hi; p_hi by masking the
//     low 13 mantissa bits (exact in fp16); one F2FP per packed pair; FMNMX3 for the maximum;
//   * P is written over S IN PLACE per 16-key group (= one k-step of P V): 8 columns of packed P_hi, 8 of packed P_lo;
//   * Q_hi lives in TMEM as the A operand of Q K^T (head_dim <= 48): six of the nine Q K^T UMMAs per key tile fetch only K
//     from shared memory (TMEM per query tile: 2 x 64 scores, O 2 hd, Q_hi hd / 2 = 248 of 256 columns at head_dim 48); it is
//     copied there by the softmax warps from the TMA-loaded shared-memory boxes. head_dim 64: all three terms from shared memory.
// Measured and rejected: 128-key tiles with one in-place score buffer (serialises Q K^T -> softmax -> P V, 0.99 ms against
// 0.91), a Veltkamp split on the FMA pipe, part of the exponentials as an FFMA2 polynomial (slower at every fraction), three score
// buffers with O in hd columns (one more UMMA per P V k-step).
// Operands: qkvh[6][B][nh][hd][Lp] fp16 = {Q_hi,Q_lo,K_hi,K_lo,V_hi,V_lo}, positions contiguous, Lp % 8 == 0, Q
// pre-multiplied by scale*log2(e). A TMA box = 64 positions (128 B) x hd rows: for Q and K the MN-major 128B-swizzle
// operand (K dim = d), for V the K-major 128B-swizzle operand (rows = d, K dim = keys) — layouts verified with
// tools/umma_probe_f16.cu. P is written back to TMEM as packed fp16 pairs (low half = even key). Reference: components.py:75-87.
#include "attention_tc.cuh"
#include <cuda_fp16.h>

namespace m2 {

constexpr int AH_THREADS = 128 + 512;      // loader, two issuers, idle | four softmax warpgroups (two per query tile)
constexpr uint32_t AH_TMEM_COLS = 512;
// TMEM per query tile (tile stride 256 columns): two score buffers of 64 columns (tile t uses buffer t & 1; P(t) is
// packed over it, 16-key group u in columns [16 u, 16 u + 16): P_hi in the first 8, P_lo in the last 8), O in 2 hd columns
// from 128, Q_hi in hd / 2 columns from 224.
constexpr uint32_t AH_COL_S = 0, AH_COL_O = 128, AH_COL_Q = 224;      // Q_hi: hd / 2 columns from 224 (head_dim <= 48)
constexpr uint32_t AH_COL_TILE = 256;
constexpr int AH_STAGES = 4;        // K / V ring depth
// phase timestamps (tools/attn_prof.py) exist only in the tools build: predicated-off CS2R / STG pairs still cost issue slots
#ifdef M2TTS_TOOLS
#define AH_PROF(cond, stmt) do { if (cond) { stmt; } } while (0)
#else
#define AH_PROF(cond, stmt) do { } while (0)
#endif

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

__host__ __device__ constexpr uint32_t ah_idesc2(int M, int N, int a_mn, int b_mn) {   // kind::f16: fp16 x fp16 -> fp32; *_mn = 1: MN-major smem operand
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t ah_idesc(int M, int N, int mn_major) { return ah_idesc2(M, N, mn_major, mn_major); }
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

// packed fp32 pairs (Blackwell FADD2 / FFMA2): one instruction for two lanes
__device__ __forceinline__ uint64_t ah_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void ah_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ah_add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t ah_sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ah_max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint32_t ah_cvt2(float lo, float hi) {      // {lo, hi} -> packed fp16 pair (low half = lo), round to nearest
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void ah_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void ah_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void ah_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ uint32_t ah_lds16(uint32_t saddr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float ws_ex2v(float x) {      // ex2 that keeps its place in the instruction stream
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Softmax organisation of the attention kernel below:
//   * warpgroup w of a query tile owns the key tiles t = w (mod 2) and their score buffer w, a thread = one query row over all
//     64 keys. Nothing is exchanged inside a tile, and while one warpgroup waits for the P V / Q K^T of its buffer the other
//     one is in the middle of its tile, so the four warps of a scheduler sit in different phases without being forced to;
//   * the per-row reference maximum m_ref (lazy rescaling) lives in shared memory; the DECISION for tile t (keep m_ref or raise
//     it and rescale O) is handed from the warpgroup of tile t-1 to the warpgroup of tile t through a one-directional named
//     barrier (bar.arrive by the producer right after its decision, bar.sync by the consumer right before its own): the chain
//     per tile is load + maximum + decision (~300 cycles), the exponentials run outside it. Each warpgroup keeps its own
//     partial row sum in the scale of the m_ref it last saw and rescales it when it sees a new one;
//   * within a tile the exponentials of 16-key group g+1 are issued between the split arithmetic of group g (MUFU and ALU / FMA
//     work interleaved in one instruction stream); groups 2, 3 come from the registers of the maximum pass, groups 0, 1 are
//     re-read from TMEM (32 score registers live at a time);
//   * P(t)-ready has one mbarrier per warpgroup (alternating arrivals must not mix in one phase); the last P V also commits to
//     a single-phase barrier for the final read of O (a parity wait is only exact for a waiter at most one phase behind).
// Reference: components.py:75-87.
// ---------------------------------------------------------------------------------------------------------------------------
// attention_hp_kernel — persistent (round 2; until then one CTA per item, attention_h_kernel in the history). One CTA per SM walks a fixed
// list of work items (item = the two — at the end of an utterance one — 128-query tiles of one (utterance, head)); every role
// (loader, issuers, softmax warpgroups) runs its own loop over the same list and all mbarrier phases simply continue from item
// to item (g = key tiles this query-tile slot has seen so far decides score buffer, owner warpgroup and parity; gk = key tiles
// the CTA has seen decides the K / V ring stage). What one CTA per item paid at every item (tools/attn_cta_prof.py, cycles of a
// 100 k-cycle CTA): launch gap 1.8 k, barrier init + TMEM allocation 1.5 k, Q load + Q_hi copy 2.4 k, first scores 0.7 k, drain
// of the last P V 2.9 k, O normalise + store 2.1 k — here the next item's Q and K / V loads, its Q_hi copy and its first two
// Q K^T run under the previous item's last tiles and epilogue. Additional barriers: q_empty (last Q K^T of an item committed:
// the Q slot may be reloaded) and o_free (both warpgroups have read O: the next item's first P V may overwrite it).
// Item order per CTA: two-tile items c, c + G, ...; the single-tile items (half the duration) go first to the CTAs that got
// one two-tile item less, two each, then round-robin (AhItems) — the balance the hardware scheduler gave the one-CTA-per-item
// kernel with its longest-first grid.
struct AhItems {
  int n_long, n_single, G, c, r, nd, stage, m;
  __device__ __forceinline__ AhItems(int n_long_, int n_single_, int G_, int c_)
      : n_long(n_long_), n_single(n_single_), G(G_), c(c_), r(n_long_ % G_), nd(G_ - n_long_ % G_), stage(0), m(0) {}
  __device__ __forceinline__ int next() {      // -1 at the end
    if (stage == 0) {
      const int idx = c + m * G;
      if (idx < n_long) { ++m; return idx; }
      stage = 1; m = 0;
    }
    if (stage == 1) {
      const int lim = min(n_single, 2 * nd);
      const int j = (c - r) + m * nd;
      if (c >= r && m < 2 && j < lim) { ++m; return n_long + j; }
      stage = 2; m = 0;
    }
    const int j = 2 * nd + c + m * G;
    if (j < n_single) { ++m; return n_long + j; }
    return -1;
  }
};
struct AhItem { int q0, head, b, ntq, Leff, nkt, row_q; bool all_masked; };
template <int HD>
__device__ __forceinline__ AhItem ah_item(int idx, int n_long, int npf, int nh, int L, const int64_t* __restrict__ lengths) {
  AhItem w;
  int qx, bh;
  if (idx < n_long) { bh = idx / npf; qx = idx % npf; }
  else { bh = idx - n_long; qx = npf; }
  w.q0 = qx * (2 * TC_BQ); w.head = bh % nh; w.b = bh / nh;
  w.ntq = w.q0 + TC_BQ < L ? 2 : 1;
  w.Leff = L; w.all_masked = false;
  if (lengths != nullptr) {
    const long long len = lengths[w.b];
    if (len <= 0) w.all_masked = true;
    else if (len < L) w.Leff = (int)len;
  }
  w.nkt = (w.Leff + TC_BK - 1) / TC_BK;
  w.row_q = (w.b * nh + w.head) * HD;
  return w;
}

template <int HD>
__global__ void __launch_bounds__(AH_THREADS, 1)
attention_hp_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ ctx, const int64_t* __restrict__ lengths, int B, int L, int nh,
                    float* __restrict__ ctx_lo, __half* __restrict__ ctx_h, long long* __restrict__ prof, int32_t* __restrict__ status, int dbg_skip) {
  static_assert(HD % 16 == 0 && HD >= 16 && HD <= 64, "16-bit split attention: head_dim in {16,32,48,64}");
  constexpr bool QT = (128 + 2 * HD + HD / 2) <= 256;             // Q_hi as a TMEM A operand (head_dim <= 48), else from shared memory
  pdl_launch_dependents();      // M2_LAUNCH_PDL: every global access below follows a pdl_wait()
#ifdef M2TTS_TOOLS
  long long cta_t0 = 0, cta_c0 = 0;
  if (prof != nullptr && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(cta_t0)); cta_c0 = clock64(); }
#endif
  constexpr uint32_t BOX = AhSmem<HD>::box;
  constexpr int KSTEPS_D = HD / 16;
  constexpr uint32_t IDESC_QK1 = ah_idesc(TC_BQ, TC_BK, 1);
  constexpr uint32_t IDESC_QKT = ah_idesc2(TC_BQ, TC_BK, 0, 1);
  constexpr uint32_t IDESC_PV2 = ah_idesc(TC_BQ, 2 * HD, 0);
  constexpr uint32_t IDESC_PV1 = ah_idesc(TC_BQ, HD, 0);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = sbase;
  const uint32_t sK = sbase + AhSmem<HD>::off_k;
  const uint32_t sV = sbase + AhSmem<HD>::off_v;
  const uint32_t sBar = sbase + AhSmem<HD>::off_bar;
  // barriers: q_full[2] s_full[2 tiles][2 buffers] p_ready[2 tiles][2 warpgroups] pv_done[2] o_done[2] | k_full[S] k_empty[S]
  // v_full[S] v_empty[S] | tmem slot | q_hi[2] | p_ready of the second half [2][2] | q_empty[2] o_free[2]
  const uint32_t bar_qf = sBar, bar_sf = sBar + 16, bar_pr = sBar + 48, bar_pv = sBar + 80, bar_done = sBar + 96;
  const uint32_t bar_kf = sBar + 112, bar_ke = bar_kf + 8 * AH_STAGES, bar_vf = bar_ke + 8 * AH_STAGES, bar_ve = bar_vf + 8 * AH_STAGES;
  const uint32_t tmem_slot = bar_ve + 8 * AH_STAGES;
  const uint32_t bar_qh = tmem_slot + 8;
  const uint32_t bar_p2 = bar_qh + 16;
  const uint32_t bar_qe = bar_p2 + 32, bar_of = bar_qe + 16;

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int n_tiles_q = (L + TC_BQ - 1) / TC_BQ, npf = n_tiles_q >> 1, n_long = npf * nh * B, n_single = (n_tiles_q & 1) * nh * B;
  const int plane = B * nh * HD;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_qf + 8 * i, 1); mbar_init(bar_sf + 16 * i, 1); mbar_init(bar_sf + 16 * i + 8, 1);
      mbar_init(bar_pr + 16 * i, 4); mbar_init(bar_pr + 16 * i + 8, 4); mbar_init(bar_pv + 8 * i, 1); mbar_init(bar_done + 8 * i, 1);
      mbar_init(bar_qh + 8 * i, 8); mbar_init(bar_p2 + 16 * i, 4); mbar_init(bar_p2 + 16 * i + 8, 4);
      mbar_init(bar_qe + 8 * i, 1); mbar_init(bar_of + 8 * i, 8);
    }
    for (int i = 0; i < AH_STAGES; ++i) {      // K / V stages are released by two commits: one per issuer, or two by issuer A for a single-tile item
      mbar_init(bar_kf + 8 * i, 1); mbar_init(bar_ke + 8 * i, 2); mbar_init(bar_vf + 8 * i, 1); mbar_init(bar_ve + 8 * i, 2);
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
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  AhItems items(n_long, n_single, (int)gridDim.x, (int)blockIdx.x);
  if (warp == 0) {
    if (lane == 0) {
      // ===== loader =====
      pdl_wait();
      int gk = 0, nq[2] = {0, 0};
      for (int idx = items.next(); idx >= 0; idx = items.next()) {
        const AhItem w = ah_item<HD>(idx, n_long, npf, nh, L, lengths);
        for (int x = 0; x < w.ntq; ++x) {
          if (nq[x] > 0) mbar_wait(bar_qe + 8 * x, (uint32_t)((nq[x] - 1) & 1));      // the slot's last Q K^T of its previous item has run
          mbar_expect_tx(bar_qf + 8 * x, AhSmem<HD>::q_bytes);
          for (int h = 0; h < 2; ++h)
            for (int j = 0; j < 2; ++j)
              tma_load_2d(sQ + (uint32_t)x * AhSmem<HD>::q_bytes + (h * 2 + j) * BOX, &tmap, w.q0 + x * TC_BQ + j * 64, h * plane + w.row_q,
                          bar_qf + 8 * x);
          ++nq[x];
        }
        for (int t = 0; t < w.nkt; ++t) {
          const int g = gk + t, st = g % AH_STAGES;
          const uint32_t par_prev = (uint32_t)(((g / AH_STAGES) - 1) & 1);
          if (g >= AH_STAGES) mbar_wait(bar_ke + 8 * st, par_prev);
          mbar_expect_tx(bar_kf + 8 * st, AhSmem<HD>::kv_bytes);
          for (int h = 0; h < 2; ++h)
            tma_load_2d(sK + (uint32_t)st * AhSmem<HD>::kv_bytes + h * BOX, &tmap, t * TC_BK, (2 + h) * plane + w.row_q, bar_kf + 8 * st);
          if (g >= AH_STAGES) mbar_wait(bar_ve + 8 * st, par_prev);
          mbar_expect_tx(bar_vf + 8 * st, AhSmem<HD>::kv_bytes);
          for (int h = 0; h < 2; ++h)
            tma_load_2d(sV + (uint32_t)st * AhSmem<HD>::kv_bytes + h * BOX, &tmap, t * TC_BK, (4 + h) * plane + w.row_q, bar_vf + 8 * st);
        }
        gk += w.nkt;
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===== UMMA issuer of query-tile slot x (whole warp, one elected lane issues) =====
    const int x = warp - 1;
    auto issue_qk = [&](int st, int buf) {
      if (dbg_skip & 1) return;      // bring-up timing experiment (tools build, M2TTS_ATT_DBG): results invalid
      const uint32_t q = sQ + (uint32_t)x * AhSmem<HD>::q_bytes, k = sK + (uint32_t)st * AhSmem<HD>::kv_bytes;
      const uint32_t d = tmem_base + (uint32_t)x * AH_COL_TILE + AH_COL_S + (uint32_t)buf * 64u;
      if (QT) {
        const uint32_t qt = tmem_base + (uint32_t)x * AH_COL_TILE + AH_COL_Q;
#pragma unroll
        for (int ks = 0; ks < KSTEPS_D; ++ks) {
          const uint64_t khi = umma_desc(k + ks * 2048u, BOX, 1024u, 2u), klo = umma_desc(k + BOX + ks * 2048u, BOX, 1024u, 2u);
          umma_f16_ts_w(d, qt + ks * 8, khi, IDESC_QKT, ks ? 1u : 0u);
          umma_f16_ts_w(d, qt + ks * 8, klo, IDESC_QKT, 1u);
          umma_f16_ss_w(d, umma_desc(q + 2 * BOX + ks * 2048u, BOX, 1024u, 2u), khi, IDESC_QK1, 1u);
        }
        return;
      }
      // head_dim 64: O takes all the columns Q_hi would need; the three product terms from shared memory (hi*hi, hi*lo, lo*hi)
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t qa = q + (term == 2 ? 2 * BOX : 0u), kb = k + (term == 1 ? BOX : 0u);
#pragma unroll
        for (int ks = 0; ks < KSTEPS_D; ++ks)
          umma_f16_ss_w(d, umma_desc(qa + ks * 2048u, BOX, 1024u, 2u), umma_desc(kb + ks * 2048u, BOX, 1024u, 2u), IDESC_QK1,
                        (term | ks) ? 1u : 0u);
      }
    };
    auto issue_pv = [&](int st, int buf, int half, uint32_t accumulate) {
      if (dbg_skip & 2) return;
      const uint32_t v = sV + (uint32_t)st * AhSmem<HD>::kv_bytes;
      const uint32_t tb = tmem_base + (uint32_t)x * AH_COL_TILE;
      const uint32_t pb = tb + AH_COL_S + (uint32_t)buf * 64u;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int ks = (half == 0 ? 2 : 0) + kk;
        umma_f16_ts_w(tb + AH_COL_O, pb + ks * 16, umma_desc(v + ks * 32u, 16u, 1024u, 2u), IDESC_PV2, (half | kk) ? 1u : accumulate);
        umma_f16_ts_w(tb + AH_COL_O, pb + ks * 16 + 8, umma_desc(v + ks * 32u, 16u, 1024u, 2u), IDESC_PV1, 1u);
      }
    };
    int gk = 0, g = 0, nq = 0;
    for (int idx = items.next(); idx >= 0; idx = items.next()) {
      const AhItem w = ah_item<HD>(idx, n_long, npf, nh, L, lengths);
      if (x < w.ntq) {
        const bool twice = w.ntq == 1;      // no second issuer for this item: release the K / V stages for both
        for (int t = 0; t < 2 && t < w.nkt; ++t) {      // prologue: the scores of key tiles 0 and 1
          const int st = (gk + t) % AH_STAGES;
          mbar_wait(bar_kf + 8 * st, (uint32_t)(((gk + t) / AH_STAGES) & 1));
          if (t == 0) {
            mbar_wait(bar_qf + 8 * x, (uint32_t)(nq & 1));
            if (QT) mbar_wait(bar_qh + 8 * x, (uint32_t)(nq & 1));
          }
          tc_fence_after();
          issue_qk(st, (g + t) & 1);
          tc_commit_w(bar_sf + 16 * x + 8 * ((g + t) & 1));
          tc_commit_w(bar_ke + 8 * st);
          if (twice) tc_commit_w(bar_ke + 8 * st);
          if (t == w.nkt - 1) tc_commit_w(bar_qe + 8 * x);
        }
        for (int t = 0; t < w.nkt; ++t) {
          const int gt = g + t, buf = gt & 1, st = (gk + t) % AH_STAGES;
          const uint32_t ph = (uint32_t)((gt >> 1) & 1);
          mbar_wait(bar_pr + 16 * x + 8 * buf, ph);
          mbar_wait(bar_vf + 8 * st, (uint32_t)(((gk + t) / AH_STAGES) & 1));
          if (t == 0 && nq > 0) mbar_wait(bar_of + 8 * x, (uint32_t)((nq - 1) & 1));      // the previous item's O has been read
          tc_fence_after();
          issue_pv(st, buf, 0, t > 0 ? 1u : 0u);
          mbar_wait(bar_p2 + 16 * x + 8 * buf, ph);
          tc_fence_after();
          issue_pv(st, buf, 1, 1u);
          tc_commit_w(bar_pv + 8 * x);
          if (t == w.nkt - 1) tc_commit_w(bar_done + 8 * x);
          tc_commit_w(bar_ve + 8 * st);
          if (twice) tc_commit_w(bar_ve + 8 * st);
          if (t + 2 < w.nkt) {                                      // score buffer `buf` is free again once PV(t) has read P
            const int s2 = (gk + t + 2) % AH_STAGES;
            mbar_wait(bar_kf + 8 * s2, (uint32_t)(((gk + t + 2) / AH_STAGES) & 1));
            tc_fence_after();
            issue_qk(s2, buf);
            tc_commit_w(bar_sf + 16 * x + 8 * buf);
            tc_commit_w(bar_ke + 8 * s2);
            if (twice) tc_commit_w(bar_ke + 8 * s2);
            if (t + 2 == w.nkt - 1) tc_commit_w(bar_qe + 8 * x);
          }
        }
        g += w.nkt;
        ++nq;
      }
      gk += w.nkt;
    }
  } else if (warp >= 4) {
    // ===== softmax warpgroups: slot x has two (warps 4-7 / 12-15 for slot A, 8-11 / 16-19 for slot B); warpgroup wg owns the key
    // tiles with (g + t) & 1 == wg and score buffer wg; thread = query row = TMEM lane, all 64 keys of the tile =====
    const int x = ((warp - 4) >> 2) & 1, wg = (warp - 4) >> 3;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)x * AH_COL_TILE;
    const uint32_t t_s = t_lane + AH_COL_S + (uint32_t)wg * 64u;      // score buffer of this warpgroup
    float* exch = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)) + AhSmem<HD>::off_exch);
    float* mref_s = exch + x * 128;               // per-row reference maximum of the running softmax
    float* lsum_s = exch + 256 + x * 128;         // final row sums: warpgroup wg writes [wg * 256 + row]
    const int hb = 1 + x * 2;                     // named barriers hb / hb + 1: hand-off of the m_ref decision of even / odd key tiles
    pdl_wait();                                   // ctx is written below
    // Q_hi of an item -> TMEM (the A operand of Q K^T) from the TMA-loaded boxes in shared memory; this warpgroup writes the d range
    // [wg hd/2, (wg+1) hd/2). All Q K^T of the slot's previous item have completed when this runs (each warpgroup has seen the
    // scores of its last key tile, and the hand-off of the last decision orders the other warpgroup behind them).
    auto copy_q = [&](const AhItem& w, int nq) {
      const int qi = w.q0 + x * TC_BQ + row;
      const uint32_t qs = sQ + (uint32_t)x * AhSmem<HD>::q_bytes + (uint32_t)(row >> 6) * BOX;
      const uint32_t p2 = (uint32_t)(row & 63) * 2u;
      mbar_wait(bar_qf + 8 * x, (uint32_t)(nq & 1));
#pragma unroll
      for (int c4 = 0; c4 < HD / 16; ++c4) {
        uint32_t wv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t d = (uint32_t)(wg * (HD / 2) + (c4 * 4 + e) * 2);
          const uint32_t lo = ah_lds16(qs + d * 128u + (p2 ^ ((d & 7u) << 4)));
          const uint32_t hi = ah_lds16(qs + (d + 1u) * 128u + (p2 ^ (((d + 1u) & 7u) << 4)));
          wv[e] = qi < L ? (lo | (hi << 16)) : 0u;
        }
        ah_st4(t_lane + AH_COL_Q + (uint32_t)(wg * (HD / 4) + c4 * 4), wv);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_qh + 8 * x) : "memory");
    };
    auto next_item = [&](AhItem& w) {             // the next item that has a query tile in this slot
      for (int idx = items.next(); idx >= 0; idx = items.next()) {
        w = ah_item<HD>(idx, n_long, npf, nh, L, lengths);
        if (x < w.ntq) return true;
      }
      return false;
    };
    AhItem w;
    bool have = next_item(w);
    int g = 0, nq = 0;
    if (QT && have) copy_q(w, 0);
    while (have) {
      const int Leff = w.Leff, nkt = w.nkt;
      const bool all_masked = w.all_masked;
      float m_c = -INFINITY;                                // the m_ref this warpgroup's row sum is scaled to
      uint64_t l2 = ah_pack(0.f, 0.f);                      // running row sum as a packed pair (even keys, odd keys)
#ifdef M2TTS_TOOLS
      const bool pw = prof != nullptr && blockIdx.x == 0 && x == 0 && wg == 0 && row == 0 && nq == 0;
#endif
      auto mask16 = [&](uint32_t* sv, int k0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = __uint_as_float(sv[j]);
          if (all_masked) a = (k0 + j < L) ? 0.f : -INFINITY;
          else if (k0 + j >= Leff) a = -INFINITY;
          sv[j] = __float_as_uint(a);
        }
      };
      auto max32 = [&](const uint32_t* u, const uint32_t* v) {
        float mx = ah_max3(__uint_as_float(u[0]), __uint_as_float(u[1]), __uint_as_float(u[2]));
#pragma unroll
        for (int j = 3; j < 15; j += 2) mx = ah_max3(mx, __uint_as_float(u[j]), __uint_as_float(u[j + 1]));
        mx = ah_max3(mx, __uint_as_float(u[15]), __uint_as_float(v[0]));
#pragma unroll
        for (int j = 1; j < 15; j += 2) mx = ah_max3(mx, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
        return fmaxf(mx, __uint_as_float(v[15]));
      };
      auto rescale_l = [&](float m_new) {                   // the row sum follows m_ref (2^-inf = 0 on the first tile: l2 is 0 then)
        const float a = ws_ex2(m_c - m_new);
        float la, lb;
        ah_unpack(l2, la, lb);
        l2 = ah_pack(la * a, lb * a);
        m_c = m_new;
      };
      for (int t = (wg - g) & 1; t < nkt; t += 2) {
        const int gt = g + t;                               // (gt & 1) == wg
#ifdef M2TTS_TOOLS
        const bool pt = pw && t >= 8 && t < 72;
        long long* pp = prof + (pt ? ((t - 8) >> 1) * 8 : 0);
#endif
        AH_PROF(pt, pp[0] = clock64());
        mbar_wait(bar_sf + 16 * x + 8 * wg, (uint32_t)((gt >> 1) & 1));
        AH_PROF(pt, pp[1] = clock64());
        __syncwarp();
        tc_fence_after();
        const int kbase = t * TC_BK;
        const bool need_mask = all_masked || kbase + TC_BK > Leff;      // key padding: only ever in the last key tile
        // pass 1: row maximum over the 64 keys; the second half of the scores stays in registers
        uint32_t sa[16], sb[16];
        tmem_ld16(t_s, sa);
        tmem_ld16(t_s + 16, sb);
        tmem_wait_ld();
        if (need_mask) { mask16(sa, kbase); mask16(sb, kbase + 16); }
        float mx = max32(sa, sb);
        tmem_ld16(t_s + 32, sa);
        tmem_ld16(t_s + 48, sb);
        tmem_wait_ld();
        if (need_mask) { mask16(sa, kbase + 32); mask16(sb, kbase + 48); }
        mx = fmaxf(mx, max32(sa, sb));
        AH_PROF(pt, pp[2] = clock64());
        // the decision of tile t-1 (other warpgroup) precedes ours
        if (t > 0) asm volatile("bar.sync %0, 256;" ::"r"(hb + ((gt - 1) & 1)) : "memory");
        float m_s;
        if (t == 0) {
          m_s = mx;
          mref_s[row] = mx;
        } else {
          m_s = mref_s[row];
          if (__any_sync(0xffffffffu, mx > m_s + 8.0f)) {
            // lazy rescale: P V(gt-1) has landed after this wait (the barrier has completed gt-1 or gt phases here, so the parity
            // test is exact), P V(gt) waits for our P and P V(gt+1) for the other warpgroup, which waits for our hand-off: O is quiescent
            mbar_wait(bar_pv + 8 * x, (uint32_t)((gt - 1) & 1));
            tc_fence_after();
            const float m_new = fmaxf(m_s, mx);
            const float alpha = ws_ex2(m_s - m_new);
#pragma unroll
            for (int c = 0; c < 2 * HD; c += 16) {
              uint32_t orr[16];
              tmem_ld16(t_lane + AH_COL_O + c, orr);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j) orr[j] = __float_as_uint(__uint_as_float(orr[j]) * alpha);
              tmem_st16(t_lane + AH_COL_O + c, orr);
            }
            mref_s[row] = m_new;
            m_s = m_new;
          }
        }
        asm volatile("bar.arrive %0, 256;" ::"r"(hb + (gt & 1)) : "memory");      // hand the decision on to the next tile
        if (m_s != m_c) rescale_l(m_s);
        AH_PROF(pt, pp[3] = clock64());
        // pass 2: p = 2^(s - m_ref) per 16-key group (= one k-step of P V), written over the group's own score columns as 8 columns
        // of packed P_hi (p with the low 13 mantissa bits masked off: exact in fp16) and 8 of packed P_lo = fp16(p - P_hi).
        const uint64_t m2 = ah_pack(m_s, m_s);
        auto exp_pair = [&](uint32_t* s, int j) {
          float d0, d1;
          ah_unpack(ah_sub2(ah_pack(__uint_as_float(s[2 * j]), __uint_as_float(s[2 * j + 1])), m2), d0, d1);
          s[2 * j] = __float_as_uint(ws_ex2v(d0)); s[2 * j + 1] = __float_as_uint(ws_ex2v(d1));
        };
        auto split_pair = [&](const uint32_t* pv, int j, uint32_t* ph, uint32_t* pl) {
          const float p0 = __uint_as_float(pv[2 * j]), p1 = __uint_as_float(pv[2 * j + 1]);
          const uint64_t pp2 = ah_pack(p0, p1);
          l2 = ah_add2(l2, pp2);
          const float h0 = __uint_as_float(__float_as_uint(p0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(p1) & 0xFFFFE000u);
          float r0, r1;
          ah_unpack(ah_sub2(pp2, ah_pack(h0, h1)), r0, r1);
          ph[j] = ah_cvt2(h0, h1);
          pl[j] = ah_cvt2(r0, r1);
        };
        auto exp_and_split = [&](uint32_t* nx, const uint32_t* cur, uint32_t col) {
          uint32_t ph[8], pl[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (nx != nullptr) exp_pair(nx, j);
            if (cur != nullptr) split_pair(cur, j, ph, pl);
          }
          if (cur != nullptr) {
            ah_st8(t_s + col, ph);
            ah_st8(t_s + col + 8, pl);
          }
        };
        exp_and_split(sa, nullptr, 0);            // group 2
        exp_and_split(sb, sa, 32);                // group 3 | group 2
        uint32_t sc[16], sd[16];
        tmem_ld16(t_s, sc);
        tmem_ld16(t_s + 16, sd);
        tmem_wait_ld();
        if (need_mask) { mask16(sc, kbase); mask16(sd, kbase + 16); }
        exp_and_split(sc, sb, 48);                // group 0 | group 3
        // groups 2 and 3 of P are stored: the issuer may start their P V k-steps while groups 0 and 1 are still being computed
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_pr + 16 * x + 8 * wg) : "memory");
        exp_and_split(sd, sc, 0);                 // group 1 | group 0
        exp_and_split(nullptr, sd, 16);           //         | group 1
        AH_PROF(pt, pp[4] = clock64());
        tmem_wait_st();
        AH_PROF(pt, pp[5] = clock64());
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_p2 + 16 * x + 8 * wg) : "memory");
      }
      // the warpgroup that did not own the last key tile takes over its decision (the final m_ref)
      if (((g + nkt - 1) & 1) != wg) {
        asm volatile("bar.sync %0, 256;" ::"r"(hb + ((g + nkt - 1) & 1)) : "memory");
        const float m_s = mref_s[row];
        if (m_s != m_c) rescale_l(m_s);
      }
      // final row sum: each warpgroup holds the sum of its own key tiles (both in the scale of the final m_ref)
      float l_run;
      {
        float la, lb;
        ah_unpack(l2, la, lb);
        l_run = la + lb;
        lsum_s[wg * 256 + row] = l_run;
        asm volatile("bar.sync %0, 256;" ::"r"(5 + x) : "memory");
        l_run += lsum_s[(wg ^ 1) * 256 + row];
      }
      // the slot's next item: its Q_hi goes to TMEM now, so its first two Q K^T run under the epilogue below
      const AhItem cur = w;
      have = next_item(w);
      if (QT && have) copy_q(w, nq + 1);
      {
        // both warpgroups normalise and store: warpgroup wg takes the columns [wg hd/2, (wg+1) hd/2) of every row
        mbar_wait(bar_done + 8 * x, (uint32_t)(nq & 1));     // the item's last PV has landed: O is complete
        __syncwarp();
        tc_fence_after();
        const int qi = cur.q0 + x * TC_BQ + row;
        const float inv = 1.0f / l_run;
        const long long orow = ((long long)cur.b * L + qi) * (nh * HD) + cur.head * HD;
        bool bad = false;
#pragma unroll
        for (int c0 = 0; c0 < HD / 2; c0 += 8) {
          const int c = wg * (HD / 2) + c0;
          uint32_t orr[8], or2[8];
          ah_ld8(t_lane + AH_COL_O + c, orr);
          ah_ld8(t_lane + AH_COL_O + HD + c, or2);
          tmem_wait_ld();
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (__uint_as_float(orr[j]) + __uint_as_float(or2[j])) * inv;
          if (qi < L) {
            if (ctx_h != nullptr) {   // fp16 hi/lo planes [2][B*L][nh*HD] for the 16-bit split out_proj (lin_h.cu)
              __half* dh = ctx_h + orow + c;
              __half* dl = dh + (long long)B * L * (nh * HD);
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) h_split2(o[2 * e], o[2 * e + 1], hi[e], lo[e], bad);
              *reinterpret_cast<uint4*>(dh) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(dl) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            } else if (ctx_lo == nullptr) {
              *reinterpret_cast<float4*>(ctx + orow + c) = make_float4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<float4*>(ctx + orow + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
            } else {   // TF32 hi/lo planes for the tensor-core out_proj
              float h[8], l[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) { h[e] = __uint_as_float(tf32_hi(o[e])); l[e] = __uint_as_float(tf32_hi(o[e] - h[e])); }
              *reinterpret_cast<float4*>(ctx + orow + c) = make_float4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<float4*>(ctx + orow + c + 4) = make_float4(h[4], h[5], h[6], h[7]);
              *reinterpret_cast<float4*>(ctx_lo + orow + c) = make_float4(l[0], l[1], l[2], l[3]);
              *reinterpret_cast<float4*>(ctx_lo + orow + c + 4) = make_float4(l[4], l[5], l[6], l[7]);
            }
          }
        }
        // O has been read: the next item's first P V may overwrite it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_of + 8 * x) : "memory");
        if (ctx_h != nullptr) h_flag(bad, status);
      }
      g += nkt;
      ++nq;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(AH_TMEM_COLS) : "memory");
  }
#ifdef M2TTS_TOOLS
  if (prof != nullptr && threadIdx.x == 0) {      // CTA lifetime (tools/attn_cta_prof.py): [start ns, end ns, SM, start clock, end clock] from word 1024 on
    long long t1; uint32_t sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    long long* wp = prof + 1024 + 5 * (long long)blockIdx.x;
    wp[0] = cta_t0; wp[1] = t1; wp[2] = sm; wp[3] = cta_c0; wp[4] = clock64();
  }
#endif
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
  const int items = ceil_div(L, 2 * TC_BQ) * nh * B;      // persistent: one CTA per SM over the item list
  dim3 grid((unsigned)(items < kNumSMs ? items : kNumSMs), 1, 1);
  M2_CUDA_OK(allow_smem(attention_hp_kernel<HD>, smem));
  M2_LAUNCH_PDL(M2TTS_STAGE_ATTENTION, attention_hp_kernel<HD>, grid, AH_THREADS, smem, s, tmap, ctx, lengths, B, L, nh, ctx_lo, ctx_h,
                tools_env_int("M2TTS_LIN_PROF_STAGE", -1) >= 0 ? nullptr : g_ws_prof, status, dbg_skip);      // the buffer belongs to tools/lin_prof.py then
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
