// attention_tc.cu — flash attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA),
// fp32-faithful through 3xTF32 splitting: every operand x is carried as x = hi + lo with both
// halves exactly representable in TF32 (10-bit mantissa), and each product is evaluated as
//   a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi        (dropped term a_lo*b_lo ~ 2^-22 |a b|)
// with fp32 accumulation in TMEM. This keeps the reference's fp32 results (components.py:75-87)
// within ~1e-6 while the O(L^2) work runs on the tensor pipe instead of FFMA.
//
// Operands (written by the QKV row-GEMM epilogue, qkv_mode 2): one fp32 array
//   qkv6[6][B][nh][hd][Lp]  = {Q_hi, Q_lo, K_hi, K_lo, V_hi, V_lo}, positions contiguous,
// Q already multiplied by scale*log2(e). One TMA tensor map over it (rows = 6*B*nh*hd,
// cols = L, row pitch Lp) with a {32 positions x hd rows} box and 128-byte swizzle, so a box is
// hd rows of 128 B: for Q and K this is the canonical MN-major SW128 UMMA layout (K dim = d),
// for V^T it is the canonical K-major SW128 layout (N dim = d, K dim = keys).
//
// CTA = 128 threads = one (utterance, head, 128-query tile); thread i owns query row i (TMEM
// lane i). Per 64-key tile:  S = Q K^T (18 UMMAs, M128 N64 K8) -> tcgen05.ld -> online softmax
// in registers -> P_hi/P_lo back to TMEM -> O_tile = P V (24 UMMAs, M128 N=hd K8, A from TMEM)
// -> tcgen05.ld -> o = o*alpha + O_tile in registers. K and V tiles are single-buffered but
// refilled by TMA as soon as the UMMAs that read them have committed; two CTAs per SM overlap
// each other's tensor and ALU phases. TMEM: 256 columns per CTA (S 64 | P_hi 64 | P_lo 64 | O 64).
#include "attention_tc.cuh"

namespace m2 {

constexpr int TC_BOX = 32;      // positions per TMA box (128 B of fp32)
constexpr int TC_THREADS = 128;
constexpr uint32_t TC_TMEM_COLS = 256;
constexpr uint32_t TC_COL_S = 0, TC_COL_PHI = 64, TC_COL_PLO = 128, TC_COL_O = 192;

template <int HD>
struct TcSmem {
  static constexpr uint32_t box_bytes = HD * 128;                   // hd rows x 32 positions x 4 B
  static constexpr uint32_t q_bytes = 2 * (TC_BQ / TC_BOX) * box_bytes;   // hi + lo
  static constexpr uint32_t kv_bytes = 2 * (TC_BK / TC_BOX) * box_bytes;  // hi + lo, per operand
  static constexpr uint32_t total = q_bytes + 2 * kv_bytes + 1024 /*align slack*/ + 64 /*barriers*/;
};

template <int HD>
__global__ void __launch_bounds__(TC_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qk, const __grid_constant__ CUtensorMap tmap_v,
                    float* __restrict__ ctx,
                    const int64_t* __restrict__ lengths, int B, int L, int nh,
                    float* __restrict__ dbg_s, float* __restrict__ dbg_o, float* __restrict__ ctx_lo) {
  static_assert(HD % 16 == 0 && HD >= 16 && HD <= 64, "tensor-core path: head_dim in {16,32,48,64}");
  constexpr uint32_t BOX = TcSmem<HD>::box_bytes;
  constexpr int QBOX = TC_BQ / TC_BOX, KBOX = TC_BK / TC_BOX;  // 4, 2
  constexpr int KSTEPS_D = HD / 8;                             // k-steps of the QK^T product
  constexpr uint32_t IDESC_QK = umma_idesc_tf32(TC_BQ, TC_BK, 1, 1);
  constexpr uint32_t IDESC_PV = umma_idesc_tf32(TC_BQ, HD, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SW128 atoms need 1024-B alignment
  const uint32_t sQ = sbase;                                  // [hi|lo][QBOX][HD rows][128 B]
  const uint32_t sK = sQ + TcSmem<HD>::q_bytes;               // [hi|lo][KBOX][HD][128 B]
  const uint32_t sV = sK + TcSmem<HD>::kv_bytes;
  const uint32_t sBar = sV + TcSmem<HD>::kv_bytes;            // 5 mbarriers + tmem slot
  const uint32_t bar_q = sBar, bar_k = sBar + 8, bar_v = sBar + 16, bar_s = sBar + 24, bar_o = sBar + 32;
  const uint32_t tmem_slot = sBar + 40;

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int q0 = blockIdx.x * TC_BQ, head = blockIdx.y, b = blockIdx.z;

  int Leff = L;
  bool all_masked = false;
  if (lengths != nullptr) {
    const long long len = lengths[b];
    if (len <= 0) all_masked = true;
    else if (len < L) Leff = (int)len;
  }
  const int nkt = (Leff + TC_BK - 1) / TC_BK;

  // rows of the six operand planes in the tensor map
  const int plane = B * nh * HD;
  const int row_q = (b * nh + head) * HD;   // + which * plane

  if (tid == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_qk) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_v) : "memory");
  }
  if (warp == 0) {
    __syncwarp();   // lane 0 just left the barrier-init branch; .sync.aligned needs the whole warp converged
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);      // warp-uniform for ptxas (uniform-register UMMA operands)
  const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes

  auto load_kv = [&](uint32_t sdst, uint32_t bar, int which_hi, int k0) {
    const CUtensorMap* map = (which_hi == 4) ? &tmap_v : &tmap_qk;
    mbar_expect_tx(bar, TcSmem<HD>::kv_bytes);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int x = 0; x < KBOX; ++x)
        tma_load_2d(sdst + (h * KBOX + x) * BOX, map, k0 + x * TC_BOX, (which_hi + h) * plane + row_q, bar);
  };
  auto issue_qk = [&]() {
    // S[128 q, 64 keys] = sum over d: Q^T and K^T boxes are MN-major (positions contiguous)
#pragma unroll
    for (int term = 0; term < 3; ++term) {          // hi*hi, hi*lo, lo*hi
      const uint32_t qa = sQ + ((term == 2) ? QBOX * BOX : 0u);
      const uint32_t kb = sK + ((term == 1) ? KBOX * BOX : 0u);
#pragma unroll
      for (int ks = 0; ks < KSTEPS_D; ++ks) {
        // MN-major, 128B swizzle / 32B atoms: LBO = next 32 positions (next box), SBO = next 4 d-rows
        const uint64_t ad = umma_desc(qa + ks * 1024u, BOX, 512u, 1u);
        const uint64_t bd = umma_desc(kb + ks * 1024u, BOX, 512u, 1u);
        umma_tf32_ss(tmem_base + TC_COL_S, ad, bd, IDESC_QK, (term | ks) ? 1u : 0u);
      }
    }
  };
  auto issue_pv = [&]() {
    // O_tile[128 q, HD] = P[128, 64 keys] (TMEM) * V[64 keys, HD]; V^T boxes are K-major (keys contiguous)
#pragma unroll
    for (int term = 0; term < 3; ++term) {          // hi*hi, hi*lo, lo*hi
      const uint32_t pa = tmem_base + ((term == 2) ? TC_COL_PLO : TC_COL_PHI);
      const uint32_t vb = sV + ((term == 1) ? KBOX * BOX : 0u);
#pragma unroll
      for (int ks = 0; ks < TC_BK / 8; ++ks) {
        const uint64_t bd = umma_desc(vb + (ks >> 2) * BOX + (ks & 3) * 32u, 16u, 1024u, 2u);
        umma_tf32_ts(tmem_base + TC_COL_O, pa + ks * 8, bd, IDESC_PV, (term | ks) ? 1u : 0u);
      }
    }
  };

  if (tid == 0) {
    mbar_expect_tx(bar_q, TcSmem<HD>::q_bytes);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int x = 0; x < QBOX; ++x)
        tma_load_2d(sQ + (h * QBOX + x) * BOX, &tmap_qk, q0 + x * TC_BOX, h * plane + row_q, bar_q);
    load_kv(sK, bar_k, 2, 0);
    load_kv(sV, bar_v, 4, 0);
    mbar_wait(bar_q, 0);
    mbar_wait(bar_k, 0);
    tc_fence_after();
    issue_qk();
    tc_commit(bar_s);
  }
  __syncwarp();

  float o[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) o[c] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  for (int t = 0; t < nkt; ++t) {
    const uint32_t par = (uint32_t)(t & 1);
    mbar_wait(bar_s, par);
    tc_fence_after();
    if (tid == 0 && t + 1 < nkt) load_kv(sK, bar_k, 2, (t + 1) * TC_BK);  // K buffer is free: QK(t) committed
    __syncwarp();

    // ---- S row -> registers ----
    uint32_t sr[TC_BK];
#pragma unroll
    for (int c = 0; c < TC_BK; c += 16) tmem_ld16(t_lane + TC_COL_S + c, sr + c);
    tmem_wait_ld();
    if (dbg_s != nullptr && t == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
#pragma unroll
      for (int j = 0; j < TC_BK; ++j) dbg_s[tid * TC_BK + j] = __uint_as_float(sr[j]);
    }
    float s[TC_BK];
    const int kbase = t * TC_BK;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < TC_BK; ++j) {
      float v = __uint_as_float(sr[j]);
      if (all_masked) v = (kbase + j < L) ? 0.f : -INFINITY;
      else if (kbase + j >= Leff) v = -INFINITY;
      s[j] = v;
      mx = fmaxf(mx, v);
    }
    const float m_new = fmaxf(m_run, mx);
    const float alpha = exp2f(m_run - m_new);
    m_run = m_new;
    float rs = 0.f;
#pragma unroll
    for (int c = 0; c < TC_BK; c += 16) {
      uint32_t ph[16], pl[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p = exp2f(s[c + j] - m_new);
        rs += p;
        ph[j] = tf32_hi(p);
        pl[j] = tf32_hi(p - __uint_as_float(ph[j]));
      }
      tmem_st16(t_lane + TC_COL_PHI + c, ph);
      tmem_st16(t_lane + TC_COL_PLO + c, pl);
    }
    l_run = l_run * alpha + rs;
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();   // every row's P is in TMEM and every S load has retired

    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_v, par);
      tc_fence_after();
      issue_pv();
      tc_commit(bar_o);
      if (t + 1 < nkt) {   // S columns are free again: queue the next tile's scores behind PV(t)
        mbar_wait(bar_k, par ^ 1u);
        tc_fence_after();
        issue_qk();
        tc_commit(bar_s);
      }
    }
    __syncwarp();

    mbar_wait(bar_o, par);
    __syncwarp();
    tc_fence_after();
    if (tid == 0 && t + 1 < nkt) load_kv(sV, bar_v, 4, (t + 1) * TC_BK);  // V buffer is free: PV(t) committed
    __syncwarp();
#pragma unroll
    for (int c = 0; c < HD; c += 16) {
      uint32_t orr[16];
      tmem_ld16(t_lane + TC_COL_O + c, orr);
      tmem_wait_ld();
      if (dbg_o != nullptr && t == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dbg_o[tid * HD + c + j] = __uint_as_float(orr[j]);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) o[c + j] = fmaf(o[c + j], alpha, __uint_as_float(orr[j]));
    }
  }

  // ---- epilogue: normalise and store this thread's query row ----
  const int qi = q0 + tid;
  if (qi < L) {
    const float inv = 1.0f / l_run;
    float* dst = ctx + ((long long)b * L + qi) * (nh * HD) + head * HD;
    if (ctx_lo == nullptr) {
#pragma unroll
      for (int c = 0; c < HD; c += 4)
        *reinterpret_cast<float4*>(dst + c) = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
    } else {   // hi/lo planes for the tensor-core out_proj
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
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}


// =================================================================================================
// Warp-specialised variant (default): one CTA per SM-slot works on TWO 128-query tiles of the same
// (utterance, head) that share every K/V tile. Roles: warp 0 = TMA loader, warp 1 = UMMA issuer,
// warps 4-7 / 8-11 = softmax warpgroups of query tile A / B (thread = query row = TMEM lane).
// The issuer alternates  PV(A,t) QK(A,t+1) | PV(B,t) QK(B,t+1)  so the tensor pipe runs one tile's
// GEMMs while the other tile's warpgroup does its softmax; K and V tiles are double-buffered rings.
// The hi/lo split is folded into the N dimension so that every UMMA is wide (a tf32 UMMA with M = 128, K = 8 costs
// ~max(N/2, 32 + N/4) cycles, measured with mma_bench.cu, so narrow ones waste the pipe):
//   S[:, 0:64]   = Q_hi [K_hi | K_lo]^T (one N = 128 UMMA per k-step; columns 64:128 hold the Q_hi K_lo^T part)
//   S[:, 0:64]  += Q_lo K_hi^T          (N = 64)
//   O[:, 0:2HD]  = P_hi [V_hi | V_lo]   (N = 2 HD);   O[:, 0:HD] += P_lo V_hi   (N = HD)
// and the softmax warpgroup adds the two column halves when it reads S and O.
// TMEM per query tile (224 columns): S (128; P_hi overwrites columns 0:64 and P_lo columns 64:128) | O (2 HD <= 96).
// =================================================================================================
constexpr int WS_THREADS = 384;
constexpr uint32_t WS_TMEM_COLS = 512;
constexpr uint32_t WS_COL_TILE = 224, WS_COL_S = 0, WS_COL_PLO = 64, WS_COL_O = 128;
constexpr int WS_KV_STAGES = 2;

template <int HD>
struct WsSmem {
  static constexpr uint32_t box_bytes = HD * 128;
  static constexpr uint32_t q_bytes = 2 * (TC_BQ / TC_BOX) * box_bytes;    // one query tile, hi + lo
  static constexpr uint32_t kv_bytes = 2 * (TC_BK / TC_BOX) * box_bytes;   // one K (or V) tile, hi + lo
  static constexpr uint32_t off_k = 2 * q_bytes;
  static constexpr uint32_t off_v = off_k + WS_KV_STAGES * kv_bytes;
  static constexpr uint32_t off_bar = off_v + WS_KV_STAGES * kv_bytes;
  static constexpr uint32_t total = off_bar + 256 + 1024 /*align slack*/;
};


template <int HD>
__global__ void __launch_bounds__(WS_THREADS, 1)
attention_ws_kernel(const __grid_constant__ CUtensorMap tmap_qk, const __grid_constant__ CUtensorMap tmap_v,
                    float* __restrict__ ctx, const int64_t* __restrict__ lengths, int B, int L, int nh,
                    float* __restrict__ ctx_lo, long long* __restrict__ prof) {
  static_assert(HD % 16 == 0 && HD >= 16 && HD <= 64, "tensor-core path: head_dim in {16,32,48,64}");
  constexpr uint32_t BOX = WsSmem<HD>::box_bytes;
  constexpr int QBOX = TC_BQ / TC_BOX, KBOX = TC_BK / TC_BOX;
  constexpr int KSTEPS_D = HD / 8;
  constexpr uint32_t IDESC_QK2 = umma_idesc_tf32(TC_BQ, 2 * TC_BK, 1, 1);   // Q_hi x [K_hi | K_lo]
  constexpr uint32_t IDESC_QK1 = umma_idesc_tf32(TC_BQ, TC_BK, 1, 1);       // Q_lo x K_hi
  constexpr uint32_t IDESC_PV2 = umma_idesc_tf32(TC_BQ, 2 * HD, 0, 0);      // P_hi x [V_hi | V_lo]
  constexpr uint32_t IDESC_PV1 = umma_idesc_tf32(TC_BQ, HD, 0, 0);          // P_lo x V_hi

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = sbase;                                   // [tile][hi|lo][QBOX][HD rows][128 B]
  const uint32_t sK = sbase + WsSmem<HD>::off_k;               // [stage][hi|lo][KBOX][HD][128 B]; V: [stage][KBOX][hi|lo][HD][128 B]
  const uint32_t sV = sbase + WsSmem<HD>::off_v;
  const uint32_t sBar = sbase + WsSmem<HD>::off_bar;
  // barriers: q_full[2] k_full[2] k_empty[2] v_full[2] v_empty[2] s_full[2] p_ready[2] o_ready[2] o_free[2]
  const uint32_t bar_qf = sBar, bar_kf = sBar + 16, bar_ke = sBar + 32, bar_vf = sBar + 48, bar_ve = sBar + 64;
  const uint32_t bar_sf = sBar + 80, bar_pr = sBar + 96, bar_or = sBar + 112, bar_of = sBar + 128;
  const uint32_t tmem_slot = sBar + 144;

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int q0 = blockIdx.x * (2 * TC_BQ), head = blockIdx.y, b = blockIdx.z;

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
      mbar_init(bar_qf + 8 * i, 1); mbar_init(bar_kf + 8 * i, 1); mbar_init(bar_ke + 8 * i, 1);
      mbar_init(bar_vf + 8 * i, 1); mbar_init(bar_ve + 8 * i, 1); mbar_init(bar_sf + 8 * i, 1);
      mbar_init(bar_pr + 8 * i, 4); mbar_init(bar_or + 8 * i, 1); mbar_init(bar_of + 8 * i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_qk) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_v) : "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(WS_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);      // warp-uniform for ptxas (uniform-register UMMA operands)

  if (warp == 0) {
    if (lane == 0) {
      // ===== loader: both query tiles once, then the K / V rings =====
      for (int x = 0; x < 2; ++x) {
        mbar_expect_tx(bar_qf + 8 * x, WsSmem<HD>::q_bytes);
        for (int h = 0; h < 2; ++h)
          for (int j = 0; j < QBOX; ++j)
            tma_load_2d(sQ + (uint32_t)x * WsSmem<HD>::q_bytes + (h * QBOX + j) * BOX, &tmap_qk, q0 + x * TC_BQ + j * TC_BOX,
                        h * plane + row_q, bar_qf + 8 * x);
      }
      for (int t = 0; t < nkt; ++t) {
        const int st = t & 1;
        const uint32_t par_prev = (uint32_t)(((t >> 1) - 1) & 1);
        if (t >= WS_KV_STAGES) mbar_wait(bar_ke + 8 * st, par_prev);
        mbar_expect_tx(bar_kf + 8 * st, WsSmem<HD>::kv_bytes);
        for (int h = 0; h < 2; ++h)
          for (int j = 0; j < KBOX; ++j)
            tma_load_2d(sK + (uint32_t)st * WsSmem<HD>::kv_bytes + (h * KBOX + j) * BOX, &tmap_qk, t * TC_BK + j * TC_BOX,
                        (2 + h) * plane + row_q, bar_kf + 8 * st);
        if (t >= WS_KV_STAGES) mbar_wait(bar_ve + 8 * st, par_prev);
        mbar_expect_tx(bar_vf + 8 * st, WsSmem<HD>::kv_bytes);
        for (int h = 0; h < 2; ++h)
          for (int j = 0; j < KBOX; ++j)
            tma_load_2d(sV + (uint32_t)st * WsSmem<HD>::kv_bytes + (j * 2 + h) * BOX, &tmap_v, t * TC_BK + j * TC_BOX,
                        (4 + h) * plane + row_q, bar_vf + 8 * st);
      }
    }
  } else if (warp == 1) {
    // ===== UMMA issuer: the whole warp runs the loop (uniform operands), one elected lane issues =====
    auto issue_qk = [&](int x, int st) {
      const uint32_t q = sQ + (uint32_t)x * WsSmem<HD>::q_bytes, k = sK + (uint32_t)st * WsSmem<HD>::kv_bytes;
      const uint32_t d = tmem_base + (uint32_t)x * WS_COL_TILE + WS_COL_S;
      // MN-major, 128B swizzle / 32B atoms: LBO = next 32 positions (next box; the K_lo boxes follow the K_hi boxes,
      // so N = 128 simply runs on into them), SBO = next 4 d-rows
#pragma unroll
      for (int ks = 0; ks < KSTEPS_D; ++ks)
        umma_tf32_ss_w(d, umma_desc(q + ks * 1024u, BOX, 512u, 1u), umma_desc(k + ks * 1024u, BOX, 512u, 1u), IDESC_QK2, ks ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < KSTEPS_D; ++ks)
        umma_tf32_ss_w(d, umma_desc(q + QBOX * BOX + ks * 1024u, BOX, 512u, 1u), umma_desc(k + ks * 1024u, BOX, 512u, 1u), IDESC_QK1, 1u);
    };
    auto issue_pv = [&](int x, int st, uint32_t accumulate) {
      const uint32_t v = sV + (uint32_t)st * WsSmem<HD>::kv_bytes;
      const uint32_t tb = tmem_base + (uint32_t)x * WS_COL_TILE;
      // V^T boxes are K-major (keys contiguous): rows = d; per 32-key box the V_hi rows are followed by the V_lo rows.
      // O accumulates in TMEM over ALL key tiles (the softmax warpgroup rescales it in place when the row maximum
      // grows by more than 2^8, see below), so only the very first UMMA of a query tile overwrites.
#pragma unroll
      for (int ks = 0; ks < TC_BK / 8; ++ks)
        umma_tf32_ts_w(tb + WS_COL_O, tb + WS_COL_S + ks * 8, umma_desc(v + (ks >> 2) * (2 * BOX) + (ks & 3) * 32u, 16u, 1024u, 2u),
                       IDESC_PV2, ks ? 1u : accumulate);
#pragma unroll
      for (int ks = 0; ks < TC_BK / 8; ++ks)
        umma_tf32_ts_w(tb + WS_COL_O, tb + WS_COL_PLO + ks * 8, umma_desc(v + (ks >> 2) * (2 * BOX) + (ks & 3) * 32u, 16u, 1024u, 2u),
                       IDESC_PV1, 1u);
    };
    mbar_wait(bar_kf, 0);
    for (int x = 0; x < 2; ++x) {
      mbar_wait(bar_qf + 8 * x, 0);
      tc_fence_after();
      issue_qk(x, 0);
      tc_commit_w(bar_sf + 8 * x);
    }
    tc_commit_w(bar_ke);
    for (int t = 0; t < nkt; ++t) {
      const int st = t & 1, sn = (t + 1) & 1;
      const uint32_t par = (uint32_t)(t & 1);
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        mbar_wait(bar_pr + 8 * x, par);                         // P(x,t) is in TMEM (and O has been rescaled if needed)
        if (x == 0) mbar_wait(bar_vf + 8 * st, (uint32_t)((t >> 1) & 1));
        tc_fence_after();
        issue_pv(x, st, t > 0 ? 1u : 0u);
        if (t + 1 == nkt) tc_commit_w(bar_or + 8 * x);          // the query tile's O is complete
        if (x == 1) tc_commit_w(bar_ve + 8 * st);
        if (t + 1 < nkt) {
          if (x == 0) { mbar_wait(bar_kf + 8 * sn, (uint32_t)(((t + 1) >> 1) & 1)); tc_fence_after(); }
          issue_qk(x, sn);                                      // S columns are free: PV(x,t) is queued ahead of it
          tc_commit_w(bar_sf + 8 * x);
          if (x == 1) tc_commit_w(bar_ke + 8 * sn);
        }
      }
    }
  } else if (warp >= 4) {
    // ===== softmax warpgroup of query tile x: thread = query row =====
    const int x = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)x * WS_COL_TILE;
    // Online softmax with LAZY rescaling: p = 2^(s - m_ref) where m_ref is a per-row reference that is only raised
    // when the running maximum exceeds it by more than 2^8 (p stays <= 256, harmless in fp32 / TF32 hi+lo), so O can
    // accumulate in TMEM across key tiles and the warpgroup touches it only on those rare occasions and at the end.
    float m_ref = -INFINITY, l_run = 0.f;
    for (int t = 0; t < nkt; ++t) {
      const uint32_t par = (uint32_t)(t & 1);
      mbar_wait(bar_sf + 8 * x, par);
      const bool pr = prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && x == 0 && row == 0 && t < 48;
      if (pr) prof[t * 8 + 0] = clock64();
      __syncwarp();
      tc_fence_after();
      uint32_t sr[TC_BK];
#pragma unroll
      for (int c = 0; c < TC_BK; c += 16) tmem_ld16(t_lane + WS_COL_S + c, sr + c);
#pragma unroll
      for (int c = 0; c < TC_BK; c += 16) {       // + the Q_hi K_lo^T part
        uint32_t s2[16];
        tmem_ld16(t_lane + WS_COL_S + TC_BK + c, s2);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) sr[c + j] = __float_as_uint(__uint_as_float(sr[c + j]) + __uint_as_float(s2[j]));
      }
      if (pr) prof[t * 8 + 1] = clock64();
      const int kbase = t * TC_BK;
      if (all_masked || kbase + TC_BK > Leff) {      // only the last key tile (or a fully masked utterance) needs masking
#pragma unroll
        for (int j = 0; j < TC_BK; ++j) {
          float v = __uint_as_float(sr[j]);
          if (all_masked) v = (kbase + j < L) ? 0.f : -INFINITY;
          else if (kbase + j >= Leff) v = -INFINITY;
          sr[j] = __float_as_uint(v);
        }
      }
      float mx = __uint_as_float(sr[0]);
#pragma unroll
      for (int j = 1; j < TC_BK; ++j) mx = fmaxf(mx, __uint_as_float(sr[j]));
      if (t == 0) {
        m_ref = mx;                                  // key 0 is never masked, so mx is finite; PV(0) overwrites O
      } else if (__any_sync(0xffffffffu, mx > m_ref + 8.0f)) {
        // S(t) complete implies PV(t-1) complete (issued ahead of it), and PV(t) waits for our arrival below:
        // O is quiescent, rescale this warp's rows in place.
        const float m_new = fmaxf(m_ref, mx);
        const float alpha = ws_ex2(m_ref - m_new);
        m_ref = m_new;
        l_run *= alpha;
#pragma unroll
        for (int c = 0; c < 2 * HD; c += 16) {
          uint32_t orr[16];
          tmem_ld16(t_lane + WS_COL_O + c, orr);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) orr[j] = __float_as_uint(__uint_as_float(orr[j]) * alpha);
          tmem_st16(t_lane + WS_COL_O + c, orr);
        }
      }
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < TC_BK; c += 16) {
        uint32_t ph[16], pl[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p = ws_ex2(__uint_as_float(sr[c + j]) - m_ref);
          rs += p;
          ph[j] = tf32_hi(p);
          pl[j] = __float_as_uint(p - __uint_as_float(ph[j]));
        }
        tmem_st16(t_lane + WS_COL_S + c, ph);       // P_hi overwrites the scores in place
        tmem_st16(t_lane + WS_COL_PLO + c, pl);
      }
      l_run += rs;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_pr + 8 * x) : "memory");
      if (pr) prof[t * 8 + 2] = clock64();
    }
    // ---- the query tile's O is complete: fetch it (both column halves), normalise, store ----
    mbar_wait(bar_or + 8 * x, 0);
    __syncwarp();
    tc_fence_after();
    float o[HD];
#pragma unroll
    for (int c = 0; c < HD; c += 16) {
      uint32_t orr[16], or2[16];
      tmem_ld16(t_lane + WS_COL_O + c, orr);
      tmem_ld16(t_lane + WS_COL_O + HD + c, or2);    // the P_hi V_lo part
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) o[c + j] = __uint_as_float(orr[j]) + __uint_as_float(or2[j]);
    }
    const int qi = q0 + x * TC_BQ + row;
    if (qi < L) {
      const float inv = 1.0f / l_run;
      float* dst = ctx + ((long long)b * L + qi) * (nh * HD) + head * HD;
      if (ctx_lo == nullptr) {
#pragma unroll
        for (int c = 0; c < HD; c += 4)
          *reinterpret_cast<float4*>(dst + c) = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
      } else {
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(WS_TMEM_COLS) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

template <int HD>
static int launch_tc_hd(const CUtensorMap& tmap_qk, const CUtensorMap& tmap_v, float* ctx, const int64_t* lengths, int B, int L, int nh,
                        cudaStream_t s, float* dbg_s, float* dbg_o, float* ctx_lo) {
  const size_t smem = TcSmem<HD>::total;
  M2_CUDA_OK(allow_smem(attention_tc_kernel<HD>, smem));
  dim3 grid(ceil_div(L, TC_BQ), nh, B);
  M2_LAUNCH(M2TTS_STAGE_ATTENTION, attention_tc_kernel<HD>, grid, TC_THREADS, smem, s, tmap_qk, tmap_v, ctx, lengths, B, L, nh, dbg_s, dbg_o, ctx_lo);
  return M2TTS_OK;
}

long long* g_ws_prof = nullptr;      // shared with attention_h.cu
template <int HD>
static int launch_ws_hd(const CUtensorMap& tmap_qk, const CUtensorMap& tmap_v, float* ctx, const int64_t* lengths, int B, int L, int nh,
                        cudaStream_t s, float* ctx_lo) {
  const size_t smem = WsSmem<HD>::total;
  M2_CUDA_OK(allow_smem(attention_ws_kernel<HD>, smem));
  dim3 grid(ceil_div(L, 2 * TC_BQ), nh, B);
  M2_LAUNCH(M2TTS_STAGE_ATTENTION, attention_ws_kernel<HD>, grid, WS_THREADS, smem, s, tmap_qk, tmap_v, ctx, lengths, B, L, nh, ctx_lo, g_ws_prof);
  return M2TTS_OK;
}

bool attention_tc_supported(int hd) { return hd == 16 || hd == 32 || hd == 48 || hd == 64; }

// qkv6: [6][B][nh][hd][Lp] fp32 (Q_hi,Q_lo,K_hi,K_lo,V_hi,V_lo), Q pre-scaled by scale*log2e.
int launch_attention_tc(const float* qkv6, float* ctx, const int64_t* lengths, int B, int L, int Lp, int nh,
                        int hd, cudaStream_t s, float* dbg_s, float* dbg_o, float* ctx_lo) {
  M2_REQUIRE(qkv6 && ctx, M2TTS_E_NULLPTR, "attention_tc: null pointer");
  M2_REQUIRE(attention_tc_supported(hd), M2TTS_E_UNSUPPORTED, "attention_tc: head_dim %d unsupported", hd);
  M2_REQUIRE(B > 0 && L > 0 && nh > 0 && B <= 65535 && nh <= 65535 && (Lp & 3) == 0 && Lp >= L, M2TTS_E_BADSHAPE,
             "attention_tc: B=%d L=%d Lp=%d nh=%d", B, L, Lp, nh);
  M2_REQUIRE((((uintptr_t)qkv6) & 15) == 0 && ((nh * hd) & 3) == 0, M2TTS_E_BADSHAPE, "attention_tc: misaligned operands");
  EncodeTiledFn enc = get_encode_fn();
  M2_REQUIRE(enc != nullptr, M2TTS_E_CUDA, "attention_tc: cuTensorMapEncodeTiled unavailable");
  CUtensorMap tmap_qk, tmap_v;
  const cuuint64_t dims[2] = {(cuuint64_t)L, (cuuint64_t)6 * B * nh * hd};
  const cuuint64_t strides[1] = {(cuuint64_t)Lp * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BOX, (cuuint32_t)hd};
  const cuuint32_t estr[2] = {1, 1};
  // Q/K boxes feed MN-major TF32 operands -> 128-B swizzle with 32-B atoms; V^T boxes are K-major -> plain 128-B swizzle
  CUresult r = enc(&tmap_qk, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)qkv6, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "attention_tc: cuTensorMapEncodeTiled (q/k) failed (%d)", (int)r);
  r = enc(&tmap_v, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)qkv6, dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "attention_tc: cuTensorMapEncodeTiled (v) failed (%d)", (int)r);
  if (dbg_s == nullptr && dbg_o == nullptr && hd <= 48) {   // warp-specialised two-tile pipeline (head_dim 64 does not fit two query tiles: single-warpgroup kernel)
    switch (hd) {
      case 16: return launch_ws_hd<16>(tmap_qk, tmap_v, ctx, lengths, B, L, nh, s, ctx_lo);
      case 32: return launch_ws_hd<32>(tmap_qk, tmap_v, ctx, lengths, B, L, nh, s, ctx_lo);
      default: return launch_ws_hd<48>(tmap_qk, tmap_v, ctx, lengths, B, L, nh, s, ctx_lo);
    }
  }
  switch (hd) {
    case 16: return launch_tc_hd<16>(tmap_qk, tmap_v, ctx, lengths, B, L, nh, s, dbg_s, dbg_o, ctx_lo);
    case 32: return launch_tc_hd<32>(tmap_qk, tmap_v, ctx, lengths, B, L, nh, s, dbg_s, dbg_o, ctx_lo);
    case 48: return launch_tc_hd<48>(tmap_qk, tmap_v, ctx, lengths, B, L, nh, s, dbg_s, dbg_o, ctx_lo);
    default: return launch_tc_hd<64>(tmap_qk, tmap_v, ctx, lengths, B, L, nh, s, dbg_s, dbg_o, ctx_lo);
  }
}

}  // namespace m2

#ifdef M2TTS_TOOLS
// bring-up: device buffer of 48 x 8 int64 receiving phase timestamps of CTA 0 / query tile A (NULL = off)
extern "C" int m2tts_attention_set_prof(long long* dev_buf) { m2::g_ws_prof = dev_buf; return M2TTS_OK; }

// Test / bring-up entry: run the tensor-core attention on caller-prepared hi/lo planes and
// optionally dump the first score tile (128x64, pre-mask, log2 domain) and the first P*V tile.
extern "C" int m2tts_attention_tc_planes(const float* qkv6, float* ctx, const int64_t* lengths, int B, int L,
                                         int Lp, int nh, int hd, float* dbg_s, float* dbg_o,
                                         m2tts_stream_t stream) {
  return m2::launch_attention_tc(qkv6, ctx, lengths, B, L, Lp, nh, hd, (cudaStream_t)stream, dbg_s, dbg_o);
}
#endif  // M2TTS_TOOLS
