// lin_h.cu — the transformer's linear layers as ONE persistent, warp-specialised tcgen05 kernel with the 16-bit split
// (fp16 hi/lo operands, fp32 accumulation; see attention_h.cu for the error model):
//   C[rows, N] = A[rows, K] @ W[N, K]^T  (+bias, ReLU, +residual)      components.py:55,70-72,90,103; tts_model.py:223-226
// A arrives as fp16 hi/lo planes [2][R][K] (LayerNorm/split kernel below, the attention epilogue, or the FFN1 epilogue),
// W as fp16 hi/lo planes [2][N][K]. Unlike lingemm_tc_kernel (rowgemm_tc.cu), which re-loaded the whole weight matrix
// for every 128-row tile and ran load -> UMMA -> epilogue serially, here
//   * every weight row is RESIDENT in shared memory for the life of the CTA (<= 110 KB),
//   * warp 0 streams 128-row A tiles through a ring with TMA, warp 1 issues the UMMAs (warp-collective), warps 2-5
//     run the epilogue out of one of two TMEM accumulator buffers while the next unit's UMMAs execute,
//   * N is processed in passes of np <= 128 columns and the hi/lo split is folded into N: per k-step
//     A_hi x [W_hi ; W_lo] (N = 2 np) and A_lo x W_hi (N = np); the epilogue adds the two column halves.
// Epilogue modes: 0 = fp32 row-major (+bias, +residual), 1 = fp16 hi/lo planes [2][R][N] (+bias, ReLU) for the next
// GEMM, 3 = the attention operand planes [6][B][nh][hd][Lp] (Q rows scaled by scale*log2 e), saturated to fp16 range.
#include "common.cuh"
#include "attention_tc.cuh"
#include <cuda_fp16.h>

namespace m2 {

constexpr int LH_BM = 128;
#ifndef LH_NG
#define LH_NG 3
#endif
constexpr int LH_G = LH_NG;                  // epilogue warpgroups: the 16-column chunks of a pass are dealt round-robin (one warp per
                                           // scheduler cannot hide the latency of its own dependent instructions)
constexpr int LH_LNU = 4;                        // row groups a LayerNorm producer warp has in flight (K <= 96 on this path)
constexpr int LH_LNW = 8;                        // LayerNorm producer warps (A = split(LN(x)) written straight into the A ring)
constexpr int LH_THREADS = 64 + 128 * LH_G + 32 * LH_LNW;      // TMA producer, issuer, epilogue warps, LayerNorm producers

struct LinHArgs {
  int R, K, N;                     // rows, inner dim, outputs
  int np, n_passes, kboxes;        // columns per pass, passes, 32-column boxes of K
  int a_stages;
  const float* bias;               // [N] or null
  int relu;
  const float* residual; int ldr;  // fp32 [R][ldr] or null (mode 0)
  float* y; int ldy;               // mode 0
  __half* y_planes;                // mode 1: [2][R][N]
  __half* qkvh; long long plane_stride; int L, nh, hd, Lp; float qscale;   // mode 3
  int mode;
  int stg2;                        // 1: two staging tiles (units alternate), when shared memory allows
  int res_tma;                     // mode 0 with staging: the residual tile is TMA-loaded into the staging tile and updated in place
  int tpose;                       // mode 3 with staging: TRANSPOSED product (weights = A, positions = N): a thread is an output channel and holds 16
                                   // consecutive positions per chunk, so the attention operand planes leave as 16-byte row pieces (see the epilogue)
  int tpu;                         // mode 3 with staging: 128-row tiles per utterance (tiles do not straddle utterances); else 0
  int stg_bytes;                   // > 0: the epilogue stages the output tile in shared memory and writes it with TMA stores (modes 0, 1)
  long long* prof;                 // bring-up: phase timestamps of CTA 0's first epilogue warp (m2tts_attention_set_prof buffer)
  int dbg;                         // bring-up timing experiments (M2TTS_LIN_DBG): 1 no stores, 2 no UMMAs, 4 no A-tile loads; results invalid
  int32_t* status;                 // fp16-plane outputs (modes 1, 3): M2TTS_ST_FP16_RANGE
  const float* ln_x; const float* ln_w; const float* ln_b; float ln_eps;      // A = split(LayerNorm(ln_x)) by the producer warps (null: A tiles by TMA)
};

// Operand rows are 32 halves = 64 bytes (64-B swizzle): K = 96 is then exactly three boxes (with 128-byte rows a
// quarter of every weight/activation tile would be zero padding, and the QKV weights would not leave room for a ring).
__device__ __forceinline__ uint64_t lh_desc(uint32_t saddr) {   // K-major, 64-B swizzle: SBO = 8 rows = 512 B
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void lh_mma_w(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void lh_tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(LH_THREADS, 1)
lin_h_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_y,
             const __grid_constant__ CUtensorMap tmap_r, const LinHArgs a) {
  pdl_launch_dependents();      // the next kernel of the stream may take the SMs this grid's CTAs leave (M2_LAUNCH_PDL)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t a_box = LH_BM * 64u;                                   // 128 rows x 32 halves
  const uint32_t a_stage = 2u * a.kboxes * a_box;                       // hi + lo
  const uint32_t w_box = (uint32_t)a.np * 64u;
  const uint32_t w_bytes = (uint32_t)a.kboxes * a.n_passes * 2u * w_box;
  const uint32_t sA = sbase;                                            // [stage][plane][kbox][128 x 64 B]
  const uint32_t sStg = sA + (uint32_t)a.a_stages * a_stage;            // output staging: 8 KB boxes of 128 rows x 64 B, 64-byte swizzle
  const uint32_t sW = sStg + (uint32_t)a.stg_bytes * (uint32_t)(1 + a.stg2);       // [kbox][pass][plane][np x 64 B]
  const uint32_t sBias = sW + w_bytes;                                  // N floats
  const uint32_t sLn = (sBias + (uint32_t)a.N * 4u + 15u) & ~15u;         // LayerNorm weight and bias of the producers: 2 x 96 floats
  const uint32_t sBar = sLn + 768u;
  // barriers: a_full[4] a_empty[4] acc_full[2] acc_empty[2] w_full
  const uint32_t bar_af = sBar, bar_ae = sBar + 32, bar_cf = sBar + 64, bar_ce = sBar + 80, bar_w = sBar + 96, bar_rs = sBar + 104, tmem_slot = sBar + 112;
  float* bias_s = reinterpret_cast<float*>(gbase + (sBias - sbase));

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int m_tiles = a.tpu > 0 ? (a.R / a.L) * a.tpu : (a.R + LH_BM - 1) / LH_BM;
  const int S = a.a_stages;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(bar_af + 8 * i, a.ln_x != nullptr ? (uint32_t)LH_LNW : 1u); mbar_init(bar_ae + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_cf + 8 * i, 1); mbar_init(bar_ce + 8 * i, 4 * LH_G); }
    mbar_init(bar_w, 1); mbar_init(bar_rs, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_y) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_r) : "memory");
  }
  for (int i = tid; i < a.N; i += LH_THREADS) bias_s[i] = a.bias != nullptr ? a.bias[i] : 0.f;
  float* ln_s = reinterpret_cast<float*>(gbase + (sLn - sbase));
  if (a.ln_x != nullptr)      // the LayerNorm parameters are weights, nobody's output: no pdl_wait needed
    for (int i = tid; i < 2 * a.K; i += LH_THREADS) ln_s[i < a.K ? i : 96 + (i - a.K)] = i < a.K ? a.ln_w[i] : a.ln_b[i - a.K];
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
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
      // ===== producer: all weight rows once, then the A tiles =====
      mbar_expect_tx(bar_w, w_bytes);
      for (int kb = 0; kb < a.kboxes; ++kb)
        for (int p = 0; p < a.n_passes; ++p)
          for (int pl = 0; pl < 2; ++pl)
            tma_load_2d(sW + (uint32_t)((kb * a.n_passes + p) * 2 + pl) * w_box, &tmap_w, kb * 32, pl * a.N + p * a.np, bar_w);
      pdl_wait();      // the weights are nobody's output; the A tiles are the previous kernel's
      int it = 0;
      for (int mt = blockIdx.x; mt < m_tiles && a.ln_x == nullptr; mt += gridDim.x, ++it) {
        const int st = it % S;
        if (it >= S) mbar_wait(bar_ae + 8 * st, (uint32_t)(((it / S) - 1) & 1));
        if (a.dbg & 4) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_af + 8 * st) : "memory"); continue; }
        mbar_expect_tx(bar_af + 8 * st, a_stage);
        const int nb = a.tpu > 0 ? a.R / a.L : 0, ub = a.tpu > 0 ? mt / a.tpu : 0, ul = a.tpu > 0 ? (mt % a.tpu) * LH_BM : 0;
        for (int pl = 0; pl < 2; ++pl)
          for (int kb = 0; kb < a.kboxes; ++kb) {
            const uint32_t dst = sA + (uint32_t)st * a_stage + (uint32_t)(pl * a.kboxes + kb) * a_box;
            if (a.tpu > 0) lh_tma_load_3d(dst, &tmap_a, kb * 32, ul, pl * nb + ub, bar_af + 8 * st);      // rows >= L of the utterance: zero-filled
            else tma_load_2d(dst, &tmap_a, kb * 32, pl * a.R + mt * LH_BM, bar_af + 8 * st);
          }
      }
    }
  } else if (warp == 1) {
    // ===== UMMA issuer (whole warp, one elected lane issues) =====
    mbar_wait(bar_w, 0);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * a.np) >> 3) << 17) | ((uint32_t)(LH_BM >> 4) << 24);   // fp16 x fp16 -> fp32, K-major
    const uint32_t idesc1 = (1u << 4) | ((uint32_t)(a.np >> 3) << 17) | ((uint32_t)(LH_BM >> 4) << 24);
    const uint32_t idescT = (1u << 4) | ((uint32_t)(LH_BM >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);      // transposed: M = 128 channels, N = 128 positions
    const int ksteps = a.K / 16;
    int it = 0, unit = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it) {
      const int st = it % S;
      mbar_wait(bar_af + 8 * st, (uint32_t)((it / S) & 1));
      tc_fence_after();
      const uint32_t aHi = sA + (uint32_t)st * a_stage, aLo = aHi + (uint32_t)a.kboxes * a_box;
      for (int p = 0; p < a.n_passes; ++p, ++unit) {
        const int buf = unit & 1;
        if (unit >= 2) mbar_wait(bar_ce + 8 * buf, (uint32_t)(((unit >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)buf * 256u;
        for (int ks = 0; ks < ksteps && !(a.dbg & 2); ++ks) {
          const uint32_t koff = (uint32_t)(ks >> 1) * a_box + (uint32_t)(ks & 1) * 32u;
          const uint32_t wb = sW + (uint32_t)(((ks >> 1) * a.n_passes + p) * 2) * w_box + (uint32_t)(ks & 1) * 32u;
          if (a.tpose) {
            // D[channel, position]: the pass's weight rows are the A operand (M = 128: the 96 rows of the pass and 32 rows of whatever
            // follows them in shared memory — lanes nobody reads), the activation tile is the B operand (N = 128 positions); the three
            // product terms accumulate into the same 128 columns
            lh_mma_w(d, lh_desc(wb), lh_desc(aHi + koff), idescT, ks ? 1u : 0u);      // W_hi x X_hi
            lh_mma_w(d, lh_desc(wb), lh_desc(aLo + koff), idescT, 1u);                // W_hi x X_lo
            lh_mma_w(d, lh_desc(wb + w_box), lh_desc(aHi + koff), idescT, 1u);        // W_lo x X_hi
            continue;
          }
          const uint64_t bd = lh_desc(wb);
          lh_mma_w(d, lh_desc(aHi + koff), bd, idesc2, ks ? 1u : 0u);     // A_hi x [W_hi ; W_lo]
          lh_mma_w(d, lh_desc(aLo + koff), bd, idesc1, 1u);               // A_lo x W_hi
        }
        tc_commit_w(bar_cf + 8 * buf);
      }
      tc_commit_w(bar_ae + 8 * st);        // the A tile is free once every pass has read it
    }
  } else if (warp >= 2 + 4 * LH_G) {
    // ===== LayerNorm producers (a.ln_x): A tile = split(LN(x rows)) written into the ring stage in the TMA box layout (per plane
    // and 32-column k-box: 128 rows x 64 B, 64-byte swizzle). 8 lanes per row (a float4 per 32 columns each), 4 rows per warp
    // pass, 32 rows per warp and tile. Replaces the ln_split_h launch and the HBM round trip of the normalised planes. =====
    if (a.ln_x != nullptr) {
      pdl_wait();
      const int pw = warp - (2 + 4 * LH_G), sub = lane & 7;
      const int n4 = a.K >> 5;                     // float4 per lane (K % 32 == 0, K <= 96)
      float amax = 0.f;
      static_assert(LH_LNW * 4 * LH_LNU == LH_BM, "LayerNorm producers: 16 rows per warp");
      // A warp owns 16 rows of every tile: four row groups of 4 rows, 8 lanes per row. The rows of the NEXT tile are loaded into
      // registers right after this tile's rows have been handed over, so their latency runs under the tile's UMMAs and epilogue
      // (the ring has a single stage next to the QKV / FFN weights: without the prefetch every tile waited for its own loads).
      float4 v[LH_LNU][3];
      bool valid[LH_LNU];
      auto load_tile = [&](int mt) {
#pragma unroll
        for (int u = 0; u < LH_LNU; ++u) {
          const int r = pw * (4 * LH_LNU) + u * 4 + (lane >> 3);      // row of the tile
          long long grow;
          if (a.tpu > 0) { const int b = mt / a.tpu, l = (mt % a.tpu) * LH_BM + r; valid[u] = l < a.L; grow = (long long)b * a.L + l; }
          else { grow = (long long)mt * LH_BM + r; valid[u] = grow < a.R; }
          const float* xr = a.ln_x + (valid[u] ? grow : 0) * a.K;
#pragma unroll
          for (int i = 0; i < 3; ++i) v[u][i] = i < n4 ? __ldg(reinterpret_cast<const float4*>(xr + 32 * i + 4 * sub)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      // The register prefetch keeps ONE tile's rows in flight per SM (48 KB), and they come from HBM. One bulk L2 prefetch per
      // tile, issued a tile period ahead of the loads by one lane, turns those loads into L2 hits (QKV 0.132 -> 0.126, FFN1 0.081
      // -> 0.076, mel projection 0.066 -> 0.058 ms per launch). What a producer warp spends per tile — ~7 k cycles for four row
      // groups of ~250 instructions — did not move with packed arithmetic (100 instructions per group), MUFU.RSQ instead of
      // the IEEE sequences, two or four groups interleaved, loads issued per group a tile ahead, L2-only loads or polling
      // back-off in the other roles (all measured, tools/lin_prof.py).
      auto prefetch_tile = [&](int mt) {
        if (pw != 0 || lane != 0 || mt >= m_tiles) return;
        long long grow; int rows;
        if (a.tpu > 0) { const int b = mt / a.tpu, l0 = (mt % a.tpu) * LH_BM; grow = (long long)b * a.L + l0; rows = a.L - l0 < LH_BM ? a.L - l0 : LH_BM; }
        else { grow = (long long)mt * LH_BM; rows = a.R - grow < LH_BM ? (int)(a.R - grow) : LH_BM; }
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.ln_x + grow * a.K), "r"((uint32_t)(rows * a.K * 4)) : "memory");
      };
      int it = 0, mt = blockIdx.x;
      if (mt < m_tiles) load_tile(mt);
      prefetch_tile(mt + (int)gridDim.x);
      for (; mt < m_tiles; ++it) {
        const int st = it % S;
        if (it >= S) mbar_wait(bar_ae + 8 * st, (uint32_t)(((it / S) - 1) & 1));
        const uint32_t stage_off = (sA - sbase) + (uint32_t)st * a_stage;
#pragma unroll
        for (int u = 0; u < LH_LNU; ++u) {
          const int r = pw * (4 * LH_LNU) + u * 4 + (lane >> 3);
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 3; ++i) s += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
          s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
          const float mean = s / (float)a.K;
          float q = 0.f;
#pragma unroll
          for (int i = 0; i < 3; ++i)
            if (i < n4) {
              const float d0 = v[u][i].x - mean, d1 = v[u][i].y - mean, d2 = v[u][i].z - mean, d3 = v[u][i].w - mean;
              q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
          q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
          const float rstd = 1.0f / sqrtf(q / (float)a.K + a.ln_eps);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            if (i < n4) {
              const int k = 32 * i + 4 * sub;
              const float4 ww = *reinterpret_cast<const float4*>(ln_s + k), bb = *reinterpret_cast<const float4*>(ln_s + 96 + k);
              uint2 hv = make_uint2(0u, 0u), lv = make_uint2(0u, 0u);
              if (valid[u]) {
                const float y0 = (v[u][i].x - mean) * rstd * ww.x + bb.x, y1 = (v[u][i].y - mean) * rstd * ww.y + bb.y;
                const float y2 = (v[u][i].z - mean) * rstd * ww.z + bb.z, y3 = (v[u][i].w - mean) * rstd * ww.w + bb.w;
                h_split_pair(f2_pack(y0, y1), hv.x, lv.x, amax);
                h_split_pair(f2_pack(y2, y3), hv.y, lv.y, amax);
              }
              // k-box i, row r, 16-byte chunk (4 sub) >> 3 = sub >> 1 swizzled by (r >> 1) & 3, 8 bytes at (sub & 1) * 8
              const uint32_t off = (uint32_t)r * 64u + ((((uint32_t)sub >> 1) ^ ((uint32_t)(r >> 1) & 3u)) << 4) + ((uint32_t)sub & 1u) * 8u;
              *reinterpret_cast<uint2*>(gbase + stage_off + (uint32_t)i * a_box + off) = hv;
              *reinterpret_cast<uint2*>(gbase + stage_off + (uint32_t)(a.kboxes + i) * a_box + off) = lv;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the UMMAs read these rows through the async proxy
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_af + 8 * st) : "memory");
        mt += (int)gridDim.x;
        if (mt < m_tiles) load_tile(mt);
        prefetch_tile(mt + (int)gridDim.x);
      }
      h_flag(h_amax_bad(amax), a.status);
    }
  } else {
    // ===== epilogue warpgroup eg: thread = row of the tile, chunks eg, eg + G, ... of every pass =====
    pdl_wait();      // before the first residual read / output write
    const int eg = (warp - 2) >> 2;
    const int qtr = warp & 3;
    const int row = qtr * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qtr * 32) << 16);
    const int H = a.nh * a.hd;
    // modes 0 / 1 with staging: the tile leaves through 64-byte-swizzled 8 KB boxes (128 rows x 16 floats, or x 32 halves per
    // plane) and TMA stores; written thread-per-row straight to global memory every store instruction touched 32 lines and
    // the stores were 2/3 of the kernel time (M2TTS_LIN_DBG=1)
    const bool tma_out = a.stg_bytes != 0;
    uint8_t* stg0 = gbase + (sStg - sbase);
    const uint32_t swz = (uint32_t)((row >> 1) & 3);
    const bool leader = warp == 2 && lane == 0;
    bool bad = false;               // an fp16-plane output left the fp16 range
    float amax = 0.f;               // running maximum of |x| of the packed-pair splits (NaN-propagating): one range check at the end
    int unit = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      const long long r = (long long)mt * LH_BM + row;
      bool valid = r < a.R;
      int b = 0, l = 0;
      if (a.tpu > 0) { b = mt / a.tpu; l = (mt % a.tpu) * LH_BM + row; valid = l < a.L; }
      else if (a.mode == 3 && valid) { b = (int)(r / a.L); l = (int)(r - (long long)b * a.L); }
      for (int p = 0; p < a.n_passes; ++p, ++unit) {
        const int buf = unit & 1;
        const int n0 = p * a.np;
        const bool pt = a.prof != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0 && unit < 48;
        if (pt) a.prof[unit * 8 + 0] = clock64();
        // with two staging tiles the units alternate and only the stores of the unit before the previous one must have drained
        const uint32_t so = a.stg2 ? (uint32_t)(unit & 1) * (uint32_t)a.stg_bytes : 0u;
        uint8_t* stg = stg0 + so;
        const uint32_t sStgU = sStg + so;
        if (a.res_tma) {                    // earlier TMA stores have read this staging tile
          if (leader) { if (a.stg2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
          asm volatile("bar.sync 1, %0;" ::"n"(128 * LH_G) : "memory");
        }
        if (a.res_tma && leader) {
          // residual tile -> staging tile (same boxes as the output; rows >= R are zero-filled). Read thread-per-row straight
          // from global memory these loads took 3-6 k cycles per unit (tools/lin_prof.py): 32 lines per load instruction.
          mbar_expect_tx(bar_rs, (uint32_t)a.np * 512u);
          for (int c = 0; c < (a.np >> 4); ++c)
            tma_load_2d(sStgU + (uint32_t)c * 8192u, &tmap_r, n0 + 16 * c, mt * LH_BM, bar_rs);
          // and the next unit's residual tile towards L2: the staging tile is single, so its load cannot start before this
          // unit's stores have drained, but it can at least find its data in L2
          int mt2 = mt, p2 = p + 1;
          if (p2 == a.n_passes) { p2 = 0; mt2 = mt + (int)gridDim.x; }
          if (mt2 < m_tiles)
            for (int c = 0; c < (a.np >> 4); ++c)
              asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
                           ::"l"(&tmap_r), "r"(p2 * a.np + 16 * c), "r"(mt2 * LH_BM) : "memory");
        }
        if (pt) a.prof[unit * 8 + 1] = clock64();
        mbar_wait(bar_cf + 8 * buf, (uint32_t)((unit >> 1) & 1));
        __syncwarp();
        tc_fence_after();
        if (a.res_tma) mbar_wait(bar_rs, (uint32_t)(unit & 1));
        else if (tma_out) {                 // no residual to fetch: the staging tile is only needed now
          if (leader) { if (a.stg2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
          asm volatile("bar.sync 1, %0;" ::"n"(128 * LH_G) : "memory");
        }
        if (pt) a.prof[unit * 8 + 2] = clock64();
        const uint32_t tb = t_lane + (uint32_t)buf * 256u;
        bool released = false;
        if (a.tpose) {
          // thread = output channel `row` of pass p (q, k or v), TMEM columns = the tile's 128 positions; a chunk = 16 positions =
          // 32 bytes of a plane row. Staging: per (plane, head, 64-position half) a box of hd rows x 128 B in the 128-byte swizzle of
          // the TMA store (the attention kernel's own load box): the 8 lanes of a quarter-warp hit 8 different 16-byte columns. With
          // positions in the lanes this epilogue issued 64 two-byte stores per chunk and thread; now four 16-byte ones.
          const bool act = row < a.np;                   // lanes 96..127 of the accumulator belong to nobody (warp-uniform)
          const int head = act ? row / a.hd : 0, dd = act ? row - head * a.hd : 0;
          const float sc = (p == 0) ? a.qscale : 1.0f;
          const uint64_t sc2 = f2_pack(sc, sc);
          const uint32_t boxb = (uint32_t)a.hd * 128u;
          for (int k = eg; k < LH_BM / 16; k += LH_G) {
            uint32_t v[16];
            if (act) { tmem_ld16(tb + 16 * k, v); tmem_wait_ld(); }
            if (k + LH_G >= LH_BM / 16) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_ce + 8 * buf) : "memory");
              released = true;
            }
            if (!act || (a.dbg & 1)) continue;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) h_split_pair(f2_mul(f2_pack_u(v[2 * j], v[2 * j + 1]), sc2), hi[j], lo[j], amax);
            uint8_t* bh = stg + (uint32_t)(head * 2 + (k >> 2)) * boxb + (uint32_t)dd * 128u;
            uint8_t* bl = bh + (uint32_t)(a.nh * 2) * boxb;
            const uint32_t c16 = (uint32_t)(k & 3) * 2u, sw = (uint32_t)dd & 7u;
            *reinterpret_cast<uint4*>(bh + ((c16 ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(bh + (((c16 + 1u) ^ sw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<uint4*>(bl + ((c16 ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(bl + (((c16 + 1u) ^ sw) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          }
        }
        for (int c0 = eg * 16; c0 < a.np && !a.tpose; c0 += 16 * LH_G) {
          uint32_t v[16], w[16];
          float4 rs[4];
          if (a.res_tma) {
            const uint8_t* bx = stg + (uint32_t)(c0 >> 4) * 8192u + (uint32_t)row * 64u;
#pragma unroll
            for (int j = 0; j < 4; ++j) rs[j] = *reinterpret_cast<const float4*>(bx + (((uint32_t)j ^ swz) << 4));
          } else if (a.mode == 0 && a.residual != nullptr && valid) {          // issue the residual loads ahead of the TMEM reads
            const float4* rp = reinterpret_cast<const float4*>(a.residual + r * a.ldr + n0 + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) rs[j] = __ldg(rp + j);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) rs[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          tmem_ld16(tb + c0, v);
          tmem_ld16(tb + a.np + c0, w);
          tmem_wait_ld();
          if (c0 + 16 * LH_G >= a.np) {     // this warpgroup's last TMEM read of the unit: hand the accumulator buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_ce + 8 * buf) : "memory");
            released = true;
          }
          if (!valid || (a.dbg & 1)) continue;
          float x[16];
          {
            const float4* bp = reinterpret_cast<const float4*>(bias_s + n0 + c0);      // 16-byte aligned: n0 + c0 is a multiple of 16
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bv = bp[j4];
              const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
              for (int e = 0; e < 4; e += 2) {      // packed pairs (FADD2): accumulator halves + bias
                const int j = 4 * j4 + e;
                float t0, t1;
                f2_unpack(f2_add(f2_add(f2_pack_u(v[j], v[j + 1]), f2_pack_u(w[j], w[j + 1])), f2_pack(bb[e], bb[e + 1])), t0, t1);
                if (a.relu) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); }
                x[j] = t0; x[j + 1] = t1;
              }
            }
          }
          if (a.mode == 0 && tma_out) {
            uint8_t* bx = stg + (uint32_t)(c0 >> 4) * 8192u + (uint32_t)row * 64u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(bx + (((uint32_t)j ^ swz) << 4)) =
                  make_float4(x[4 * j] + rs[j].x, x[4 * j + 1] + rs[j].y, x[4 * j + 2] + rs[j].z, x[4 * j + 3] + rs[j].w);
          } else if (a.mode == 0) {
            float4* yp = reinterpret_cast<float4*>(a.y + r * a.ldy + n0 + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              yp[j] = make_float4(x[4 * j] + rs[j].x, x[4 * j + 1] + rs[j].y, x[4 * j + 2] + rs[j].z, x[4 * j + 3] + rs[j].w);
          } else if (a.mode == 1) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) h_split_pair(f2_pack(x[2 * j], x[2 * j + 1]), hi[j], lo[j], amax);
            if (tma_out) {
              uint8_t* bh = stg + (uint32_t)(c0 >> 5) * 8192u + (uint32_t)row * 64u;
              uint8_t* bl = bh + (uint32_t)(a.np >> 5) * 8192u;
              const uint32_t p0 = (c0 & 16) ? 2u : 0u;
              *reinterpret_cast<uint4*>(bh + ((p0 ^ swz) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(bh + (((p0 + 1u) ^ swz) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
              *reinterpret_cast<uint4*>(bl + ((p0 ^ swz) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              *reinterpret_cast<uint4*>(bl + (((p0 + 1u) ^ swz) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            } else {
              uint4* hp = reinterpret_cast<uint4*>(a.y_planes + r * a.N + n0 + c0);
              uint4* lp = reinterpret_cast<uint4*>(a.y_planes + ((long long)a.R + r) * a.N + n0 + c0);
              hp[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); hp[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
              lp[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); lp[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
          } else {
            // head_dim is a multiple of 16, so the 16 columns of a chunk belong to ONE (q|k|v, head): the index
            // arithmetic (two integer divisions) is done once per chunk, not per column
            int which = p, head = 0, d0 = 0;
            if (!tma_out) {     // direct stores need (q|k|v, head, d); with staging a pass is exactly q, k or v and the box does the rest
              const int n = n0 + c0;
              which = n / H;
              const int rem = n - which * H;
              head = rem / a.hd; d0 = rem - head * a.hd;
            }
            const float sc = (which == 0) ? a.qscale : 1.0f;
            if (tma_out) {      // staging [plane][np d-rows][128 positions]: a warp writes 64 contiguous bytes per row
              __half* sp = reinterpret_cast<__half*>(stg) + (uint32_t)c0 * 128u + (uint32_t)row;
              // packed conversions (F2FP / HADD2.F32 run at full rate; the scalar cvt goes through the quarter-rate XU pipe and
              // made this epilogue XU-bound)
              const uint64_t sc2 = f2_pack(sc, sc);
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                uint32_t hw, lw;
                h_split_pair(f2_mul(f2_pack(x[j], x[j + 1]), sc2), hw, lw, amax);
                const __half2 h = *reinterpret_cast<const __half2*>(&hw), lo = *reinterpret_cast<const __half2*>(&lw);
                sp[j * 128] = __low2half(h);
                sp[(j + 1) * 128] = __high2half(h);
                sp[j * 128 + a.np * 128] = __low2half(lo);
                sp[(j + 1) * 128 + a.np * 128] = __high2half(lo);
              }
            } else {
              __half* hp = a.qkvh + (long long)(2 * which) * a.plane_stride + (((long long)b * a.nh + head) * a.hd + d0) * a.Lp + l;
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float t0 = x[j] * sc, t1 = x[j + 1] * sc;
                h_chk(t0, bad); h_chk(t1, bad);
                const __half2 h = __floats2half2_rn(t0, t1);
                const float2 hf = __half22float2(h);
                const __half2 lo = __floats2half2_rn(t0 - hf.x, t1 - hf.y);
                hp[(long long)j * a.Lp] = __low2half(h);
                hp[(long long)(j + 1) * a.Lp] = __high2half(h);
                hp[(long long)j * a.Lp + a.plane_stride] = __low2half(lo);
                hp[(long long)(j + 1) * a.Lp + a.plane_stride] = __high2half(lo);
              }
            }
          }
        }
        if (pt) a.prof[unit * 8 + 3] = clock64();
        if (!released) {                    // narrow pass: this warpgroup had no chunk, the buffer still needs its arrival
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_ce + 8 * buf) : "memory");
        }
        if (tma_out) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync 2, %0;" ::"n"(128 * LH_G) : "memory");
          if (leader && !(a.dbg & 1)) {
            const int r0 = mt * LH_BM;
            if (a.mode == 3 && a.tpose) {      // pass p = q, k or v: one box of 64 positions x hd rows per (plane, head, half)
              const int nbat = a.R / a.L, l0 = (mt % a.tpu) * LH_BM;
              for (int pl = 0; pl < 2; ++pl)
                for (int hh = 0; hh < a.nh; ++hh)
                  for (int hf = 0; hf < 2; ++hf)
                    if (l0 + 64 * hf < a.L)
                      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                   ::"l"(&tmap_y), "r"(l0 + 64 * hf), "r"((((2 * p + pl) * nbat + b) * a.nh + hh) * a.hd),
                                     "r"(sStgU + (uint32_t)(((pl * a.nh + hh) * 2 + hf) * a.hd * 128)) : "memory");
            } else if (a.mode == 3) {      // pass p = q, k or v: one box of 128 positions x hd rows per (plane, head)
              const int nbat = a.R / a.L, l0 = (mt % a.tpu) * LH_BM;
              for (int pl = 0; pl < 2; ++pl)
                for (int hh = 0; hh < a.nh; ++hh)
                  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                               ::"l"(&tmap_y), "r"(l0), "r"((((2 * p + pl) * nbat + b) * a.nh + hh) * a.hd),
                                 "r"(sStgU + (uint32_t)((pl * a.np + hh * a.hd) * 256)) : "memory");
            } else if (a.mode == 0) {
              for (int c = 0; c < (a.np >> 4); ++c)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(&tmap_y), "r"(n0 + 16 * c), "r"(r0), "r"(sStgU + (uint32_t)c * 8192u) : "memory");
            } else {
              const int nb = a.np >> 5;
              for (int pl = 0; pl < 2; ++pl)
                for (int c = 0; c < nb; ++c)
                  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                               ::"l"(&tmap_y), "r"(n0 + 32 * c), "r"(r0), "r"(pl), "r"(sStgU + (uint32_t)(pl * nb + c) * 8192u) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (pt) a.prof[unit * 8 + 4] = clock64();
      }
    }
    if (tma_out && leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    h_flag(bad || h_amax_bad(amax), a.status);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// LayerNorm over the last dim (optional) + fp16 hi/lo split: x [R,K] fp32 -> planes [2][R][K] fp16.
// 8 lanes per row (float4 each, up to 8 per lane: K <= 256), 4 rows per warp pass, grid-stride over row groups.
__global__ void __launch_bounds__(256) ln_split_h_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bvec, __half* __restrict__ planes,
                                                         long long R, int K, float eps, int32_t* __restrict__ status) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n4 = K >> 5;                       // float4 per lane (K % 32 == 0)
  bool bad = false;
  for (long long row = warp_id * 4 + (lane >> 3); row < ((R + 3) & ~3LL); row += warps * 4) {
    const bool valid = row < R;
    const float* xr = x + (valid ? row : 0) * K;
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < n4) { v[i] = __ldg(reinterpret_cast<const float4*>(xr + 32 * i + 4 * sub)); s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    float mean = 0.f, rstd = 1.f;
    if (w != nullptr) {
      s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
      mean = s / (float)K;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < n4) {
          const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
          q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
      q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
      rstd = 1.0f / sqrtf(q / (float)K + eps);
    }
    if (!valid) continue;
    __half* hp = planes + row * K;
    __half* lp = planes + (R + row) * K;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < n4) {
        const int k = 32 * i + 4 * sub;
        float a[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        if (w != nullptr) {
          const float4 ww = __ldg(reinterpret_cast<const float4*>(w + k)), bb = __ldg(reinterpret_cast<const float4*>(bvec + k));
          a[0] = (a[0] - mean) * rstd * ww.x + bb.x; a[1] = (a[1] - mean) * rstd * ww.y + bb.y;
          a[2] = (a[2] - mean) * rstd * ww.z + bb.z; a[3] = (a[3] - mean) * rstd * ww.w + bb.w;
        }
        uint2 hv, lv;
        h_split2(a[0], a[1], hv.x, lv.x, bad);
        h_split2(a[2], a[3], hv.y, lv.y, bad);
        *reinterpret_cast<uint2*>(hp + k) = hv;
        *reinterpret_cast<uint2*>(lp + k) = lv;
      }
  }
  h_flag(bad, status);
}

// W [N,K] fp32 -> fp16 planes [2][N][K]; several matrices per launch (blockIdx.y = job)
struct WSplitHJobs { const float* src[4]; __half* dst[4]; long long n[4]; int32_t* status; };
__global__ void w_split_h_kernel(WSplitHJobs jobs) {
  const float* s = jobs.src[blockIdx.y];
  __half* d = jobs.dst[blockIdx.y];
  const long long n = jobs.n[blockIdx.y];
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = s[i];
    h_chk(v, bad);
    const __half h = __float2half_rn(v);
    d[i] = h;
    d[n + i] = __float2half_rn(v - __half2float(h));
  }
  h_flag(bad, jobs.status);
}

// ---- host ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn5)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn5 lh_encode_fn() {
  static EncodeTiledFn5 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn5)p;
  }
  return fn;
}

static int lh_passes(int N) { return (N + 127) / 128; }
bool linear_h_eligible(int K, int N) {
  if (K % 32 != 0 || K > 256 || N % 16 != 0 || N > 512) return false;
  const int p = lh_passes(N);
  if (N % p != 0 || (N / p) % 16 != 0) return false;
  const int kboxes = K / 32;
  const size_t w = (size_t)kboxes * 2 * N * 64, a1 = (size_t)2 * kboxes * LH_BM * 64;
  return w + a1 + (size_t)N * 4 + 2048 <= 225 * 1024;
}

int launch_ln_split_h(const float* x, const float* w, const float* b, void* planes, long long R, int K, float eps, int32_t* status, cudaStream_t s) {
  M2_REQUIRE(K % 32 == 0 && K <= 256 && (((uintptr_t)x) & 15) == 0, M2TTS_E_UNSUPPORTED, "ln_split_h: K=%d", K);
  long long blocks = (R + 31) / 32;            // 8 warps x 4 rows per block pass
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  M2_LAUNCH_PDL(M2TTS_STAGE_LAYERNORM, ln_split_h_kernel, (unsigned)blocks, 256, 0, s, x, w, b, (__half*)planes, R, K, eps, status);
  return M2TTS_OK;
}

int launch_w_split_h(const float* const* src, void* const* dst, const long long* n, int jobs, int32_t* status, cudaStream_t s) {
  M2_REQUIRE(jobs >= 1 && jobs <= 4, M2TTS_E_BADSHAPE, "w_split_h: 1..4 jobs");
  WSplitHJobs j{};
  j.status = status;
  long long mx = 1;
  for (int i = 0; i < jobs; ++i) { j.src[i] = src[i]; j.dst[i] = (__half*)dst[i]; j.n[i] = n[i]; if (n[i] > mx) mx = n[i]; }
  dim3 grid((unsigned)((mx + 255) / 256 > 256 ? 256 : (mx + 255) / 256), jobs);
  M2_LAUNCH(M2TTS_STAGE_PACK, w_split_h_kernel, grid, 256, 0, s, j);
  return M2TTS_OK;
}

extern long long* g_ws_prof;      // attention_tc.cu (m2tts_attention_set_prof)

// a_planes: fp16 [2][R][K]; w_planes: fp16 [2][N][K]
int launch_linear_h(const void* a_planes, const void* w_planes, const LinHParams& q, int stage, cudaStream_t s) {
  M2_REQUIRE(linear_h_eligible(q.K, q.N), M2TTS_E_UNSUPPORTED, "linear_h: K=%d N=%d not eligible", q.K, q.N);
  if (q.ln_x != nullptr) a_planes = w_planes;      // unused: the tensor map below is never dereferenced
  M2_REQUIRE((((uintptr_t)a_planes) & 15) == 0 && (((uintptr_t)w_planes) & 15) == 0 && (q.K & 7) == 0, M2TTS_E_BADSHAPE,
             "linear_h: misaligned operands");
  EncodeTiledFn5 enc = lh_encode_fn();
  M2_REQUIRE(enc != nullptr, M2TTS_E_CUDA, "linear_h: cuTensorMapEncodeTiled unavailable");
  LinHArgs a{};
  a.R = q.R; a.K = q.K; a.N = q.N;
  a.n_passes = lh_passes(q.N); a.np = q.N / a.n_passes; a.kboxes = q.K / 32;
  a.bias = q.bias; a.relu = q.relu; a.residual = q.residual; a.ldr = q.ldr; a.y = q.y; a.ldy = q.ldy;
  a.y_planes = (__half*)q.y_planes; a.qkvh = (__half*)q.qkvh; a.plane_stride = q.plane_stride;
  a.L = q.L; a.nh = q.nh; a.hd = q.hd; a.Lp = q.Lp; a.qscale = q.qscale; a.mode = q.mode;
  { static int ps = -2; if (ps == -2) ps = tools_env_int("M2TTS_LIN_PROF_STAGE", -1); a.prof = (ps >= 0 && ps == stage) ? g_ws_prof : nullptr; }
  { static int dbg = -1; if (dbg < 0) dbg = tools_env_int("M2TTS_LIN_DBG", 0); a.dbg = dbg; }
  a.status = q.status;
  a.ln_x = q.ln_x; a.ln_w = q.ln_w; a.ln_b = q.ln_b; a.ln_eps = q.ln_eps;
  M2_REQUIRE(q.ln_x == nullptr || (q.ln_w != nullptr && q.ln_b != nullptr && (((uintptr_t)q.ln_x) & 15) == 0 && q.K <= 96), M2TTS_E_BADSHAPE,
             "linear_h: LayerNorm producer needs weight, bias, 16-byte aligned rows and K <= 96");
  const size_t a_stage = (size_t)2 * a.kboxes * LH_BM * 64, w_bytes = (size_t)a.kboxes * 2 * a.N * 64;
  const bool stage3 = q.mode == 3 && a.np == q.nh * q.hd && q.L > 0 && q.R % q.L == 0 && q.plane_stride == (long long)(q.R / q.L) * q.nh * q.hd * q.Lp &&
                      (q.Lp & 7) == 0 && (((uintptr_t)q.qkvh) & 15) == 0;
  a.stg_bytes = (q.mode == 0 && (q.ldy & 3) == 0 && (((uintptr_t)q.y) & 15) == 0) || (q.mode == 1 && a.np % 32 == 0) || stage3 ? a.np * 512 : 0;
  // transposed QKV product (see the issuer): its A operand is 128 weight rows starting at a pass's first row, i.e. up to 128 - np rows
  // past the pass — for the last pass past the weight images. 8 KB of slack at the end of the allocation keep those reads (of lanes
  // nobody uses) inside this CTA's shared-memory window.
  static int use_t = -1;
  if (use_t < 0) use_t = tools_env_int("M2TTS_LIN_TPOSE", 1);
  const bool want_t = stage3 && use_t != 0 && a.np <= 128 && (q.hd & 7) == 0 && q.bias == nullptr && !q.relu;      // the QKV projection has no bias (components.py:51)
  size_t fixed = w_bytes + (size_t)a.stg_bytes + (size_t)a.N * 4 + 16 + 768 + 256 + 1024 + (want_t ? 8192 : 0);      // + bias, LayerNorm parameters, barriers, alignment
  int st = (int)((225 * 1024 - fixed) / a_stage);
  if (st < 1 && a.stg_bytes != 0) {      // no room for the staging tile: the epilogue stores directly
    fixed -= (size_t)a.stg_bytes;
    a.stg_bytes = 0;
    st = (int)((225 * 1024 - fixed) / a_stage);
  }
  a.stg2 = 0;
  if (a.stg_bytes != 0 && 225 * 1024 >= fixed + (size_t)a.stg_bytes + a_stage) {      // room for a second staging tile (and at least one A stage)
    a.stg2 = 1;
    fixed += (size_t)a.stg_bytes;
    st = (int)((225 * 1024 - fixed) / a_stage);
  }
  a.a_stages = st > 4 ? 4 : st;
  a.tpu = (q.mode == 3 && a.stg_bytes != 0) ? ceil_div(q.L, LH_BM) : 0;
  a.tpose = (a.tpu > 0 && want_t) ? 1 : 0;
  M2_REQUIRE(a.a_stages >= 1, M2TTS_E_UNSUPPORTED, "linear_h: operands do not fit shared memory");
  const size_t smem = fixed + (size_t)a.a_stages * a_stage;
  CUtensorMap ta, tw;
  const cuuint32_t estr[2] = {1, 1};
  if (a.tpu > 0) {      // per-utterance tiles: {K, L, plane * B + b}
    const cuuint64_t dims[3] = {(cuuint64_t)a.K, (cuuint64_t)q.L, (cuuint64_t)2 * (a.R / q.L)};
    const cuuint64_t strides[2] = {(cuuint64_t)a.K * 2, (cuuint64_t)q.L * a.K * 2};
    const cuuint32_t box[3] = {32u, (cuuint32_t)LH_BM, 1u};
    const cuuint32_t es3[3] = {1, 1, 1};
    const CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(a_planes), dims, strides, box, es3,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "linear_h: tensor map (A, per utterance) failed (%d)", (int)r);
  } else {
    const cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)2 * a.R};
    const cuuint64_t strides[1] = {(cuuint64_t)a.K * 2};
    const cuuint32_t box[2] = {32u, (cuuint32_t)LH_BM};
    const CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a_planes), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "linear_h: tensor map (A) failed (%d)", (int)r);
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)2 * a.N};
    const cuuint64_t strides[1] = {(cuuint64_t)a.K * 2};
    const cuuint32_t box[2] = {32u, (cuuint32_t)a.np};
    const CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w_planes), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "linear_h: tensor map (W) failed (%d)", (int)r);
  }
  CUtensorMap ty = ta, tr = ta;      // placeholders when the epilogue stores / loads directly
  a.res_tma = 0;
  if (a.stg_bytes != 0 && q.mode == 0 && q.residual != nullptr && (q.ldr & 3) == 0 && (((uintptr_t)q.residual) & 15) == 0) {
    const cuuint64_t dims[2] = {(cuuint64_t)a.N, (cuuint64_t)a.R};
    const cuuint64_t strides[1] = {(cuuint64_t)q.ldr * 4};
    const cuuint32_t box[2] = {16u, (cuuint32_t)LH_BM};
    const CUresult r = enc(&tr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(q.residual), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "linear_h: tensor map (residual) failed (%d)", (int)r);
    a.res_tma = 1;
  }
  if (a.stg_bytes != 0 && q.mode == 0) {
    const cuuint64_t dims[2] = {(cuuint64_t)a.N, (cuuint64_t)a.R};
    const cuuint64_t strides[1] = {(cuuint64_t)q.ldy * 4};
    const cuuint32_t box[2] = {16u, (cuuint32_t)LH_BM};
    const CUresult r = enc(&ty, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, q.y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "linear_h: tensor map (Y) failed (%d)", (int)r);
  } else if (a.stg_bytes != 0 && q.mode == 3) {      // operand planes as {positions (clipped at L), all d-rows}
    const cuuint64_t dims[2] = {(cuuint64_t)q.L, (cuuint64_t)6 * (a.R / q.L) * q.nh * q.hd};
    const cuuint64_t strides[1] = {(cuuint64_t)q.Lp * 2};
    const cuuint32_t box[2] = {a.tpose ? 64u : (cuuint32_t)LH_BM, (cuuint32_t)q.hd};
    const CUresult r = enc(&ty, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, q.qkvh, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           a.tpose ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "linear_h: tensor map (attention planes) failed (%d)", (int)r);
  } else if (a.stg_bytes != 0) {
    const cuuint64_t dims[3] = {(cuuint64_t)a.N, (cuuint64_t)a.R, 2};
    const cuuint64_t strides[2] = {(cuuint64_t)a.N * 2, (cuuint64_t)a.R * a.N * 2};
    const cuuint32_t box[3] = {32u, (cuuint32_t)LH_BM, 1u};
    const cuuint32_t es3[3] = {1, 1, 1};
    const CUresult r = enc(&ty, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, q.y_planes, dims, strides, box, es3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "linear_h: tensor map (Y planes) failed (%d)", (int)r);
  }
  M2_CUDA_OK(allow_smem(lin_h_kernel, smem));
  const int m_tiles = a.tpu > 0 ? (a.R / q.L) * a.tpu : ceil_div(a.R, LH_BM);
  const int grid = m_tiles < kNumSMs ? m_tiles : kNumSMs;
  M2_LAUNCH_PDL(stage, lin_h_kernel, grid, LH_THREADS, smem, s, ta, tw, ty, tr, a);
  return M2TTS_OK;
}

}  // namespace m2
