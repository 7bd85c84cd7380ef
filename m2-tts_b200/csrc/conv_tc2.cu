// conv_tc2.cu — persistent, warp-specialised tap-GEMM for the vocoder convolutions (tcgen05 + TMEM + TMA).
// Same math and operand layouts as conv_tc.cu (see there), organised for throughput:
//   * one CTA per SM loops over (utterance, 120-position) tiles of ONE output-channel tile, so the packed
//     weight images stay RESIDENT in shared memory whenever they fit (every layer but the widest ones);
//   * warp 0 = TMA producer (activation boxes, and weight chunks when streaming), warp 1 = UMMA issuer,
//     warps 2-5 / 6-9 = two epilogue groups that take alternate tiles (= alternate TMEM accumulator buffers),
//     so two epilogues and the next tile's UMMAs are in flight at once; ring of activation stages between producer and issuer;
//   * epilogue: TMEM -> staging tile in shared memory -> out[t] = sum_tap D_tap[t + shift_tap] (+bias,
//     LeakyReLU, +residual) -> global (hi/lo planes for the next tensor-core layer, or plain fp32).
#include "conv_tc.cuh"
#include <math.h>

namespace m2 {

constexpr int P_THREADS = 320;            // producer warp, UMMA warp, 2 x 4 epilogue warps
constexpr int P_STAGING = 64 * 128 * 4;   // per epilogue group: 3 taps x 16 columns (or 2r x 8 channel rows) x 128 rows

__device__ __forceinline__ void p_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void p_epi_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }
__device__ __forceinline__ void p_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

struct PersistSmem {   // byte offsets from the 1024-aligned base, computed on the host
  uint32_t a_ring, w_region, staging, bars;
  uint32_t stage_bytes;      // bytes per ring stage (activations, + weights when streaming)
  uint32_t w_stage;          // bytes of one chunk's weight image (hi + lo)
  uint32_t total;
};

__global__ void __launch_bounds__(P_THREADS, 1)
tapgemm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const TapGemmArgs a, const PersistSmem L, int* dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ct_smem_u32(smem_raw) + 1023u) & ~1023u;
  float* stage_f = reinterpret_cast<float*>(smem_raw + (sbase - ct_smem_u32(smem_raw)) + L.staging);
  const int S = a.ring_stages;
  const uint32_t bar_full = sbase + L.bars;            // [S]
  const uint32_t bar_empty = bar_full + 8 * S;         // [S]
  const uint32_t bar_wfull = bar_empty + 8 * S;        // resident weights landed
  const uint32_t bar_accf = bar_wfull + 8;             // [2] accumulator buffer full
  const uint32_t bar_acce = bar_accf + 16;             // [2] accumulator buffer drained by the epilogue
  const uint32_t tmem_slot = bar_acce + 16;
  const uint32_t w_plane_bytes = (uint32_t)a.rows_total * 64u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // tile schedule: this CTA owns output-channel tile `ntile` and every cpg-th (utterance, position) tile
  const int ntile = blockIdx.x % a.n_tiles;
  const int first = blockIdx.x / a.n_tiles;
  const int cpg = gridDim.x / a.n_tiles;
  const int total_mt = a.B * a.m_tiles;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { ct_mbar_init(bar_full + 8 * s, 1); ct_mbar_init(bar_empty + 8 * s, 1); }
    ct_mbar_init(bar_wfull, 1);
    ct_mbar_init(bar_accf, 1); ct_mbar_init(bar_accf + 8, 1);
    ct_mbar_init(bar_acce, 4); ct_mbar_init(bar_acce + 8, 4);     // one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const float* wsrc = a.wblob + (size_t)ntile * a.n_chunks * (size_t)(2 * a.rows_total * 16);

  if (warp == 0) {
    if (lane == 0) {
      // ===== producer =====
      if (a.w_resident) {
        ct_expect_tx(bar_wfull, (uint32_t)a.n_chunks * L.w_stage);
        for (int c = 0; c < a.n_chunks; ++c)
          ct_bulk(sbase + L.w_region + (uint32_t)c * L.w_stage, wsrc + (size_t)c * (2 * a.rows_total * 16), L.w_stage, bar_wfull);
      }
      int it = 0;
      for (int j = first; j < total_mt; j += cpg) {
        const int b = j / a.m_tiles, start = (j % a.m_tiles) * CT_STEP - CT_HALO;
        for (int c = 0; c < a.n_chunks; ++c, ++it) {
          const int s = it % S;
          if (it >= S) ct_wait(bar_empty + 8 * s, (uint32_t)((it / S - 1) & 1), dbg, 1, c);
          const uint32_t sA = sbase + L.a_ring + (uint32_t)s * L.stage_bytes, full = bar_full + 8 * s;
          ct_expect_tx(full, CT_A_STAGE + (a.w_resident ? 0u : L.w_stage));
#pragma unroll
          for (int plane = 0; plane < 2; ++plane) {
            const int row = (plane * a.B + b) * a.CI + c * CT_CK;
#pragma unroll
            for (int x = 0; x < 4; ++x)
              ct_tma_2d(sA + (uint32_t)(plane * 4 + x) * CT_ABOX, &tmap_a, start + 32 * x, row, full);
          }
          if (!a.w_resident) ct_bulk(sA + CT_A_STAGE, wsrc + (size_t)c * (2 * a.rows_total * 16), L.w_stage, full);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== UMMA issuer =====
      if (a.w_resident) { ct_wait(bar_wfull, 0, dbg, 4, 0); }
      int it = 0, t = 0;
      for (int j = first; j < total_mt; j += cpg, ++t) {
        const int buf = t & 1;
        if (t >= 2) ct_wait(bar_acce + 8 * buf, (uint32_t)((t / 2 - 1) & 1), dbg, 5, t);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t dbase = tmem_base + (uint32_t)(buf * a.n_cols);
        for (int c = 0; c < a.n_chunks; ++c, ++it) {
          const int s = it % S;
          ct_wait(bar_full + 8 * s, (uint32_t)((it / S) & 1), dbg, 2, c);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sA = sbase + L.a_ring + (uint32_t)s * L.stage_bytes;
          const uint32_t sW = a.w_resident ? (sbase + L.w_region + (uint32_t)c * L.w_stage) : (sA + CT_A_STAGE);
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(a.tap_rows[tap] >> 3) << 17) |
                                   ((uint32_t)(CT_BM >> 4) << 24);
            const uint32_t wrow_off = (uint32_t)(a.tap_wrow[tap] >> 3) * 512u;
#pragma unroll
            for (int term = 0; term < 3; ++term) {           // hi*hi, hi*lo, lo*hi
              const uint32_t ap = (term == 2) ? 1u : 0u, wp = (term == 1) ? 1u : 0u;
#pragma unroll
              for (int ks = 0; ks < CT_CK / 8; ++ks) {
                const uint64_t ad = ct_desc(sA + (ap * 4) * CT_ABOX + ks * 1024u, CT_ABOX, 512u, 1u);
                const uint64_t bd = ct_desc(sW + wp * w_plane_bytes + wrow_off + ks * 256u, 128u, 512u, 0u);
                ct_mma(dbase + (uint32_t)a.tap_dcol[tap], ad, bd, idesc, (c | term | ks) ? 1u : 0u);
              }
            }
          }
          ct_commit(bar_empty + 8 * s);
        }
        ct_commit(bar_accf + 8 * buf);
      }
    }
  } else {
    // ===== epilogue: group g = (warp-2)/4 takes tiles t with t % 2 == g, i.e. always accumulator buffer g =====
    const int grp = (warp - 2) >> 2;
    const int qtr = warp & 3;                            // TMEM lane quarter this warp may access
    const int m = qtr * 32 + lane;                       // GEMM row = input position start + m
    const int buf = grp;
    float* stg = stage_f + grp * (P_STAGING / 4);
    const uint32_t t_lane = tmem_base + ((uint32_t)(qtr * 32) << 16) + (uint32_t)(buf * a.n_cols);
    const int co0 = ntile * a.co_tile;
    int t = 0;
    for (int j = first; j < total_mt; j += cpg, ++t) {
      if ((t & 1) != grp) continue;
      const int b = j / a.m_tiles, start = (j % a.m_tiles) * CT_STEP - CT_HALO;
      const int q = start + m;
      const bool own = (m >= CT_HALO) && (m < CT_BM - CT_HALO) && (q < a.L_in);
      ct_wait(bar_accf + 8 * buf, (uint32_t)((t / 2) & 1), dbg, 3, t);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

      if (a.r == 1) {
        const int s0 = a.tap_shift[0], s1 = a.tap_shift[1], s2 = a.tap_shift[2];
        for (int c0 = 0; c0 < a.co_tile; c0 += 16) {
          // residual loads first: 32 independent read-only loads in flight while the tile is staged
          float rsd[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) rsd[c] = 0.f;
          if (own && a.res_hi != nullptr) {
            const float* __restrict__ rh = a.res_hi + ((size_t)b * a.CO + co0 + c0) * a.Lp_res + q;
            const float* __restrict__ rl = a.res_lo + ((size_t)b * a.CO + co0 + c0) * a.Lp_res + q;
#pragma unroll
            for (int c = 0; c < 16; ++c) rsd[c] = __ldg(rh + (size_t)c * a.Lp_res) + __ldg(rl + (size_t)c * a.Lp_res);
          }
          uint32_t v[3][16];
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) p_ld16(t_lane + (uint32_t)(a.tap_dcol[tap] + c0), v[tap]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0 + 16 >= a.co_tile) {   // last TMEM read of this tile: hand the accumulator buffer back
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) p_arrive(bar_acce + 8 * buf);
          }
#pragma unroll
          for (int tap = 0; tap < 3; ++tap)
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) stg[((tap * 16 + jj) << 7) + m] = __uint_as_float(v[tap][jj]);
          p_epi_sync(grp);
          if (own) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const int co = co0 + c0 + c;
              float x = stg[((c) << 7) + m + s0] + stg[((16 + c) << 7) + m + s1] + stg[((32 + c) << 7) + m + s2] + __ldg(a.bias + co);
              if (a.act == 1) x = x > 0.f ? x : 0.1f * x;
              x += rsd[c];
              const size_t oo = ((size_t)b * a.CO + co) * a.Lp_out + q;
              if (a.out_lo != nullptr) { const float h = ct_hi(x); a.out_hi[oo] = h; a.out_lo[oo] = ct_hi(x - h); }
              else a.out_hi[oo] = x;
            }
          }
          p_epi_sync(grp);
        }
      } else {
        // transposed conv, r in {2,4}: D0 [0, r*ct) phase-major; D1 [r*ct, +r/2*ct) phases < r/2 (row q-1);
        // D2 next r/2*ct columns, phases >= r/2 (row q+1). Staging rows: group g of 8 channels:
        // g < r: D0 phase g; g >= r: D1/D2 phase g - r.
        const int ct = a.co_tile, r = a.r, hr = a.r / 2, groups = 2 * a.r;
        for (int c0 = 0; c0 < ct; c0 += 8) {
          uint32_t v[8][8];
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (g < groups) ct_ld8(t_lane + (uint32_t)(g * ct + c0), v[g]);   // D0|D1|D2 are contiguous: column = g*ct + c
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0 + 8 >= ct) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) p_arrive(bar_acce + 8 * buf);
          }
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (g < groups) {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) stg[((g * 8 + jj) << 7) + m] = __uint_as_float(v[g][jj]);
            }
          p_epi_sync(grp);
          if (own) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int co = co0 + c0 + jj;
              const float bv = __ldg(a.bias + co);
              float x[4];
#pragma unroll
              for (int p = 0; p < 4; ++p) {
                if (p < r) {
                  float tv = stg[((p * 8 + jj) << 7) + m] + bv;
                  tv += stg[(((r + p) * 8 + jj) << 7) + m + (p < hr ? -1 : 1)];
                  x[p] = tv > 0.f ? tv : 0.1f * tv;
                } else {
                  x[p] = 0.f;
                }
              }
              const size_t oo = ((size_t)b * a.CO + co) * a.Lp_out + (size_t)r * q;
              if (r == 4) {
                if (a.out_lo != nullptr) {
                  float h[4], l[4];
#pragma unroll
                  for (int p = 0; p < 4; ++p) { h[p] = ct_hi(x[p]); l[p] = ct_hi(x[p] - h[p]); }
                  *reinterpret_cast<float4*>(a.out_hi + oo) = make_float4(h[0], h[1], h[2], h[3]);
                  *reinterpret_cast<float4*>(a.out_lo + oo) = make_float4(l[0], l[1], l[2], l[3]);
                } else {
                  *reinterpret_cast<float4*>(a.out_hi + oo) = make_float4(x[0], x[1], x[2], x[3]);
                }
              } else {
                if (a.out_lo != nullptr) {
                  const float h0 = ct_hi(x[0]), h1 = ct_hi(x[1]);
                  *reinterpret_cast<float2*>(a.out_hi + oo) = make_float2(h0, h1);
                  *reinterpret_cast<float2*>(a.out_lo + oo) = make_float2(ct_hi(x[0] - h0), ct_hi(x[1] - h1));
                } else {
                  *reinterpret_cast<float2*>(a.out_hi + oo) = make_float2(x[0], x[1]);
                }
              }
            }
          }
          p_epi_sync(grp);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
  }
}

int launch_tapgemm_persistent(const CUtensorMap& tmap, TapGemmArgs& a, int stage, cudaStream_t s) {
  for (int j = 0; j < 3; ++j)
    M2_REQUIRE(a.tap_shift[j] >= -CT_HALO && a.tap_shift[j] <= CT_HALO, M2TTS_E_UNSUPPORTED,
               "conv_tc: tap shift %d exceeds the halo", a.tap_shift[j]);
  M2_REQUIRE(2 * a.n_cols <= 512, M2TTS_E_UNSUPPORTED, "conv_tc: %d accumulator columns do not double-buffer", a.n_cols);
  a.m_tiles = ceil_div(a.L_in, CT_STEP);
  a.tmem_cols = 32;
  while (a.tmem_cols < 2 * a.n_cols) a.tmem_cols <<= 1;
  PersistSmem L{};
  L.w_stage = 2u * (uint32_t)a.rows_total * 64u;
  const uint32_t staging = 2u * (uint32_t)P_STAGING;   // one staging tile per epilogue group
  const uint32_t budget = 226u * 1024u - 1024u /*alignment*/ - 256u /*barriers*/ - staging;
  const uint32_t w_all = (uint32_t)a.n_chunks * L.w_stage;
  if (w_all + 2u * CT_A_STAGE <= budget) {
    a.w_resident = 1;
    L.stage_bytes = CT_A_STAGE;
    int st = (int)((budget - w_all) / CT_A_STAGE);
    a.ring_stages = st > 6 ? 6 : st;
    L.a_ring = 0; L.w_region = (uint32_t)a.ring_stages * CT_A_STAGE;
    L.staging = L.w_region + ((w_all + 1023u) & ~1023u);
  } else {
    a.w_resident = 0;
    L.stage_bytes = CT_A_STAGE + ((L.w_stage + 1023u) & ~1023u);
    int st = (int)(budget / L.stage_bytes);
    M2_REQUIRE(st >= 2, M2TTS_E_UNSUPPORTED, "conv_tc: weight chunk of %u B does not fit a 2-stage ring", L.w_stage);
    a.ring_stages = st > 4 ? 4 : st;
    L.a_ring = 0; L.w_region = 0;
    L.staging = (uint32_t)a.ring_stages * L.stage_bytes;
  }
  L.bars = L.staging + staging;
  L.total = L.bars + 256u + 1024u;
  M2_REQUIRE(L.total <= 227u * 1024u, M2TTS_E_UNSUPPORTED, "conv_tc: %u B of shared memory", L.total);
  M2_CUDA_OK(allow_smem(tapgemm_persistent_kernel, L.total));
  const int total_mt = a.B * a.m_tiles;
  int cpg = kNumSMs / a.n_tiles;
  if (cpg < 1) cpg = 1;
  if (cpg > total_mt) cpg = total_mt;
  const int grid = cpg * a.n_tiles;
  M2_LAUNCH(stage, tapgemm_persistent_kernel, grid, P_THREADS, L.total, s, tmap, a, L, debug_words_device());
  return M2TTS_OK;
}

}  // namespace m2
