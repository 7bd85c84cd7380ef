// conv_tc2.cu — persistent, warp-specialised tap-GEMM kernel for the vocoder convolutions
// (tcgen05 + TMEM + TMA). Math and operand layouts: see conv_tc.cu. Organisation:
//   * one CTA per SM loops over (utterance, 120-position) tiles of ONE output-channel tile, so the packed
//     weight images stay RESIDENT in shared memory whenever they fit (every layer but the widest ones);
//   * activations stay PLAIN fp32 in HBM. warp 0 streams them with TMA (box = 32 positions x 16 channels,
//     128B swizzle / 32B atoms) into a raw ring. kind::tf32 UMMAs TRUNCATE fp32 operands (they ignore the low 13
//     mantissa bits; measured: tools/tf32_trunc_probe.py), so the raw tile IS the hi operand; warps 2-3 only
//     compute the lo tile x - trunc(x) (same swizzled addresses: a flat element-wise pass) and hand it to the
//     tensor pipe through fence.proxy.async + mbarrier; a raw stage is released by the UMMAs that read it; warp 1 issues the UMMAs (3 taps x {hi*hi, hi*lo, lo*hi} x 2 k-steps
//     per 16-channel chunk) into one of up to four TMEM accumulator buffers;
//   * warps 4-7 / 8-11 are two epilogue groups taking alternate tiles: TMEM -> staging tile in shared memory
//     -> out[t] = sum_tap D_tap[t + shift_tap] + bias (LeakyReLU, + residual) -> plain fp32 in HBM.
// HBM traffic is therefore the algorithmic minimum per layer (input once, residual once, output once).
#include "conv_tc.cuh"
#include <cuda_fp16.h>
#include <math.h>

namespace m2 {

constexpr int P_THREADS = 384;            // producer, UMMA issuer, 2 splitter warps, 2 x 4 epilogue warps
constexpr int P_SPLIT_WARPS = 2;

__device__ __forceinline__ void p_epi_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }

struct PersistSmem {   // byte offsets from the 1024-aligned base, computed on the host
  uint32_t raw_ring, split_ring, w_region, staging, bars;
  uint32_t split_stage_bytes;   // hi+lo tile (+ weight chunk when streaming)
  uint32_t w_stage;             // bytes of one chunk's weight image (hi + lo)
  uint32_t staging_group;       // bytes of one epilogue group's staging tile
  uint32_t total;
};

// Transposed-conv epilogue of one tile (R = upsampling rate, compile time so the accumulator arrays stay in registers).
// Only the D1/D2 groups cross rows; the D0 groups stay in registers. Row-major staging (one padded row of 8 R floats per
// thread) with 128-bit accesses, as in the Conv1d branch.
template <int R>
__device__ __forceinline__ void convT_epilogue(const TapGemmArgs& a, float* stg, uint32_t t_lane, uint32_t bar_release, int grp,
                                               int m, int lane, bool own, int b, int q, int co0) {
  constexpr int HR = R / 2, SP = 36;
  const int ct = a.co_tile;
  const int rm = max(m - 1, 0), rp = min(m + 1, CT_BM - 1);
  for (int c0 = 0; c0 < ct; c0 += 8) {
    uint32_t v[2 * R][8];
#pragma unroll
    for (int g = 0; g < 2 * R; ++g) ct_ld8(t_lane + (uint32_t)(g * ct + c0), v[g]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (c0 + 8 >= ct) {      // last TMEM read of this tile: hand the accumulator buffer back
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) ct_arrive(bar_release);
    }
    uint4* myrow = reinterpret_cast<uint4*>(stg + m * SP);
#pragma unroll
    for (int p = 0; p < R; ++p) {
      myrow[2 * p] = make_uint4(v[R + p][0], v[R + p][1], v[R + p][2], v[R + p][3]);
      myrow[2 * p + 1] = make_uint4(v[R + p][4], v[R + p][5], v[R + p][6], v[R + p][7]);
    }
    p_epi_sync(grp);
    if (own) {
      float nb[R][8];            // neighbour-row contribution per phase: row q-1 for p < R/2, row q+1 otherwise
#pragma unroll
      for (int p = 0; p < R; ++p) {
        const float4* np4 = reinterpret_cast<const float4*>(stg + (p < HR ? rm : rp) * SP + 8 * p);
        const float4 n0 = np4[0], n1 = np4[1];
        nb[p][0] = n0.x; nb[p][1] = n0.y; nb[p][2] = n0.z; nb[p][3] = n0.w;
        nb[p][4] = n1.x; nb[p][5] = n1.y; nb[p][6] = n1.z; nb[p][7] = n1.w;
      }
      if (a.out_cl == 2) {
        // channel-last fp16 hi/lo planes [2][B][L_out][CO] for the 16-bit split ResBlock kernel: 8 channels = one 16-byte
        // chunk per (output position, plane)
        bool bad = false;
#pragma unroll
        for (int p = 0; p < R; ++p) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float t0 = __uint_as_float(v[p][2 * e]) + __ldg(a.bias + co0 + c0 + 2 * e) + nb[p][2 * e];
            float t1 = __uint_as_float(v[p][2 * e + 1]) + __ldg(a.bias + co0 + c0 + 2 * e + 1) + nb[p][2 * e + 1];
            t0 = t0 > 0.f ? t0 : 0.1f * t0; t1 = t1 > 0.f ? t1 : 0.1f * t1;
            h_split2(t0, t1, hw[e], lw[e], bad);
          }
          __half* oh = reinterpret_cast<__half*>(a.out) + ((size_t)b * a.L_out + (size_t)R * q + p) * a.CO + co0 + c0;
          *reinterpret_cast<uint4*>(oh) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(oh + (size_t)a.B * a.L_out * a.CO) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        h_flag(bad, a.status);
      } else {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int co = co0 + c0 + jj;
        const float bv = __ldg(a.bias + co);
        float x[R];
#pragma unroll
        for (int p = 0; p < R; ++p) {
          const float tv = __uint_as_float(v[p][jj]) + bv + nb[p][jj];
          x[p] = tv > 0.f ? tv : 0.1f * tv;
        }
        float* op = a.out + ((size_t)b * a.CO + co) * a.Lp_out + (size_t)R * q;
        if (R == 4) *reinterpret_cast<float4*>(op) = make_float4(x[0], x[1], x[2], x[3]);
        else *reinterpret_cast<float2*>(op) = make_float2(x[0], x[1]);
      }
      }
    }
    p_epi_sync(grp);
  }
}

__global__ void __launch_bounds__(P_THREADS, 1)
tapgemm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const TapGemmArgs a, const PersistSmem L, int* dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ct_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - ct_smem_u32(smem_raw));     // generic pointer to the aligned base
  const int R = a.raw_stages, S = a.split_stages, NB = a.n_bufs;
  const uint32_t bar_rawf = sbase + L.bars;              // [R] raw tile landed (TMA tx)
  const uint32_t bar_rawe = bar_rawf + 8 * R;            // [R] raw tile consumed by both splitter warps
  const uint32_t bar_splf = bar_rawe + 8 * R;            // [S] hi/lo tile written by both splitter warps
  const uint32_t bar_sple = bar_splf + 8 * S;            // [S] hi/lo tile (and streamed weights) consumed by the UMMAs
  const uint32_t bar_wf = bar_sple + 8 * S;              // [S] streamed weight chunk landed / [0] resident weights landed
  const uint32_t bar_accf = bar_wf + 8 * S;              // [NB] accumulator buffer complete
  const uint32_t bar_acce = bar_accf + 8 * NB;           // [NB] accumulator buffer drained by the epilogue
  const uint32_t tmem_slot = bar_acce + 8 * NB;
  const uint32_t w_plane_bytes = (uint32_t)a.rows_total * 64u;

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  // tile schedule: this CTA owns output-channel tile `ntile` and every cpg-th (utterance, position) tile
  const int ntile = blockIdx.x % a.n_tiles;
  const int first = blockIdx.x / a.n_tiles;
  const int cpg = gridDim.x / a.n_tiles;
  const int total_mt = a.B * a.m_tiles;

  if (tid == 0) {
    for (int s = 0; s < R; ++s) { ct_mbar_init(bar_rawf + 8 * s, 1); ct_mbar_init(bar_rawe + 8 * s, 1); }
    for (int s = 0; s < S; ++s) {
      ct_mbar_init(bar_splf + 8 * s, 1); ct_mbar_init(bar_sple + 8 * s, 1); ct_mbar_init(bar_wf + 8 * s, 1);
    }
    for (int s = 0; s < NB; ++s) { ct_mbar_init(bar_accf + 8 * s, 1); ct_mbar_init(bar_acce + 8 * s, 4); }
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
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);      // warp-uniform for ptxas (uniform-register UMMA operands)
  const float* wsrc = a.wblob + (size_t)ntile * a.n_chunks * (size_t)(2 * a.rows_total * 16);

  if (warp == 0) {
    if (lane == 0) {
      // ===== producer: plain fp32 activation boxes (+ weight chunks when they are not resident) =====
      if (a.w_resident) {
        ct_expect_tx(bar_wf, (uint32_t)a.n_chunks * L.w_stage);
        for (int c = 0; c < a.n_chunks; ++c)
          ct_bulk(sbase + L.w_region + (uint32_t)c * L.w_stage, wsrc + (size_t)c * (2 * a.rows_total * 16), L.w_stage, bar_wf);
      }
      uint32_t rs = 0, rs_round = 0, ss = 0, ss_round = 0;
      const bool resident = a.w_resident != 0;
      for (int j = first; j < total_mt; j += cpg) {
        const int b = j / a.m_tiles, start = (j % a.m_tiles) * CT_STEP - CT_HALO;
        for (int c = 0; c < a.n_chunks; ++c) {
          if (rs_round > 0) ct_wait(bar_rawe + 8 * rs, (rs_round - 1) & 1u, dbg, 1, c);
          const uint32_t dst = sbase + L.raw_ring + rs * CT_RAW_STAGE, full = bar_rawf + 8 * rs;
          ct_expect_tx(full, CT_RAW_STAGE);
          const int row = b * a.CI + c * CT_CK;
#pragma unroll
          for (int x = 0; x < 4; ++x) ct_tma_2d(dst + (uint32_t)x * CT_ABOX, &tmap_a, start + 32 * x, row, full);
          if (!resident) {
            if (ss_round > 0) ct_wait(bar_sple + 8 * ss, (ss_round - 1) & 1u, dbg, 6, c);
            ct_expect_tx(bar_wf + 8 * ss, L.w_stage);
            ct_bulk(sbase + L.split_ring + ss * L.split_stage_bytes + CT_A_STAGE,
                    wsrc + (size_t)c * (2 * a.rows_total * 16), L.w_stage, bar_wf + 8 * ss);
          }
          if (++rs == (uint32_t)R) { rs = 0; ++rs_round; }
          if (++ss == (uint32_t)S) { ss = 0; ++ss_round; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===== UMMA issuer: the whole warp runs the loop with warp-uniform operands and one elected lane issues
      // (ct_mma_w). All taps read the SAME activation rows (their shifts are applied to accumulator rows in the
      // epilogue) and their weight rows / accumulator columns are contiguous, so one UMMA with N = rows_total covers
      // them: a chunk is 3 terms x 2 k-steps = 6 UMMAs of N up to 256 instead of 18 narrow ones. =====
      if (a.w_resident) ct_wait(bar_wf, 0, dbg, 4, 0);
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(a.rows_total >> 3) << 17) |
                             ((uint32_t)(CT_BM >> 4) << 24);
      // activations: MN-major, 128B swizzle / 32B atoms: LBO = next 32 positions (next box), SBO = next 4 channels
      const uint64_t a_tmpl = ct_desc(0u, CT_ABOX, 512u, 1u);
      // weights: K-major, no swizzle (8 x 16 B core matrices): LBO = next 4 channels, SBO = next 8 rows
      const uint64_t w_tmpl = ct_desc(0u, 128u, 512u, 0u);
      const uint32_t wplane16 = w_plane_bytes >> 4;
      const uint32_t n_chunks = (uint32_t)a.n_chunks;
      const bool resident = a.w_resident != 0;
      uint32_t ss = 0, ss_par = 0;            // lo-ring stage and its phase parity
      uint32_t rs = 0;                        // raw-ring stage (the hi operand)
      uint32_t buf = 0, buf_round = 0;        // accumulator buffer and how many times the ring of buffers wrapped
      for (int j = first; j < total_mt; j += cpg) {
        if (buf_round > 0) ct_wait(bar_acce + 8 * buf, (buf_round - 1) & 1u, dbg, 5, j);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t dbase = tmem_base + buf * (uint32_t)a.n_cols;
        for (uint32_t c = 0; c < n_chunks; ++c) {
          ct_wait(bar_splf + 8 * ss, ss_par, dbg, 2, (int)c);
          if (!resident) ct_wait(bar_wf + 8 * ss, ss_par, dbg, 7, (int)c);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sLo = sbase + L.split_ring + ss * L.split_stage_bytes;
          const uint32_t sHi = sbase + L.raw_ring + rs * CT_RAW_STAGE;
          const uint32_t sW = resident ? (sbase + L.w_region + c * L.w_stage) : (sLo + CT_A_STAGE);
          const uint64_t ahi0 = a_tmpl | (uint64_t)((sHi >> 4) & 0x3FFFu);
          const uint64_t alo0 = a_tmpl | (uint64_t)((sLo >> 4) & 0x3FFFu);
          const uint64_t bd0 = w_tmpl | (uint64_t)((sW >> 4) & 0x3FFFu);
#pragma unroll
          for (int term = 0; term < 3; ++term) {           // hi*hi, hi*lo, lo*hi
            const uint32_t ap = (term == 2) ? 1u : 0u, wp = (term == 1) ? 1u : 0u;
#pragma unroll
            for (int ks = 0; ks < CT_CK / 8; ++ks) {
              const uint64_t ad = (ap ? alo0 : ahi0) + (uint64_t)((ks * 1024u) >> 4);
              const uint64_t bd = bd0 + (uint64_t)(wp * wplane16 + ks * 16u);
              ct_mma_w(dbase, ad, bd, idesc, (c | (uint32_t)term | (uint32_t)ks) ? 1u : 0u);
            }
          }
          ct_commit_w(bar_sple + 8 * ss);
          ct_commit_w(bar_rawe + 8 * rs);           // the raw (= hi) stage is free once these UMMAs have read it
          if (++ss == (uint32_t)S) { ss = 0; ss_par ^= 1u; }
          if (++rs == (uint32_t)R) rs = 0;
        }
        ct_commit_w(bar_accf + 8 * buf);
        if (++buf == (uint32_t)NB) { buf = 0; ++buf_round; }
      }
    }
  } else if (warp < 2 + P_SPLIT_WARPS) {
    // ===== splitters: raw fp32 tile -> TF32 hi tile + lo tile (flat, the swizzle is address-preserving) =====
    // The two splitter warps take ALTERNATE chunks (a whole 8 KB tile each), so the fixed cost of a hand-off (two
    // barrier waits, the proxy fence, the arrive) of one warp overlaps the other warp's arithmetic.
    const int sw = warp - 2;
    uint32_t rs = 0, rs_par = 0, ss = 0, ss_round = 0, g = 0;
    for (int j = first; j < total_mt; j += cpg) {
      for (int c = 0; c < a.n_chunks; ++c, ++g) {
        if ((g & 1u) == (uint32_t)sw) {
          const bool prs = a.prof != nullptr && blockIdx.x == 0 && sw == 0 && lane == 0 && j == first + 8 * cpg && c < 8;
          if (prs) a.prof[512 + c * 4 + 0] = clock64();
          ct_wait(bar_rawf + 8 * rs, rs_par, dbg, 8, c);
          if (prs) a.prof[512 + c * 4 + 1] = clock64();
          if (ss_round > 0) ct_wait(bar_sple + 8 * ss, (ss_round - 1) & 1u, dbg, 9, c);
          if (prs) a.prof[512 + c * 4 + 2] = clock64();
          __syncwarp();
          const float4* src = reinterpret_cast<const float4*>(gbase + L.raw_ring + rs * CT_RAW_STAGE);
          float4* dlo = reinterpret_cast<float4*>(gbase + L.split_ring + ss * L.split_stage_bytes);
          // 512 float4 per tile, 16 per lane, in two batches of 8 loads (the compiler cannot hoist shared-memory loads
          // over shared-memory stores)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = src[h * 256 + k * 32 + lane];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float4 l;
              l.x = v[k].x - ct_hi(v[k].x); l.y = v[k].y - ct_hi(v[k].y); l.z = v[k].z - ct_hi(v[k].z); l.w = v[k].w - ct_hi(v[k].w);
              dlo[h * 256 + k * 32 + lane] = l;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor pipe
          __syncwarp();
          if (lane == 0) ct_arrive(bar_splf + 8 * ss);
          if (prs) a.prof[512 + c * 4 + 3] = clock64();
        }
        if (++rs == (uint32_t)R) { rs = 0; rs_par ^= 1u; }
        if (++ss == (uint32_t)S) { ss = 0; ++ss_round; }
      }
    }
  } else {
    // ===== epilogue: group g = (warp-4)/4 takes tiles t with t % 2 == g; accumulator buffer = t % NB =====
    const int grp = (warp - 4) >> 2;
    const int qtr = warp & 3;                            // TMEM lane quarter this warp may access
    const int m = qtr * 32 + lane;                       // GEMM row = input position start + m
    float* stg = reinterpret_cast<float*>(gbase + L.staging + (uint32_t)grp * L.staging_group);
    const int co0 = ntile * a.co_tile;
    int t = 0;
    for (int j = first; j < total_mt; j += cpg, ++t) {
      if ((t & 1) != grp) continue;
      const int buf = t % NB;
      const uint32_t t_lane = tmem_base + ((uint32_t)(qtr * 32) << 16) + (uint32_t)(buf * a.n_cols);
      const int b = j / a.m_tiles, start = (j % a.m_tiles) * CT_STEP - CT_HALO;
      const int q = start + m;
      const bool own = (m >= CT_HALO) && (m < CT_BM - CT_HALO) && (q < a.L_in);
      const bool pr = a.prof != nullptr && blockIdx.x == 0 && grp == 0 && m == 0 && t < 128;
      if (pr) a.prof[t * 4 + 0] = clock64();
      if (a.residual != nullptr) {
        // pull this warp's residual lines (one 128-B line per output channel) into L2 while the tile's UMMAs are
        // still running: the loads below then cost an L2 hit instead of a DRAM round trip per 16-channel round
        const int qw = start + qtr * 32;
        if (qw < a.L_in) {
          const float* rp = a.residual + ((size_t)b * a.CO + co0) * a.Lp_res + (qw < 0 ? 0 : qw);
          for (int c = lane; c < a.co_tile; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (size_t)c * a.Lp_res));
        }
      }
      ct_wait(bar_accf + 8 * buf, (uint32_t)((t / NB) & 1), dbg, 3, t);
      if (pr) a.prof[t * 4 + 1] = clock64();
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

      if (a.r == 1) {
        // Conv1d: out[m] = D0[m - d] + D1[m] + D2[m + d]. D1 stays in registers; D0 and D2 go through a row-major
        // staging tile (one padded row of 32 floats per thread) with 128-bit shared-memory accesses: 16 vector
        // instructions per 16 channels instead of 96 scalar ones (the epilogue was bound by shared-memory
        // instruction issue, measured with tools/tapgemm_prof.py).
        constexpr int SP = 36;                                  // staging row pitch in floats (conflict-free for 128-bit)
        const int r0 = min(max(m + a.tap_shift[0], 0), CT_BM - 1), r2 = min(max(m + a.tap_shift[2], 0), CT_BM - 1);
        for (int c0 = 0; c0 < a.co_tile; c0 += 16) {
          // residual loads first: 16 independent read-only loads in flight while the tile is staged
          float rsd[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) rsd[c] = 0.f;
          if (own && a.residual != nullptr) {
            const float* __restrict__ rp = a.residual + ((size_t)b * a.CO + co0 + c0) * a.Lp_res + q;
#pragma unroll
            for (int c = 0; c < 16; ++c) rsd[c] = __ldg(rp + (size_t)c * a.Lp_res);
          }
          uint32_t v[3][16];
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) ct_ld16(t_lane + (uint32_t)(a.tap_dcol[tap] + c0), v[tap]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0 + 16 >= a.co_tile) {   // last TMEM read of this tile: hand the accumulator buffer back
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) ct_arrive(bar_acce + 8 * buf);
          }
          uint4* myrow = reinterpret_cast<uint4*>(stg + m * SP);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            myrow[j4] = make_uint4(v[0][4 * j4], v[0][4 * j4 + 1], v[0][4 * j4 + 2], v[0][4 * j4 + 3]);
            myrow[4 + j4] = make_uint4(v[2][4 * j4], v[2][4 * j4 + 1], v[2][4 * j4 + 2], v[2][4 * j4 + 3]);
          }
          p_epi_sync(grp);
          if (own) {
            const float4* d0p = reinterpret_cast<const float4*>(stg + r0 * SP);
            const float4* d2p = reinterpret_cast<const float4*>(stg + r2 * SP + 16);
            const float4* bp = reinterpret_cast<const float4*>(a.bias + co0 + c0);
            float xo[16];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 d0 = d0p[j4], d2 = d2p[j4], bb = __ldg(bp + j4);
              xo[4 * j4 + 0] = d0.x + __uint_as_float(v[1][4 * j4 + 0]) + d2.x + bb.x;
              xo[4 * j4 + 1] = d0.y + __uint_as_float(v[1][4 * j4 + 1]) + d2.y + bb.y;
              xo[4 * j4 + 2] = d0.z + __uint_as_float(v[1][4 * j4 + 2]) + d2.z + bb.z;
              xo[4 * j4 + 3] = d0.w + __uint_as_float(v[1][4 * j4 + 3]) + d2.w + bb.w;
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float x = xo[c];
              if (a.act == 1) x = x > 0.f ? x : 0.1f * x;
              xo[c] = x + rsd[c];
            }
            if (a.out_cl == 2) {  // channel-last fp16 hi/lo planes for the 16-bit split fused stages (same bytes as fp32)
              __half* oh = reinterpret_cast<__half*>(a.out) + ((size_t)b * a.L_out + q) * a.CO + co0 + c0;
              __half* ol = oh + (size_t)a.B * a.L_out * a.CO;
              bool bad = false;
#pragma unroll
              for (int j8 = 0; j8 < 2; ++j8) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  h_split2(xo[8 * j8 + 2 * e], xo[8 * j8 + 2 * e + 1], hw[e], lw[e], bad);
                }
                *reinterpret_cast<uint4*>(oh + 8 * j8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(ol + 8 * j8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              }
              h_flag(bad, a.status);
            } else if (a.out_cl) {       // channel-last rows for the fused narrow stages: 64 contiguous bytes per thread
              float4* op = reinterpret_cast<float4*>(a.out + ((size_t)b * a.L_out + q) * a.CO + co0 + c0);
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) op[c4] = make_float4(xo[4 * c4], xo[4 * c4 + 1], xo[4 * c4 + 2], xo[4 * c4 + 3]);
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c) a.out[((size_t)b * a.CO + co0 + c0 + c) * a.Lp_out + q] = xo[c];
            }
          }
          p_epi_sync(grp);
        }
      } else {
        // transposed conv, r in {2,4}: accumulator columns D0 [0, r*ct) phase-major, D1 [r*ct, +r/2*ct) phases < r/2
        // (needs row q-1), D2 next r/2*ct columns, phases >= r/2 (needs row q+1); column = g*ct + channel with
        // g < r: D0 phase g, g >= r: D1/D2 phase g - r. Staging rows: g*8 + channel.
        if (a.r == 4) convT_epilogue<4>(a, stg, t_lane, bar_acce + 8 * buf, grp, m, lane, own, b, q, co0);
        else convT_epilogue<2>(a, stg, t_lane, bar_acce + 8 * buf, grp, m, lane, own, b, q, co0);
      }
      if (pr) a.prof[t * 4 + 2] = clock64();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
  }
}

static long long* g_tg_prof = nullptr;      // set only by the tools build (m2tts_tapgemm_set_prof)
int launch_tapgemm_persistent(const CUtensorMap& tmap, TapGemmArgs& a, int stage, cudaStream_t s) {
  a.prof = g_tg_prof;
  for (int j = 0; j < 3; ++j)
    M2_REQUIRE(a.tap_shift[j] >= -CT_HALO && a.tap_shift[j] <= CT_HALO, M2TTS_E_UNSUPPORTED,
               "conv_tc: tap shift %d exceeds the halo", a.tap_shift[j]);
  M2_REQUIRE(2 * a.n_cols <= 512, M2TTS_E_UNSUPPORTED, "conv_tc: %d accumulator columns do not double-buffer", a.n_cols);
  a.m_tiles = ceil_div(a.L_in, CT_STEP);
  a.n_bufs = 512 / a.n_cols;
  if (a.n_bufs > 4) a.n_bufs = 4;
  a.tmem_cols = 32;
  while (a.tmem_cols < a.n_bufs * a.n_cols) a.tmem_cols <<= 1;
  PersistSmem L{};
  L.w_stage = 2u * (uint32_t)a.rows_total * 64u;
  L.staging_group = (a.r == 1 ? 48u : (uint32_t)(2 * a.r * 8)) * 128u * 4u;   // conv: 3 taps x 16 columns; convT: 2r x 8 channels
  const uint32_t staging = 2u * L.staging_group;
  // Ring depths. The raw ring is what is in flight from HBM (8 KB per stage): it takes every stage the budget allows
  // (up to 8); the lo ring (consumed within a chunk's UMMAs) stays at two stages.
  const uint32_t fixed0 = 1024u /*alignment*/ + 512u /*barriers*/ + staging;
  const uint32_t budget0 = 227u * 1024u - fixed0;
  const uint32_t w_all = (uint32_t)a.n_chunks * L.w_stage;
  const uint32_t w_al = (w_all + 1023u) & ~1023u;
  if (w_al + 2u * CT_A_STAGE + 3u * CT_RAW_STAGE <= budget0) {
    a.w_resident = 1;
    L.split_stage_bytes = CT_A_STAGE;
    a.split_stages = 2;
    int rs = (int)((budget0 - w_al - 2u * CT_A_STAGE) / CT_RAW_STAGE);
    a.raw_stages = rs > 8 ? 8 : rs;
  } else {
    a.w_resident = 0;
    L.split_stage_bytes = CT_A_STAGE + ((L.w_stage + 1023u) & ~1023u);
    M2_REQUIRE(2u * L.split_stage_bytes + 3u * CT_RAW_STAGE <= budget0, M2TTS_E_UNSUPPORTED,
               "conv_tc: weight chunk of %u B does not fit a 2-stage ring", L.w_stage);
    a.split_stages = 2;
    int rs = (int)((budget0 - 2u * L.split_stage_bytes) / CT_RAW_STAGE);
    a.raw_stages = rs > 8 ? 8 : rs;
    // a third lo stage (the streamed weights travel with it) helps more than raw stages beyond 4
    if (a.raw_stages >= 4 + (int)((L.split_stage_bytes + CT_RAW_STAGE - 1) / CT_RAW_STAGE)) {
      a.split_stages = 3;
      a.raw_stages -= (int)((L.split_stage_bytes + CT_RAW_STAGE - 1) / CT_RAW_STAGE);
    }
  }
  L.raw_ring = 0;
  L.split_ring = (uint32_t)a.raw_stages * CT_RAW_STAGE;
  L.w_region = L.split_ring + (uint32_t)a.split_stages * L.split_stage_bytes;
  L.staging = L.w_region + (a.w_resident ? ((w_all + 1023u) & ~1023u) : 0u);
  L.bars = L.staging + staging;
  L.total = L.bars + 512u + 1024u;
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

// bring-up: device buffer of >= 576 int64 receiving epilogue / splitter timestamps of CTA 0 (NULL = off)
#ifdef M2TTS_TOOLS
extern "C" int m2tts_tapgemm_set_prof(long long* dev_buf) { m2::g_tg_prof = dev_buf; return M2TTS_OK; }
#endif
