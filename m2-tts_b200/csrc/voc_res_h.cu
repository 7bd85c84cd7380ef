// voc_res_h.cu — a whole LightweightResBlock (components.py:177-200) for C = 64 channels as ONE persistent tcgen05 kernel,
// channel-last, 16-bit split (fp16 hi/lo operands, fp32 accumulation; see voc_fused_h.cu / attention_h.cu):
//   u [B][L][C] (fp16 hi/lo planes)  ->  y = u + conv2(lrelu(conv1(u), 0.1))
// The two convolutions of a ResBlock used to be two tap-GEMM launches with the intermediate v = lrelu(conv1(u)) and the
// residual travelling through HBM (stage 1: 0.9 GB each way); here v never leaves the SM and u is read once.
// For C = 64 the fp16 weight images of both convolutions (96 KB) stay resident next to two input slots and V.
// (C = 128 would need 384 KB of weights: that stage keeps the tap-GEMM kernel.)
// Tile: 128 V rows (i <-> t = Ts + i), outputs i in [1, 127); X rows j <-> t = Ts - 1 + j (136 loaded, zero outside
// the utterance = the convolution's padding). A tap is the descriptor start address moved by whole rows.
// Warp roles: 0 TMA producer | 1 UMMA issuer (warp-collective) | 2-9 two epilogue warpgroups (one channel half each).
#include "conv_tc.cuh"
#include "attention_tc.cuh"
#include <cuda_fp16.h>
#include <math.h>

namespace m2 {

struct ResHArgs {
  int B, L;
  int tiles_per_utt, total_tiles;
  const __half* wblob;
  const float* bias1; const float* bias2;
  __half* out_h; long long out_plane;    // fp16 hi/lo planes [2][B][L][C], or
  float* out_f;                          // fp32 channel-last [B][L][C]
  long long* prof;                       // bring-up: phase timestamps of CTA 0's first epilogue warp (m2tts_attention_set_prof buffer)
  int dbg_nostore;                       // bring-up timing experiment (M2TTS_DBG_NOSTORE=1): skip the plane stores, results invalid
  int32_t* status;                       // M2TTS_ST_FP16_RANGE when V or the output planes leave the fp16 range
};

template <int C>
struct RhCfg {
  static constexpr int RB = C * 2;                   // bytes of a row (C = 64 -> 128 B, 128-byte swizzle)
  static constexpr int XR = 136, VR = 136, NOUT = 126;
  static constexpr uint32_t XPL = XR * RB, VPL = VR * RB;
  static constexpr uint32_t O_X = 0;                 // [2 slots][plane]
  static constexpr uint32_t O_V = 2 * 2 * XPL;
  static constexpr uint32_t OFF_W = O_V + 2 * VPL;
  static constexpr uint32_t WPART = 2 * C * RB;      // [W_hi rows ; W_lo rows] of one tap
  static constexpr uint32_t WBYTES = 6 * WPART;
  static constexpr uint32_t OFF_CONST = OFF_W + WBYTES;
  static constexpr uint32_t OFF_BAR = OFF_CONST + 1024;
  static constexpr uint32_t TOTAL = OFF_BAR + 128 + 1024;
#ifndef RH_NG
#define RH_NG 2
#endif
  static constexpr int NG = RH_NG;                   // epilogue warpgroups, C / NG channels each
  static constexpr int THREADS = 64 + 128 * NG;
  static constexpr int T_C1 = 0, T_C2 = 2 * C;
  static_assert(C == 64 || C == 32, "fused ResBlock: C in {32, 64}");
  static_assert(TOTAL <= 227 * 1024, "fused ResBlock: shared memory");
};

__device__ __forceinline__ void rh_tma_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
template <int RB>
__device__ __forceinline__ uint64_t rh_desc(uint32_t addr) {      // K-major rows of RB bytes with the RB-byte swizzle, SBO = 8 rows
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((8u * RB) >> 4) << 32) | (1ull << 46) | ((uint64_t)(RB == 128 ? 2 : 4) << 61);
}
template <int RB>
__device__ __forceinline__ uint32_t rh_swz(int row, int chunk) {   // byte offset of 16-byte chunk `chunk` of row `row`
  return RB == 128 ? (uint32_t)row * 128u + ((uint32_t)(chunk ^ (row & 7)) << 4) : (uint32_t)row * 64u + ((uint32_t)(chunk ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ void rh_mma_w(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void rh_split8(const float* x, uint4& hi, uint4& lo, bool& bad) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) h_split2(x[2 * e], x[2 * e + 1], h[e], l[e], bad);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void rh_join8(const uint4& hi, const uint4& lo, float* x) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h[e]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&l[e]));
    x[2 * e] = a.x + b.x; x[2 * e + 1] = a.y + b.y;
  }
}
__device__ __forceinline__ void rh_ld_sum16(uint32_t t_main, uint32_t t_corr, float* v) {
  uint32_t a[16], b[16];
  ct_ld16(t_main, a);
  ct_ld16(t_corr, b);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(a[j]) + __uint_as_float(b[j]);
}

template <int C>
__global__ void __launch_bounds__(RhCfg<C>::THREADS, 1)
voc_res_h_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y, const ResHArgs a, int* dbg) {
  pdl_launch_dependents();      // M2_LAUNCH_PDL: every access to another kernel's data follows a pdl_wait()
  using K = RhCfg<C>;
  constexpr int RB = K::RB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ct_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - ct_smem_u32(smem_raw));
  const uint32_t bars = sbase + K::OFF_BAR;
  // x_full[2] x_free[2] acc_c1 v_ready acc_c2 w_full
  const uint32_t bar_xf = bars, bar_xe = bars + 16, bar_c1 = bars + 32, bar_vr = bars + 40, bar_c2 = bars + 48, bar_w = bars + 56;
  const uint32_t tmem_slot = bars + 64;
  float* consts = reinterpret_cast<float*>(gbase + K::OFF_CONST);   // [0,C) b1 | [C,2C) b2

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int n_iter = (a.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int it) { return it * (int)gridDim.x + (int)blockIdx.x; };

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { ct_mbar_init(bar_xf + 8 * s, 1); ct_mbar_init(bar_xe + 8 * s, a.out_h != nullptr ? 1 : 4 * K::NG); }
    ct_mbar_init(bar_c1, 1); ct_mbar_init(bar_vr, 4 * K::NG); ct_mbar_init(bar_c2, 1); ct_mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  for (int i = tid; i < 2 * C; i += K::THREADS) consts[i] = i < C ? a.bias1[i] : a.bias2[i - C];
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);      // warp-uniform for ptxas (uniform-register UMMA operands)

  if (warp == 0) {
    if (lane == 0) {
      ct_expect_tx(bar_w, K::WBYTES);
      for (uint32_t off = 0; off < K::WBYTES; off += 8192u)
        ct_bulk(sbase + K::OFF_W + off, reinterpret_cast<const uint8_t*>(a.wblob) + off, 8192u, bar_w);
      pdl_wait();
      for (int it = 0; it < n_iter; ++it) {
        const int g = tile_of(it);
        if (g >= a.total_tiles) break;
        const int b = g / a.tiles_per_utt, k = g % a.tiles_per_utt;
        const int Ts = k * K::NOUT - 1;
        const int slot = it & 1, use = it >> 1;
        if (use > 0) ct_wait(bar_xe + 8 * slot, (uint32_t)((use - 1) & 1), dbg, 1, it);     // EPI3 of the tile two back has read its residual
        ct_expect_tx(bar_xf + 8 * slot, 2 * K::XPL);
        const uint32_t dst = sbase + K::O_X + (uint32_t)slot * 2 * K::XPL;
        rh_tma_4d(dst, &tmap_x, 0, Ts - 1, b, 0, bar_xf + 8 * slot);
        rh_tma_4d(dst + K::XPL, &tmap_x, 0, Ts - 1, b, 1, bar_xf + 8 * slot);
      }
    }
  } else if (warp == 1) {
    // ===== UMMA issuer: D[i, (main|corr, co)] = sum_tap A[i + tap, :] W_tap; per (tap, k-step) A_hi x [W_hi;W_lo], A_lo x W_hi =====
    ct_wait(bar_w, 0, dbg, 2, 0);
    const uint32_t sW = sbase + K::OFF_W;
    const uint32_t id_2c = (1u << 4) | ((uint32_t)((2 * C) >> 3) << 17) | (8u << 24), id_c = (1u << 4) | ((uint32_t)(C >> 3) << 17) | (8u << 24);
    auto issue_conv = [&](uint32_t sA, uint32_t plane_bytes, int conv) {
      const uint32_t d = tmem_base + (uint32_t)(conv == 0 ? K::T_C1 : K::T_C2);
#pragma unroll
      for (int tap = 0; tap < 3; ++tap)
#pragma unroll
        for (int ks = 0; ks < C / 16; ++ks) {
          const uint32_t a_hi = sA + (uint32_t)tap * RB + (uint32_t)ks * 32u;
          const uint64_t bd = rh_desc<RB>(sW + (uint32_t)(conv * 3 + tap) * K::WPART + (uint32_t)ks * 32u);
          rh_mma_w(d, rh_desc<RB>(a_hi), bd, id_2c, (tap | ks) ? 1u : 0u);
          rh_mma_w(d, rh_desc<RB>(a_hi + plane_bytes), bd, id_c, 1u);
        }
    };
    if (tile_of(0) < a.total_tiles) {
      ct_wait(bar_xf, 0u, dbg, 3, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_conv(sbase + K::O_X, K::XPL, 0);
      ct_commit_w(bar_c1);
    }
    for (int it = 0; it < n_iter; ++it) {
      if (tile_of(it) >= a.total_tiles) break;
      ct_wait(bar_vr, (uint32_t)(it & 1), dbg, 4, it);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_conv(sbase + K::O_V, K::VPL, 1);
      ct_commit_w(bar_c2);
      if (it + 1 < n_iter && tile_of(it + 1) < a.total_tiles) {      // conv1 of the next tile runs under this tile's last epilogue
        const int slot = (it + 1) & 1, use = (it + 1) >> 1;
        ct_wait(bar_xf + 8 * slot, (uint32_t)(use & 1), dbg, 3, it + 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_conv(sbase + K::O_X + (uint32_t)slot * 2 * K::XPL, K::XPL, 0);
        ct_commit_w(bar_c1);
      }
    }
  } else {
    // ===== epilogue warpgroup g: thread m owns TMEM lane m and channels [CG g, CG g + CG) =====
    pdl_wait();
    const int g = (warp - 2) >> 2;
    const int qtr = warp & 3;
    const int m = qtr * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qtr * 32) << 16);
    uint8_t* Vb = gbase + K::O_V;
    const float* b1 = consts, *b2 = consts + C;
    constexpr int CG = C / K::NG;
    const int cg0 = g * CG;
    float amax = 0.f;      // max |value| written as fp16 planes (NaN sticks): the fp16-range check
    for (int it = 0; it < n_iter; ++it) {
      const int gt = tile_of(it);
      if (gt >= a.total_tiles) break;
      const uint32_t par = (uint32_t)(it & 1);
      const int b = gt / a.tiles_per_utt, k = gt % a.tiles_per_utt;
      const int Ts = k * K::NOUT - 1;
      const int t = Ts + m;
      const bool inside = t >= 0 && t < a.L;
      const uint8_t* Xb = gbase + K::O_X + (uint32_t)(it & 1) * 2 * K::XPL;

      const bool pt = a.prof != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0 && it < 48;
      if (pt) a.prof[it * 8 + 0] = clock64();
      // ---- EPI2: conv1 accumulator -> V = lrelu(. + bias), zero outside the utterance, fp16 hi/lo rows ----
      ct_wait(bar_c1, par, dbg, 9, it);
      if (pt) a.prof[it * 8 + 1] = clock64();
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        const float keep = inside ? 1.f : 0.f;
        const bool all_keep = __all_sync(0xffffffffu, inside);
#pragma unroll
        for (int c0 = cg0; c0 < cg0 + CG; c0 += 16) {
          uint64_t v[8];
          ct_ld_sum16_pairs(t_lane + (uint32_t)(K::T_C1 + c0), t_lane + (uint32_t)(K::T_C1 + C + c0), b1 + c0, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = f2_lrelu01(v[j]);
          if (!all_keep) {                      // only the tiles at the ends of an utterance have rows to zero
            const uint64_t k2 = f2_pack(keep, keep);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = f2_mul(v[j], k2);
          }
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) h_split_pair(v[j], hi[j], lo[j], amax);
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8) {
            const uint32_t off = rh_swz<RB>(m + 1, (c0 >> 3) + j8);
            *reinterpret_cast<uint4*>(Vb + off) = make_uint4(hi[4 * j8], hi[4 * j8 + 1], hi[4 * j8 + 2], hi[4 * j8 + 3]);
            *reinterpret_cast<uint4*>(Vb + K::VPL + off) = make_uint4(lo[4 * j8], lo[4 * j8 + 1], lo[4 * j8 + 2], lo[4 * j8 + 3]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) ct_arrive(bar_vr);
      if (pt) a.prof[it * 8 + 2] = clock64();

      // planes mode: the previous tile's output left through TMA stores that read its input slot; once they have read it the
      // slot goes back to the producer (the leader has nothing else to do while conv2 runs)
      const bool leader = warp == 2 && lane == 0;
      if (a.out_h != nullptr && leader && it > 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        ct_arrive(bar_xe + 8 * ((it - 1) & 1));
      }

      // ---- EPI3: conv2 accumulator + bias + u (row m + 1 of the input tile) -> output ----
      ct_wait(bar_c2, par, dbg, 10, it);
      if (pt) a.prof[it * 8 + 3] = clock64();
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = cg0; c0 < cg0 + CG; c0 += 16) {
        uint64_t y[8];
        ct_ld_sum16_pairs(t_lane + (uint32_t)(K::T_C2 + c0), t_lane + (uint32_t)(K::T_C2 + C + c0), b2 + c0, y);
#pragma unroll
        for (int j8 = 0; j8 < 2; ++j8) {      // + the residual u (hi + lo, exact)
          const uint32_t off = rh_swz<RB>(m + 1, (c0 >> 3) + j8);
          const uint4 uh = *reinterpret_cast<const uint4*>(Xb + off), ul = *reinterpret_cast<const uint4*>(Xb + K::XPL + off);
          y[4 * j8] = f2_add(y[4 * j8], h_join_pair(uh.x, ul.x));
          y[4 * j8 + 1] = f2_add(y[4 * j8 + 1], h_join_pair(uh.y, ul.y));
          y[4 * j8 + 2] = f2_add(y[4 * j8 + 2], h_join_pair(uh.z, ul.z));
          y[4 * j8 + 3] = f2_add(y[4 * j8 + 3], h_join_pair(uh.w, ul.w));
        }
        if (a.out_h != nullptr) {
          // planes: y overwrites this thread's own residual chunks in the input slot (same swizzled addresses); the rows
          // leave as two TMA stores below (thread-per-row global stores touch 32 lines per instruction: 0.2 ms)
          // (rows 0 and 127 are computed from the two V rows outside the tile — never written — and are not stored: they
          // must not raise the range flag)
          uint8_t* Xw = gbase + K::O_X + (uint32_t)(it & 1) * 2 * K::XPL;
          float amax_y = 0.f;
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) h_split_pair(y[j], hi[j], lo[j], amax_y);
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8) {
            const uint32_t off = rh_swz<RB>(m + 1, (c0 >> 3) + j8);
            *reinterpret_cast<uint4*>(Xw + off) = make_uint4(hi[4 * j8], hi[4 * j8 + 1], hi[4 * j8 + 2], hi[4 * j8 + 3]);
            *reinterpret_cast<uint4*>(Xw + K::XPL + off) = make_uint4(lo[4 * j8], lo[4 * j8 + 1], lo[4 * j8 + 2], lo[4 * j8 + 3]);
          }
          if (m >= 1 && m <= K::NOUT) amax = amax_nan3(amax, amax_y, 0.f);
        } else if (inside && m >= 1 && m < 1 + K::NOUT) {
          const size_t o = ((size_t)b * a.L + t) * C + c0;
          uint64_t* op = reinterpret_cast<uint64_t*>(a.out_f + o);
#pragma unroll
          for (int j = 0; j < 8; ++j) op[j] = y[j];
        }
      }
      if (pt) a.prof[it * 8 + 4] = clock64();
      if (a.out_h != nullptr) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(128 * K::NG) : "memory");
        if (leader && !a.dbg_nostore) {      // output rows m = 1 .. NOUT sit in slot rows 2 .. NOUT + 1; TMA clips positions >= L
          const uint32_t s0 = sbase + K::O_X + (uint32_t)(it & 1) * 2 * K::XPL + 2u * RB;
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                       ::"l"(&tmap_y), "r"(0), "r"(k * K::NOUT), "r"(b), "r"(0), "r"(s0) : "memory");
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                       ::"l"(&tmap_y), "r"(0), "r"(k * K::NOUT), "r"(b), "r"(1), "r"(s0 + K::XPL) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (pt) a.prof[it * 8 + 5] = clock64();
        continue;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) ct_arrive(bar_xe + 8 * (it & 1));      // this input slot (the residual) has been read
    }
    if (a.out_h != nullptr && warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    h_flag(h_amax_bad(amax), a.status);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// weight image: per (conv, tap) a part [W_hi rows ; W_lo rows] x C, K-major rows with the 128-byte swizzle
struct RhPackArgs { const float* w1; const float* w2; __half* blob; int C; int32_t* status; };
__global__ void rh_wpack_kernel(RhPackArgs p) {
  const int C = p.C, RB = 2 * C;
  const int per = 2 * C * C, total = 6 * per;
  bool bad = false;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int part = idx / per, e = idx - part * per;
    const int n = e / C, k = e % C;
    const int co = n % C, lo = n / C;
    const float* w = part < 3 ? p.w1 : p.w2;
    const float v = w[((size_t)co * C + k) * 3 + (part % 3)];
    h_chk(v, bad);
    const __half h = __float2half_rn(v);
    const uint32_t off = (uint32_t)part * (2 * C * RB) + (uint32_t)n * RB + ((((uint32_t)k >> 3) ^ (RB == 128 ? (uint32_t)(n & 7) : (uint32_t)((n >> 1) & 3))) << 4) + (uint32_t)(k & 7) * 2u;
    p.blob[off >> 1] = lo ? __float2half_rn(v - __half2float(h)) : h;
  }
  h_flag(bad, p.status);
}

typedef CUresult (*EncodeTiledFn7)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn7 rh_encode_fn() {
  static EncodeTiledFn7 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn7)p;
  }
  return fn;
}

extern long long* g_ws_prof;      // attention_tc.cu (m2tts_attention_set_prof)
bool voc_res_h_eligible(int C, int dil) { return (C == 64 || C == 32) && dil == 1; }
size_t voc_res_h_wblob_bytes(int C) { return C == 64 ? RhCfg<64>::WBYTES : (C == 32 ? RhCfg<32>::WBYTES : 0); }

// uh: fp16 hi/lo planes channel-last [2][B][L][C] (u_plane elements apart); output planes (out_h/out_plane) or fp32 CL (out_f)
template <int C>
static int launch_voc_res_h_t(const void* uh, long long u_plane, const float* w1, const float* b1, const float* w2, const float* b2, void* wblob,
                     void* out_h, long long out_plane, float* out_f, int B, int L, int stage, int32_t* status, cudaStream_t s) {
  if (w1 != nullptr) {      // (re)write the weight image; w1 == nullptr: wblob already holds it
    M2_REQUIRE(w2 != nullptr && (((uintptr_t)wblob) & 15) == 0, M2TTS_E_BADSHAPE, "voc_res_h: pack arguments");
    RhPackArgs p{w1, w2, (__half*)wblob, C, status};
    M2_LAUNCH(M2TTS_STAGE_PACK, rh_wpack_kernel, ceil_div(12 * C * C, 256), 256, 0, s, p);
  }
  if (uh == nullptr) return M2TTS_OK;      // pack only
  M2_REQUIRE((((uintptr_t)uh) & 15) == 0 && (((uintptr_t)wblob) & 15) == 0 && (u_plane & 7) == 0 && (out_plane & 7) == 0, M2TTS_E_BADSHAPE,
             "voc_res_h: misaligned pointers");
  M2_REQUIRE(B > 0 && L > 0 && (out_h != nullptr || out_f != nullptr), M2TTS_E_BADSHAPE, "voc_res_h: B=%d L=%d", B, L);
  using K = RhCfg<C>;
  EncodeTiledFn7 enc = rh_encode_fn();
  M2_REQUIRE(enc != nullptr, M2TTS_E_CUDA, "voc_res_h: cuTensorMapEncodeTiled unavailable");
  CUtensorMap tmap;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B, 2};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)L * C * 2, (cuuint64_t)u_plane * 2};
  const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)K::XR, 1u, 1u};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(uh), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, (C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "voc_res_h: cuTensorMapEncodeTiled failed (%d)", (int)r);
  ResHArgs a{};
  a.B = B; a.L = L; a.wblob = (const __half*)wblob; a.bias1 = b1; a.bias2 = b2;
  a.out_h = (__half*)out_h; a.out_plane = out_plane;
  a.prof = g_ws_prof;
  { static int ns = -1; if (ns < 0) ns = tools_env_int("M2TTS_DBG_NOSTORE", 0) == 1 ? 1 : 0; a.dbg_nostore = ns; }
  a.out_f = out_f; a.status = status;
  a.tiles_per_utt = ceil_div(L, K::NOUT);
  a.total_tiles = B * a.tiles_per_utt;
  const int grid = a.total_tiles < kNumSMs ? a.total_tiles : kNumSMs;
  M2_CUDA_OK(allow_smem(voc_res_h_kernel<C>, K::TOTAL));
  CUtensorMap tmap_y = tmap;      // placeholder when the output is fp32
  if (out_h != nullptr) {
    const cuuint64_t ydims[4] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B, 2};
    const cuuint64_t ystr[3] = {(cuuint64_t)C * 2, (cuuint64_t)L * C * 2, (cuuint64_t)out_plane * 2};
    const cuuint32_t ybox[4] = {(cuuint32_t)C, (cuuint32_t)K::NOUT, 1u, 1u};
    const cuuint32_t yes[4] = {1, 1, 1, 1};
    const CUresult ry = enc(&tmap_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, out_h, ydims, ystr, ybox, yes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            (C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B), CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(ry == CUDA_SUCCESS, M2TTS_E_CUDA, "voc_res_h: cuTensorMapEncodeTiled (output) failed (%d)", (int)ry);
  }
  M2_LAUNCH_PDL(stage, voc_res_h_kernel<C>, grid, K::THREADS, K::TOTAL, s, tmap, tmap_y, a, debug_words_device());
  return M2TTS_OK;
}

int launch_voc_res_h(const void* uh, long long u_plane, const float* w1, const float* b1, const float* w2, const float* b2, void* wblob,
                     void* out_h, long long out_plane, float* out_f, int B, int C, int L, int stage, int32_t* status, cudaStream_t s) {
  M2_REQUIRE(C == 64 || C == 32, M2TTS_E_UNSUPPORTED, "voc_res_h: C=%d (32 or 64)", C);
  if (C == 64) return launch_voc_res_h_t<64>(uh, u_plane, w1, b1, w2, b2, wblob, out_h, out_plane, out_f, B, L, stage, status, s);
  return launch_voc_res_h_t<32>(uh, u_plane, w1, b1, w2, b2, wblob, out_h, out_plane, out_f, B, L, stage, status, s);
}

}  // namespace m2

using namespace m2;

extern "C" size_t m2tts_resblock_fused_h_workspace_bytes(int B, int C, int L) {
  if ((C != 64 && C != 32) || B <= 0 || L <= 0) return 0;
  return align_up(voc_res_h_wblob_bytes(C), 256) + align_up((size_t)B * L * C * 4, 256) + 512;    // weight image + input planes
}

// y = x + conv2(leaky_relu(conv1(x), 0.1)) (components.py:196-200), x / y fp32 CHANNEL-LAST [B][L][C], C = 64, k = 3, dilation 1.
extern "C" int m2tts_resblock_fused_h(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* y,
                                      int B, int C, int L, int32_t* status, void* workspace, size_t workspace_bytes, m2tts_stream_t stream) {
  M2_REQUIRE(x && w1 && b1 && w2 && b2 && y && workspace, M2TTS_E_NULLPTR, "resblock_fused_h: null pointer");
  M2_REQUIRE(C == 64 || C == 32, M2TTS_E_UNSUPPORTED, "resblock_fused_h: C=%d (32 or 64)", C);
  Carver cv(workspace, workspace_bytes);
  __half* wblob = cv.take<__half>(voc_res_h_wblob_bytes(C) / 2);
  const long long n = (long long)B * L * C;
  __half* planes = cv.take<__half>((size_t)2 * n);
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "resblock_fused_h: workspace too small or misaligned");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_split_planes_h(x, planes, n, status, s);
  if (rc) return rc;
  return launch_voc_res_h(planes, n, w1, b1, w2, b2, wblob, nullptr, 0, y, B, C, L, M2TTS_STAGE_VOC_RES1, status, s);
}
