// voc_up_h.cu — ConvTranspose1d(CI, CI/2, kernel 8, stride 4, padding 2) + leaky_relu(0.1) of the two wide vocoder stages
// (components.py:225-241, the `ups` of SimplifiedHiFiGAN) on channel-last fp16 hi/lo planes, 16-bit split, fp32 accumulation,
// as a persistent tcgen05 kernel. Polyphase form: output sample t = 4 q + p has exactly two taps,
//   y[4q+p] = x[q] . W[:, :, p+2]  +  (p < 2 ? x[q-1] . W[:, :, p+6] : x[q+1] . W[:, :, p-2]),
// so for one phase p the stage is a GEMM over the input rows with two row-shifted A operands (a shift moves the start
// address of the K-major swizzled descriptor by whole rows) and ONE accumulator. Each CTA owns one (phase, COT output
// channels) pair with its weights resident in shared memory (64 KB for both shapes) and strides over 128-row tiles of
// the input; the input tile streams through a 3- or 4-stage ring of 64-channel k-blocks (hi + lo plane per stage); the output tile
// leaves through a swizzled staging buffer and two TMA stores (one per plane).
//   x : planes [2][B][L][CI] (the previous kernel wrote them)      y : planes [2][B][4L][CI/2]
// Warp roles: 0 TMA producer | 1 UMMA issuer (warp-collective) | 2.. epilogue warpgroups (thread = input row q, 16 channels).
#include "conv_tc.cuh"
#include "attention_tc.cuh"
#include <cuda_fp16.h>
#include <math.h>

namespace m2 {

struct UpHArgs {
  int B, L;
  int tiles_per_utt, total_tiles, n_tiles, co_tiles;
  const __half* wblob;                   // [n_tile][tap][k-block][hi rows ; lo rows][64] swizzled image
  const float* bias;
  __half* out_h; long long out_plane;
  int dbg_mode;                          // bring-up timing experiments: 1 no stores, 2 no UMMAs, 4 no TMA loads (results invalid)
  int32_t* status;                       // M2TTS_ST_FP16_RANGE when the output planes leave the fp16 range
};

template <int CI, int COT>
struct UhCfg {
  static constexpr int CO = CI / 2, KB = CI / 64;
  static constexpr int XR = 136, NQ = 128;
  static constexpr uint32_t XPL = XR * 128;                 // one plane (hi or lo) of one 64-channel k-block
  static constexpr uint32_t WKB = 2 * COT * 128;            // [hi rows ; lo rows] of one (tap, k-block)
  static constexpr uint32_t WBYTES = 2 * KB * WKB;
  static constexpr int STG_PLANES = WBYTES > 64 * 1024 ? 1 : 2;      // output staging: both planes at once, or one after the other
  static constexpr int SPL = WBYTES > 64 * 1024 ? 1 : 2;             // planes per ring stage (2: hi and lo UMMAs interleave on one B descriptor)
  static constexpr int NST = WBYTES > 64 * 1024 ? 4 : 3;
  static constexpr uint32_t STAGE = SPL * XPL;
  static constexpr int UNITS = KB * 2 / SPL;                // ring units per tile
  static constexpr uint32_t OFF_W = NST * STAGE;
  static constexpr uint32_t OPL = 128 * COT * 2;            // output staging: one plane of one tile, rows of COT halves (swizzled)
  static constexpr uint32_t OFF_STG = OFF_W + WBYTES;
  static constexpr uint32_t OFF_CONST = OFF_STG + STG_PLANES * OPL;
  static constexpr uint32_t OFF_BAR = OFF_CONST + 256;
  static constexpr uint32_t TOTAL = OFF_BAR + 256 + 1024;
  static constexpr int G = COT / 16;                        // epilogue warpgroups, 16 channels each
  static constexpr int THREADS = 64 + 128 * G;
  static constexpr uint32_t TMEM_COLS = 4 * COT;            // two accumulator buffers of (main | corr)
  static_assert(TOTAL <= 227 * 1024, "voc_up_h: shared memory");
  static_assert(COT == 32 || COT == 64, "voc_up_h: channel tile");
  static_assert(NST <= 6, "voc_up_h: barrier block");
};

__device__ __forceinline__ void uh_tma_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t uh_desc(uint32_t addr) {      // K-major, 128-byte swizzle, SBO = 1024 B
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void uh_mma_w(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

template <int CI, int COT>
__global__ void __launch_bounds__(UhCfg<CI, COT>::THREADS, 1)
voc_up_h_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y, const UpHArgs a, int* dbg) {
  pdl_launch_dependents();      // M2_LAUNCH_PDL: every access to another kernel's data follows a pdl_wait()
  using K = UhCfg<CI, COT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ct_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - ct_smem_u32(smem_raw));
  const uint32_t bars = sbase + K::OFF_BAR;
  // full[6] empty[6] acc_full[2] acc_free[2] w_full
  const uint32_t bar_f = bars, bar_e = bars + 48, bar_cf = bars + 96, bar_ce = bars + 112, bar_w = bars + 128, tmem_slot = bars + 136;
  float* bias_s = reinterpret_cast<float*>(gbase + K::OFF_CONST);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int ntile = blockIdx.x % a.n_tiles, first = blockIdx.x / a.n_tiles, cpg = gridDim.x / a.n_tiles;
  const int ph = ntile / a.co_tiles, co0 = (ntile % a.co_tiles) * COT;

  if (tid == 0) {
    for (int s = 0; s < K::NST; ++s) { ct_mbar_init(bar_f + 8 * s, 1); ct_mbar_init(bar_e + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { ct_mbar_init(bar_cf + 8 * s, 1); ct_mbar_init(bar_ce + 8 * s, 4 * K::G); }
    ct_mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_y) : "memory");
  }
  if (tid < COT) bias_s[tid] = a.bias[co0 + tid];
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(K::TMEM_COLS) : "memory");
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
        ct_bulk(sbase + K::OFF_W + off, reinterpret_cast<const uint8_t*>(a.wblob) + (size_t)ntile * K::WBYTES + off, 8192u, bar_w);
      pdl_wait();
      int u = 0;
      for (int g = first; g < a.total_tiles; g += cpg) {
        const int b = g / a.tiles_per_utt, q0 = (g % a.tiles_per_utt) * K::NQ;      // X row j <-> input row q0 - 1 + j
        for (int kp = 0; kp < K::UNITS; ++kp, ++u) {       // unit = (k-block, plane) or (k-block, both planes)
          const int st = u % K::NST, use = u / K::NST;
          if (use > 0) ct_wait(bar_e + 8 * st, (uint32_t)((use - 1) & 1), dbg, 1, u);
          if (a.dbg_mode & 4) { ct_arrive(bar_f + 8 * st); continue; }
          ct_expect_tx(bar_f + 8 * st, K::STAGE);
          const uint32_t dst = sbase + (uint32_t)st * K::STAGE;
          if (K::SPL == 1) {
            uh_tma_4d(dst, &tmap_x, (kp >> 1) * 64, q0 - 1, b, kp & 1, bar_f + 8 * st);
          } else {
            uh_tma_4d(dst, &tmap_x, kp * 64, q0 - 1, b, 0, bar_f + 8 * st);
            uh_tma_4d(dst + K::XPL, &tmap_x, kp * 64, q0 - 1, b, 1, bar_f + 8 * st);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== UMMA issuer: D[m, (main|corr, co)] = sum_{tap, k} A[m + shift(tap), k] W_tap[k, co] =====
    ct_wait(bar_w, 0, dbg, 2, 0);
    const uint32_t sW = sbase + K::OFF_W;
    const uint32_t id_2n = (1u << 4) | ((uint32_t)((2 * COT) >> 3) << 17) | (8u << 24), id_n = (1u << 4) | ((uint32_t)(COT >> 3) << 17) | (8u << 24);
    const uint32_t shift1 = ph < 2 ? 0u : 2u * 128u;                 // tap 1 reads x[q-1] (phases 0, 1) or x[q+1] (phases 2, 3)
    int it = 0, u = 0;
    for (int g = first; g < a.total_tiles; g += cpg, ++it) {
      const int ab = it & 1, ause = it >> 1;
      if (ause > 0) ct_wait(bar_ce + 8 * ab, (uint32_t)((ause - 1) & 1), dbg, 4, it);
      const uint32_t d = tmem_base + (uint32_t)ab * (2 * COT);
      for (int kp = 0; kp < K::UNITS; ++kp, ++u) {
        const int st = u % K::NST, use = u / K::NST;
        const int kb = K::SPL == 1 ? (kp >> 1) : kp;
        ct_wait(bar_f + 8 * st, (uint32_t)(use & 1), dbg, 3, u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sX = sbase + (uint32_t)st * K::STAGE;
        if (!(a.dbg_mode & 2)) {
          if (K::SPL == 2) {
#pragma unroll
            for (int tap = 0; tap < 2; ++tap)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t a_hi = sX + (tap == 0 ? 128u : shift1) + (uint32_t)ks * 32u;
                const uint64_t bd = uh_desc(sW + (uint32_t)(tap * K::KB + kb) * K::WKB + (uint32_t)ks * 32u);
                uh_mma_w(d, uh_desc(a_hi), bd, id_2n, (kp | tap | ks) ? 1u : 0u);        // A_hi x [W_hi ; W_lo]
                uh_mma_w(d, uh_desc(a_hi + K::XPL), bd, id_n, 1u);                       // A_lo x W_hi
              }
          } else if ((kp & 1) == 0) {
#pragma unroll
            for (int tap = 0; tap < 2; ++tap)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)                                           // A_hi x [W_hi ; W_lo]
                uh_mma_w(d, uh_desc(sX + (tap == 0 ? 128u : shift1) + (uint32_t)ks * 32u),
                         uh_desc(sW + (uint32_t)(tap * K::KB + kb) * K::WKB + (uint32_t)ks * 32u), id_2n, (kp | tap | ks) ? 1u : 0u);
          } else {
#pragma unroll
            for (int tap = 0; tap < 2; ++tap)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)                                           // A_lo x W_hi
                uh_mma_w(d, uh_desc(sX + (tap == 0 ? 128u : shift1) + (uint32_t)ks * 32u),
                         uh_desc(sW + (uint32_t)(tap * K::KB + kb) * K::WKB + (uint32_t)ks * 32u), id_n, 1u);
          }
        }
        ct_commit_w(bar_e + 8 * st);
      }
      ct_commit_w(bar_cf + 8 * ab);
    }
  } else {
    // ===== epilogue warpgroup eg: thread m = input row of the tile -> output row 4 (q0 + m) + phase, channels [16 eg, 16 eg + 16) =====
    // (several warpgroups: one warp per scheduler cannot hide the latency of its own dependent instructions)
    pdl_wait();
    const int eg = (warp - 2) >> 2;
    const int qtr = warp & 3;
    const int m = qtr * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qtr * 32) << 16);
    const float* bs = bias_s + eg * 16;
    // the tile's output goes through a swizzled staging buffer and leaves as two TMA stores (whole 2 COT-byte row pieces
    // at a row stride of 4 CO halves: thread-per-row global stores would touch 32 lines per instruction)
    uint8_t* stg = gbase + K::OFF_STG;
    constexpr uint32_t ORB = COT * 2;
    const uint32_t sw = ORB == 128 ? (uint32_t)(m & 7) : (uint32_t)((m >> 1) & 3);
    const bool leader = warp == 2 && lane == 0;
    float amax = 0.f;      // max |value| written as fp16 planes (NaN sticks): the fp16-range check
    int it = 0;
    for (int g = first; g < a.total_tiles; g += cpg, ++it) {
      const int ab = it & 1, ause = it >> 1;
      const int b = g / a.tiles_per_utt, q0 = (g % a.tiles_per_utt) * K::NQ;
      ct_wait(bar_cf + 8 * ab, (uint32_t)(ause & 1), dbg, 9, it);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t vm[16], vc[16];
      const uint32_t col = (uint32_t)(ab * 2 * COT + eg * 16);
      ct_ld16(t_lane + col, vm);
      ct_ld16(t_lane + col + COT, vc);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) ct_arrive(bar_ce + 8 * ab);
      uint4 hv[2], lv[2];
      {      // packed fp32 pairs (common.cuh): main + correction halves + bias, LeakyReLU, hi/lo split
        const uint64_t* bp = reinterpret_cast<const uint64_t*>(bs);
        uint32_t hw[8], lw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint64_t y = f2_lrelu01(f2_add(f2_add(f2_pack_u(vm[2 * j], vm[2 * j + 1]), f2_pack_u(vc[2 * j], vc[2 * j + 1])), bp[j]));
          h_split_pair(y, hw[j], lw[j], amax);
        }
        hv[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); hv[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
        lv[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]); lv[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
      }
      const uint32_t off0 = (uint32_t)m * ORB + ((((uint32_t)(eg * 2)) ^ sw) << 4), off1 = (uint32_t)m * ORB + ((((uint32_t)(eg * 2 + 1)) ^ sw) << 4);
      const uint32_t s0 = sbase + K::OFF_STG;
#pragma unroll
      for (int pass = 0; pass < 3 - K::STG_PLANES; ++pass) {      // one pass with both planes staged, or hi then lo through one buffer
        if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // earlier stores have read the staging buffer
        asm volatile("bar.sync 1, %0;" ::"n"(128 * K::G) : "memory");
        if (K::STG_PLANES == 2) {
          *reinterpret_cast<uint4*>(stg + off0) = hv[0]; *reinterpret_cast<uint4*>(stg + off1) = hv[1];
          *reinterpret_cast<uint4*>(stg + K::OPL + off0) = lv[0]; *reinterpret_cast<uint4*>(stg + K::OPL + off1) = lv[1];
        } else {
          *reinterpret_cast<uint4*>(stg + off0) = pass == 0 ? hv[0] : lv[0]; *reinterpret_cast<uint4*>(stg + off1) = pass == 0 ? hv[1] : lv[1];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 2, %0;" ::"n"(128 * K::G) : "memory");
        if (leader && !(a.dbg_mode & 1)) {
          if (K::STG_PLANES == 2) {
            asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                         ::"l"(&tmap_y), "r"(co0), "r"(ph), "r"(q0), "r"(b), "r"(0), "r"(s0) : "memory");
            asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                         ::"l"(&tmap_y), "r"(co0), "r"(ph), "r"(q0), "r"(b), "r"(1), "r"(s0 + K::OPL) : "memory");
          } else {
            asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                         ::"l"(&tmap_y), "r"(co0), "r"(ph), "r"(q0), "r"(b), "r"(pass), "r"(s0) : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    h_flag(h_amax_bad(amax), a.status);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(K::TMEM_COLS) : "memory");
  }
}

// weight image: [n_tile = phase * co_tiles + co_tile][tap][k-block][W_hi rows (COT) ; W_lo rows (COT)][64 k], 128-byte swizzle.
// ConvTranspose1d weight [CI][CO][8]: tap 0 <-> kernel index p + 2 (x[q]); tap 1 <-> p + 6 (x[q-1], p < 2) or p - 2 (x[q+1], p >= 2).
struct UhPackArgs { const float* w; __half* blob; int CI, COT; int32_t* status; };
__global__ void uh_wpack_kernel(UhPackArgs p) {
  const int CO = p.CI / 2, KB = p.CI / 64, co_tiles = CO / p.COT;
  const int total = 4 * CO * 2 * p.CI * 2;       // phases x co x taps x ci x (hi, lo)
  const uint32_t wkb = 2u * p.COT * 128u, wbytes = 2u * KB * wkb;
  bool bad = false;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int e = idx;
    const int k = e % p.CI; e /= p.CI;
    const int n = e % (2 * p.COT); e /= (2 * p.COT);
    const int tap = e % 2; const int ntile = e / 2;
    const int ph = ntile / co_tiles, co = (ntile % co_tiles) * p.COT + n % p.COT, lo = n / p.COT;
    const int kw = tap == 0 ? ph + 2 : (ph < 2 ? ph + 6 : ph - 2);
    const float v = p.w[((size_t)k * CO + co) * 8 + kw];
    h_chk(v, bad);
    const __half h = __float2half_rn(v);
    const int kb = k >> 6, kk = k & 63;
    const uint32_t off = (uint32_t)ntile * wbytes + (uint32_t)(tap * KB + kb) * wkb + (uint32_t)n * 128u +
                         ((((uint32_t)kk >> 3) ^ (uint32_t)(n & 7)) << 4) + (uint32_t)(kk & 7) * 2u;
    p.blob[off >> 1] = lo ? __float2half_rn(v - __half2float(h)) : h;
  }
  h_flag(bad, p.status);
}

typedef CUresult (*EncodeTiledFn9)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn9 uh_encode_fn() {
  static EncodeTiledFn9 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn9)p;
  }
  return fn;
}

static int g_uh_dbg = 0;
bool voc_up_h_eligible(int CI, int CO, int r) { return r == 4 && CO * 2 == CI && (CI == 256 || CI == 128 || CI == 64); }
static int uh_cot(int CI) { return CI == 64 ? 32 : 64; }      // output channels per CTA
size_t voc_up_h_wblob_bytes(int CI) { return (size_t)8 * CI * (CI / 2) * 2 * 2; }      // every weight once, hi + lo

template <int CI, int COT>
static int launch_up_h_t(const CUtensorMap& tmap, const CUtensorMap& tmap_y, UpHArgs a, int stage, cudaStream_t s) {
  using K = UhCfg<CI, COT>;
  a.co_tiles = K::CO / COT;
  a.n_tiles = 4 * a.co_tiles;
  int cpg = kNumSMs / a.n_tiles;
  if (cpg > a.total_tiles) cpg = a.total_tiles;
  M2_CUDA_OK(allow_smem(voc_up_h_kernel<CI, COT>, K::TOTAL));
  M2_LAUNCH_PDL(stage, (voc_up_h_kernel<CI, COT>), cpg * a.n_tiles, K::THREADS, K::TOTAL, s, tmap, tmap_y, a, debug_words_device());
  return M2TTS_OK;
}

// xh: fp16 hi/lo planes channel-last [2][B][L][CI] (x_plane elements apart) -> out_h planes [2][B][4L][CI/2] = lrelu(convT(x) + bias)
int launch_voc_up_h(const void* xh, long long x_plane, const float* w, const float* bias, void* wblob, void* out_h, long long out_plane,
                    int B, int CI, int L, int stage, int32_t* status, cudaStream_t s) {
  M2_REQUIRE(voc_up_h_eligible(CI, CI / 2, 4), M2TTS_E_UNSUPPORTED, "voc_up_h: CI=%d (256, 128 or 64)", CI);
  const int COT = uh_cot(CI);
  if (w != nullptr) {      // (re)write the weight image; w == nullptr: wblob already holds it
    M2_REQUIRE((((uintptr_t)wblob) & 15) == 0, M2TTS_E_BADSHAPE, "voc_up_h: misaligned weight image");
    UhPackArgs p{w, (__half*)wblob, CI, COT, status};
    M2_LAUNCH(M2TTS_STAGE_PACK, uh_wpack_kernel, ceil_div(16 * CI * (CI / 2), 256), 256, 0, s, p);
  }
  if (xh == nullptr) return M2TTS_OK;      // pack only
  M2_REQUIRE((((uintptr_t)xh) & 15) == 0 && (((uintptr_t)wblob) & 15) == 0 && (((uintptr_t)out_h) & 15) == 0 && (x_plane & 7) == 0 && (out_plane & 7) == 0,
             M2TTS_E_BADSHAPE, "voc_up_h: misaligned pointers");
  M2_REQUIRE(B > 0 && L > 0, M2TTS_E_BADSHAPE, "voc_up_h: B=%d L=%d", B, L);
  EncodeTiledFn9 enc = uh_encode_fn();
  M2_REQUIRE(enc != nullptr, M2TTS_E_CUDA, "voc_up_h: cuTensorMapEncodeTiled unavailable");
  CUtensorMap tmap;
  const cuuint64_t dims[4] = {(cuuint64_t)CI, (cuuint64_t)L, (cuuint64_t)B, 2};
  const cuuint64_t strides[3] = {(cuuint64_t)CI * 2, (cuuint64_t)L * CI * 2, (cuuint64_t)x_plane * 2};
  const cuuint32_t box[4] = {64u, 136u, 1u, 1u};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(xh), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "voc_up_h: cuTensorMapEncodeTiled failed (%d)", (int)r);
  UpHArgs a{};
  a.B = B; a.L = L; a.wblob = (const __half*)wblob; a.bias = bias; a.out_h = (__half*)out_h; a.out_plane = out_plane;
  a.tiles_per_utt = ceil_div(L, 128);
  a.total_tiles = B * a.tiles_per_utt;
  a.dbg_mode = g_uh_dbg; a.status = status;
  // output planes [2][B][4L][CO] seen as {CO, phase, q, B, plane}: one store = one phase of 128 input rows, COT channels
  CUtensorMap tmap_y;
  {
    const int CO = CI / 2;
    const cuuint64_t ydims[5] = {(cuuint64_t)CO, 4, (cuuint64_t)L, (cuuint64_t)B, 2};
    const cuuint64_t ystr[4] = {(cuuint64_t)CO * 2, (cuuint64_t)4 * CO * 2, (cuuint64_t)4 * L * CO * 2, (cuuint64_t)out_plane * 2};
    const cuuint32_t ybox[5] = {(cuuint32_t)COT, 1u, 128u, 1u, 1u};
    const cuuint32_t yes[5] = {1, 1, 1, 1, 1};
    const CUresult ry = enc(&tmap_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, out_h, ydims, ystr, ybox, yes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            COT == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(ry == CUDA_SUCCESS, M2TTS_E_CUDA, "voc_up_h: cuTensorMapEncodeTiled (output) failed (%d)", (int)ry);
  }
  if (CI == 64) return launch_up_h_t<64, 32>(tmap, tmap_y, a, stage, s);
  return CI == 256 ? launch_up_h_t<256, 64>(tmap, tmap_y, a, stage, s) : launch_up_h_t<128, 64>(tmap, tmap_y, a, stage, s);
}

void voc_up_h_set_debug(int m) { g_uh_dbg = m; }

}  // namespace m2

using namespace m2;

#ifdef M2TTS_TOOLS
// bring-up only (tools/up_h_prof.py): timing experiments with parts of the kernel switched off; results are invalid while set
extern "C" int m2tts_voc_up_h_set_debug(int mode) { voc_up_h_set_debug(mode); return M2TTS_OK; }
#endif

namespace {
__global__ void uh_join_planes_kernel(const __half* planes, long long n, float* y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __half2float(planes[i]) + __half2float(planes[n + i]);
}
}  // namespace

extern "C" size_t m2tts_conv_transpose_x4_h_workspace_bytes(int B, int CI, int L) {
  if (!voc_up_h_eligible(CI, CI / 2, 4) || B <= 0 || L <= 0) return 0;
  return align_up(voc_up_h_wblob_bytes(CI), 256) + align_up((size_t)B * L * CI * 4, 256) + align_up((size_t)B * L * CI * 8, 256) + 1024;
}

// y = leaky_relu(conv_transpose1d(x, w, b, stride 4, padding 2), 0.1) (components.py:225-241, one `ups` layer + its activation):
// x fp32 CHANNEL-LAST [B][L][CI], w [CI][CI/2][8] (state_dict layout), y fp32 channel-last [B][4L][CI/2]; CI in {128, 256}.
extern "C" int m2tts_conv_transpose_x4_h(const float* x, const float* w, const float* b, float* y, int B, int CI, int L,
                                         int32_t* status, void* workspace, size_t workspace_bytes, m2tts_stream_t stream) {
  M2_REQUIRE(x && w && b && y && workspace, M2TTS_E_NULLPTR, "conv_transpose_x4_h: null pointer");
  M2_REQUIRE(voc_up_h_eligible(CI, CI / 2, 4), M2TTS_E_UNSUPPORTED, "conv_transpose_x4_h: CI=%d (64, 128 or 256)", CI);
  Carver cv(workspace, workspace_bytes);
  __half* wblob = cv.take<__half>(voc_up_h_wblob_bytes(CI) / 2);
  const long long n_in = (long long)B * L * CI, n_out = (long long)B * 4 * L * (CI / 2);
  __half* xp = cv.take<__half>((size_t)2 * n_in);
  __half* yp = cv.take<__half>((size_t)2 * n_out);
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "conv_transpose_x4_h: workspace too small or misaligned");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_split_planes_h(x, xp, n_in, status, s);
  if (rc) return rc;
  if ((rc = launch_voc_up_h(xp, n_in, w, b, wblob, yp, n_out, B, CI, L, M2TTS_STAGE_VOC_UP, status, s))) return rc;
  M2_LAUNCH(M2TTS_STAGE_PACK, uh_join_planes_kernel, 1184, 256, 0, s, yp, n_out, y);
  return M2TTS_OK;
}
