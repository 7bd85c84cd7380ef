// voc_conv_h.cu — Conv1d(C, C, k = 3, dilation 1) for C = 128, channel-last, 16-bit split (fp16 hi/lo operands, fp32
// accumulation), as a persistent tcgen05 kernel — the two convolutions of the widest ResBlock (components.py:181-200),
// whose weights (384 KB as fp16 hi/lo) are too large for the fused ResBlock kernel of voc_res_h.cu:
//   y = act(conv(x) + bias) (+ residual)
// x, residual: fp16 hi/lo planes channel-last [2][B][L][C]; y: the same planes, or fp32 CHANNEL-FIRST [B][C][Lp]
// (what the next upsampling tap-GEMM reads). Against the tap-GEMM kernel there is no splitter (the producer already
// wrote the operand planes), K = 16 per UMMA, one accumulator for all taps (a tap is a row shift of the descriptor
// start address) and therefore a plain thread-per-row epilogue.
// Each CTA owns 64 output channels (weights resident: 96 KB; a kind::f16 UMMA costs ~60 cycles in this kernel whatever its
// N <= 128, so N is made as large as the weights allow) and strides over 128-row tiles; the input tile streams through a
// 3-stage ring of 64-channel k-blocks (hi + lo plane per stage) and two TMEM accumulator buffers keep UMMA and epilogue
// of consecutive tiles overlapped.
// Warp roles: 0 TMA producer | 1 UMMA issuer (warp-collective) | 2-17 four epilogue warpgroups (thread = row, 16 channels).
#include "conv_tc.cuh"
#include "attention_tc.cuh"
#include <cuda_fp16.h>
#include <math.h>

namespace m2 {

struct ConvHArgs {
  int B, L, Lp_out;
  int CO;                                // output channels = row length of the output (and residual) planes
  int ks1;                               // 16-channel k-steps of the second k-block (4, or fewer when CI < 128: the input conv)
  int nkb;                               // 64-channel k-blocks that carry data (1 when CI == 64: the second one is neither loaded nor multiplied)
  int tiles_per_utt, total_tiles, n_tiles;
  const __half* wblob;                   // [n_tile][tap][hi rows ; lo rows][C] swizzled image
  const float* bias;
  int act;                               // 0 none, 1 leaky_relu(0.1)
  const __half* res_h; long long res_plane;     // optional residual planes
  __half* out_h; long long out_plane;    // planes out, or
  float* out_cf;                         // fp32 channel-first [B][C][Lp_out]
  int32_t* status;                       // M2TTS_ST_FP16_RANGE when the output planes leave the fp16 range
};

constexpr int CH_C = 128, CH_NT = 64;                // channels, output channels per CTA
constexpr int CH_RB = 128;                           // bytes of one k-block row (64 halves)
constexpr int CH_KB = CH_C / 64;                     // k-blocks per row
constexpr int CH_XR = 136, CH_NOUT = 128;
constexpr uint32_t CH_XKB = CH_XR * CH_RB;           // one k-block of one plane of the input tile
constexpr uint32_t CH_STAGE = 2 * CH_XKB;            // ring stage: hi + lo plane of one k-block
constexpr int CH_NST = 3;
constexpr uint32_t CH_WKB = 2 * CH_NT * CH_RB;       // [hi rows ; lo rows] of one (tap, k-block)
constexpr uint32_t CH_WBYTES = 3 * CH_KB * CH_WKB;   // 96 KB
constexpr uint32_t CH_OFF_W = CH_NST * CH_STAGE;
constexpr uint32_t CH_OFF_CONST = CH_OFF_W + CH_WBYTES;
constexpr uint32_t CH_OFF_BAR = CH_OFF_CONST + 256;
constexpr uint32_t CH_TOTAL = CH_OFF_BAR + 128 + 1024;
constexpr int CH_G = CH_NT / 16;                     // epilogue warpgroups, 16 channels each
constexpr int CH_THREADS = 64 + 128 * CH_G;
constexpr uint32_t CH_TMEM = 4 * CH_NT;               // two accumulator buffers of (main | corr)
static_assert(CH_TOTAL <= 227 * 1024, "voc_conv_h: shared memory");

__device__ __forceinline__ void ch_tma_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t ch_desc(uint32_t addr) {      // K-major, 128-byte swizzle, SBO = 1024 B
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void ch_mma_w(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(CH_THREADS, 1)
voc_conv_h_kernel(const __grid_constant__ CUtensorMap tmap_x, const ConvHArgs a, int* dbg) {
  pdl_launch_dependents();      // M2_LAUNCH_PDL: every access to another kernel's data follows a pdl_wait()
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ct_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - ct_smem_u32(smem_raw));
  const uint32_t bars = sbase + CH_OFF_BAR;
  // x_full[3] x_free[3] acc_full[2] acc_free[2] w_full
  const uint32_t bar_xf = bars, bar_xe = bars + 24, bar_cf = bars + 48, bar_ce = bars + 64, bar_w = bars + 80, tmem_slot = bars + 88;
  float* bias_s = reinterpret_cast<float*>(gbase + CH_OFF_CONST);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int ntile = blockIdx.x % a.n_tiles, first = blockIdx.x / a.n_tiles, cpg = gridDim.x / a.n_tiles;
  const int co0 = ntile * CH_NT;

  if (tid == 0) {
    for (int s = 0; s < CH_NST; ++s) { ct_mbar_init(bar_xf + 8 * s, 1); ct_mbar_init(bar_xe + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { ct_mbar_init(bar_cf + 8 * s, 1); ct_mbar_init(bar_ce + 8 * s, 4 * CH_G); }
    ct_mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  if (tid < CH_NT) bias_s[tid] = a.bias[co0 + tid];
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(CH_TMEM) : "memory");
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
      ct_expect_tx(bar_w, CH_WBYTES);
      for (uint32_t off = 0; off < CH_WBYTES; off += 8192u)
        ct_bulk(sbase + CH_OFF_W + off, reinterpret_cast<const uint8_t*>(a.wblob) + (size_t)ntile * CH_WBYTES + off, 8192u, bar_w);
      pdl_wait();
      int u = 0;
      for (int g = first; g < a.total_tiles; g += cpg) {
        const int b = g / a.tiles_per_utt, k = g % a.tiles_per_utt;
        const int Ts = k * CH_NOUT;                    // output row 0 of the tile; X row j <-> t = Ts - 1 + j
        for (int kb = 0; kb < a.nkb; ++kb, ++u) {
          const int st = u % CH_NST, use = u / CH_NST;
          if (use > 0) ct_wait(bar_xe + 8 * st, (uint32_t)((use - 1) & 1), dbg, 1, u);
          ct_expect_tx(bar_xf + 8 * st, CH_STAGE);
          const uint32_t dst = sbase + (uint32_t)st * CH_STAGE;
          ch_tma_4d(dst, &tmap_x, kb * 64, Ts - 1, b, 0, bar_xf + 8 * st);
          ch_tma_4d(dst + CH_XKB, &tmap_x, kb * 64, Ts - 1, b, 1, bar_xf + 8 * st);
        }
      }
    }
  } else if (warp == 1) {
    // ===== UMMA issuer: D[i, (main|corr, co)] = sum_tap A[i + tap, :] W_tap =====
    ct_wait(bar_w, 0, dbg, 2, 0);
    const uint32_t sW = sbase + CH_OFF_W;
    const uint32_t id_2n = (1u << 4) | ((uint32_t)((2 * CH_NT) >> 3) << 17) | (8u << 24), id_n = (1u << 4) | ((uint32_t)(CH_NT >> 3) << 17) | (8u << 24);
    int it = 0, u = 0;
    for (int g = first; g < a.total_tiles; g += cpg, ++it) {
      const int slot = it & 1, ause = it >> 1;
      if (ause > 0) ct_wait(bar_ce + 8 * slot, (uint32_t)((ause - 1) & 1), dbg, 4, it);     // accumulator buffer `slot` drained
      const uint32_t d = tmem_base + (uint32_t)slot * (2 * CH_NT);
      for (int kb = 0; kb < a.nkb; ++kb, ++u) {
        const int st = u % CH_NST, use = u / CH_NST;
        ct_wait(bar_xf + 8 * st, (uint32_t)(use & 1), dbg, 3, u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sX = sbase + (uint32_t)st * CH_STAGE;
        const int nks = kb == 0 ? 4 : a.ks1;
#pragma unroll
        for (int tap = 0; tap < 3; ++tap)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks >= nks) break;
            const uint32_t a_hi = sX + (uint32_t)tap * CH_RB + (uint32_t)ks * 32u;
            const uint64_t bd = ch_desc(sW + (uint32_t)(tap * CH_KB + kb) * CH_WKB + (uint32_t)ks * 32u);
            ch_mma_w(d, ch_desc(a_hi), bd, id_2n, (kb | tap | ks) ? 1u : 0u);       // A_hi x [W_hi ; W_lo]
            ch_mma_w(d, ch_desc(a_hi + CH_XKB), bd, id_n, 1u);                      // A_lo x W_hi
          }
        ct_commit_w(bar_xe + 8 * st);
      }
      ct_commit_w(bar_cf + 8 * slot);
    }
  } else {
    // ===== epilogue warpgroup eg: thread m = output row of the tile, channels [16 eg, 16 eg + 16) =====
    pdl_wait();
    // (several warpgroups: one warp per scheduler cannot hide the latency of its own dependent instructions)
    const int eg = (warp - 2) >> 2;
    const int qtr = warp & 3;
    const int m = qtr * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qtr * 32) << 16);
    const float* bs = bias_s + eg * 16;
    float amax = 0.f;      // max |value| written as fp16 planes (NaN sticks): the fp16-range check
    int it = 0;
    for (int g = first; g < a.total_tiles; g += cpg, ++it) {
      const int slot = it & 1, use = it >> 1;
      const int b = g / a.tiles_per_utt, k = g % a.tiles_per_utt;
      const int t = k * CH_NOUT + m;
      const bool valid = t < a.L;
      const size_t o = ((size_t)b * a.L + t) * a.CO + co0 + eg * 16;
      // residual planes of this row (32 B + 32 B): requested before the accumulator wait (fetching them a whole tile ahead
      // was measured and changes nothing: with the residual the kernel moves 1.34 GB and sits at ~55 % of the HBM peak)
      uint4 rh[2], rl[2];
      if (a.res_h != nullptr && valid) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          rh[j] = __ldg(reinterpret_cast<const uint4*>(a.res_h + o) + j);
          rl[j] = __ldg(reinterpret_cast<const uint4*>(a.res_h + a.res_plane + o) + j);
        }
      }
      ct_wait(bar_cf + 8 * slot, (uint32_t)(use & 1), dbg, 9, it);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t vm[16], vc[16];
      ct_ld16(t_lane + (uint32_t)(slot * 2 * CH_NT + eg * 16), vm);
      ct_ld16(t_lane + (uint32_t)(slot * 2 * CH_NT + CH_NT + eg * 16), vc);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) ct_arrive(bar_ce + 8 * slot);
      if (!valid) continue;
      // packed fp32 pairs (common.cuh): main + correction halves + bias, LeakyReLU, residual, hi/lo split
      uint64_t y[8];
      const uint64_t* bp = reinterpret_cast<const uint64_t*>(bs);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y[j] = f2_add(f2_add(f2_pack_u(vm[2 * j], vm[2 * j + 1]), f2_pack_u(vc[2 * j], vc[2 * j + 1])), bp[j]);
        if (a.act == 1) y[j] = f2_lrelu01(y[j]);
      }
      if (a.res_h != nullptr) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          y[4 * j] = f2_add(y[4 * j], h_join_pair(rh[j].x, rl[j].x));
          y[4 * j + 1] = f2_add(y[4 * j + 1], h_join_pair(rh[j].y, rl[j].y));
          y[4 * j + 2] = f2_add(y[4 * j + 2], h_join_pair(rh[j].z, rl[j].z));
          y[4 * j + 3] = f2_add(y[4 * j + 3], h_join_pair(rh[j].w, rl[j].w));
        }
      }
      if (a.out_h != nullptr) {
        uint32_t hw[8], lw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h_split_pair(y[j], hw[j], lw[j], amax);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          *(reinterpret_cast<uint4*>(a.out_h + o) + j) = make_uint4(hw[4 * j], hw[4 * j + 1], hw[4 * j + 2], hw[4 * j + 3]);
          *(reinterpret_cast<uint4*>(a.out_h + a.out_plane + o) + j) = make_uint4(lw[4 * j], lw[4 * j + 1], lw[4 * j + 2], lw[4 * j + 3]);
        }
      } else {
        float* op = a.out_cf + ((size_t)b * a.CO + co0 + eg * 16) * a.Lp_out + t;     // channel-first: coalesced across the warp's rows
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float y0, y1;
          f2_unpack(y[j], y0, y1);
          op[(size_t)(2 * j) * a.Lp_out] = y0;
          op[(size_t)(2 * j + 1) * a.Lp_out] = y1;
        }
      }
    }
    h_flag(h_amax_bad(amax), a.status);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CH_TMEM) : "memory");
  }
}

// weight image: [n_tile][tap][k-block][W_hi rows (32) ; W_lo rows (32)][64 k], K-major rows with the 128-byte swizzle
struct ChPackArgs { const float* w; __half* blob; int CI, CO; int32_t* status; };
__global__ void ch_wpack_kernel(ChPackArgs p) {
  const int total = p.CO * CH_C * 3 * 2;
  bool bad = false;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int e = idx;
    const int k = e % CH_C; e /= CH_C;
    const int n = e % (2 * CH_NT); e /= (2 * CH_NT);
    const int tap = e % 3; const int ntile = e / 3;
    const int lo = n / CH_NT, co = ntile * CH_NT + n % CH_NT;
    const float v = k < p.CI ? p.w[((size_t)co * p.CI + k) * 3 + tap] : 0.f;      // k >= CI: zero padding
    h_chk(v, bad);
    const __half h = __float2half_rn(v);
    const int kb = k >> 6, kk = k & 63;
    const uint32_t off = (uint32_t)ntile * CH_WBYTES + (uint32_t)(tap * CH_KB + kb) * CH_WKB + (uint32_t)n * CH_RB +
                         ((((uint32_t)kk >> 3) ^ (uint32_t)(n & 7)) << 4) + (uint32_t)(kk & 7) * 2u;
    p.blob[off >> 1] = lo ? __float2half_rn(v - __half2float(h)) : h;
  }
  h_flag(bad, p.status);
}

typedef CUresult (*EncodeTiledFn8)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn8 ch_encode_fn() {
  static EncodeTiledFn8 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn8)p;
  }
  return fn;
}

bool voc_conv_h_eligible(int C, int dil) { return C == CH_C && dil == 1; }
bool voc_conv_h_io_eligible(int CI, int CO) { return CI >= 64 && CI <= CH_C && CI % 16 == 0 && CO >= CH_NT && CO % CH_NT == 0; }
size_t voc_conv_h_wblob_bytes(int C) { return C % CH_NT == 0 ? (size_t)(C / CH_NT) * CH_WBYTES : 0; }      // C = output channels

// xh: fp16 hi/lo planes channel-last [2][B][L][CI] (x_plane apart), 64 <= CI <= 128 (channels CI..127 are zero-filled by TMA and
// have zero weights); CO output channels (multiple of 64); residual planes [2][B][L][CO] optional; output planes or fp32 channel-first.
int launch_voc_conv_h(const void* xh, long long x_plane, const float* w, const float* bias, void* wblob, const void* res_h,
                      long long res_plane, void* out_h, long long out_plane, float* out_cf, int Lp_out, int B, int CI, int CO, int L, int act,
                      int stage, int32_t* status, cudaStream_t s) {
  M2_REQUIRE(voc_conv_h_io_eligible(CI, CO), M2TTS_E_UNSUPPORTED, "voc_conv_h: CI=%d CO=%d", CI, CO);
  if (w != nullptr) {      // (re)write the weight image; w == nullptr: wblob already holds it
    M2_REQUIRE((((uintptr_t)wblob) & 15) == 0, M2TTS_E_BADSHAPE, "voc_conv_h: misaligned weight image");
    ChPackArgs p{w, (__half*)wblob, CI, CO, status};
    M2_LAUNCH(M2TTS_STAGE_PACK, ch_wpack_kernel, ceil_div(CO * CH_C * 6, 256), 256, 0, s, p);
  }
  if (xh == nullptr) return M2TTS_OK;      // pack only
  const int C = CI;
  M2_REQUIRE((((uintptr_t)xh) & 15) == 0 && (((uintptr_t)wblob) & 15) == 0 && (x_plane & 7) == 0 && (out_plane & 7) == 0 && (res_plane & 7) == 0,
             M2TTS_E_BADSHAPE, "voc_conv_h: misaligned pointers");
  M2_REQUIRE(B > 0 && L > 0 && (out_h != nullptr || out_cf != nullptr), M2TTS_E_BADSHAPE, "voc_conv_h: B=%d L=%d", B, L);
  EncodeTiledFn8 enc = ch_encode_fn();
  M2_REQUIRE(enc != nullptr, M2TTS_E_CUDA, "voc_conv_h: cuTensorMapEncodeTiled unavailable");
  CUtensorMap tmap;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B, 2};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)L * C * 2, (cuuint64_t)x_plane * 2};
  const cuuint32_t box[4] = {64u, (cuuint32_t)CH_XR, 1u, 1u};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(xh), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "voc_conv_h: cuTensorMapEncodeTiled failed (%d)", (int)r);
  ConvHArgs a{};
  a.B = B; a.L = L; a.Lp_out = Lp_out; a.wblob = (const __half*)wblob; a.bias = bias; a.act = act;
  a.res_h = (const __half*)res_h; a.res_plane = res_plane; a.out_h = (__half*)out_h; a.out_plane = out_plane; a.out_cf = out_cf;
  a.CO = CO; a.ks1 = (CI - 64 + 15) / 16; a.nkb = CI > 64 ? 2 : 1; a.status = status;
  a.n_tiles = CO / CH_NT;
  a.tiles_per_utt = ceil_div(L, CH_NOUT);
  a.total_tiles = B * a.tiles_per_utt;
  int cpg = kNumSMs / a.n_tiles;
  if (cpg > a.total_tiles) cpg = a.total_tiles;
  const int grid = cpg * a.n_tiles;
  M2_CUDA_OK(allow_smem(voc_conv_h_kernel, CH_TOTAL));
  M2_LAUNCH_PDL(stage, voc_conv_h_kernel, grid, CH_THREADS, CH_TOTAL, s, tmap, a, debug_words_device());
  return M2TTS_OK;
}

}  // namespace m2

using namespace m2;

namespace {
__global__ void ch_join_planes_kernel(const __half* planes, long long n, float* y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __half2float(planes[i]) + __half2float(planes[n + i]);
}
}  // namespace

extern "C" size_t m2tts_conv1d_k3_h_workspace_bytes(int B, int C, int L) {
  if (!voc_conv_h_eligible(C, 1) || B <= 0 || L <= 0) return 0;
  return align_up(voc_conv_h_wblob_bytes(C), 256) + 3 * align_up((size_t)B * L * C * 4, 256) + 1024;   // weight image + x, residual, y planes
}

// y = act(conv1d(x, w, b, padding = 1)) (+ residual) (components.py:196-200, one convolution of the ResBlock), C = 128, k = 3.
// x / residual fp32 CHANNEL-LAST [B][L][C]; y fp32 channel-first [B][C][L] (out_cl = 0) or channel-last [B][L][C] (out_cl = 1, through
// the fp16 hi/lo planes the kernel hands to the next 16-bit split kernel).
extern "C" int m2tts_conv1d_k3_h(const float* x, const float* w, const float* b, const float* residual, float* y, int B, int C, int L,
                                 int act, int out_cl, int32_t* status, void* workspace, size_t workspace_bytes, m2tts_stream_t stream) {
  M2_REQUIRE(x && w && b && y && workspace, M2TTS_E_NULLPTR, "conv1d_k3_h: null pointer");
  M2_REQUIRE(voc_conv_h_eligible(C, 1), M2TTS_E_UNSUPPORTED, "conv1d_k3_h: C=%d (128)", C);
  M2_REQUIRE(act >= 0 && act <= 1 && out_cl >= 0 && out_cl <= 1, M2TTS_E_BADSHAPE, "conv1d_k3_h: act=%d out_cl=%d", act, out_cl);
  Carver cv(workspace, workspace_bytes);
  __half* wblob = cv.take<__half>(voc_conv_h_wblob_bytes(C) / 2);
  const long long n = (long long)B * L * C;
  __half* xp = cv.take<__half>((size_t)2 * n);
  __half* rp = cv.take<__half>((size_t)2 * n);
  __half* yp = cv.take<__half>((size_t)2 * n);
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "conv1d_k3_h: workspace too small or misaligned");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_split_planes_h(x, xp, n, status, s);
  if (rc) return rc;
  if (residual != nullptr && (rc = launch_split_planes_h(residual, rp, n, status, s))) return rc;
  rc = launch_voc_conv_h(xp, n, w, b, wblob, residual ? rp : nullptr, n, out_cl ? yp : nullptr, n, out_cl ? nullptr : y, L, B, C, C, L, act,
                         M2TTS_STAGE_VOC_RES1, status, s);
  if (rc) return rc;
  if (out_cl) M2_LAUNCH(M2TTS_STAGE_VOC_RES1, ch_join_planes_kernel, 1184, 256, 0, s, yp, n, y);
  return M2TTS_OK;
}
