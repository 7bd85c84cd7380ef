// voc_fused_h.cu — the fused narrow vocoder stage of voc_fused.cu with the 16-BIT split (fp16 hi/lo operands, fp32
// accumulation in TMEM; error model: attention_h.cu):
//   x [B][L][2C] -> u = lrelu(ConvTranspose1d(2C -> C, k=4, s=2, p=1)(x)) -> y = u + conv2(lrelu(conv1(u)))
//                -> (last stage) audio = tanh(Conv1d(C -> 1, k=3)(y))              tts_model.py:255-263,272,289-295
// What the 16-bit split buys here:
//   * inter-stage tensors travel as fp16 hi/lo planes [2][B][L][C] — the SAME bytes as fp32 — written by the producer's
//     epilogue, so the tile arrives from TMA ready to be a UMMA operand: no splitter warps, no raw buffer;
//   * operands are half the bytes: the C = 32 stage fits a 128-position tile (the TF32 kernel could only afford 64 and
//     wasted half of every transposed-conv UMMA), with a second X slot so the next tile's load is always in flight;
//   * K = 16 per UMMA: half the UMMA count of the TF32 kernel.
// Same organisation otherwise: channel-last rows are the K-major A operand, a convolution tap is the descriptor start
// address moved by whole rows, taps / phases / hi-lo folded into N, U and V written by the epilogue warps straight
// into swizzled operand rows. fp16 needs |x| <= 65504: every producer checks its values and raises M2TTS_ST_FP16_RANGE.
#include "conv_tc.cuh"
#include "attention_tc.cuh"
#include <cuda_fp16.h>
#include <math.h>

namespace m2 {

struct FusedHArgs {
  int B, L_in, L_out;
  int tiles_per_utt, total_tiles;
  const __half* wblob;                   // packed fp16 weight image (fh_wpack_kernel)
  const float* bias_up; const float* bias1; const float* bias2;
  const float* out_w; const float* out_b;   // FINAL: Conv1d(C,1,3) weight [1][C][3] and bias
  __half* out_h;                         // non-final: fp16 hi/lo planes [2][B][L_out][C]
  float* out_f;                          // non-final alternative: fp32 channel-last [B][L_out][C]; FINAL: audio [B][L_out]
  int dbg_nostore;                       // bring-up timing experiment (M2TTS_DBG_NOSTORE=1): skip the plane stores, results invalid
  long long out_plane;                   // elements between the hi and the lo plane of out_h
  int32_t* status;                       // M2TTS_ST_FP16_RANGE when U, V or the output planes leave the fp16 range
  int c_real;                            // channels that exist (8 in a 16-channel kernel: the stage-1 model's last stage; the rest are zero weights / biases)
  int tma_out;                           // C = 32 planes output: the tile leaves through V's shared-memory rows and two TMA stores
  long long* prof;                       // tools build: phase timestamps of CTA 0, context 0, first epilogue warp (m2tts_attention_set_prof buffer)
};
#ifdef M2TTS_TOOLS
#define FH_PROF(k) do { if (pt) a.prof[it * 8 + (k)] = clock64(); } while (0)
#define FH_PROF2(k) do { if (pt) a.prof[512 + it * 8 + (k)] = clock64(); } while (0)      // EPI1 in detail
#else
#define FH_PROF(k) do { } while (0)
#define FH_PROF2(k) do { } while (0)
#endif

template <int C, int NCTX, bool FINAL>
struct FhCfg {
  static constexpr int NQ = 128;
  // epilogue warpgroups per context: GP = 2 take the two phases of the transposed conv (EPI1) / the two 128-row halves of U
  // (EPI2, EPI3), GC = C / 16 the 16-channel chunks. With one warpgroup per context (C = 16) or two (C = 32) the kernels
  // ran 2 epilogue warps per scheduler and were bound by the latency of their own dependent chains (ncu: issue slots
  // 38-47 % busy, tensor pipe 16-27 %)
  static constexpr int GP = 2, GC = C / 16, G = GP * GC;
  static constexpr int XSLOTS = (NCTX == 1) ? 2 : 1; // input tile slots per context
  static constexpr int CI = 2 * C;
  static constexpr int XRB = CI * 2;                 // bytes of an input row (128 -> 128B swizzle, 64 -> 64B swizzle)
  static constexpr int URB = 64;                     // bytes of a U/V row (C = 16 rows are padded to 64 B)
  static constexpr int UROWS = 2 * NQ, HALVES = UROWS / 128;
  static constexpr int XR = NQ + 8;                  // input rows per tile (row j <-> q = Qs - 1 + j)
  static constexpr uint32_t XPL = XR * XRB;          // one plane of the input tile
  static constexpr uint32_t UPL = (UROWS + 8) * URB; // one plane of U or V (row i stored at index i + 1)
  static constexpr uint32_t O_X = 0;                 // [slot][plane]
  static constexpr uint32_t O_U = XSLOTS * 2 * XPL;
  static constexpr uint32_t O_V = O_U + 2 * UPL;
  static constexpr uint32_t CTX = O_V + 2 * UPL;
  static constexpr int ILO = FINAL ? 4 : 2, IHI = UROWS - ILO, NOUT = IHI - ILO;
  // weight image, bytes. Every part stacks [W_hi rows ; W_lo rows]; rows are K-major with the swizzle of their operand.
  static constexpr uint32_t W_UP0 = 0;                              // 4C rows x XRB: [p0 hi | p0 lo | p1 hi | p1 lo]
  static constexpr uint32_t W_UPM = 4 * C * XRB;                    // 2C rows: row q-1, phase 0
  static constexpr uint32_t W_UPP = W_UPM + 2 * C * XRB;            // 2C rows: row q+1, phase 1
  static constexpr uint32_t W_C1 = W_UPP + 2 * C * XRB;             // 3 taps x (2C rows x 64 B)
  static constexpr uint32_t W_C2 = W_C1 + 3 * 2 * C * URB;
  static constexpr uint32_t WBYTES = W_C2 + 3 * 2 * C * URB;
  static constexpr int T_UP = 0, T_C1 = 4 * C, T_C2 = 4 * C + HALVES * 2 * C, TCOLS_CTX = 4 * C + 4 * HALVES * C;
  static constexpr int NBAR = 2 * XSLOTS + 7;        // x_full[S] x_free[S] acc_up u_ready acc_c1[2] v_ready acc_c2[2]
  static constexpr uint32_t OFF_W = NCTX * CTX;
  static constexpr uint32_t OFF_CONST = OFF_W + WBYTES;
  static constexpr uint32_t OFF_EXCH = OFF_CONST + 1024;
  static constexpr uint32_t OFF_BAR = OFF_EXCH + (FINAL ? NCTX * GC * 3 * UROWS * 4 : 0);   // FINAL: partial tap sums [ctx][channel group][tap][row]
  static constexpr uint32_t TOTAL = OFF_BAR + 8 * (NCTX * NBAR + 1) + 16 + 1024;
  static constexpr int THREADS = 64 + 128 * NCTX * G;
  static_assert(C == 16 || C == 32, "fused stage: C in {16,32}");
  static_assert(CTX % 1024 == 0 && XPL % 512 == 0 && WBYTES % 1024 == 0, "operand alignment");
  static_assert(TOTAL <= 227 * 1024, "fused stage: shared memory");
  static_assert(NCTX * TCOLS_CTX <= 512, "fused stage: TMEM columns");
};

__device__ __forceinline__ void fh_tma_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
template <int ROWB>
__device__ __forceinline__ uint64_t fh_desc_tmpl() {     // K-major swizzled operand: SBO = 8 rows; start address added by the caller
  return ((uint64_t)1 << 16) | ((uint64_t)((8u * ROWB) >> 4) << 32) | (1ull << 46) | ((uint64_t)(ROWB == 128 ? 2 : 4) << 61);
}
__device__ __forceinline__ uint32_t fh_swz64(int row, int chunk) {    // byte offset of 16-byte chunk `chunk` of 64-byte row `row`
  return (uint32_t)row * 64u + ((uint32_t)(chunk ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ uint32_t fh_idesc(int N) {     // kind::f16, fp16 x fp16 -> fp32, both K-major, M = 128
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void fh_mma_w(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void fh_group_sync(int ctx, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(ctx + 1), "r"(threads) : "memory"); }
__device__ __forceinline__ float fh_lrelu(float v) { return v > 0.f ? v : 0.1f * v; }
__device__ __forceinline__ void fh_ld_sum16(uint32_t t_main, uint32_t t_corr, float* v) {
  uint32_t a[16], b[16];
  ct_ld16(t_main, a);
  ct_ld16(t_corr, b);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(a[j]) + __uint_as_float(b[j]);
}
// 8 floats -> one 16-byte chunk of fp16 hi and one of fp16 lo; `bad` records a value outside the fp16 range
__device__ __forceinline__ void fh_split8(const float* x, uint4& hi, uint4& lo, bool& bad) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) h_split2(x[2 * e], x[2 * e + 1], h[e], l[e], bad);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void fh_join8(const uint4& hi, const uint4& lo, float* x) {     // hi + lo -> 8 floats
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h[e]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&l[e]));
    x[2 * e] = a.x + b.x; x[2 * e + 1] = a.y + b.y;
  }
}

// accumulator chunk of 16 columns (main + correction halves) -> 8 packed fp32 pairs, + bias pairs from shared memory
__device__ __forceinline__ void fh_ld_sum16_pairs(uint32_t t_main, uint32_t t_corr, const float* bias16, uint64_t* v) {
  uint32_t a[16], b[16];
  ct_ld16(t_main, a);
  ct_ld16(t_corr, b);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const uint64_t* bp = reinterpret_cast<const uint64_t*>(bias16);      // 8-byte aligned: bias16 = consts + multiple of 16 floats
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = f2_add(f2_add(f2_pack_u(a[2 * j], a[2 * j + 1]), f2_pack_u(b[2 * j], b[2 * j + 1])), bp[j]);
}
// 8 pairs -> two 16-byte chunks of fp16 hi and two of fp16 lo, stored at 64-byte row `row`, chunks ch0, ch0 + 1 of planes P, P + UPL
__device__ __forceinline__ void fh_split_store16(const uint64_t* v, uint8_t* plane_hi, uint32_t plane_stride, int row, int ch0, float& amax) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) h_split_pair(v[j], hi[j], lo[j], amax);
#pragma unroll
  for (int j8 = 0; j8 < 2; ++j8) {
    const uint32_t off = fh_swz64(row, ch0 + j8);
    *reinterpret_cast<uint4*>(plane_hi + off) = make_uint4(hi[4 * j8], hi[4 * j8 + 1], hi[4 * j8 + 2], hi[4 * j8 + 3]);
    *reinterpret_cast<uint4*>(plane_hi + plane_stride + off) = make_uint4(lo[4 * j8], lo[4 * j8 + 1], lo[4 * j8 + 2], lo[4 * j8 + 3]);
  }
}

template <int C, int NCTX, bool FINAL>
__global__ void __launch_bounds__(FhCfg<C, NCTX, FINAL>::THREADS, 1)
voc_stage_fused_h_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y, const FusedHArgs a, int* dbg) {
  pdl_launch_dependents();      // M2_LAUNCH_PDL: every access to another kernel's data follows a pdl_wait()
  using K = FhCfg<C, NCTX, FINAL>;
  constexpr int CI = K::CI, XRB = K::XRB, URB = K::URB, HALVES = K::HALVES, XR = K::XR, G = K::G, GC = K::GC, XS = K::XSLOTS, NQ = K::NQ;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ct_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - ct_smem_u32(smem_raw));
  const uint32_t bars = sbase + K::OFF_BAR;
  auto bar = [&](int ctx, int which) { return bars + 8u * (uint32_t)(ctx * K::NBAR + which); };
  // barrier indices inside a context
  constexpr int X_FULL = 0, X_FREE = XS, ACC_UP = 2 * XS, U_READY = 2 * XS + 1, ACC_C1 = 2 * XS + 2, V_READY = 2 * XS + 4, ACC_C2 = 2 * XS + 5;
  const uint32_t bar_w = bars + 8u * (NCTX * K::NBAR);
  const uint32_t tmem_slot = bar_w + 8;
  float* consts = reinterpret_cast<float*>(gbase + K::OFF_CONST);   // [0,C) b_up | [C,2C) b1 | [2C,3C) b2 | [3C,6C) out_w | [6C] out_b

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int n_iter = (a.total_tiles + (int)gridDim.x * NCTX - 1) / ((int)gridDim.x * NCTX);
  auto tile_of = [&](int it, int ctx) { return (it * (int)gridDim.x + (int)blockIdx.x) * NCTX + ctx; };

  if (tid == 0) {
    for (int c = 0; c < NCTX; ++c) {
      for (int s = 0; s < XS; ++s) { ct_mbar_init(bar(c, X_FULL + s), 1); ct_mbar_init(bar(c, X_FREE + s), 1); }
      ct_mbar_init(bar(c, ACC_UP), 1);    ct_mbar_init(bar(c, U_READY), 4 * G);
      ct_mbar_init(bar(c, ACC_C1), 1);    ct_mbar_init(bar(c, ACC_C1 + 1), 1);
      ct_mbar_init(bar(c, V_READY), 4 * G);
      ct_mbar_init(bar(c, ACC_C2), 1);    ct_mbar_init(bar(c, ACC_C2 + 1), 1);
    }
    ct_mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  for (int i = tid; i < 6 * C + 1; i += K::THREADS) {
    float v = 0.f;
    const int cr = a.c_real;      // channels >= cr are padding: zero biases and output-conv weights
    if (i < C) v = i < cr ? a.bias_up[i] : 0.f;
    else if (i < 2 * C) v = i - C < cr ? a.bias1[i - C] : 0.f;
    else if (i < 3 * C) v = i - 2 * C < cr ? a.bias2[i - 2 * C] : 0.f;
    else if (FINAL && i < 6 * C) { const int e = i - 3 * C; v = e % C < cr ? a.out_w[(e % C) * 3 + e / C] : 0.f; }   // [tap][ci]
    else if (FINAL) v = a.out_b[0];
    consts[i] = v;
  }
  constexpr uint32_t TMEM_COLS = (NCTX * K::TCOLS_CTX <= 128) ? 128u : (NCTX * K::TCOLS_CTX <= 256 ? 256u : 512u);
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
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
      // ===== producer: weights once, then the hi and lo planes of one input tile per (iteration, context) =====
      ct_expect_tx(bar_w, K::WBYTES);
      for (uint32_t off = 0; off < K::WBYTES; off += 8192u) {
        const uint32_t n = K::WBYTES - off < 8192u ? K::WBYTES - off : 8192u;
        ct_bulk(sbase + K::OFF_W + off, reinterpret_cast<const uint8_t*>(a.wblob) + off, n, bar_w);
      }
      pdl_wait();
      for (int it = 0; it < n_iter; ++it)
        for (int c = 0; c < NCTX; ++c) {
          const int g = tile_of(it, c);
          if (g >= a.total_tiles) continue;
          const int b = g / a.tiles_per_utt, k = g % a.tiles_per_utt;
          const int Qs = k * (K::NOUT / 2) - K::ILO / 2;
          const int slot = it % XS, use = it / XS;
          if (use > 0) ct_wait(bar(c, X_FREE + slot), (uint32_t)((use - 1) & 1), dbg, 1, it);   // the up-GEMM that read this slot is done
          ct_expect_tx(bar(c, X_FULL + slot), 2 * K::XPL);
          const uint32_t dst = sbase + (uint32_t)c * K::CTX + K::O_X + (uint32_t)slot * 2 * K::XPL;
          fh_tma_4d(dst, &tmap_x, 0, Qs - 1, b, 0, bar(c, X_FULL + slot));
          fh_tma_4d(dst + K::XPL, &tmap_x, 0, Qs - 1, b, 1, bar(c, X_FULL + slot));
        }
    }
  } else if (warp == 1) {
    // ===== UMMA issuer: the whole warp runs the loop with warp-uniform operands, one elected lane issues =====
    ct_wait(bar_w, 0, dbg, 2, 0);
    const uint32_t sW = sbase + K::OFF_W;
    const uint64_t x_tmpl = fh_desc_tmpl<XRB>(), u_tmpl = fh_desc_tmpl<URB>();
    const uint32_t id_4c = fh_idesc(4 * C), id_2c = fh_idesc(2 * C), id_c = fh_idesc(C);
    auto dsc = [](uint64_t tmpl, uint32_t addr) -> uint64_t { return tmpl | (uint64_t)((addr >> 4) & 0x3FFFu); };
    // transposed conv: D[q, (phase, main|corr, co)]; 7 UMMAs per 16-channel k-step
    auto issue_up = [&](int c, int slot) {
      const uint32_t sX = sbase + (uint32_t)c * K::CTX + K::O_X + (uint32_t)slot * 2 * K::XPL;
      const uint32_t d = tmem_base + (uint32_t)(c * K::TCOLS_CTX + K::T_UP);
#pragma unroll
      for (int ks = 0; ks < CI / 16; ++ks) {
        auto xa = [&](int plane, int shift) -> uint64_t {
          return dsc(x_tmpl, sX + (uint32_t)plane * K::XPL + (uint32_t)(1 + shift) * XRB + (uint32_t)ks * 32u);
        };
        auto wb = [&](uint32_t part, int r0) -> uint64_t { return dsc(x_tmpl, sW + part + (uint32_t)r0 * XRB + (uint32_t)ks * 32u); };
        fh_mma_w(d, xa(0, 0), wb(K::W_UP0, 0), id_4c, ks ? 1u : 0u);          // row q: A_hi x [p0 hi | p0 lo | p1 hi | p1 lo]
        fh_mma_w(d, xa(1, 0), wb(K::W_UP0, 0), id_c, 1u);                     //        A_lo x p0 hi
        fh_mma_w(d + 2 * C, xa(1, 0), wb(K::W_UP0, 2 * C), id_c, 1u);         //        A_lo x p1 hi
        fh_mma_w(d, xa(0, -1), wb(K::W_UPM, 0), id_2c, 1u);                   // row q-1: phase 0
        fh_mma_w(d, xa(1, -1), wb(K::W_UPM, 0), id_c, 1u);
        fh_mma_w(d + 2 * C, xa(0, 1), wb(K::W_UPP, 0), id_2c, 1u);            // row q+1: phase 1
        fh_mma_w(d + 2 * C, xa(1, 1), wb(K::W_UPP, 0), id_c, 1u);
      }
      ct_commit_w(bar(c, ACC_UP));
      ct_commit_w(bar(c, X_FREE + slot));
    };
    // ResBlock conv: D[i, (main|corr, co)] = sum_tap A[i + tap - 1, :] W_tap; 2 UMMAs per (tap, k-step)
    auto issue_conv = [&](int c, int conv) {
      const uint32_t sA = sbase + (uint32_t)c * K::CTX + (conv == 0 ? K::O_U : K::O_V);
      const uint32_t wpart = conv == 0 ? K::W_C1 : K::W_C2;
#pragma unroll
      for (int h = 0; h < HALVES; ++h) {
        const uint32_t d = tmem_base + (uint32_t)(c * K::TCOLS_CTX + (conv == 0 ? K::T_C1 : K::T_C2) + h * 2 * C);
#pragma unroll
        for (int tap = 0; tap < 3; ++tap)
#pragma unroll
          for (int ks = 0; ks < C / 16; ++ks) {
            const uint32_t a_hi = sA + (uint32_t)(128 * h + tap) * URB + (uint32_t)ks * 32u;
            const uint64_t bd = dsc(u_tmpl, sW + wpart + (uint32_t)tap * (2 * C * URB) + (uint32_t)ks * 32u);
            fh_mma_w(d, dsc(u_tmpl, a_hi), bd, id_2c, (tap | ks) ? 1u : 0u);
            fh_mma_w(d, dsc(u_tmpl, a_hi + K::UPL), bd, id_c, 1u);
          }
        ct_commit_w(bar(c, (conv == 0 ? ACC_C1 : ACC_C2) + h));
      }
    };
    if (NCTX == 1) {
      // one context, two input slots: up(it+1) is issued behind conv2(it) and runs under this tile's last epilogue
      if (tile_of(0, 0) < a.total_tiles) {
        ct_wait(bar(0, X_FULL), 0u, dbg, 3, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_up(0, 0);
      }
      for (int it = 0; it < n_iter; ++it) {
        if (tile_of(it, 0) >= a.total_tiles) break;
        const uint32_t par = (uint32_t)(it & 1);
        ct_wait(bar(0, U_READY), par, dbg, 4, it);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_conv(0, 0);
        ct_wait(bar(0, V_READY), par, dbg, 5, it);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_conv(0, 1);
        if (it + 1 < n_iter && tile_of(it + 1, 0) < a.total_tiles) {
          const int slot = (it + 1) % XS, use = (it + 1) / XS;
          ct_wait(bar(0, X_FULL + slot), (uint32_t)(use & 1), dbg, 3, it + 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_up(0, slot);
        }
      }
    } else {
      for (int it = 0; it < n_iter; ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        for (int c = 0; c < NCTX; ++c) {
          if (tile_of(it, c) >= a.total_tiles) continue;
          ct_wait(bar(c, X_FULL), par, dbg, 3, it);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_up(c, 0);
        }
#pragma unroll
        for (int conv = 0; conv < 2; ++conv)
          for (int c = 0; c < NCTX; ++c) {
            if (tile_of(it, c) >= a.total_tiles) continue;
            ct_wait(bar(c, conv == 0 ? U_READY : V_READY), par, dbg, 4 + conv, it);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_conv(c, conv);
          }
      }
    }
  } else {
    // ===== epilogue warpgroup g of context c: thread m owns TMEM lane m =====
    pdl_wait();
    const int eg = (warp - 2) >> 2;
    const int c = eg / G, g = eg % G;
    const int gp = g / GC, gc = g % GC;             // phase / row half, 16-channel chunk
    const int qtr = warp & 3;
    const int m = qtr * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qtr * 32) << 16) + (uint32_t)(c * K::TCOLS_CTX);
    uint8_t* Ub = gbase + (uint32_t)c * K::CTX + K::O_U;
    uint8_t* Vb = gbase + (uint32_t)c * K::CTX + K::O_V;
    float* exch = reinterpret_cast<float*>(gbase + K::OFF_EXCH) + c * GC * 3 * K::UROWS;
    const float* b_up = consts, *b1 = consts + C, *b2 = consts + 2 * C, *ow = consts + 3 * C;
    const int c0 = 16 * gc;                          // this warpgroup's channels
    float amax = 0.f;                                // max |value| written as fp16 planes (NaN sticks): the fp16-range check
    for (int it = 0; it < n_iter; ++it) {
      const int gt = tile_of(it, c);
      if (gt >= a.total_tiles) break;
      const uint32_t par = (uint32_t)(it & 1);
      const int b = gt / a.tiles_per_utt, k = gt % a.tiles_per_utt;
      const int Qs = k * (K::NOUT / 2) - K::ILO / 2;
      const int Ts = 2 * Qs;                               // output position of U row 0
#ifdef M2TTS_TOOLS
      const bool pt = a.prof != nullptr && blockIdx.x == 0 && eg == 0 && qtr == 0 && lane == 0 && it < 60;
#endif
      FH_PROF(0);

      // ---- EPI1: transposed-conv accumulator -> U = lrelu(. + bias), zero outside the utterance, fp16 hi/lo rows ----
      ct_wait(bar(c, ACC_UP), par, dbg, 8, it);
      FH_PROF(1);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        const int q = Qs + m;
        const float keep = (q >= 0 && q < a.L_in) ? 1.f : 0.f;
        const int p = gp;                                  // each warpgroup takes one phase and 16 channels
        const int row = 2 * m + p + 1;
        uint64_t v[8];
        FH_PROF2(0);
        fh_ld_sum16_pairs(t_lane + (uint32_t)(K::T_UP + p * 2 * C + c0), t_lane + (uint32_t)(K::T_UP + p * 2 * C + C + c0), b_up + c0, v);
        FH_PROF2(1);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = f2_lrelu01(v[j]);
        if (!__all_sync(0xffffffffu, keep != 0.f)) {       // only the tiles at the ends of an utterance have rows to zero
          const uint64_t k2 = f2_pack(keep, keep);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = f2_mul(v[j], k2);
        }
        fh_split_store16(v, Ub, K::UPL, row, c0 >> 3, amax);
        FH_PROF2(2);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      FH_PROF2(3);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      FH_PROF2(4);
      __syncwarp();
      if (lane == 0) ct_arrive(bar(c, U_READY));
      FH_PROF(2);

      // ---- EPI2: conv1 accumulator -> V = lrelu(. + bias), zero outside the utterance, fp16 hi/lo rows ----
      if (!FINAL && a.tma_out && it > 0) {      // the previous tile's output left through V: its TMA stores must have read it
        if (eg == 0 && qtr == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        fh_group_sync(c, 128 * G);
      }
      {
        const int h = gp;                                    // each warpgroup takes one 128-row half and 16 channels
        ct_wait(bar(c, ACC_C1 + h), par, dbg, 9, it);
        FH_PROF(3);
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int i = 128 * h + m;
        const int t = Ts + i;
        const float keep = (t >= 0 && t < a.L_out) ? 1.f : 0.f;
        // rows 0 and UROWS - 1 of V are computed from the two U rows outside the tile (never written: whatever shared memory
        // held) and only feed output rows the tile discards (ILO >= 2): they must not raise the range flag
        float amax_v = 0.f;
        uint64_t v[8];
        fh_ld_sum16_pairs(t_lane + (uint32_t)(K::T_C1 + h * 2 * C + c0), t_lane + (uint32_t)(K::T_C1 + h * 2 * C + C + c0), b1 + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = f2_lrelu01(v[j]);
        if (!__all_sync(0xffffffffu, keep != 0.f)) {
          const uint64_t k2 = f2_pack(keep, keep);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = f2_mul(v[j], k2);
        }
        fh_split_store16(v, Vb, K::UPL, i + 1, c0 >> 3, amax_v);
        if (i >= 1 && i <= K::UROWS - 2) amax = amax_nan3(amax, amax_v, 0.f);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) ct_arrive(bar(c, V_READY));
      FH_PROF(4);

      // ---- EPI3: conv2 accumulator + bias + U -> stage output (or the 1-channel output conv + tanh) ----
      {
        const int h = gp;
        ct_wait(bar(c, ACC_C2 + h), par, dbg, 10, it);
        if (!FINAL && a.tma_out) ct_wait(bar(c, ACC_C2 + (h ^ 1)), par, dbg, 10, it);      // V is overwritten below: both halves of conv2 must have read it
        FH_PROF(5);
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int i = 128 * h + m;
        const int t = Ts + i;
        const bool inside = (t >= 0 && t < a.L_out);
        uint64_t y[8];
        fh_ld_sum16_pairs(t_lane + (uint32_t)(K::T_C2 + h * 2 * C + c0), t_lane + (uint32_t)(K::T_C2 + h * 2 * C + C + c0), b2 + c0, y);
#pragma unroll
        for (int j8 = 0; j8 < 2; ++j8) {      // + the residual U (hi + lo, exact)
          const uint32_t off = fh_swz64(i + 1, (c0 >> 3) + j8);
          const uint4 uh = *reinterpret_cast<const uint4*>(Ub + off), ul = *reinterpret_cast<const uint4*>(Ub + K::UPL + off);
          y[4 * j8] = f2_add(y[4 * j8], h_join_pair(uh.x, ul.x));
          y[4 * j8 + 1] = f2_add(y[4 * j8 + 1], h_join_pair(uh.y, ul.y));
          y[4 * j8 + 2] = f2_add(y[4 * j8 + 2], h_join_pair(uh.z, ul.z));
          y[4 * j8 + 3] = f2_add(y[4 * j8 + 3], h_join_pair(uh.w, ul.w));
        }
        if (!FINAL && a.tma_out) {
          // planes through TMA: y goes to row i + 2 of V (dead once conv2 has run; index i + 2 puts the first stored row, i = ILO,
          // on a 128-byte boundary) in V's own swizzled layout and the rows [ILO, IHI) leave as one box per plane below —
          // thread-per-row global stores touched 32 lines per instruction (EPI3 + its barrier: half of the tile period)
          float amax_y = 0.f;
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) h_split_pair(y[j], hi[j], lo[j], amax_y);
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8) {
            const uint32_t off = fh_swz64(i + 2, (c0 >> 3) + j8);
            *reinterpret_cast<uint4*>(Vb + off) = make_uint4(hi[4 * j8], hi[4 * j8 + 1], hi[4 * j8 + 2], hi[4 * j8 + 3]);
            *reinterpret_cast<uint4*>(Vb + K::UPL + off) = make_uint4(lo[4 * j8], lo[4 * j8 + 1], lo[4 * j8 + 2], lo[4 * j8 + 3]);
          }
          if (inside && i >= K::ILO && i < K::IHI) amax = amax_nan3(amax, amax_y, 0.f);
        } else if (!FINAL) {
          if (inside && i >= K::ILO && i < K::IHI) {
            const size_t o = ((size_t)b * a.L_out + t) * C + c0;
            if (a.out_h != nullptr) {        // fp16 hi/lo planes for the next fused stage
              uint32_t hi[8], lo[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) h_split_pair(y[j], hi[j], lo[j], amax);
#pragma unroll
              for (int j8 = 0; j8 < 2; ++j8) {
                if (!a.dbg_nostore) *reinterpret_cast<uint4*>(a.out_h + o + 8 * j8) = make_uint4(hi[4 * j8], hi[4 * j8 + 1], hi[4 * j8 + 2], hi[4 * j8 + 3]);
                if (!a.dbg_nostore) *reinterpret_cast<uint4*>(a.out_h + a.out_plane + o + 8 * j8) = make_uint4(lo[4 * j8], lo[4 * j8 + 1], lo[4 * j8 + 2], lo[4 * j8 + 3]);
              }
            } else {                         // plain fp32 channel-last
              uint64_t* op = reinterpret_cast<uint64_t*>(a.out_f + o);
#pragma unroll
              for (int j = 0; j < 8; ++j) op[j] = y[j];
            }
          }
        } else {      // the output conv zero-pads y outside the utterance; every channel group contributes its channels
          uint64_t d0 = f2_pack(0.f, 0.f), d1 = d0, d2 = d0;
          const uint64_t* ow2 = reinterpret_cast<const uint64_t*>(ow + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            d0 = f2_fma(ow2[j], y[j], d0);
            d1 = f2_fma(ow2[C / 2 + j], y[j], d1);
            d2 = f2_fma(ow2[C + j], y[j], d2);
          }
          float e0, e1, e2, e3, e4, e5;
          f2_unpack(d0, e0, e1); f2_unpack(d1, e2, e3); f2_unpack(d2, e4, e5);
          float* e = exch + gc * 3 * K::UROWS;
          e[i] = inside ? e0 + e1 : 0.f;
          e[K::UROWS + i] = inside ? e2 + e3 : 0.f;
          e[2 * K::UROWS + i] = inside ? e4 + e5 : 0.f;
        }
      }
      FH_PROF(6);
      if (!FINAL && a.tma_out) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      fh_group_sync(c, 128 * G);              // U reads done (the next EPI1 may overwrite it); exch complete
      if (!FINAL && a.tma_out && eg == 0 && qtr == 0 && lane == 0 && !a.dbg_nostore) {
        // output rows i in [ILO, IHI) = positions k NOUT ..., V rows ILO + 2 ...; TMA clips positions >= L_out
        const uint32_t s0 = sbase + (uint32_t)c * K::CTX + K::O_V + (uint32_t)(K::ILO + 2) * URB;
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                     ::"l"(&tmap_y), "r"(0), "r"(k * K::NOUT), "r"(b), "r"(0), "r"(s0) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                     ::"l"(&tmap_y), "r"(0), "r"(k * K::NOUT), "r"(b), "r"(1), "r"(s0 + K::UPL) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (FINAL) {
        if (gc == 0) {
          const int i = 128 * gp + m;
          const int t = Ts + i;
          if (i >= K::ILO && i < K::IHI && t < a.L_out) {
            float s = consts[6 * C];
#pragma unroll
            for (int gg = 0; gg < GC; ++gg) {
              const float* e = exch + gg * 3 * K::UROWS;
              s += e[i - 1] + e[K::UROWS + i] + e[2 * K::UROWS + i + 1];
            }
            a.out_f[(size_t)b * a.L_out + t] = tanhf(s);
          }
        }
        fh_group_sync(c, 128 * G);            // exch reads done before the next tile rewrites it
      }
    }
    if (!FINAL && a.tma_out && eg == 0 && qtr == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    h_flag(h_amax_bad(amax), a.status);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- weight image: K-major rows with the swizzle of the operand they multiply; parts stack [hi rows ; lo rows] ----
struct FhPackArgs { const float* up_w; const float* w1; const float* w2; __half* blob; int C; int32_t* status; int c_real; };
__global__ void fh_wpack_kernel(FhPackArgs p) {
  const int C = p.C, CI = 2 * C, XRB = CI * 2;
  const int n_up0 = 4 * C * CI, n_upm = 2 * C * CI, n_conv = 2 * C * C;
  const int total = n_up0 + 2 * n_upm + 6 * n_conv;
  const uint32_t b_upm = 4 * C * XRB, b_upp = b_upm + 2 * C * XRB, b_c = b_upp + 2 * C * XRB;
  bool bad = false;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int e = idx, n, k, lo, rowb;
    uint32_t base;
    float v;
    if (e < n_up0) {                       // rows [p0 hi | p0 lo | p1 hi | p1 lo], each C rows
      base = 0; rowb = XRB; n = e / CI; k = e % CI;
      const int ph = n / (2 * C), co = n % C; lo = (n / C) & 1;
      v = (k < 2 * p.c_real && co < p.c_real) ? p.up_w[((size_t)k * p.c_real + co) * 4 + ph + 1] : 0.f;
    } else if (e < n_up0 + 2 * n_upm) {    // row q-1 (kernel tap 3) then row q+1 (kernel tap 0): [hi | lo]
      e -= n_up0;
      const int which = e / n_upm; e -= which * n_upm;
      base = which == 0 ? b_upm : b_upp; rowb = XRB; n = e / CI; k = e % CI;
      const int co = n % C; lo = n / C;
      v = (k < 2 * p.c_real && co < p.c_real) ? p.up_w[((size_t)k * p.c_real + co) * 4 + (which == 0 ? 3 : 0)] : 0.f;
    } else {                               // conv1 taps 0..2, conv2 taps 0..2: [hi | lo], 64-byte rows
      e -= n_up0 + 2 * n_upm;
      const int part = e / n_conv; e -= part * n_conv;
      base = b_c + (uint32_t)part * (2 * C * 64); rowb = 64; n = e / C; k = e % C;
      const int co = n % C; lo = n / C;
      const float* w = part < 3 ? p.w1 : p.w2;
      v = (k < p.c_real && co < p.c_real) ? w[((size_t)co * p.c_real + k) * 3 + (part % 3)] : 0.f;
    }
    h_chk(v, bad);
    const __half h = __float2half_rn(v);
    const uint32_t sw = rowb == 128 ? (uint32_t)(n & 7) : (uint32_t)((n >> 1) & 3);
    const uint32_t off = base + (uint32_t)n * rowb + ((((uint32_t)k >> 3) ^ sw) << 4) + (uint32_t)(k & 7) * 2u;
    p.blob[off >> 1] = lo ? __float2half_rn(v - __half2float(h)) : h;
  }
  h_flag(bad, p.status);
}

// fp32 channel-last rows -> fp16 hi/lo planes (stand-alone entry / producers that are not ours) and back
__global__ void fh_split_planes_kernel(const float* __restrict__ x, __half* __restrict__ planes, long long n, int32_t* __restrict__ status) {
  pdl_launch_dependents();
  pdl_wait();
  bool bad = false;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += (long long)gridDim.x * blockDim.x * 8) {
    float v[8];
    const float4 a = *reinterpret_cast<const float4*>(x + i), b = *reinterpret_cast<const float4*>(x + i + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    uint4 hi, lo;
    fh_split8(v, hi, lo, bad);
    *reinterpret_cast<uint4*>(planes + i) = hi;
    *reinterpret_cast<uint4*>(planes + n + i) = lo;
  }
  h_flag(bad, status);
}

typedef CUresult (*EncodeTiledFn6)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn6 fh_encode_fn() {
  static EncodeTiledFn6 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn6)p;
  }
  return fn;
}

extern long long* g_ws_prof;      // attention_tc.cu (m2tts_attention_set_prof)

template <int C, int NCTX, bool FINAL>
static int launch_fh(const __half* xh, long long x_plane, FusedHArgs a, int stage, cudaStream_t s) {
  using K = FhCfg<C, NCTX, FINAL>;
  EncodeTiledFn6 enc = fh_encode_fn();
  M2_REQUIRE(enc != nullptr, M2TTS_E_CUDA, "voc_fused_h: cuTensorMapEncodeTiled unavailable");
  CUtensorMap tmap;
  const int ci_real = 2 * a.c_real;      // the box covers CI channels; those beyond ci_real are zero-filled by TMA
  const cuuint64_t dims[4] = {(cuuint64_t)ci_real, (cuuint64_t)a.L_in, (cuuint64_t)a.B, 2};
  const cuuint64_t strides[3] = {(cuuint64_t)ci_real * 2, (cuuint64_t)a.L_in * ci_real * 2, (cuuint64_t)x_plane * 2};
  const cuuint32_t box[4] = {(cuuint32_t)K::CI, (cuuint32_t)K::XR, 1u, 1u};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(xh), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, K::XRB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "voc_fused_h: cuTensorMapEncodeTiled failed (%d)", (int)r);
  CUtensorMap tmap_y = tmap;      // placeholder unless the output planes leave through TMA
  a.tma_out = 0;
  if (!FINAL && C == 32 && a.out_h != nullptr && (((uintptr_t)a.out_h) & 15) == 0) {
    const cuuint64_t ydims[4] = {(cuuint64_t)C, (cuuint64_t)a.L_out, (cuuint64_t)a.B, 2};
    const cuuint64_t ystrides[3] = {(cuuint64_t)C * 2, (cuuint64_t)a.L_out * C * 2, (cuuint64_t)a.out_plane * 2};
    const cuuint32_t ybox[4] = {(cuuint32_t)C, (cuuint32_t)K::NOUT, 1u, 1u};
    const CUresult ry = enc(&tmap_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.out_h, ydims, ystrides, ybox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(ry == CUDA_SUCCESS, M2TTS_E_CUDA, "voc_fused_h: cuTensorMapEncodeTiled (output planes) failed (%d)", (int)ry);
    a.tma_out = 1;
  }
  a.tiles_per_utt = ceil_div(a.L_out, K::NOUT);
  a.total_tiles = a.B * a.tiles_per_utt;
  int grid = ceil_div(a.total_tiles, NCTX);
  if (grid > kNumSMs) grid = kNumSMs;
  M2_CUDA_OK(allow_smem(voc_stage_fused_h_kernel<C, NCTX, FINAL>, K::TOTAL));
  M2_LAUNCH_PDL(stage, (voc_stage_fused_h_kernel<C, NCTX, FINAL>), grid, K::THREADS, K::TOTAL, s, tmap, tmap_y, a, debug_words_device());
  return M2TTS_OK;
}

size_t voc_fused_h_wblob_bytes(int C) { return C == 32 ? FhCfg<32, 1, false>::WBYTES : (C == 16 || C == 8 ? FhCfg<16, 2, false>::WBYTES : 0); }
// C = 8 runs in the 16-channel kernel with zero-padded weights, only as the LAST stage (the output is the waveform, not planes)
bool voc_fused_h_eligible(int C, int r, int dil, bool final_stage) { return r == 2 && dil == 1 && (C == 16 || C == 32 || (C == 8 && final_stage)); }

// xh: fp16 hi/lo planes, channel-last [2][B][L_in][2C] (x_plane elements apart). Output: out_h (fp16 hi/lo planes
// [2][B][2L][C], out_plane apart) or out_f (fp32 channel-last [B][2L][C]); with out_w != null: audio fp32 [B][2L] in out_f.
int launch_voc_stage_fused_h(const void* xh, long long x_plane, const float* up_w, const float* up_b, const float* w1, const float* b1,
                             const float* w2, const float* b2, const float* out_w, const float* out_b, void* wblob,
                             void* out_h, long long out_plane, float* out_f, int B, int C, int L_in, int stage, int32_t* status, cudaStream_t s) {
  M2_REQUIRE(C == 16 || C == 32 || C == 8, M2TTS_E_UNSUPPORTED, "voc_fused_h: C=%d (8, 16 or 32)", C);
  const int c_real = C;
  if (C == 8) C = 16;      // zero-padded to the 16-channel kernel
  if (up_w != nullptr) {      // (re)write the weight image; up_w == nullptr: wblob already holds it
    M2_REQUIRE(w1 != nullptr && w2 != nullptr && (((uintptr_t)wblob) & 15) == 0, M2TTS_E_BADSHAPE, "voc_fused_h: pack arguments");
    FhPackArgs p{up_w, w1, w2, (__half*)wblob, C, status, c_real};
    M2_LAUNCH(M2TTS_STAGE_PACK, fh_wpack_kernel, ceil_div(28 * C * C, 256), 256, 0, s, p);
  }
  if (xh == nullptr) return M2TTS_OK;      // pack only
  M2_REQUIRE((((uintptr_t)xh) & 15) == 0 && (((uintptr_t)wblob) & 15) == 0 && (x_plane & 7) == 0 && (out_plane & 7) == 0, M2TTS_E_BADSHAPE,
             "voc_fused_h: misaligned pointers");
  M2_REQUIRE(B > 0 && L_in > 0 && (long long)B * L_in * 2 < (1ll << 31), M2TTS_E_BADSHAPE, "voc_fused_h: B=%d L=%d", B, L_in);
  M2_REQUIRE(out_h != nullptr || out_f != nullptr, M2TTS_E_NULLPTR, "voc_fused_h: no output");
  FusedHArgs a{};
  a.B = B; a.L_in = L_in; a.L_out = 2 * L_in; a.wblob = (const __half*)wblob; a.bias_up = up_b; a.bias1 = b1; a.bias2 = b2;
  a.out_w = out_w; a.out_b = out_b; a.out_h = (__half*)out_h; a.out_f = out_f; a.out_plane = out_plane; a.c_real = c_real;
  { static int ns = -1; if (ns < 0) ns = tools_env_int("M2TTS_DBG_NOSTORE", 0) == 1 ? 1 : 0; a.dbg_nostore = ns; }
  a.status = status;
#ifdef M2TTS_TOOLS
  a.prof = g_ws_prof;
#endif
  const bool fin = out_w != nullptr;
  if (fin) M2_REQUIRE(out_f != nullptr, M2TTS_E_NULLPTR, "voc_fused_h: the last stage writes fp32 audio");
  M2_REQUIRE(c_real == C || fin, M2TTS_E_UNSUPPORTED, "voc_fused_h: C=%d only as the last stage", c_real);
  const __half* x = (const __half*)xh;
  if (C == 16) return fin ? launch_fh<16, 2, true>(x, x_plane, a, stage, s) : launch_fh<16, 2, false>(x, x_plane, a, stage, s);
  return fin ? launch_fh<32, 1, true>(x, x_plane, a, stage, s) : launch_fh<32, 1, false>(x, x_plane, a, stage, s);
}

int launch_split_planes_h(const float* x, void* planes, long long n, int32_t* status, cudaStream_t s) {
  M2_REQUIRE((n & 7) == 0 && (((uintptr_t)x) & 15) == 0 && (((uintptr_t)planes) & 15) == 0, M2TTS_E_BADSHAPE, "split_planes: n=%lld", n);
  long long blocks = (n / 8 + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  M2_LAUNCH_PDL(M2TTS_STAGE_PACK, fh_split_planes_kernel, (unsigned)blocks, 256, 0, s, x, (__half*)planes, n, status);
  return M2TTS_OK;
}

}  // namespace m2

using namespace m2;

extern "C" size_t m2tts_vocoder_stage_fused_h_workspace_bytes(int B, int C, int L) {
  if ((C != 16 && C != 32) || B <= 0 || L <= 0) return 0;
  return align_up(voc_fused_h_wblob_bytes(C), 256) + align_up((size_t)B * L * 2 * C * 4, 256) + 512;   // weight image + input planes
}

// Same contract as m2tts_vocoder_stage_fused (fp32 channel-last in, fp32 channel-last or audio out), 16-bit split inside.
extern "C" int m2tts_vocoder_stage_fused_h(const float* x, const float* up_w, const float* up_b, const float* res1_w,
                                           const float* res1_b, const float* res2_w, const float* res2_b,
                                           const float* out_w, const float* out_b, float* y, int B, int C, int L,
                                           int32_t* status, void* workspace, size_t workspace_bytes, m2tts_stream_t stream) {
  M2_REQUIRE(x && up_w && up_b && res1_w && res1_b && res2_w && res2_b && y && workspace, M2TTS_E_NULLPTR,
             "vocoder_stage_fused_h: null pointer");
  M2_REQUIRE((out_w == nullptr) == (out_b == nullptr), M2TTS_E_NULLPTR, "vocoder_stage_fused_h: out_w/out_b must both be set or both null");
  M2_REQUIRE(C == 16 || C == 32, M2TTS_E_UNSUPPORTED, "vocoder_stage_fused_h: C=%d (16 or 32)", C);
  Carver cv(workspace, workspace_bytes);
  __half* wblob = cv.take<__half>(voc_fused_h_wblob_bytes(C) / 2);
  const long long n = (long long)B * L * 2 * C;
  __half* planes = cv.take<__half>((size_t)2 * n);
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "vocoder_stage_fused_h: workspace too small or misaligned");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_split_planes_h(x, planes, n, status, s);
  if (rc) return rc;
  return launch_voc_stage_fused_h(planes, n, up_w, up_b, res1_w, res1_b, res2_w, res2_b, out_w, out_b, wblob, nullptr, 0, y, B, C, L,
                                  M2TTS_STAGE_VOC_FUSED, status, s);
}
