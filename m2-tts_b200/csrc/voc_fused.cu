// voc_fused.cu — one whole narrow vocoder stage as ONE persistent tcgen05 kernel, channel-last activations:
//   x [B][L][2C]  ->  u = lrelu(ConvTranspose1d(2C -> C, k=4, s=2, p=1)(x))            tts_model.py:255-263,291
//                 ->  y = u + conv2(lrelu(conv1(u)))            (LightweightResBlock)   components.py:196-200
//                 ->  (last stage) audio = tanh(Conv1d(C -> 1, k=3)(y))                 tts_model.py:272,295
// so a stage reads its input once and writes its output once (HBM traffic = algorithmic minimum); u and the
// ResBlock intermediate never leave the SM. fp32-faithful through 3xTF32 (see conv_tc.cu).
//
// Layout: activations are position-major rows of C floats (channel-last), i.e. the K-major UMMA operand with
// 128-byte (C = 32) or 64-byte (C = 16) swizzled rows. A convolution tap is a ROW shift of the A operand, applied
// by moving the descriptor start address by whole rows — the swizzle is a function of the shared-memory address,
// so any row offset is legal (measured: rowshift_probe.cu) — and all taps accumulate into ONE TMEM accumulator.
//
// Tile = NQ input positions (GEMM rows of the transposed conv) -> U rows i in [0, 2NQ) (output position
// t = 2*Qs + i), V = lrelu(conv1(U)) valid on [1, 2NQ-1), Y valid on [2, 2NQ-2); tiles overlap by the halo.
// Warp roles: 0 TMA producer (raw fp32 rows, zero-filled outside the utterance) | 1 UMMA issuer | 2-3 hi/lo
// splitters | 4.. epilogue groups (one per tile context). With NCTX = 2 two tiles are in flight per CTA and the
// issuer interleaves their GEMM phases, so one context's epilogue overlaps the other's tensor work.
#include "conv_tc.cuh"
#include <math.h>

namespace m2 {

#ifndef FS_UP_SPLIT
#define FS_UP_SPLIT 0   // 1: issue the next tile's transposed-conv GEMM in two halves around conv2; 0: after conv2
#endif

struct FusedStageArgs {
  int B, L_in, L_out;
  int tiles_per_utt, total_tiles;
  const float* wblob;                    // packed weight image (fs_wpack_kernel)
  const float* bias_up; const float* bias1; const float* bias2;
  const float* out_w; const float* out_b;   // FINAL: Conv1d(C,1,3) weight [1][C][3] and bias
  float* out;                            // [B][L_out][C] channel-last, or FINAL: audio [B][L_out]
  long long* prof;                       // optional: per-tile phase timestamps of CTA 0 / context 0 (bring-up)
};

// NCTX = 2 (C = 16): two tile contexts per CTA, each with its own epilogue warpgroup; V reuses the dead X split.
// NCTX = 1 (C = 32): one context (the 114 KB of weights leave room for no more); the raw tile is split IN PLACE so
//   the next tile's load and transposed-conv GEMM overlap this tile's ResBlock phases (V has its own buffer), and
//   G = 2 epilogue warpgroups share every epilogue (by phase in EPI1, by channel half in EPI2/EPI3).
template <int C, int NQ, int NCTX, bool FINAL>
struct FsCfg {
  static constexpr bool OVL = NCTX == 1;
  static constexpr int G = (OVL && !FINAL) ? 2 : 1;  // epilogue warpgroups per context
  static constexpr int CI = 2 * C;
  static constexpr int KB = CI / 32;                 // 128-byte k-blocks of an input row
  static constexpr int ROWB = C * 4;                 // bytes of one U/V row
  static constexpr int UROWS = 2 * NQ;
  static constexpr int HALVES = UROWS / 128;
  static constexpr int XR = NQ + 8;                  // input rows per tile in shared memory (row j <-> q = Qs - 1 + j)
  static constexpr uint32_t XRAW = KB * XR * 128;    // raw fp32 tile = one plane of the split
  static constexpr uint32_t UPL = (UROWS + 8) * ROWB;  // one plane of U or V (row i stored at index i + 1)
  // byte offsets inside a context
  static constexpr uint32_t O_RAW = 0;
  static constexpr uint32_t O_XS = OVL ? 0 : XRAW;                 // hi plane (lo plane follows at + XRAW)
  static constexpr uint32_t O_V = OVL ? 2 * XRAW : XRAW;           // V aliases the X split unless OVL
  static constexpr uint32_t O_U = OVL ? 2 * XRAW + 2 * UPL : 3 * XRAW;
  static constexpr uint32_t CTX = O_U + 2 * UPL;
  static constexpr int ILO = FINAL ? 4 : 2, IHI = UROWS - ILO;
  static constexpr int NOUT = IHI - ILO;             // output positions per tile
  // weight image, floats. Every part stacks [W_hi rows ; W_lo rows] so that A_hi x part is ONE wide UMMA.
  static constexpr int W_UP0 = 0;                    // 4C rows x CI: [p0 hi | p0 lo | p1 hi | p1 lo]
  static constexpr int W_UPM = 4 * C * CI;           // 2C rows x CI: row q-1, phase 0
  static constexpr int W_UPP = 6 * C * CI;           // 2C rows x CI: row q+1, phase 1
  static constexpr int W_C1 = 8 * C * CI;            // 3 taps x (2C rows x C)
  static constexpr int W_C2 = W_C1 + 6 * C * C;
  static constexpr int WFLOATS = W_C2 + 6 * C * C;   // = 28 C^2
  static constexpr uint32_t WBYTES = WFLOATS * 4u;
  // TMEM columns per context: up [p0 main | p0 corr | p1 main | p1 corr], conv halves [main | corr]
  static constexpr int T_UP = 0, T_C1 = 4 * C, T_C2 = 4 * C + HALVES * 2 * C;
  static constexpr int TCOLS_CTX = 4 * C + 4 * HALVES * C;
  static constexpr int NBAR = 11;                    // barriers per context
  static constexpr uint32_t OFF_W = NCTX * CTX;
  static constexpr uint32_t OFF_CONST = OFF_W + WBYTES;          // biases, output-conv weights
  static constexpr uint32_t OFF_EXCH = OFF_CONST + 1024;         // FINAL: p0/p2 exchange [NCTX][2][UROWS]
  static constexpr uint32_t OFF_BAR = OFF_EXCH + (FINAL ? NCTX * 2 * UROWS * 4 : 0);
  static constexpr uint32_t TOTAL = OFF_BAR + 8 * (NCTX * NBAR + 1) + 16 + 1024 /*alignment slack*/;
  static constexpr int THREADS = 128 + 128 * NCTX * G;
  static_assert(C == 16 || C == 32, "fused stage: C in {16,32}");
  static_assert(NQ == 64 || NQ == 128, "fused stage: NQ in {64,128}");
  static_assert(OVL || 2 * UPL <= 2 * XRAW, "V must fit in the dead X split region");
  static_assert(TOTAL <= 227 * 1024, "fused stage: shared memory");
  static_assert(NCTX * TCOLS_CTX <= 512, "fused stage: TMEM columns");
};

__device__ __forceinline__ void fs_tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// K-major swizzled A-operand descriptor template (start address added by the caller): SBO = 8 rows
template <int ROWB>
__device__ __forceinline__ uint64_t fs_adesc() {
  return ((uint64_t)1 << 16) | ((uint64_t)((8u * ROWB) >> 4) << 32) | (1ull << 46) | ((uint64_t)(ROWB == 128 ? 2 : 4) << 61);
}
template <int ROWB>
__device__ __forceinline__ uint32_t fs_swz(int row, int chunk) {     // byte offset of 16-byte chunk `chunk` of row `row`
  if (ROWB == 128) return (uint32_t)row * 128u + ((uint32_t)(chunk ^ (row & 7)) << 4);
  return (uint32_t)row * 64u + ((uint32_t)(chunk ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ uint32_t fs_idesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void fs_group_sync(int ctx, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(ctx + 1), "r"(threads) : "memory"); }
__device__ __forceinline__ float fs_lrelu(float v) { return v > 0.f ? v : 0.1f * v; }
// 16 accumulator columns (main) + the matching 16 "corr" columns (A_hi x W_lo part) -> 16 sums
__device__ __forceinline__ void fs_ld_sum16(uint32_t t_main, uint32_t t_corr, float* v) {
  uint32_t a[16], b[16];
  ct_ld16(t_main, a);
  ct_ld16(t_corr, b);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(a[j]) + __uint_as_float(b[j]);
}

template <int C, int NQ, int NCTX, bool FINAL>
__global__ void __launch_bounds__(FsCfg<C, NQ, NCTX, FINAL>::THREADS, 1)
voc_stage_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const FusedStageArgs a, int* dbg) {
  using K = FsCfg<C, NQ, NCTX, FINAL>;
  constexpr int CI = K::CI, KB = K::KB, ROWB = K::ROWB, HALVES = K::HALVES, XR = K::XR, G = K::G;
  constexpr bool OVL = K::OVL;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ct_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - ct_smem_u32(smem_raw));
  const uint32_t bars = sbase + K::OFF_BAR;
  auto bar = [&](int ctx, int which) { return bars + 8u * (uint32_t)(ctx * K::NBAR + which); };
  // XS_FREE: the transposed-conv GEMM (OVL) / conv2 (otherwise) has finished with the X split (resp. V) region
  enum { XRAW_FULL = 0, XRAW_EMPTY, XS_FULL, XS_FREE, ACC_UP, U_READY, ACC_C1 /*+h*/, V_READY = 8, ACC_C2 /*+h*/ };
  const uint32_t bar_w = bars + 8u * (NCTX * K::NBAR);
  const uint32_t tmem_slot = bar_w + 8;
  float* consts = reinterpret_cast<float*>(gbase + K::OFF_CONST);   // [0,C) b_up | [C,2C) b1 | [2C,3C) b2 | [3C,6C) out_w | [6C] out_b

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int n_iter = (a.total_tiles + (int)gridDim.x * NCTX - 1) / ((int)gridDim.x * NCTX);
  auto tile_of = [&](int it, int ctx) { return (it * (int)gridDim.x + (int)blockIdx.x) * NCTX + ctx; };

  if (tid == 0) {
    for (int c = 0; c < NCTX; ++c) {
      ct_mbar_init(bar(c, XRAW_FULL), 1); ct_mbar_init(bar(c, XRAW_EMPTY), 2);
      ct_mbar_init(bar(c, XS_FULL), 2);   ct_mbar_init(bar(c, XS_FREE), 1);
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
    if (i < C) v = a.bias_up[i];
    else if (i < 2 * C) v = a.bias1[i - C];
    else if (i < 3 * C) v = a.bias2[i - 2 * C];
    else if (FINAL && i < 6 * C) { const int e = i - 3 * C; v = a.out_w[(e % C) * 3 + e / C]; }   // [tap][ci]
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
      // ===== producer: weights once, then one raw fp32 tile per (iteration, context) =====
      ct_expect_tx(bar_w, K::WBYTES);
      for (uint32_t off = 0; off < K::WBYTES; off += 8192u) {
        const uint32_t n = K::WBYTES - off < 8192u ? K::WBYTES - off : 8192u;
        ct_bulk(sbase + K::OFF_W + off, reinterpret_cast<const uint8_t*>(a.wblob) + off, n, bar_w);
      }
      for (int it = 0; it < n_iter; ++it)
        for (int c = 0; c < NCTX; ++c) {
          const int g = tile_of(it, c);
          if (g >= a.total_tiles) continue;
          const int b = g / a.tiles_per_utt, k = g % a.tiles_per_utt;
          const int Qs = k * (K::NOUT / 2) - K::ILO / 2;
          // the landing buffer is free once the splitters have read it (separate raw buffer) or, when the split is
          // done in place, once the transposed-conv GEMM of the previous tile has consumed it
          if (it > 0) ct_wait(bar(c, OVL ? XS_FREE : XRAW_EMPTY), (uint32_t)((it - 1) & 1), dbg, 1, it);
          ct_expect_tx(bar(c, XRAW_FULL), K::XRAW);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            fs_tma_3d(sbase + (uint32_t)c * K::CTX + K::O_RAW + (uint32_t)kb * (XR * 128), &tmap_x, kb * 32, Qs - 1, b, bar(c, XRAW_FULL));
        }
    }
  } else if (warp == 1) {
    // ===== UMMA issuer: the whole warp runs the loop with warp-uniform operands, one elected lane issues =====
    ct_wait(bar_w, 0, dbg, 2, 0);
    const uint32_t sW = sbase + K::OFF_W;
    const uint64_t w_tmpl = ct_desc(0u, 128u, 512u, 0u);
    const uint64_t x_tmpl = fs_adesc<128>();
    const uint64_t u_tmpl = fs_adesc<ROWB>();
    const uint32_t id_4c = fs_idesc(4 * C), id_2c = fs_idesc(2 * C), id_c = fs_idesc(C);
    // B descriptor: rows [r0, ...) of a part image with NR rows in total, k-step ks
    auto wdesc = [&](int part_floats, int NR, int r0, int ks) -> uint64_t {
      const uint32_t addr = sW + (uint32_t)part_floats * 4u + (uint32_t)(ks >> 1) * (uint32_t)(NR * 64) + (uint32_t)(r0 >> 3) * 512u +
                            (uint32_t)(ks & 1) * 256u;
      return w_tmpl | (uint64_t)((addr >> 4) & 0x3FFFu);
    };
    // transposed conv: D[q, (phase, main|corr, co)]; 7 UMMAs per k-step
    auto issue_up = [&](int c, int part) {      // part 0 / 1: first / second half of the k-steps, 2: all
      const uint32_t sX = sbase + (uint32_t)c * K::CTX + K::O_XS;
      const uint32_t d = tmem_base + (uint32_t)(c * K::TCOLS_CTX + K::T_UP);
      const int ks0 = part == 1 ? CI / 16 : 0, ks1 = part == 0 ? CI / 16 : CI / 8;
#pragma unroll
      for (int ks = 0; ks < CI / 8; ++ks) {
        if (ks < ks0 || ks >= ks1) continue;
        const uint32_t koff = (uint32_t)(ks >> 2) * (XR * 128) + (uint32_t)(ks & 3) * 32u;
        auto adesc = [&](int plane, int shift) -> uint64_t {
          const uint32_t aaddr = sX + (uint32_t)plane * K::XRAW + koff + (uint32_t)(1 + shift) * 128u;
          return x_tmpl | (uint64_t)((aaddr >> 4) & 0x3FFFu);
        };
        // row q: all phases.  A_hi x [p0 hi | p0 lo | p1 hi | p1 lo];  A_lo x p0 hi, A_lo x p1 hi
        ct_mma_w(d, adesc(0, 0), wdesc(K::W_UP0, 4 * C, 0, ks), id_4c, ks ? 1u : 0u);
        ct_mma_w(d, adesc(1, 0), wdesc(K::W_UP0, 4 * C, 0, ks), id_c, 1u);
        ct_mma_w(d + 2 * C, adesc(1, 0), wdesc(K::W_UP0, 4 * C, 2 * C, ks), id_c, 1u);
        // row q-1: phase 0.  A_hi x [hi | lo], A_lo x hi
        ct_mma_w(d, adesc(0, -1), wdesc(K::W_UPM, 2 * C, 0, ks), id_2c, 1u);
        ct_mma_w(d, adesc(1, -1), wdesc(K::W_UPM, 2 * C, 0, ks), id_c, 1u);
        // row q+1: phase 1
        ct_mma_w(d + 2 * C, adesc(0, 1), wdesc(K::W_UPP, 2 * C, 0, ks), id_2c, 1u);
        ct_mma_w(d + 2 * C, adesc(1, 1), wdesc(K::W_UPP, 2 * C, 0, ks), id_c, 1u);
      }
      if (part == 0) return;
      ct_commit_w(bar(c, ACC_UP));
      if (OVL) ct_commit_w(bar(c, XS_FREE));       // the in-place split buffer may take the next raw tile
    };
    // ResBlock conv: D[i, (main|corr, co)] = sum_tap A[i + tap - 1, :] W_tap; 6 UMMAs per k-step
    auto issue_conv = [&](int c, int conv) {
      const uint32_t sA = sbase + (uint32_t)c * K::CTX + (conv == 0 ? K::O_U : K::O_V);
      const int wpart = conv == 0 ? K::W_C1 : K::W_C2;
#pragma unroll
      for (int h = 0; h < HALVES; ++h) {
        const uint32_t d = tmem_base + (uint32_t)(c * K::TCOLS_CTX + (conv == 0 ? K::T_C1 : K::T_C2) + h * 2 * C);
#pragma unroll
        for (int tap = 0; tap < 3; ++tap)
#pragma unroll
          for (int ks = 0; ks < C / 8; ++ks) {
            const uint32_t a_hi = sA + (uint32_t)(128 * h + tap) * ROWB + (uint32_t)ks * 32u;
            ct_mma_w(d, u_tmpl | (uint64_t)((a_hi >> 4) & 0x3FFFu), wdesc(wpart + tap * 2 * C * C, 2 * C, 0, ks), id_2c, (tap | ks) ? 1u : 0u);
            ct_mma_w(d, u_tmpl | (uint64_t)(((a_hi + K::UPL) >> 4) & 0x3FFFu), wdesc(wpart + tap * 2 * C * C, 2 * C, 0, ks), id_c, 1u);
          }
        ct_commit_w(bar(c, (conv == 0 ? ACC_C1 : ACC_C2) + h));
      }
      if (conv == 1 && !OVL) ct_commit_w(bar(c, XS_FREE));   // V (= the X split region) is dead once conv2 has run
    };
    if (OVL) {
      // one context: the transposed-conv GEMM of tile it+1 is issued in two halves, after conv1(it) and after
      // conv2(it), so that it runs under this tile's epilogues without holding up conv2 for its whole length
      if (tile_of(0, 0) < a.total_tiles) {
        ct_wait(bar(0, XS_FULL), 0u, dbg, 3, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_up(0, 2);
      }
      for (int it = 0; it < n_iter; ++it) {
        if (tile_of(it, 0) >= a.total_tiles) break;
        const uint32_t par = (uint32_t)(it & 1);
        ct_wait(bar(0, U_READY), par, dbg, 4, it);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_conv(0, 0);
        const bool more = it + 1 < n_iter && tile_of(it + 1, 0) < a.total_tiles;
        if (more && FS_UP_SPLIT) {
          ct_wait(bar(0, XS_FULL), par ^ 1u, dbg, 3, it + 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_up(0, 0);
        }
        ct_wait(bar(0, V_READY), par, dbg, 5, it);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_conv(0, 1);
        if (more) {
          if (!FS_UP_SPLIT) {
            ct_wait(bar(0, XS_FULL), par ^ 1u, dbg, 3, it + 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          issue_up(0, FS_UP_SPLIT ? 1 : 2);
        }
      }
    } else {
      for (int it = 0; it < n_iter; ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        for (int c = 0; c < NCTX; ++c) {
          if (tile_of(it, c) >= a.total_tiles) continue;
          ct_wait(bar(c, XS_FULL), par, dbg, 3, it);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_up(c, 2);
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
  } else if (warp < 4) {
    // ===== splitters: raw fp32 tile -> TF32 hi tile + exact remainder (flat: the swizzle is address preserving) =====
    const int sw = warp - 2;
    for (int it = 0; it < n_iter; ++it)
      for (int c = 0; c < NCTX; ++c) {
        if (tile_of(it, c) >= a.total_tiles) continue;
        ct_wait(bar(c, XRAW_FULL), (uint32_t)(it & 1), dbg, 6, it);
        if (!OVL && it > 0) ct_wait(bar(c, XS_FREE), (uint32_t)((it - 1) & 1), dbg, 7, it);
        __syncwarp();
        const float4* src = reinterpret_cast<const float4*>(gbase + (uint32_t)c * K::CTX + K::O_RAW);
        float4* dhi = reinterpret_cast<float4*>(gbase + (uint32_t)c * K::CTX + K::O_XS);     // == src when the split is in place
        float4* dlo = dhi + K::XRAW / 16;
        constexpr int N4 = K::XRAW / 16;
#pragma unroll 4
        for (int i = sw * 32 + lane; i < N4; i += 64) {
          const float4 v = src[i];
          float4 h, l;
          h.x = ct_hi(v.x); h.y = ct_hi(v.y); h.z = ct_hi(v.z); h.w = ct_hi(v.w);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          dhi[i] = h; dlo[i] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) { ct_arrive(bar(c, XS_FULL)); if (!OVL) ct_arrive(bar(c, XRAW_EMPTY)); }
      }
  } else {
    // ===== epilogue warpgroup g of context c: thread m owns TMEM lane m =====
    const int eg = (warp - 4) >> 2;                 // warpgroup index
    const int c = eg / G, g = eg % G;
    const int qtr = warp & 3;
    const int m = qtr * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qtr * 32) << 16) + (uint32_t)(c * K::TCOLS_CTX);
    uint8_t* Ub = gbase + (uint32_t)c * K::CTX + K::O_U;
    uint8_t* Vb = gbase + (uint32_t)c * K::CTX + K::O_V;
    float* exch = reinterpret_cast<float*>(gbase + K::OFF_EXCH) + c * 2 * K::UROWS;
    const float* b_up = consts, *b1 = consts + C, *b2 = consts + 2 * C, *ow = consts + 3 * C;
    constexpr int CG = C / G;                        // channels per warpgroup in EPI2 / EPI3
    const int cg0 = g * CG;
    for (int it = 0; it < n_iter; ++it) {
      const int gt = tile_of(it, c);
      if (gt >= a.total_tiles) break;
      const uint32_t par = (uint32_t)(it & 1);
      const int b = gt / a.tiles_per_utt, k = gt % a.tiles_per_utt;
      const int Qs = k * (K::NOUT / 2) - K::ILO / 2;
      const int Ts = 2 * Qs;                               // output position of U row 0

      // ---- EPI1: transposed-conv accumulator -> U = lrelu(. + bias), zero outside the utterance, hi/lo rows ----
      ct_wait(bar(c, ACC_UP), par, dbg, 8, it);
      const bool pr = a.prof != nullptr && blockIdx.x == 0 && eg == 0 && m == 0 && it < 64;
      if (pr) a.prof[it * 8 + 0] = clock64();
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (m < NQ) {                                        // warp-uniform (NQ is a multiple of 32)
        const int q = Qs + m;
        const float keep = (q >= 0 && q < a.L_in) ? 1.f : 0.f;
#pragma unroll
        for (int pp = 0; pp < 2 / G; ++pp) {
          const int p = G == 2 ? g : pp;                   // with two warpgroups each takes one phase
          const int row = 2 * m + p + 1;
#pragma unroll
          for (int c0 = 0; c0 < C; c0 += 16) {
            float v[16];
            fs_ld_sum16(t_lane + (uint32_t)(K::T_UP + p * 2 * C + c0), t_lane + (uint32_t)(K::T_UP + p * 2 * C + C + c0), v);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              float4 h, l;
              const float x0 = fs_lrelu(v[4 * j4 + 0] + b_up[c0 + 4 * j4 + 0]) * keep;
              const float x1 = fs_lrelu(v[4 * j4 + 1] + b_up[c0 + 4 * j4 + 1]) * keep;
              const float x2 = fs_lrelu(v[4 * j4 + 2] + b_up[c0 + 4 * j4 + 2]) * keep;
              const float x3 = fs_lrelu(v[4 * j4 + 3] + b_up[c0 + 4 * j4 + 3]) * keep;
              h.x = ct_hi(x0); h.y = ct_hi(x1); h.z = ct_hi(x2); h.w = ct_hi(x3);
              l.x = x0 - h.x; l.y = x1 - h.y; l.z = x2 - h.z; l.w = x3 - h.w;
              const uint32_t off = fs_swz<ROWB>(row, (c0 >> 2) + j4);
              *reinterpret_cast<float4*>(Ub + off) = h;
              *reinterpret_cast<float4*>(Ub + K::UPL + off) = l;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) ct_arrive(bar(c, U_READY));
      if (pr) a.prof[it * 8 + 1] = clock64();

      // ---- EPI2: conv1 accumulator -> V = lrelu(. + bias), zero outside the utterance, hi/lo rows ----
#pragma unroll
      for (int h = 0; h < HALVES; ++h) {
        ct_wait(bar(c, ACC_C1 + h), par, dbg, 9, it);
        if (pr && h == 0) a.prof[it * 8 + 2] = clock64();
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int i = 128 * h + m;
        const int t = Ts + i;
        const float keep = (t >= 0 && t < a.L_out) ? 1.f : 0.f;
#pragma unroll
        for (int c0 = cg0; c0 < cg0 + CG; c0 += 16) {
          float v[16];
          fs_ld_sum16(t_lane + (uint32_t)(K::T_C1 + h * 2 * C + c0), t_lane + (uint32_t)(K::T_C1 + h * 2 * C + C + c0), v);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            float4 hh, l;
            const float x0 = fs_lrelu(v[4 * j4 + 0] + b1[c0 + 4 * j4 + 0]) * keep;
            const float x1 = fs_lrelu(v[4 * j4 + 1] + b1[c0 + 4 * j4 + 1]) * keep;
            const float x2 = fs_lrelu(v[4 * j4 + 2] + b1[c0 + 4 * j4 + 2]) * keep;
            const float x3 = fs_lrelu(v[4 * j4 + 3] + b1[c0 + 4 * j4 + 3]) * keep;
            hh.x = ct_hi(x0); hh.y = ct_hi(x1); hh.z = ct_hi(x2); hh.w = ct_hi(x3);
            l.x = x0 - hh.x; l.y = x1 - hh.y; l.z = x2 - hh.z; l.w = x3 - hh.w;
            const uint32_t off = fs_swz<ROWB>(i + 1, (c0 >> 2) + j4);
            *reinterpret_cast<float4*>(Vb + off) = hh;
            *reinterpret_cast<float4*>(Vb + K::UPL + off) = l;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) ct_arrive(bar(c, V_READY));
      if (pr) a.prof[it * 8 + 3] = clock64();

      // ---- EPI3: conv2 accumulator + bias + U -> stage output (or the 1-channel output conv + tanh) ----
      float p1[HALVES];
#pragma unroll
      for (int h = 0; h < HALVES; ++h) {
        ct_wait(bar(c, ACC_C2 + h), par, dbg, 10, it);
        if (pr && h == 0) a.prof[it * 8 + 4] = clock64();
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int i = 128 * h + m;
        const int t = Ts + i;
        const bool inside = (t >= 0 && t < a.L_out);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
        for (int c0 = cg0; c0 < cg0 + CG; c0 += 16) {
          float y[16];
          fs_ld_sum16(t_lane + (uint32_t)(K::T_C2 + h * 2 * C + c0), t_lane + (uint32_t)(K::T_C2 + h * 2 * C + C + c0), y);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint32_t off = fs_swz<ROWB>(i + 1, (c0 >> 2) + j4);
            const float4 uh = *reinterpret_cast<const float4*>(Ub + off);
            const float4 ul = *reinterpret_cast<const float4*>(Ub + K::UPL + off);
            y[4 * j4 + 0] += b2[c0 + 4 * j4 + 0] + (uh.x + ul.x);
            y[4 * j4 + 1] += b2[c0 + 4 * j4 + 1] + (uh.y + ul.y);
            y[4 * j4 + 2] += b2[c0 + 4 * j4 + 2] + (uh.z + ul.z);
            y[4 * j4 + 3] += b2[c0 + 4 * j4 + 3] + (uh.w + ul.w);
          }
          if (!FINAL) {
            if (inside && i >= K::ILO && i < K::IHI) {
              float4* op = reinterpret_cast<float4*>(a.out + ((size_t)b * a.L_out + t) * C + c0);
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) op[j4] = make_float4(y[4 * j4], y[4 * j4 + 1], y[4 * j4 + 2], y[4 * j4 + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              d0 = fmaf(ow[c0 + j], y[j], d0);
              d1 = fmaf(ow[C + c0 + j], y[j], d1);
              d2 = fmaf(ow[2 * C + c0 + j], y[j], d2);
            }
          }
        }
        if (FINAL) {      // the output conv zero-pads y outside the utterance
          exch[i] = inside ? d0 : 0.f;
          exch[K::UROWS + i] = inside ? d2 : 0.f;
          p1[h] = d1;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      fs_group_sync(c, 128 * G);              // U reads done (the next EPI1 may overwrite it); exch complete
      if (FINAL) {
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          const int i = 128 * h + m;
          const int t = Ts + i;
          if (i >= K::ILO && i < K::IHI && t < a.L_out) {
            const float s = consts[6 * C] + exch[i - 1] + p1[h] + exch[K::UROWS + i + 1];
            a.out[(size_t)b * a.L_out + t] = tanhf(s);
          }
        }
        fs_group_sync(c, 128 * G);            // exch reads done before the next tile rewrites it
      }
      if (pr) a.prof[it * 8 + 5] = clock64();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- weight image: K-major no-swizzle core matrices; every part stacks hi rows and lo rows:
//      part [k/16][n/8][(k%16)/4][n%8][k%4] with n over NR rows (see FsCfg::W_*) ----
struct FsPackArgs { const float* up_w; const float* w1; const float* w2; float* blob; int C; };
__global__ void fs_wpack_kernel(FsPackArgs p) {
  const int C = p.C, CI = 2 * C;
  const int n_up0 = 4 * C * CI, n_upm = 2 * C * CI, n_conv = 2 * C * C;
  const int total = n_up0 + 2 * n_upm + 6 * n_conv;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int base, NR, Kd, e = idx;
    float v;
    int n, k, lo;
    if (e < n_up0) {                       // rows [p0 hi | p0 lo | p1 hi | p1 lo], each C rows
      base = 0; NR = 4 * C; Kd = CI; n = e / Kd; k = e % Kd;
      const int ph = n / (2 * C), co = n % C; lo = (n / C) & 1;
      v = p.up_w[((size_t)k * C + co) * 4 + ph + 1];
    } else if (e < n_up0 + 2 * n_upm) {    // row q-1 (kernel tap 3) then row q+1 (kernel tap 0): [hi | lo]
      e -= n_up0;
      const int which = e / n_upm; e -= which * n_upm;
      base = n_up0 + which * n_upm; NR = 2 * C; Kd = CI; n = e / Kd; k = e % Kd;
      const int co = n % C; lo = n / C;
      v = p.up_w[((size_t)k * C + co) * 4 + (which == 0 ? 3 : 0)];
    } else {                               // conv1 taps 0..2, conv2 taps 0..2: [hi | lo]
      e -= n_up0 + 2 * n_upm;
      const int part = e / n_conv; e -= part * n_conv;
      base = n_up0 + 2 * n_upm + part * n_conv; NR = 2 * C; Kd = C; n = e / Kd; k = e % Kd;
      const int co = n % C; lo = n / C;
      const float* w = part < 3 ? p.w1 : p.w2;
      v = w[((size_t)co * C + k) * 3 + (part % 3)];
    }
    const int off = ((k >> 4) * (NR * 64) + (n >> 3) * 512 + ((k & 15) >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4) >> 2;
    const float h = ct_hi(v);
    p.blob[base + off] = lo ? ct_hi(v - h) : h;
  }
}

typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn4 fs_encode_fn() {
  static EncodeTiledFn4 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn4)p;
  }
  return fn;
}

template <int C, int NQ, int NCTX, bool FINAL>
static int launch_fs(const float* x, FusedStageArgs a, int stage, cudaStream_t s) {
  using K = FsCfg<C, NQ, NCTX, FINAL>;
  EncodeTiledFn4 enc = fs_encode_fn();
  M2_REQUIRE(enc != nullptr, M2TTS_E_CUDA, "voc_fused: cuTensorMapEncodeTiled unavailable");
  CUtensorMap tmap;
  const cuuint64_t dims[3] = {(cuuint64_t)K::CI, (cuuint64_t)a.L_in, (cuuint64_t)a.B};
  const cuuint64_t strides[2] = {(cuuint64_t)K::CI * 4, (cuuint64_t)a.L_in * K::CI * 4};
  const cuuint32_t box[3] = {32u, (cuuint32_t)K::XR, 1u};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)x, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "voc_fused: cuTensorMapEncodeTiled failed (%d)", (int)r);
  a.tiles_per_utt = ceil_div(a.L_out, K::NOUT);
  a.total_tiles = a.B * a.tiles_per_utt;
  int grid = ceil_div(a.total_tiles, NCTX);
  if (grid > kNumSMs) grid = kNumSMs;
  M2_CUDA_OK(allow_smem(voc_stage_fused_kernel<C, NQ, NCTX, FINAL>, K::TOTAL));
  M2_LAUNCH(stage, (voc_stage_fused_kernel<C, NQ, NCTX, FINAL>), grid, K::THREADS, K::TOTAL, s, tmap, a, debug_words_device());
  return M2TTS_OK;
}

static long long* g_fs_prof = nullptr;
bool voc_fused_eligible(int C, int r, int dil) { return (C == 16 || C == 32) && r == 2 && dil == 1; }
size_t voc_fused_wblob_floats(int C) { return (size_t)2 * 14 * C * C; }

// x: channel-last [B][L_in][2C] (16-byte aligned); out: channel-last [B][2 L_in][C], or audio [B][2 L_in] when out_w != null.
int launch_voc_stage_fused(const float* x, const float* up_w, const float* up_b, const float* w1, const float* b1,
                           const float* w2, const float* b2, const float* out_w, const float* out_b, float* wblob,
                           float* out, int B, int C, int L_in, int stage, cudaStream_t s) {
  M2_REQUIRE(C == 16 || C == 32, M2TTS_E_UNSUPPORTED, "voc_fused: C=%d (16 or 32)", C);
  if (up_w != nullptr) {      // (re)write the weight image; up_w == nullptr: wblob already holds it
    M2_REQUIRE(w1 != nullptr && w2 != nullptr && (((uintptr_t)wblob) & 15) == 0, M2TTS_E_BADSHAPE, "voc_fused: pack arguments");
    FsPackArgs p{up_w, w1, w2, wblob, C};
    M2_LAUNCH(M2TTS_STAGE_PACK, fs_wpack_kernel, ceil_div(28 * C * C, 256), 256, 0, s, p);
  }
  if (x == nullptr) return M2TTS_OK;      // pack only
  M2_REQUIRE((((uintptr_t)x) & 15) == 0 && (((uintptr_t)out) & 15) == 0 && (((uintptr_t)wblob) & 15) == 0, M2TTS_E_BADSHAPE,
             "voc_fused: misaligned pointers");
  M2_REQUIRE(B > 0 && L_in > 0 && (long long)B * L_in * 2 < (1ll << 31), M2TTS_E_BADSHAPE, "voc_fused: B=%d L=%d", B, L_in);
  FusedStageArgs a{};
  a.B = B; a.L_in = L_in; a.L_out = 2 * L_in; a.wblob = wblob; a.bias_up = up_b; a.bias1 = b1; a.bias2 = b2;
  a.out_w = out_w; a.out_b = out_b; a.out = out; a.prof = g_fs_prof;
  const bool fin = out_w != nullptr;
  if (C == 16) return fin ? launch_fs<16, 128, 2, true>(x, a, stage, s) : launch_fs<16, 128, 2, false>(x, a, stage, s);
  return fin ? launch_fs<32, 64, 1, true>(x, a, stage, s) : launch_fs<32, 64, 1, false>(x, a, stage, s);
}

}  // namespace m2

using namespace m2;

// bring-up: device buffer of 64 x 8 int64 that receives phase timestamps (NULL = off)
#ifdef M2TTS_TOOLS
extern "C" int m2tts_vocoder_stage_fused_set_prof(long long* dev_buf) { m2::g_fs_prof = dev_buf; return M2TTS_OK; }
#endif

extern "C" size_t m2tts_vocoder_stage_fused_workspace_bytes(int C) {
  if (C != 16 && C != 32) return 0;
  return align_up(voc_fused_wblob_floats(C) * sizeof(float), 256) + 256;
}

extern "C" int m2tts_vocoder_stage_fused(const float* x, const float* up_w, const float* up_b, const float* res1_w,
                                         const float* res1_b, const float* res2_w, const float* res2_b,
                                         const float* out_w, const float* out_b, float* y, int B, int C, int L,
                                         void* workspace, size_t workspace_bytes, m2tts_stream_t stream) {
  M2_REQUIRE(x && up_w && up_b && res1_w && res1_b && res2_w && res2_b && y && workspace, M2TTS_E_NULLPTR,
             "vocoder_stage_fused: null pointer");
  M2_REQUIRE((out_w == nullptr) == (out_b == nullptr), M2TTS_E_NULLPTR, "vocoder_stage_fused: out_w/out_b must both be set or both null");
  Carver cv(workspace, workspace_bytes);
  float* wblob = cv.take<float>(voc_fused_wblob_floats(C == 16 || C == 32 ? C : 16));
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "vocoder_stage_fused: workspace too small or misaligned");
  return launch_voc_stage_fused(x, up_w, up_b, res1_w, res1_b, res2_w, res2_b, out_w, out_b, wblob, y, B, C, L,
                                M2TTS_STAGE_VOC_RES1, (cudaStream_t)stream);
}
