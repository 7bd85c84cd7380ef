// conv_tc.cu — vocoder convolutions as implicit GEMM on tcgen05/TMEM fed by TMA ("tap-GEMM"),
// fp32-faithful through 3xTF32 splitting (see attention_tc.cu for the error model).
//
//   D[m, n] = sum_taps sum_ci  X[ci, m + shift_tap] * W_tap[n, ci]
//
//   Conv1d k=3 (dilation d):   taps (-d, 0, +d), W_tap[n=co, ci] = w[co, ci, tap]      (tts_model.py:246, components.py:181-190)
//   ConvTranspose1d k=2r, s=r: polyphase — output sample r*q+p of channel co is D[q, p*ct + co]:
//        tap 0 (shift 0):  all r phases,      W[(p,co), ci] = w[ci, co, p + r/2]
//        tap 1 (shift -1): phases p <  r/2,   W[(p,co), ci] = w[ci, co, p + r/2 + r]
//        tap 2 (shift +1): phases p >= r/2,   W[(p,co), ci] = w[ci, co, p - r/2]         (tts_model.py:255-263)
//
// Operands: activations live in HBM as two planes (hi, lo) [2][B][C][Lp], channel-first with
// positions contiguous, so a TMA box {32 positions x 16 channels} lands in shared memory as the
// MN-major 128B-swizzle/32B-atom UMMA layout (the only MN-major layout tcgen05 takes for 32-bit
// operands); the tap shift is just the box's column coordinate and TMA's out-of-bounds zero fill IS
// the convolution's zero padding. Weights are pre-packed per (channel tile, 16-channel chunk) as
// the exact shared-memory image of a K-major no-swizzle operand (hi and lo planes) and arrive with
// one cp.async.bulk per chunk. CTA = 128 threads: lane 0 of warp 0 produces (TMA), lane 0 of
// warp 1 issues UMMAs (M128, N = rows of the tap, K8; 3 split terms), all four warps run the
// epilogue (TMEM -> bias/activation/residual -> global, hi/lo planes or plain fp32).
#include "common.cuh"
#include "conv_tc.cuh"
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace m2 {

// (constants, TapGemmArgs and the PTX helpers live in conv_tc.cuh)

// D_tap[m, n] = sum_ci X[ci, start + m] * W_tap[n, ci] for the three taps (separate TMEM column ranges);
// the epilogue forms out[t] = sum_tap D_tap[t - start + shift_tap] through a shared-memory staging tile.
__global__ void __launch_bounds__(CT_THREADS, 2)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const TapGemmArgs a, int* dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ct_smem_u32(smem_raw) + 1023u) & ~1023u;
  float* stage_f = reinterpret_cast<float*>(smem_raw + (sbase - ct_smem_u32(smem_raw)));   // epilogue staging (reuses the ring)
  const uint32_t w_plane_bytes = (uint32_t)a.rows_total * 64u;
  const uint32_t w_stage = 2u * w_plane_bytes;
  const uint32_t stage_bytes = CT_A_STAGE + ((w_stage + 1023u) & ~1023u);
  const uint32_t sBar = sbase + CT_STAGES * stage_bytes;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * CT_STAGES, bar_acc = sBar + 16 * CT_STAGES;
  const uint32_t tmem_slot = bar_acc + 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int start = blockIdx.x * CT_STEP - CT_HALO;   // first input position of this tile (multiple of 4)
  const int ntile = blockIdx.y, b = blockIdx.z;

  if (tid == 0) {
    for (int s = 0; s < CT_STAGES; ++s) { ct_mbar_init(bar_full + 8 * s, 1); ct_mbar_init(bar_empty + 8 * s, 1); }
    ct_mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
  }
  if (warp == 0) {
    __syncwarp();   // lane 0 just left the barrier-init branch; .sync.aligned needs the whole warp converged
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0 && lane == 0) {
    // ===== producer: 8 TMA boxes (hi/lo x 4 x 32 positions) + one bulk copy of the weight image per chunk =====
    const float* wsrc = a.wblob + (size_t)ntile * a.n_chunks * (size_t)(2 * a.rows_total * 16);
    for (int c = 0; c < a.n_chunks; ++c) {
      const int s = c % CT_STAGES;
      if (c >= CT_STAGES) ct_wait(bar_empty + 8 * s, (uint32_t)((c / CT_STAGES - 1) & 1), dbg, 1, c);
      const uint32_t sA = sbase + s * stage_bytes, sW = sA + CT_A_STAGE, full = bar_full + 8 * s;
      ct_expect_tx(full, CT_A_STAGE + w_stage);
#pragma unroll
      for (int plane = 0; plane < 2; ++plane) {
        const int row = (plane * a.B + b) * a.CI + c * CT_CK;
#pragma unroll
        for (int x = 0; x < 4; ++x)
          ct_tma_2d(sA + (uint32_t)(plane * 4 + x) * CT_ABOX, &tmap_a, start + 32 * x, row, full);
      }
      ct_bulk(sW, wsrc + (size_t)c * (2 * a.rows_total * 16), w_stage, full);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== UMMA issuer =====
    for (int c = 0; c < a.n_chunks; ++c) {
      const int s = c % CT_STAGES;
      ct_wait(bar_full + 8 * s, (uint32_t)((c / CT_STAGES) & 1), dbg, 2, c);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sA = sbase + s * stage_bytes, sW = sA + CT_A_STAGE;
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
            ct_mma(tmem_base + (uint32_t)a.tap_dcol[tap], ad, bd, idesc, (c | term | ks) ? 1u : 0u);
          }
        }
      }
      ct_commit(bar_empty + 8 * s);
    }
    ct_commit(bar_acc);
  }
  __syncwarp();

  // ===== epilogue: thread = GEMM row m = input position start + m =====
  ct_wait(bar_acc, 0, dbg, 3, 0);
  __syncwarp();     // threads leave the polling loop one by one; tcgen05.ld is .sync.aligned
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int m = tid;
  const int q = start + m;
  const bool own = (m >= CT_HALO) && (m < CT_BM - CT_HALO) && (q < a.L_in);
  const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
  const int co0 = ntile * a.co_tile;

  if (a.r == 1) {
    // staging tile S[tap][32 cols][128 rows]
    const int s0 = a.tap_shift[0], s1 = a.tap_shift[1], s2 = a.tap_shift[2];
    for (int c0 = 0; c0 < a.co_tile; c0 += 32) {
#pragma unroll
      for (int tap = 0; tap < 3; ++tap)
#pragma unroll
        for (int cc = 0; cc < 32; cc += 8) {
          uint32_t v[8];
          ct_ld8(t_lane + (uint32_t)(a.tap_dcol[tap] + c0 + cc), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) stage_f[((tap * 32 + cc + j) << 7) + m] = __uint_as_float(v[j]);
        }
      __syncthreads();
      if (own) {
        // residual loads first (read-only path, independent of the stores below) so their latency overlaps
        float rsd[32];
        if (a.res_hi != nullptr) {
          const float* __restrict__ rh = a.res_hi + ((size_t)b * a.CO + co0 + c0) * a.Lp_res + q;
          const float* __restrict__ rl = a.res_lo + ((size_t)b * a.CO + co0 + c0) * a.Lp_res + q;
#pragma unroll
          for (int c = 0; c < 32; ++c) rsd[c] = __ldg(rh + (size_t)c * a.Lp_res) + __ldg(rl + (size_t)c * a.Lp_res);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) rsd[c] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int co = co0 + c0 + c;
          float x = stage_f[((c) << 7) + m + s0] + stage_f[((32 + c) << 7) + m + s1] + stage_f[((64 + c) << 7) + m + s2] +
                    __ldg(a.bias + co);
          if (a.act == 1) x = x > 0.f ? x : 0.1f * x;
          x += rsd[c];
          const size_t oo = ((size_t)b * a.CO + co) * a.Lp_out + q;
          if (a.out_lo != nullptr) { const float h = ct_hi(x); a.out_hi[oo] = h; a.out_lo[oo] = ct_hi(x - h); }
          else a.out_hi[oo] = x;
        }
      }
      __syncthreads();
    }
  } else {
    // transposed conv (r == 4): accumulator columns D0 [0,4ct) = (phase, channel), D1 [4ct,6ct) phases 0-1 (needs row q-1),
    // D2 [6ct,8ct) phases 2-3 (needs row q+1). Staging tile S[64][128]: 0-31 D0 (p*8+j), 32-47 D1, 48-63 D2.
    const int ct = a.co_tile;
    for (int c0 = 0; c0 < ct; c0 += 8) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        // g 0..3: D0 phase g; 4..5: D1 phase g-4; 6..7: D2 phase g-6 (+2)
        const int col = (g < 4) ? g * ct : (g < 6 ? 4 * ct + (g - 4) * ct : 6 * ct + (g - 6) * ct);
        uint32_t v[8];
        ct_ld8(t_lane + (uint32_t)(col + c0), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) stage_f[((g * 8 + j) << 7) + m] = __uint_as_float(v[j]);
      }
      __syncthreads();
      if (own) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int co = co0 + c0 + j;
          const float bv = __ldg(a.bias + co);
          float x[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float t = stage_f[((p * 8 + j) << 7) + m] + bv;
            t += (p < 2) ? stage_f[(((4 + p) * 8 + j) << 7) + m - 1] : stage_f[(((6 + p - 2) * 8 + j) << 7) + m + 1];
            x[p] = t > 0.f ? t : 0.1f * t;
          }
          const size_t oo = ((size_t)b * a.CO + co) * a.Lp_out + (size_t)4 * q;
          if (a.out_lo != nullptr) {
            float h[4], l[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) { h[p] = ct_hi(x[p]); l[p] = ct_hi(x[p] - h[p]); }
            *reinterpret_cast<float4*>(a.out_hi + oo) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(a.out_lo + oo) = make_float4(l[0], l[1], l[2], l[3]);
          } else {
            *reinterpret_cast<float4*>(a.out_hi + oo) = make_float4(x[0], x[1], x[2], x[3]);
          }
        }
      }
      __syncthreads();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
}

// ---- weight packing: build the shared-memory images (K-major, no swizzle: 8x16-byte core matrices) ----
struct WPackArgs {
  const float* w; float* blob;
  int mode;     // 0: Conv1d weight [CO][CI][3], 1: ConvTranspose1d weight [CI][CO][2r]
  int CI, CO, r, co_tile, rows_total, n_chunks, n_tiles;
};
__global__ void tc_wpack_kernel(WPackArgs p) {
  const size_t img = (size_t)p.rows_total * 16;
  const size_t total = (size_t)p.n_tiles * p.n_chunks * 2 * img;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t e = idx % img;
    const int plane = (int)((idx / img) & 1);
    const int chunk = (int)((idx / (2 * img)) % p.n_chunks);
    const int ntile = (int)(idx / (2 * img * p.n_chunks));
    const int g = (int)(e >> 7), rem = (int)(e & 127);
    const int n = g * 8 + ((rem & 31) >> 2), k = (rem >> 5) * 4 + (rem & 3);
    const int ci = chunk * CT_CK + k;
    const int ct = p.co_tile;
    float v;
    if (p.mode == 0) {
      const int tap = n / ct, co = ntile * ct + n % ct;
      v = p.w[((size_t)co * p.CI + ci) * 3 + tap];
    } else {
      const int r = p.r;
      int kk, c;
      if (n < r * ct) { kk = n / ct + r / 2; c = n % ct; }
      else if (n < r * ct + (r / 2) * ct) { const int n2 = n - r * ct; kk = n2 / ct + r / 2 + r; c = n2 % ct; }
      else { const int n2 = n - r * ct - (r / 2) * ct; kk = n2 / ct; c = n2 % ct; }
      v = p.w[((size_t)ci * p.CO + ntile * ct + c) * (2 * r) + kk];
    }
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    p.blob[idx] = plane == 0 ? h : __uint_as_float(__float_as_uint(v - h) & 0xFFFFE000u);
  }
}

// plain fp32 [B][C][L] (pitch L) -> planes [2][B][C][Lp]
__global__ void tc_split_planes_kernel(const float* __restrict__ x, float* __restrict__ planes, long long rows, int L, int Lp) {
  const long long total = rows * Lp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Lp;
    const int t = (int)(i - row * Lp);
    const float v = t < L ? x[row * L + t] : 0.f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    planes[i] = h;
    planes[total + i] = __uint_as_float(__float_as_uint(v - h) & 0xFFFFE000u);
  }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn2 ct_encode_fn() {
  static EncodeTiledFn2 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn2)p;
  }
  return fn;
}

static int next_pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }

static bool tapgemm_v1() { static int v = -1; if (v < 0) { const char* e = getenv("M2TTS_TAPGEMM"); v = (e && strcmp(e, "v1") == 0) ? 1 : 0; } return v == 1; }
// dilation <= 4 is checked at launch. The persistent kernel takes any CO that is a multiple of 16 (<= 64 or a
// multiple of 64); the first-generation kernel (M2TTS_TAPGEMM=v1) needs CO % 64 == 0.
bool conv3_tc_eligible(int CI, int CO) {
  if (CI % CT_CK != 0 || CO % 16 != 0 || CO < 16) return false;
  if (tapgemm_v1()) return CO % 64 == 0;
  return CO % 64 == 0 || CO < 64;
}
bool convT_tc_eligible(int CI, int CO, int r) {
  if (CI % CT_CK != 0) return false;
  if (r == 4) return CO % 32 == 0 && CO >= 32;
  if (r == 2 && !tapgemm_v1()) return CO % 16 == 0 && CO >= 16;
  return false;
}
size_t conv3_tc_wblob_floats(int CI, int CO) { return (size_t)2 * 3 * CO * CI; }
size_t convT_tc_wblob_floats(int CI, int CO, int r) { return (size_t)2 * 2 * r * CO * CI; }

static int launch_tapgemm(const float* x_planes, int Lp_in, TapGemmArgs& a, int n_tiles, int stage, cudaStream_t s) {
  EncodeTiledFn2 enc = ct_encode_fn();
  M2_REQUIRE(enc != nullptr, M2TTS_E_CUDA, "conv_tc: cuTensorMapEncodeTiled unavailable");
  M2_REQUIRE((Lp_in & 3) == 0 && (((uintptr_t)x_planes) & 15) == 0, M2TTS_E_BADSHAPE, "conv_tc: input planes misaligned");
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)a.L_in, (cuuint64_t)2 * a.B * a.CI};
  const cuuint64_t strides[1] = {(cuuint64_t)Lp_in * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)CT_CK};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)x_planes, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  a.n_tiles = n_tiles;
  if (!tapgemm_v1()) return launch_tapgemm_persistent(tmap, a, stage, s);
  a.tmem_cols = next_pow2_cols(a.n_cols);
  const uint32_t w_stage = 2u * (uint32_t)a.rows_total * 64u;
  const size_t smem = (size_t)CT_STAGES * (CT_A_STAGE + ((w_stage + 1023u) & ~1023u)) + 1024 + 128;
  M2_REQUIRE(smem <= 227 * 1024, M2TTS_E_UNSUPPORTED, "conv_tc: tile needs %zu B of shared memory", smem);
  M2_CUDA_OK(allow_smem(tapgemm_kernel, smem));
  for (int j = 0; j < 3; ++j)
    M2_REQUIRE(a.tap_shift[j] >= -CT_HALO && a.tap_shift[j] <= CT_HALO, M2TTS_E_UNSUPPORTED, "conv_tc: tap shift %d exceeds the halo", a.tap_shift[j]);
  dim3 grid(ceil_div(a.L_in, CT_STEP), n_tiles, a.B);
  M2_LAUNCH(stage, tapgemm_kernel, grid, CT_THREADS, smem, s, tmap, a, debug_words_device());
  return M2TTS_OK;
}

static int launch_wpack(const WPackArgs& p, cudaStream_t s) {
  const size_t total = (size_t)p.n_tiles * p.n_chunks * 2 * p.rows_total * 16;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  M2_LAUNCH(M2TTS_STAGE_PACK, tc_wpack_kernel, grid, 256, 0, s, p);
  return M2TTS_OK;
}

// Conv1d k=3 on planes. wblob: conv3_tc_wblob_floats(CI,CO) floats of scratch.
int launch_conv3_tc(const float* x_planes, int Lp_in, const float* w, float* wblob, const float* bias,
                    const float* res_hi, const float* res_lo, int Lp_res, float* out_hi, float* out_lo, int Lp_out,
                    int B, int CI, int CO, int L, int dil, int act, int stage, cudaStream_t s) {
  M2_REQUIRE(conv3_tc_eligible(CI, CO), M2TTS_E_UNSUPPORTED, "conv3_tc: CI=%d CO=%d not eligible", CI, CO);
  const int ct = (CO % 64 == 0) ? 64 : CO, n_tiles = CO / ct;   // <= 192 accumulator columns, double-buffered in TMEM
  WPackArgs p{w, wblob, 0, CI, CO, 1, ct, 3 * ct, CI / CT_CK, n_tiles};
  int rc = launch_wpack(p, s);
  if (rc) return rc;
  TapGemmArgs a{};
  a.CI = CI; a.L_in = L; a.B = B; a.n_chunks = CI / CT_CK;
  for (int j = 0; j < 3; ++j) { a.tap_shift[j] = (j - 1) * dil; a.tap_rows[j] = ct; a.tap_wrow[j] = j * ct; a.tap_dcol[j] = j * ct; }
  a.rows_total = 3 * ct; a.n_cols = 3 * ct; a.wblob = wblob; a.r = 1; a.co_tile = ct; a.CO = CO; a.L_out = L; a.Lp_out = Lp_out;
  a.bias = bias; a.act = act; a.res_hi = res_hi; a.res_lo = res_lo; a.Lp_res = Lp_res; a.out_hi = out_hi; a.out_lo = out_lo;
  return launch_tapgemm(x_planes, Lp_in, a, n_tiles, stage, s);
}

// ConvTranspose1d(k=2r, stride r, pad r/2) + leaky_relu on planes, r == 4.
int launch_convT_tc(const float* x_planes, int Lp_in, const float* w, float* wblob, const float* bias, float* out_hi,
                    float* out_lo, int Lp_out, int B, int CI, int CO, int L, int r, cudaStream_t s) {
  M2_REQUIRE(convT_tc_eligible(CI, CO, r), M2TTS_E_UNSUPPORTED, "convT_tc: CI=%d CO=%d r=%d not eligible", CI, CO, r);
  const int ct = (CO % 32 == 0) ? 32 : 16, n_tiles = CO / ct;
  WPackArgs p{w, wblob, 1, CI, CO, r, ct, 2 * r * ct, CI / CT_CK, n_tiles};
  int rc = launch_wpack(p, s);
  if (rc) return rc;
  TapGemmArgs a{};
  a.CI = CI; a.L_in = L; a.B = B; a.n_chunks = CI / CT_CK;
  a.tap_shift[0] = 0;  a.tap_rows[0] = r * ct;       a.tap_wrow[0] = 0;                        a.tap_dcol[0] = 0;
  a.tap_shift[1] = -1; a.tap_rows[1] = (r / 2) * ct; a.tap_wrow[1] = r * ct;                   a.tap_dcol[1] = r * ct;
  a.tap_shift[2] = 1;  a.tap_rows[2] = (r / 2) * ct; a.tap_wrow[2] = r * ct + (r / 2) * ct;    a.tap_dcol[2] = r * ct + (r / 2) * ct;
  a.rows_total = 2 * r * ct; a.n_cols = 2 * r * ct; a.wblob = wblob; a.r = r; a.co_tile = ct; a.CO = CO;
  a.L_out = r * L; a.Lp_out = Lp_out; a.bias = bias; a.act = 1; a.out_hi = out_hi; a.out_lo = out_lo;
  return launch_tapgemm(x_planes, Lp_in, a, n_tiles, M2TTS_STAGE_VOC_UP, s);
}

int launch_split_planes(const float* x, float* planes, long long rows, int L, int Lp, cudaStream_t s) {
  long long total = rows * Lp;
  int grid = (int)((total + 255) / 256 > kNumSMs * 16 ? kNumSMs * 16 : (total + 255) / 256);
  M2_LAUNCH(M2TTS_STAGE_PACK, tc_split_planes_kernel, grid, 256, 0, s, x, planes, rows, L, Lp);
  return M2TTS_OK;
}

}  // namespace m2

using namespace m2;

// ---- stand-alone entry points (unit tests / users with plain fp32 tensors) ---------------------------
extern "C" size_t m2tts_conv_tc_workspace_bytes(int B, int CI, int CO, int L, int r) {
  if (B <= 0 || CI <= 0 || CO <= 0 || L <= 0 || r <= 0) return 0;
  const size_t Lp = (size_t)((L + 3) & ~3);
  const size_t planes = 2 * (size_t)B * (CI + CO) * Lp;   // input planes (+ residual planes)
  const size_t wb = (size_t)2 * 2 * (r > 3 ? r : 3) * CO * CI;
  return (planes + wb) * sizeof(float) + 8 * 256;
}

extern "C" int m2tts_conv1d_k3_tc(const float* x, const float* w, const float* bias, const float* residual, float* y,
                                  int B, int CI, int CO, int L, int dilation, int act, void* workspace,
                                  size_t workspace_bytes, m2tts_stream_t stream) {
  M2_REQUIRE(x && w && bias && y && workspace, M2TTS_E_NULLPTR, "conv1d_k3_tc: null pointer");
  M2_REQUIRE(B > 0 && L > 0 && dilation >= 1 && (act == 0 || act == 1), M2TTS_E_BADSHAPE, "conv1d_k3_tc: bad arguments");
  M2_REQUIRE(conv3_tc_eligible(CI, CO), M2TTS_E_UNSUPPORTED, "conv1d_k3_tc: needs CI %% 16 == 0 and CO a multiple of 16 that is < 64 or a multiple of 64 (CI=%d CO=%d)", CI, CO);
  const int Lp = (L + 3) & ~3;
  Carver cv(workspace, workspace_bytes);
  float* planes = cv.take<float>(2 * (size_t)B * CI * Lp);
  float* rplanes = cv.take<float>(residual ? 2 * (size_t)B * CO * Lp : 0);
  float* wblob = cv.take<float>(conv3_tc_wblob_floats(CI, CO));
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "conv1d_k3_tc: workspace too small or misaligned");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_split_planes(x, planes, (long long)B * CI, L, Lp, s);
  if (rc) return rc;
  const float* rh = nullptr; const float* rl = nullptr;
  if (residual != nullptr) {
    if ((rc = launch_split_planes(residual, rplanes, (long long)B * CO, L, Lp, s))) return rc;
    rh = rplanes; rl = rplanes + (size_t)B * CO * Lp;
  }
  return launch_conv3_tc(planes, Lp, w, wblob, bias, rh, rl, Lp, y, nullptr, L, B, CI, CO, L, dilation, act,
                         M2TTS_STAGE_VOC_RES1, s);
}

extern "C" int m2tts_conv_transpose1d_lrelu_tc(const float* x, const float* w, const float* bias, float* y, int B, int CI,
                                               int CO, int L, int r, void* workspace, size_t workspace_bytes,
                                               m2tts_stream_t stream) {
  M2_REQUIRE(x && w && bias && y && workspace, M2TTS_E_NULLPTR, "conv_transpose1d_tc: null pointer");
  M2_REQUIRE(B > 0 && L > 0, M2TTS_E_BADSHAPE, "conv_transpose1d_tc: bad arguments");
  M2_REQUIRE(convT_tc_eligible(CI, CO, r), M2TTS_E_UNSUPPORTED,
             "conv_transpose1d_tc: needs r == 4 (CO %% 32 == 0) or r == 2 (CO %% 16 == 0), CI %% 16 == 0 (CI=%d CO=%d r=%d)", CI, CO, r);
  const int Lp = (L + 3) & ~3;
  Carver cv(workspace, workspace_bytes);
  float* planes = cv.take<float>(2 * (size_t)B * CI * Lp);
  float* wblob = cv.take<float>(convT_tc_wblob_floats(CI, CO, r));
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "conv_transpose1d_tc: workspace too small or misaligned");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_split_planes(x, planes, (long long)B * CI, L, Lp, s);
  if (rc) return rc;
  return launch_convT_tc(planes, Lp, w, wblob, bias, y, nullptr, r * L, B, CI, CO, L, r, s);
}
