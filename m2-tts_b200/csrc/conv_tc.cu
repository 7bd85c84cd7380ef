// conv_tc.cu — vocoder convolutions as implicit GEMM on tcgen05/TMEM fed by TMA ("tap-GEMM"),
// fp32-faithful through 3xTF32 splitting (see attention_tc.cu for the error model). This file: weight packing,
// tensor maps, launch parameters and the C-ABI entry points; the kernel itself is conv_tc2.cu.
//
//   D_tap[m, n] = sum_ci X[ci, m] * W_tap[n, ci]            (one accumulator range per tap, M = 128 positions)
//   out[t]      = sum_tap D_tap[t + shift_tap]              (taps shift accumulator ROWS in the epilogue: TMA box
//                                                            rows must start 16-B aligned, so inputs cannot shift)
//
//   Conv1d k=3 (dilation d <= 4): taps (-d, 0, +d), W_tap[n=co, ci] = w[co, ci, tap]   (tts_model.py:246, components.py:181-190)
//   ConvTranspose1d k=2r, s=r:    polyphase — output sample r*q+p of channel co:
//        tap 0 (row q):    all r phases,     W[(p,co), ci] = w[ci, co, p + r/2]
//        tap 1 (row q-1):  phases p <  r/2,  W[(p,co), ci] = w[ci, co, p + r/2 + r]
//        tap 2 (row q+1):  phases p >= r/2,  W[(p,co), ci] = w[ci, co, p - r/2]          (tts_model.py:255-263)
//
// Activations are plain fp32 [B][C][L] (channel-first, positions contiguous): a TMA box {32 positions x 16 channels}
// with the 128B-swizzle/32B-atom mode lands as the MN-major UMMA layout (the only MN-major layout tcgen05 takes for
// 32-bit operands); out-of-bounds zero fill IS the convolution's zero padding. Weights are pre-packed per
// (channel tile, 16-channel chunk) as the exact shared-memory image of a K-major no-swizzle operand (hi and lo).
#include "common.cuh"
#include "conv_tc.cuh"
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace m2 {

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

// Any channel count that is a multiple of 16 (output: < 64 or a multiple of 64); dilation <= 4 is checked at launch.
bool conv3_tc_eligible(int CI, int CO) {
  return CI % CT_CK == 0 && CO % 16 == 0 && CO >= 16 && (CO % 64 == 0 || CO < 64);
}
bool convT_tc_eligible(int CI, int CO, int r) {
  if (CI % CT_CK != 0) return false;
  if (r == 4) return CO % 32 == 0 && CO >= 32;
  if (r == 2) return CO % 16 == 0 && CO >= 16;
  return false;
}
size_t conv3_tc_wblob_floats(int CI, int CO) { return (size_t)2 * 3 * CO * CI; }
size_t convT_tc_wblob_floats(int CI, int CO, int r) { return (size_t)2 * 2 * r * CO * CI; }

// x: plain fp32 [B][CI][Lp_in] (Lp_in % 4 == 0, 16-B aligned)
static int launch_tapgemm(const float* x, int Lp_in, TapGemmArgs& a, int n_tiles, int stage, cudaStream_t s) {
  EncodeTiledFn2 enc = ct_encode_fn();
  M2_REQUIRE(enc != nullptr, M2TTS_E_CUDA, "conv_tc: cuTensorMapEncodeTiled unavailable");
  M2_REQUIRE((Lp_in & 3) == 0 && (((uintptr_t)x) & 15) == 0, M2TTS_E_BADSHAPE, "conv_tc: input needs a row pitch that is a multiple of 4 floats and 16-B alignment");
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)a.L_in, (cuuint64_t)a.B * a.CI};
  const cuuint64_t strides[1] = {(cuuint64_t)Lp_in * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)CT_CK};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)x, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  a.n_tiles = n_tiles;
  return launch_tapgemm_persistent(tmap, a, stage, s);
}

static int launch_wpack(const WPackArgs& p, cudaStream_t s) {
  const size_t total = (size_t)p.n_tiles * p.n_chunks * 2 * p.rows_total * 16;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  M2_LAUNCH(M2TTS_STAGE_PACK, tc_wpack_kernel, grid, 256, 0, s, p);
  return M2TTS_OK;
}

// Conv1d k=3 on plain fp32. wblob: conv3_tc_wblob_floats(CI,CO) floats of scratch. residual (plain, pitch Lp_res) may be null.
int launch_conv3_tc(const float* x, int Lp_in, const float* w, float* wblob, const float* bias, const float* residual,
                    int Lp_res, float* out, int Lp_out, int B, int CI, int CO, int L, int dil, int act, int stage,
                    cudaStream_t s, int out_cl, int32_t* status) {
  M2_REQUIRE(conv3_tc_eligible(CI, CO), M2TTS_E_UNSUPPORTED, "conv3_tc: CI=%d CO=%d not eligible", CI, CO);
  const int ct = (CO % 64 == 0) ? 64 : CO, n_tiles = CO / ct;   // <= 192 accumulator columns per tile
  if (w != nullptr) {      // (re)write the weight image; w == nullptr: wblob already holds it
    WPackArgs p{w, wblob, 0, CI, CO, 1, ct, 3 * ct, CI / CT_CK, n_tiles};
    int rc = launch_wpack(p, s);
    if (rc) return rc;
  }
  if (x == nullptr) return M2TTS_OK;      // pack only
  TapGemmArgs a{};
  a.status = status;
  a.CI = CI; a.L_in = L; a.B = B; a.n_chunks = CI / CT_CK;
  for (int j = 0; j < 3; ++j) { a.tap_shift[j] = (j - 1) * dil; a.tap_rows[j] = ct; a.tap_wrow[j] = j * ct; a.tap_dcol[j] = j * ct; }
  a.rows_total = 3 * ct; a.n_cols = 3 * ct; a.wblob = wblob; a.r = 1; a.co_tile = ct; a.CO = CO; a.L_out = L; a.Lp_out = Lp_out;
  a.bias = bias; a.act = act; a.residual = residual; a.Lp_res = Lp_res; a.out = out; a.out_cl = out_cl;
  return launch_tapgemm(x, Lp_in, a, n_tiles, stage, s);
}

// ConvTranspose1d(k=2r, stride r, pad r/2) + leaky_relu on plain fp32, r in {2,4}.
int launch_convT_tc(const float* x, int Lp_in, const float* w, float* wblob, const float* bias, float* out, int Lp_out,
                    int B, int CI, int CO, int L, int r, cudaStream_t s, int out_cl, int32_t* status) {
  M2_REQUIRE(convT_tc_eligible(CI, CO, r), M2TTS_E_UNSUPPORTED, "convT_tc: CI=%d CO=%d r=%d not eligible", CI, CO, r);
  const int ct = (CO % 32 == 0) ? 32 : 16, n_tiles = CO / ct;
  if (w != nullptr) {      // (re)write the weight image; w == nullptr: wblob already holds it
    WPackArgs p{w, wblob, 1, CI, CO, r, ct, 2 * r * ct, CI / CT_CK, n_tiles};
    int rc = launch_wpack(p, s);
    if (rc) return rc;
  }
  if (x == nullptr) return M2TTS_OK;      // pack only
  M2_REQUIRE((Lp_out % r) == 0 && (Lp_out & 3) == 0, M2TTS_E_BADSHAPE, "convT_tc: output pitch %d", Lp_out);
  TapGemmArgs a{};
  a.status = status;
  a.CI = CI; a.L_in = L; a.B = B; a.n_chunks = CI / CT_CK;
  a.tap_shift[0] = 0;  a.tap_rows[0] = r * ct;       a.tap_wrow[0] = 0;                        a.tap_dcol[0] = 0;
  a.tap_shift[1] = -1; a.tap_rows[1] = (r / 2) * ct; a.tap_wrow[1] = r * ct;                   a.tap_dcol[1] = r * ct;
  a.tap_shift[2] = 1;  a.tap_rows[2] = (r / 2) * ct; a.tap_wrow[2] = r * ct + (r / 2) * ct;    a.tap_dcol[2] = r * ct + (r / 2) * ct;
  a.rows_total = 2 * r * ct; a.n_cols = 2 * r * ct; a.wblob = wblob; a.r = r; a.co_tile = ct; a.CO = CO;
  a.L_out = r * L; a.Lp_out = Lp_out; a.bias = bias; a.act = 1; a.out = out; a.out_cl = out_cl;
  return launch_tapgemm(x, Lp_in, a, n_tiles, M2TTS_STAGE_VOC_UP, s);
}

// copies rows of L floats to rows of Lp floats (zero tail) — only used by the stand-alone entry points when L % 4 != 0
__global__ void tc_repitch_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int L, int Lp) {
  const long long total = rows * Lp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Lp;
    const int t = (int)(i - row * Lp);
    y[i] = t < L ? x[row * L + t] : 0.f;
  }
}
static int launch_repitch(const float* x, float* y, long long rows, int L, int Lp, cudaStream_t s) {
  const long long total = rows * Lp;
  const int grid = (int)((total + 255) / 256 > kNumSMs * 16 ? kNumSMs * 16 : (total + 255) / 256);
  M2_LAUNCH(M2TTS_STAGE_PACK, tc_repitch_kernel, grid, 256, 0, s, x, y, rows, L, Lp);
  return M2TTS_OK;
}

}  // namespace m2

using namespace m2;

// ---- stand-alone entry points (plain fp32 tensors in and out) ---------------------------------------
extern "C" size_t m2tts_conv_tc_workspace_bytes(int B, int CI, int CO, int L, int r) {
  if (B <= 0 || CI <= 0 || CO <= 0 || L <= 0 || r <= 0) return 0;
  const size_t Lp = (size_t)((L + 3) & ~3);
  const size_t copy = (size_t)B * CI * Lp;                       // re-pitched input when L % 4 != 0
  const size_t wb = (size_t)2 * 2 * (r > 3 ? r : 3) * CO * CI;
  return (copy + wb) * sizeof(float) + 8 * 256;
}

extern "C" int m2tts_conv1d_k3_tc(const float* x, const float* w, const float* bias, const float* residual, float* y,
                                  int B, int CI, int CO, int L, int dilation, int act, void* workspace,
                                  size_t workspace_bytes, m2tts_stream_t stream) {
  M2_REQUIRE(x && w && bias && y && workspace, M2TTS_E_NULLPTR, "conv1d_k3_tc: null pointer");
  M2_REQUIRE(B > 0 && L > 0 && dilation >= 1 && (act == 0 || act == 1), M2TTS_E_BADSHAPE, "conv1d_k3_tc: bad arguments");
  M2_REQUIRE(conv3_tc_eligible(CI, CO), M2TTS_E_UNSUPPORTED,
             "conv1d_k3_tc: needs CI %% 16 == 0 and CO a multiple of 16 that is < 64 or a multiple of 64 (CI=%d CO=%d)", CI, CO);
  const int Lp = (L + 3) & ~3;
  Carver cv(workspace, workspace_bytes);
  float* xcopy = cv.take<float>((size_t)B * CI * Lp);
  float* wblob = cv.take<float>(conv3_tc_wblob_floats(CI, CO));
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "conv1d_k3_tc: workspace too small or misaligned");
  cudaStream_t s = (cudaStream_t)stream;
  const float* xin = x;
  if (Lp != L || (((uintptr_t)x) & 15) != 0) {
    int rc = launch_repitch(x, xcopy, (long long)B * CI, L, Lp, s);
    if (rc) return rc;
    xin = xcopy;
  }
  return launch_conv3_tc(xin, Lp, w, wblob, bias, residual, L, y, L, B, CI, CO, L, dilation, act, M2TTS_STAGE_VOC_RES1, s);
}

extern "C" int m2tts_conv_transpose1d_lrelu_tc(const float* x, const float* w, const float* bias, float* y, int B, int CI,
                                               int CO, int L, int r, void* workspace, size_t workspace_bytes,
                                               m2tts_stream_t stream) {
  M2_REQUIRE(x && w && bias && y && workspace, M2TTS_E_NULLPTR, "conv_transpose1d_tc: null pointer");
  M2_REQUIRE(B > 0 && L > 0, M2TTS_E_BADSHAPE, "conv_transpose1d_tc: bad arguments");
  M2_REQUIRE(convT_tc_eligible(CI, CO, r), M2TTS_E_UNSUPPORTED,
             "conv_transpose1d_tc: needs r == 4 (CO %% 32 == 0) or r == 2 (CO %% 16 == 0), CI %% 16 == 0 (CI=%d CO=%d r=%d)", CI, CO, r);
  M2_REQUIRE(((r * L) & 3) == 0 && (((uintptr_t)y) & 15) == 0, M2TTS_E_UNSUPPORTED,
             "conv_transpose1d_tc: r*L must be a multiple of 4 and y 16-B aligned");
  const int Lp = (L + 3) & ~3;
  Carver cv(workspace, workspace_bytes);
  float* xcopy = cv.take<float>((size_t)B * CI * Lp);
  float* wblob = cv.take<float>(convT_tc_wblob_floats(CI, CO, r));
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "conv_transpose1d_tc: workspace too small or misaligned");
  cudaStream_t s = (cudaStream_t)stream;
  const float* xin = x;
  if (Lp != L || (((uintptr_t)x) & 15) != 0) {
    int rc = launch_repitch(x, xcopy, (long long)B * CI, L, Lp, s);
    if (rc) return rc;
    xin = xcopy;
  }
  return launch_convT_tc(xin, Lp, w, wblob, bias, y, r * L, B, CI, CO, L, r, s);
}
