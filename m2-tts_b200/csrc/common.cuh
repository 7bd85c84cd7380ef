// common.cuh — shared host/device helpers for the m2tts_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/m2tts_b200.h"

#ifndef __CUDA_ARCH__
#define M2_HOST_ONLY 1
#endif

namespace m2 {

// ---- error reporting (thread local message, never throws) -------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define M2_CUDA_OK(call)                                                      \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) return m2::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define M2_REQUIRE(cond, code, ...)   \
  do {                                \
    if (!(cond)) {                    \
      m2::set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)

// ---- launch accounting + optional per-stage CUDA-event timers ---------------
void note_launch(int stage, cudaStream_t s, bool begin);

struct StageScope {
  int stage; cudaStream_t s;
  StageScope(int st, cudaStream_t stream) : stage(st), s(stream) { note_launch(stage, s, true); }
  ~StageScope() { note_launch(stage, s, false); }
};

// Launch `kernel` and return M2TTS_E_CUDA from the enclosing function on a
// launch-configuration error.
#define M2_LAUNCH(stage, kernel, grid, block, smem, stream, ...)              \
  do {                                                                        \
    {                                                                         \
      m2::StageScope _scope((stage), (stream));                               \
      kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);             \
    }                                                                         \
    cudaError_t _le = cudaGetLastError();                                     \
    if (_le != cudaSuccess) return m2::cuda_fail(_le, #kernel, __FILE__, __LINE__); \
  } while (0)

// The same with programmatic dependent launch (PDL): the kernel may be scheduled while its predecessor in the stream is still
// draining — its CTAs take the SMs the predecessor's CTAs leave — and runs its prologue (barrier init, TMEM allocation,
// tensor-map prefetch, weight loads) there. ONLY for kernels that call pdl_wait() in every CTA before their first access to
// anything another kernel of the stream reads or writes, and pdl_launch_dependents() early.
#define M2_LAUNCH_PDL(stage, kernel, grid, block, smem, strm_, ...)            \
  do {                                                                        \
    {                                                                         \
      m2::StageScope _scope((stage), (strm_));                                \
      cudaLaunchConfig_t _cfg = {};                                           \
      _cfg.gridDim = dim3(grid); _cfg.blockDim = dim3(block);                 \
      _cfg.dynamicSmemBytes = (smem); _cfg.stream = (strm_);                 \
      cudaLaunchAttribute _at[1];                                             \
      _at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;         \
      _at[0].val.programmaticStreamSerializationAllowed = m2::pdl_enabled();  \
      _cfg.attrs = _at; _cfg.numAttrs = 1;                                    \
      cudaError_t _pe = cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__);       \
      if (_pe != cudaSuccess) { (void)cudaGetLastError(); return m2::cuda_fail(_pe, #kernel, __FILE__, __LINE__); } \
    }                                                                         \
    cudaError_t _le = cudaGetLastError();                                     \
    if (_le != cudaSuccess) return m2::cuda_fail(_le, #kernel, __FILE__, __LINE__); \
  } while (0)
int pdl_enabled();      // runtime.cu: 1 unless M2TTS_PDL=0 (tools build: A/B measurements)

// opt a kernel into > 48 KB of dynamic shared memory (once per instantiation)
template <typename K>
inline cudaError_t allow_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// bump allocator over the caller's workspace
struct Carver {
  char* base; size_t size; size_t off;
  Carver(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <typename T> T* take(size_t count) {
    off = align_up(off, 256);
    T* r = (T*)(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= size && (((uintptr_t)base) % 256 == 0); }
};

// SM count of the CURRENT device (148 on B200), cached per device index
int num_sms();
#define kNumSMs (m2::num_sms())

// `precision` argument of the entry points -> M2TTS_PREC_SPLIT16 / _FFMA / _TF32 (M2TTS_PREC_DEFAULT reads M2TTS_PRECISION once)
int resolve_precision(int precision);

// Bring-up switches (environment variables, prof buffers) exist only in the tools build (-DM2TTS_TOOLS); the product
// library never reads them.
#ifdef M2TTS_TOOLS
int tools_env_int(const char* name, int dflt);
#else
inline int tools_env_int(const char*, int dflt) { return dflt; }
#endif

// ---- device helpers ---------------------------------------------------------
#ifdef __CUDACC__
// programmatic dependent launch (see M2_LAUNCH_PDL)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// fp16-range guard of the 16-bit split. A producer converts without clamping (overflow -> inf, NaN -> NaN) and records
// the violation: one FSETP per value instead of the two FMNMX of a clamp. The flagged call's output is invalid and the
// caller re-runs it with M2TTS_PREC_TF32 (include/m2tts_b200.h, "Status word").
__device__ __forceinline__ void h_chk(float v, bool& bad) { bad |= !(fabsf(v) <= 65504.f); }
__device__ __forceinline__ void h_flag(bool bad, int32_t* status) {
  if (bad && status != nullptr) atomicOr(status, (int32_t)M2TTS_ST_FP16_RANGE);
}
// two floats -> packed fp16 hi pair and packed fp16 lo pair, x = hi + lo (22 significant bits)
__device__ __forceinline__ void h_split2(float a0, float a1, uint32_t& hi, uint32_t& lo, bool& bad) {
  h_chk(a0, bad);
  h_chk(a1, bad);
  const __half2 h = __floats2half2_rn(a0, a1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// ---- packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2: one instruction for two values) -------------------------------------
// The epilogues of the 16-bit split kernels turn every accumulator element into an fp16 hi/lo pair; per step that is 2.4 G
// elements in the vocoder alone, and at ~10 instructions per element the kernels were bound by the issue slots of their
// epilogue warps (tools/fused_h_prof.py). These helpers bring an element to ~5 instructions.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ uint64_t f2_pack_u(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// running maximum of |x| that keeps a NaN once it has seen one: the fp16-range check of a whole epilogue is one comparison
// `!(amax <= 65504)` at the end instead of one FSETP per element
__device__ __forceinline__ float amax_nan3(float m, float a, float b) {
  float r;
  asm("{\n\t.reg .f32 x, y;\n\tabs.f32 x, %2;\n\tabs.f32 y, %3;\n\tmax.NaN.f32 %0, %1, x, y;\n\t}" : "=f"(r) : "f"(m), "f"(a), "f"(b));
  return r;
}
// LeakyReLU(0.1) of a pair: max(x, 0.1 x) (slope < 1)
__device__ __forceinline__ uint64_t f2_lrelu01(uint64_t x) {
  float a, b, c, d;
  f2_unpack(x, a, b);
  f2_unpack(f2_mul(x, f2_pack(0.1f, 0.1f)), c, d);
  return f2_pack(fmaxf(a, c), fmaxf(b, d));
}
// pair -> packed fp16 hi and packed fp16 lo with x = hi + lo (hi = fp16(x), lo = fp16(x - hi): 22 significant bits), 6
// instructions per pair (F2FP, 2 x HADD2.F32, FADD2, F2FP, FMNMX3) against 9 of h_split2 with its per-value range checks.
// (A Veltkamp split in packed arithmetic — c = 8193 x, hi = c - (c - x) — does not survive ptxas: it contracts mul.rn.f32x2 +
// sub.rn.f32x2 into FFMA2, which computes c - x exactly and returns hi = x; tools/ubench/split_test.cu.)
__device__ __forceinline__ void h_split_pair(uint64_t x, uint32_t& hi, uint32_t& lo, float& amax) {
  float x0, x1, r0, r1;
  f2_unpack(x, x0, x1);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  f2_unpack(f2_sub(x, f2_pack(hf.x, hf.y)), r0, r1);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
  amax = amax_nan3(amax, x0, x1);
}
__device__ __forceinline__ bool h_amax_bad(float amax) { return !(amax <= 65504.f); }
// packed fp16 hi + packed fp16 lo -> fp32 pair
__device__ __forceinline__ uint64_t h_join_pair(uint32_t hi, uint32_t lo) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lo));
  return f2_add(f2_pack(a.x, a.y), f2_pack(b.x, b.y));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
#endif

}  // namespace m2

// ---- internal launchers (one per .cu file), all return M2TTS_* codes --------
namespace m2 {

struct RowGemmArgs {
  const float* x; int ldx;            // [R,K]
  const float* ln_w; const float* ln_b; float eps;  // LN prologue if ln_w != null
  const float* wt;                    // packed W^T [K,N]
  const float* bias;                  // [N] or null
  const float* residual; int ldr;     // [R,N] or null
  int relu;
  float* y; int ldy;                  // ROWMAJOR output
  // QKV-split output (if qkv_mode): q,k as [B,nh,hd,Lp], v as [B,nh,L,hd]
  int qkv_mode; float* q; float* k; float* v; int L; int Lp; int nh; int hd;
  // qkv_mode 2: q points at six hi/lo planes [6][B,nh,hd,Lp] (plane_stride floats apart), Q scaled by qscale
  float qscale; long long plane_stride;
  int R; int K; int N;
  int stage;
};
int launch_rowgemm(const RowGemmArgs& a, cudaStream_t s);
// transposes W [N,K] -> wt [K,N] for up to 4 matrices in one launch
struct PackJob { const float* src; float* dst; int N; int K; };
int launch_pack_transpose(const PackJob* jobs, int n_jobs, cudaStream_t s);

int launch_attention(const float* q, const float* k, const float* v, float* ctx,
                     const int64_t* lengths, int B, int L, int Lp, int nh, int hd,
                     cudaStream_t s);
bool attention_tc_supported(int hd);
int launch_attention_tc(const float* qkv6, float* ctx, const int64_t* lengths, int B, int L, int Lp,
                        int nh, int hd, cudaStream_t s, float* dbg_s = nullptr, float* dbg_o = nullptr,
                        float* ctx_lo = nullptr);  // ctx_lo != null: write ctx as TF32 hi/lo planes (ctx = hi plane)
// 16-bit split attention (attention_h.cu): qkvh = six fp16 planes [6][B][nh][hd][Lp], Lp % 8 == 0
int launch_attention_h(const void* qkvh, float* ctx, const int64_t* lengths, int B, int L, int Lp, int nh, int hd,
                       cudaStream_t s, float* ctx_lo = nullptr, void* ctx_half_planes = nullptr,   // ctx_half_planes: fp16 [2][B*L][nh*hd]
                       int32_t* status = nullptr);
// ---- persistent 16-bit split linear layers (lin_h.cu): operands as fp16 hi/lo planes ----
struct LinHParams {
  int R, K, N;
  const float* bias; int relu;
  const float* residual; int ldr;  // fp32 [R][ldr] or null (mode 0)
  float* y; int ldy;               // mode 0: fp32 [R][ldy]
  void* y_planes;                  // mode 1: fp16 hi/lo planes [2][R][N]
  void* qkvh; long long plane_stride; int L, nh, hd, Lp; float qscale;   // mode 3: attention operand planes
  int mode;
  int32_t* status;                 // M2TTS_ST_FP16_RANGE when an fp16-plane output leaves the fp16 range (may be null)
  // A = split(LayerNorm(ln_x)) computed by the kernel's own producer warps from fp32 rows [R][K] (a_planes is ignored then):
  // no ln_split_h launch, no round trip of the normalised planes through HBM
  const float* ln_x; const float* ln_w; const float* ln_b; float ln_eps;
};
bool linear_h_eligible(int K, int N);
int launch_ln_split_h(const float* x, const float* w, const float* b, void* planes, long long R, int K, float eps, int32_t* status, cudaStream_t s);
int launch_w_split_h(const float* const* src, void* const* dst, const long long* n, int jobs, int32_t* status, cudaStream_t s);
int launch_linear_h(const void* a_planes, const void* w_planes, const LinHParams& q, int stage, cudaStream_t s);
// ---- tensor-core linear layers (rowgemm_tc.cu) ----
struct LinTcArgs {
  int R, L, K, N, n_tile;          // rows total, rows per utterance, inner dim, outputs, outputs per CTA
  int kboxes;                      // 32-column boxes resident per pass (set by the launcher)
  const float* bias;               // [N] or null
  int relu;
  const float* residual; int ldr;  // plain fp32 [R, ldr] or null
  float* y; int ldy;               // plain output [R, ldy] (mode 0)
  float* y_planes;                 // mode 1: hi/lo planes [2][R][N]
  // mode 2: attention operand planes [6][B][nh][hd][Lp] (fp32 holding TF32 hi/lo); mode 3: the same as fp16 hi/lo
  float* qkv6; long long plane_stride; int nh, hd, Lp; float qscale;
  int mode;
  int32_t* status;                 // mode 3 (fp16 planes): M2TTS_ST_FP16_RANGE
};

bool linear_tc_eligible(int K, int N);
int launch_ln_split(const float* x, const float* w, const float* b, float* planes, long long R, int K, float eps, cudaStream_t s);
int launch_w_split(const float* const* src, float* const* dst, const long long* n, int jobs, cudaStream_t s);
int launch_linear_tc(const float* a_planes, const float* w_planes, LinTcArgs a, int B, int stage, cudaStream_t s);   // a.status: see LinTcArgs
int* debug_words_device();  // pinned mapped scratch for hang diagnostics (may be null)

// tensor-core ("tap-GEMM") convolutions on plain fp32 [B][C][Lp] (conv_tc.cu / conv_tc2.cu)
bool conv3_tc_eligible(int CI, int CO);
bool convT_tc_eligible(int CI, int CO, int r);
size_t conv3_tc_wblob_floats(int CI, int CO);
size_t convT_tc_wblob_floats(int CI, int CO, int r);
// Every conv launcher below takes the module weight `w` AND its image buffer `wblob`: with w != nullptr the image is (re)written
// by a small pack kernel first; w == nullptr means wblob already holds the image (m2tts_vocoder_pack). x == nullptr: pack only.
int launch_conv3_tc(const float* x, int Lp_in, const float* w, float* wblob, const float* bias, const float* residual,
                    int Lp_res, float* out, int Lp_out, int B, int CI, int CO, int L, int dil, int act, int stage,
                    cudaStream_t s, int out_cl = 0, int32_t* status = nullptr);   // out_cl 1: write channel-last fp32 [B][L][CO]; 2: channel-last fp16 hi/lo planes [2][B][L][CO]
int launch_convT_tc(const float* x, int Lp_in, const float* w, float* wblob, const float* bias, float* out, int Lp_out,
                    int B, int CI, int CO, int L, int r, cudaStream_t s, int out_cl = 0, int32_t* status = nullptr);   // out_cl 2: fp16 hi/lo planes, channel-last
// whole ResBlock for C = 64 as one kernel, channel-last fp16 hi/lo planes in (voc_res_h.cu)
bool voc_res_h_eligible(int C, int dil);
size_t voc_res_h_wblob_bytes(int C);
int launch_voc_res_h(const void* uh, long long u_plane, const float* w1, const float* b1, const float* w2, const float* b2, void* wblob,
                     void* out_h, long long out_plane, float* out_f, int B, int C, int L, int stage, int32_t* status, cudaStream_t s);
// one Conv1d(C, C, 3) for C = 128 on channel-last fp16 hi/lo planes (voc_conv_h.cu); output planes or fp32 channel-first
bool voc_conv_h_eligible(int C, int dil);
bool voc_conv_h_io_eligible(int CI, int CO);      // 64 < CI <= 128 (zero-padded to 128), CO a multiple of 64: also the input conv
size_t voc_conv_h_wblob_bytes(int CO);
int launch_voc_conv_h(const void* xh, long long x_plane, const float* w, const float* bias, void* wblob, const void* res_h,
                      long long res_plane, void* out_h, long long out_plane, float* out_cf, int Lp_out, int B, int CI, int CO, int L, int act,
                      int stage, int32_t* status, cudaStream_t s);
// ConvTranspose1d(CI, CI/2, 8, stride 4, padding 2) + leaky_relu on channel-last fp16 hi/lo planes (voc_up_h.cu)
bool voc_up_h_eligible(int CI, int CO, int r);
size_t voc_up_h_wblob_bytes(int CI);
int launch_voc_up_h(const void* xh, long long x_plane, const float* w, const float* bias, void* wblob, void* out_h, long long out_plane,
                    int B, int CI, int L, int stage, int32_t* status, cudaStream_t s);
// 16-bit split flavour (voc_fused_h.cu): input as fp16 hi/lo planes [2][B][L][2C]
size_t voc_fused_h_wblob_bytes(int C);
int launch_voc_stage_fused_h(const void* xh, long long x_plane, const float* up_w, const float* up_b, const float* w1, const float* b1,
                             const float* w2, const float* b2, const float* out_w, const float* out_b, void* wblob,
                             void* out_h, long long out_plane, float* out_f, int B, int C, int L_in, int stage, int32_t* status, cudaStream_t s);
int launch_split_planes_h(const float* x, void* planes, long long n, int32_t* status, cudaStream_t s);
// fused narrow stage on channel-last activations (voc_fused.cu): upsample x2 + ResBlock (+ output conv + tanh)
bool voc_fused_eligible(int C, int r, int dil);
bool voc_fused_h_eligible(int C, int r, int dil, bool final_stage);      // voc_fused_h.cu: C = 8 only as the last stage (zero-padded to 16)
size_t voc_fused_wblob_floats(int C);
int launch_voc_stage_fused(const float* x, const float* up_w, const float* up_b, const float* w1, const float* b1,
                           const float* w2, const float* b2, const float* out_w, const float* out_b, float* wblob,
                           float* out, int B, int C, int L_in, int stage, cudaStream_t s);
}  // namespace m2
