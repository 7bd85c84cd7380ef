// rowgemm_tc.cu — the transformer's linear layers on tcgen05/TMEM fed by TMA, fp32-faithful via 3xTF32.
//   C[rows, N] = A[rows, K] @ W[N, K]^T   (+bias, ReLU, +residual; or the attention operand planes)
// A arrives as hi/lo planes [2][R][K] (K-major rows; written by the LayerNorm/split kernel below, by the
// attention kernel's epilogue, or by the FFN1 epilogue), W as hi/lo planes [2][N][K] (nn.Linear layout).
// Both are K-major, so every TMA box {32 k x rows} is the canonical 128-byte-swizzle UMMA operand.
// K <= 256 is tiny here, so a CTA loads its whole 128-row x K slab and one N tile of W in one shot,
// issues K/8 x 3 UMMAs (M128, N = tile, K8), and the 128 threads run the epilogue (thread = row).
// Reference ops: components.py:55,70-72 (qkv), :90 (out_proj), :103 (ffn), tts_model.py:223-226.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>

namespace m2 {

constexpr int LG_BM = 128;
constexpr int LG_THREADS = 128;
constexpr int LG_KB = 3;            // 32-column boxes per pass (96 inner-dim columns)

__device__ __forceinline__ uint32_t lg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lg_wait(uint32_t bar, uint32_t parity, int* dbg, int code) {
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  if (dbg != nullptr) { dbg[0] = code; dbg[2] = blockIdx.x; dbg[3] = blockIdx.y; dbg[4] = blockIdx.z; __threadfence_system(); }
  __trap();
}
__device__ __forceinline__ void lg_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t lg_desc_k_sw128(uint32_t saddr) {   // K-major, 128-B swizzle: SBO = 1024 B, LBO unused
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void lg_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ float lg_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

__global__ void __launch_bounds__(LG_THREADS, 1)
lingemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const LinTcArgs a, uint32_t tmem_cols, int* dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (lg_smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_box = LG_BM * 128u, w_box = (uint32_t)a.n_tile * 128u;
  const uint32_t sA = sbase;                                   // [plane][kbox][128 rows x 128 B]
  const uint32_t sW = sA + 2u * a.kboxes * a_box;              // [plane][kbox][n_tile rows x 128 B]
  const uint32_t sBar = sW + 2u * a.kboxes * w_box;
  const uint32_t bar_full = sBar, bar_acc = sBar + 8, bar_empty = sBar + 16, tmem_slot = sBar + 24;

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int l0 = blockIdx.x * LG_BM, ntile = blockIdx.y, b = blockIdx.z;
  const int row0 = b * a.L + l0;
  const int n0 = ntile * a.n_tile;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_full));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_acc));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_empty));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);      // warp-uniform for ptxas (uniform-register UMMA operands)

  if (warp == 0) {
    // K is consumed in passes of up to LG_KB boxes (96 columns) through the same buffers. The whole warp runs the
    // loop (warp-uniform operands) and one elected lane issues each UMMA; lane 0 alone drives the TMA.
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.n_tile >> 3) << 17) | ((uint32_t)(LG_BM >> 4) << 24);
    const int total_boxes = (a.K + 31) / 32;
    const int passes = (total_boxes + a.kboxes - 1) / a.kboxes;
    for (int p = 0; p < passes; ++p) {
      const int kb0 = p * a.kboxes;
      const int nb = min(a.kboxes, total_boxes - kb0);
      if (p > 0) lg_wait(bar_empty, (uint32_t)((p - 1) & 1), dbg, 13);   // previous pass's UMMAs have read the buffers
      if (tid == 0) {
        const uint32_t bytes = 2u * nb * (a_box + w_box);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_full), "r"(bytes) : "memory");
        for (int plane = 0; plane < 2; ++plane)
          for (int kb = 0; kb < nb; ++kb) {
            lg_tma_2d(sA + (uint32_t)(plane * a.kboxes + kb) * a_box, &tmap_a, (kb0 + kb) * 32, plane * a.R + row0, bar_full);
            lg_tma_2d(sW + (uint32_t)(plane * a.kboxes + kb) * w_box, &tmap_w, (kb0 + kb) * 32, plane * a.N + n0, bar_full);
          }
      }
      __syncwarp();
      lg_wait(bar_full, (uint32_t)(p & 1), dbg, 11);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int ksteps = (min(a.K - kb0 * 32, nb * 32) + 7) / 8;
      for (int term = 0; term < 3; ++term) {               // hi*hi, hi*lo, lo*hi
        const uint32_t ap = (term == 2) ? 1u : 0u, wp = (term == 1) ? 1u : 0u;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t ad = lg_desc_k_sw128(sA + (ap * a.kboxes + (ks >> 2)) * a_box + (ks & 3) * 32u);
          const uint64_t bd = lg_desc_k_sw128(sW + (wp * a.kboxes + (ks >> 2)) * w_box + (ks & 3) * 32u);
          const uint32_t acc = (p | term | ks) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
      }
      asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                   "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar_empty) : "memory");
    }
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar_acc) : "memory");
  }
  __syncwarp();
  lg_wait(bar_acc, 0, dbg, 12);
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ===== epilogue: thread = row =====
  const int l = l0 + tid;
  const bool valid = l < a.L;
  const long long row = (long long)row0 + tid;
  const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
  if (a.mode == 3) {
    // Attention operand planes as fp16 hi/lo (attention_h.cu), positions contiguous: the accumulator tile is
    // transposed through shared memory (the operand buffers are dead once the accumulator is complete) so that
    // every global store is 16 bytes of 8 consecutive positions instead of 2-byte scalars.
    __half* stg = reinterpret_cast<__half*>(smem_raw + (sbase - lg_smem_u32(smem_raw)));     // [plane][n_tile][128]
    const int H = a.nh * a.hd, nt = a.n_tile;
    bool bad = false;
    for (int c0 = 0; c0 < nt; c0 += 16) {
      uint32_t v[16];
      lg_ld16(t_lane + c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = n0 + c0 + j;
        float t = __uint_as_float(v[j]);
        if (a.bias != nullptr) t += __ldg(a.bias + n);
        if (n < H) t *= a.qscale;                                  // Q rows carry scale * log2(e)
        t = valid ? t : 0.f;
        h_chk(t, bad);                                             // fp16 range (M2TTS_ST_FP16_RANGE)
        const __half h = __float2half_rn(t);
        stg[(c0 + j) * LG_BM + tid] = h;
        stg[(nt + c0 + j) * LG_BM + tid] = __float2half_rn(t - __half2float(h));
      }
    }
    h_flag(bad, a.status);
    __syncthreads();
    __half* base = reinterpret_cast<__half*>(a.qkv6);
    for (int idx = tid; idx < 2 * nt * 16; idx += LG_THREADS) {
      const int chunk = idx & 15, nl = (idx >> 4) % nt, pl = idx / (16 * nt);
      const int lpos = l0 + chunk * 8;
      if (lpos >= a.Lp) continue;
      const int n = n0 + nl;
      const int which = n / H, rem = n - which * H;
      const int head = rem / a.hd, d = rem - head * a.hd;
      const uint4 val = *reinterpret_cast<const uint4*>(stg + (pl * nt + nl) * LG_BM + chunk * 8);
      *reinterpret_cast<uint4*>(base + (long long)(2 * which + pl) * a.plane_stride + (((long long)b * a.nh + head) * a.hd + d) * a.Lp + lpos) = val;
    }
  } else
  for (int c0 = 0; c0 < a.n_tile; c0 += 16) {
    uint32_t v[16];
    __syncwarp();   // rows past the end skip the stores below; reconverge before the .sync.aligned load
    lg_ld16(t_lane + c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (!valid) continue;
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int n = n0 + c0 + j;
      float t = __uint_as_float(v[j]);
      if (a.bias != nullptr) t += __ldg(a.bias + n);
      if (a.relu) t = fmaxf(t, 0.f);
      x[j] = t;
    }
    if (a.mode == 0) {
      if (a.residual != nullptr) {
        const float4* rp = reinterpret_cast<const float4*>(a.residual + row * a.ldr + n0 + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float4 r = __ldg(rp + j); x[4 * j] += r.x; x[4 * j + 1] += r.y; x[4 * j + 2] += r.z; x[4 * j + 3] += r.w; }
      }
      float4* yp = reinterpret_cast<float4*>(a.y + row * a.ldy + n0 + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) yp[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    } else if (a.mode == 1) {
      float4* hp = reinterpret_cast<float4*>(a.y_planes + row * a.N + n0 + c0);
      float4* lp = reinterpret_cast<float4*>(a.y_planes + ((long long)a.R + row) * a.N + n0 + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float h[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { h[e] = lg_hi(x[4 * j + e]); lo[e] = lg_hi(x[4 * j + e] - h[e]); }
        hp[j] = make_float4(h[0], h[1], h[2], h[3]);
        lp[j] = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
    } else {
      const int H = a.nh * a.hd;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = n0 + c0 + j;
        const int which = n / H, rem = n - which * H;
        const int head = rem / a.hd, d = rem - head * a.hd;
        const float t = (which == 0) ? x[j] * a.qscale : x[j];
        const float h = lg_hi(t);
        float* hp = a.qkv6 + (long long)(2 * which) * a.plane_stride + (((long long)b * a.nh + head) * a.hd + d) * a.Lp + l;
        hp[0] = h;
        hp[a.plane_stride] = lg_hi(t - h);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

// LayerNorm over the last dim (optional) + hi/lo split: x [R,K] -> planes [2][R][K]. One warp per row.
__global__ void ln_split_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bvec,
                                float* __restrict__ planes, long long R, int K, float eps) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* xr = x + row * K;
  float mean = 0.f, rstd = 1.f;
  if (w != nullptr) {
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += xr[k];
    mean = warp_sum(s) / (float)K;
    float q = 0.f;
    for (int k = lane; k < K; k += 32) { const float d = xr[k] - mean; q += d * d; }
    rstd = 1.0f / sqrtf(warp_sum(q) / (float)K + eps);
  }
  float* hp = planes + row * K;
  float* lp = planes + (R + row) * K;
  for (int k = lane; k < K; k += 32) {
    float v = xr[k];
    if (w != nullptr) v = (v - mean) * rstd * __ldg(w + k) + __ldg(bvec + k);
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hp[k] = h;
    lp[k] = __uint_as_float(__float_as_uint(v - h) & 0xFFFFE000u);
  }
}

// W [N,K] fp32 -> planes [2][N][K]; several matrices per launch (blockIdx.y = job)
struct WSplitJobs { const float* src[4]; float* dst[4]; long long n[4]; };
__global__ void w_split_kernel(WSplitJobs jobs) {
  const float* s = jobs.src[blockIdx.y];
  float* d = jobs.dst[blockIdx.y];
  const long long n = jobs.n[blockIdx.y];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = s[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    d[i] = h;
    d[n + i] = __uint_as_float(__float_as_uint(v - h) & 0xFFFFE000u);
  }
}

// ---- host ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 lg_encode_fn() {
  static EncodeTiledFn3 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn3)p;
  }
  return fn;
}

static inline int min_i(int a, int b) { return a < b ? a : b; }
static int pick_n_tile(int N, int K) {
  // largest tile (multiple of 16, <= 256) dividing N whose operands fit beside the 128-row A slab
  const int kboxes = min_i((K + 31) / 32, LG_KB);
  const int cands[] = {256, 192, 160, 144, 128, 96, 80, 64, 48, 32, 16};
  for (int c : cands) {
    if (N % c) continue;
    const size_t smem = 2ull * kboxes * (LG_BM * 128 + c * 128) + 1024 + 64;
    if (smem <= 225 * 1024) return c;
  }
  return 0;
}

bool linear_tc_eligible(int K, int N) { return K % 8 == 0 && K <= 256 && N % 16 == 0 && pick_n_tile(N, K) > 0; }

int launch_ln_split(const float* x, const float* w, const float* b, float* planes, long long R, int K, float eps,
                    cudaStream_t s) {
  const int wpb = 8;
  M2_LAUNCH(M2TTS_STAGE_LAYERNORM, ln_split_kernel, (unsigned)((R + wpb - 1) / wpb), wpb * 32, 0, s, x, w, b, planes, R, K, eps);
  return M2TTS_OK;
}

int launch_w_split(const float* const* src, float* const* dst, const long long* n, int jobs, cudaStream_t s) {
  M2_REQUIRE(jobs >= 1 && jobs <= 4, M2TTS_E_BADSHAPE, "w_split: 1..4 jobs");
  WSplitJobs j{};
  long long mx = 1;
  for (int i = 0; i < jobs; ++i) { j.src[i] = src[i]; j.dst[i] = dst[i]; j.n[i] = n[i]; if (n[i] > mx) mx = n[i]; }
  dim3 grid((unsigned)((mx + 255) / 256 > 256 ? 256 : (mx + 255) / 256), jobs);
  M2_LAUNCH(M2TTS_STAGE_PACK, w_split_kernel, grid, 256, 0, s, j);
  return M2TTS_OK;
}

// a_planes [2][R][K], w_planes [2][N][K]; rows are grouped in utterances of L rows (R = B*L).
int launch_linear_tc(const float* a_planes, const float* w_planes, LinTcArgs a, int B, int stage, cudaStream_t s) {
  M2_REQUIRE(linear_tc_eligible(a.K, a.N), M2TTS_E_UNSUPPORTED, "linear_tc: K=%d N=%d not eligible", a.K, a.N);
  M2_REQUIRE((a.K & 3) == 0 && (((uintptr_t)a_planes) & 15) == 0 && (((uintptr_t)w_planes) & 15) == 0, M2TTS_E_BADSHAPE,
             "linear_tc: misaligned operands");
  EncodeTiledFn3 enc = lg_encode_fn();
  M2_REQUIRE(enc != nullptr, M2TTS_E_CUDA, "linear_tc: cuTensorMapEncodeTiled unavailable");
  a.n_tile = pick_n_tile(a.N, a.K);
  a.kboxes = min_i((a.K + 31) / 32, LG_KB);
  CUtensorMap ta, tw;
  const cuuint32_t estr[2] = {1, 1};
  {
    const cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)2 * a.R};
    const cuuint64_t strides[1] = {(cuuint64_t)a.K * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)LG_BM};
    const CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a_planes, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "linear_tc: tensor map (A) failed (%d)", (int)r);
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)2 * a.N};
    const cuuint64_t strides[1] = {(cuuint64_t)a.K * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)a.n_tile};
    const CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w_planes, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M2_REQUIRE(r == CUDA_SUCCESS, M2TTS_E_CUDA, "linear_tc: tensor map (W) failed (%d)", (int)r);
  }
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < a.n_tile) tmem_cols <<= 1;
  const size_t smem = 2ull * a.kboxes * (LG_BM * 128 + a.n_tile * 128) + 1024 + 64;
  M2_CUDA_OK(allow_smem(lingemm_tc_kernel, smem));
  dim3 grid(ceil_div(a.L, LG_BM), a.N / a.n_tile, B);
  M2_LAUNCH(stage, lingemm_tc_kernel, grid, LG_THREADS, smem, s, ta, tw, a, tmem_cols, debug_words_device());
  return M2TTS_OK;
}

}  // namespace m2
