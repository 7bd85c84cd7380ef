// umma_probe_f16.cu — bring-up probe for tcgen05.mma kind::f16 operand layouts (fp16 inputs, fp32 accumulate), the
// building block of the 16-bit split-precision kernels. One CTA computes D[128,N] = A[128,K] B[N,K]^T, K-step 16.
//   mode 0: A MN-major 128B-swizzle in shared memory, B MN-major 128B-swizzle   (attention Q K^T)
//   mode 1: A from TMEM (two fp16 per 32-bit column, low half = even k), B K-major 128B-swizzle   (attention P V)
//   mode 2: A K-major 128B-swizzle in shared memory, B K-major 128B-swizzle
#include "../common.cuh"
#include <cuda_fp16.h>

namespace m2 {

__device__ __forceinline__ uint32_t ph_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t ph_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) | ((uint64_t)lt << 61);
}
// MN-major, 128B swizzle, 16-bit: 64 mn contiguous (128 B), 8 k rows per 1024-B atom; mn groups at `lbo`, k groups at 1024 B
__device__ __forceinline__ uint32_t ph_off_mn(int mn, int k, uint32_t lbo) {
  const uint32_t in = (uint32_t)(k & 7) * 128u + (uint32_t)(mn & 63) * 2u;
  return (uint32_t)(mn >> 6) * lbo + (uint32_t)(k >> 3) * 1024u + (in ^ (((in >> 7) & 7u) << 4));
}
// K-major, 128B swizzle, 16-bit: 64 k contiguous (128 B) per row, 8 rows per 1024-B atom
__device__ __forceinline__ uint32_t ph_off_k(int mn, int k, uint32_t kblock_bytes) {
  const uint32_t in = (uint32_t)(mn & 7) * 128u + (uint32_t)(k & 63) * 2u;
  return (uint32_t)(k >> 6) * kblock_bytes + (uint32_t)(mn >> 3) * 1024u + (in ^ (((in >> 7) & 7u) << 4));
}

__global__ void __launch_bounds__(128) umma_probe_f16_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                             float* __restrict__ D, int N, int K, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ph_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ph_smem_u32(smem_raw));
  const uint32_t a_bytes = 32 * 1024, b_bytes = 32 * 1024;
  const uint32_t sA = base, sB = base + a_bytes, sBar = sB + b_bytes, slot = sBar + 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < (a_bytes + b_bytes) / 4; i += 128) reinterpret_cast<float*>(gen)[i] = 0.f;
  __syncthreads();
  const uint32_t lbo_a = (uint32_t)(K / 8) * 1024u, lbo_b = (uint32_t)(K / 8) * 1024u;
  if (mode == 0) {
    for (int i = tid; i < 128 * K; i += 128) *reinterpret_cast<__half*>(gen + ph_off_mn(i / K, i % K, lbo_a)) = __float2half(A[i]);
    for (int i = tid; i < N * K; i += 128) *reinterpret_cast<__half*>(gen + a_bytes + ph_off_mn(i / K, i % K, lbo_b)) = __float2half(Bm[i]);
  } else {
    if (mode == 2)
      for (int i = tid; i < 128 * K; i += 128) *reinterpret_cast<__half*>(gen + ph_off_k(i / K, i % K, 16 * 1024)) = __float2half(A[i]);
    for (int i = tid; i < N * K; i += 128) *reinterpret_cast<__half*>(gen + a_bytes + ph_off_k(i / K, i % K, (uint32_t)(N / 8) * 1024u)) = __float2half(Bm[i]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sBar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t acol = 256;     // A operand columns in TMEM (mode 1)
  if (mode == 1) {
    // thread = row m: column c holds (A[m, 2c], A[m, 2c+1]) with the even k in the low half
    for (int c = 0; c < K / 2; c += 8) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) {
        const __half2 h2 = __floats2half2_rn(A[tid * K + 2 * (c + j)], A[tid * K + 2 * (c + j) + 1]);
        r[j] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   ::"r"(lane_base + acol + c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) {
    const uint32_t mn = mode == 0 ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (mn << 15) | (mn << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t ad = 0, bd;
      if (mode == 0) {
        ad = ph_desc(sA + ks * 2048u, lbo_a, 1024u, 2u);
        bd = ph_desc(sB + ks * 2048u, lbo_b, 1024u, 2u);
      } else {
        if (mode == 2) ad = ph_desc(sA + (ks >> 2) * 16384u + (ks & 3) * 32u, 16u, 1024u, 2u);
        bd = ph_desc(sB + (ks >> 2) * (uint32_t)(N / 8) * 1024u + (ks & 3) * 32u, 16u, 1024u, 2u);
      }
      const uint32_t acc = ks ? 1u : 0u;
      if (mode == 1)
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(tmem), "r"(tmem + acol + (uint32_t)ks * 8u), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                 ::"r"(sBar) : "memory");
  }
  __syncwarp();
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(sBar), "r"(0u) : "memory");
    if (ok) break;
    if (it > (1u << 22)) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < N; c += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(lane_base + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[tid * N + c + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace m2

using namespace m2;

// A [128,K], B [N,K], D [128,N] device fp32 (A/B values must be fp16-representable for an exact check).
extern "C" int m2tts_umma_probe_f16(const float* A, const float* Bm, float* D, int N, int K, int mode, m2tts_stream_t stream) {
  M2_REQUIRE(A && Bm && D, M2TTS_E_NULLPTR, "umma_probe_f16: null pointer");
  M2_REQUIRE(N >= 16 && N <= 128 && N % 16 == 0 && K >= 16 && K <= 64 && K % 16 == 0 && mode >= 0 && mode <= 2, M2TTS_E_BADSHAPE,
             "umma_probe_f16: N=%d K=%d mode=%d", N, K, mode);
  const size_t smem = 64 * 1024 + 1024 + 64;
  M2_CUDA_OK(allow_smem(umma_probe_f16_kernel, smem));
  M2_LAUNCH(M2TTS_STAGE_PROBE, umma_probe_f16_kernel, 1, 128, smem, (cudaStream_t)stream, A, Bm, D, N, K, mode);
  return M2TTS_OK;
}
