// umma_probe.cu — bring-up probe for tcgen05.mma operand layouts (test infrastructure of the
// kernels, exported for tests/test_gpu_umma_probe.py). One CTA computes D[128,N] = A[128,K] B[N,K]^T
// with kind::tf32, writing the logical fp32 matrices into shared memory itself (no TMA) in one of
// the canonical UMMA layouts, so descriptor semantics can be checked against a host matmul.
//   mode 0: K-major,  128-byte swizzle   mode 1: MN-major, 128-byte swizzle
//   mode 2: K-major,  no swizzle         mode 3: MN-major, no swizzle
#include "../common.cuh"

namespace m2 {

__device__ __forceinline__ uint32_t pr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct ProbeOperand {
  int mode;            // see above
  uint32_t lbo, sbo;   // byte offsets used to lay the data out
  uint32_t dlbo, dsbo; // byte offsets written into the descriptor (normally the same)
  uint32_t kstep_bytes;  // start-address advance per K=8 step
  int mn_major_flag;   // bit for the instruction descriptor
};

// byte offset of logical element (mn, k) inside the operand's smem region
__device__ __forceinline__ uint32_t probe_offset(const ProbeOperand& o, int mn, int k) {
  const int ks = k >> 3, kr = k & 7;
  uint32_t off;
  switch (o.mode) {
    case 0: {  // K-major SW128: rows = mn at 128-B pitch, 8-row groups at SBO; 32 k per 128-B row
      const uint32_t in = (uint32_t)(mn & 7) * 128u + (uint32_t)(k & 31) * 4u;
      off = (uint32_t)(mn >> 3) * o.sbo + (uint32_t)(k >> 5) * o.lbo + (in ^ (((in >> 7) & 7u) << 4));
      return off;
    }
    case 1: {  // MN-major SW128: 32 mn contiguous (128 B), k rows at 128-B pitch, mn atoms at LBO, k groups at SBO
      const uint32_t in = (uint32_t)kr * 128u + (uint32_t)(mn & 31) * 4u;
      off = (uint32_t)(mn >> 5) * o.lbo + (uint32_t)ks * o.sbo + (in ^ (((in >> 7) & 7u) << 4));
      return off;
    }
    case 4: {  // MN-major, 128-B swizzle with 32-B atomicity (the only MN-major layout for 32-bit operands):
               // 32 mn contiguous, 4 k rows per 512-B atom, 32-B chunks XORed with (k row & 3); mn atoms at LBO,
               // 4-row k groups at SBO
      const uint32_t in = (uint32_t)(k & 3) * 128u + (uint32_t)(mn & 31) * 4u;
      return (uint32_t)(mn >> 5) * o.lbo + (uint32_t)(k >> 2) * o.sbo + (in ^ (((in >> 7) & 3u) << 5));
    }
    case 2:    // K-major no swizzle: core matrix = 8 mn rows x 16 B (4 k); mn groups at SBO, k chunks at LBO
      return (uint32_t)(mn >> 3) * o.sbo + (uint32_t)(k >> 2) * o.lbo + (uint32_t)(mn & 7) * 16u + (uint32_t)(k & 3) * 4u;
    default:   // MN-major no swizzle: core matrix = 8 k rows x 16 B (4 mn); mn chunks at SBO, k groups at LBO
      return (uint32_t)(mn >> 2) * o.sbo + (uint32_t)ks * o.lbo + (uint32_t)kr * 16u + (uint32_t)(mn & 3) * 4u;
  }
}

__global__ void __launch_bounds__(128) umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                         float* __restrict__ D, int N, int K, ProbeOperand pa,
                                                         ProbeOperand pb, uint32_t a_bytes, uint32_t b_bytes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (pr_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - pr_smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + a_bytes, sBar = sB + b_bytes, slot = sBar + 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < (a_bytes + b_bytes) / 4; i += 128) reinterpret_cast<float*>(gen)[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < 128 * K; i += 128) {
    const int mn = i / K, k = i % K;
    *reinterpret_cast<float*>(gen + probe_offset(pa, mn, k)) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int mn = i / K, k = i % K;
    *reinterpret_cast<float*>(gen + a_bytes + probe_offset(pb, mn, k)) = Bm[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sBar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)pa.mn_major_flag << 15) |
                           ((uint32_t)pb.mn_major_flag << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint64_t lt_a = (pa.mode == 4) ? 1ull : (pa.mode < 2) ? 2ull : 0ull;
    const uint64_t lt_b = (pb.mode == 4) ? 1ull : (pb.mode < 2) ? 2ull : 0ull;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t aa = sA + ks * pa.kstep_bytes, bb = sB + ks * pb.kstep_bytes;
      const uint64_t ad = (uint64_t)((aa >> 4) & 0x3FFFu) | ((uint64_t)((pa.dlbo >> 4) & 0x3FFFu) << 16) |
                          ((uint64_t)((pa.dsbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (lt_a << 61);
      const uint64_t bd = (uint64_t)((bb >> 4) & 0x3FFFu) | ((uint64_t)((pb.dlbo >> 4) & 0x3FFFu) << 16) |
                          ((uint64_t)((pb.dsbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (lt_b << 61);
      const uint32_t acc = ks ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sBar) : "memory");
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
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
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
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

}  // namespace m2

using namespace m2;

// A [128,K], B [N,K], D [128,N] device fp32. a[]/b[] = {mode, layout_lbo, layout_sbo, kstep_bytes,
// mn_major_flag, region_bytes, desc_lbo, desc_sbo}.
extern "C" int m2tts_umma_probe(const float* A, const float* Bm, float* D, int N, int K, const int* a, const int* b,
                                m2tts_stream_t stream) {
  M2_REQUIRE(A && Bm && D && a && b, M2TTS_E_NULLPTR, "umma_probe: null pointer");
  M2_REQUIRE(N >= 8 && N <= 256 && N % 8 == 0 && K >= 8 && K % 8 == 0, M2TTS_E_BADSHAPE, "umma_probe: N=%d K=%d", N, K);
  ProbeOperand pa{a[0], (uint32_t)a[1], (uint32_t)a[2], (uint32_t)a[6], (uint32_t)a[7], (uint32_t)a[3], a[4]};
  ProbeOperand pb{b[0], (uint32_t)b[1], (uint32_t)b[2], (uint32_t)b[6], (uint32_t)b[7], (uint32_t)b[3], b[4]};
  const size_t smem = (size_t)a[5] + b[5] + 1024 + 64;
  M2_REQUIRE(smem <= 200 * 1024, M2TTS_E_UNSUPPORTED, "umma_probe: operands too large");
  M2_CUDA_OK(allow_smem(umma_probe_kernel, smem));
  M2_LAUNCH(M2TTS_STAGE_PROBE, umma_probe_kernel, 1, 128, smem, (cudaStream_t)stream, A, Bm, D, N, K, pa, pb,
            (uint32_t)a[5], (uint32_t)b[5]);
  return M2TTS_OK;
}
