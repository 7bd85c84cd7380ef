// mma_bench.cu — micro-benchmark of tcgen05.mma kind::tf32 issue/execution cost per operand configuration
// (bring-up tool behind tools/mma_bench.py; the numbers size the tiles of the conv / attention kernels).
// One CTA, one issuing thread, `n` back-to-back M=128 x N x K=8 UMMAs accumulating into `nacc` TMEM accumulators in
// turn; cycles from the first issue to the completion of the last (tcgen05.commit -> mbarrier), by clock64.
//   mode 0: A K-major 128B-swizzle (smem)      B K-major no-swizzle      (channel-last convolutions, C = 32)
//   mode 1: A MN-major 128B/32B-atom (smem)    B MN-major 128B/32B-atom  (attention Q K^T, tap-GEMM convolutions)
//   mode 2: A from TMEM                         B K-major 128B-swizzle    (attention P V)
//   mode 3: A K-major 64B-swizzle (smem)       B K-major no-swizzle      (channel-last convolutions, C = 16)
//   mode 4: A K-major 128B-swizzle (smem)      B K-major 128B-swizzle    (linear layers)
//   mode 5: as mode 4 with kind::f16 (K = 16)                              (16-bit split kernels)
//   mode 6: as mode 5 with the A start address moved by one 128-byte row  (row-shifted conv taps)
#include "../common.cuh"

namespace m2 {

__device__ __forceinline__ uint32_t mb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mb_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) | ((uint64_t)lt << 61);
}

__global__ void __launch_bounds__(128) mma_bench_kernel(int mode, int N, int n, int nacc, int elect, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (mb_smem_u32(smem_raw) + 1023u) & ~1023u;
  float* gen = reinterpret_cast<float*>(smem_raw + (base - mb_smem_u32(smem_raw)));
  const uint32_t sA = base, sB = base + 64 * 1024, sBar = base + 128 * 1024, slot = sBar + 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32 * 1024; i += 128) gen[i] = 0.f;
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
  const uint32_t a_mn = mode == 1 ? 1u : 0u, b_mn = mode == 1 ? 1u : 0u;
  const bool f16 = mode >= 5;
  const uint32_t idesc = f16 ? ((1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24))
                             : ((1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24));
  long long t0 = 0, t1 = 0;
  bool issuer = false;
  if (warp == 0 && elect == 2) {
    // CUTLASS-style: the WHOLE warp runs the issue loop (operands stay warp-uniform), one elected lane issues
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (mode == 0 || mode == 4 || mode == 5) ad[ks] = mb_desc(sA + ks * 32u, 16u, 1024u, 2u);
      else if (mode == 6) ad[ks] = mb_desc(sA + 128u * (1u + (ks & 1)) + ks * 32u, 16u, 1024u, 2u);
      else if (mode == 1) ad[ks] = mb_desc(sA + ks * 1024u, 6144u, 512u, 1u);
      else ad[ks] = mb_desc(sA + (ks & 1) * 32u, 16u, 512u, 4u);
      if (mode == 0 || mode == 3) bd[ks] = mb_desc(sB + (ks >> 1) * (uint32_t)(N * 64) + (ks & 1) * 256u, 128u, 512u, 0u);
      else if (mode == 1) bd[ks] = mb_desc(sB + ks * 1024u, 6144u, 512u, 1u);
      else bd[ks] = mb_desc(sB + ks * 32u, 16u, 1024u, 2u);
    }
    const uint32_t d0 = tmem, d1 = tmem + (nacc > 1 ? (uint32_t)N : 0u);
    t0 = clock64();
    if (mode == 2) {
      for (int i = 0; i < n; i += 4) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"((ks & 1) ? d1 : d0), "r"(tmem + 256u + (uint32_t)ks * 8u), "l"(bd[ks]), "r"(idesc), "r"((uint32_t)(i > 0)) : "memory");
      }
    } else if (f16) {
      for (int i = 0; i < n; i += 4) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"((ks & 1) ? d1 : d0), "l"(ad[ks]), "l"(bd[ks]), "r"(idesc), "r"((uint32_t)(i > 0)) : "memory");
      }
    } else {
    for (int i = 0; i < n; i += 4) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"((ks & 1) ? d1 : d0), "l"(ad[ks]), "l"(bd[ks]), "r"(idesc), "r"((uint32_t)(i > 0)) : "memory");
    }
    }
    t1 = clock64();
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                 ::"r"(sBar) : "memory");
    for (uint32_t it = 0;; ++it) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(sBar), "r"(0u) : "memory");
      if (ok) break;
      if (it > (1u << 24)) __trap();
    }
    const long long t2 = clock64();
    if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp == 0) {
    if (elect) {
      uint32_t is_leader;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_leader));
      issuer = is_leader != 0;
    } else {
      issuer = tid == 0;
    }
  }
  if (issuer) {
    // everything loop-invariant is precomputed: the loop body is 4 x (UMMA) with constant operands
    uint64_t ad[4], bd[4];
    uint32_t at[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      at[ks] = tmem + 256u + (uint32_t)ks * 8u;
      if (mode == 0 || mode == 4) ad[ks] = mb_desc(sA + ks * 32u, 16u, 1024u, 2u);
      else if (mode == 1) ad[ks] = mb_desc(sA + ks * 1024u, 6144u, 512u, 1u);
      else if (mode == 3) ad[ks] = mb_desc(sA + (ks & 1) * 32u, 16u, 512u, 4u);
      else ad[ks] = 0;
      if (mode == 0 || mode == 3) bd[ks] = mb_desc(sB + (ks >> 1) * (uint32_t)(N * 64) + (ks & 1) * 256u, 128u, 512u, 0u);
      else if (mode == 1) bd[ks] = mb_desc(sB + ks * 1024u, 6144u, 512u, 1u);
      else bd[ks] = mb_desc(sB + ks * 32u, 16u, 1024u, 2u);
    }
    const uint32_t d0 = tmem, d1 = tmem + (nacc > 1 ? (uint32_t)N : 0u);
    t0 = clock64();
    if (mode == 2) {
      for (int i = 0; i < n; i += 4) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"((ks & 1) ? d1 : d0), "r"(at[ks]), "l"(bd[ks]), "r"(idesc), "r"((uint32_t)(i > 0)) : "memory");
      }
    } else {
      for (int i = 0; i < n; i += 4) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"((ks & 1) ? d1 : d0), "l"(ad[ks]), "l"(bd[ks]), "r"(idesc), "r"((uint32_t)(i > 0)) : "memory");
      }
    }
    t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sBar) : "memory");
    for (uint32_t it = 0;; ++it) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(sBar), "r"(0u) : "memory");
      if (ok) break;
      if (it > (1u << 24)) __trap();
    }
    const long long t2 = clock64();
    out[0] = t1 - t0;   // issue time
    out[1] = t2 - t0;   // issue + execution
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace m2

using namespace m2;

extern "C" int m2tts_mma_bench(int mode, int N, int n, int nacc, int elect, long long* out_dev, m2tts_stream_t stream) {
  M2_REQUIRE(out_dev && mode >= 0 && mode <= 6 && N >= 16 && N <= 256 && N % 16 == 0 && n > 0 && nacc >= 1 && nacc * N <= 256,
             M2TTS_E_BADSHAPE, "mma_bench: bad arguments");
  const size_t smem = 128 * 1024 + 1024 + 64;
  M2_CUDA_OK(allow_smem(mma_bench_kernel, smem));
  M2_LAUNCH(M2TTS_STAGE_PROBE, mma_bench_kernel, 1, 128, smem, (cudaStream_t)stream, mode, N, n, nacc, elect, out_dev);
  return M2TTS_OK;
}
