// rowshift_probe.cu — bring-up probe: K-major swizzled A operand whose descriptor start address is moved by
// whole rows inside the swizzle pattern (how the channel-last convolutions apply their tap shifts).
// A_full [R_TOTAL rows][K] is written to shared memory at `rowbytes` pitch (128 B -> SWIZZLE_128B, 64 B ->
// SWIZZLE_64B, chunk index XORed with address bits [7..]); the MMA reads 128 rows starting at row `shift`:
//   D[m, n] = sum_k A_full[shift + m, k] * B[n, k],   B in the K-major no-swizzle core-matrix layout.
// `base_offset` is written into descriptor bits 49-51 so both conventions can be checked.
#include "../common.cuh"

namespace m2 {

__device__ __forceinline__ uint32_t rp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) rowshift_probe_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                             float* __restrict__ D, int rows_total, int N, int K,
                                                             int rowbytes, int shift, int base_offset) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (rp_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - rp_smem_u32(smem_raw));
  const uint32_t a_bytes = (uint32_t)rows_total * rowbytes;
  const uint32_t b_bytes = (uint32_t)N * K * 4u;
  const int kblocks = (K * 4 + rowbytes - 1) / rowbytes;       // K wider than one row -> consecutive [rows][rowbytes] blocks
  const uint32_t sA = base, sB = base + (((uint32_t)kblocks * a_bytes + 1023u) & ~1023u), sBar = sB + b_bytes, slot = sBar + 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int kper = rowbytes / 4;
  const uint32_t mask = rowbytes == 128 ? 7u : 3u;
  for (int i = tid; i < rows_total * K; i += 128) {
    const int r = i / K, k = i % K;
    const int kb = k / kper, kk = k % kper;
    const uint32_t addr = (uint32_t)kb * a_bytes + (uint32_t)r * rowbytes;          // row start (pattern is address based)
    const uint32_t chunk = ((uint32_t)kk >> 2) ^ ((addr >> 7) & mask);
    *reinterpret_cast<float*>(gen + addr + chunk * 16u + (kk & 3) * 4u) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const uint32_t off = (uint32_t)(k >> 4) * (uint32_t)(N * 64) + (uint32_t)(n >> 3) * 512u + (uint32_t)((k & 15) >> 2) * 128u +
                         (uint32_t)(n & 7) * 16u + (uint32_t)(k & 3) * 4u;
    *reinterpret_cast<float*>(gen + (sB - base) + off) = Bm[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint64_t lt_a = rowbytes == 128 ? 2ull : 4ull;           // SWIZZLE_128B / SWIZZLE_64B
    const uint32_t sbo_a = 8u * (uint32_t)rowbytes;
    for (int ks = 0; ks < K / 8; ++ks) {
      const int kb = (ks * 8) / kper, kin = (ks * 8) % kper;
      const uint32_t aa = sA + (uint32_t)kb * a_bytes + (uint32_t)shift * rowbytes + (uint32_t)kin * 4u;
      const uint32_t bb = sB + (uint32_t)(ks >> 1) * (uint32_t)(N * 64) + (uint32_t)(ks & 1) * 256u;
      const uint64_t ad = (uint64_t)((aa >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)((sbo_a >> 4) & 0x3FFFu) << 32) |
                          (1ull << 46) | ((uint64_t)(base_offset & 7) << 49) | (lt_a << 61);
      const uint64_t bd = (uint64_t)((bb >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46);
      const uint32_t acc = ks ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
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

// A [rows_total,K], B [N,K], D [128,N] device fp32; rows_total >= shift + 128.
extern "C" int m2tts_rowshift_probe(const float* A, const float* Bm, float* D, int rows_total, int N, int K, int rowbytes,
                                    int shift, int base_offset, m2tts_stream_t stream) {
  M2_REQUIRE(A && Bm && D, M2TTS_E_NULLPTR, "rowshift_probe: null pointer");
  M2_REQUIRE((rowbytes == 128 || rowbytes == 64) && N % 8 == 0 && N >= 8 && N <= 256 && K % 16 == 0 && shift >= 0 &&
                 rows_total >= shift + 128 && rows_total % 8 == 0,
             M2TTS_E_BADSHAPE, "rowshift_probe: bad arguments");
  const int kblocks = (K * 4 + rowbytes - 1) / rowbytes;
  const size_t smem = (size_t)kblocks * rows_total * rowbytes + (size_t)N * K * 4 + 4096;
  M2_REQUIRE(smem <= 200 * 1024, M2TTS_E_UNSUPPORTED, "rowshift_probe: operands too large");
  M2_CUDA_OK(allow_smem(rowshift_probe_kernel, smem));
  M2_LAUNCH(M2TTS_STAGE_PROBE, rowshift_probe_kernel, 1, 128, smem, (cudaStream_t)stream, A, Bm, D, rows_total, N, K, rowbytes,
            shift, base_offset);
  return M2TTS_OK;
}
