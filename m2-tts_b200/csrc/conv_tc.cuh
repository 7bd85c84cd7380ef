// conv_tc.cuh — shared definitions of the tensor-core tap-GEMM convolutions (conv_tc.cu = host side and
// weight packing, conv_tc2.cu = the persistent kernel).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace m2 {

constexpr int CT_BM = 128;            // GEMM rows (input positions) per tile, including the halo
constexpr int CT_HALO = 4;            // rows on each side that are computed but not stored (|tap shift| <= 4)
constexpr int CT_STEP = CT_BM - 2 * CT_HALO;   // 120 output positions per tile; tile starts stay 16-B aligned for TMA
constexpr int CT_CK = 16;             // input channels per pipeline chunk
constexpr uint32_t CT_ABOX = CT_CK * 128;      // one TMA box: 16 channel rows x 32 positions x 4 B
constexpr uint32_t CT_RAW_STAGE = 4u * CT_ABOX;    // 128 positions of fp32 as delivered by TMA = 8 KB
constexpr uint32_t CT_A_STAGE = CT_RAW_STAGE;      // the lo tile (x - trunc_tf32(x)); the raw tile itself is the hi operand

struct TapGemmArgs {
  int CI, L_in, B, n_chunks;
  int tap_shift[3], tap_rows[3], tap_wrow[3], tap_dcol[3];
  int rows_total;            // weight rows per (chunk, plane) image
  int n_cols;                // accumulator columns per tile
  int tmem_cols;             // allocation (power of two), n_bufs * n_cols <= tmem_cols
  int n_bufs;                // accumulator buffers in TMEM (2..4)
  const float* wblob;        // [n_tile][chunk][plane][rows_total][16] (image order)
  int r, co_tile, CO;
  int L_out, Lp_out;
  const float* bias;
  int act;                   // 0 none, 1 leaky_relu(0.1)
  const float* residual; int Lp_res;   // plain fp32 [B][CO][Lp_res] or null
  float* out;                // plain fp32 [B][CO][Lp_out]  (out_cl: channel-last [B][L_out][CO], conv only)
  int out_cl;
  int n_tiles, m_tiles, w_resident;
  int raw_stages, split_stages;
  long long* prof;           // optional bring-up timestamps (CTA 0, epilogue group 0)
  int32_t* status;           // out_cl == 2 (fp16 hi/lo planes out): M2TTS_ST_FP16_RANGE
};

// ---- PTX helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ct_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ct_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ct_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ct_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ct_wait(uint32_t bar, uint32_t parity, int* dbg, int code, int chunk) {
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  if (dbg != nullptr) {
    dbg[0] = code; dbg[1] = chunk; dbg[2] = blockIdx.x; dbg[3] = blockIdx.y; dbg[4] = blockIdx.z; dbg[5] = threadIdx.x;
    __threadfence_system();
  }
  __trap();
}
// NOTE: with 4-byte elements the innermost TMA coordinate must be a multiple of 4 (16-byte aligned box rows);
// an unaligned coordinate raises "illegal instruction" — which is why the taps shift OUTPUT rows, not input boxes.
__device__ __forceinline__ void ct_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void ct_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ct_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ct_mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
// Warp-collective variants (whole warp calls with warp-uniform operands, one elected lane issues): ptxas keeps the
// descriptors in uniform registers and emits a predicated UTCHMMA instead of an ELECT/BRA.U.ANY loop per UMMA.
__device__ __forceinline__ void ct_commit_w(uint32_t bar) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ct_mma_w(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t ct_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ void ct_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void ct_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// accumulator chunk of 16 columns (main + correction halves of a split-folded-into-N GEMM) -> 8 packed fp32 pairs, + the bias
// pairs at bias16 (shared memory, 8-byte aligned)
__device__ __forceinline__ void ct_ld_sum16_pairs(uint32_t t_main, uint32_t t_corr, const float* bias16, uint64_t* v) {
  uint32_t a[16], b[16];
  ct_ld16(t_main, a);
  ct_ld16(t_corr, b);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const uint64_t* bp = reinterpret_cast<const uint64_t*>(bias16);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = f2_add(f2_add(f2_pack_u(a[2 * j], a[2 * j + 1]), f2_pack_u(b[2 * j], b[2 * j + 1])), bp[j]);
}
__device__ __forceinline__ float ct_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

int launch_tapgemm_persistent(const CUtensorMap& tmap, TapGemmArgs& a, int stage, cudaStream_t s);

}  // namespace m2
