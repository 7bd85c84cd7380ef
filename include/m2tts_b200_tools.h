/*
 * m2tts_b200_tools.h — bring-up and measurement hooks. NOT part of the product ABI.
 *
 * These symbols exist only in m2-tts_b200/lib/libm2tts_b200_tools.so, which is the product sources compiled with
 * -DM2TTS_TOOLS plus the probe kernels under m2-tts_b200/csrc/tools/ (`make -C m2-tts_b200/csrc tools`). The product
 * library libm2tts_b200.so exports none of them. tools/*.py load the tools library through M2TTS_B200_LIB.
 */
#ifndef M2TTS_B200_TOOLS_H_
#define M2TTS_B200_TOOLS_H_

#include "m2tts_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* clock64 phase timestamps of CTA 0 (48 x 8 int64 device buffer, NULL = off): attention / linear kernels, the TF32 fused
 * vocoder stage, the TF32 tap-GEMM */
int m2tts_attention_set_prof(long long* dev_buf);
int m2tts_vocoder_stage_fused_set_prof(long long* dev_buf);
int m2tts_tapgemm_set_prof(long long* dev_buf);
/* switch parts of the upsampling kernel off for timing experiments (1 no stores, 2 no UMMAs, 4 no TMA loads; results
 * are invalid while non-zero). The environment switches M2TTS_ATT_DBG / M2TTS_LIN_DBG / M2TTS_DBG_NOSTORE /
 * M2TTS_LIN_PROF_STAGE are likewise only read by the tools build. */
int m2tts_voc_up_h_set_debug(int mode);

/* TF32 attention on caller-prepared hi/lo planes, optionally dumping the first score / PV tile */
int m2tts_attention_tc_planes(const float* qkv6, float* ctx, const int64_t* lengths, int B, int L, int Lp, int nh,
                              int hd, float* dbg_s, float* dbg_o, m2tts_stream_t stream);

/* K-major swizzled UMMA A operand whose descriptor start address is moved by whole rows inside the swizzle pattern */
int m2tts_rowshift_probe(const float* A, const float* Bm, float* D, int rows_total, int N, int K, int rowbytes,
                         int shift, int base_offset, m2tts_stream_t stream);
/* tcgen05.mma cost per operand configuration */
int m2tts_mma_bench(int mode, int N, int n, int nacc, int elect, long long* out_dev, m2tts_stream_t stream);
/* operand-layout probes. kind::tf32: a[] / b[] = {mode, layout_lbo, layout_sbo, kstep_bytes, mn_major_flag, region_bytes,
 * desc_lbo, desc_sbo}; kind::f16: mode 0 MN-major smem x MN-major smem, 1 TMEM x K-major, 2 K-major x K-major */
int m2tts_umma_probe(const float* A, const float* Bm, float* D, int N, int K, const int* a, const int* b, m2tts_stream_t stream);
int m2tts_umma_probe_f16(const float* A, const float* Bm, float* D, int N, int K, int mode, m2tts_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* M2TTS_B200_TOOLS_H_ */
