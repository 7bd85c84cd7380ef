/*
 * m2tts_b200.h — C ABI of the B200-native (sm_100a) synthesis hot path of m2-tts.
 *
 * The reference (Ryannasr11/m2-tts) is pure Python/PyTorch and has NO native
 * interface; the hot path is the eval-mode forward of src/models/tts_model.py
 * and src/models/components.py.  Every entry point below replaces one module
 * forward of the reference (cited as file:line relative to the reference repo)
 * and is what a ctypes / cffi / pybind stub in that module would bind.
 *
 * Conventions
 *   - plain C, no torch types: raw DEVICE pointers + sizes + a CUDA stream
 *     handle (cudaStream_t passed as void*; NULL = legacy default stream);
 *   - the caller owns every buffer (inputs, outputs, workspace, packed weight
 *     images, the status word); the library allocates nothing and keeps no
 *     pointer after a call returns;
 *   - all work is enqueued on the caller's stream on the CURRENT CUDA device,
 *     no implicit synchronisation (the only host read on the path is the
 *     caller's own read-back of `t_max` after m2tts_length_regulate_count when
 *     max_length is not given);
 *   - return value: 0 = ok, negative = error (M2TTS_E_*); the message is
 *     available per thread from m2tts_last_error_string(); nothing throws;
 *   - floating point tensors are contiguous fp32 unless strides are passed;
 *     `lengths`/`ids` are int64 as on the reference's Python API;
 *   - re-entrant: no mutable global state except the opt-in stage timers and
 *     the launch counter (both atomics / guarded). Kernel selection is a
 *     per-call argument (`precision`), not a process-wide switch.
 *
 * Status word.  Entry points that take `int32_t* status` (device memory, may be
 * NULL = do not report) OR the M2TTS_ST_* bits below into it with an atomic on
 * the rare path; the caller zeroes it and reads it whenever it synchronises.
 * A set bit means the OUTPUT OF THAT CALL IS NOT VALID:
 *   M2TTS_ST_FP16_RANGE  an operand of the 16-bit split (precision 0) left the
 *                        fp16 range (|x| > 65504) or was not finite: run the
 *                        call again with M2TTS_PREC_TF32 (no range limit);
 *   M2TTS_ST_BAD_ID      m2tts_embed_posenc saw an id outside [0, vocab): the
 *                        reference's nn.Embedding raises IndexError.
 *
 * Packed weights.  The tensor-core kernels read weights from re-arranged images
 * (fp16 or TF32 hi/lo planes, swizzled shared-memory images). m2tts_*_pack
 * writes the images of one module into a caller-owned device buffer of
 * m2tts_*_pack_bytes; the forward entry points take that buffer as `packed`
 * (valid for the same shapes and precision until the weights change). With
 * packed == NULL the forward packs into its workspace on every call (stateless
 * form, a few extra small launches).
 */
#ifndef M2TTS_B200_H_
#define M2TTS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* m2tts_stream_t; /* cudaStream_t */

enum {
  M2TTS_OK = 0,
  M2TTS_E_BADSHAPE = -1,    /* non-positive / inconsistent sizes            */
  M2TTS_E_UNSUPPORTED = -2, /* dimension outside what the kernels cover     */
  M2TTS_E_WORKSPACE = -3,   /* workspace too small / misaligned             */
  M2TTS_E_CUDA = -4,        /* a CUDA runtime call or launch failed         */
  M2TTS_E_NULLPTR = -5      /* a required pointer is NULL                   */
};

/* bits of the device status word */
enum {
  M2TTS_ST_FP16_RANGE = 1,
  M2TTS_ST_BAD_ID = 2
};

/* `precision` argument of the GEMM-shaped entry points. Every choice is fp32-faithful (max-abs 1e-4 against the
 * reference's fp32, tests/test_gpu_parity.py); they differ in speed and operand range. */
enum {
  M2TTS_PREC_DEFAULT = -1, /* SPLIT16 unless the environment overrides it for A/B measurements
                              (M2TTS_PRECISION=split16|ffma|tf32, read once)                                   */
  M2TTS_PREC_SPLIT16 = 0,  /* tcgen05 kind::f16, operands as fp16 hi + lo, 3 products, fp32 accumulate.
                              |operand| <= 65504 or M2TTS_ST_FP16_RANGE is raised                              */
  M2TTS_PREC_FFMA = 1,     /* fp32 FFMA kernels (any shape)                                                    */
  M2TTS_PREC_TF32 = 2      /* tcgen05 kind::tf32, operands as TF32 hi + lo, 3 products; no range limit         */
};

/* stage ids for the opt-in per-kernel timers (m2tts_stage_timing_*) */
enum {
  M2TTS_STAGE_EMBED = 0,
  M2TTS_STAGE_PACK = 1,
  M2TTS_STAGE_LN_QKV = 2,
  M2TTS_STAGE_ATTENTION = 3,
  M2TTS_STAGE_OUTPROJ = 4,
  M2TTS_STAGE_FFN1 = 5,
  M2TTS_STAGE_FFN2 = 6,
  M2TTS_STAGE_LN_PROJ = 7,
  M2TTS_STAGE_LAYERNORM = 8,
  M2TTS_STAGE_DURPRED = 9,
  M2TTS_STAGE_LR_COUNT = 10,
  M2TTS_STAGE_LR_GATHER = 11,
  M2TTS_STAGE_VOC_IN = 12,
  M2TTS_STAGE_VOC_UP = 13,
  M2TTS_STAGE_VOC_RES1 = 14,
  M2TTS_STAGE_VOC_RES2 = 15,
  M2TTS_STAGE_VOC_OUT = 16,
  M2TTS_STAGE_PROBE = 17,
  M2TTS_STAGE_VOC_FUSED = 18,
  M2TTS_NUM_STAGES = 19
};

/* ---- weights: raw views of the reference's state_dict tensors ------------- */

/* One TransformerEncoderLayer (components.py:106-140). Shapes as in the
 * state_dict: qkv_w [3H,H] (row order [3, heads, head_dim], components.py:70),
 * out_w [H,H], ffn1_w [F,H], ffn2_w [H,F]. */
typedef struct {
  const float* norm1_w; const float* norm1_b;
  const float* qkv_w;
  const float* out_w;   const float* out_b;
  const float* norm2_w; const float* norm2_b;
  const float* ffn1_w;  const float* ffn1_b;
  const float* ffn2_w;  const float* ffn2_b;
} m2tts_layer_weights;

/* VariancePredictor inside DurationPredictor (components.py:203-223,
 * tts_model.py:92-117): two ConvBlocks (Conv1d k=3 + eval BatchNorm1d + ReLU)
 * and a 1x1 projection. conv*_w [H,H,3], bn vectors [H], proj_w [1,H,1]. */
typedef struct {
  const float* conv_w[2]; const float* conv_b[2];
  const float* bn_w[2];   const float* bn_b[2];
  const float* bn_mean[2]; const float* bn_var[2];
  const float* proj_w;    const float* proj_b;
  float bn_eps;
} m2tts_durpred_weights;

/* SimpleVocoder (tts_model.py:231-297). in_w [C,M,3]; up_w[j] [c_j, c_j/2, 2r_j]
 * (ConvTranspose1d layout), r = {4,4,2,2}; res*_w[j] [c_j/2, c_j/2, 3];
 * out_w [1, C/16, 3]. */
typedef struct {
  const float* in_w;  const float* in_b;
  const float* up_w[4];   const float* up_b[4];
  const float* res1_w[4]; const float* res1_b[4];
  const float* res2_w[4]; const float* res2_b[4];
  const float* out_w; const float* out_b;
  int res_dilation[4]; /* LightweightResBlock conv1 dilation (default 1) */
} m2tts_vocoder_weights;

/* ---- library services ------------------------------------------------------ */

int m2tts_version(void);
const char* m2tts_last_error_string(void);
/* number of kernel launches this library has issued since load */
uint64_t m2tts_launch_count(void);
/* opt-in CUDA-event timers around every kernel launch, keyed by stage id.
 * enable(1) resets and starts collecting, enable(0) stops. read() synchronises
 * the recorded events and returns summed milliseconds + launch counts. */
int m2tts_stage_timing_enable(int on);
int m2tts_stage_timing_read(float* ms_sum, int* launches, int n_stages);
/* Diagnostics: the tensor-core kernels bound every mbarrier wait; on a timeout they store
 * {code, chunk, blockIdx.x, blockIdx.y, blockIdx.z} in pinned host memory and trap. */
int m2tts_debug_words(int* out, int n);
/* fp32 FFMA peak probe (bench.py's roofline denominator for the FFMA kernels): `iters` dependent-chain FFMAs per
 * thread on a full grid; writes nothing but a checksum; returns the flop count through *flops. */
int m2tts_ffma_probe(float* sink, int iters, double* flops, m2tts_stream_t stream);

/* ---- text encoder pieces (tts_model.py:57-89) ------------------------------ */

/* x[b,s,:] = emb[ids[b,s],:]*sqrt(H) + pe[s,:]  (tts_model.py:78-80,
 * components.py:39); mask[b,s] = s < lengths[b] (components.py:226-241) when
 * both `lengths` and `mask` are non-NULL. An id outside [0, vocab) raises
 * M2TTS_ST_BAD_ID in *status (its row is computed from a clamped id and is not valid). */
int m2tts_embed_posenc(const int64_t* ids, const float* emb, const float* pe,
                       const int64_t* lengths, float* x, uint8_t* mask,
                       int B, int S, int H, int vocab, int32_t* status, m2tts_stream_t stream);

/* bytes of scratch one transformer layer needs for [B,L,H] with ffn dim F */
size_t m2tts_transformer_workspace_bytes(int B, int L, int H, int F);
/* packed images of one layer's four weight matrices for `precision` */
size_t m2tts_transformer_pack_bytes(int H, int F, int precision);
int m2tts_transformer_pack(const m2tts_layer_weights* w, int H, int F, int precision, void* packed,
                           size_t packed_bytes, int32_t* status, m2tts_stream_t stream);

/* One pre-LN transformer layer, eval mode (components.py:131-140):
 *   x1 = x + out_proj(softmax(mask(q k^T / sqrt(hd))) v),  q,k,v = split(qkv(LN1(x)))
 *   y  = x1 + W2 relu(W1 LN2(x1) + b1) + b2
 * `lengths` (int64 [B]) masks KEYS only with the reference's finite -1e9 fill
 * (components.py:77-81); NULL = no mask (decoder). x_in may equal x_out.
 * `packed`: NULL or the buffer written by m2tts_transformer_pack for the same (H, F, precision). */
int m2tts_transformer_layer(const m2tts_layer_weights* w, const void* packed, const float* x_in,
                            float* x_out, const int64_t* lengths, int B, int L,
                            int H, int num_heads, int F, float ln_eps, int precision, int32_t* status,
                            void* workspace, size_t workspace_bytes,
                            m2tts_stream_t stream);

/* y = LayerNorm(x) over the last dim (tts_model.py:87) */
int m2tts_layernorm(const float* x, const float* w, const float* b, float* y,
                    int rows, int H, float eps, m2tts_stream_t stream);

/* y[rows,N] = LayerNorm(x) @ W^T + bias (tts_model.py:223-226: decoder.norm then
 * mel_projection, W [N,H]). workspace: m2tts_ln_proj_rows_workspace_bytes(rows,H,N) lets the call run on
 * the tensor cores (normalised rows as hi/lo planes); m2tts_ln_proj_workspace_bytes(H,N) is the FFMA minimum.
 * `packed`: NULL or the image written by m2tts_ln_proj_pack for the same (H, N, precision). */
size_t m2tts_ln_proj_workspace_bytes(int H, int N);
size_t m2tts_ln_proj_rows_workspace_bytes(int rows, int H, int N);
size_t m2tts_ln_proj_pack_bytes(int H, int N, int precision);
int m2tts_ln_proj_pack(const float* W, int H, int N, int precision, void* packed, size_t packed_bytes,
                       int32_t* status, m2tts_stream_t stream);
int m2tts_layernorm_proj(const float* x, const float* ln_w, const float* ln_b,
                         const float* W, const float* bias, const void* packed, float* y, int rows,
                         int H, int N, float eps, int precision, int32_t* status, void* workspace,
                         size_t workspace_bytes, m2tts_stream_t stream);

/* ---- duration predictor (tts_model.py:99-117) ------------------------------ */
/* enc [B,S,H] -> dur [B,S] = softplus(proj(ConvBlock(ConvBlock(enc^T)))); H a multiple of 4 (M2TTS_E_UNSUPPORTED otherwise) */
int m2tts_duration_predictor(const m2tts_durpred_weights* w, const float* enc,
                             float* dur, int B, int S, int H,
                             m2tts_stream_t stream);

/* ---- length regulator (tts_model.py:126-178) ------------------------------- */
/* Pass 1: n[b,s] = trunc(dur[b,s]) if > 0 else 0 (python int(), tts_model.py:150-151);
 * cum[b,s] = inclusive prefix sum (int32, saturating); frames[b] = sum_s n[b,s];
 * *t_max = max_b max(1, frames[b]) (tts_model.py:158-166); *status bit0 = a NaN
 * duration was seen (python raises ValueError), bit1 = +-inf (OverflowError),
 * bit2 = a frame count overflowed int32 (this word is the regulator's own, not the
 * M2TTS_ST_* word). All outputs are device memory. */
int m2tts_length_regulate_count(const float* dur, int B, int S, int32_t* cum,
                                int32_t* frames, int32_t* t_max, int32_t* status,
                                m2tts_stream_t stream);
/* Pass 2: out[b,j,:] = enc[b,s(j),:] with s(j) = first s: cum[b,s] > j, zero rows
 * for j >= frames[b]; truncated at T (tts_model.py:168-178). index[b,j] = s(j)
 * or -1 (optional, may be NULL). */
int m2tts_length_regulate_gather(const float* enc, const int32_t* cum,
                                 const int32_t* frames, float* out,
                                 int32_t* index, int B, int S, int H, int T,
                                 m2tts_stream_t stream);

/* ---- vocoder (tts_model.py:279-297) ---------------------------------------- */
size_t m2tts_vocoder_workspace_bytes(int B, int T, int M, int C);
/* packed images of every convolution of the vocoder for `precision` */
size_t m2tts_vocoder_pack_bytes(int M, int C, int precision);
int m2tts_vocoder_pack(const m2tts_vocoder_weights* w, int M, int C, int precision, void* packed,
                       size_t packed_bytes, int32_t* status, m2tts_stream_t stream);
/* Which kernels m2tts_vocoder_forward runs for (M, C, precision, res_dilation[4] or NULL = all 1): kinds[0] = input conv
 * (0 fp32 FFMA, 1 TF32 tap-GEMM, 2 channel-last 16-bit split), kinds[1..4] = stages 0..3 (0 FFMA; 1 TF32 tap-GEMMs; 2 fused
 * TF32; 3 fused 16-bit split; 4 / 5 voc_up_h + fused ResBlock / two conv kernels; 6 / 7 the same behind a TF32 upsampler).
 * A pure function of its arguments — the same choice m2tts_vocoder_pack makes. */
int m2tts_vocoder_plan(int M, int C, int precision, const int* res_dilation, int* kinds);
/* mel element (b,m,t) is read at mel[b*stride_b + m*stride_m + t*stride_t]
 * (so both a contiguous [B,M,T] tensor and the transposed view of the decoder's
 * [B,T,M] output, tts_model.py:390, are accepted without a copy).
 * audio [B,1,64*T] contiguous. `packed`: NULL or the buffer written by
 * m2tts_vocoder_pack for the same (M, C, precision) and res_dilation. */
int m2tts_vocoder_forward(const m2tts_vocoder_weights* w, const void* packed, const float* mel,
                          int64_t stride_b, int64_t stride_m, int64_t stride_t,
                          float* audio, int B, int T, int M, int C, int precision, int32_t* status,
                          void* workspace, size_t workspace_bytes,
                          m2tts_stream_t stream);

/* Per-stage entry points (unit tests / profiling). Channel-first fp32.
 * conv1d k=3, "same" zero padding = dilation (components.py:181-190,
 * tts_model.py:246,272). act: 0 none, 1 leaky_relu(0.1), 2 tanh.
 * residual (may be NULL) is added AFTER the conv (components.py:200).
 * x element (b,ci,t) at x[b*xs_b + ci*xs_c + t*xs_t]; y contiguous [B,CO,L].
 * workspace: m2tts_conv_workspace_bytes(CI,CO,3). */
size_t m2tts_conv_workspace_bytes(int CI, int CO, int taps);
int m2tts_conv1d_k3(const float* x, int64_t xs_b, int64_t xs_c, int64_t xs_t,
                    const float* w, const float* bias, const float* residual,
                    float* y, int B, int CI, int CO, int L, int dilation, int act,
                    void* workspace, size_t workspace_bytes, m2tts_stream_t stream);
/* ConvTranspose1d(CI, CO, k=2r, stride=r, padding=r/2) + leaky_relu(0.1)
 * (tts_model.py:255-263,291), r in {2,4}; x [B,CI,L] -> y [B,CO,r*L]. */
int m2tts_conv_transpose1d_lrelu(const float* x, const float* w, const float* bias,
                                 float* y, int B, int CI, int CO, int L, int r,
                                 m2tts_stream_t stream);

/* Tensor-core (tcgen05, 3xTF32) variants of the two per-stage entry points, plain fp32 tensors in
 * and out, same semantics (act: 0 none, 1 leaky_relu(0.1); residual added after the conv).
 * Eligibility: conv CI % 16 == 0, CO a multiple of 16 that is < 64 or a multiple of 64, dilation <= 4;
 * transposed conv CI % 16 == 0 and (r == 4, CO % 32 == 0) or (r == 2, CO % 16 == 0); else M2TTS_E_UNSUPPORTED. workspace: m2tts_conv_tc_workspace_bytes(B, CI, CO, L, r). */
size_t m2tts_conv_tc_workspace_bytes(int B, int CI, int CO, int L, int r);
int m2tts_conv1d_k3_tc(const float* x, const float* w, const float* bias, const float* residual, float* y,
                       int B, int CI, int CO, int L, int dilation, int act, void* workspace,
                       size_t workspace_bytes, m2tts_stream_t stream);
int m2tts_conv_transpose1d_lrelu_tc(const float* x, const float* w, const float* bias, float* y, int B,
                                    int CI, int CO, int L, int r, void* workspace, size_t workspace_bytes,
                                    m2tts_stream_t stream);

/* One whole narrow vocoder stage fused into a single tcgen05 kernel, CHANNEL-LAST activations
 * (tts_model.py:289-292 for one (upsample, resblock) pair, plus :295 for the last stage):
 *   x [B][L][2C]  ->  u = leaky_relu(ConvTranspose1d(2C, C, k=4, s=2, p=1)(x), 0.1)
 *                 ->  y = u + conv2(leaky_relu(conv1(u), 0.1))            (components.py:196-200)
 *   out_w == NULL: y written channel-last [B][2L][C];
 *   out_w != NULL: audio [B][2L] = tanh(Conv1d(C, 1, k=3, p=1)(y))  (out_w [1][C][3], out_b [1]).
 * C in {16, 32}; weights in the reference's state_dict layouts (up_w [2C][C][4], res*_w [C][C][3]).
 * TF32 split; workspace: m2tts_vocoder_stage_fused_workspace_bytes(C). */
size_t m2tts_vocoder_stage_fused_workspace_bytes(int C);
int m2tts_vocoder_stage_fused(const float* x, const float* up_w, const float* up_b, const float* res1_w,
                              const float* res1_b, const float* res2_w, const float* res2_b,
                              const float* out_w, const float* out_b, float* y, int B, int C, int L,
                              void* workspace, size_t workspace_bytes, m2tts_stream_t stream);

/* The same stage with the 16-bit split: what m2tts_vocoder_forward uses for its narrow stages by default.
 * C in {16, 32}, and C = 8 as the LAST stage only (out_w != NULL: runs zero-padded in the 16-channel kernel — the stage-1 model).
 * Same arguments plus the status word; workspace: m2tts_vocoder_stage_fused_h_workspace_bytes(B, C, L). */
size_t m2tts_vocoder_stage_fused_h_workspace_bytes(int B, int C, int L);
int m2tts_vocoder_stage_fused_h(const float* x, const float* up_w, const float* up_b, const float* res1_w,
                                const float* res1_b, const float* res2_w, const float* res2_b,
                                const float* out_w, const float* out_b, float* y, int B, int C, int L,
                                int32_t* status, void* workspace, size_t workspace_bytes, m2tts_stream_t stream);

/* A whole LightweightResBlock (components.py:177-200), y = x + conv2(leaky_relu(conv1(x), 0.1)), as one tcgen05 kernel with
 * the 16-bit split; x / y fp32 CHANNEL-LAST [B][L][C], C in {32, 64}, kernel 3, dilation 1; weights in state_dict layout [C][C][3].
 * workspace: m2tts_resblock_fused_h_workspace_bytes(B, C, L). */
size_t m2tts_resblock_fused_h_workspace_bytes(int B, int C, int L);
int m2tts_resblock_fused_h(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* y,
                           int B, int C, int L, int32_t* status, void* workspace, size_t workspace_bytes,
                           m2tts_stream_t stream);

/* One convolution of the widest LightweightResBlock (components.py:181-200): y = act(conv1d(x, w, b, padding 1)) (+ residual),
 * C = 128, kernel 3, dilation 1, 16-bit split, channel-last operands. x / residual fp32 CHANNEL-LAST [B][L][C]; act 0 none,
 * 1 leaky_relu(0.1); y fp32 channel-first [B][C][L] (out_cl 0) or channel-last [B][L][C] (out_cl 1).
 * workspace: m2tts_conv1d_k3_h_workspace_bytes(B, C, L). */
size_t m2tts_conv1d_k3_h_workspace_bytes(int B, int C, int L);
int m2tts_conv1d_k3_h(const float* x, const float* w, const float* b, const float* residual, float* y, int B, int C, int L,
                      int act, int out_cl, int32_t* status, void* workspace, size_t workspace_bytes, m2tts_stream_t stream);

/* One upsampling layer of the vocoder with its activation (tts_model.py:255-263,291):
 * y = leaky_relu(conv_transpose1d(x, w, b, stride 4, padding 2), 0.1), kernel 8, CI in {64, 128, 256}, CO = CI / 2, 16-bit split,
 * channel-last operands: x fp32 [B][L][CI], w [CI][CO][8] (state_dict layout), y fp32 [B][4L][CO].
 * workspace: m2tts_conv_transpose_x4_h_workspace_bytes(B, CI, L). */
size_t m2tts_conv_transpose_x4_h_workspace_bytes(int B, int CI, int L);
int m2tts_conv_transpose_x4_h(const float* x, const float* w, const float* b, float* y, int B, int CI, int L,
                              int32_t* status, void* workspace, size_t workspace_bytes, m2tts_stream_t stream);

/* ---- after the path: waveform -> PCM16 on the device (scripts/synthesize.py:147-155 saves through
 * src/utils/audio.py:154-180; clip to [-1,1], x32767, round half to even). audio/pcm: n samples, device memory. */
int m2tts_pcm16(const float* audio, int16_t* pcm, long long n, m2tts_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* M2TTS_B200_H_ */
