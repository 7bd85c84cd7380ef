// transformer.cu — C-ABI entry points for one pre-LN transformer layer and the decoder's
// final LayerNorm + mel projection. Composition (all on the caller's stream):
//   [weight images: per call, or once by m2tts_transformer_pack] -> [LN1 + QKV] -> flash attention -> [out_proj + residual]
//                -> [LN2 + FFN1 + ReLU] -> [FFN2 + residual]
// Reference: components.py:131-140 (TransformerEncoderLayer._forward), :59-90, :103.
#include "common.cuh"
#include <math.h>

using namespace m2;

namespace {
struct LayerWs {
  float *wqkv_t, *wo_t, *w1_t, *w2_t;  // packed W^T
  float *q, *k, *v, *ctx, *x1, *hid;
  float *w_planes[4];   // tensor-core path: hi/lo planes of qkv / out_proj / ffn1 / ffn2 weights
  float *xn;            // tensor-core path: LayerNorm output as hi/lo planes [2][R][H]
  int Lp;
};

bool carve_layer(void* ws, size_t bytes, int B, int L, int H, int F, LayerWs* o) {
  Carver cv(ws, bytes);
  const int Lp = (L + 3) & ~3;
  o->Lp = Lp;
  o->wqkv_t = cv.take<float>((size_t)H * 3 * H);
  o->wo_t = cv.take<float>((size_t)H * H);
  o->w1_t = cv.take<float>((size_t)H * F);
  o->w2_t = cv.take<float>((size_t)F * H);
  // six planes: the tensor-core attention wants {Q,K,V} x {hi,lo}; the FFMA kernel uses the first three
  o->q = cv.take<float>((size_t)6 * B * H * Lp);
  o->k = o->q + (size_t)B * H * Lp;
  o->v = o->k + (size_t)B * H * Lp;
  o->ctx = cv.take<float>((size_t)2 * B * L * H);   // plain [R,H] or hi/lo planes
  o->x1 = cv.take<float>((size_t)B * L * H);
  o->hid = cv.take<float>((size_t)2 * B * L * F);   // plain [R,F] or hi/lo planes
  o->xn = cv.take<float>((size_t)2 * B * L * H);
  o->w_planes[0] = cv.take<float>((size_t)2 * 3 * H * H);
  o->w_planes[1] = cv.take<float>((size_t)2 * H * H);
  o->w_planes[2] = cv.take<float>((size_t)2 * F * H);
  o->w_planes[3] = cv.take<float>((size_t)2 * H * F);
  return cv.ok();
}
}  // namespace

// Which family of kernels a layer of this shape runs at this precision (pack and forward must agree).
namespace {
enum { FAM_H = 0, FAM_TF32 = 1, FAM_FFMA = 2 };
int layer_family(int prec, int H, int F, int hd) {
  if (prec == M2TTS_PREC_SPLIT16 && attention_tc_supported(hd) && linear_h_eligible(H, 3 * H) && linear_h_eligible(H, H) &&
      linear_h_eligible(H, F) && linear_h_eligible(F, H) && (hd % 16 == 0))
    return FAM_H;
  if (prec != M2TTS_PREC_FFMA && attention_tc_supported(hd) && linear_tc_eligible(H, 3 * H) && linear_tc_eligible(H, H) &&
      linear_tc_eligible(H, F) && linear_tc_eligible(F, H) && (H % 16 == 0))
    return FAM_TF32;
  return FAM_FFMA;
}
// images of the four weight matrices: fp16 / TF32 hi-lo planes (2 x elements) or the FFMA transposes (1 x elements)
size_t layer_pack_floats(int H, int F) { return 2 * ((size_t)3 * H * H + (size_t)H * H + 2 * (size_t)H * F); }

int pack_layer(const m2tts_layer_weights* w, int fam, int H, int F, float* const* dst, int32_t* status, cudaStream_t s) {
  const float* srcs[4] = {w->qkv_w, w->out_w, w->ffn1_w, w->ffn2_w};
  const long long ns[4] = {(long long)3 * H * H, (long long)H * H, (long long)F * H, (long long)H * F};
  if (fam == FAM_H) {
    void* d[4] = {dst[0], dst[1], dst[2], dst[3]};
    return launch_w_split_h(srcs, d, ns, 4, status, s);
  }
  if (fam == FAM_TF32) return launch_w_split(srcs, dst, ns, 4, s);
  PackJob jobs[4] = {{w->qkv_w, dst[0], 3 * H, H}, {w->out_w, dst[1], H, H}, {w->ffn1_w, dst[2], F, H}, {w->ffn2_w, dst[3], H, F}};
  return launch_pack_transpose(jobs, 4, s);
}
void carve_pack(float* base, int H, int F, float** dst) {
  dst[0] = base;
  dst[1] = dst[0] + (size_t)2 * 3 * H * H;
  dst[2] = dst[1] + (size_t)2 * H * H;
  dst[3] = dst[2] + (size_t)2 * F * H;
}
}  // namespace

extern "C" size_t m2tts_transformer_workspace_bytes(int B, int L, int H, int F) {
  if (B <= 0 || L <= 0 || H <= 0 || F <= 0) return 0;
  const size_t Lp = (size_t)((L + 7) & ~7);
  size_t fl = 3 * ((size_t)H * 3 * H + (size_t)H * H + 2 * (size_t)H * F) + 6 * (size_t)B * H * Lp +
              5 * (size_t)B * L * H + 2 * (size_t)B * L * F;
  return fl * sizeof(float) + 32 * 256;
}

extern "C" size_t m2tts_transformer_pack_bytes(int H, int F, int precision) {
  (void)precision;
  if (H <= 0 || F <= 0) return 0;
  return align_up(layer_pack_floats(H, F) * sizeof(float), 256);
}

extern "C" int m2tts_transformer_pack(const m2tts_layer_weights* w, int H, int F, int precision, void* packed, size_t packed_bytes,
                                      int32_t* status, m2tts_stream_t stream) {
  M2_REQUIRE(w && packed, M2TTS_E_NULLPTR, "transformer_pack: null pointer");
  M2_REQUIRE(w->qkv_w && w->out_w && w->ffn1_w && w->ffn2_w, M2TTS_E_NULLPTR, "transformer_pack: null weight pointer");
  M2_REQUIRE(H > 0 && F > 0, M2TTS_E_BADSHAPE, "transformer_pack: H=%d F=%d", H, F);
  M2_REQUIRE(packed_bytes >= m2tts_transformer_pack_bytes(H, F, precision) && (((uintptr_t)packed) & 255) == 0, M2TTS_E_WORKSPACE,
             "transformer_pack: buffer too small (%zu B) or not 256-B aligned", packed_bytes);
  // head_dim does not enter the images, only the choice of kernel family: the images are written for the family a
  // tensor-core head dim gets; a forward call whose head dim lands in another family ignores `packed` and packs per call
  float* dst[4];
  carve_pack((float*)packed, H, F, dst);
  const int fam = layer_family(resolve_precision(precision), H, F, 16);
  return pack_layer(w, fam, H, F, dst, status, (cudaStream_t)stream);
}

extern "C" int m2tts_transformer_layer(const m2tts_layer_weights* w, const void* packed, const float* x_in, float* x_out,
                                       const int64_t* lengths, int B, int L, int H, int num_heads, int F,
                                       float ln_eps, int precision, int32_t* status, void* workspace, size_t workspace_bytes,
                                       m2tts_stream_t stream) {
  M2_REQUIRE(w && x_in && x_out && workspace, M2TTS_E_NULLPTR, "transformer_layer: null pointer");
  M2_REQUIRE(w->norm1_w && w->norm1_b && w->qkv_w && w->out_w && w->out_b && w->norm2_w && w->norm2_b &&
                 w->ffn1_w && w->ffn1_b && w->ffn2_w && w->ffn2_b,
             M2TTS_E_NULLPTR, "transformer_layer: null weight pointer");
  M2_REQUIRE(B > 0 && L > 0 && H > 0 && F > 0 && num_heads > 0, M2TTS_E_BADSHAPE,
             "transformer_layer: B=%d L=%d H=%d heads=%d F=%d", B, L, H, num_heads, F);
  M2_REQUIRE(H % num_heads == 0, M2TTS_E_BADSHAPE, "transformer_layer: hidden_dim %d not divisible by %d heads",
             H, num_heads);
  const int hd = H / num_heads;
  M2_REQUIRE(hd % 8 == 0 && hd <= 64, M2TTS_E_UNSUPPORTED,
             "transformer_layer: head_dim %d unsupported (multiples of 8 up to 64)", hd);
  M2_REQUIRE(H <= 256 && F <= 256, M2TTS_E_UNSUPPORTED, "transformer_layer: H=%d F=%d exceed 256", H, F);
  M2_REQUIRE((long long)B * L < (1ll << 31), M2TTS_E_UNSUPPORTED, "transformer_layer: B*L too large");
  LayerWs ws;
  M2_REQUIRE(carve_layer(workspace, workspace_bytes, B, L, H, F, &ws), M2TTS_E_WORKSPACE,
             "transformer_layer: workspace too small (%zu B, need %zu) or not 256-B aligned", workspace_bytes,
             m2tts_transformer_workspace_bytes(B, L, H, F));
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  const int prec = resolve_precision(precision);
  const int fam = layer_family(prec, H, F, hd);
  // weight images: the caller's packed buffer when it holds this family's images, else written here
  float* wimg[4];
  if (packed != nullptr && fam == layer_family(prec, H, F, 16)) {
    M2_REQUIRE((((uintptr_t)packed) & 255) == 0, M2TTS_E_WORKSPACE, "transformer_layer: packed weights must be 256-B aligned");
    carve_pack((float*)const_cast<void*>(packed), H, F, wimg);
  } else {
    wimg[0] = ws.w_planes[0]; wimg[1] = ws.w_planes[1]; wimg[2] = ws.w_planes[2]; wimg[3] = ws.w_planes[3];
    if ((rc = pack_layer(w, fam, H, F, wimg, status, s))) return rc;
  }

  const int R = B * L;
  // ---- 16-bit split path (default): every GEMM of the layer on tcgen05 with fp16 hi/lo operand planes, persistent
  //      weight-resident linear kernels (lin_h.cu) and the warp-specialised attention (attention_h.cu) ----
  if (fam == FAM_H) {
    const int Lp = (L + 7) & ~7;
    const bool ln_fused = H <= 96;      // LayerNorm + split by the GEMM's own producer warps (lin_h.cu), else a separate ln_split_h launch
    if (!ln_fused && (rc = launch_ln_split_h(x_in, w->norm1_w, w->norm1_b, ws.xn, R, H, ln_eps, status, s))) return rc;
    {  // attention operand planes = split(LN1(x) Wqkv^T)
      LinHParams q{};
      q.R = R; q.K = H; q.N = 3 * H; q.mode = 3; q.qkvh = ws.q; q.plane_stride = (long long)B * H * Lp; q.status = status;
      q.L = L; q.nh = num_heads; q.hd = hd; q.Lp = Lp; q.qscale = (float)((1.0 / sqrt((double)hd)) * 1.4426950408889634);
      if (ln_fused) { q.ln_x = x_in; q.ln_w = w->norm1_w; q.ln_b = w->norm1_b; q.ln_eps = ln_eps; }
      if ((rc = launch_linear_h(ws.xn, wimg[0], q, M2TTS_STAGE_LN_QKV, s))) return rc;
    }
    if ((rc = launch_attention_h(ws.q, nullptr, lengths, B, L, Lp, num_heads, hd, s, nullptr, ws.ctx, status))) return rc;
    {  // x1 = x + ctx Wo^T + bo
      LinHParams q{};
      q.R = R; q.K = H; q.N = H; q.mode = 0; q.bias = w->out_b; q.residual = x_in; q.ldr = H; q.y = ws.x1; q.ldy = H;
      if ((rc = launch_linear_h(ws.ctx, wimg[1], q, M2TTS_STAGE_OUTPROJ, s))) return rc;
    }
    if (!ln_fused && (rc = launch_ln_split_h(ws.x1, w->norm2_w, w->norm2_b, ws.xn, R, H, ln_eps, status, s))) return rc;
    {  // hid = relu(LN2(x1) W1^T + b1) as planes
      LinHParams q{};
      q.R = R; q.K = H; q.N = F; q.mode = 1; q.bias = w->ffn1_b; q.relu = 1; q.y_planes = ws.hid; q.status = status;
      if (ln_fused) { q.ln_x = ws.x1; q.ln_w = w->norm2_w; q.ln_b = w->norm2_b; q.ln_eps = ln_eps; }
      if ((rc = launch_linear_h(ws.xn, wimg[2], q, M2TTS_STAGE_FFN1, s))) return rc;
    }
    {  // y = x1 + hid W2^T + b2
      LinHParams q{};
      q.R = R; q.K = F; q.N = H; q.mode = 0; q.bias = w->ffn2_b; q.residual = ws.x1; q.ldr = H; q.y = x_out; q.ldy = H;
      if ((rc = launch_linear_h(ws.hid, wimg[3], q, M2TTS_STAGE_FFN2, s))) return rc;
    }
    return M2TTS_OK;
  }
  // ---- TF32 split: every GEMM of the layer on tcgen05 (3xTF32), operands as hi/lo planes; with M2TTS_PREC_SPLIT16 on
  //      shapes the persistent 16-bit linear kernels do not take, the attention still runs the 16-bit split ----
  if (fam == FAM_TF32) {
    if ((rc = launch_ln_split(x_in, w->norm1_w, w->norm1_b, ws.xn, R, H, ln_eps, s))) return rc;
    const bool half_planes = prec == M2TTS_PREC_SPLIT16 && (hd % 16 == 0);
    const int Lp = half_planes ? ((L + 7) & ~7) : ws.Lp;
    {  // attention operand planes = split(LN1(x) Wqkv^T)
      LinTcArgs a{};
      a.R = R; a.L = L; a.K = H; a.N = 3 * H; a.mode = half_planes ? 3 : 2; a.qkv6 = ws.q; a.plane_stride = (long long)B * H * Lp;
      a.nh = num_heads; a.hd = hd; a.Lp = Lp; a.qscale = (float)((1.0 / sqrt((double)hd)) * 1.4426950408889634); a.status = status;
      if ((rc = launch_linear_tc(ws.xn, wimg[0], a, B, M2TTS_STAGE_LN_QKV, s))) return rc;
    }
    if (half_planes) {
      if ((rc = launch_attention_h(ws.q, ws.ctx, lengths, B, L, Lp, num_heads, hd, s, ws.ctx + (size_t)R * H))) return rc;
    } else {
      if ((rc = launch_attention_tc(ws.q, ws.ctx, lengths, B, L, ws.Lp, num_heads, hd, s, nullptr, nullptr,
                                    ws.ctx + (size_t)R * H))) return rc;
    }
    {  // x1 = x + ctx Wo^T + bo
      LinTcArgs a{};
      a.R = R; a.L = L; a.K = H; a.N = H; a.mode = 0; a.bias = w->out_b; a.residual = x_in; a.ldr = H; a.y = ws.x1; a.ldy = H;
      if ((rc = launch_linear_tc(ws.ctx, wimg[1], a, B, M2TTS_STAGE_OUTPROJ, s))) return rc;
    }
    if ((rc = launch_ln_split(ws.x1, w->norm2_w, w->norm2_b, ws.xn, R, H, ln_eps, s))) return rc;
    {  // hid = relu(LN2(x1) W1^T + b1) as planes
      LinTcArgs a{};
      a.R = R; a.L = L; a.K = H; a.N = F; a.mode = 1; a.bias = w->ffn1_b; a.relu = 1; a.y_planes = ws.hid;
      if ((rc = launch_linear_tc(ws.xn, wimg[2], a, B, M2TTS_STAGE_FFN1, s))) return rc;
    }
    {  // y = x1 + hid W2^T + b2
      LinTcArgs a{};
      a.R = R; a.L = L; a.K = F; a.N = H; a.mode = 0; a.bias = w->ffn2_b; a.residual = ws.x1; a.ldr = H; a.y = x_out; a.ldy = H;
      if ((rc = launch_linear_tc(ws.hid, wimg[3], a, B, M2TTS_STAGE_FFN2, s))) return rc;
    }
    return M2TTS_OK;
  }

  // ---- fp32 FFMA linear layers (any shape); attention on the TF32 tensor-core kernel unless FFMA was asked for ----
  const bool use_tc = prec != M2TTS_PREC_FFMA && attention_tc_supported(hd);
  {  // q,k,v = split(LN1(x) Wqkv^T)
    RowGemmArgs a{};
    a.x = x_in; a.ldx = H; a.ln_w = w->norm1_w; a.ln_b = w->norm1_b; a.eps = ln_eps;
    a.wt = wimg[0]; a.qkv_mode = use_tc ? 2 : 1; a.q = ws.q; a.k = ws.k; a.v = ws.v;
    a.qscale = (float)((1.0 / sqrt((double)hd)) * 1.4426950408889634);
    a.plane_stride = (long long)B * H * ws.Lp;
    a.L = L; a.Lp = ws.Lp; a.nh = num_heads; a.hd = hd; a.R = R; a.K = H; a.N = 3 * H;
    a.stage = M2TTS_STAGE_LN_QKV;
    if ((rc = launch_rowgemm(a, s))) return rc;
  }
  if (use_tc) {
    if ((rc = launch_attention_tc(ws.q, ws.ctx, lengths, B, L, ws.Lp, num_heads, hd, s))) return rc;
  } else {
    if ((rc = launch_attention(ws.q, ws.k, ws.v, ws.ctx, lengths, B, L, ws.Lp, num_heads, hd, s))) return rc;
  }
  {  // x1 = x + ctx Wo^T + bo
    RowGemmArgs a{};
    a.x = ws.ctx; a.ldx = H; a.wt = wimg[1]; a.bias = w->out_b; a.residual = x_in; a.ldr = H;
    a.y = ws.x1; a.ldy = H; a.R = R; a.K = H; a.N = H; a.stage = M2TTS_STAGE_OUTPROJ;
    if ((rc = launch_rowgemm(a, s))) return rc;
  }
  {  // hid = relu(LN2(x1) W1^T + b1)
    RowGemmArgs a{};
    a.x = ws.x1; a.ldx = H; a.ln_w = w->norm2_w; a.ln_b = w->norm2_b; a.eps = ln_eps;
    a.wt = wimg[2]; a.bias = w->ffn1_b; a.relu = 1; a.y = ws.hid; a.ldy = F;
    a.R = R; a.K = H; a.N = F; a.stage = M2TTS_STAGE_FFN1;
    if ((rc = launch_rowgemm(a, s))) return rc;
  }
  {  // y = x1 + hid W2^T + b2
    RowGemmArgs a{};
    a.x = ws.hid; a.ldx = F; a.wt = wimg[3]; a.bias = w->ffn2_b; a.residual = ws.x1; a.ldr = H;
    a.y = x_out; a.ldy = H; a.R = R; a.K = F; a.N = H; a.stage = M2TTS_STAGE_FFN2;
    if ((rc = launch_rowgemm(a, s))) return rc;
  }
  return M2TTS_OK;
}

namespace {
int ln_proj_family(int prec, int H, int N) {
  if (prec == M2TTS_PREC_SPLIT16 && linear_h_eligible(H, N)) return FAM_H;
  if (prec != M2TTS_PREC_FFMA && linear_tc_eligible(H, N) && (N % 16 == 0) && (H % 4 == 0)) return FAM_TF32;
  return FAM_FFMA;
}
int pack_ln_proj(const float* W, int fam, int H, int N, float* wt, int32_t* status, cudaStream_t s) {
  const float* srcs[1] = {W};
  const long long ns[1] = {(long long)N * H};
  if (fam == FAM_H) { void* d[1] = {wt}; return launch_w_split_h(srcs, d, ns, 1, status, s); }
  if (fam == FAM_TF32) { float* d[1] = {wt}; return launch_w_split(srcs, d, ns, 1, s); }
  PackJob job{W, wt, N, H};
  return launch_pack_transpose(&job, 1, s);
}
}  // namespace

extern "C" size_t m2tts_ln_proj_workspace_bytes(int H, int N) {
  if (H <= 0 || N <= 0) return 0;
  return align_up((size_t)3 * H * N * sizeof(float), 256) + 1024;   // W^T (FFMA) or W hi/lo planes (tensor cores)
}

extern "C" size_t m2tts_ln_proj_rows_workspace_bytes(int rows, int H, int N) {
  if (rows <= 0 || H <= 0 || N <= 0) return 0;
  return m2tts_ln_proj_workspace_bytes(H, N) + align_up((size_t)2 * rows * H * sizeof(float), 256) + 256;
}

extern "C" size_t m2tts_ln_proj_pack_bytes(int H, int N, int precision) {
  (void)precision;
  if (H <= 0 || N <= 0) return 0;
  return align_up((size_t)3 * H * N * sizeof(float), 256);
}

extern "C" int m2tts_ln_proj_pack(const float* W, int H, int N, int precision, void* packed, size_t packed_bytes, int32_t* status,
                                  m2tts_stream_t stream) {
  M2_REQUIRE(W && packed, M2TTS_E_NULLPTR, "ln_proj_pack: null pointer");
  M2_REQUIRE(H > 0 && N > 0, M2TTS_E_BADSHAPE, "ln_proj_pack: H=%d N=%d", H, N);
  M2_REQUIRE(packed_bytes >= m2tts_ln_proj_pack_bytes(H, N, precision) && (((uintptr_t)packed) & 255) == 0, M2TTS_E_WORKSPACE,
             "ln_proj_pack: buffer too small (%zu B) or not 256-B aligned", packed_bytes);
  return pack_ln_proj(W, ln_proj_family(resolve_precision(precision), H, N), H, N, (float*)packed, status, (cudaStream_t)stream);
}

extern "C" int m2tts_layernorm_proj(const float* x, const float* ln_w, const float* ln_b, const float* W,
                                    const float* bias, const void* packed, float* y, int rows, int H, int N, float eps,
                                    int precision, int32_t* status, void* workspace, size_t workspace_bytes, m2tts_stream_t stream) {
  M2_REQUIRE(x && ln_w && ln_b && W && y && workspace, M2TTS_E_NULLPTR, "layernorm_proj: null pointer");
  M2_REQUIRE(rows > 0 && H > 0 && N > 0, M2TTS_E_BADSHAPE, "layernorm_proj: rows=%d H=%d N=%d", rows, H, N);
  Carver cv(workspace, workspace_bytes);
  float* wt = cv.take<float>((size_t)3 * H * N);
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "layernorm_proj: workspace too small or not 256-B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  // tensor-core path when the caller's workspace also has room for the normalised rows as hi/lo planes
  float* xn = cv.take<float>((size_t)2 * rows * H);
  int fam = ln_proj_family(resolve_precision(precision), H, N);
  if (!cv.ok()) fam = FAM_FFMA;
  if (packed != nullptr && cv.ok()) {      // images written by m2tts_ln_proj_pack (same family rule; FFMA needs no xn)
    M2_REQUIRE((((uintptr_t)packed) & 255) == 0, M2TTS_E_WORKSPACE, "layernorm_proj: packed weights must be 256-B aligned");
    wt = (float*)const_cast<void*>(packed);
  } else if ((rc = pack_ln_proj(W, fam, H, N, wt, status, s))) {
    return rc;
  }
  if (fam == FAM_H) {
    LinHParams q{};
    q.R = rows; q.K = H; q.N = N; q.mode = 0; q.bias = bias; q.y = y; q.ldy = N; q.status = status;
    if (H <= 96) { q.ln_x = x; q.ln_w = ln_w; q.ln_b = ln_b; q.ln_eps = eps; }      // LayerNorm + split by the GEMM's producer warps
    else if ((rc = launch_ln_split_h(x, ln_w, ln_b, xn, rows, H, eps, status, s))) return rc;
    return launch_linear_h(xn, wt, q, M2TTS_STAGE_LN_PROJ, s);
  }
  if (fam == FAM_TF32) {
    if ((rc = launch_ln_split(x, ln_w, ln_b, xn, rows, H, eps, s))) return rc;
    LinTcArgs a{};
    a.R = rows; a.L = rows; a.K = H; a.N = N; a.mode = 0; a.bias = bias; a.y = y; a.ldy = N;
    return launch_linear_tc(xn, wt, a, 1, M2TTS_STAGE_LN_PROJ, s);
  }
  RowGemmArgs a{};
  a.x = x; a.ldx = H; a.ln_w = ln_w; a.ln_b = ln_b; a.eps = eps; a.wt = wt; a.bias = bias;
  a.y = y; a.ldy = N; a.R = rows; a.K = H; a.N = N; a.stage = M2TTS_STAGE_LN_PROJ;
  return launch_rowgemm(a, s);
}
