// transformer.cu — C-ABI entry points for one pre-LN transformer layer and the decoder's
// final LayerNorm + mel projection. Composition (all on the caller's stream):
//   pack weights -> [LN1 + QKV] -> flash attention -> [out_proj + residual]
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

extern "C" size_t m2tts_transformer_workspace_bytes(int B, int L, int H, int F) {
  if (B <= 0 || L <= 0 || H <= 0 || F <= 0) return 0;
  const size_t Lp = (size_t)((L + 3) & ~3);
  size_t fl = 3 * ((size_t)H * 3 * H + (size_t)H * H + 2 * (size_t)H * F) + 6 * (size_t)B * H * Lp +
              5 * (size_t)B * L * H + 2 * (size_t)B * L * F;
  return fl * sizeof(float) + 32 * 256;
}

extern "C" int m2tts_transformer_layer(const m2tts_layer_weights* w, const float* x_in, float* x_out,
                                       const int64_t* lengths, int B, int L, int H, int num_heads, int F,
                                       float ln_eps, void* workspace, size_t workspace_bytes,
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

  const int R = B * L;
  // ---- 16-bit split path (default): every GEMM of the layer on tcgen05 with fp16 hi/lo operand planes, persistent
  //      weight-resident linear kernels (lin_h.cu) and the warp-specialised attention (attention_h.cu) ----
  if (attention_mode() == 0 && attention_tc_supported(hd) && linear_h_eligible(H, 3 * H) && linear_h_eligible(H, H) &&
      linear_h_eligible(H, F) && linear_h_eligible(F, H) && (hd % 16 == 0)) {
    const float* srcs[4] = {w->qkv_w, w->out_w, w->ffn1_w, w->ffn2_w};
    void* wpl[4] = {ws.w_planes[0], ws.w_planes[1], ws.w_planes[2], ws.w_planes[3]};
    const long long ns[4] = {(long long)3 * H * H, (long long)H * H, (long long)F * H, (long long)H * F};
    if ((rc = launch_w_split_h(srcs, wpl, ns, 4, s))) return rc;
    if ((rc = launch_ln_split_h(x_in, w->norm1_w, w->norm1_b, ws.xn, R, H, ln_eps, s))) return rc;
    const int Lp = (L + 7) & ~7;
    {  // attention operand planes = split(LN1(x) Wqkv^T)
      LinHParams q{};
      q.R = R; q.K = H; q.N = 3 * H; q.mode = 3; q.qkvh = ws.q; q.plane_stride = (long long)B * H * Lp;
      q.L = L; q.nh = num_heads; q.hd = hd; q.Lp = Lp; q.qscale = (float)((1.0 / sqrt((double)hd)) * 1.4426950408889634);
      if ((rc = launch_linear_h(ws.xn, wpl[0], q, M2TTS_STAGE_LN_QKV, s))) return rc;
    }
    if ((rc = launch_attention_h(ws.q, nullptr, lengths, B, L, Lp, num_heads, hd, s, nullptr, ws.ctx))) return rc;
    {  // x1 = x + ctx Wo^T + bo
      LinHParams q{};
      q.R = R; q.K = H; q.N = H; q.mode = 0; q.bias = w->out_b; q.residual = x_in; q.ldr = H; q.y = ws.x1; q.ldy = H;
      if ((rc = launch_linear_h(ws.ctx, wpl[1], q, M2TTS_STAGE_OUTPROJ, s))) return rc;
    }
    if ((rc = launch_ln_split_h(ws.x1, w->norm2_w, w->norm2_b, ws.xn, R, H, ln_eps, s))) return rc;
    {  // hid = relu(LN2(x1) W1^T + b1) as planes
      LinHParams q{};
      q.R = R; q.K = H; q.N = F; q.mode = 1; q.bias = w->ffn1_b; q.relu = 1; q.y_planes = ws.hid;
      if ((rc = launch_linear_h(ws.xn, wpl[2], q, M2TTS_STAGE_FFN1, s))) return rc;
    }
    {  // y = x1 + hid W2^T + b2
      LinHParams q{};
      q.R = R; q.K = F; q.N = H; q.mode = 0; q.bias = w->ffn2_b; q.residual = ws.x1; q.ldr = H; q.y = x_out; q.ldy = H;
      if ((rc = launch_linear_h(ws.hid, wpl[3], q, M2TTS_STAGE_FFN2, s))) return rc;
    }
    return M2TTS_OK;
  }
  // ---- tensor-core path: every GEMM of the layer on tcgen05 (3xTF32), operands as hi/lo planes ----
  if (attention_mode() != 1 && attention_tc_supported(hd) && linear_tc_eligible(H, 3 * H) && linear_tc_eligible(H, H) &&
      linear_tc_eligible(H, F) && linear_tc_eligible(F, H) && (H % 16 == 0)) {
    const float* srcs[4] = {w->qkv_w, w->out_w, w->ffn1_w, w->ffn2_w};
    const long long ns[4] = {(long long)3 * H * H, (long long)H * H, (long long)F * H, (long long)H * F};
    if ((rc = launch_w_split(srcs, ws.w_planes, ns, 4, s))) return rc;
    if ((rc = launch_ln_split(x_in, w->norm1_w, w->norm1_b, ws.xn, R, H, ln_eps, s))) return rc;
    const bool half_planes = attention_mode() == 0;      // 16-bit split attention (default): fp16 hi/lo planes
    const int Lp = half_planes ? ((L + 7) & ~7) : ws.Lp;
    {  // attention operand planes = split(LN1(x) Wqkv^T)
      LinTcArgs a{};
      a.R = R; a.L = L; a.K = H; a.N = 3 * H; a.mode = half_planes ? 3 : 2; a.qkv6 = ws.q; a.plane_stride = (long long)B * H * Lp;
      a.nh = num_heads; a.hd = hd; a.Lp = Lp; a.qscale = (float)((1.0 / sqrt((double)hd)) * 1.4426950408889634);
      if ((rc = launch_linear_tc(ws.xn, ws.w_planes[0], a, B, M2TTS_STAGE_LN_QKV, s))) return rc;
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
      if ((rc = launch_linear_tc(ws.ctx, ws.w_planes[1], a, B, M2TTS_STAGE_OUTPROJ, s))) return rc;
    }
    if ((rc = launch_ln_split(ws.x1, w->norm2_w, w->norm2_b, ws.xn, R, H, ln_eps, s))) return rc;
    {  // hid = relu(LN2(x1) W1^T + b1) as planes
      LinTcArgs a{};
      a.R = R; a.L = L; a.K = H; a.N = F; a.mode = 1; a.bias = w->ffn1_b; a.relu = 1; a.y_planes = ws.hid;
      if ((rc = launch_linear_tc(ws.xn, ws.w_planes[2], a, B, M2TTS_STAGE_FFN1, s))) return rc;
    }
    {  // y = x1 + hid W2^T + b2
      LinTcArgs a{};
      a.R = R; a.L = L; a.K = F; a.N = H; a.mode = 0; a.bias = w->ffn2_b; a.residual = ws.x1; a.ldr = H; a.y = x_out; a.ldy = H;
      if ((rc = launch_linear_tc(ws.hid, ws.w_planes[3], a, B, M2TTS_STAGE_FFN2, s))) return rc;
    }
    return M2TTS_OK;
  }

  PackJob jobs[4] = {{w->qkv_w, ws.wqkv_t, 3 * H, H}, {w->out_w, ws.wo_t, H, H},
                     {w->ffn1_w, ws.w1_t, F, H}, {w->ffn2_w, ws.w2_t, H, F}};
  if ((rc = launch_pack_transpose(jobs, 4, s))) return rc;

  const bool use_tc = attention_mode() != 1 && attention_tc_supported(hd);
  {  // q,k,v = split(LN1(x) Wqkv^T)
    RowGemmArgs a{};
    a.x = x_in; a.ldx = H; a.ln_w = w->norm1_w; a.ln_b = w->norm1_b; a.eps = ln_eps;
    a.wt = ws.wqkv_t; a.qkv_mode = use_tc ? 2 : 1; a.q = ws.q; a.k = ws.k; a.v = ws.v;
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
    a.x = ws.ctx; a.ldx = H; a.wt = ws.wo_t; a.bias = w->out_b; a.residual = x_in; a.ldr = H;
    a.y = ws.x1; a.ldy = H; a.R = R; a.K = H; a.N = H; a.stage = M2TTS_STAGE_OUTPROJ;
    if ((rc = launch_rowgemm(a, s))) return rc;
  }
  {  // hid = relu(LN2(x1) W1^T + b1)
    RowGemmArgs a{};
    a.x = ws.x1; a.ldx = H; a.ln_w = w->norm2_w; a.ln_b = w->norm2_b; a.eps = ln_eps;
    a.wt = ws.w1_t; a.bias = w->ffn1_b; a.relu = 1; a.y = ws.hid; a.ldy = F;
    a.R = R; a.K = H; a.N = F; a.stage = M2TTS_STAGE_FFN1;
    if ((rc = launch_rowgemm(a, s))) return rc;
  }
  {  // y = x1 + hid W2^T + b2
    RowGemmArgs a{};
    a.x = ws.hid; a.ldx = F; a.wt = ws.w2_t; a.bias = w->ffn2_b; a.residual = ws.x1; a.ldr = H;
    a.y = x_out; a.ldy = H; a.R = R; a.K = F; a.N = H; a.stage = M2TTS_STAGE_FFN2;
    if ((rc = launch_rowgemm(a, s))) return rc;
  }
  return M2TTS_OK;
}

extern "C" size_t m2tts_ln_proj_workspace_bytes(int H, int N) {
  if (H <= 0 || N <= 0) return 0;
  return align_up((size_t)3 * H * N * sizeof(float), 256) + 1024;   // W^T (FFMA) or W hi/lo planes (tensor cores)
}

extern "C" size_t m2tts_ln_proj_rows_workspace_bytes(int rows, int H, int N) {
  if (rows <= 0 || H <= 0 || N <= 0) return 0;
  return m2tts_ln_proj_workspace_bytes(H, N) + align_up((size_t)2 * rows * H * sizeof(float), 256) + 256;
}

extern "C" int m2tts_layernorm_proj(const float* x, const float* ln_w, const float* ln_b, const float* W,
                                    const float* bias, float* y, int rows, int H, int N, float eps,
                                    void* workspace, size_t workspace_bytes, m2tts_stream_t stream) {
  M2_REQUIRE(x && ln_w && ln_b && W && y && workspace, M2TTS_E_NULLPTR, "layernorm_proj: null pointer");
  M2_REQUIRE(rows > 0 && H > 0 && N > 0, M2TTS_E_BADSHAPE, "layernorm_proj: rows=%d H=%d N=%d", rows, H, N);
  Carver cv(workspace, workspace_bytes);
  float* wt = cv.take<float>((size_t)3 * H * N);
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "layernorm_proj: workspace too small or not 256-B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  {  // tensor-core path when the caller's workspace also has room for the normalised rows as hi/lo planes
    float* xn = cv.take<float>((size_t)2 * rows * H);
    if (attention_mode() == 0 && cv.ok() && linear_h_eligible(H, N)) {   // 16-bit split (default)
      const float* srcs[1] = {W};
      void* dsts[1] = {wt};
      const long long ns[1] = {(long long)N * H};
      if ((rc = launch_w_split_h(srcs, dsts, ns, 1, s))) return rc;
      if ((rc = launch_ln_split_h(x, ln_w, ln_b, xn, rows, H, eps, s))) return rc;
      LinHParams q{};
      q.R = rows; q.K = H; q.N = N; q.mode = 0; q.bias = bias; q.y = y; q.ldy = N;
      return launch_linear_h(xn, wt, q, M2TTS_STAGE_LN_PROJ, s);
    }
    if (attention_mode() != 1 && cv.ok() && linear_tc_eligible(H, N) && (N % 16 == 0) && (H % 4 == 0)) {
      const float* srcs[1] = {W};
      float* dsts[1] = {wt};
      const long long ns[1] = {(long long)N * H};
      if ((rc = launch_w_split(srcs, dsts, ns, 1, s))) return rc;
      if ((rc = launch_ln_split(x, ln_w, ln_b, xn, rows, H, eps, s))) return rc;
      LinTcArgs a{};
      a.R = rows; a.L = rows; a.K = H; a.N = N; a.mode = 0; a.bias = bias; a.y = y; a.ldy = N;
      return launch_linear_tc(xn, wt, a, 1, M2TTS_STAGE_LN_PROJ, s);
    }
  }
  PackJob job{W, wt, N, H};
  if ((rc = launch_pack_transpose(&job, 1, s))) return rc;
  RowGemmArgs a{};
  a.x = x; a.ldx = H; a.ln_w = ln_w; a.ln_b = ln_b; a.eps = eps; a.wt = wt; a.bias = bias;
  a.y = y; a.ldy = N; a.R = rows; a.K = H; a.N = N; a.stage = M2TTS_STAGE_LN_PROJ;
  return launch_rowgemm(a, s);
}
