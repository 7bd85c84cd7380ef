// durpred.cu — the duration predictor as one fused kernel.
// Reference: tts_model.py:99-117 (transpose -> VariancePredictor -> squeeze -> softplus),
// components.py:154-174 (Conv1d k=3 pad=1 -> BatchNorm1d(eval, running stats) -> ReLU),
// components.py:214-223 (two ConvBlocks then a 1x1 Conv1d to one channel).
// One CTA = one utterance x 32 phoneme positions; the two hidden activations live in shared
// memory with a +-2 / +-1 halo. Each Conv1d zero-pads ITS OWN input, so hidden activations at
// positions outside [0,S) are forced to 0 (not relu(bn(bias))).
#include "common.cuh"

namespace m2 {

constexpr int DP_TS = 32;        // positions per CTA
constexpr int DP_THREADS = 256;

// out[p][co] = relu(bn(bias[co] + sum_{ci,j} w[co,ci,j] * in[p+j][ci])) for p in [0,n_out);
// global position of out row p is g0 + p; rows outside [0,S) are zeroed.
__device__ __forceinline__ void dp_conv_bn_relu(const float* __restrict__ in, float* __restrict__ out,
                                                const float* __restrict__ w, const float* __restrict__ bias,
                                                const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                const float* __restrict__ bn_mean, const float* __restrict__ bn_var,
                                                float eps, int H, int n_out, int g0, int S) {
  const int nq = (n_out + 3) / 4;
  for (int it = threadIdx.x; it < H * nq; it += DP_THREADS) {
    const int co = it % H, p0 = (it / H) * 4;
    const float* wr = w + (long long)co * H * 3;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ci = 0; ci < H; ++ci) {
      const float w0 = __ldg(wr + ci * 3), w1 = __ldg(wr + ci * 3 + 1), w2 = __ldg(wr + ci * 3 + 2);
      float xv[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) xv[r] = in[(p0 + r) * H + ci];  // rows up to n_out+1 exist (+3 slack rows)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(xv[r], w0, fmaf(xv[r + 1], w1, fmaf(xv[r + 2], w2, acc[r])));
    }
    const float invstd = 1.0f / sqrtf(__ldg(bn_var + co) + eps);
    const float g = __ldg(bn_w + co), be = __ldg(bn_b + co), mu = __ldg(bn_mean + co), cb = __ldg(bias + co);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int p = p0 + r;
      if (p >= n_out) break;
      const int gpos = g0 + p;
      float v = ((acc[r] + cb) - mu) * invstd * g + be;
      v = fmaxf(v, 0.f);
      out[p * H + co] = (gpos >= 0 && gpos < S) ? v : 0.f;
    }
  }
}

__global__ void __launch_bounds__(DP_THREADS) durpred_kernel(m2tts_durpred_weights w, const float* __restrict__ enc,
                                                             float* __restrict__ dur, int S, int H) {
  extern __shared__ __align__(16) float smem[];
  // rows: xs TS+4 (+4 slack), h1 TS+2 (+4 slack), h2 TS
  float* xs = smem;
  float* h1 = xs + (DP_TS + 8) * H;
  float* h2 = h1 + (DP_TS + 6) * H;
  const int b = blockIdx.y, s0 = blockIdx.x * DP_TS;
  const float* eb = enc + (long long)b * S * H;

  for (int idx = threadIdx.x; idx < (DP_TS + 8) * H; idx += DP_THREADS) {
    const int p = idx / H, c = idx % H;
    const int g = s0 - 2 + p;
    xs[idx] = (p < DP_TS + 4 && g >= 0 && g < S) ? eb[(long long)g * H + c] : 0.f;
  }
  for (int idx = threadIdx.x; idx < 4 * H; idx += DP_THREADS) h1[(DP_TS + 2) * H + idx] = 0.f;  // slack rows
  __syncthreads();
  dp_conv_bn_relu(xs, h1, w.conv_w[0], w.conv_b[0], w.bn_w[0], w.bn_b[0], w.bn_mean[0], w.bn_var[0],
                  w.bn_eps, H, DP_TS + 2, s0 - 1, S);
  __syncthreads();
  dp_conv_bn_relu(h1, h2, w.conv_w[1], w.conv_b[1], w.bn_w[1], w.bn_b[1], w.bn_mean[1], w.bn_var[1],
                  w.bn_eps, H, DP_TS, s0, S);
  __syncthreads();
  // 1x1 projection to one channel + softplus (beta=1, threshold=20), one warp per position
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int p = warp; p < DP_TS; p += DP_THREADS / 32) {
    const int g = s0 + p;
    if (g >= S) break;
    float acc = 0.f;
    for (int c = lane; c < H; c += 32) acc = fmaf(h2[p * H + c], __ldg(w.proj_w + c), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float z = acc + __ldg(w.proj_b);
      dur[(long long)b * S + g] = (z > 20.f) ? z : log1pf(expf(z));
    }
  }
}

}  // namespace m2

using namespace m2;

extern "C" int m2tts_duration_predictor(const m2tts_durpred_weights* w, const float* enc, float* dur,
                                        int B, int S, int H, m2tts_stream_t stream) {
  M2_REQUIRE(w && enc && dur, M2TTS_E_NULLPTR, "duration_predictor: null pointer");
  M2_REQUIRE(B > 0 && S > 0 && H > 0 && B <= 65535, M2TTS_E_BADSHAPE, "duration_predictor: B=%d S=%d H=%d", B, S, H);
  for (int i = 0; i < 2; ++i)
    M2_REQUIRE(w->conv_w[i] && w->conv_b[i] && w->bn_w[i] && w->bn_b[i] && w->bn_mean[i] && w->bn_var[i],
               M2TTS_E_NULLPTR, "duration_predictor: null weight in conv block %d", i);
  M2_REQUIRE(w->proj_w && w->proj_b, M2TTS_E_NULLPTR, "duration_predictor: null projection");
  const size_t smem = (size_t)((DP_TS + 8) + (DP_TS + 6) + DP_TS) * H * sizeof(float);
  M2_REQUIRE(smem <= 227 * 1024, M2TTS_E_UNSUPPORTED, "duration_predictor: hidden dim %d too large", H);
  M2_CUDA_OK(allow_smem(durpred_kernel, smem));
  dim3 grid(ceil_div(S, DP_TS), B);
  M2_LAUNCH(M2TTS_STAGE_DURPRED, durpred_kernel, grid, DP_THREADS, smem, (cudaStream_t)stream, *w, enc, dur, S, H);
  return M2TTS_OK;
}
