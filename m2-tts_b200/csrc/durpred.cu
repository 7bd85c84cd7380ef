// durpred.cu — the duration predictor as one fused kernel.
// Reference: tts_model.py:99-117 (transpose -> VariancePredictor -> squeeze -> softplus),
// components.py:154-174 (Conv1d k=3 pad=1 -> BatchNorm1d(eval, running stats) -> ReLU),
// components.py:214-223 (two ConvBlocks then a 1x1 Conv1d to one channel).
//
// One CTA = one utterance x DP_TS phoneme positions; the two hidden activations live in shared memory with a
// +-2 / +-1 halo. Each Conv1d zero-pads ITS OWN input, so hidden activations at positions outside [0,S) are forced
// to 0 (not relu(bn(bias))).
//
// Round 2: the weights of the layer at hand are staged in shared memory as [ci][tap][co] (a chunk of output channels when
// the whole layer does not fit), the activations channel-major [c][position], and a thread owns a 4 (co) x 4 (position)
// register tile: per input channel 3 LDS.128 of weights + LDS.128 + LDS.64 of inputs feed 48 FFMA. The first version read
// every weight with a scalar __ldg at a stride of 3 H floats across the warp (32 cache lines per load) and did 12 FFMA per
// 9 loads: 125 us for 4 x 48 phonemes, the most expensive kernel of small calls.
#include "common.cuh"

namespace m2 {

constexpr int DP_TS = 32;        // positions per CTA
constexpr int DP_THREADS = 256;
constexpr int DP_P = DP_TS + 8;  // row pitch (positions) of the channel-major activations: multiple of 4, covers halo + slack

struct DpLayer {
  const float* w; const float* bias; const float* bn_w; const float* bn_b; const float* bn_mean; const float* bn_var;
};

// Stage w[co0 .. co0+coc)[ci][tap] -> ws[(ci*3+tap)*(coc+4) + co]. Global reads are coalesced (the source is contiguous in
// (co, ci, tap)); the transposed shared-memory writes are 4-way conflicted at worst (pitch coc + 4).
__device__ __forceinline__ void dp_stage_weights(const float* __restrict__ w, float* __restrict__ ws, int H, int co0, int coc) {
  const int n = coc * H * 3, pitch = coc + 4;
  const float* src = w + (long long)co0 * H * 3;
  for (int i = threadIdx.x; i < n; i += DP_THREADS) {
    const int co = i / (H * 3), r = i - co * (H * 3);
    ws[r * pitch + co] = __ldg(src + i);
  }
}

// out[co][p] = relu(bn(bias[co] + sum_{ci,j} w[co,ci,j] * in[ci][p+j])) for p in [0,n_out), channel-major rows of DP_P
// positions; global position of out column p is g0 + p; columns outside [0,S) are zeroed. in must hold columns
// [0, round_up(n_out,4) + 2). OUT_PM: write position-major out[p][co] (row pitch H) instead, for the projection.
template <bool OUT_PM>
__device__ __forceinline__ void dp_conv_bn_relu(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ ws,
                                                const DpLayer& ly, float eps, int H, int coc_max, int n_out, int g0, int S) {
  const int nq = (n_out + 3) / 4;
  for (int co0 = 0; co0 < H; co0 += coc_max) {
    const int coc = min(coc_max, H - co0), pitch = coc + 4, ng = coc / 4;
    __syncthreads();                                   // previous users of ws / producers of `in` are done
    dp_stage_weights(ly.w, ws, H, co0, coc);
    __syncthreads();
    for (int it = threadIdx.x; it < ng * nq; it += DP_THREADS) {
      const int g = it % ng, p0 = (it / ng) * 4;
      float acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[a][r] = 0.f;
      const float* wp = ws + 4 * g;
      const float* xp = in + p0;
#pragma unroll 2
      for (int ci = 0; ci < H; ++ci) {
        const float4 x03 = *reinterpret_cast<const float4*>(xp + ci * DP_P);
        const float2 x45 = *reinterpret_cast<const float2*>(xp + ci * DP_P + 4);
        const float xv[6] = {x03.x, x03.y, x03.z, x03.w, x45.x, x45.y};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + (ci * 3 + j) * pitch);
          const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[a][r] = fmaf(xv[r + j], wv[a], acc[a][r]);
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int co = co0 + 4 * g + a;
        const float invstd = 1.0f / sqrtf(__ldg(ly.bn_var + co) + eps);
        const float gm = __ldg(ly.bn_w + co), be = __ldg(ly.bn_b + co), mu = __ldg(ly.bn_mean + co), cb = __ldg(ly.bias + co);
        float v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int gpos = g0 + p0 + r;
          const float t = fmaxf(((acc[a][r] + cb) - mu) * invstd * gm + be, 0.f);
          v[r] = (p0 + r < n_out && gpos >= 0 && gpos < S) ? t : 0.f;
        }
        if (OUT_PM) {
#pragma unroll
          for (int r = 0; r < 4; ++r) out[(p0 + r) * H + co] = v[r];
        } else {
          *reinterpret_cast<float4*>(out + co * DP_P + p0) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(DP_THREADS) durpred_kernel(m2tts_durpred_weights w, const float* __restrict__ enc,
                                                             float* __restrict__ dur, int S, int H, int coc_max) {
  extern __shared__ __align__(16) float smem[];
  // channel-major activations [H][DP_P]: xs columns = positions s0-2 .. , h1 columns = positions s0-1 .. ; h2 is
  // position-major [DP_TS][H] and reuses xs (dead after layer 1); then the weight chunk
  float* xs = smem;
  float* h1 = xs + H * DP_P;
  float* ws = h1 + H * DP_P;
  float* h2 = xs;
  const int b = blockIdx.y, s0 = blockIdx.x * DP_TS;
  const float* eb = enc + (long long)b * S * H;

  // enc rows [s0-2, s0+DP_TS+6) -> xs[c][p] (zero outside [0,S)): coalesced reads along c, transposed writes
  for (int idx = threadIdx.x; idx < DP_P * H; idx += DP_THREADS) {
    const int p = idx / H, c = idx - p * H;
    const int g = s0 - 2 + p;
    xs[c * DP_P + p] = (g >= 0 && g < S) ? __ldg(eb + (long long)g * H + c) : 0.f;
  }
  for (int idx = threadIdx.x; idx < DP_P * H; idx += DP_THREADS) h1[idx] = 0.f;      // columns beyond n_out stay zero
  const DpLayer l0{w.conv_w[0], w.conv_b[0], w.bn_w[0], w.bn_b[0], w.bn_mean[0], w.bn_var[0]};
  const DpLayer l1{w.conv_w[1], w.conv_b[1], w.bn_w[1], w.bn_b[1], w.bn_mean[1], w.bn_var[1]};
  dp_conv_bn_relu<false>(xs, h1, ws, l0, w.bn_eps, H, coc_max, DP_TS + 2, s0 - 1, S);
  dp_conv_bn_relu<true>(h1, h2, ws, l1, w.bn_eps, H, coc_max, DP_TS, s0, S);
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
  M2_REQUIRE(H % 4 == 0, M2TTS_E_UNSUPPORTED, "duration_predictor: hidden dim %d must be a multiple of 4", H);
  for (int i = 0; i < 2; ++i)
    M2_REQUIRE(w->conv_w[i] && w->conv_b[i] && w->bn_w[i] && w->bn_b[i] && w->bn_mean[i] && w->bn_var[i],
               M2TTS_E_NULLPTR, "duration_predictor: null weight in conv block %d", i);
  M2_REQUIRE(w->proj_w && w->proj_b, M2TTS_E_NULLPTR, "duration_predictor: null projection");
  // shared memory: two channel-major activation tiles + one weight chunk [H*3][coc+4]; the whole layer when it fits
  const size_t act = (size_t)2 * H * DP_P * sizeof(float);
  M2_REQUIRE(act + (size_t)H * 3 * 8 * sizeof(float) <= 227 * 1024, M2TTS_E_UNSUPPORTED, "duration_predictor: hidden dim %d too large", H);
  int coc = H;
  // <= 112 KB keeps two CTAs per SM (their staging and compute phases overlap)
  while (coc > 4 && act + (size_t)H * 3 * (coc + 4) * sizeof(float) > 112 * 1024) coc = ((coc / 2) + 3) & ~3;
  const size_t smem = act + (size_t)H * 3 * (coc + 4) * sizeof(float);
  M2_CUDA_OK(allow_smem(durpred_kernel, smem));
  dim3 grid(ceil_div(S, DP_TS), B);
  M2_LAUNCH(M2TTS_STAGE_DURPRED, durpred_kernel, grid, DP_THREADS, smem, (cudaStream_t)stream, *w, enc, dur, S, H, coc);
  return M2TTS_OK;
}
