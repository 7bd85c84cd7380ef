// length_regulator.cu — integer scan + gather replacement for the reference's Python double loop
// (tts_model.py:126-178): `int(durations[b,s].item())` truncates toward zero, only counts > 0 are
// expanded (:150-151), an utterance whose counts are all zero becomes ONE zero row (:158-160),
// the batch is zero-padded / truncated to max_length (:165-176).
// Pass 1 (count): per utterance, n = trunc(d) clamped at 0, inclusive int32 prefix sum, frame
// total, batch max. Pass 2 (gather): output row j copies encoder row s(j) = first s with
// cum[s] > j (binary search in shared memory), rows past the total are zero. Bit-exact: the
// data path is integer index arithmetic plus fp32 copies.
#include "common.cuh"
#include <limits.h>

namespace m2 {

constexpr int LR_THREADS = 256;

__global__ void __launch_bounds__(LR_THREADS) lr_count_kernel(const float* __restrict__ dur, int S,
                                                              int32_t* __restrict__ cum,
                                                              int32_t* __restrict__ frames,
                                                              int32_t* __restrict__ t_max,
                                                              int32_t* __restrict__ status) {
  __shared__ long long warp_tot[LR_THREADS / 32];
  __shared__ long long carry_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* d = dur + (long long)b * S;
  int32_t* c = cum + (long long)b * S;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  int st = 0;
  for (int base = 0; base < S; base += LR_THREADS) {
    const int s = base + tid;
    long long n = 0;
    if (s < S) {
      const float v = d[s];
      if (isnan(v)) st |= 1;
      else if (isinf(v)) st |= 2;
      else if (v >= 2147483648.0f) { st |= 4; n = INT_MAX; }
      else if (v >= 1.0f) n = (long long)v;  // C cast == Python int(): truncation toward zero
    }
    // inclusive scan inside the warp, then across warps
    long long x = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    long long pre = carry_s;
    for (int w = 0; w < warp; ++w) pre += warp_tot[w];
    const long long incl = pre + x;
    if (s < S) {
      if (incl > INT_MAX) { st |= 4; c[s] = INT_MAX; }
      else c[s] = (int32_t)incl;
    }
    __syncthreads();
    if (tid == LR_THREADS - 1) carry_s = incl;
    __syncthreads();
  }
  if (tid == 0) {
    long long tot = carry_s;
    if (tot > INT_MAX) { tot = INT_MAX; st |= 4; }
    frames[b] = (int32_t)tot;
    atomicMax(t_max, (int32_t)(tot < 1 ? 1 : tot));
  }
  if (st) atomicOr(status, st);
}

constexpr int LR_ROWS = 64;  // output rows per CTA

__global__ void __launch_bounds__(LR_THREADS) lr_gather_kernel(const float* __restrict__ enc,
                                                               const int32_t* __restrict__ cum,
                                                               const int32_t* __restrict__ frames,
                                                               float* __restrict__ out,
                                                               int32_t* __restrict__ index, int S, int H, int T) {
  extern __shared__ int32_t cs[];  // cum[b, :]
  const int b = blockIdx.y, j0 = blockIdx.x * LR_ROWS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t* cb = cum + (long long)b * S;
  for (int s = tid; s < S; s += LR_THREADS) cs[s] = cb[s];
  __syncthreads();
  const int nfr = frames[b];
  const float* eb = enc + (long long)b * S * H;
  const bool vec = ((H & 3) == 0) && ((((uintptr_t)enc) & 15) == 0) && ((((uintptr_t)out) & 15) == 0);
  for (int r = warp; r < LR_ROWS; r += LR_THREADS / 32) {
    const int j = j0 + r;
    if (j >= T) break;
    int src = -1;
    if (j < nfr) {
      int lo = 0, hi = S - 1;  // first s with cs[s] > j; exists because cs[S-1] == nfr > j
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cs[mid] > j) hi = mid; else lo = mid + 1;
      }
      src = lo;
    }
    float* orow = out + ((long long)b * T + j) * H;
    if (vec) {
      const float4* er = reinterpret_cast<const float4*>(eb + (long long)(src < 0 ? 0 : src) * H);
      float4* o4 = reinterpret_cast<float4*>(orow);
      for (int c = lane; c < H / 4; c += 32) o4[c] = (src >= 0) ? er[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const float* er = eb + (long long)(src < 0 ? 0 : src) * H;
      for (int c = lane; c < H; c += 32) orow[c] = (src >= 0) ? er[c] : 0.f;
    }
    if (index != nullptr && lane == 0) index[(long long)b * T + j] = src;
  }
}

}  // namespace m2

using namespace m2;

extern "C" int m2tts_length_regulate_count(const float* dur, int B, int S, int32_t* cum, int32_t* frames,
                                           int32_t* t_max, int32_t* status, m2tts_stream_t stream) {
  M2_REQUIRE(dur && cum && frames && t_max && status, M2TTS_E_NULLPTR, "length_regulate_count: null pointer");
  M2_REQUIRE(B > 0 && S > 0, M2TTS_E_BADSHAPE, "length_regulate_count: B=%d S=%d", B, S);
  cudaStream_t s = (cudaStream_t)stream;
  M2_CUDA_OK(cudaMemsetAsync(t_max, 0, sizeof(int32_t), s));
  M2_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
  M2_LAUNCH(M2TTS_STAGE_LR_COUNT, lr_count_kernel, B, LR_THREADS, 0, s, dur, S, cum, frames, t_max, status);
  return M2TTS_OK;
}

extern "C" int m2tts_length_regulate_gather(const float* enc, const int32_t* cum, const int32_t* frames,
                                            float* out, int32_t* index, int B, int S, int H, int T,
                                            m2tts_stream_t stream) {
  M2_REQUIRE(enc && cum && frames && out, M2TTS_E_NULLPTR, "length_regulate_gather: null pointer");
  M2_REQUIRE(B > 0 && S > 0 && H > 0 && T > 0 && B <= 65535, M2TTS_E_BADSHAPE,
             "length_regulate_gather: B=%d S=%d H=%d T=%d", B, S, H, T);
  const size_t smem = (size_t)S * sizeof(int32_t);
  M2_REQUIRE(smem <= 200 * 1024, M2TTS_E_UNSUPPORTED, "length_regulate_gather: S=%d too long", S);
  M2_CUDA_OK(allow_smem(lr_gather_kernel, smem));
  dim3 grid(ceil_div(T, LR_ROWS), B);
  M2_LAUNCH(M2TTS_STAGE_LR_GATHER, lr_gather_kernel, grid, LR_THREADS, smem, (cudaStream_t)stream, enc, cum,
            frames, out, index, S, H, T);
  return M2TTS_OK;
}
