// elementwise.cu — bandwidth-bound helpers: embedding + positional encoding (+ padding
// mask), stand-alone LayerNorm, weight transposition ("pack") for the row GEMMs.
#include "common.cuh"

namespace m2 {

// x[b,s,h] = emb[id,h] * sqrt(H) + pe[s,h]   (reference: tts_model.py:78-80, components.py:39)
// mask[b,s] = s < lengths[b]                 (reference: components.py:238-240)
__global__ void embed_posenc_kernel(const int64_t* __restrict__ ids, const float* __restrict__ emb,
                                    const float* __restrict__ pe, const int64_t* __restrict__ lengths,
                                    float* __restrict__ x, uint8_t* __restrict__ mask,
                                    int B, int S, int H, int vocab, float scale, int32_t* __restrict__ status) {
  const long long total = (long long)B * S * H;
  bool bad_id = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(i % H);
    const long long row = i / H;
    const int s = (int)(row % S);
    const int b = (int)(row / S);
    long long id = ids[row];
    if (id < 0 || id >= vocab) {      // nn.Embedding raises IndexError (tts_model.py:78): report, and stay inside the table
      bad_id = true;
      id = id < 0 ? 0 : vocab - 1;
    }
    x[i] = __fadd_rn(__fmul_rn(emb[id * H + h], scale), pe[(long long)s * H + h]);  // mul then add, as torch
    if (h == 0 && mask != nullptr && lengths != nullptr) mask[row] = (uint8_t)((long long)s < lengths[b]);
  }
  if (bad_id && status != nullptr) atomicOr(status, (int32_t)M2TTS_ST_BAD_ID);
}

// One warp per row; biased variance, eps inside the sqrt (nn.LayerNorm).
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ b, float* __restrict__ y,
                                 int rows, int H, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (long long)warp * H;
  float s = 0.f;
  for (int h = lane; h < H; h += 32) s += xr[h];
  const float mean = warp_sum(s) / (float)H;
  float v = 0.f;
  for (int h = lane; h < H; h += 32) { const float d = xr[h] - mean; v += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)H + eps);
  float* yr = y + (long long)warp * H;
  for (int h = lane; h < H; h += 32) yr[h] = (xr[h] - mean) * rstd * w[h] + b[h];
}

// dst[k*N + n] = src[n*K + k]; 32x32 tiles through shared memory; blockIdx.z = job.
struct PackJobs { PackJob j[8]; };
__global__ void pack_transpose_kernel(PackJobs jobs) {
  __shared__ float tile[32][33];
  const PackJob jb = jobs.j[blockIdx.z];
  const int tiles_k = (jb.K + 31) / 32, tiles_n = (jb.N + 31) / 32;
  for (int t = blockIdx.x; t < tiles_k * tiles_n; t += gridDim.x) {
    const int k0 = (t % tiles_k) * 32, n0 = (t / tiles_k) * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int n = n0 + r, k = k0 + threadIdx.x;
      tile[r][threadIdx.x] = (n < jb.N && k < jb.K) ? jb.src[(long long)n * jb.K + k] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int k = k0 + r, n = n0 + threadIdx.x;
      if (k < jb.K && n < jb.N) jb.dst[(long long)k * jb.N + n] = tile[threadIdx.x][r];
    }
    __syncthreads();
  }
}

int launch_pack_transpose(const PackJob* jobs, int n_jobs, cudaStream_t s) {
  M2_REQUIRE(n_jobs >= 1 && n_jobs <= 8, M2TTS_E_BADSHAPE, "pack_transpose: 1..8 jobs");
  PackJobs pj;
  int max_tiles = 1;
  for (int i = 0; i < n_jobs; ++i) {
    pj.j[i] = jobs[i];
    int t = ceil_div(jobs[i].K, 32) * ceil_div(jobs[i].N, 32);
    if (t > max_tiles) max_tiles = t;
  }
  dim3 grid(max_tiles < 64 ? max_tiles : 64, 1, n_jobs), block(32, 8);
  M2_LAUNCH(M2TTS_STAGE_PACK, pack_transpose_kernel, grid, block, 0, s, pj);
  return M2TTS_OK;
}

// float waveform -> PCM16 (clip to [-1, 1], scale by 32767, round half to even like numpy.round): the wav writer's
// conversion done where the samples already are, so only 2 bytes per sample cross PCIe (utils/audio.py).
__global__ void pcm16_kernel(const float* __restrict__ x, int16_t* __restrict__ y, long long n) {
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    short4 o;
    o.x = (short)__float2int_rn(fminf(fmaxf(v.x, -1.f), 1.f) * 32767.f);
    o.y = (short)__float2int_rn(fminf(fmaxf(v.y, -1.f), 1.f) * 32767.f);
    o.z = (short)__float2int_rn(fminf(fmaxf(v.z, -1.f), 1.f) * 32767.f);
    o.w = (short)__float2int_rn(fminf(fmaxf(v.w, -1.f), 1.f) * 32767.f);
    reinterpret_cast<short4*>(y)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    y[i] = (int16_t)__float2int_rn(fminf(fmaxf(x[i], -1.f), 1.f) * 32767.f);
  }
}

}  // namespace m2

using namespace m2;

extern "C" int m2tts_pcm16(const float* audio, int16_t* pcm, long long n, m2tts_stream_t stream) {
  M2_REQUIRE(audio && pcm, M2TTS_E_NULLPTR, "pcm16: null pointer");
  M2_REQUIRE(n > 0, M2TTS_E_BADSHAPE, "pcm16: n=%lld", n);
  M2_REQUIRE((((uintptr_t)audio) & 15) == 0 && (((uintptr_t)pcm) & 7) == 0, M2TTS_E_BADSHAPE, "pcm16: misaligned pointers");
  long long blocks = ((n >> 2) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  M2_LAUNCH(M2TTS_STAGE_EMBED, pcm16_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, audio, pcm, n);
  return M2TTS_OK;
}

extern "C" int m2tts_embed_posenc(const int64_t* ids, const float* emb, const float* pe,
                                  const int64_t* lengths, float* x, uint8_t* mask, int B, int S,
                                  int H, int vocab, int32_t* status, m2tts_stream_t stream) {
  M2_REQUIRE(ids && emb && pe && x, M2TTS_E_NULLPTR, "embed_posenc: null pointer");
  M2_REQUIRE(B > 0 && S > 0 && H > 0 && vocab > 0, M2TTS_E_BADSHAPE,
             "embed_posenc: B=%d S=%d H=%d vocab=%d must be positive", B, S, H, vocab);
  const long long total = (long long)B * S * H;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  M2_LAUNCH(M2TTS_STAGE_EMBED, embed_posenc_kernel, grid, 256, 0, (cudaStream_t)stream, ids, emb, pe,
            lengths, x, mask, B, S, H, vocab, (float)sqrt((double)H), status);
  return M2TTS_OK;
}

extern "C" int m2tts_layernorm(const float* x, const float* w, const float* b, float* y, int rows,
                               int H, float eps, m2tts_stream_t stream) {
  M2_REQUIRE(x && w && b && y, M2TTS_E_NULLPTR, "layernorm: null pointer");
  M2_REQUIRE(rows > 0 && H > 0, M2TTS_E_BADSHAPE, "layernorm: rows=%d H=%d", rows, H);
  const int warps_per_block = 8;
  M2_LAUNCH(M2TTS_STAGE_LAYERNORM, layernorm_kernel, ceil_div(rows, warps_per_block),
            warps_per_block * 32, 0, (cudaStream_t)stream, x, w, b, y, rows, H, eps);
  return M2TTS_OK;
}
