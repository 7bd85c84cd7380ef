// vocoder.cu — fp32 kernels for SimpleVocoder (tts_model.py:231-297), channel-first [B,C,L]:
//   conv3      : Conv1d(k=3, dilation d, zero "same" padding) + bias (+LeakyReLU(0.1) | tanh)
//                (+ residual added after the conv)             components.py:181-200, tts_model.py:246,272
//   convT      : ConvTranspose1d(k=2r, stride r, padding r/2) + bias + LeakyReLU(0.1), r in {2,4}
//                written as a polyphase filter: output sample r*q+p has exactly two taps,
//                x[q]*w[p+r/2] and x[q-1]*w[p+3r/2] (p<r/2) or x[q+1]*w[p-r/2] (p>=r/2)
//                                                             tts_model.py:255-263,291
//   conv3_co1  : the 1-channel output conv + tanh              tts_model.py:272,295
// CTA = 128 threads arranged TX (time) x TY (out-channel groups). A thread owns 8 time steps
// (two runs of 4, so each quarter-warp's LDS.128 is contiguous) x 8 output values per step and
// streams input channels through shared memory in chunks of CK: 192 FFMA per 12 smem loads.
#include "common.cuh"
#include <cuda_fp16.h>
#include <math.h>

namespace m2 {

constexpr int VC_THREADS = 128;

struct ConvArgs {
  const float* x; long long xs_b, xs_c, xs_t;  // element strides of the input
  const float* wp;        // conv3: packed [CI][3][CO]; convT: native [CI][CO][2r]
  const float* bias;
  const float* residual;  // [B,CO,L] or null
  float* y;               // [B,CO,L_out]
  int CI, CO, L, dil, act;
  int y_pitch;            // output row pitch in floats (0 = L)
};

__device__ __forceinline__ float vc_act(float v, int act) {
  if (act == 1) return v > 0.f ? v : 0.1f * v;
  if (act == 2) return tanhf(v);
  return v;
}

// ---------------------------------------------------------------------------------------------
template <int TY, bool DIL1>
__global__ void __launch_bounds__(VC_THREADS) conv3_kernel(ConvArgs a) {
  constexpr int TX = VC_THREADS / TY;
  constexpr int TT = 8 * TX;       // time steps per CTA
  constexpr int COB = 8 * TY;      // output channels per CTA
  constexpr int CK = (TY >= 4) ? 16 : 8;
  extern __shared__ __align__(16) float smem[];
  const int dil = DIL1 ? 1 : a.dil;
  const int XW = TT + 2 * dil;            // staged columns: t = t0 - dil + c
  const int XST = (XW + 3) & ~3;          // row stride (floats), keeps 16-B alignment
  float* xs = smem;                       // [CK][XST]
  float* ws = smem + CK * XST;            // [CK][3][COB]

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int t0 = blockIdx.x * TT, co0 = blockIdx.y * COB, b = blockIdx.z;
  const int L = a.L, CI = a.CI, CO = a.CO;
  const float* xb = a.x + (long long)b * a.xs_b;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;

  const bool wvec = ((CO & 3) == 0) && ((((uintptr_t)a.wp) & 15) == 0);

  for (int ci0 = 0; ci0 < CI; ci0 += CK) {
    __syncthreads();
    // ---- stage inputs (zero outside [0,L) and beyond CI) ----
    if (a.xs_t == 1) {
      for (int idx = tid; idx < CK * XW; idx += VC_THREADS) {
        const int ci = idx / XW, c = idx - ci * XW;
        const int t = t0 - dil + c;
        float v = 0.f;
        if (ci0 + ci < CI && t >= 0 && t < L) v = xb[(long long)(ci0 + ci) * a.xs_c + t];
        xs[ci * XST + c] = v;
      }
    } else {  // channel-contiguous input (the decoder's [B,T,M] mel): walk channels fastest
      for (int idx = tid; idx < CK * XW; idx += VC_THREADS) {
        const int c = idx / CK, ci = idx - c * CK;
        const int t = t0 - dil + c;
        float v = 0.f;
        if (ci0 + ci < CI && t >= 0 && t < L) v = xb[(long long)(ci0 + ci) * a.xs_c + (long long)t * a.xs_t];
        xs[ci * XST + c] = v;
      }
    }
    // ---- stage weights: ws[ci][j][co] <- wp[(ci0+ci)*3 + j][co0 + co] ----
    if (wvec) {
      constexpr int C4 = COB / 4;
      for (int idx = tid; idx < CK * 3 * C4; idx += VC_THREADS) {
        const int row = idx / C4, c = (idx - row * C4) * 4;
        const int ci = row / 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ci0 + ci < CI && co0 + c < CO)
          v = *reinterpret_cast<const float4*>(a.wp + ((long long)ci0 * 3 + row) * CO + co0 + c);
        *reinterpret_cast<float4*>(ws + row * COB + c) = v;
      }
    } else {
      for (int idx = tid; idx < CK * 3 * COB; idx += VC_THREADS) {
        const int row = idx / COB, c = idx - row * COB;
        const int ci = row / 3;
        ws[idx] = (ci0 + ci < CI && co0 + c < CO) ? a.wp[((long long)ci0 * 3 + row) * CO + co0 + c] : 0.f;
      }
    }
    __syncthreads();

#pragma unroll 2
    for (int ci = 0; ci < CK; ++ci) {
      const float* xr = xs + ci * XST;
      const float* wr = ws + ci * 3 * COB + ty * 8;
      if constexpr (DIL1) {
        // columns 4tx..4tx+5 cover t-1..t+4 of the left run; +TT/2 for the right run
        const float4 l4 = *reinterpret_cast<const float4*>(xr + 4 * tx);
        const float2 l2 = *reinterpret_cast<const float2*>(xr + 4 * tx + 4);
        const float4 r4 = *reinterpret_cast<const float4*>(xr + TT / 2 + 4 * tx);
        const float2 r2 = *reinterpret_cast<const float2*>(xr + TT / 2 + 4 * tx + 4);
        const float xl[6] = {l4.x, l4.y, l4.z, l4.w, l2.x, l2.y};
        const float xq[6] = {r4.x, r4.y, r4.z, r4.w, r2.x, r2.y};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 w0 = *reinterpret_cast<const float4*>(wr + j * COB);
          const float4 w1 = *reinterpret_cast<const float4*>(wr + j * COB + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              acc[i][c] = fmaf(xl[i + j], wv[c], acc[i][c]);
              acc[4 + i][c] = fmaf(xq[i + j], wv[c], acc[4 + i][c]);
            }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 w0 = *reinterpret_cast<const float4*>(wr + j * COB);
          const float4 w1 = *reinterpret_cast<const float4*>(wr + j * COB + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int tl = 4 * tx + (i & 3) + (i >> 2) * (TT / 2);
            const float xv = xr[tl + j * dil];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(xv, wv[c], acc[i][c]);
          }
        }
      }
    }
  }

  // ---- epilogue: bias, activation, residual, store (float4 along time when aligned) ----
  const int P = a.y_pitch ? a.y_pitch : L;
  const bool svec = ((P & 3) == 0) && ((L & 3) == 0) && ((((uintptr_t)a.y) & 15) == 0) &&
                    (a.residual == nullptr || (((uintptr_t)a.residual) & 15) == 0);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int co = co0 + ty * 8 + c;
    if (co >= CO) continue;
    const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
    const long long rowoff = ((long long)b * CO + co) * P;
    const long long resoff = ((long long)b * CO + co) * L;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = t0 + 4 * tx + h * (TT / 2);
      if (t >= L) continue;
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = vc_act(acc[h * 4 + i][c] + bv, a.act);
      if (svec && t + 3 < L) {
        if (a.residual) {
          const float4 r = *reinterpret_cast<const float4*>(a.residual + resoff + t);
          v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        *reinterpret_cast<float4*>(a.y + rowoff + t) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (t + i < L) {
            a.y[rowoff + t + i] = v[i] + (a.residual ? a.residual[resoff + t + i] : 0.f);
          }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <int R, int TY>
__global__ void __launch_bounds__(VC_THREADS) convT_kernel(ConvArgs a) {
  constexpr int TX = VC_THREADS / TY;
  constexpr int TQ = 8 * TX;        // input positions per CTA
  constexpr int COT = 8 / R;        // output channels per thread
  constexpr int COB = COT * TY;     // output channels per CTA
  constexpr int K2 = 2 * R;         // kernel taps
  constexpr int CK = (TY >= 4) ? 16 : 8;
  constexpr int XST = TQ + 8;       // columns: q = q0 - 1 + c, c in [0, TQ+2)
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                 // [CK][XST]
  float* ws = smem + CK * XST;      // [CK][COB][K2]

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int q0 = blockIdx.x * TQ, co0 = blockIdx.y * COB, b = blockIdx.z;
  const int L = a.L, CI = a.CI, CO = a.CO;
  const float* xb = a.x + (long long)b * a.xs_b;

  float acc[2][4][COT][R];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < COT; ++c)
#pragma unroll
        for (int p = 0; p < R; ++p) acc[h][i][c][p] = 0.f;

  for (int ci0 = 0; ci0 < CI; ci0 += CK) {
    __syncthreads();
    if (a.xs_t == 1) {
      for (int idx = tid; idx < CK * (TQ + 2); idx += VC_THREADS) {
        const int ci = idx / (TQ + 2), c = idx - ci * (TQ + 2);
        const int q = q0 - 1 + c;
        float v = 0.f;
        if (ci0 + ci < CI && q >= 0 && q < L) v = xb[(long long)(ci0 + ci) * a.xs_c + q];
        xs[ci * XST + c] = v;
      }
    } else {  // channel-last input (the output of a fused stage): walk channels fastest
      for (int idx = tid; idx < CK * (TQ + 2); idx += VC_THREADS) {
        const int c = idx / CK, ci = idx - c * CK;
        const int q = q0 - 1 + c;
        float v = 0.f;
        if (ci0 + ci < CI && q >= 0 && q < L) v = xb[(long long)(ci0 + ci) * a.xs_c + (long long)q * a.xs_t];
        xs[ci * XST + c] = v;
      }
    }
    for (int idx = tid; idx < CK * COB * K2; idx += VC_THREADS) {
      const int ci = idx / (COB * K2), rem = idx - ci * (COB * K2);
      const int col = rem / K2, kk = rem - col * K2;
      float v = 0.f;
      if (ci0 + ci < CI && co0 + col < CO) v = a.wp[((long long)(ci0 + ci) * CO + co0 + col) * K2 + kk];
      ws[idx] = v;
    }
    __syncthreads();

#pragma unroll 2
    for (int ci = 0; ci < CK; ++ci) {
      const float* xr = xs + ci * XST;
      float xv[2][6];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 v4 = *reinterpret_cast<const float4*>(xr + h * (TQ / 2) + 4 * tx);
        const float2 v2 = *reinterpret_cast<const float2*>(xr + h * (TQ / 2) + 4 * tx + 4);
        xv[h][0] = v4.x; xv[h][1] = v4.y; xv[h][2] = v4.z; xv[h][3] = v4.w; xv[h][4] = v2.x; xv[h][5] = v2.y;
      }
      const float* wr = ws + (ci * COB + ty * COT) * K2;
#pragma unroll
      for (int c = 0; c < COT; ++c) {
        float wv[K2];
#pragma unroll
        for (int kk = 0; kk < K2; kk += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(wr + c * K2 + kk);
          wv[kk] = t4.x; wv[kk + 1] = t4.y; wv[kk + 2] = t4.z; wv[kk + 3] = t4.w;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int p = 0; p < R; ++p) {
              // xv[h][i+1] = x[q], xv[h][i] = x[q-1], xv[h][i+2] = x[q+1]
              float s = fmaf(xv[h][i + 1], wv[p + R / 2], acc[h][i][c][p]);
              if (p < R / 2) s = fmaf(xv[h][i], wv[p + R / 2 + R], s);
              else s = fmaf(xv[h][i + 2], wv[p - R / 2], s);
              acc[h][i][c][p] = s;
            }
      }
    }
  }

  const long long Lo = (long long)R * L;
  const bool svec = ((Lo & 3) == 0) && ((((uintptr_t)a.y) & 15) == 0);
#pragma unroll
  for (int c = 0; c < COT; ++c) {
    const int co = co0 + ty * COT + c;
    if (co >= CO) continue;
    const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
    float* yr = a.y + ((long long)b * CO + co) * Lo;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = q0 + h * (TQ / 2) + 4 * tx;
      if (q >= L) continue;
      float v[4 * R];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int p = 0; p < R; ++p) {
          const float t = acc[h][i][c][p] + bv;
          v[i * R + p] = t > 0.f ? t : 0.1f * t;
        }
      if (svec && q + 3 < L) {
#pragma unroll
        for (int e = 0; e < 4 * R; e += 4)
          *reinterpret_cast<float4*>(yr + (long long)R * q + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4 * R; ++e)
          if (q + e / R < L) yr[(long long)R * q + e] = v[e];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// y[b,0,t] = tanh(bias + sum_{ci,j} w[0,ci,j] * x[b,ci,t+j-1]); 4 outputs per thread, x read
// straight from global (contiguous [B,CI,L]); purely bandwidth-bound.
__global__ void __launch_bounds__(256) conv3_co1_tanh_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ bias,
                                                             float* __restrict__ y, int CI, int L) {
  extern __shared__ float wsm[];  // [CI][3]
  for (int i = threadIdx.x; i < CI * 3; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (t >= L) return;
  const float* xb = x + (long long)b * CI * L;
  const bool vec = ((L & 3) == 0) && ((((uintptr_t)x) & 15) == 0) && ((((uintptr_t)y) & 15) == 0) && (t + 3 < L);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ci = 0; ci < CI; ++ci) {
    const float* xr = xb + (long long)ci * L;
    float xv[6];
    if (vec) {
      const float4 c4 = *reinterpret_cast<const float4*>(xr + t);
      xv[1] = c4.x; xv[2] = c4.y; xv[3] = c4.z; xv[4] = c4.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[1 + i] = (t + i < L) ? xr[t + i] : 0.f;
    }
    xv[0] = (t > 0) ? xr[t - 1] : 0.f;
    xv[5] = (t + 4 < L) ? xr[t + 4] : 0.f;
    const float w0 = wsm[ci * 3], w1 = wsm[ci * 3 + 1], w2 = wsm[ci * 3 + 2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = fmaf(xv[i], w0, fmaf(xv[i + 1], w1, fmaf(xv[i + 2], w2, acc[i])));
  }
  const float bv = bias ? bias[0] : 0.f;
  float* yb = y + (long long)b * L;
  if (vec) {
    *reinterpret_cast<float4*>(yb + t) = make_float4(tanhf(acc[0] + bv), tanhf(acc[1] + bv), tanhf(acc[2] + bv), tanhf(acc[3] + bv));
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (t + i < L) yb[t + i] = tanhf(acc[i] + bv);
  }
}

// mel element (b,m,t) at x[b*sb + m*sm + t*st]  ->  y[b][m][t] with row pitch Lp (zero tail), through a 32x32 tile so
// both sides are coalesced whichever of m / t is contiguous in x. Feeds the tensor-core input convolution, whose TMA
// needs channel-first rows with a 16-byte pitch (the decoder hands the vocoder a [B,T,M] tensor, tts_model.py:390).
__global__ void __launch_bounds__(256) mel_to_channel_first_kernel(const float* __restrict__ x, long long sb, long long sm, long long st,
                                                                    float* __restrict__ y, int M, int T, int Lp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  const float* xb = x + (long long)b * sb;
  if (st == 1) {          // t contiguous: read rows of t
    for (int r = ty; r < 32; r += 8) {
      const int m = m0 + r, t = t0 + tx;
      tile[r][tx] = (m < M && t < T) ? xb[(long long)m * sm + t] : 0.f;
    }
  } else {                // m contiguous (or generic): read rows of m
    for (int r = ty; r < 32; r += 8) {
      const int t = t0 + r, m = m0 + tx;
      tile[tx][r] = (m < M && t < T) ? xb[(long long)m * sm + (long long)t * st] : 0.f;
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int m = m0 + r, t = t0 + tx;
    if (m < M && t < Lp) y[((long long)b * M + m) * Lp + t] = tile[r][tx];
  }
}

// x [B, M, T] fp32 with element strides -> fp16 hi/lo planes channel-last [2][B][T][M] (the input of the channel-last
// 16-bit split input convolution); a 32 x 32 tile goes through shared memory so that both sides are coalesced whichever
// of m / t is contiguous in x.
__global__ void __launch_bounds__(256) mel_to_planes_kernel(const float* __restrict__ x, long long sb, long long sm, long long st,
                                                             __half* __restrict__ planes, long long plane, int M, int T,
                                                             int32_t* __restrict__ status) {
  __shared__ float tile[32][33];      // [t][m]
  bool bad = false;
  const int b = blockIdx.z, t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  const float* xb = x + (long long)b * sb;
  if (st == 1) {          // t contiguous: read rows of t
    for (int r = ty; r < 32; r += 8) {
      const int m = m0 + r, t = t0 + tx;
      tile[tx][r] = (m < M && t < T) ? xb[(long long)m * sm + t] : 0.f;
    }
  } else {                // m contiguous (or generic): read rows of m
    for (int r = ty; r < 32; r += 8) {
      const int t = t0 + r, m = m0 + tx;
      tile[r][tx] = (m < M && t < T) ? xb[(long long)m * sm + (long long)t * st] : 0.f;
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, m = m0 + tx;
    if (m < M && t < T) {
      const float v = tile[r][tx];
      h_chk(v, bad);
      const __half h = __float2half_rn(v);
      const long long o = ((long long)b * T + t) * M + m;
      planes[o] = h;
      planes[plane + o] = __float2half_rn(v - __half2float(h));
    }
  }
  h_flag(bad, status);
}

// w[CO][CI][3] -> wp[CI][3][CO], several convolutions per launch (blockIdx.y = job)
struct ConvPackJob { const float* src; float* dst; int CO, CI; };
struct ConvPackJobs { ConvPackJob j[12]; };
__global__ void conv_pack_kernel(ConvPackJobs jobs) {
  const ConvPackJob jb = jobs.j[blockIdx.y];
  const int total = jb.CO * jb.CI * 3;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gridDim.x * blockDim.x) {
    const int co = o % jb.CO, row = o / jb.CO;  // row = ci*3 + j
    jb.dst[o] = jb.src[(long long)co * jb.CI * 3 + row];
  }
}

static int launch_conv_pack(const ConvPackJob* jobs, int n, cudaStream_t s) {
  ConvPackJobs pj;
  int mx = 1;
  for (int i = 0; i < n; ++i) { pj.j[i] = jobs[i]; const int t = jobs[i].CO * jobs[i].CI * 3; if (t > mx) mx = t; }
  dim3 grid(ceil_div(mx, 256) < 128 ? ceil_div(mx, 256) : 128, n);
  M2_LAUNCH(M2TTS_STAGE_PACK, conv_pack_kernel, grid, 256, 0, s, pj);
  return M2TTS_OK;
}

// ---- launchers ------------------------------------------------------------------------------
template <int TY, bool DIL1>
static int launch_conv3_t(const ConvArgs& a, int B, int stage, cudaStream_t s) {
  constexpr int TX = VC_THREADS / TY, TT = 8 * TX, COB = 8 * TY, CK = (TY >= 4) ? 16 : 8;
  const int dil = DIL1 ? 1 : a.dil;
  const int XST = (TT + 2 * dil + 3) & ~3;
  const size_t smem = (size_t)(CK * XST + CK * 3 * COB) * sizeof(float);
  M2_REQUIRE(smem <= 227 * 1024, M2TTS_E_UNSUPPORTED, "conv1d_k3: dilation %d too large", dil);
  M2_CUDA_OK(allow_smem(conv3_kernel<TY, DIL1>, smem));
  dim3 grid(ceil_div(a.L, TT), ceil_div(a.CO, COB), B);
  M2_LAUNCH(stage, (conv3_kernel<TY, DIL1>), grid, VC_THREADS, smem, s, a);
  return M2TTS_OK;
}

static int launch_conv3(const ConvArgs& a, int B, int stage, cudaStream_t s) {
  M2_REQUIRE(B > 0 && B <= 65535 && a.CI > 0 && a.CO > 0 && a.L > 0 && a.dil >= 1, M2TTS_E_BADSHAPE,
             "conv1d_k3: B=%d CI=%d CO=%d L=%d dil=%d", B, a.CI, a.CO, a.L, a.dil);
  const int ty = a.CO > 32 ? 8 : (a.CO > 16 ? 4 : (a.CO > 8 ? 2 : 1));
  if (a.dil == 1) {
    switch (ty) {
      case 8: return launch_conv3_t<8, true>(a, B, stage, s);
      case 4: return launch_conv3_t<4, true>(a, B, stage, s);
      case 2: return launch_conv3_t<2, true>(a, B, stage, s);
      default: return launch_conv3_t<1, true>(a, B, stage, s);
    }
  }
  switch (ty) {
    case 8: return launch_conv3_t<8, false>(a, B, stage, s);
    case 4: return launch_conv3_t<4, false>(a, B, stage, s);
    case 2: return launch_conv3_t<2, false>(a, B, stage, s);
    default: return launch_conv3_t<1, false>(a, B, stage, s);
  }
}

template <int R, int TY>
static int launch_convT_t(const ConvArgs& a, int B, cudaStream_t s) {
  constexpr int TX = VC_THREADS / TY, TQ = 8 * TX, COB = (8 / R) * TY, CK = (TY >= 4) ? 16 : 8;
  const size_t smem = (size_t)(CK * (TQ + 8) + CK * COB * 2 * R) * sizeof(float);
  M2_CUDA_OK(allow_smem(convT_kernel<R, TY>, smem));
  dim3 grid(ceil_div(a.L, TQ), ceil_div(a.CO, COB), B);
  M2_LAUNCH(M2TTS_STAGE_VOC_UP, (convT_kernel<R, TY>), grid, VC_THREADS, smem, s, a);
  return M2TTS_OK;
}

static int launch_convT(const ConvArgs& a, int B, int r, cudaStream_t s) {
  M2_REQUIRE(B > 0 && B <= 65535 && a.CI > 0 && a.CO > 0 && a.L > 0, M2TTS_E_BADSHAPE,
             "conv_transpose1d: B=%d CI=%d CO=%d L=%d", B, a.CI, a.CO, a.L);
  M2_REQUIRE(r == 2 || r == 4, M2TTS_E_UNSUPPORTED, "conv_transpose1d: stride %d unsupported (2 or 4)", r);
  if (r == 4) {  // 2 channels per thread
    if (a.CO > 8) return launch_convT_t<4, 8>(a, B, s);
    if (a.CO > 4) return launch_convT_t<4, 4>(a, B, s);
    return launch_convT_t<4, 2>(a, B, s);
  }
  if (a.CO > 16) return launch_convT_t<2, 8>(a, B, s);  // 4 channels per thread
  if (a.CO > 8) return launch_convT_t<2, 4>(a, B, s);
  return launch_convT_t<2, 2>(a, B, s);
}

}  // namespace m2

using namespace m2;

extern "C" size_t m2tts_conv_workspace_bytes(int CI, int CO, int taps) {
  if (CI <= 0 || CO <= 0 || taps <= 0) return 0;
  return align_up((size_t)CI * CO * taps * sizeof(float), 256) + 256;
}

extern "C" int m2tts_conv1d_k3(const float* x, int64_t xs_b, int64_t xs_c, int64_t xs_t, const float* w,
                               const float* bias, const float* residual, float* y, int B, int CI, int CO,
                               int L, int dilation, int act, void* workspace, size_t workspace_bytes,
                               m2tts_stream_t stream) {
  M2_REQUIRE(x && w && y && workspace, M2TTS_E_NULLPTR, "conv1d_k3: null pointer");
  M2_REQUIRE(act >= 0 && act <= 2, M2TTS_E_BADSHAPE, "conv1d_k3: act=%d", act);
  M2_REQUIRE(B > 0 && CI > 0 && CO > 0 && L > 0 && dilation >= 1, M2TTS_E_BADSHAPE,
             "conv1d_k3: B=%d CI=%d CO=%d L=%d dil=%d", B, CI, CO, L, dilation);
  Carver cv(workspace, workspace_bytes);
  float* wp = cv.take<float>((size_t)CI * CO * 3);
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "conv1d_k3: workspace too small or not 256-B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  ConvPackJob job{w, wp, CO, CI};
  int rc = launch_conv_pack(&job, 1, s);
  if (rc) return rc;
  ConvArgs a{x, xs_b, xs_c, xs_t, wp, bias, residual, y, CI, CO, L, dilation, act};
  return launch_conv3(a, B, M2TTS_STAGE_VOC_RES1, s);
}

extern "C" int m2tts_conv_transpose1d_lrelu(const float* x, const float* w, const float* bias, float* y,
                                            int B, int CI, int CO, int L, int r, m2tts_stream_t stream) {
  M2_REQUIRE(x && w && y, M2TTS_E_NULLPTR, "conv_transpose1d: null pointer");
  ConvArgs a{x, (long long)CI * L, L, 1, w, bias, nullptr, y, CI, CO, L, 1, 1};
  return launch_convT(a, B, r, (cudaStream_t)stream);
}

// ---- the vocoder as a chain of per-stage kernels ------------------------------------------------------------------------
namespace {

static const int kRates[4] = {4, 4, 2, 2};      // tts_model.py:244

// image of the input conv: the tap-GEMM layout or 96 KB per 64 output channels of the channel-last kernel, whichever is larger
size_t in_img_floats(int M, int C) {
  const size_t tc = conv3_tc_wblob_floats(M, C), hh = voc_conv_h_io_eligible(M, C) ? (size_t)C / 64 * 98304 / sizeof(float) : 0;
  return tc > hh ? tc : hh;
}
// image of a fused narrow stage (C = 8 runs zero-padded in the 16-channel kernel)
size_t fs_img_floats(int c) { return voc_fused_wblob_floats(c == 32 ? 32 : ((c == 16 || c == 8) ? 16 : 0)); }

// Per-stage kernel choice, a function of (M, C, precision, dilations) only, so that m2tts_vocoder_pack and
// m2tts_vocoder_forward agree on which weight images exist.
//   IN_H / S_UPH_* / S_FUSED_H : channel-last 16-bit split kernels (fp16 hi/lo planes between stages)
//   *_TC / S_FUSED_TF32        : TF32-split tcgen05 kernels (tap-GEMMs on channel-first fp32; fused narrow stages channel-last)
//   *_FFMA                     : fp32 FFMA kernels (any shape, any input strides)
enum { IN_FFMA = 0, IN_TC = 1, IN_H = 2 };
enum { S_FFMA = 0, S_TC = 1, S_FUSED_TF32 = 2, S_FUSED_H = 3, S_UPH_RESH = 4, S_UPH_CONVH = 5, S_TC_RESH = 6, S_TC_CONVH = 7 };

struct VocPlan {
  int in_kind;
  int kind[4];
  bool out_planes[4];      // stage j hands fp16 hi/lo planes (channel-last) to stage j + 1
  bool in_planes;          // the input conv hands planes to stage 0
};

bool stage_reads_planes(int k) { return k == S_FUSED_H || k == S_UPH_RESH || k == S_UPH_CONVH; }
bool stage_is_tc_family(int k) { return k == S_TC || k == S_UPH_RESH || k == S_UPH_CONVH || k == S_TC_RESH || k == S_TC_CONVH; }

VocPlan make_plan(int M, int C, int prec, const int* res_dilation) {
  VocPlan p{};
  p.in_kind = IN_FFMA;
  for (int j = 0; j < 4; ++j) { p.kind[j] = S_FFMA; p.out_planes[j] = false; }
  if (prec == M2TTS_PREC_FFMA) return p;
  const bool h = prec == M2TTS_PREC_SPLIT16;
  // The all-16-bit-split chain: every stage has a channel-last kernel that reads and writes fp16 hi/lo planes — transposed conv
  // (voc_up_h: 256 / 128 / 64 input channels, stride 4) + ResBlock (voc_conv_h x 2 for C = 128, voc_res_h for C = 64 / 32), or a
  // fused narrow stage (voc_stage_fused_h: C = 32 / 16, stride 2; C = 8 zero-padded to 16 as the last stage). Covers the
  // stage2_quality vocoder (C = 256) and the stage1_poc one (C = 128); anything else falls through to the mixed plan below.
  if (h && C % 64 == 0 && voc_conv_h_io_eligible(M, C) && (size_t)C / 64 * 98304 <= in_img_floats(M, C) * sizeof(float)) {
    bool all = true;
    int kinds[4];
    int ci2 = C;
    for (int j = 0; j < 4 && all; ++j, ci2 /= 2) {
      const int c = ci2 / 2;
      const int dil = res_dilation[j] > 0 ? res_dilation[j] : 1;
      if (voc_up_h_eligible(ci2, c, kRates[j]) && voc_res_h_eligible(c, dil)) kinds[j] = S_UPH_RESH;
      else if (voc_up_h_eligible(ci2, c, kRates[j]) && voc_conv_h_eligible(c, dil)) kinds[j] = S_UPH_CONVH;
      else if (voc_fused_h_eligible(c, kRates[j], dil, j == 3)) kinds[j] = S_FUSED_H;
      else all = false;
    }
    if (all) {
      p.in_kind = IN_H;
      p.in_planes = true;
      for (int j = 0; j < 4; ++j) { p.kind[j] = kinds[j]; p.out_planes[j] = j < 3; }
      return p;
    }
  }
  // pass 1: the TF32-era skeleton — wide stages as a prefix of tap-GEMM stages, narrow stages fused
  int base[4];      // 0 ffma, 1 tc, 2 fused
  int ci = C;
  for (int j = 0; j < 4; ++j, ci /= 2) {
    const int c = ci / 2;
    const int dil = res_dilation[j] > 0 ? res_dilation[j] : 1;
    const bool tc_ok = (j == 0 || base[j - 1] == 1) && convT_tc_eligible(ci, c, kRates[j]) && conv3_tc_eligible(c, c) && dil <= 4;
    const bool fused_ok = j >= 1 && base[j - 1] != 0 && voc_fused_eligible(c, kRates[j], dil);
    base[j] = fused_ok ? 2 : (tc_ok ? 1 : 0);
  }
  // pass 2: 16-bit split refinements of the tap-GEMM stages
  ci = C;
  for (int j = 0; j < 4; ++j, ci /= 2) {
    const int c = ci / 2;
    const int dil = res_dilation[j] > 0 ? res_dilation[j] : 1;
    if (base[j] == 0) { p.kind[j] = S_FFMA; continue; }
    if (base[j] == 2) { p.kind[j] = h ? S_FUSED_H : S_FUSED_TF32; continue; }
    p.kind[j] = S_TC;
    if (!h) continue;
    const bool next_fused = j + 1 < 4 && base[j + 1] == 2;
    const bool resh = voc_res_h_eligible(c, dil) && next_fused;      // C = 64: whole ResBlock in one kernel, output as planes
    const bool convh = voc_conv_h_eligible(c, dil);                  // C = 128: two channel-last conv launches
    if (!resh && !convh) continue;
    const bool producer_tc = j == 0 ? conv3_tc_eligible(M, C) : stage_is_tc_family(p.kind[j - 1]);
    const bool uph = voc_up_h_eligible(ci, c, kRates[j]) && producer_tc;
    p.kind[j] = resh ? (uph ? S_UPH_RESH : S_TC_RESH) : (uph ? S_UPH_CONVH : S_TC_CONVH);
  }
  for (int j = 0; j < 4; ++j)
    p.out_planes[j] = h && j + 1 < 4 && stage_reads_planes(p.kind[j + 1]) &&
                      (p.kind[j] == S_FUSED_H || p.kind[j] == S_UPH_RESH || p.kind[j] == S_TC_RESH || p.kind[j] == S_UPH_CONVH || p.kind[j] == S_TC_CONVH);
  p.in_planes = h && stage_reads_planes(p.kind[0]);
  if (p.kind[0] != S_FFMA && conv3_tc_eligible(M, C)) p.in_kind = IN_TC;
  if (p.in_planes && voc_conv_h_io_eligible(M, C) && (size_t)C / 64 * 98304 <= conv3_tc_wblob_floats(M, C) * sizeof(float)) p.in_kind = IN_H;
  if (p.in_planes && p.in_kind == IN_FFMA) {      // no producer of planes for stage 0: fall back to the tap-GEMM upsampler there
    p.in_planes = false;
    if (p.kind[0] == S_UPH_RESH) p.kind[0] = S_TC_RESH;
    if (p.kind[0] == S_UPH_CONVH) p.kind[0] = S_TC_CONVH;
  }
  return p;
}

// weight images of one vocoder, carved from the caller's `packed` buffer or from the workspace
struct VocBlobs {
  float* in_ffma; float* in_img;
  float* r1_ffma[4]; float* r2_ffma[4];
  float* r1_img[4]; float* r2_img[4]; float* up_img[4]; float* fs_img[4];
};

bool carve_blobs(Carver& cv, int M, int C, VocBlobs* o) {
  o->in_ffma = cv.take<float>((size_t)C * M * 3);
  o->in_img = cv.take<float>(in_img_floats(M, C));
  for (int j = 0, c = C / 2; j < 4; ++j, c /= 2) {
    o->r1_ffma[j] = cv.take<float>((size_t)c * c * 3);
    o->r2_ffma[j] = cv.take<float>((size_t)c * c * 3);
    o->r1_img[j] = cv.take<float>(conv3_tc_wblob_floats(c, c));
    o->r2_img[j] = cv.take<float>(conv3_tc_wblob_floats(c, c));
    o->up_img[j] = cv.take<float>(convT_tc_wblob_floats(2 * c, c, kRates[j]));
    o->fs_img[j] = cv.take<float>(fs_img_floats(c));
  }
  return cv.ok();
}

size_t blobs_bytes(int M, int C) {
  size_t wts = (size_t)C * M * 3 + in_img_floats(M, C);
  for (int j = 0, c = C / 2; j < 4; ++j, c /= 2)
    wts += 2 * (size_t)c * c * 3 + 2 * conv3_tc_wblob_floats(c, c) + convT_tc_wblob_floats(2 * c, c, kRates[j]) +
           fs_img_floats(c);
  return align_up(wts * sizeof(float), 256) + 40 * 256;
}

int check_weights(const m2tts_vocoder_weights* w, const char* who) {
  M2_REQUIRE(w != nullptr, M2TTS_E_NULLPTR, "%s: null weights", who);
  M2_REQUIRE(w->in_w && w->in_b && w->out_w && w->out_b, M2TTS_E_NULLPTR, "%s: null weights", who);
  for (int j = 0; j < 4; ++j)
    M2_REQUIRE(w->up_w[j] && w->up_b[j] && w->res1_w[j] && w->res1_b[j] && w->res2_w[j] && w->res2_b[j],
               M2TTS_E_NULLPTR, "%s: null weights in stage %d", who, j);
  return M2TTS_OK;
}

// run == false: only the weight images are written (m2tts_vocoder_pack); pack == false: the images are already in `bl`.
int vocoder_chain(const m2tts_vocoder_weights* w, const VocPlan& plan, const VocBlobs& bl, bool pack, bool run, const float* mel,
                  int64_t stride_b, int64_t stride_m, int64_t stride_t, float* audio, int B, int T, int M, int C, float* bufA, float* bufB,
                  float* bufC, int32_t* status, cudaStream_t s) {
  int rc;
  auto W = [&](const float* p) { return pack ? p : (const float*)nullptr; };      // weight pointer handed to a launcher (null: image is ready)
  // FFMA weight layout [CI][3][CO], only for the stages that run FFMA kernels
  if (pack) {
    ConvPackJob jobs[9];
    int n = 0;
    if (plan.in_kind == IN_FFMA) jobs[n++] = ConvPackJob{w->in_w, bl.in_ffma, C, M};
    for (int j = 0, c = C / 2; j < 4; ++j, c /= 2)
      if (plan.kind[j] == S_FFMA) {
        jobs[n++] = ConvPackJob{w->res1_w[j], bl.r1_ffma[j], c, c};
        jobs[n++] = ConvPackJob{w->res2_w[j], bl.r2_ffma[j], c, c};
      }
    if (n > 0 && (rc = launch_conv_pack(jobs, n, s))) return rc;
  }
  const float* nx = nullptr;      // "no input": pack-only call of a launcher
  // ---- input conv: mel (strided) -> bufA ----
  int L = T, c_in = C;
  int Lp = (plan.kind[0] != S_FFMA) ? ((L + 3) & ~3) : L;      // row pitch of channel-first tensors feeding TMA: multiple of 4 floats
  bool cl = false;                                            // layout of the current activation (bufA): channel-last?
  if (plan.in_kind == IN_H) {
    // channel-last 16-bit split input conv: mel -> fp16 hi/lo planes [2][B][T][M] (bufC), then the conv kernel of the wide
    // ResBlocks with CI = M zero-padded to 128 (the padding costs nothing: TMA zero-fills, the k-steps beyond M are skipped)
    __half* mp = reinterpret_cast<__half*>(bufC);
    const long long mplane = (long long)B * T * M;
    if (run) {
      if (stride_m == 1 && stride_t == M && stride_b == (int64_t)T * M && (((uintptr_t)mel) & 15) == 0 && (mplane & 7) == 0) {
        // the decoder's own [B,T,M] output (tts_model.py:390 passes its transpose view): already channel-last, a flat split
        if ((rc = launch_split_planes_h(mel, mp, mplane, status, s))) return rc;
      } else {
        dim3 grid(ceil_div(T, 32), ceil_div(M, 32), B);
        M2_LAUNCH(M2TTS_STAGE_VOC_IN, mel_to_planes_kernel, grid, 256, 0, s, mel, (long long)stride_b, (long long)stride_m, (long long)stride_t,
                  mp, mplane, M, T, status);
      }
    }
    if ((rc = launch_voc_conv_h(run ? mp : nullptr, mplane, W(w->in_w), w->in_b, bl.in_img, nullptr, 0, bufA, (long long)B * T * C, nullptr, 0, B, M, C,
                                T, 0, M2TTS_STAGE_VOC_IN, status, s))) return rc;
  } else if (plan.in_kind == IN_TC) {
    // tensor-core input conv: channel-first copy of the mel with a 16-byte row pitch (bufC), then the tap-GEMM
    const float* xin = mel;
    if (run && !(stride_t == 1 && stride_m == Lp && stride_b == (int64_t)M * Lp && (((uintptr_t)mel) & 15) == 0)) {
      dim3 grid(ceil_div(Lp, 32), ceil_div(M, 32), B);
      M2_LAUNCH(M2TTS_STAGE_VOC_IN, mel_to_channel_first_kernel, grid, 256, 0, s, mel, (long long)stride_b, (long long)stride_m,
                (long long)stride_t, bufC, M, T, Lp);
      xin = bufC;
    }
    if ((rc = launch_conv3_tc(run ? xin : nx, Lp, W(w->in_w), bl.in_img, w->in_b, nullptr, 0, bufA, Lp, B, M, C, T, 1, 0, M2TTS_STAGE_VOC_IN, s,
                              plan.in_planes ? 2 : 0, status))) return rc;
  } else if (run) {
    ConvArgs a{mel, stride_b, stride_m, stride_t, bl.in_ffma, w->in_b, nullptr, bufA, M, C, T, 1, 0};
    a.y_pitch = Lp;
    if ((rc = launch_conv3(a, B, M2TTS_STAGE_VOC_IN, s))) return rc;
  }
  bool audio_done = false;
  for (int j = 0; j < 4; ++j) {
    const int r = kRates[j], c = c_in / 2, Lo = L * r;
    const int dil = w->res_dilation[j] > 0 ? w->res_dilation[j] : 1;
    const int kind = plan.kind[j];
    const bool last = j == 3;
    const bool next_cl = !last && (plan.kind[j + 1] == S_FUSED_H || plan.kind[j + 1] == S_FUSED_TF32 || stage_reads_planes(plan.kind[j + 1]));
    const long long in_plane = (long long)B * L * c_in, out_plane = (long long)B * Lo * c;
    if (kind == S_FUSED_H) {
      // bufA = fp16 hi/lo planes, channel-last [2][B][L][c_in] -> bufB planes [2][B][Lo][c] when the next stage reads planes,
      // plain fp32 channel-last otherwise, or straight to the waveform
      const bool planes_out = plan.out_planes[j];
      if ((rc = launch_voc_stage_fused_h(run ? (const void*)bufA : nullptr, in_plane, W(w->up_w[j]), w->up_b[j], W(w->res1_w[j]), w->res1_b[j],
                                         W(w->res2_w[j]), w->res2_b[j], last ? w->out_w : nullptr, last ? w->out_b : nullptr, bl.fs_img[j],
                                         planes_out ? (void*)bufB : nullptr, out_plane, last ? audio : (planes_out ? nullptr : bufB), B, c, L,
                                         M2TTS_STAGE_VOC_FUSED, status, s))) return rc;
      if (last) audio_done = true;
      else { float* t = bufA; bufA = bufB; bufB = t; }
      cl = true;
    } else if (kind == S_FUSED_TF32) {
      // bufA channel-last [B][L][c_in] -> bufB channel-last [B][Lo][c] (or straight to the waveform)
      if ((rc = launch_voc_stage_fused(run ? bufA : nx, W(w->up_w[j]), w->up_b[j], W(w->res1_w[j]), w->res1_b[j], W(w->res2_w[j]), w->res2_b[j],
                                       last ? w->out_w : nullptr, last ? w->out_b : nullptr, bl.fs_img[j], last ? audio : bufB, B, c, L,
                                       M2TTS_STAGE_VOC_FUSED, s))) return rc;
      if (last) audio_done = true;
      else { float* t = bufA; bufA = bufB; bufB = t; }
      cl = true;
    } else if (kind == S_UPH_RESH || kind == S_TC_RESH) {
      // C = 64: the upsampler writes fp16 hi/lo planes channel-last (bufB) and the whole ResBlock is ONE 16-bit split kernel
      // (bufB -> bufA planes): v = lrelu(conv1(u)) and the residual never travel through HBM
      if (kind == S_UPH_RESH) rc = launch_voc_up_h(run ? (const void*)bufA : nullptr, in_plane, W(w->up_w[j]), w->up_b[j], bl.up_img[j], bufB, out_plane, B,
                                                   c_in, L, M2TTS_STAGE_VOC_UP, status, s);
      else rc = launch_convT_tc(run ? bufA : nx, Lp, W(w->up_w[j]), bl.up_img[j], w->up_b[j], bufB, Lo, B, c_in, c, L, r, s, 2, status);
      if (rc) return rc;
      if ((rc = launch_voc_res_h(run ? (const void*)bufB : nullptr, out_plane, W(w->res1_w[j]), w->res1_b[j], W(w->res2_w[j]), w->res2_b[j], bl.r1_img[j],
                                 bufA, out_plane, nullptr, B, c, Lo, M2TTS_STAGE_VOC_RES1, status, s))) return rc;
      Lp = Lo;
      cl = true;
    } else if (kind == S_UPH_CONVH || kind == S_TC_CONVH) {
      // C = 128: the upsampler writes fp16 hi/lo planes channel-last (bufB); the two convolutions of the ResBlock run on those
      // planes (no splitter, one accumulator for the three taps); conv2 writes what the next stage reads
      if (kind == S_UPH_CONVH) rc = launch_voc_up_h(run ? (const void*)bufA : nullptr, in_plane, W(w->up_w[j]), w->up_b[j], bl.up_img[j], bufB, out_plane, B,
                                                    c_in, L, M2TTS_STAGE_VOC_UP, status, s);
      else rc = launch_convT_tc(run ? bufA : nx, Lp, W(w->up_w[j]), bl.up_img[j], w->up_b[j], bufB, Lo, B, c_in, c, L, r, s, 2, status);
      if (rc) return rc;
      const bool po = plan.out_planes[j];
      if ((rc = launch_voc_conv_h(run ? (const void*)bufB : nullptr, out_plane, W(w->res1_w[j]), w->res1_b[j], bl.r1_img[j], nullptr, 0, bufC, out_plane,
                                  nullptr, 0, B, c, c, Lo, 1, M2TTS_STAGE_VOC_RES1, status, s))) return rc;
      if ((rc = launch_voc_conv_h(run ? (const void*)bufC : nullptr, out_plane, W(w->res2_w[j]), w->res2_b[j], bl.r2_img[j], bufB, out_plane,
                                  po ? (void*)bufA : nullptr, out_plane, po ? nullptr : bufA, Lo, B, c, c, Lo, 0, M2TTS_STAGE_VOC_RES2, status, s))) return rc;
      Lp = Lo;
      cl = po;
    } else if (kind == S_TC) {
      // bufA (pitch Lp) -> up -> bufB -> conv1 -> bufC -> conv2 (+ residual bufB) -> bufA; Lo = r*L is a multiple of 4
      if ((rc = launch_convT_tc(run ? bufA : nx, Lp, W(w->up_w[j]), bl.up_img[j], w->up_b[j], bufB, Lo, B, c_in, c, L, r, s))) return rc;
      if ((rc = launch_conv3_tc(run ? bufB : nx, Lo, W(w->res1_w[j]), bl.r1_img[j], w->res1_b[j], nullptr, 0, bufC, Lo, B, c, c, Lo, dil, 1,
                                M2TTS_STAGE_VOC_RES1, s))) return rc;
      const int ocl = next_cl ? (plan.kind[j + 1] == S_FUSED_H ? 2 : 1) : 0;
      if ((rc = launch_conv3_tc(run ? bufC : nx, Lo, W(w->res2_w[j]), bl.r2_img[j], w->res2_b[j], bufB, Lo, bufA, Lo, B, c, c, Lo, 1, 0,
                                M2TTS_STAGE_VOC_RES2, s, ocl, status))) return rc;
      Lp = Lo;
      cl = next_cl;
    } else if (run) {
      {  // bufB = lrelu(convT(bufA)); bufA is channel-first, or channel-last after a fused stage
        ConvArgs a{bufA, (long long)c_in * L, cl ? 1 : L, cl ? c_in : 1, w->up_w[j], w->up_b[j], nullptr, bufB, c_in, c, L, 1, 1};
        if ((rc = launch_convT(a, B, r, s))) return rc;
      }
      {  // bufC = lrelu(conv1(bufB))
        ConvArgs a{bufB, (long long)c * Lo, Lo, 1, bl.r1_ffma[j], w->res1_b[j], nullptr, bufC, c, c, Lo, dil, 1};
        if ((rc = launch_conv3(a, B, M2TTS_STAGE_VOC_RES1, s))) return rc;
      }
      {  // bufA = conv2(bufC) + bufB
        ConvArgs a{bufC, (long long)c * Lo, Lo, 1, bl.r2_ffma[j], w->res2_b[j], bufB, bufA, c, c, Lo, 1, 0};
        if ((rc = launch_conv3(a, B, M2TTS_STAGE_VOC_RES2, s))) return rc;
      }
      Lp = Lo;
      cl = false;
    }
    L = Lo; c_in = c;
  }
  // output conv + tanh: bufA [B,C/16,64T] -> audio [B,1,64T] (fused into the last stage when that stage is FUSED)
  if (!audio_done && run) {
    const int threads = 256;
    dim3 grid(ceil_div(ceil_div(L, 4), threads), B);
    M2_LAUNCH(M2TTS_STAGE_VOC_OUT, conv3_co1_tanh_kernel, grid, threads, (size_t)c_in * 3 * sizeof(float), s, bufA,
              w->out_w, w->out_b, audio, c_in, L);
  }
  return M2TTS_OK;
}

}  // namespace

extern "C" size_t m2tts_vocoder_pack_bytes(int M, int C, int precision) {
  (void)precision;      // one layout covers every precision (the images of the kernels a precision does not use stay unwritten)
  if (M <= 0 || C < 16 || C % 16 != 0) return 0;
  return blobs_bytes(M, C);
}

extern "C" int m2tts_vocoder_plan(int M, int C, int precision, const int* res_dilation, int* kinds) {
  M2_REQUIRE(kinds != nullptr, M2TTS_E_NULLPTR, "vocoder_plan: null pointer");
  M2_REQUIRE(M > 0 && C >= 16 && C % 16 == 0, M2TTS_E_UNSUPPORTED, "vocoder_plan: M=%d C=%d", M, C);
  const int one[4] = {1, 1, 1, 1};
  const VocPlan plan = make_plan(M, C, resolve_precision(precision), res_dilation != nullptr ? res_dilation : one);
  kinds[0] = plan.in_kind;
  for (int j = 0; j < 4; ++j) kinds[1 + j] = plan.kind[j];
  return M2TTS_OK;
}

extern "C" size_t m2tts_vocoder_workspace_bytes(int B, int T, int M, int C) {
  if (B <= 0 || T <= 0 || M <= 0 || C < 16) return 0;
  // widest activation: 4*C*T floats per utterance (+ row-pitch padding of the first tensor); three ping-pong buffers,
  // plus room for the weight images when the caller passes no packed buffer
  const size_t act = align_up((size_t)B * C * ((size_t)T + 4) * 4 * sizeof(float), 256);
  return 3 * act + blobs_bytes(M, C);
}

extern "C" int m2tts_vocoder_pack(const m2tts_vocoder_weights* w, int M, int C, int precision, void* packed, size_t packed_bytes,
                                  int32_t* status, m2tts_stream_t stream) {
  int rc = check_weights(w, "vocoder_pack");
  if (rc) return rc;
  M2_REQUIRE(packed != nullptr, M2TTS_E_NULLPTR, "vocoder_pack: null buffer");
  M2_REQUIRE(M > 0 && C >= 16 && C % 16 == 0, M2TTS_E_UNSUPPORTED, "vocoder_pack: M=%d C=%d", M, C);
  Carver cv(packed, packed_bytes);
  VocBlobs bl;
  M2_REQUIRE(carve_blobs(cv, M, C, &bl), M2TTS_E_WORKSPACE, "vocoder_pack: buffer too small (%zu B, need %zu) or not 256-B aligned",
             packed_bytes, blobs_bytes(M, C));
  const VocPlan plan = make_plan(M, C, resolve_precision(precision), w->res_dilation);
  return vocoder_chain(w, plan, bl, true, false, nullptr, 0, 0, 0, nullptr, 1, 1, M, C, nullptr, nullptr, nullptr, status, (cudaStream_t)stream);
}

extern "C" int m2tts_vocoder_forward(const m2tts_vocoder_weights* w, const void* packed, const float* mel, int64_t stride_b,
                                     int64_t stride_m, int64_t stride_t, float* audio, int B, int T, int M,
                                     int C, int precision, int32_t* status, void* workspace, size_t workspace_bytes,
                                     m2tts_stream_t stream) {
  M2_REQUIRE(w && mel && audio && workspace, M2TTS_E_NULLPTR, "vocoder_forward: null pointer");
  M2_REQUIRE(B > 0 && T > 0 && M > 0, M2TTS_E_BADSHAPE, "vocoder_forward: B=%d T=%d M=%d", B, T, M);
  M2_REQUIRE(B <= 65535, M2TTS_E_UNSUPPORTED, "vocoder_forward: at most 65535 utterances per call (B=%d)", B);
  M2_REQUIRE(C >= 16 && C % 16 == 0, M2TTS_E_UNSUPPORTED,
             "vocoder_forward: hidden_channels=%d must be a positive multiple of 16", C);
  int rc = check_weights(w, "vocoder_forward");
  if (rc) return rc;
  Carver cv(workspace, workspace_bytes);
  const size_t act = (size_t)B * C * ((size_t)T + 4) * 4;   // floats
  float* bufA = cv.take<float>(act);
  float* bufB = cv.take<float>(act);
  float* bufC = cv.take<float>(act);
  VocBlobs bl;
  if (packed != nullptr) {
    M2_REQUIRE((((uintptr_t)packed) & 255) == 0, M2TTS_E_WORKSPACE, "vocoder_forward: packed weights must be 256-B aligned");
    Carver pc(const_cast<void*>(packed), blobs_bytes(M, C));
    carve_blobs(pc, M, C, &bl);
  } else {
    carve_blobs(cv, M, C, &bl);
  }
  M2_REQUIRE(cv.ok(), M2TTS_E_WORKSPACE, "vocoder_forward: workspace too small (%zu B) or not 256-B aligned", workspace_bytes);
  const VocPlan plan = make_plan(M, C, resolve_precision(precision), w->res_dilation);
  return vocoder_chain(w, plan, bl, packed == nullptr, true, mel, stride_b, stride_m, stride_t, audio, B, T, M, C, bufA, bufB, bufC, status,
                       (cudaStream_t)stream);
}
