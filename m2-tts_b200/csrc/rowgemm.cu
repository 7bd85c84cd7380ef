// rowgemm.cu — fp32 row-block GEMM family for the transformer layers:
//   y[R,N] = epilogue( prologue(x)[R,K] @ W^T[K,N] )
// prologue: optional LayerNorm over K (the whole row lives in the CTA's smem tile);
// epilogue: +bias, ReLU, +residual, row-major store — or the Q/K/V split store that
// lays Q,K out head-major and d-major ([B,nh,hd,Lp]) and V as [B,nh,L,hd] so the
// attention kernel's tiles are plain 2-D sub-blocks.
// Reference ops replaced: components.py:55,70-72 (qkv), :90 (out_proj), :103 (ffn),
// :133,137 (norm1/norm2), tts_model.py:223-226 (decoder.norm + mel_projection).
//
// Tiling: CTA = 128 threads, 128 rows x (8*TN) cols per pass, thread tile 8 x TN,
// operands staged k-major in shared memory (A^T [K][128+4], W^T chunk [K][8*TN]).
#include "common.cuh"

namespace m2 {

constexpr int RG_BM = 128;
constexpr int RG_AST = RG_BM + 4;  // smem row stride of the A^T tile (floats)
constexpr int RG_THREADS = 128;

template <int TN>
__global__ void __launch_bounds__(RG_THREADS) rowgemm_kernel(RowGemmArgs a) {
  constexpr int BN = 8 * TN;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                      // [K][RG_AST]
  float* Ws = smem + (size_t)a.K * RG_AST;  // [K][BN]

  const int tid = threadIdx.x;
  const int tx = tid & 15;   // row group: rows 4tx..4tx+3 and 64+4tx..64+4tx+3 (conflict-free LDS.128)
  const int ty = tid >> 4;   // col group: cols ty*TN .. +TN-1
  const int bidx = blockIdx.y;           // utterance (qkv mode) or 0
  const int l0 = blockIdx.x * RG_BM;     // first row of this tile inside the utterance
  const int rows_here = min(RG_BM, a.L - l0);
  const long long row0 = (long long)bidx * a.L + l0;
  const int K = a.K, N = a.N;

  // ---- stage the 128 x K input tile transposed; one thread owns one row ----
  {
    const int m = tid;
    const bool valid = m < rows_here;
    const float* xr = a.x + (row0 + m) * (long long)a.ldx;
    const bool vec = ((a.ldx & 3) == 0) && ((((uintptr_t)a.x) & 15) == 0) && ((K & 3) == 0);
    if (valid && vec) {
      for (int k = 0; k < K; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + k);
        As[(k + 0) * RG_AST + m] = v.x; As[(k + 1) * RG_AST + m] = v.y;
        As[(k + 2) * RG_AST + m] = v.z; As[(k + 3) * RG_AST + m] = v.w;
      }
    } else {
      for (int k = 0; k < K; ++k) As[k * RG_AST + m] = valid ? xr[k] : 0.f;
    }
    if (a.ln_w != nullptr && valid) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += As[k * RG_AST + m];
      const float mean = s / (float)K;
      float q = 0.f;
      for (int k = 0; k < K; ++k) { const float d = As[k * RG_AST + m] - mean; q += d * d; }
      const float rstd = 1.0f / sqrtf(q / (float)K + a.eps);
      for (int k = 0; k < K; ++k)
        As[k * RG_AST + m] = (As[k * RG_AST + m] - mean) * rstd * __ldg(a.ln_w + k) + __ldg(a.ln_b + k);
    }
  }

  const bool wvec = ((N & 3) == 0) && ((((uintptr_t)a.wt) & 15) == 0);

  for (int n0 = 0; n0 < N; n0 += BN) {
    __syncthreads();  // A tile ready (first pass) / previous chunk fully consumed
    // ---- stage W^T chunk [K][BN] (zero beyond N) ----
    if (wvec) {
      constexpr int C4 = BN / 4;
      for (int idx = tid; idx < K * C4; idx += RG_THREADS) {
        const int k = idx / C4, c = (idx % C4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + c < N) v = *reinterpret_cast<const float4*>(a.wt + (long long)k * N + n0 + c);
        *reinterpret_cast<float4*>(Ws + k * BN + c) = v;
      }
    } else {
      for (int idx = tid; idx < K * BN; idx += RG_THREADS) {
        const int k = idx / BN, c = idx % BN;
        Ws[idx] = (n0 + c < N) ? a.wt[(long long)k * N + n0 + c] : 0.f;
      }
    }
    __syncthreads();

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const float* ap = As + tx * 4;
    const float* wp = Ws + ty * TN;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(ap + k * RG_AST);
      const float4 a1 = *reinterpret_cast<const float4*>(ap + k * RG_AST + 64);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
#pragma unroll
      for (int j = 0; j < TN; j += 2) {
        const float2 t = *reinterpret_cast<const float2*>(wp + k * BN + j);
        bv[j] = t.x; bv[j + 1] = t.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }

    // ---- epilogue ----
    const int nbase = n0 + ty * TN;
    if (!a.qkv_mode) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = tx * 4 + (i & 3) + (i >> 2) * 64;
        if (m >= rows_here) continue;
        const long long row = row0 + m;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int n = nbase + j;
          if (n >= N) continue;
          float v = acc[i][j];
          if (a.bias) v += __ldg(a.bias + n);
          if (a.relu) v = fmaxf(v, 0.f);
          if (a.residual) v += a.residual[row * a.ldr + n];
          a.y[row * a.ldy + n] = v;
        }
      }
    } else {
      const int H = a.nh * a.hd;
      if (a.qkv_mode == 2) {
        // tensor-core attention operands: six planes [Q_hi,Q_lo,K_hi,K_lo,V_hi,V_lo][B,nh,hd,Lp],
        // each value split as hi + lo with both halves exact in TF32; Q carries scale*log2(e).
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int n = nbase + j;
          if (n >= N) continue;
          const int which = n / H, rem = n - which * H;
          const int head = rem / a.hd, d = rem - head * a.hd;
          const float sc = (which == 0) ? a.qscale : 1.0f;
          float* hi_p = a.q + (long long)(2 * which) * a.plane_stride +
                        (((long long)bidx * a.nh + head) * a.hd + d) * a.Lp + l0 + tx * 4;
          float* lo_p = hi_p + a.plane_stride;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (l0 + tx * 4 + h * 64 >= a.Lp) continue;
            float hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float v = acc[h * 4 + i][j] * sc;
              hi[i] = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
              lo[i] = __uint_as_float(__float_as_uint(v - hi[i]) & 0xFFFFE000u);
            }
            *reinterpret_cast<float4*>(hi_p + h * 64) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(lo_p + h * 64) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 2) {
        const int n = nbase + j;
        if (n >= N) continue;
        const int which = n / H, rem = n - which * H;
        const int head = rem / a.hd, d = rem - head * a.hd;
        if (which < 2) {
          // Q / K: [B, nh, hd, Lp], two runs of 4 consecutive l per (thread, n)
          float* dst = (which == 0 ? a.q : a.k) + (((long long)bidx * a.nh + head) * a.hd + d) * a.Lp + l0 + tx * 4;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            float* dj = dst + (long long)jj * a.Lp;  // d and d+1 stay inside one head (hd even)
            if (l0 + tx * 4 < a.Lp)
              *reinterpret_cast<float4*>(dj) = make_float4(acc[0][j + jj], acc[1][j + jj], acc[2][j + jj], acc[3][j + jj]);
            if (l0 + tx * 4 + 64 < a.Lp)
              *reinterpret_cast<float4*>(dj + 64) = make_float4(acc[4][j + jj], acc[5][j + jj], acc[6][j + jj], acc[7][j + jj]);
          }
        } else {
          // V: [B, nh, L, hd], two consecutive d per store
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = tx * 4 + (i & 3) + (i >> 2) * 64;
            if (m >= rows_here) continue;
            float* dst = a.v + (((long long)bidx * a.nh + head) * a.L + l0 + m) * a.hd + d;
            *reinterpret_cast<float2*>(dst) = make_float2(acc[i][j], acc[i][j + 1]);
          }
        }
      }
    }
  }
}

template <int TN>
static int launch_tn(const RowGemmArgs& a, cudaStream_t s, int nb) {
  const size_t smem = ((size_t)a.K * RG_AST + (size_t)a.K * 8 * TN) * sizeof(float);
  M2_REQUIRE(smem <= 227 * 1024, M2TTS_E_UNSUPPORTED, "rowgemm: K=%d needs %zu B of shared memory", a.K, smem);
  M2_CUDA_OK(allow_smem(rowgemm_kernel<TN>, smem));
  dim3 grid(ceil_div(a.L, RG_BM), nb);
  M2_LAUNCH(a.stage, rowgemm_kernel<TN>, grid, RG_THREADS, smem, s, a);
  return M2TTS_OK;
}

int launch_rowgemm(const RowGemmArgs& in, cudaStream_t s) {
  RowGemmArgs a = in;
  M2_REQUIRE(a.x && a.wt, M2TTS_E_NULLPTR, "rowgemm: null operand");
  M2_REQUIRE(a.R > 0 && a.K > 0 && a.N > 0, M2TTS_E_BADSHAPE, "rowgemm: R=%d K=%d N=%d", a.R, a.K, a.N);
  M2_REQUIRE(a.K <= 256, M2TTS_E_UNSUPPORTED, "rowgemm: inner dim %d > 256 not supported", a.K);
  int nb = 1;
  if (a.qkv_mode) {
    M2_REQUIRE(a.q && (a.qkv_mode == 2 || (a.k && a.v)), M2TTS_E_NULLPTR, "rowgemm: null q/k/v");
    M2_REQUIRE(a.L > 0 && a.R % a.L == 0 && (a.Lp & 3) == 0 && a.Lp >= a.L && (a.hd & 1) == 0 &&
                   a.N == 3 * a.nh * a.hd,
               M2TTS_E_BADSHAPE, "rowgemm qkv: L=%d Lp=%d nh=%d hd=%d N=%d", a.L, a.Lp, a.nh, a.hd, a.N);
    nb = a.R / a.L;
  } else {
    M2_REQUIRE(a.y, M2TTS_E_NULLPTR, "rowgemm: null output");
    a.L = a.R;
  }
  // pick the column-tile width that wastes the fewest padded columns (ties: wider)
  int best_tn = 8; long long best_pad = -1;
  const int cands[3] = {8, 6, 4};
  for (int c = 0; c < 3; ++c) {
    const int bn = 8 * cands[c];
    const long long pad = (long long)ceil_div(a.N, bn) * bn;
    if (best_pad < 0 || pad < best_pad) { best_pad = pad; best_tn = cands[c]; }
  }
  switch (best_tn) {
    case 8: return launch_tn<8>(a, s, nb);
    case 6: return launch_tn<6>(a, s, nb);
    default: return launch_tn<4>(a, s, nb);
  }
}

}  // namespace m2
