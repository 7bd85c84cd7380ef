// attention.cu — fp32 flash-style (online softmax) multi-head attention.
// Replaces components.py:75-87: scores = q k^T * scale; masked_fill(key >= len, -1e9);
// softmax; attn @ v — without ever materialising the [B,heads,L,L] score tensor
// (6 GB per decoder layer at the stage2 10 s configuration).
//
// Operand layouts (written by the QKV row-GEMM epilogue):
//   q, k : [B, nh, hd, Lp]  (d-major: a 2-D tile is hd rows x 64/128 contiguous positions)
//   v    : [B, nh, L, hd]
//   ctx  : [B, L, nh*hd]    (what out_proj consumes)
// CTA = 128 threads = one (utterance, head, 128-query tile); key tiles of 64 stream through
// a 2-deep cp.async ring. Each thread owns an 8x8 score sub-tile (rows {4tx..+3, 64+4tx..+3},
// keys {4ty..+3, 32+4ty..+3}: every LDS.128 of a quarter-warp is contiguous) and an 8 x hd/8
// output sub-tile; P goes through shared memory once per tile.
#include "common.cuh"
#include <math.h>

namespace m2 {

constexpr int AT_BQ = 128, AT_BK = 64, AT_THREADS = 128;
constexpr int AT_PST = AT_BQ + 4;  // P^T row stride: conflict-free 128-bit stores

template <int HD>
struct AttSmem {
  static constexpr int q_floats = HD * AT_BQ;
  static constexpr int k_floats = HD * AT_BK;   // per buffer
  static constexpr int v_floats = AT_BK * HD;   // per buffer
  static constexpr int p_floats = AT_BK * AT_PST;
  static constexpr size_t bytes = (size_t)(q_floats + 2 * k_floats + 2 * v_floats + p_floats) * sizeof(float);
};

template <int HD>
__device__ __forceinline__ void att_issue_kv_tile(float* Ks, float* Vs, const float* __restrict__ kg,
                                                  const float* __restrict__ vg, int k0, int L, int Lp,
                                                  int tid) {
  // K tile: HD rows x 64 positions (16 chunks of 16 B per row)
  for (int idx = tid; idx < HD * 16; idx += AT_THREADS) {
    const int d = idx >> 4, c = (idx & 15) * 4;
    float* dst = Ks + d * AT_BK + c;
    if (k0 + c < Lp) cp_async16(dst, kg + (long long)d * Lp + k0 + c);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // V tile: 64 keys x HD contiguous floats
  constexpr int CV = HD / 4;
  for (int idx = tid; idx < AT_BK * CV; idx += AT_THREADS) {
    const int kk = idx / CV, c = (idx % CV) * 4;
    float* dst = Vs + kk * HD + c;
    if (k0 + kk < L) cp_async16(dst, vg + (long long)(k0 + kk) * HD + c);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int HD>
__global__ void __launch_bounds__(AT_THREADS, (AttSmem<HD>::bytes > 110 * 1024) ? 1 : 2)
attention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                 float* __restrict__ ctx, const int64_t* __restrict__ lengths, int L, int Lp, int nh,
                 float qscale_log2) {
  static_assert(HD % 8 == 0 && HD >= 8 && HD <= 64, "head_dim must be a multiple of 8 in [8,64]");
  constexpr int DT = HD / 8;  // output dims per thread
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                                 // [HD][128]
  float* Ks = Qs + AttSmem<HD>::q_floats;           // [2][HD][64]
  float* Vs = Ks + 2 * AttSmem<HD>::k_floats;       // [2][64][HD]
  float* Ps = Vs + 2 * AttSmem<HD>::v_floats;       // [64][132]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int tx = warp * 4 + (lane >> 3);  // 0..15 query group
  const int ty = lane & 7;                // 0..7  key group / output-dim group
  const int q0 = blockIdx.x * AT_BQ;
  const int head = blockIdx.y, b = blockIdx.z;

  const long long bh = (long long)b * nh + head;
  const float* qg = q + bh * HD * Lp;
  const float* kg = k + bh * HD * Lp;
  const float* vg = v + bh * (long long)L * HD;

  // key range and mask semantics
  int Leff = L;
  bool all_masked = false;
  if (lengths != nullptr) {
    const long long len = lengths[b];
    if (len <= 0) all_masked = true;          // every score is -1e9 -> uniform softmax over all L keys
    else if (len < L) Leff = (int)len;        // keys >= len contribute exp(-1e9 - m) == 0 exactly
  }
  const int nkt = (Leff + AT_BK - 1) / AT_BK;

  att_issue_kv_tile<HD>(Ks, Vs, kg, vg, 0, L, Lp, tid);
  cp_async_commit();

  // Q tile, pre-scaled by scale*log2(e) so the softmax runs on exp2
  for (int idx = tid; idx < HD * 32; idx += AT_THREADS) {
    const int d = idx >> 5, c = (idx & 31) * 4;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + c < Lp) {
      val = *reinterpret_cast<const float4*>(qg + (long long)d * Lp + q0 + c);
      val.x *= qscale_log2; val.y *= qscale_log2; val.z *= qscale_log2; val.w *= qscale_log2;
    }
    *reinterpret_cast<float4*>(Qs + d * AT_BQ + c) = val;
  }

  float o[8][DT];
  float m_run[8], l_run[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    m_run[i] = -INFINITY; l_run[i] = 0.f;
#pragma unroll
    for (int c = 0; c < DT; ++c) o[i][c] = 0.f;
  }

  for (int t = 0; t < nkt; ++t) {
    const int buf = t & 1;
    cp_async_wait<0>();
    __syncthreads();  // tile t (and Q) visible; everyone is done with tile t-1 and with Ps
    if (t + 1 < nkt) {
      att_issue_kv_tile<HD>(Ks + (buf ^ 1) * AttSmem<HD>::k_floats, Vs + (buf ^ 1) * AttSmem<HD>::v_floats,
                            kg, vg, (t + 1) * AT_BK, L, Lp, tid);
      cp_async_commit();
    }
    const float* Kb = Ks + buf * AttSmem<HD>::k_floats;
    const float* Vb = Vs + buf * AttSmem<HD>::v_floats;

    // ---- S = Q K^T (log2 domain) ----
    float s[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < HD; ++d) {
      const float4 a0 = *reinterpret_cast<const float4*>(Qs + d * AT_BQ + tx * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(Qs + d * AT_BQ + 64 + tx * 4);
      const float4 b0 = *reinterpret_cast<const float4*>(Kb + d * AT_BK + ty * 4);
      const float4 b1 = *reinterpret_cast<const float4*>(Kb + d * AT_BK + 32 + ty * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
    }

    // ---- mask ----
    const int kbase = t * AT_BK;
    if (all_masked) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kk = kbase + ty * 4 + (j & 3) + (j >> 2) * 32;
        const float fill = (kk < L) ? 0.f : -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][j] = fill;
      }
    } else if (kbase + AT_BK > Leff) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kk = kbase + ty * 4 + (j & 3) + (j >> 2) * 32;
        if (kk >= Leff) {
#pragma unroll
          for (int i = 0; i < 8; ++i) s[i][j] = -INFINITY;
        }
      }
    }

    // ---- online softmax: row max over the 8 key-group lanes, rescale, P -> smem ----
    float alpha[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float mx = s[i][0];
#pragma unroll
      for (int j = 1; j < 8; ++j) mx = fmaxf(mx, s[i][j]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      const float m_new = fmaxf(m_run[i], mx);
      alpha[i] = exp2f(m_run[i] - m_new);
      m_run[i] = m_new;
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[i][j] = exp2f(s[i][j] - m_new); rs += s[i][j]; }
      l_run[i] = l_run[i] * alpha[i] + rs;
#pragma unroll
      for (int c = 0; c < DT; ++c) o[i][c] *= alpha[i];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = ty * 4 + (j & 3) + (j >> 2) * 32;
      float* pr = Ps + kk * AT_PST + tx * 4;
      *reinterpret_cast<float4*>(pr) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
      *reinterpret_cast<float4*>(pr + 64) = make_float4(s[4][j], s[5][j], s[6][j], s[7][j]);
    }
    __syncthreads();  // P visible

    // ---- O += P V ----
#pragma unroll 4
    for (int kk = 0; kk < AT_BK; ++kk) {
      const float4 p0 = *reinterpret_cast<const float4*>(Ps + kk * AT_PST + tx * 4);
      const float4 p1 = *reinterpret_cast<const float4*>(Ps + kk * AT_PST + 64 + tx * 4);
      const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      float vv[DT];
      const float* vr = Vb + kk * HD + ty * DT;
      if constexpr (DT % 4 == 0) {
#pragma unroll
        for (int c = 0; c < DT; c += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(vr + c);
          vv[c] = t4.x; vv[c + 1] = t4.y; vv[c + 2] = t4.z; vv[c + 3] = t4.w;
        }
      } else if constexpr (DT % 2 == 0) {
#pragma unroll
        for (int c = 0; c < DT; c += 2) {
          const float2 t2 = *reinterpret_cast<const float2*>(vr + c);
          vv[c] = t2.x; vv[c + 1] = t2.y;
        }
      } else {
#pragma unroll
        for (int c = 0; c < DT; ++c) vv[c] = vr[c];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < DT; ++c) o[i][c] = fmaf(pv[i], vv[c], o[i][c]);
    }
  }

  // ---- finalise: full row sums across the key-group lanes, normalise, store ----
  const int H = nh * HD;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float l = l_run[i];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    l += __shfl_xor_sync(0xffffffffu, l, 4);
    const int qi = q0 + tx * 4 + (i & 3) + (i >> 2) * 64;
    if (qi < L) {
      const float inv = 1.0f / l;
      float* dst = ctx + ((long long)b * L + qi) * H + head * HD + ty * DT;
      if constexpr (DT % 2 == 0) {
#pragma unroll
        for (int c = 0; c < DT; c += 2)
          *reinterpret_cast<float2*>(dst + c) = make_float2(o[i][c] * inv, o[i][c + 1] * inv);
      } else {
#pragma unroll
        for (int c = 0; c < DT; ++c) dst[c] = o[i][c] * inv;
      }
    }
  }
}

template <int HD>
static int launch_hd(const float* q, const float* k, const float* v, float* ctx, const int64_t* lengths,
                     int B, int L, int Lp, int nh, cudaStream_t s) {
  const size_t smem = AttSmem<HD>::bytes;
  M2_CUDA_OK(allow_smem(attention_kernel<HD>, smem));
  dim3 grid(ceil_div(L, AT_BQ), nh, B);
  const float qscale_log2 = (float)((1.0 / sqrt((double)HD)) * 1.4426950408889634);
  M2_LAUNCH(M2TTS_STAGE_ATTENTION, attention_kernel<HD>, grid, AT_THREADS, smem, s, q, k, v, ctx, lengths,
            L, Lp, nh, qscale_log2);
  return M2TTS_OK;
}

int launch_attention(const float* q, const float* k, const float* v, float* ctx, const int64_t* lengths,
                     int B, int L, int Lp, int nh, int hd, cudaStream_t s) {
  M2_REQUIRE(q && k && v && ctx, M2TTS_E_NULLPTR, "attention: null pointer");
  M2_REQUIRE(B > 0 && L > 0 && nh > 0 && B <= 65535 && nh <= 65535, M2TTS_E_BADSHAPE,
             "attention: B=%d L=%d nh=%d", B, L, nh);
  switch (hd) {
    case 8: return launch_hd<8>(q, k, v, ctx, lengths, B, L, Lp, nh, s);
    case 16: return launch_hd<16>(q, k, v, ctx, lengths, B, L, Lp, nh, s);
    case 24: return launch_hd<24>(q, k, v, ctx, lengths, B, L, Lp, nh, s);
    case 32: return launch_hd<32>(q, k, v, ctx, lengths, B, L, Lp, nh, s);
    case 40: return launch_hd<40>(q, k, v, ctx, lengths, B, L, Lp, nh, s);
    case 48: return launch_hd<48>(q, k, v, ctx, lengths, B, L, Lp, nh, s);
    case 56: return launch_hd<56>(q, k, v, ctx, lengths, B, L, Lp, nh, s);
    case 64: return launch_hd<64>(q, k, v, ctx, lengths, B, L, Lp, nh, s);
    default:
      set_error("attention: head_dim %d unsupported (multiples of 8 up to 64)", hd);
      return M2TTS_E_UNSUPPORTED;
  }
}

}  // namespace m2
