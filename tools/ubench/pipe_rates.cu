// Bring-up microbenchmark (GPU box): issue cost per warp instruction of the instructions the attention softmax is made of,
// one CTA on one SM, W warps per scheduler. nvcc -arch=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 2048
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 * 1.1f, a2 = a0 * 1.2f, a3 = a0 * 1.3f, a4 = a0 * 1.4f, a5 = a0 * 1.5f, a6 = a0 * 1.6f, a7 = a0 * 1.7f;
  const float c = seed * 0.5f;
  uint64_t p0, p1, p2, p3, pc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
  asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
  asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c));
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    if (OP == 0) {        // FADD x8 independent
      a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c;
    } else if (OP == 1) { // FADD2 x4 (8 flops)
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pc));
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pc));
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pc));
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pc));
    } else if (OP == 2) { // MUFU.EX2 x8
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
    } else if (OP == 3) { // LOP3 x8
      asm volatile("and.b32 %0, %0, 0xFFFFE0FF;" : "+f"(a0)); asm volatile("and.b32 %0, %0, 0xFFFFE0FF;" : "+f"(a1));
      asm volatile("and.b32 %0, %0, 0xFFFFE0FF;" : "+f"(a2)); asm volatile("and.b32 %0, %0, 0xFFFFE0FF;" : "+f"(a3));
      asm volatile("and.b32 %0, %0, 0xFFFFE0FF;" : "+f"(a4)); asm volatile("and.b32 %0, %0, 0xFFFFE0FF;" : "+f"(a5));
      asm volatile("and.b32 %0, %0, 0xFFFFE0FF;" : "+f"(a6)); asm volatile("and.b32 %0, %0, 0xFFFFE0FF;" : "+f"(a7));
    } else if (OP == 4) { // F2FP x4 (pack pairs), results fed back
      uint32_t r;
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a0), "f"(a1)); a0 = __uint_as_float(r);
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a2), "f"(a3)); a2 = __uint_as_float(r);
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a4), "f"(a5)); a4 = __uint_as_float(r);
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a6), "f"(a7)); a6 = __uint_as_float(r);
    } else if (OP == 5) { // FMNMX3 x4
      asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(a1), "f"(c)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(a3), "f"(c));
      asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a4) : "f"(a5), "f"(c)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a6) : "f"(a7), "f"(c));
    } else if (OP == 6) { // FFMA2 x4
      asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p0) : "l"(pc)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p1) : "l"(pc));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p2) : "l"(pc)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p3) : "l"(pc));
    } else if (OP == 7) { // FFMA x8 (3-reg)
      a0 = fmaf(a0, c, a1); a1 = fmaf(a1, c, a2); a2 = fmaf(a2, c, a3); a3 = fmaf(a3, c, a4);
      a4 = fmaf(a4, c, a5); a5 = fmaf(a5, c, a6); a6 = fmaf(a6, c, a7); a7 = fmaf(a7, c, a0);
    } else if (OP == 8) { // the softmax pair mix: 3 FADD2 + 2 MUFU + 2 LOP3 + 2 F2FP, two independent pairs
      uint64_t d, l; uint32_t r; float x0, x1;
      asm volatile("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(p0), "l"(pc));
      asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(d));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x0), "f"(x1));
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(d));
      float h0, h1;
      asm volatile("and.b32 %0, %1, 0xFFFFE000;" : "=f"(h0) : "f"(x0)); asm volatile("and.b32 %0, %1, 0xFFFFE000;" : "=f"(h1) : "f"(x1));
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(l) : "f"(h0), "f"(h1));
      asm volatile("sub.f32x2 %0, %1, %2;" : "=l"(l) : "l"(d), "l"(l));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(h1), "f"(h0)); a6 += __uint_as_float(r);
      asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(l));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0)); a7 += __uint_as_float(r);
    }
  }
  const long long t1 = clock64();
  float s0, s1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(p0));
  float s2, s3;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(s2), "=f"(s3) : "l"(p2));
  out[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + s0 + s1 + s2 + s3;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int n_instr, float* out, long long* cyc) {
  for (int warps = 4; warps <= 16; warps *= 2) {
    k<OP><<<1, warps * 32>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    // warps / 4 warps per scheduler, each issuing n_instr per iteration
    printf("%-28s %2d warps/SM (%d per scheduler): %.2f cycles per warp-instruction per scheduler\n", name, warps, warps / 4,
           (double)c / ITERS / n_instr / (warps / 4));
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  run<0>("FADD", 8, out, cyc);
  run<1>("FADD2 (f32x2)", 4, out, cyc);
  run<6>("FFMA2 (f32x2)", 4, out, cyc);
  run<7>("FFMA (3 reg)", 8, out, cyc);
  run<2>("MUFU.EX2", 8, out, cyc);
  run<3>("LOP3", 8, out, cyc);
  run<4>("F2FP.F16.F32.PACK", 4, out, cyc);
  run<5>("FMNMX3", 4, out, cyc);
  run<8>("softmax pair mix (9 instr)", 9, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
