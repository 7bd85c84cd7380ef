// Checks common.cuh's h_split_pair against h_split2 on random values: x - (hi + lo) and the two hi planes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../m2-tts_b200/csrc -I../../include split_test.cu -o split_test
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
__global__ void k(const float* x, float* out, int n) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i >= n) return;
  uint32_t h1, l1, h2, l2;
  bool bad = false;
  float amax = 0.f;
  m2::h_split2(x[i], x[i + 1], h1, l1, bad);
  m2::h_split_pair(m2::f2_pack(x[i], x[i + 1]), h2, l2, amax);
  const float2 a = __half22float2(*reinterpret_cast<__half2*>(&h1)), b = __half22float2(*reinterpret_cast<__half2*>(&l1));
  const float2 c = __half22float2(*reinterpret_cast<__half2*>(&h2)), d = __half22float2(*reinterpret_cast<__half2*>(&l2));
  out[4 * i] = x[i] - (a.x + b.x); out[4 * i + 1] = x[i] - (c.x + d.x); out[4 * i + 2] = a.x - c.x; out[4 * i + 3] = amax;
  out[4 * i + 4] = x[i + 1] - (a.y + b.y); out[4 * i + 5] = x[i + 1] - (c.y + d.y); out[4 * i + 6] = a.y - c.y; out[4 * i + 7] = amax;
}
int main() {
  const int n = 1 << 16;
  float *x, *o;
  cudaMallocManaged(&x, n * 4); cudaMallocManaged(&o, n * 16);
  srand(1);
  for (int i = 0; i < n; ++i) x[i] = ((float)rand() / RAND_MAX - 0.5f) * powf(10.f, (float)(rand() % 9 - 5));
  k<<<n / 2 / 256, 256>>>(x, o, n);
  cudaDeviceSynchronize();
  double e1 = 0, e2 = 0, dh = 0;
  for (int i = 0; i < n; ++i) {
    const double r = fabs(x[i]) + 1e-30;
    e1 = fmax(e1, fabs(o[4 * i]) / r); e2 = fmax(e2, fabs(o[4 * i + 1]) / r); dh = fmax(dh, fabs(o[4 * i + 2]) / r);
  }
  printf("max rel |x - (hi+lo)|: h_split2 %.3e, h_split_pair %.3e; max rel hi difference %.3e; %s\n", e1, e2, dh, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
