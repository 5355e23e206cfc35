// Micro-benchmark: issue rate of MUFU.EX2 for ONE warp per SM sub-partition (the situation of a lone softmax
// warpgroup), alone and interleaved with packed / scalar FMA-pipe work.  Prints cycles per MUFU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_issue mufu_issue.cu && ./mufu_issue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NF2, int NF1, int WARPS_PER_SMSP>
__global__ void k(float* out, long long* cyc, int iters) {
  float m[8];
  float2 a[8];
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = -0.001f * (threadIdx.x + i); a[i] = make_float2(0.5f + i, 0.25f * i); s[i] = 0.1f * i; }
  const float2 c1 = make_float2(1.0001f, 0.9999f), c2 = make_float2(1e-3f, -1e-3f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
#pragma unroll
      for (int j = 0; j < NF2; ++j) {
        unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&a[(i + j) & 7]);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(*reinterpret_cast<const unsigned long long*>(&c1)),
                     "l"(*reinterpret_cast<const unsigned long long*>(&c2)));
        *reinterpret_cast<unsigned long long*>(&a[(i + j) & 7]) = r;
      }
#pragma unroll
      for (int j = 0; j < NF1; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[(i + j) & 7]) : "f"(1.0001f), "f"(1e-3f));
    }
  }
  const long long t1 = clock64();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += m[i] + a[i].x + a[i].y + s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NF2, int NF1, int W>
void run() {
  float* d; long long* c; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 8);
  const int iters = 2000;
  k<NF2, NF1, W><<<148, 128 * W>>>(d, c, iters);
  k<NF2, NF1, W><<<148, 128 * W>>>(d, c, iters);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("warps/SMSP %d: per MUFU: +%d FFMA2 +%d FFMA: %.2f cycles per MUFU (per warp)\n", W, NF2, NF1, double(h) / (iters * 8.0));
  cudaFree(d); cudaFree(c);
}
int main() {
  run<0, 0, 1>(); run<1, 0, 1>(); run<2, 0, 1>(); run<3, 0, 1>(); run<4, 0, 1>(); run<0, 2, 1>(); run<0, 4, 1>(); run<0, 6, 1>();
  run<0, 0, 2>(); run<2, 0, 2>(); run<4, 0, 2>(); run<0, 4, 2>();
  return 0;
}
