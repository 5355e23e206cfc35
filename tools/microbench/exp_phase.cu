// Micro-benchmark of the softmax exp phase of attention_spatial.cu in isolation: one warp per SM sub-partition (a lone
// warpgroup) or two, 128 scores per thread, variants that drop one ingredient at a time to expose its marginal cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -o exp_phase exp_phase.cu && ./exp_phase
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));
  const float2 j = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __fadd2_rn(x, make_float2(-j.x, -j.y));
  float2 q = __ffma2_rn(make_float2(5.508868381e-02f, 5.508868381e-02f), f, make_float2(2.426040515e-01f, 2.426040515e-01f));
  q = __ffma2_rn(q, f, make_float2(6.932762417e-01f, 6.932762417e-01f));
  q = __ffma2_rn(q, f, make_float2(9.999289404e-01f, 9.999289404e-01f));
  float2 r;
  r.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return r;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// MASK: poly share as in the kernel (3 = 25 %, 1 = 50 %, 255 = none); SUM / PACK / SCALE: keep that ingredient
template <int MASK, bool SUM, bool PACK, bool SCALE>
__global__ void __launch_bounds__(256, 1) k(const float* __restrict__ in, uint32_t* __restrict__ out, long long* cyc, int iters) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[i * 256 + threadIdx.x];
  const float2 sc2 = make_float2(0.18f, 0.18f), nmb2 = make_float2(-3.f, -3.f);
  float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 128; c += 32) {
      uint32_t pk[16];
#pragma unroll
      for (int k2 = 0; k2 < 32; k2 += 4) {
        float2 x0 = make_float2(s[c + k2], s[c + k2 + 1]), x1 = make_float2(s[c + k2 + 2], s[c + k2 + 3]);
        if (SCALE) { x0 = __ffma2_rn(x0, sc2, nmb2); x1 = __ffma2_rn(x1, sc2, nmb2); }
        if ((((c + k2) >> 1) & MASK) == MASK) x0 = exp2_poly2(x0);
        else { x0.x = exp2f(x0.x); x0.y = exp2f(x0.y); }
        if (((((c + k2) >> 1) + 1) & MASK) == MASK) x1 = exp2_poly2(x1);
        else { x1.x = exp2f(x1.x); x1.y = exp2f(x1.y); }
        if (SUM) { la = __fadd2_rn(la, x0); lb = __fadd2_rn(lb, x1); }
        if (PACK) { pk[k2 >> 1] = pack2(x0.x, x0.y); pk[(k2 >> 1) + 1] = pack2(x1.x, x1.y); }
        else { pk[k2 >> 1] = __float_as_uint(x0.x) ^ __float_as_uint(x0.y); pk[(k2 >> 1) + 1] = __float_as_uint(x1.x) ^ __float_as_uint(x1.y); }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) acc ^= pk[i];      // stands in for tcgen05.st (keeps the results live)
    }
    // feed a data dependence back so iterations cannot be merged
    s[it & 127] += __uint_as_float(acc & 0x3fu) * 1e-30f;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(la.x + la.y + lb.x + lb.y);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}


// Asymmetric split: warps of warpgroup 0 (threads 0..127) send NP0/8 of their pairs to the polynomial, warpgroup 1 NP1/8.
template <int NP>
__device__ __forceinline__ uint32_t phase(const float (&s)[128], float2& la, float2& lb) {
  const float2 sc2 = make_float2(0.18f, 0.18f), nmb2 = make_float2(-3.f, -3.f);
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < 128; c += 32) {
    uint32_t pk[16];
#pragma unroll
    for (int k2 = 0; k2 < 32; k2 += 4) {
      float2 x0 = __ffma2_rn(make_float2(s[c + k2], s[c + k2 + 1]), sc2, nmb2);
      float2 x1 = __ffma2_rn(make_float2(s[c + k2 + 2], s[c + k2 + 3]), sc2, nmb2);
      if ((((c + k2) >> 1) & 7) < NP) x0 = exp2_poly2(x0);
      else { x0.x = exp2f(x0.x); x0.y = exp2f(x0.y); }
      if (((((c + k2) >> 1) + 1) & 7) < NP) x1 = exp2_poly2(x1);
      else { x1.x = exp2f(x1.x); x1.y = exp2f(x1.y); }
      la = __fadd2_rn(la, x0); lb = __fadd2_rn(lb, x1);
      pk[k2 >> 1] = pack2(x0.x, x0.y); pk[(k2 >> 1) + 1] = pack2(x1.x, x1.y);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= pk[i];
  }
  return acc;
}
template <int NP0, int NP1>
__global__ void __launch_bounds__(256, 1) k2w(const float* __restrict__ in, uint32_t* __restrict__ out, long long* cyc, int iters) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[i * 256 + threadIdx.x];
  float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x < 128) {
    for (int it = 0; it < iters; ++it) { acc ^= phase<NP0>(s, la, lb); s[it & 127] += __uint_as_float(acc & 0x3fu) * 1e-30f; }
  } else {
    for (int it = 0; it < iters; ++it) { acc ^= phase<NP1>(s, la, lb); s[it & 127] += __uint_as_float(acc & 0x3fu) * 1e-30f; }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(la.x + la.y + lb.x + lb.y);
  if ((threadIdx.x & 127) == 0 && blockIdx.x == 0) cyc[threadIdx.x >> 7] = t1 - t0;
}
template <int NP0, int NP1>
void run2(const float* in) {
  uint32_t* d; long long* c; cudaMalloc(&d, 148 * 256 * 4); cudaMalloc(&c, 16);
  const int iters = 200;
  k2w<NP0, NP1><<<148, 256>>>(in, d, c, iters);
  k2w<NP0, NP1><<<148, 256>>>(in, d, c, iters);
  long long h[2]; cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
  printf("poly eighths WG0 %d / WG1 %d: WG0 %7.0f, WG1 %7.0f cycles per tile phase (both running)\n", NP0, NP1, double(h[0]) / iters, double(h[1]) / iters);
  cudaFree(d); cudaFree(c);
}

template <int MASK, bool SUM, bool PACK, bool SCALE>
void run(const char* name, int warps_per_smsp, const float* in) {
  uint32_t* d; long long* c; cudaMalloc(&d, 148 * 256 * 4); cudaMalloc(&c, 8);
  const int iters = 200;
  k<MASK, SUM, PACK, SCALE><<<148, 128 * warps_per_smsp>>>(in, d, c, iters);
  k<MASK, SUM, PACK, SCALE><<<148, 128 * warps_per_smsp>>>(in, d, c, iters);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps/SMSP %d: %7.0f cycles per 128x128 tile phase (%.1f per score per warp)\n", name, warps_per_smsp,
         double(h) / iters, double(h) / iters / 128);
  cudaFree(d); cudaFree(c);
}
int main() {
  float* in; cudaMalloc(&in, 128 * 256 * 4);
  float* h = new float[128 * 256];
  for (int i = 0; i < 128 * 256; ++i) h[i] = -float(i % 37) * 0.3f;
  cudaMemcpy(in, h, 128 * 256 * 4, cudaMemcpyHostToDevice);
  for (int w = 1; w <= 2; ++w) {
    run<3, true, true, true>("kernel mix (25% poly, sum, pack, scale)", w, in);
    run<255, true, true, true>("no poly", w, in);
    run<1, true, true, true>("50% poly", w, in);
    run<3, false, true, true>("25% poly, no sum", w, in);
    run<3, true, false, true>("25% poly, no pack", w, in);
    run<3, true, true, false>("25% poly, no scale", w, in);
    run<255, false, false, false>("MUFU only", w, in);
    run<255, false, true, false>("MUFU + pack", w, in);
    run<255, true, false, false>("MUFU + sum", w, in);
  }
  run2<2, 2>(in); run2<3, 3>(in); run2<0, 4>(in); run2<0, 5>(in); run2<0, 6>(in); run2<1, 4>(in); run2<1, 5>(in); run2<2, 4>(in); run2<0, 8>(in); run2<4, 4>(in);
  return 0;
}
