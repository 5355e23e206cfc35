// Key-frame scale/shift alignment of the long-video driver (video_depth.py:216-252, utils/util.py:40-74) on device:
// a one-pass 4-sum reduction (warp shuffles, per-CTA double partials summed in a fixed order: bit-reproducible
// run to run, no atomics), a tiny solve kernel that keeps
// (scale, shift) in device memory (no host round trip), and a fused affine + clamp + cross-fade kernel.
#include "../../include/vda.h"
#include "common.cuh"

namespace vda {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums: [0]=sum p*p  [1]=sum p  [2]=sum p*t  [3]=sum t   (a_11 = n is known)
__global__ void __launch_bounds__(256) lsq_sums_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                       long long n, double* __restrict__ sums) {
  float s_pp = 0.f, s_p = 0.f, s_pt = 0.f, s_t = 0.f;
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 p = reinterpret_cast<const float4*>(pred)[i];
    const float4 t = reinterpret_cast<const float4*>(target)[i];
    s_pp += p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w;
    s_p += p.x + p.y + p.z + p.w;
    s_pt += p.x * t.x + p.y * t.y + p.z * t.z + p.w * t.w;
    s_t += t.x + t.y + t.z + t.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = n4 << 2; i < n; ++i) {
      const float p = pred[i], t = target[i];
      s_pp += p * p; s_p += p; s_pt += p * t; s_t += t;
    }
  }
  __shared__ double sh[4][8];
  double v[4] = {s_pp, s_p, s_pt, s_t};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = warp_sum_d(v[k]);
    if (lane == 0) sh[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) a += sh[threadIdx.x][w];
    sums[static_cast<size_t>(blockIdx.x) * 4 + threadIdx.x] = a;      // per-CTA partial, reduced by the solve kernel
  }
}

__global__ void lsq_solve_kernel(const double* __restrict__ partials, int n_part, long long n,
                                 float* __restrict__ scale_shift) {
  double sums[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < n_part; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) sums[k] += partials[static_cast<size_t>(i) * 4 + k];
  // utils/util.py:51-62 evaluated on float32 sums like the reference (np.float32 scalars)
  const float a00 = static_cast<float>(sums[0]), a01 = static_cast<float>(sums[1]), a11 = static_cast<float>(n);
  const float b0 = static_cast<float>(sums[2]), b1 = static_cast<float>(sums[3]);
  const float det = a00 * a11 - a01 * a01;
  float x0 = 1.f, x1 = 0.f;
  if (det != 0.f) {
    x0 = (a11 * b0 - a01 * b1) / det;
    x1 = (-a01 * b0 + a00 * b1) / det;
  }
  scale_shift[0] = x0;
  scale_shift[1] = x1;
}

__global__ void __launch_bounds__(256) affine_clamp_blend_kernel(const float* __restrict__ x, const float* __restrict__ ss,
                                                                 const float* __restrict__ prev, const float* __restrict__ bw,
                                                                 float* __restrict__ out, int frames, long long hw) {
  const float s = ss[0], t = ss[1];
  const long long total = static_cast<long long>(frames) * hw;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = x[i] * s + t;
    v = v < 0.f ? 0.f : v;
    if (bw) {
      const float w = bw[i / hw];
      v = prev[i] * (1.f - w) + v * w;
    }
    out[i] = v;
  }
}

}  // namespace vda

using namespace vda;

extern "C" int vda_lsq_scale_shift(const float* pred, const float* target, int64_t n, float* scale_shift, double* scratch,
                                   void* stream) {
  VDA_CHECK(n > 0, "lsq: empty input");
  VDA_CHECK((reinterpret_cast<uintptr_t>(pred) & 15) == 0 && (reinterpret_cast<uintptr_t>(target) & 15) == 0,
            "lsq: inputs must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long g = (n / 4 + 255) / 256;
  if (g > VDA_LSQ_MAX_PARTIALS) g = VDA_LSQ_MAX_PARTIALS;
  if (g < 1) g = 1;
  lsq_sums_kernel<<<static_cast<unsigned>(g), 256, 0, st>>>(pred, target, n, scratch);
  lsq_solve_kernel<<<1, 1, 0, st>>>(scratch, static_cast<int>(g), n, scale_shift);
  VDA_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vda_affine_clamp_blend(const float* x, const float* scale_shift, const float* prev, const float* blend_w,
                                      float* out, int frames, int64_t hw, void* stream) {
  VDA_CHECK(frames > 0 && hw > 0, "affine: empty input");
  VDA_CHECK(!blend_w || prev, "affine: blend weights given without prev frames");
  const long long total = static_cast<long long>(frames) * hw;
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  affine_clamp_blend_kernel<<<static_cast<unsigned>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, scale_shift, prev, blend_w, out, frames, hw);
  VDA_CUDA(cudaGetLastError());
  return 0;
}
