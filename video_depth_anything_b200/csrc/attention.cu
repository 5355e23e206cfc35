// Temporal attention (the spatial ViT attention lives in attention_spatial.cu).
//
//     Temporal attention (motion_module/motion_module.py:230-297, motion_module/attention.py:182-211):
//     32-frame sequences at every spatial position, 8 heads of d = C/8.  Bandwidth-bound; one CTA per
//     position, one warp per head, K/V staged in shared memory, fp32 math.
#include "../../include/vda.h"
#include "common.cuh"

namespace vda {

// ---------------------------------------------------------------------------------------------
// temporal attention: CTA = (spatial position, group of `hpc` heads); one warp per head; lane = query frame
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
temporal_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Tn, int hw, int C, int heads, int hpc) {
  extern __shared__ float sm_t[];   // [3][Tn][W+1] fp32 (q | k | v) for this CTA's head group, W = hpc*dh
  const int pos = blockIdx.x;
  const int dh = C / heads;
  const int Wc = hpc * dh;                    // columns handled by this CTA
  const int col0 = blockIdx.y * Wc;
  const int ldc = Wc + 1;                     // odd pitch: lane-strided reads are conflict-free
  // cooperative load: per frame, three contiguous segments of Wc elements
  const int vec_per_seg = Wc / 8;
  const int total = 3 * Tn * vec_per_seg;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int v8 = idx % vec_per_seg;
    const int rest = idx / vec_per_seg;
    const int f = rest % Tn, which = rest / Tn;
    const T* src = qkv + (static_cast<long long>(f) * hw + pos) * (3LL * C) + which * C + col0 + v8 * 8;
    const uint4 u = *reinterpret_cast<const uint4*>(src);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float* dst = sm_t + (which * Tn + f) * ldc + v8 * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 p = H16<T>::unpack2(w[i]);
      dst[2 * i] = p.x;
      dst[2 * i + 1] = p.y;
    }
  }
  __syncthreads();
  const float* sq = sm_t;
  const float* sk = sq + Tn * ldc;
  const float* sv = sk + Tn * ldc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = rsqrtf(static_cast<float>(dh));
  if (warp < hpc && lane < Tn) {
    const int co = warp * dh;
    float sc[32];
    float mx = -INFINITY;
    const float* qrow = sq + lane * ldc + co;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float a = -INFINITY;
      if (j < Tn) {
        const float* krow = sk + j * ldc + co;
        a = 0.f;
        for (int c = 0; c < dh; ++c) a = fmaf(qrow[c], krow[c], a);
        a *= scale;
      }
      sc[j] = a;
      mx = fmaxf(mx, a);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      sc[j] = __expf(sc[j] - mx);
      sum += sc[j];
    }
    const float inv = 1.f / sum;
    T* orow = out + (static_cast<long long>(lane) * hw + pos) * C + col0 + co;
    for (int c = 0; c < dh; c += 2) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < Tn) {
          a0 = fmaf(sc[j], sv[j * ldc + co + c], a0);
          a1 = fmaf(sc[j], sv[j * ldc + co + c + 1], a1);
        }
      }
      *reinterpret_cast<uint32_t*>(orow + c) = H16<T>::pack2(a0 * inv, a1 * inv);
    }
  }
}

}  // namespace vda

using namespace vda;

extern "C" int vda_attention_temporal(const void* qkv, void* out, int T, int hw, int C, int heads, int dtype,
                                      void* stream) {
  VDA_CHECK(T > 0 && T <= 32, "temporal attention supports 1..32 frames (temporal_max_len, dpt_temporal.py:38), got %d", T);
  VDA_CHECK(C % 8 == 0 && C % heads == 0 && (C / heads) % 8 == 0 && heads <= 8,
            "bad channel/head split C=%d heads=%d (head dim must be a multiple of 8, heads <= 8)", C, heads);
  const int dh = C / heads;
  int hpc = heads;   // heads per CTA: largest divisor of `heads` whose fp32 q|k|v tile stays <= 50 KB
  while (hpc > 1 && (heads % hpc != 0 || static_cast<size_t>(3) * T * (hpc * dh + 1) * 4 > 50 * 1024)) --hpc;
  const size_t smem = static_cast<size_t>(3) * T * (hpc * dh + 1) * sizeof(float);
  VDA_CHECK(smem <= 100 * 1024, "temporal attention tile does not fit shared memory (head dim %d)", dh);
  dim3 grid(hw, heads / hpc);
  const int threads = 32 * hpc < 64 ? 64 : 32 * hpc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == VDA_BF16) {
    auto k = temporal_attention_kernel<__nv_bfloat16>;
    VDA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    k<<<grid, threads, smem, st>>>(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), T, hw, C,
                                   heads, hpc);
  } else {
    auto k = temporal_attention_kernel<__half>;
    VDA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    k<<<grid, threads, smem, st>>>(static_cast<const __half*>(qkv), static_cast<__half*>(out), T, hw, C, heads, hpc);
  }
  VDA_CUDA(cudaGetLastError());
  return 0;
}
