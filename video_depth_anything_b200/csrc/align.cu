// Key-frame scale/shift alignment of the long-video driver (video_depth.py:216-252, utils/util.py:40-74) on device:
// a one-pass 4-sum reduction (warp shuffles, per-CTA double partials summed in a fixed order: bit-reproducible
// run to run, no atomics), a tiny solve kernel that keeps
// (scale, shift) in device memory (no host round trip), and a fused affine + clamp + cross-fade kernel.
#include "../../include/vda.h"
#include "common.cuh"
#include <cooperative_groups.h>
#include <vector>

namespace vda {

namespace cg = cooperative_groups;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The four running sums of utils/util.py:40-62 for one thread: [pp]=sum p*p  [p]=sum p  [pt]=sum p*t  [t]=sum t
// (a_11 = n is known).  Explicit fmaf / add sequence: the per-window kernel and the fused chain kernel below must
// produce the same bits (1 GPU and N GPUs give identical videos), whatever the compiler would contract.
struct LsqAcc { float pp, p, pt, t; };
__device__ __forceinline__ void lsq_acc(LsqAcc& a, float p, float t) {
  a.pp = fmaf(p, p, a.pp);
  a.p += p;
  a.pt = fmaf(p, t, a.pt);
  a.t += t;
}
// affine + clamp of the alignment (video_depth.py:234-250): one fused multiply-add, then max(., 0)
__device__ __forceinline__ float affine_clamp(float x, float s, float t) {
  const float v = fmaf(x, s, t);
  return v < 0.f ? 0.f : v;
}
// 4 consecutive floats: one 16-byte load when the address allows it (the result is the same either way)
__device__ __forceinline__ float4 ld_quad(const float* __restrict__ base, long long e, bool vec) {
  if (vec) return *reinterpret_cast<const float4*>(base + e);
  return make_float4(base[e], base[e + 1], base[e + 2], base[e + 3]);
}
// per-CTA reduction of the thread sums to doubles, stored as this CTA's partial (fixed order: warp butterflies, then
// the warps in index order)
__device__ __forceinline__ void lsq_store_partial(const LsqAcc& a, double* __restrict__ sums) {
  __shared__ double sh[4][8];
  double v[4] = {a.pp, a.p, a.pt, a.t};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = warp_sum_d(v[k]);
    if (lane == 0) sh[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double acc = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) acc += sh[threadIdx.x][w];
    sums[static_cast<size_t>(blockIdx.x) * 4 + threadIdx.x] = acc;
  }
  __syncthreads();   // sh may be reused by the caller's next round
}
// (scale, shift) from the per-CTA partials, by ONE warp in a fixed order (lane l sums partials l, l+32, ... ascending,
// then a butterfly): every lane returns the same values.  utils/util.py:51-62 evaluated on float32 sums like the
// reference (np.float32 scalars).
__device__ __forceinline__ float2 lsq_solve_warp(const double* __restrict__ partials, int n_part, long long n) {
  const int lane = threadIdx.x & 31;
  double sums[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = lane; i < n_part; i += 32)
#pragma unroll
    for (int k = 0; k < 4; ++k) sums[k] += partials[static_cast<size_t>(i) * 4 + k];
#pragma unroll
  for (int k = 0; k < 4; ++k) sums[k] = warp_sum_d(sums[k]);
  const float a00 = static_cast<float>(sums[0]), a01 = static_cast<float>(sums[1]), a11 = static_cast<float>(n);
  const float b0 = static_cast<float>(sums[2]), b1 = static_cast<float>(sums[3]);
  const float det = a00 * a11 - a01 * a01;
  float x0 = 1.f, x1 = 0.f;
  if (det != 0.f) {
    x0 = __fdiv_rn(a11 * b0 - a01 * b1, det);
    x1 = __fdiv_rn(-a01 * b0 + a00 * b1, det);
  }
  return make_float2(x0, x1);
}

__global__ void __launch_bounds__(256) lsq_sums_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                       long long n, double* __restrict__ sums) {
  LsqAcc a = {0.f, 0.f, 0.f, 0.f};
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 p = ld_quad(pred, 4 * i, true);
    const float4 t = ld_quad(target, 4 * i, true);
    lsq_acc(a, p.x, t.x); lsq_acc(a, p.y, t.y); lsq_acc(a, p.z, t.z); lsq_acc(a, p.w, t.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) lsq_acc(a, pred[i], target[i]);
  lsq_store_partial(a, sums);
}

__global__ void __launch_bounds__(32) lsq_solve_kernel(const double* __restrict__ partials, int n_part, long long n,
                                                       float* __restrict__ scale_shift) {
  const float2 ss = lsq_solve_warp(partials, n_part, n);
  if (threadIdx.x == 0) {
    scale_shift[0] = ss.x;
    scale_shift[1] = ss.y;
  }
}

// The whole sequential (scale, shift) recurrence of a video in ONE cooperative kernel (video_depth.py:216-252): window
// k is fitted on its slots 0 / 1 against (window 0's slot 0, the ALIGNED slot 12 of window k-1), and the aligned slot
// 12 is a function of window k-1's (scale, shift) -- a chain over all K windows that the multi-GPU driver has to walk
// once every rank's anchor frames (slots 0, 1, 12) are gathered.  As separate launches that is 4 kernels per window
// (sums, solve, table copy, re-alignment of the reference frame); here a grid of the same CTAs walks the chain with one
// grid-wide barrier per window: the aligned reference frame is never materialised (it is recomputed on the fly from
// the raw anchor and the previous (scale, shift), the same fmaf + clamp), every CTA solves the 2x2 system itself from
// the double-buffered partials.  Thread / CTA mapping, accumulation order and solve are those of lsq_sums_kernel /
// lsq_solve_kernel, so the table is bit-identical to the one WindowAligner computes window by window.
// anchors: [K][3][hw] raw fp32 (slots 0, 1, 12); table: [K][2]; partials: [2][gridDim.x * 4] doubles.
__global__ void __launch_bounds__(256) align_chain_kernel(const float* __restrict__ anchors, long long hw, int K,
                                                          float* __restrict__ table, double* __restrict__ partials) {
  cg::grid_group grid = cg::this_grid();
  __shared__ float2 ss_sh;
  const long long n = 2 * hw, n4 = n >> 2;
  const float* ref0 = anchors;                         // window 0, slot 0: never rescaled (copied unclamped, :222-225)
  float s_prev = 1.f, t_prev = 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) { table[0] = 1.f; table[1] = 0.f; }
  for (int k = 1; k < K; ++k) {
    const float* pred = anchors + static_cast<long long>(k) * 3 * hw;               // slots 0 and 1, contiguous
    const float* raw12 = anchors + (static_cast<long long>(k - 1) * 3 + 2) * hw;    // slot 12 of window k-1
    const bool vec = (hw & 3) == 0;   // every frame starts 16-byte aligned (cudaMalloc'd base)
    // target element e of the flat [2 hw] reference: e < hw: ref0[e]; else the aligned key frame of window k-1
    auto tgt = [&](long long e) -> float {
      if (e < hw) return ref0[e];
      const float x = raw12[e - hw];
      return k == 1 ? x : affine_clamp(x, s_prev, t_prev);
    };
    LsqAcc a = {0.f, 0.f, 0.f, 0.f};
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long e = 4 * i;
      const float4 p = ld_quad(pred, e, vec);
      float4 t;
      if (vec) {   // a quad never straddles the two reference frames
        if (e < hw) t = ld_quad(ref0, e, true);
        else {
          t = ld_quad(raw12, e - hw, true);
          if (k > 1) {
            t.x = affine_clamp(t.x, s_prev, t_prev); t.y = affine_clamp(t.y, s_prev, t_prev);
            t.z = affine_clamp(t.z, s_prev, t_prev); t.w = affine_clamp(t.w, s_prev, t_prev);
          }
        }
      } else {
        t = make_float4(tgt(e), tgt(e + 1), tgt(e + 2), tgt(e + 3));
      }
      lsq_acc(a, p.x, t.x); lsq_acc(a, p.y, t.y); lsq_acc(a, p.z, t.z); lsq_acc(a, p.w, t.w);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
      for (long long e = n4 << 2; e < n; ++e) lsq_acc(a, pred[e], tgt(e));
    double* part = partials + static_cast<size_t>(k & 1) * gridDim.x * 4;
    lsq_store_partial(a, part);
    __threadfence();
    grid.sync();
    if (threadIdx.x < 32) {
      const float2 ss = lsq_solve_warp(part, static_cast<int>(gridDim.x), n);
      if (threadIdx.x == 0) ss_sh = ss;
    }
    __syncthreads();
    s_prev = ss_sh.x;
    t_prev = ss_sh.y;
    if (blockIdx.x == 0 && threadIdx.x == 0) { table[2 * k] = s_prev; table[2 * k + 1] = t_prev; }
    __syncthreads();   // ss_sh is rewritten in the next round
  }
}

__global__ void __launch_bounds__(256) affine_clamp_blend_kernel(const float* __restrict__ x, const float* __restrict__ ss,
                                                                 const float* __restrict__ prev, const float* __restrict__ bw,
                                                                 float* __restrict__ out, int frames, long long hw) {
  const float s = ss[0], t = ss[1];
  const long long total = static_cast<long long>(frames) * hw;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = affine_clamp(x[i], s, t);
    if (bw) {
      const float w = bw[i / hw];
      v = fmaf(v, w, prev[i] * (1.f - w));
    }
    out[i] = v;
  }
}

static long long lsq_grid(long long n) {
  long long g = (n / 4 + 255) / 256;
  if (g > VDA_LSQ_MAX_PARTIALS) g = VDA_LSQ_MAX_PARTIALS;
  if (g < 1) g = 1;
  return g;
}

}  // namespace vda

using namespace vda;

extern "C" int vda_lsq_scale_shift(const float* pred, const float* target, int64_t n, float* scale_shift, double* scratch,
                                   void* stream) {
  VDA_CHECK(n > 0, "lsq: empty input");
  VDA_CHECK((reinterpret_cast<uintptr_t>(pred) & 15) == 0 && (reinterpret_cast<uintptr_t>(target) & 15) == 0,
            "lsq: inputs must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long g = lsq_grid(n);
  lsq_sums_kernel<<<static_cast<unsigned>(g), 256, 0, st>>>(pred, target, n, scratch);
  lsq_solve_kernel<<<1, 32, 0, st>>>(scratch, static_cast<int>(g), n, scale_shift);
  VDA_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vda_align_chain(const float* anchors, int n_windows, int64_t hw, int affine, float* table, double* scratch,
                               void* stream) {
  VDA_CHECK(n_windows > 0 && hw > 0, "align_chain: empty input");
  VDA_CHECK((reinterpret_cast<uintptr_t>(anchors) & 15) == 0, "align_chain: anchors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!affine || n_windows == 1) {    // metric model: identity alignment (metric_depth/.../video_depth.py:132)
    std::vector<float> ones(static_cast<size_t>(n_windows) * 2);
    for (int k = 0; k < n_windows; ++k) { ones[2 * k] = 1.f; ones[2 * k + 1] = 0.f; }
    VDA_CUDA(cudaMemcpyAsync(table, ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    VDA_CUDA(cudaStreamSynchronize(st));   // `ones` is a pageable temporary
    return 0;
  }
  const long long n = 2 * static_cast<long long>(hw);
  const long long g = lsq_grid(n);
  int per_sm = 0;
  VDA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, align_chain_kernel, 256, 0));
  VDA_CHECK(static_cast<long long>(per_sm) * sm_count() >= g, "align_chain: %lld CTAs cannot be co-resident (%d per SM)", g,
            per_sm);
  long long hw_l = hw;
  int K = n_windows;
  void* args[] = {(void*)&anchors, (void*)&hw_l, (void*)&K, (void*)&table, (void*)&scratch};
  VDA_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(align_chain_kernel), dim3(static_cast<unsigned>(g)),
                                       dim3(256), args, 0, st));
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
