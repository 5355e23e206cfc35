// Sequence evaluation on the device (benchmark/eval/eval.py:67-122 `eval_depthcrafter` with benchmark/eval/metric.py:
// abs_relative_difference :3-13, rmse_linear :30-41, delta1_acc :68-87): masked least-squares scale/shift of the
// predicted disparity against 1/gt over the whole sequence, clip, disparity -> depth, clip to the dataset's maximum
// depth, per-frame masked means, mean over the frames that have valid pixels.  The reference does this in float64
// (NumPy lstsq, float64 tensors), so all arithmetic here is double, with the reference's own divisions; five double
// divisions per pixel make the kernels FP64-pipe bound, not HBM bound: 1.8 ms for a 110 x 375 x 1242 sequence (one pass
// over pred + gt for the sums, one for the metrics).  Every reduction has a fixed order (per-thread strided sums,
// shared-memory tree, sequential combination of the block partials): results are bit-reproducible.
#include "../../include/vda.h"
#include "common.cuh"

namespace vda {

constexpr int EV_THREADS = 256;

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* dst) {
  __shared__ double sh[EV_THREADS];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    sh[threadIdx.x] = v[k];
    __syncthreads();
    for (int o = EV_THREADS / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) dst[k] = sh[0];
    __syncthreads();
  }
}

// partials[block][5] = sum p^2, sum p, n, sum p g, sum g   with p = max(pred, 1e-3), g = 1 / (gt + 1e-8)   (eval.py:86-93)
template <typename G>
__global__ void __launch_bounds__(EV_THREADS)
eval_lsq_partials_kernel(const float* __restrict__ pred, const G* __restrict__ gt, long long n, double max_depth,
                         double* __restrict__ partials) {
  double s[5] = {0, 0, 0, 0, 0};
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double g0 = static_cast<double>(gt[i]);
    if (g0 > 1e-3 && g0 < max_depth) {
      const double p = fmax(static_cast<double>(pred[i]), 1e-3);
      const double g = 1.0 / (g0 + 1e-8);
      s[0] += p * p; s[1] += p; s[2] += 1.0; s[3] += p * g; s[4] += g;
    }
  }
  block_reduce_store<5>(s, partials + static_cast<size_t>(blockIdx.x) * 5);
}

// normal equations of [p 1] [scale shift]^T = g  (the reference solves the same system with lstsq)
__global__ void eval_lsq_solve_kernel(const double* __restrict__ partials, int nparts, double* __restrict__ ss) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s[5] = {0, 0, 0, 0, 0};
  for (int b = 0; b < nparts; ++b)
    for (int k = 0; k < 5; ++k) s[k] += partials[b * 5 + k];
  const double det = s[0] * s[2] - s[1] * s[1];
  if (det != 0.0) {
    ss[0] = (s[2] * s[3] - s[1] * s[4]) / det;
    ss[1] = (s[0] * s[4] - s[1] * s[3]) / det;
  } else {
    ss[0] = 0.0;
    ss[1] = 0.0;
  }
}

// partials[frame][slab][4] = sum |d - gt| / gt, sum (d - gt)^2, #(max(d/gt, gt/d) < 1.25), n   over the valid pixels,
// d = clip(1 / clip(scale p + shift, 1e-3), 1e-3, max_depth)   (eval.py:94-104)
template <typename G>
__global__ void __launch_bounds__(EV_THREADS)
eval_metric_partials_kernel(const float* __restrict__ pred, const G* __restrict__ gt, long long hw,
                            const double* __restrict__ ss, double max_depth, double* __restrict__ partials) {
  const int frame = blockIdx.y;
  const double scale = ss[0], shift = ss[1];
  const float* pf = pred + static_cast<long long>(frame) * hw;
  const G* gf = gt + static_cast<long long>(frame) * hw;
  double s[4] = {0, 0, 0, 0};
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < hw;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double g = static_cast<double>(gf[i]);
    if (g > 1e-3 && g < max_depth) {
      const double a = fmax(scale * fmax(static_cast<double>(pf[i]), 1e-3) + shift, 1e-3);
      const double d = fmin(fmax(1.0 / a, 1e-3), max_depth);
      const double e = d - g;
      s[0] += fabs(e) / g;
      s[1] += e * e;
      s[2] += fmax(d / g, g / d) < 1.25 ? 1.0 : 0.0;   // the reference's own divisions (an equivalent product test could
                                                       // flip a pixel that sits within an ulp of the threshold)
      s[3] += 1.0;
    }
  }
  block_reduce_store<4>(s, partials + (static_cast<size_t>(frame) * gridDim.x + blockIdx.x) * 4);
}

// out[3] = mean over frames with valid pixels of (abs_rel, rmse, delta1)   (metric.py: per-frame sums / n, then .mean())
__global__ void eval_metric_finalize_kernel(const double* __restrict__ partials, int frames, int slabs,
                                            double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double acc[3] = {0, 0, 0};
  int kept = 0;
  for (int f = 0; f < frames; ++f) {
    double s[4] = {0, 0, 0, 0};
    for (int b = 0; b < slabs; ++b)
      for (int k = 0; k < 4; ++k) s[k] += partials[(static_cast<size_t>(f) * slabs + b) * 4 + k];
    if (s[3] > 0.0) {
      acc[0] += s[0] / s[3];
      acc[1] += sqrt(s[1] / s[3]);
      acc[2] += s[2] / s[3];
      ++kept;
    }
  }
  for (int k = 0; k < 3; ++k) out[k] = kept > 0 ? acc[k] / kept : 0.0;
}

}  // namespace vda

using namespace vda;

extern "C" int vda_eval_sequence(const float* pred, const void* gt, int gt_f64, int frames, int64_t hw, double max_depth,
                                 double* out, double* scale_shift, double* scratch, void* stream) {
  VDA_CHECK(frames > 0 && hw > 0 && max_depth > 0, "eval: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = static_cast<long long>(frames) * hw;
  int nparts = static_cast<int>((n + EV_THREADS - 1) / EV_THREADS);
  if (nparts > VDA_EVAL_LSQ_PARTIALS) nparts = VDA_EVAL_LSQ_PARTIALS;
  int slabs = static_cast<int>((hw + EV_THREADS - 1) / EV_THREADS);
  if (slabs > VDA_EVAL_SLABS) slabs = VDA_EVAL_SLABS;
  double* lsq_part = scratch;
  double* met_part = scratch + static_cast<size_t>(VDA_EVAL_LSQ_PARTIALS) * 5;
  if (gt_f64) {
    eval_lsq_partials_kernel<double><<<nparts, EV_THREADS, 0, st>>>(pred, static_cast<const double*>(gt), n, max_depth, lsq_part);
    eval_lsq_solve_kernel<<<1, 32, 0, st>>>(lsq_part, nparts, scale_shift);
    eval_metric_partials_kernel<double><<<dim3(slabs, frames), EV_THREADS, 0, st>>>(pred, static_cast<const double*>(gt), hw,
                                                                                  scale_shift, max_depth, met_part);
  } else {
    eval_lsq_partials_kernel<float><<<nparts, EV_THREADS, 0, st>>>(pred, static_cast<const float*>(gt), n, max_depth, lsq_part);
    eval_lsq_solve_kernel<<<1, 32, 0, st>>>(lsq_part, nparts, scale_shift);
    eval_metric_partials_kernel<float><<<dim3(slabs, frames), EV_THREADS, 0, st>>>(pred, static_cast<const float*>(gt), hw,
                                                                                 scale_shift, max_depth, met_part);
  }
  eval_metric_finalize_kernel<<<1, 32, 0, st>>>(met_part, frames, slabs, out);
  VDA_CUDA(cudaGetLastError());
  return 0;
}
