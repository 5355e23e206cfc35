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

// ---------------------------------------------------------------------------------------------
// Temporal alignment error (benchmark/eval/eval_tae.py:60-107 `tae_torch`, driven by eval_TAE :109-213): the aligned
// depth of frame s is un-projected with the intrinsics, moved to frame d by the relative pose, projected back, rounded
// to pixel indices and SCATTERED into an empty map of frame d (`depth_proj[valid_Y, valid_X] = valid_Z`); the error is
// the masked mean of |depth_d - depth_proj| / depth_d over pixels where both are positive.
// The scatter has duplicates: several source pixels land on one target pixel.  The reference's assignment keeps the
// LAST writer in source order (NumPy / single-threaded index_put semantics), so the winner of a target pixel is the
// source pixel with the largest flat index: pass 1 takes atomicMax of the source index per target pixel (integer,
// order-independent, exact), pass 2 recomputes the winner's projected depth and reduces in a fixed order.
// All arithmetic in double like the reference (float64 tensors); one job = one ordered (source, target) frame pair.
// params per job: [0..8] R (row-major), [9..11] t, [12..15] fx, fy, cx, cy.
// ---------------------------------------------------------------------------------------------
struct TaePoint { long long ix, iy; double z; bool ok; };
__device__ __forceinline__ TaePoint tae_project(const double* __restrict__ prm, double depth, int x, int y, int W, int H) {
  const double fx = prm[12], fy = prm[13], cx = prm[14], cy = prm[15];
  // eval_tae.py:72-75   X = (xx - cx) * depth1 / fx
  const double X = (static_cast<double>(x) - cx) * depth / fx;
  const double Y = (static_cast<double>(y) - cy) * depth / fy;
  const double Z = depth;
  // :80   points3d @ R.T + T
  const double xw = X * prm[0] + Y * prm[1] + Z * prm[2] + prm[9];
  const double yw = X * prm[3] + Y * prm[4] + Z * prm[5] + prm[10];
  const double zw = X * prm[6] + Y * prm[7] + Z * prm[8] + prm[11];
  // :83-88   project, round half to even (torch.round), to integer
  const double xp = rint(xw * fx / zw + cx);
  const double yp = rint(yw * fy / zw + cy);
  TaePoint r;
  r.z = zw;
  r.ok = xp >= 0.0 && xp < static_cast<double>(W) && yp >= 0.0 && yp < static_cast<double>(H);   // false for NaN / inf
  r.ix = r.ok ? static_cast<long long>(xp) : 0;
  r.iy = r.ok ? static_cast<long long>(yp) : 0;
  return r;
}

__global__ void __launch_bounds__(EV_THREADS)
tae_scatter_kernel(const double* __restrict__ depth, const double* __restrict__ params, const int* __restrict__ src_idx,
                   int H, int W, int* __restrict__ winners) {
  const int job = blockIdx.y;
  const long long hw = static_cast<long long>(H) * W;
  const double* d = depth + static_cast<long long>(src_idx[job]) * hw;
  const double* prm = params + static_cast<size_t>(job) * 16;
  int* win = winners + static_cast<long long>(job) * hw;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < hw;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>(i / W), x = static_cast<int>(i - static_cast<long long>(y) * W);
    const TaePoint pt = tae_project(prm, d[i], x, y, W, H);
    if (pt.ok) atomicMax(win + pt.iy * W + pt.ix, static_cast<int>(i));
  }
}

// partials[job][block][2] = (sum |depth_d - proj| / depth_d, n) over the valid target pixels
__global__ void __launch_bounds__(EV_THREADS)
tae_error_kernel(const double* __restrict__ depth, const uint8_t* __restrict__ masks, const double* __restrict__ params,
                 const int* __restrict__ src_idx, const int* __restrict__ dst_idx, int H, int W,
                 const int* __restrict__ winners, double* __restrict__ partials) {
  const int job = blockIdx.y;
  const long long hw = static_cast<long long>(H) * W;
  const double* ds = depth + static_cast<long long>(src_idx[job]) * hw;
  const double* dd = depth + static_cast<long long>(dst_idx[job]) * hw;
  const uint8_t* mk = masks ? masks + static_cast<long long>(dst_idx[job]) * hw : nullptr;
  const double* prm = params + static_cast<size_t>(job) * 16;
  const int* win = winners + static_cast<long long>(job) * hw;
  double s[2] = {0, 0};
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < hw;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = win[i];
    if (w < 0) continue;                                   // depth_proj stays 0 here (:95)
    const int y = w / W, x = w - y * W;
    const double proj = tae_project(prm, ds[w], x, y, W, H).z;
    const double g = dd[i];
    if (proj > 0.0 && g > 0.0 && (!mk || mk[i])) {         // :103
      s[0] += fabs(g - proj) / g;                          // compute_errors_torch(gt = depth2, pred = depth_proj)
      s[1] += 1.0;
    }
  }
  block_reduce_store<2>(s, partials + (static_cast<size_t>(job) * gridDim.x + blockIdx.x) * 2);
}

__global__ void tae_finalize_kernel(const double* __restrict__ partials, int jobs, int slabs, double* __restrict__ out) {
  const int job = blockIdx.x * blockDim.x + threadIdx.x;
  if (job >= jobs) return;
  double s0 = 0, s1 = 0;
  for (int b = 0; b < slabs; ++b) {
    s0 += partials[(static_cast<size_t>(job) * slabs + b) * 2];
    s1 += partials[(static_cast<size_t>(job) * slabs + b) * 2 + 1];
  }
  out[job] = s1 > 0.0 ? s0 / s1 : 0.0;                     // `return 0` when nothing is valid (:91-92, :104-105)
}

// aligned depth of the whole sequence (eval_tae.py:152-160): clip(1 / clip(scale * clip(pred, 1e-3) + shift, 1e-3), 1e-3, max)
__global__ void __launch_bounds__(EV_THREADS)
eval_aligned_depth_kernel(const float* __restrict__ pred, const double* __restrict__ ss, double max_depth, long long n,
                          double* __restrict__ out) {
  const double scale = ss[0], shift = ss[1];
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double a = fmax(scale * fmax(static_cast<double>(pred[i]), 1e-3) + shift, 1e-3);
    out[i] = fmin(fmax(1.0 / a, 1e-3), max_depth);
  }
}

}  // namespace vda

using namespace vda;

extern "C" int vda_eval_aligned_depth(const float* pred, const double* scale_shift, double max_depth, int64_t n, double* out,
                                      void* stream) {
  VDA_CHECK(n > 0 && max_depth > 0, "aligned depth: bad arguments");
  long long g = (n + EV_THREADS - 1) / EV_THREADS;
  if (g > 148 * 16) g = 148 * 16;
  eval_aligned_depth_kernel<<<static_cast<unsigned>(g), EV_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, scale_shift, max_depth, n, out);
  VDA_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vda_eval_tae(const double* depth, const uint8_t* masks, const double* params, const int32_t* src_idx,
                            const int32_t* dst_idx, int jobs, int H, int W, int32_t* winners, double* partials, double* out,
                            void* stream) {
  VDA_CHECK(jobs > 0 && H > 0 && W > 0 && static_cast<long long>(H) * W < (1ll << 31), "tae: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long hw = static_cast<long long>(H) * W;
  int slabs = static_cast<int>((hw + EV_THREADS - 1) / EV_THREADS);
  if (slabs > VDA_EVAL_SLABS) slabs = VDA_EVAL_SLABS;
  VDA_CUDA(cudaMemsetAsync(winners, 0xFF, static_cast<size_t>(jobs) * hw * sizeof(int32_t), st));   // -1: no writer
  int sg = static_cast<int>((hw + EV_THREADS - 1) / EV_THREADS);
  if (sg > 148 * 8) sg = 148 * 8;
  tae_scatter_kernel<<<dim3(sg, jobs), EV_THREADS, 0, st>>>(depth, params, src_idx, H, W, winners);
  tae_error_kernel<<<dim3(slabs, jobs), EV_THREADS, 0, st>>>(depth, masks, params, src_idx, dst_idx, H, W, winners, partials);
  tae_finalize_kernel<<<(jobs + 63) / 64, 64, 0, st>>>(partials, jobs, slabs, out);
  VDA_CUDA(cudaGetLastError());
  return 0;
}

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
