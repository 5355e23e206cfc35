// Bandwidth-bound normalisation kernels: LayerNorm (warp per row, fp32 statistics, 128-bit accesses) and
// GroupNorm over NHWC activations (per-slab shifted (mean, M2) partials, fixed-order Chan merge, then apply; no atomics).
#include "../../include/vda.h"
#include "common.cuh"

namespace vda {

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (C <= 1024), two-pass statistics in fp32.
// ---------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 8;  // float4 per lane -> C <= 32*4*8 = 1024

template <typename T, bool IN_F32>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ in, T* __restrict__ out, const float* __restrict__ w,
                 const float* __restrict__ b, float eps, long long rows, int C, int drop_group,
                 const float* __restrict__ pe, int pe_rows_per_frame, int pe_frames) {
  pdl_trigger();   // a PDL-launched successor (GEMM) may be scheduled as this grid drains
  const int lane = threadIdx.x & 31;
  // rows are visited last-to-first: the producer (a GEMM epilogue walking the row tiles upwards) wrote the highest
  // rows last, so they are still in the 126 MB L2, and the consumer GEMM starts at row 0, which this kernel writes last
  const long long row = rows - 1 - (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (row < 0) return;
  long long orow = row;
  if (drop_group > 0) {
    if (row % drop_group == 0) return;          // cls token: not consumed by the head (use_clstoken=False)
    orow = row - row / drop_group - 1;
  }
  const int nv = C >> 2;                          // float4 groups per row
  float4 v[LN_MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int g = lane + 32 * i;
    if (g < nv) {
      if (IN_F32) {
        v[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + row * C)[g];
      } else {
        const uint2 u = reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(in) + row * C)[g];
        const float2 a = H16<T>::unpack2(u.x), c = H16<T>::unpack2(u.y);
        v[i] = make_float4(a.x, a.y, c.x, c.y);
      }
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(sum) / C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int g = lane + 32 * i;
    if (g < nv) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      sq += dx * dx + dy * dy + dz * dz + dw * dw;
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / C + eps);
  const float* perow = nullptr;
  if (pe) perow = pe + static_cast<long long>((row / pe_rows_per_frame) % pe_frames) * C;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int g = lane + 32 * i;
    if (g < nv) {
      const float4 ww = reinterpret_cast<const float4*>(w)[g];
      const float4 bb = reinterpret_cast<const float4*>(b)[g];
      float4 r;
      r.x = (v[i].x - mean) * rstd * ww.x + bb.x;
      r.y = (v[i].y - mean) * rstd * ww.y + bb.y;
      r.z = (v[i].z - mean) * rstd * ww.z + bb.z;
      r.w = (v[i].w - mean) * rstd * ww.w + bb.w;
      if (perow) {
        const float4 pp = reinterpret_cast<const float4*>(perow)[g];
        r.x += pp.x; r.y += pp.y; r.z += pp.z; r.w += pp.w;
      }
      uint2 u;
      u.x = H16<T>::pack2(r.x, r.y);
      u.y = H16<T>::pack2(r.z, r.w);
      reinterpret_cast<uint2*>(out + orow * C)[g] = u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row statistics + 16-bit copy of an fp32 row matrix, in the layout the residual GEMM epilogue (gemm.cu SPEC 4) writes:
// per row, `parts` partials (mean, M2) over `part_cols` consecutive columns each.  Used once per forward for the token
// matrix after patch embedding; every later LayerNorm input of the encoder comes out of a proj / fc2 epilogue.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
rowstats_cast_kernel(const float* __restrict__ in, T* __restrict__ out16, float2* __restrict__ stats, long long rows,
                     int C, int parts, int part_cols) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* src = in + row * C;
  for (int pt = 0; pt < parts; ++pt) {
    const int c0 = pt * part_cols;
    const float shift = src[c0];
    float s1 = 0.f, s2 = 0.f;
    for (int c = 4 * lane; c < part_cols; c += 128) {        // part_cols % 4 == 0
      const float4 v = *reinterpret_cast<const float4*>(src + c0 + c);
      const float dx = v.x - shift, dy = v.y - shift, dz = v.z - shift, dw = v.w - shift;
      s1 += (dx + dy) + (dz + dw);
      s2 += dx * dx + dy * dy + dz * dz + dw * dw;
      uint2 u;
      u.x = H16<T>::pack2(v.x, v.y);
      u.y = H16<T>::pack2(v.z, v.w);
      *reinterpret_cast<uint2*>(out16 + row * C + c0 + c) = u;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      const float md = s1 / part_cols;
      stats[row * parts + pt] = make_float2(shift + md, fmaxf(s2 - s1 * md, 0.f));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm over [frames, hw, C] (NHWC).  Pass 1: every CTA reduces a slab of pixels of one frame to per-group
// (mean, M2 = sum of squared deviations) partials; pass 1b merges the slabs of a frame in a fixed order with Chan's
// parallel-variance update (no atomics anywhere: results are bit-reproducible run to run); pass 2: normalise + affine.
// The slab sums are taken of d = x - K_g with K_g = the group's first channel at the slab's first pixel, so that
// M2 = sum d^2 - (sum d)^2 / n cancels against (mean - K_g)^2 ~ var instead of mean^2: a group whose |mean| is far
// above its standard deviation (ReLU'd activations of real checkpoints) keeps its variance (round 1 used
// E[x^2] - mean^2 in fp32).
// stats layout: partials [frames][slabs][groups][2] = (mean, M2), then totals [frames][groups][2] = (mean, var).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const T* __restrict__ in, float* __restrict__ partials, int hw, int C, int groups,
                       int pix_per_cta) {
  // blockDim.x is a multiple of vecs = C/8, so a thread always visits the same 8-channel column and can
  // keep its four channel-pair partial sums in registers.
  __shared__ float sh_s[256][4], sh_q[256][4];
  const int frame = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  const int cpg = C / groups;
  const int vecs = C >> 3;
  const long long total = static_cast<long long>(p1 - p0) * vecs;
  const T* base = in + (static_cast<long long>(frame) * hw + p0) * C;
  // shift of each of this thread's channel pairs: first channel of the pair's group at the slab's first pixel
  const int c0t = (threadIdx.x % vecs) * 8;
  float kk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) kk[i] = H16<T>::to_f(base[((c0t + 2 * i) / cpg) * cpg]);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(base + idx * 8);
    const uint32_t wds[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = H16<T>::unpack2(wds[i]);
      const float dx = f.x - kk[i], dy = f.y - kk[i];
      s[i] += dx + dy;
      q[i] += dx * dx + dy * dy;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { sh_s[threadIdx.x][i] = s[i]; sh_q[threadIdx.x][i] = q[i]; }
  __syncthreads();
  if (threadIdx.x < groups) {
    // thread g sums, in thread order, every channel pair that belongs to group g (cpg is even: a pair never
    // straddles groups; all of them were shifted by the same K_g)
    const int g = threadIdx.x;
    float ts = 0.f, tq = 0.f;
    for (int t = 0; t < static_cast<int>(blockDim.x); ++t) {
      const int c0 = (t % vecs) * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if ((c0 + 2 * i) / cpg == g) { ts += sh_s[t][i]; tq += sh_q[t][i]; }
    }
    const float n = static_cast<float>(p1 - p0) * cpg;
    const float kg = H16<T>::to_f(base[g * cpg]);
    const float md = ts / n;                              // mean of d
    float* dst = partials + ((static_cast<size_t>(frame) * gridDim.x + blockIdx.x) * groups + g) * 2;
    dst[0] = kg + md;
    dst[1] = fmaxf(tq - ts * md, 0.f);
  }
}

__global__ void groupnorm_reduce_kernel(const float* __restrict__ partials, float* __restrict__ totals, int slabs,
                                        int groups, int hw, int pix_per_cta, int cpg) {
  const int frame = blockIdx.x, g = threadIdx.x;
  if (g >= groups) return;
  // Chan et al.: merge (n, mean, M2) of the slabs in slab order
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int sl = 0; sl < slabs; ++sl) {
    const float* src = partials + ((static_cast<size_t>(frame) * slabs + sl) * groups + g) * 2;
    const float nb = static_cast<float>(min(pix_per_cta, hw - sl * pix_per_cta)) * cpg;
    const float delta = src[0] - mean;
    const float nn = n + nb;
    mean += delta * (nb / nn);
    m2 += src[1] + delta * delta * (n * nb / nn);
    n = nn;
  }
  totals[(frame * groups + g) * 2] = mean;
  totals[(frame * groups + g) * 2 + 1] = m2 / n;
}

template <typename T>
__global__ void __launch_bounds__(256)
groupnorm_apply_kernel(const T* __restrict__ in, T* __restrict__ out, const float* __restrict__ stats,
                       const float* __restrict__ w, const float* __restrict__ b, float eps, int frames, int hw, int C,
                       int groups) {
  const int vecs = C >> 3;
  const long long total = static_cast<long long>(frames) * hw * vecs;
  const int cpg = C / groups;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vcol = static_cast<int>(idx % vecs);
    const int frame = static_cast<int>(idx / (static_cast<long long>(hw) * vecs));
    const int c0 = vcol * 8;
    const uint4 u = *reinterpret_cast<const uint4*>(in + idx * 8);
    const uint32_t wds[4] = {u.x, u.y, u.z, u.w};
    uint32_t res[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + 2 * i;
      const int g = c / cpg;     // cpg is even, so both channels of the pair share a group
      const float mean = stats[(frame * groups + g) * 2], var = stats[(frame * groups + g) * 2 + 1];
      const float rstd = rsqrtf(var + eps);
      const float2 f = H16<T>::unpack2(wds[i]);
      res[i] = H16<T>::pack2((f.x - mean) * rstd * w[c] + b[c], (f.y - mean) * rstd * w[c + 1] + b[c + 1]);
    }
    *reinterpret_cast<uint4*>(out + idx * 8) = make_uint4(res[0], res[1], res[2], res[3]);
  }
}

}  // namespace vda

using namespace vda;

extern "C" int vda_layernorm(const void* in, int in_f32, void* out, const float* w, const float* b, float eps,
                             int64_t rows, int C, int dtype, int drop_group, const float* pe, int pe_rows_per_frame,
                             int pe_frames, void* stream) {
  VDA_CHECK(C % 4 == 0 && C <= 128 * LN_MAXV, "LayerNorm: C (%d) must be a multiple of 4 and <= %d", C, 128 * LN_MAXV);
  VDA_CHECK(rows > 0, "LayerNorm: no rows");
  VDA_CHECK(!pe || (pe_rows_per_frame > 0 && pe_frames > 0), "LayerNorm: bad positional-encoding arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
#define LN_LAUNCH(TT, F32)                                                                              \
  layernorm_kernel<TT, F32><<<grid, wpb * 32, 0, st>>>(in, static_cast<TT*>(out), w, b, eps, rows, C,   \
                                                       drop_group, pe, pe_rows_per_frame, pe_frames)
  if (dtype == VDA_BF16) {
    if (in_f32) LN_LAUNCH(__nv_bfloat16, true); else LN_LAUNCH(__nv_bfloat16, false);
  } else {
    if (in_f32) LN_LAUNCH(__half, true); else LN_LAUNCH(__half, false);
  }
#undef LN_LAUNCH
  VDA_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vda_rowstats_cast(const float* in, void* out16, float* stats, int64_t rows, int C, int parts, int part_cols,
                                 int dtype, void* stream) {
  VDA_CHECK(rows > 0 && parts > 0 && part_cols > 0 && parts * part_cols == C && part_cols % 4 == 0,
            "rowstats_cast: bad layout (C=%d parts=%d part_cols=%d)", C, parts, part_cols);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
  if (dtype == VDA_BF16)
    rowstats_cast_kernel<__nv_bfloat16><<<grid, wpb * 32, 0, st>>>(in, static_cast<__nv_bfloat16*>(out16),
                                                                  reinterpret_cast<float2*>(stats), rows, C, parts, part_cols);
  else
    rowstats_cast_kernel<__half><<<grid, wpb * 32, 0, st>>>(in, static_cast<__half*>(out16), reinterpret_cast<float2*>(stats),
                                                           rows, C, parts, part_cols);
  VDA_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vda_groupnorm(const void* in, void* out, const float* w, const float* b, float eps, int frames, int hw,
                             int C, int groups, float* stats, int dtype, void* stream) {
  VDA_CHECK(groups > 0 && groups <= 64 && C % groups == 0 && C % 8 == 0 && (C / groups) % 2 == 0,
            "GroupNorm: unsupported C=%d groups=%d", C, groups);
  const int vecs = C / 8;
  VDA_CHECK(vecs <= 256, "GroupNorm: C (%d) too large", C);
  const int sthreads = (256 / vecs) * vecs;   // multiple of vecs (see groupnorm_stats_kernel)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // ~ 4 CTAs per SM worth of slabs per launch
  int slabs = (148 * 4 + frames - 1) / frames;
  if (slabs < 1) slabs = 1;
  int pix_per_cta = (hw + slabs - 1) / slabs;
  if (pix_per_cta < 8) pix_per_cta = 8;
  slabs = (hw + pix_per_cta - 1) / pix_per_cta;
  dim3 grid(slabs, frames);
  float* totals = stats + static_cast<size_t>(frames) * slabs * groups * 2;   // fits VDA_GN_STATS_FLOATS
  const long long total = static_cast<long long>(frames) * hw * (C / 8);
  unsigned g2 = static_cast<unsigned>((total + 255) / 256);
  if (g2 > 148u * 16u) g2 = 148u * 16u;
  if (dtype == VDA_BF16) {
    groupnorm_stats_kernel<__nv_bfloat16><<<grid, sthreads, 0, st>>>(static_cast<const __nv_bfloat16*>(in), stats, hw, C, groups, pix_per_cta);
    groupnorm_reduce_kernel<<<frames, 64, 0, st>>>(stats, totals, slabs, groups, hw, pix_per_cta, C / groups);
    groupnorm_apply_kernel<__nv_bfloat16><<<g2, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out),
                                                              totals, w, b, eps, frames, hw, C, groups);
  } else {
    groupnorm_stats_kernel<__half><<<grid, sthreads, 0, st>>>(static_cast<const __half*>(in), stats, hw, C, groups, pix_per_cta);
    groupnorm_reduce_kernel<<<frames, 64, 0, st>>>(stats, totals, slabs, groups, hw, pix_per_cta, C / groups);
    groupnorm_apply_kernel<__half><<<g2, 256, 0, st>>>(static_cast<const __half*>(in), static_cast<__half*>(out), totals, w, b,
                                                       eps, frames, hw, C, groups);
  }
  VDA_CUDA(cudaGetLastError());
  return 0;
}
