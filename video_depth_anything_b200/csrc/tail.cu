// Fused depth-head tail (dpt_temporal.py:94-100 with dpt.py:118-124):
//
//   depth = relu( w2 . relu( conv3x3( bilinear_align_corners(o1 -> OH x OW) ) + b1 ) + b2 )      o1: [n, IH, IW, C] 16-bit
//
// Unfused, the upsampled C-channel map (69 MB per 518^2 frame at C = 128) is written once and then re-read nine
// times (once per tap) by the implicit-GEMM conv: the N = 32 conv is then bound by L2 -> smem traffic, not by the
// tensor pipe.  Here every CTA (persistent, 544 threads) owns 8 x 16 output-pixel tiles:
//   warps 1-12  producers: interpolate the 10 x 18 halo tile of the UPSAMPLED map straight from o1 (L1/L2 resident)
//               into shared memory, once per tile, in the no-swizzle K-major core-matrix layout
//               [halo row][8-channel chunk][halo column][8 ch]  (16-byte units);
//   warp 0      loads the whole 32 x 9C weight matrix once (TMA, SWIZZLE_128B) and issues the tcgen05.mma stream:
//               the nine taps are just nine start addresses into the same halo tile (A descriptor: LBO = chunk
//               stride, SBO = halo-row stride), 9 * C/16 UMMAs of 128 x 32 x 16 per tile, fp32 accumulators in TMEM
//               (two stages);
//   warps 13-16 epilogue: one thread per pixel, relu(acc + b1) . w2 + b2 -> relu -> fp32 depth.
// Out-of-image halo pixels are zeros (the conv's padding); interpolation arithmetic is identical to
// vda_bilinear_nhwc (fp32 weights, one rounding to 16 bit), so the result matches the unfused path.
#include "../../include/vda.h"
#include "common.cuh"

namespace vda {

namespace tl {
constexpr int TW = 8, TH = 16;            // output tile (pixels): 8 columns x 16 rows = 128 UMMA rows
constexpr int HW_ = TW + 2, HH = TH + 2;   // halo tile
constexpr int XU = 11;                    // 16-byte units per (halo row, chunk): 10 used + 1 pad (bank spread)
constexpr int NPROD = 12;                 // producer warps (the interpolation is ALU work: ~100 instructions per 16-byte unit)
constexpr int THREADS = 32 * (1 + NPROD + 4);
constexpr int N = 32;                     // conv output channels
}  // namespace tl

struct TailParams {
  const void* in;       // o1
  float* out;           // depth
  const float* bias;    // [32]
  const float* w2;      // [32]
  float b2;
  int n_img, IH, IW, OH, OW, C;
  int tiles_x, tiles_y, num_tiles;
  float sy, sx;
};

// no-swizzle K-major operand: 8-row x 16-byte core matrices; LBO = stride between the two K chunks of one MMA,
// SBO = stride between 8-row groups
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

template <typename T, int CC>
__global__ void __launch_bounds__(tl::THREADS, 1)
tail_fused_kernel(const __grid_constant__ CUtensorMap tmW, const TailParams p) {
  using namespace tl;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t b_full, a_full[2], a_empty[2], t_full[2], t_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[N], s_w2[N];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KC = CC >> 3;                                // 8-channel chunks (compile time: index math by shifts)
  const uint32_t chunk_stride = XU * 16;                     // LBO
  const uint32_t row_stride = static_cast<uint32_t>(KC) * chunk_stride;   // SBO: next halo row
  const uint32_t halo_bytes = HH * row_stride;
  constexpr int kblocks = 9 * (CC >> 6);                     // 64-wide weight tiles
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t offA = static_cast<uint32_t>(kblocks) * 4096u;   // weights first (1024-aligned 4 KB tiles)

  if (threadIdx.x < N) {
    s_bias[threadIdx.x] = p.bias[threadIdx.x];
    s_w2[threadIdx.x] = p.w2[threadIdx.x];
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    mbar_init(&b_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 32 * NPROD);
      mbar_init(&a_empty[s], 1);
      mbar_init(&t_full[s], 1);
      mbar_init(&t_empty[s], 128);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================================ weights + MMA issuer ================================
    if (elect_one()) {
      mbar_arrive_expect_tx(&b_full, static_cast<uint32_t>(kblocks) * 4096u);
      for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(smem_gen + kb * 4096, &tmW, &b_full, kb * 64, 0);
    }
    __syncwarp();
    mbar_wait(&b_full, 0);
    const uint32_t idesc = umma_idesc(H16<T>::kUmmaFmt, N);
    constexpr int ksteps = CC >> 4;                          // UMMAs per tap
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1u;
      mbar_wait(&t_empty[buf], ph ^ 1u);
      mbar_wait(&a_full[buf], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t abase = smem_base + offA + buf * halo_bytes;
        const uint32_t d_tmem = tmem_base + buf * N;
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - dy * 3;
          const uint32_t atap = abase + dy * row_stride + dx * 16;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t da = umma_desc_nosw(atap + 2 * ks * chunk_stride, chunk_stride, row_stride);
            const int kb = tap * (CC >> 6) + (ks >> 2);
            const uint64_t db = umma_desc_sw128(smem_base + kb * 4096) + 2u * (ks & 3);
            umma_f16(d_tmem, da, db, idesc, (tap | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&a_empty[buf]);
        umma_commit(&t_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp <= NPROD) {
    // ================================ producers: upsampled halo tile =====================
    const int ptid = threadIdx.x - 32;                       // 0 .. 32*NPROD-1
    const T* in = reinterpret_cast<const T*>(p.in);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1u;
      const int img = tile / per_img;
      const int rem = tile - img * per_img;
      const int Y0 = (rem / p.tiles_x) * TH - 1, X0 = (rem % p.tiles_x) * TW - 1;   // halo origin (output coords)
      const T* src = in + static_cast<long long>(img) * p.IH * p.IW * CC;
      mbar_wait(&a_empty[buf], ph ^ 1u);
      const uint32_t abase = smem_base + offA + buf * halo_bytes;
      // A unit is one halo pixel x 16 channels (two 8-channel chunks): the pixel's coordinates, interpolation weights and
      // tap offsets are computed once per 16 channels (they were half of a unit's instructions when a unit was 8
      // channels).  UB units per thread are in flight at once: all 8*UB neighbour loads are issued before any arithmetic
      // (one L1/L2 round trip per batch instead of one per unit)
      constexpr int UB = 2, KC2 = KC / 2;
      constexpr int units2 = HH * HW_ * KC2;
      for (int u0 = ptid; u0 < units2; u0 += 32 * NPROD * UB) {
        uint4 a[UB][4][2];
        float wy[UB], wx[UB];
        uint32_t dst[UB];
        int state[UB];                                         // 0: no unit, 1: zero (padding), 2: interpolate
#pragma unroll
        for (int b = 0; b < UB; ++b) {
          const int u = u0 + b * 32 * NPROD;
          state[b] = 0;
          wy[b] = wx[b] = 0.f;
          if (u < units2) {
            const int kc = 2 * (u % KC2);
            const int px = u / KC2;
            const int hy = px / HW_, hx = px - hy * HW_;
            const int Y = Y0 + hy, X = X0 + hx;
            dst[b] = abase + hy * row_stride + kc * chunk_stride + hx * 16;
            state[b] = 1;
            if (Y >= 0 && Y < p.OH && X >= 0 && X < p.OW) {
              state[b] = 2;
              const float fy = p.sy * Y, fx = p.sx * X;
              const int y0 = min(static_cast<int>(fy), p.IH - 1), x0 = min(static_cast<int>(fx), p.IW - 1);
              const int y1 = min(y0 + 1, p.IH - 1), x1 = min(x0 + 1, p.IW - 1);
              wy[b] = fy - y0;
              wx[b] = fx - x0;
              const T* bp = src + kc * 8;
              const T* t00 = bp + (y0 * p.IW + x0) * CC;
              const T* t01 = bp + (y0 * p.IW + x1) * CC;
              const T* t10 = bp + (y1 * p.IW + x0) * CC;
              const T* t11 = bp + (y1 * p.IW + x1) * CC;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                a[b][0][h] = *reinterpret_cast<const uint4*>(t00 + 8 * h);
                a[b][1][h] = *reinterpret_cast<const uint4*>(t01 + 8 * h);
                a[b][2][h] = *reinterpret_cast<const uint4*>(t10 + 8 * h);
                a[b][3][h] = *reinterpret_cast<const uint4*>(t11 + 8 * h);
              }
            }
          }
        }
#pragma unroll
        for (int b = 0; b < UB; ++b) {
          if (state[b] == 0) continue;
          const float ly = wy[b], lx = wx[b];
          const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint4 res = make_uint4(0u, 0u, 0u, 0u);
            if (state[b] == 2) {
              const uint32_t p00[4] = {a[b][0][h].x, a[b][0][h].y, a[b][0][h].z, a[b][0][h].w};
              const uint32_t p01[4] = {a[b][1][h].x, a[b][1][h].y, a[b][1][h].z, a[b][1][h].w};
              const uint32_t p10[4] = {a[b][2][h].x, a[b][2][h].y, a[b][2][h].z, a[b][2][h].w};
              const uint32_t p11[4] = {a[b][3][h].x, a[b][3][h].y, a[b][3][h].z, a[b][3][h].w};
              uint32_t r[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f00 = H16<T>::unpack2(p00[i]), f01 = H16<T>::unpack2(p01[i]);
                const float2 f10 = H16<T>::unpack2(p10[i]), f11 = H16<T>::unpack2(p11[i]);
                r[i] = H16<T>::pack2(w00 * f00.x + w01 * f01.x + w10 * f10.x + w11 * f11.x,
                                     w00 * f00.y + w01 * f01.y + w10 * f10.y + w11 * f11.y);
              }
              res = make_uint4(r[0], r[1], r[2], r[3]);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst[b] + h * chunk_stride), "r"(res.x), "r"(res.y),
                         "r"(res.z), "r"(res.w)
                         : "memory");
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the UMMA
      mbar_arrive(&a_full[buf]);
    }
  } else {
    // ================================ epilogue ===========================================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int ty = r >> 3, tx = r & 7;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1u;
      const int img = tile / per_img;
      const int rem = tile - img * per_img;
      const int Y = (rem / p.tiles_x) * TH + ty, X = (rem % p.tiles_x) * TW + tx;
      mbar_wait(&t_full[buf], ph);
      tc_fence_after();
      uint32_t rr[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * N, rr);
      tmem_ld_wait32(rr);
      tc_fence_before();
      mbar_arrive(&t_empty[buf]);
      float acc = p.b2;
#pragma unroll
      for (int i = 0; i < 32; ++i) acc = fmaf(fmaxf(__uint_as_float(rr[i]) + s_bias[i], 0.f), s_w2[i], acc);
      if (Y < p.OH && X < p.OW) p.out[(static_cast<long long>(img) * p.OH + Y) * p.OW + X] = fmaxf(acc, 0.f);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace vda

using namespace vda;

extern "C" int vda_tail_fused(const void* in, const void* w, const float* bias, const float* w2, float b2, float* out,
                              int n_img, int IH, int IW, int OH, int OW, int C, int dtype, void* stream) {
  VDA_CHECK(dtype == VDA_BF16 || dtype == VDA_FP16, "bad dtype %d", dtype);
  VDA_CHECK(C == 64 || C == 128, "tail: C (%d) must be 64 or 128 (output_conv1 channels padded to 64)", C);
  VDA_CHECK(n_img > 0 && IH > 0 && IW > 0 && OH > 0 && OW > 0, "tail: bad shape");
  VDA_CHECK((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
            "tail: in / w must be 16-byte aligned");
  CUtensorMap tm;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(9) * C, tl::N};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(9) * C * 2};
  cuuint32_t box[2] = {64, tl::N};
  if (make_tensor_map(&tm, dtype, w, 2, dims, strides, box)) return 1;
  TailParams p;
  p.in = in; p.out = out; p.bias = bias; p.w2 = w2; p.b2 = b2;
  p.n_img = n_img; p.IH = IH; p.IW = IW; p.OH = OH; p.OW = OW; p.C = C;
  p.tiles_x = (OW + tl::TW - 1) / tl::TW;
  p.tiles_y = (OH + tl::TH - 1) / tl::TH;
  p.num_tiles = n_img * p.tiles_x * p.tiles_y;
  p.sy = OH > 1 ? static_cast<float>(IH - 1) / (OH - 1) : 0.f;
  p.sx = OW > 1 ? static_cast<float>(IW - 1) / (OW - 1) : 0.f;
  const size_t halo = static_cast<size_t>(tl::HH) * (C / 8) * tl::XU * 16;
  const size_t smem = static_cast<size_t>(9) * (C / 64) * 4096 + 2 * halo + 1024;
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define TAIL_LAUNCH(TT, CCV)                                                                              \
  do {                                                                                                    \
    auto k = tail_fused_kernel<TT, CCV>;                                                                  \
    VDA_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(k), smem));   /* per (kernel, device) */     \
    k<<<grid, tl::THREADS, smem, st>>>(tm, p);                                                            \
  } while (0)
  if (dtype == VDA_BF16) {
    if (C == 128) TAIL_LAUNCH(__nv_bfloat16, 128); else TAIL_LAUNCH(__nv_bfloat16, 64);
  } else {
    if (C == 128) TAIL_LAUNCH(__half, 128); else TAIL_LAUNCH(__half, 64);
  }
#undef TAIL_LAUNCH
  VDA_CUDA(cudaGetLastError());
  return 0;
}
