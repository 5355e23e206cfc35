// Fused attention kernels.
//
// (1) Spatial ViT attention (dinov2_layers/attention.py:49-62): softmax(q k^T / sqrt(64)) v per (frame, head),
//     N ~ 1370..2443 tokens, d = 64.  Flash-style: 128 queries per CTA (8 warps x 16 rows), 64-key tiles
//     double-buffered with cp.async, scores/probabilities never leave registers, online softmax in fp32.
//     Tensor-core math is warp-level mma.sync m16n8k16 in this revision (HMMA on sm_100a); the tcgen05/TMEM
//     version is the next step for this kernel.
// (2) Temporal attention (motion_module/motion_module.py:230-297, motion_module/attention.py:182-211):
//     32-frame sequences at every spatial position, 8 heads of d = C/8.  Bandwidth-bound; one CTA per
//     position, one warp per head, K/V staged in shared memory, fp32 math.
#include "../../include/vda.h"
#include "common.cuh"

namespace vda {

// ---------------------------------------------------------------------------------------------
// warp-level MMA helpers
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                                        uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// spatial attention
// ---------------------------------------------------------------------------------------------
constexpr int SA_BM = 128, SA_BN = 64, SA_D = 64, SA_THREADS = 256;

// byte offset of 16-byte chunk `c` (0..7) of row `r` in a [rows][64 x 16-bit] tile, XOR-swizzled
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return static_cast<uint32_t>((r * 8 + (c ^ (r & 7))) * 16); }

template <typename T>
__global__ void __launch_bounds__(SA_THREADS)
spatial_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N, int heads) {
  __shared__ __align__(128) uint8_t sQ[SA_BM * SA_D * 2];
  __shared__ __align__(128) uint8_t sK[2][SA_BN * SA_D * 2];
  __shared__ __align__(128) uint8_t sV[2][SA_BN * SA_D * 2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * SA_BM;
  const int head = blockIdx.y, frame = blockIdx.z;
  const long long row_stride = 3LL * heads * SA_D;
  const T* base = qkv + static_cast<long long>(frame) * N * row_stride + head * SA_D;
  const T* gQ = base;
  const T* gK = base + heads * SA_D;
  const T* gV = base + 2 * heads * SA_D;
  const uint32_t sQa = smem_u32(sQ);
  const uint32_t sKa[2] = {smem_u32(sK[0]), smem_u32(sK[1])};
  const uint32_t sVa[2] = {smem_u32(sV[0]), smem_u32(sV[1])};

  // Q tile: 128 rows x 8 chunks
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * SA_THREADS;
    const int r = idx >> 3, c = idx & 7;
    const int gr = min(q0 + r, N - 1);
    cp_async16(sQa + sw_off(r, c), gQ + gr * row_stride + c * 8);
  }
  auto load_kv = [&](int tile, int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * SA_THREADS;
      const int r = idx >> 3, c = idx & 7;
      const int gr = min(tile * SA_BN + r, N - 1);
      cp_async16(sKa[buf] + sw_off(r, c), gK + gr * row_stride + c * 8);
      cp_async16(sVa[buf] + sw_off(r, c), gV + gr * row_stride + c * 8);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  const int ntiles = (N + SA_BN - 1) / SA_BN;
  const float sc = 0.125f * 1.4426950408889634f;  // d^-0.5 * log2(e)
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  uint32_t qf[4][4];

  for (int j = 0; j < ntiles; ++j) {
    const int buf = j & 1;
    cp_async_wait<0>();
    __syncthreads();
    if (j + 1 < ntiles) {
      load_kv(j + 1, buf ^ 1);
      cp_async_commit();
    }
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int c = ks * 2 + (lane >> 4);
        ldsm_x4(qf[ks], sQa + sw_off(r, c));
      }
    }
    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[i][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kb[4];
        const int r = np * 16 + (lane & 7) + 8 * (lane >> 4);
        const int c = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(kb, sKa[buf] + sw_off(r, c));
        mma16816<T>(s[2 * np], qf[ks], kb[0], kb[1]);
        mma16816<T>(s[2 * np + 1], qf[ks], kb[2], kb[3]);
      }
    }
    // ---- mask keys beyond N (last tile only) ----
    if (j == ntiles - 1) {
      const int kbase = j * SA_BN + 2 * (lane & 3);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kbase + 8 * i + (e & 1) >= N) s[i][e] = -INFINITY;
    }
    // ---- online softmax (rows g and g+8 of this warp's 16) ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
    }
    float alpha[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      const float m_new = fmaxf(m_run[h], mx[h]);
      alpha[h] = exp2f((m_run[h] - m_new) * sc);
      m_run[h] = m_new;
      l_run[h] *= alpha[h];
    }
    const float mb0 = m_run[0] * sc, mb1 = m_run[1] * sc;
    uint32_t pf[4][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = exp2f(fmaf(s[i][0], sc, -mb0)), p1 = exp2f(fmaf(s[i][1], sc, -mb0));
      const float p2 = exp2f(fmaf(s[i][2], sc, -mb1)), p3 = exp2f(fmaf(s[i][3], sc, -mb1));
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      pf[i >> 1][(i & 1) * 2 + 0] = H16<T>::pack2(p0, p1);
      pf[i >> 1][(i & 1) * 2 + 1] = H16<T>::pack2(p2, p3);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
      o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vb[4];
        const int r = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int c = dp * 2 + (lane >> 4);
        ldsm_x4_t(vb, sVa[buf] + sw_off(r, c));
        mma16816<T>(o[2 * dp], pf[kk], vb[0], vb[1]);
        mma16816<T>(o[2 * dp + 1], pf[kk], vb[2], vb[3]);
      }
    }
  }
  // ---- finalize ----
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int g = lane >> 2, t = lane & 3;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  T* obase = out + static_cast<long long>(frame) * N * heads * SA_D + head * SA_D;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int col = i * 8 + 2 * t;
    if (r0 < N)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(r0) * heads * SA_D + col) =
          H16<T>::pack2(o[i][0] * inv0, o[i][1] * inv0);
    if (r1 < N)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(r1) * heads * SA_D + col) =
          H16<T>::pack2(o[i][2] * inv1, o[i][3] * inv1);
  }
}

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

extern "C" int vda_attention_spatial(const void* qkv, void* out, int frames, int N, int heads, int dtype,
                                     void* stream) {
  VDA_CHECK(frames > 0 && N > 0 && heads > 0, "bad attention shape");
  VDA_CHECK((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
            "qkv/out must be 16-byte aligned");
  dim3 grid((N + SA_BM - 1) / SA_BM, heads, frames);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == VDA_BF16)
    spatial_attention_kernel<__nv_bfloat16><<<grid, SA_THREADS, 0, st>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), N, heads);
  else
    spatial_attention_kernel<__half><<<grid, SA_THREADS, 0, st>>>(static_cast<const __half*>(qkv),
                                                                  static_cast<__half*>(out), N, heads);
  VDA_CUDA(cudaGetLastError());
  return 0;
}

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
