// Temporal attention (the spatial ViT attention lives in attention_spatial.cu).
//
// motion_module/motion_module.py:230-297 with motion_module/attention.py:182-211: at every spatial position an
// independent 32-frame sequence, 8 heads of d = C/8:  out = softmax(q k^T d^-1/2) v.
//
// The op is HBM-bound (4 * T*hw*C 16-bit elements moved for 4*T*T*d FLOP per (position, head): AI = T/2 = 16), so
// the kernel is organised around memory: one CTA per (position, group of heads), the q|k|v row segments of the
// 32 frames are copied with 16-byte cp.async into XOR-swizzled shared-memory tiles (512 B..1 KB contiguous per
// segment), several CTAs per SM keep the loads of the next CTA in flight while this one computes.  One warp per
// head: S = Q K^T (32x32) and O = P V on warp-level tensor-core MMAs (m16n8k16, fp32 accumulate; 32x32 tiles are
// far too small for a 128-row tcgen05 tile and the tensor work is < 10% of the HBM time either way), softmax in
// fp32 registers with quad shuffles, O staged through the (dead) Q tile so the global stores are 16-byte and
// row-contiguous.  Head dims that are not a multiple of 16 (ViT-S: 24, 48, 8) are zero-padded in shared memory.
#include "../../include/vda.h"
#include "common.cuh"

namespace vda {

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
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// One head tile: 32 rows (frames) x DHP columns of 16-bit, row pitch DHP*2 bytes, 16-byte chunks XOR-swizzled so
// that the 8 rows of an ldmatrix 8x8 block fall into 8 different bank groups for every pitch.
// ---------------------------------------------------------------------------------------------
template <int DHP>
__device__ __forceinline__ uint32_t ta_off(int row, int chunk) {
  constexpr int CH = DHP / 8;                       // chunks per row: 2, 4, 8 or 16
  constexpr int RPL = CH >= 8 ? 1 : 8 / CH;         // rows per 128-byte line
  constexpr int MASK = CH >= 8 ? 7 : CH - 1;
  return static_cast<uint32_t>(row * (DHP * 2) + ((chunk ^ ((row / RPL) & MASK)) << 4));
}

template <typename T, int DHP>
__global__ void __launch_bounds__(256)
temporal_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Tn, int hw, int C, int heads, int hpc) {
  extern __shared__ __align__(128) uint8_t sm_t[];   // [hpc][3 (q|k|v)][32][DHP] 16-bit
  constexpr int TILE = 32 * DHP * 2;
  constexpr int CH = DHP / 8;
  const int pos = blockIdx.x;
  const int dh = C / heads;
  const int h0 = blockIdx.y * hpc;
  const int dch = dh / 8;                            // valid 16-byte chunks per head row
  const uint32_t sbase = smem_u32(sm_t);

  // ---- cooperative load: (head, which, frame, chunk); rows >= Tn and pad chunks are zero-filled ----
  const int total = hpc * 3 * 32 * CH;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int c = idx % CH;
    int rest = idx / CH;
    const int h = rest % hpc;                        // consecutive threads: chunks, then heads (contiguous in global)
    rest /= hpc;
    const int f = rest % 32, which = rest / 32;
    const uint32_t dst = sbase + (h * 3 + which) * TILE + ta_off<DHP>(f, c);
    if (f < Tn && c < dch) {
      const T* src = qkv + (static_cast<long long>(f) * hw + pos) * (3LL * C) + which * C + (h0 + h) * dh + c * 8;
      cp_async16(dst, src);
    } else {
      asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0) : "memory");
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= hpc) return;
  const uint32_t sQ = sbase + (warp * 3 + 0) * TILE, sK = sQ + TILE, sV = sK + TILE;
  const int g = lane >> 2, t4 = lane & 3;

  // ---- S = Q K^T : 32 x 32, fp32 ----
  float s[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < DHP / 16; ++ks) {
    uint32_t a[2][4], b[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) ldsm_x4(a[mt], sQ + ta_off<DHP>(mt * 16 + (lane & 15), ks * 2 + (lane >> 4)));
#pragma unroll
    for (int np = 0; np < 2; ++np)
      ldsm_x4(b[np], sK + ta_off<DHP>(np * 16 + (lane & 7) + 8 * (lane >> 4), ks * 2 + ((lane >> 3) & 1)));
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        mma16816<T>(s[mt][2 * np], a[mt], b[np][0], b[np][1]);
        mma16816<T>(s[mt][2 * np + 1], a[mt], b[np][2], b[np][3]);
      }
  }
  // ---- softmax over the key frames (columns); thread holds rows mt*16+g and mt*16+g+8, cols nt*8 + 2*t4 + {0,1} ----
  const float sc = rsqrtf(static_cast<float>(dh)) * 1.4426950408889634f;
  uint32_t pf[2][2][4];                               // P as A fragments: [m-tile][k-step of 16 keys]
  float inv[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {                  // hh = 0: row g, 1: row g + 8
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = nt * 8 + 2 * t4 + e;
          float v = s[mt][nt][2 * hh + e];
          if (col >= Tn) v = -INFINITY;
          s[mt][nt][2 * hh + e] = v;
          mx = fmaxf(mx, v);
        }
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float mb = mx * sc;
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float pv = exp2f(fmaf(s[mt][nt][2 * hh + e], sc, -mb));
          s[mt][nt][2 * hh + e] = pv;
          sum += pv;
        }
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      inv[mt][hh] = 1.f / sum;
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      pf[mt][kk][0] = H16<T>::pack2(s[mt][2 * kk][0], s[mt][2 * kk][1]);
      pf[mt][kk][1] = H16<T>::pack2(s[mt][2 * kk][2], s[mt][2 * kk][3]);
      pf[mt][kk][2] = H16<T>::pack2(s[mt][2 * kk + 1][0], s[mt][2 * kk + 1][1]);
      pf[mt][kk][3] = H16<T>::pack2(s[mt][2 * kk + 1][2], s[mt][2 * kk + 1][3]);
    }
  }
  // ---- O = P V, 64 columns of d at a time; O -> (dead) Q tile ----
  constexpr int DSTEP = DHP < 64 ? DHP : 64;
#pragma unroll
  for (int d0 = 0; d0 < DHP; d0 += DSTEP) {
    float o[2][DSTEP / 8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < DSTEP / 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int dp = 0; dp < DSTEP / 16; ++dp) {
        uint32_t vb[4];
        ldsm_x4_t(vb, sV + ta_off<DHP>(kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), d0 / 8 + dp * 2 + (lane >> 4)));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816<T>(o[mt][2 * dp], pf[mt][kk], vb[0], vb[1]);
          mma16816<T>(o[mt][2 * dp + 1], pf[mt][kk], vb[2], vb[3]);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < DSTEP / 8; ++nt) {
        const int r0 = mt * 16 + g, ch = d0 / 8 + nt;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + ta_off<DHP>(r0, ch) + t4 * 4),
                     "r"(H16<T>::pack2(o[mt][nt][0] * inv[mt][0], o[mt][nt][1] * inv[mt][0]))
                     : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + ta_off<DHP>(r0 + 8, ch) + t4 * 4),
                     "r"(H16<T>::pack2(o[mt][nt][2] * inv[mt][1], o[mt][nt][3] * inv[mt][1]))
                     : "memory");
      }
  }
  __syncwarp();
  // ---- coalesced 16-byte stores of the head's [Tn, dh] block ----
  for (int idx = lane; idx < Tn * dch; idx += 32) {
    const int f = idx / dch, c = idx - f * dch;
    uint4 u;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                 : "r"(sQ + ta_off<DHP>(f, c)) : "memory");
    *reinterpret_cast<uint4*>(out + (static_cast<long long>(f) * hw + pos) * C + (h0 + warp) * dh + c * 8) = u;
  }
}

template <typename T, int DHP>
static int launch_temporal(const void* qkv, void* out, int T_, int hw, int C, int heads, cudaStream_t st) {
  // heads per CTA: largest divisor of `heads` whose q|k|v tiles stay <= 48 KB (>= 4 CTAs per SM)
  int hpc = heads;
  while (hpc > 1 && (heads % hpc != 0 || static_cast<size_t>(hpc) * 3 * 32 * DHP * 2 > 48 * 1024)) --hpc;
  const size_t smem = static_cast<size_t>(hpc) * 3 * 32 * DHP * 2;
  auto k = temporal_attention_kernel<T, DHP>;
  VDA_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(k), 64 * 1024));   // per (kernel, device)
  dim3 grid(hw, heads / hpc);
  const int threads = 32 * hpc < 128 ? 128 : 32 * hpc;    // >= 4 warps share the load loop
  k<<<grid, threads, smem, st>>>(static_cast<const T*>(qkv), static_cast<T*>(out), T_, hw, C, heads, hpc);
  VDA_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int dispatch_temporal(const void* qkv, void* out, int T_, int hw, int C, int heads, cudaStream_t st) {
  const int dh = C / heads;
  if (dh <= 16) return launch_temporal<T, 16>(qkv, out, T_, hw, C, heads, st);
  if (dh <= 32) return launch_temporal<T, 32>(qkv, out, T_, hw, C, heads, st);
  if (dh <= 64) return launch_temporal<T, 64>(qkv, out, T_, hw, C, heads, st);
  return launch_temporal<T, 128>(qkv, out, T_, hw, C, heads, st);
}

}  // namespace vda

using namespace vda;

extern "C" int vda_attention_temporal(const void* qkv, void* out, int T, int hw, int C, int heads, int dtype,
                                      void* stream) {
  VDA_CHECK(T > 0 && T <= 32, "temporal attention supports 1..32 frames (temporal_max_len, dpt_temporal.py:38), got %d", T);
  VDA_CHECK(dtype == VDA_BF16 || dtype == VDA_FP16, "bad dtype %d", dtype);
  VDA_CHECK(C % 8 == 0 && C % heads == 0 && (C / heads) % 8 == 0 && heads <= 8 && C / heads <= 128,
            "bad channel/head split C=%d heads=%d (head dim must be a multiple of 8, <= 128; heads <= 8)", C, heads);
  VDA_CHECK((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
            "qkv/out must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == VDA_BF16) return dispatch_temporal<__nv_bfloat16>(qkv, out, T, hw, C, heads, st);
  return dispatch_temporal<__half>(qkv, out, T, hw, C, heads, st);
}
