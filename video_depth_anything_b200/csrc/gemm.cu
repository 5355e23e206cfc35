// tcgen05 / TMEM / TMA GEMM for sm_100a with fused epilogues, also used as an implicit-GEMM
// 3x3 convolution (A tiles gathered by 4-D TMA boxes with out-of-bounds zero fill = padding).
//
//   out[M,N] = A[M,K] . Wt[N,K]^T      A, Wt: bf16|fp16, K-major;  fp32 accumulators in TMEM
//
// One persistent CTA per SM, 192 threads, warp-specialised:
//   warp 0     : TMA producer (one lane)   smem ring of `stages` x {A 128x64, B block_n x 64}, SWIZZLE_128B
//   warp 1     : TMEM allocator + tcgen05.mma issuer (one lane), UMMA 128 x block_n x 16
//   warps 2..5 : epilogue; tcgen05.ld of the 128-lane accumulator (one row per thread), fused math,
//                vectorised global stores.  Two accumulator stages in TMEM overlap the epilogue of
//                tile i with the MMAs of tile i+1.
// block_n (16..256, multiple of 16) is a run-time parameter: it only appears in the instruction
// descriptor, the B tensor map and loop bounds.
#include <stdarg.h>
#include <stdio.h>
#include <mutex>

#include "../../include/vda.h"
#include "common.cuh"

namespace vda {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int kEpiWarps = 8;                 // two warps per TMEM lane quadrant, each takes half of the columns
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 8;
constexpr uint32_t kABytes = BLOCK_M * BLOCK_K * 2;

struct GemmDev {
  int M, N, K;
  int block_n, num_k_blocks, tiles_m, tiles_n, num_tiles, stages;
  uint32_t stage_bytes, tmem_cols, acc_stride;
  // conv addressing
  int H, W, bw, bh, tiles_x, tiles_y, cblocks;
  // epilogue
  const float* bias;
  const float* gamma;
  int act;
  const void* res1;
  long long ldr1;
  int res1_f32;
  const void* res2;
  void* out;
  long long ldo;
  int out_f32;
  void* out_relu;
  int row_group;
  int geglu_half;
  int convt_s, convt_co, in_h, in_w;
  const float* tail_w;
  float tail_b;
};

struct RowInfo {
  bool valid;
  long long out_row;   // row of out / res2 / out_relu (and res1 unless row_group)
  long long res1_row;
};

template <typename T>
__device__ __forceinline__ void store8(T* dst, const float* v) {
  uint4 u;
  u.x = H16<T>::pack2(v[0], v[1]);
  u.y = H16<T>::pack2(v[2], v[3]);
  u.z = H16<T>::pack2(v[4], v[5]);
  u.w = H16<T>::pack2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}
template <typename T>
__device__ __forceinline__ void load8(const T* src, float* v) {
  uint4 u = *reinterpret_cast<const uint4*>(src);
  float2 a = H16<T>::unpack2(u.x), b = H16<T>::unpack2(u.y), c = H16<T>::unpack2(u.z), d = H16<T>::unpack2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void load8f(const float* src, float* v) {
  float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8f(float* dst, const float* v) {
  *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// ---- LINEAR epilogue on 16 consecutive columns starting at global column `col` ------------------
// All global loads are issued before any arithmetic / store (out may alias res1, so the compiler cannot hoist
// loads over stores by itself): one round trip of memory latency per 16 columns instead of one per load.
template <typename T>
__device__ __forceinline__ void epi_linear16(const GemmDev& p, const RowInfo& ri, int col, int ncols, float* v) {
  float r1[16], r2[16];
  const bool two = ncols > 8;
  if (p.res1) {
    if (p.res1_f32) {
      const float* r = reinterpret_cast<const float*>(p.res1) + ri.res1_row * p.ldr1 + col;
      load8f(r, r1); if (two) load8f(r + 8, r1 + 8);
    } else {
      const T* r = reinterpret_cast<const T*>(p.res1) + ri.res1_row * p.ldr1 + col;
      load8<T>(r, r1); if (two) load8<T>(r + 8, r1 + 8);
    }
  }
  if (p.res2) {
    const T* r = reinterpret_cast<const T*>(p.res2) + ri.out_row * p.ldo + col;
    load8<T>(r, r2); if (two) load8<T>(r + 8, r2 + 8);
  }
  if (p.bias) {   // bias / gamma are tiny, L1-resident vectors: short latency, loaded where they are used
    float t[8];
    load8f(p.bias + col, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += t[i];
    if (two) {
      load8f(p.bias + col + 8, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 + i] += t[i];
    }
  }
  if (p.gamma) {
    float t[8];
    load8f(p.gamma + col, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= t[i];
    if (two) {
      load8f(p.gamma + col + 8, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 + i] *= t[i];
    }
  }
  if (p.act == VDA_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
  } else if (p.act == VDA_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (p.res1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += r1[i];
  }
  if (p.res2) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += r2[i];
  }
  if (p.out_f32) {
    float* o = reinterpret_cast<float*>(p.out) + ri.out_row * p.ldo + col;
    store8f(o, v); if (two) store8f(o + 8, v + 8);
  } else {
    T* o = reinterpret_cast<T*>(p.out) + ri.out_row * p.ldo + col;
    store8<T>(o, v); if (two) store8<T>(o + 8, v + 8);
  }
  if (p.out_relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    T* o = reinterpret_cast<T*>(p.out_relu) + ri.out_row * p.ldo + col;
    store8<T>(o, v); if (two) store8<T>(o + 8, v + 8);
  }
}

template <typename T, int EPI, bool CONV>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // 1024-byte aligned operand ring (SWIZZLE_128B atoms are 8 rows x 128 B)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 32 * kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.tiles_n;
        const int m_blk = tile / p.tiles_n;
        int img = 0, x0 = 0, y0 = 0;
        if (CONV) {
          const int per_img = p.tiles_x * p.tiles_y;
          img = m_blk / per_img;
          const int rem = m_blk - img * per_img;
          y0 = (rem / p.tiles_x) * p.bh;
          x0 = (rem % p.tiles_x) * p.bw;
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], p.stage_bytes);
          uint8_t* sA = smem_gen + static_cast<size_t>(stage) * p.stage_bytes;
          uint8_t* sB = sA + kABytes;
          if (CONV) {
            const int tap = kb / p.cblocks;
            const int cb = kb - tap * p.cblocks;
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            tma_load_4d(sA, &tmA, &full_bar[stage], cb * BLOCK_K, x0 + dx, y0 + dy, img);
          } else {
            tma_load_2d(sA, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          }
          tma_load_2d(sB, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * p.block_n);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(H16<T>::kUmmaFmt, static_cast<uint32_t>(p.block_n));
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * p.acc_stride;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_base + stage * p.stage_bytes;
          const uint64_t da = umma_desc_sw128(sA);
          const uint64_t db = umma_desc_sw128(sA + kABytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            // +32 bytes (16 elements) along K inside the 128-byte swizzle atom = +2 in the address field
            umma_f16(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[as]);       // accumulator ready for the epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ================================ epilogue ====================================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int eh = (warp - 2) >> 2;         // which half of the tile's columns this warp drains
    const int r = q * 32 + lane;            // row of the tile owned by this thread
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n_blk = tile % p.tiles_n;
      const int m_blk = tile / p.tiles_n;
      RowInfo ri;
      if (CONV) {
        const int per_img = p.tiles_x * p.tiles_y;
        const int img = m_blk / per_img;
        const int rem = m_blk - img * per_img;
        const int y = (rem / p.tiles_x) * p.bh + r / p.bw;
        const int x = (rem % p.tiles_x) * p.bw + r % p.bw;
        ri.valid = (y < p.H) && (x < p.W);
        ri.out_row = (static_cast<long long>(img) * p.H + y) * p.W + x;
        ri.res1_row = ri.out_row;
      } else {
        const long long m = static_cast<long long>(m_blk) * BLOCK_M + r;
        ri.valid = m < p.M;
        ri.out_row = m;
        ri.res1_row = m;
        if (p.row_group > 0) {
          ri.out_row = m + m / p.row_group + 1;
          ri.res1_row = m % p.row_group + 1;
        }
        if (EPI == VDA_EPI_CONVT) {
          const int per_img = p.in_h * p.in_w;
          const int img = static_cast<int>(m / per_img);
          const int rem = static_cast<int>(m - static_cast<long long>(img) * per_img);
          const int y = rem / p.in_w, x = rem - y * p.in_w;
          // row index of output pixel (img, y*S, x*S) in the upsampled map
          ri.out_row = (static_cast<long long>(img) * p.in_h * p.convt_s + static_cast<long long>(y) * p.convt_s) *
                           (p.in_w * p.convt_s) + static_cast<long long>(x) * p.convt_s;
        }
      }
      const int col_base = n_blk * p.block_n;
      if (EPI == VDA_EPI_LINEAR && ri.valid && (p.res1 || p.res2)) {
        // pull this thread's residual segment towards L2 while the tile's MMAs are still running
        const int nch = p.block_n >> 4;
        const int c_lo = col_base + (eh ? (nch + 1) / 2 : 0) * 16;
        int c_hi = col_base + (eh ? nch : (nch + 1) / 2) * 16;
        if (c_hi > p.N) c_hi = p.N;
        if (p.res1) {
          const int esz = p.res1_f32 ? 4 : 2;
          const char* base = reinterpret_cast<const char*>(p.res1) + (ri.res1_row * p.ldr1 + c_lo) * esz;
          for (int off = 0; off < (c_hi - c_lo) * esz; off += 128) prefetch_l2(base + off);
        }
        if (p.res2) {
          const char* base = reinterpret_cast<const char*>(p.res2) + (ri.out_row * p.ldo + c_lo) * 2;
          for (int off = 0; off < (c_hi - c_lo) * 2; off += 128) prefetch_l2(base + off);
        }
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * p.acc_stride;

      if (EPI == VDA_EPI_LINEAR || EPI == VDA_EPI_CONVT) {
        const int nch = p.block_n >> 4;
        const int ch_begin = eh ? (nch + 1) / 2 : 0, ch_end = eh ? nch : (nch + 1) / 2;
        auto process = [&](const uint32_t (&rr)[16], int c0) {
          if (!ri.valid) return;
          if (EPI == VDA_EPI_LINEAR) {
            const int col = col_base + c0;
            if (col < p.N) {
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]);
              epi_linear16<T>(p, ri, col, p.N - col, v);
            }
            return;
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int col = col_base + c0 + 8 * g;
            if (col < p.N) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(rr[8 * g + i]);
              const int kk = col / p.convt_co;
              const int co = col - kk * p.convt_co;
              const int ky = kk / p.convt_s, kx = kk - ky * p.convt_s;
              float t[8];
              load8f(p.bias + co, t);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += t[i];
              const long long orow = ri.out_row + static_cast<long long>(ky) * (p.in_w * p.convt_s) + kx;
              store8<T>(reinterpret_cast<T*>(p.out) + orow * p.ldo + co, v);
            }
          }
        };
        // software-pipelined TMEM drain: the load of chunk i+1 is in flight while chunk i is processed
        uint32_t ra[16], rb[16];
        if (ch_begin < ch_end) tmem_ld16(t_row + ch_begin * 16, ra);
        for (int ch = ch_begin; ch < ch_end; ch += 2) {
          tmem_ld_wait16(ra);
          if (ch + 1 < ch_end) tmem_ld16(t_row + (ch + 1) * 16, rb);
          process(ra, ch * 16);
          if (ch + 1 < ch_end) {
            tmem_ld_wait16(rb);
            if (ch + 2 < ch_end) tmem_ld16(t_row + (ch + 2) * 16, ra);
            process(rb, (ch + 1) * 16);
          }
        }
      } else if (EPI == VDA_EPI_GEGLU) {
        const int half = p.geglu_half;
        const int nch = half >> 4;
        const int ch_begin = eh ? (nch + 1) / 2 : 0, ch_end = eh ? nch : (nch + 1) / 2;
        for (int c0 = ch_begin * 16; c0 < ch_end * 16; c0 += 16) {
          uint32_t ra[16], rg[16];
          tmem_ld16(t_row + c0, ra);
          tmem_ld16(t_row + half + c0, rg);
          tmem_ld_wait16(ra);
          tmem_ld_wait16(rg);
          if (ri.valid) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int pc = col_base + c0 + 8 * g;          // packed column of `a`
              if (pc < p.N) {
                float ba[8], bg[8], v[8];
                load8f(p.bias + pc, ba);
                load8f(p.bias + pc + half, bg);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float a = __uint_as_float(ra[8 * g + i]) + ba[i];
                  const float gt = __uint_as_float(rg[8 * g + i]) + bg[i];
                  v[i] = a * gelu_erf(gt);
                }
                const int oc = n_blk * half + c0 + 8 * g;
                store8<T>(reinterpret_cast<T*>(p.out) + ri.out_row * p.ldo + oc, v);
              }
            }
          }
        }
      } else {  // VDA_EPI_TAIL: block_n == N == 32
        if (eh == 0) {   // 32 columns only: one warp per quadrant does the whole row
          float acc = p.tail_b;
          uint32_t r0[16], r1[16];
          tmem_ld16(t_row, r0);
          tmem_ld16(t_row + 16, r1);
          tmem_ld_wait16(r0);
          tmem_ld_wait16(r1);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float h = fmaxf(__uint_as_float(r0[i]) + __ldg(p.bias + i), 0.f);
            acc = fmaf(h, __ldg(p.tail_w + i), acc);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float h = fmaxf(__uint_as_float(r1[i]) + __ldg(p.bias + 16 + i), 0.f);
            acc = fmaf(h, __ldg(p.tail_w + 16 + i), acc);
          }
          if (ri.valid) reinterpret_cast<float*>(p.out)[ri.out_row] = fmaxf(acc, 0.f);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

int make_tensor_map(CUtensorMap* m, int dtype, const void* base, int rank, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  VDA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, dtype == VDA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VDA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu box %u %u)",
            static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return 0;
}

static int g_sm_count = 0;
int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

static int pick_block_n(int N, int tiles_m) {
  int bn;
  if (N <= 256) {
    bn = (N + 15) / 16 * 16;
  } else {
    bn = 128;
    for (int c = 256; c >= 128; c -= 16)
      if (N % c == 0) { bn = c; break; }
  }
  // fill the machine: halve wide tiles while there are fewer tiles than SMs
  while (bn >= 128 && bn % 32 == 0 && tiles_m * ((N + bn - 1) / bn) < sm_count()) bn /= 2;
  return bn;
}

template <typename T, int EPI, bool CONV>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& d, size_t smem, cudaStream_t st) {
  auto kfn = gemm_kernel<T, EPI, CONV>;
  static size_t attr_smem = 0;   // per instantiation: largest dynamic smem opted in so far
  if (smem > attr_smem) {
    VDA_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_smem = smem;
  }
  const int grid = d.num_tiles < sm_count() ? d.num_tiles : sm_count();
  kfn<<<grid, kThreads, smem, st>>>(tmA, tmB, d);
  VDA_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int dispatch(const vda_gemm_params* p, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& d,
                    size_t smem, cudaStream_t st) {
  const bool conv = p->a_mode == VDA_A_CONV3;
  switch (p->epilogue) {
    case VDA_EPI_LINEAR:
      return conv ? launch<T, VDA_EPI_LINEAR, true>(tmA, tmB, d, smem, st)
                  : launch<T, VDA_EPI_LINEAR, false>(tmA, tmB, d, smem, st);
    case VDA_EPI_GEGLU:
      VDA_CHECK(!conv, "GEGLU epilogue is only defined for plain GEMMs");
      return launch<T, VDA_EPI_GEGLU, false>(tmA, tmB, d, smem, st);
    case VDA_EPI_CONVT:
      VDA_CHECK(!conv, "CONVT epilogue is only defined for plain GEMMs");
      return launch<T, VDA_EPI_CONVT, false>(tmA, tmB, d, smem, st);
    case VDA_EPI_TAIL:
      return conv ? launch<T, VDA_EPI_TAIL, true>(tmA, tmB, d, smem, st)
                  : launch<T, VDA_EPI_TAIL, false>(tmA, tmB, d, smem, st);
  }
  set_error("unknown epilogue %d", p->epilogue);
  return 1;
}

}  // namespace vda

using namespace vda;

extern "C" int vda_gemm(const vda_gemm_params* p, void* stream) {
  VDA_CHECK(p != nullptr, "null params");
  VDA_CHECK(p->M > 0 && p->N > 0 && p->K > 0, "bad GEMM shape %d %d %d", p->M, p->N, p->K);
  VDA_CHECK(p->dtype == VDA_BF16 || p->dtype == VDA_FP16, "bad dtype %d", p->dtype);
  VDA_CHECK(p->N % 8 == 0, "N (%d) must be a multiple of 8", p->N);
  VDA_CHECK(p->K % 8 == 0, "K (%d) must be a multiple of 8", p->K);
  VDA_CHECK((reinterpret_cast<uintptr_t>(p->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->Wt) & 15) == 0,
            "A / Wt must be 16-byte aligned");
  VDA_CHECK(p->out != nullptr && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0, "out must be 16-byte aligned");
  const bool conv = p->a_mode == VDA_A_CONV3;

  GemmDev d = {};
  d.M = p->M; d.N = p->N; d.K = p->K;
  d.num_k_blocks = (p->K + BLOCK_K - 1) / BLOCK_K;
  CUtensorMap tmA, tmB;

  if (conv) {
    VDA_CHECK(p->C % 64 == 0 && p->K == 9 * p->C, "conv mode needs C %% 64 == 0 and K == 9*C (C=%d K=%d)", p->C, p->K);
    VDA_CHECK(p->M == p->n_img * p->H * p->W, "conv mode: M must be n_img*H*W");
    d.H = p->H; d.W = p->W;
    // pick the 128-pixel box shape (bw x bh) wasting the fewest out-of-image pixels
    const int cand[5][2] = {{16, 8}, {32, 4}, {8, 16}, {64, 2}, {128, 1}};
    long long best = -1;
    for (int i = 0; i < 5; ++i) {
      const long long tx = (p->W + cand[i][0] - 1) / cand[i][0], ty = (p->H + cand[i][1] - 1) / cand[i][1];
      if (best < 0 || tx * ty < best) { best = tx * ty; d.bw = cand[i][0]; d.bh = cand[i][1]; }
    }
    d.tiles_x = (p->W + d.bw - 1) / d.bw;
    d.tiles_y = (p->H + d.bh - 1) / d.bh;
    d.tiles_m = d.tiles_x * d.tiles_y * p->n_img;
    d.cblocks = p->C / 64;
    cuuint64_t dims[4] = {(cuuint64_t)p->C, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->n_img};
    cuuint64_t strides[3] = {(cuuint64_t)p->C * 2, (cuuint64_t)p->C * 2 * p->W, (cuuint64_t)p->C * 2 * p->W * p->H};
    cuuint32_t box[4] = {64, (cuuint32_t)d.bw, (cuuint32_t)d.bh, 1};
    if (make_tensor_map(&tmA, p->dtype, p->A, 4, dims, strides, box)) return 1;
  } else {
    VDA_CHECK(p->lda >= p->K && p->lda % 8 == 0, "lda (%lld) must be >= K and a multiple of 8", (long long)p->lda);
    d.tiles_m = (p->M + BLOCK_M - 1) / BLOCK_M;
    cuuint64_t dims[2] = {(cuuint64_t)p->K, (cuuint64_t)p->M};
    cuuint64_t strides[1] = {(cuuint64_t)p->lda * 2};
    cuuint32_t box[2] = {64, 128};
    if (make_tensor_map(&tmA, p->dtype, p->A, 2, dims, strides, box)) return 1;
  }

  if (p->epilogue == VDA_EPI_GEGLU) {
    VDA_CHECK(p->geglu_half > 0 && p->geglu_half % 16 == 0 && p->geglu_half <= 128 &&
                  p->N % (2 * p->geglu_half) == 0,
              "GEGLU: N must be a multiple of 2*geglu_half (<=256)");
    d.block_n = 2 * p->geglu_half;
  } else if (p->epilogue == VDA_EPI_TAIL) {
    VDA_CHECK(p->N == 32 && p->tail_w && p->bias, "TAIL epilogue needs N == 32, bias and tail_w");
    d.block_n = 32;
  } else {
    d.block_n = pick_block_n(p->N, d.tiles_m);
  }
  if (p->epilogue == VDA_EPI_CONVT) {
    VDA_CHECK(p->convt_s > 0 && p->convt_co % 8 == 0 && p->N == p->convt_s * p->convt_s * p->convt_co && p->bias &&
                  p->M % (p->in_h * p->in_w) == 0,
              "CONVT: inconsistent shape");
  }
  if (p->epilogue == VDA_EPI_LINEAR) {
    VDA_CHECK(p->ldo % 8 == 0 && (!p->res1 || p->ldr1 % 8 == 0), "ldo / ldr1 must be multiples of 8");
  }
  d.tiles_n = (p->N + d.block_n - 1) / d.block_n;
  d.num_tiles = d.tiles_m * d.tiles_n;
  {
    cuuint64_t dims[2] = {(cuuint64_t)p->K, (cuuint64_t)p->N};
    cuuint64_t strides[1] = {(cuuint64_t)p->K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)d.block_n};
    if (make_tensor_map(&tmB, p->dtype, p->Wt, 2, dims, strides, box)) return 1;
  }
  d.stage_bytes = kABytes + static_cast<uint32_t>(d.block_n) * BLOCK_K * 2;
  // block_n*128 is a multiple of 2048 only when block_n % 16 == 0 -> every stage stays 1024-byte aligned
  d.stage_bytes = (d.stage_bytes + 1023u) & ~1023u;
  int stages = static_cast<int>((227u * 1024u - 2048u) / d.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  d.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * d.stage_bytes + 1024;
  uint32_t cols = 32;
  while (cols < 2u * d.block_n) cols <<= 1;
  d.tmem_cols = cols;
  d.acc_stride = cols / 2;

  d.bias = p->bias; d.gamma = p->gamma; d.act = p->act;
  d.res1 = p->res1; d.ldr1 = p->ldr1; d.res1_f32 = p->res1_f32; d.res2 = p->res2;
  d.out = p->out; d.ldo = p->ldo; d.out_f32 = p->out_f32; d.out_relu = p->out_relu;
  d.row_group = p->row_group; d.geglu_half = p->geglu_half;
  d.convt_s = p->convt_s; d.convt_co = p->convt_co; d.in_h = p->in_h; d.in_w = p->in_w;
  d.tail_w = p->tail_w; d.tail_b = p->tail_b;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->dtype == VDA_BF16) return dispatch<__nv_bfloat16>(p, tmA, tmB, d, smem, st);
  return dispatch<__half>(p, tmA, tmB, d, smem, st);
}
