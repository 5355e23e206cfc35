// Spatial ViT attention (dinov2_layers/attention.py:49-62) as a tcgen05 / TMEM / TMA flash-attention kernel.
//
//   out[f, n, h, :] = softmax_n'( q[f,n,h,:] . k[f,n',h,:] / sqrt(64) ) v[f,n',h,:]       d = 64, N ~ 1370..2443
//
// Persistent kernel, one CTA (320 threads) per SM; a work item is (frame, head, pair of 128-query tiles):
//   warp 0      TMA producer: Q tiles once per item, K / V tiles (128 keys x 64) through two 3-slot rings
//   warp 1      tcgen05.mma issuer + TMEM owner:  S = Q K^T  (SS form, 128 x kv x 64)  and  O += P V  (TS form: P is
//               read from TMEM, V from smem as an MN-major operand -- no transpose of V anywhere)
//   warps 2-5   softmax warpgroup of query tile A      \  one thread per query row: the 128 scores of a row are read
//   warps 6-9   softmax warpgroup of query tile B      /  from TMEM straight into that thread's registers (no shuffles)
// TMEM (512 columns): S_A | S_B (128 fp32 columns each), O_A | O_B (64), P_A | P_B (64: 128 packed 16-bit keys).
// Because a warpgroup copies S to registers before it starts the exponentials, the issuer refills S with the next
// key tile immediately (s_free); P has its own columns, so S(j+1) never waits for O += P(j) V(j).  The two query
// tiles share every K/V tile (halves the L2 -> smem traffic) and keep the MUFU busy while the other warpgroup syncs.
// Online softmax in fp32 with lazy rescaling: O / l are only rescaled when the running row maximum grows by more
// than 2^8 (rare after the first key tiles), done by the owning warp through a TMEM round trip.
// The kernel is exp-bound (MUFU.EX2: 16/clk/SM vs 8192 dense FLOP/clk/SM, d = 64), see DESIGN.md.
#include "../../include/vda.h"
#include "common.cuh"

namespace vda {

namespace sa {
constexpr int BM = 128;            // queries per tile
constexpr int BN = 128;            // keys per tile
constexpr int D = 64;              // head dim
constexpr int KS = 3, VS = 3;      // K / V ring depth
constexpr int THREADS = 320;
constexpr uint32_t TILE_BYTES = BM * D * 2;   // 16 KB: one Q, K or V tile
constexpr uint32_t SMEM_BYTES = (2 + KS + VS) * TILE_BYTES + 1024;
// TMEM columns
constexpr uint32_t COL_S = 0, COL_O = 256, COL_P = 384;
constexpr float RESCALE_LOG2 = 8.0f;
}  // namespace sa

struct SaParams {
  int N, heads, frames;
  int n_qt;            // query tiles per (frame, head)
  int n_pairs;         // full pairs of query tiles per (frame, head)
  int n_items;         // FH * n_pairs + (n_qt odd ? FH : 0)
  int n_kv;            // key tiles
  void* out;
};

struct SaItem {
  int frame, head, q0;   // q0: first query row of tile A
  bool has_b;
};

__device__ __forceinline__ SaItem sa_decode(const SaParams& p, int item) {
  SaItem it;
  const int fh_total = p.frames * p.heads;
  int fh, tile;
  if (item < fh_total * p.n_pairs) {
    fh = item / p.n_pairs;
    tile = 2 * (item - fh * p.n_pairs);
    it.has_b = true;
  } else {
    fh = item - fh_total * p.n_pairs;
    tile = p.n_qt - 1;
    it.has_b = false;
  }
  it.frame = fh / p.heads;
  it.head = fh - it.frame * p.heads;
  it.q0 = tile * sa::BM;
  return it;
}

__device__ __forceinline__ int sa_kv_cols(const SaParams& p, int j) {   // columns of key tile j, rounded up to 32
  const int valid = p.N - j * sa::BN;
  return valid >= sa::BN ? sa::BN : ((valid + 31) & ~31);
}

template <typename T>
__global__ void __launch_bounds__(sa::THREADS, 1)
spatial_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const SaParams p) {
  using namespace sa;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, q_empty;
  __shared__ __align__(8) uint64_t k_full[KS], k_empty[KS], v_full[VS], v_empty[VS];
  __shared__ __align__(8) uint64_t s_ready[2], s_free[2], p_ready[2], o_done[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // smem map: Q_A | Q_B | K[KS] | V[VS]
  const uint32_t offQ = 0, offK = 2 * TILE_BYTES, offV = (2 + KS) * TILE_BYTES;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(&q_full, 1);
    mbar_init(&q_empty, 1);
    for (int s = 0; s < KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < VS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_ready[t], 1);
      mbar_init(&s_free[t], 128);
      mbar_init(&p_ready[t], 128);
      mbar_init(&o_done[t], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, it_n = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it_n) {
        const SaItem it = sa_decode(p, item);
        mbar_wait(&q_empty, (it_n & 1u) ^ 1u);
        mbar_arrive_expect_tx(&q_full, it.has_b ? 2 * TILE_BYTES : TILE_BYTES);
        tma_load_4d(smem_gen + offQ, &tmQKV, &q_full, 0, it.head, it.q0, it.frame);
        if (it.has_b) tma_load_4d(smem_gen + offQ + TILE_BYTES, &tmQKV, &q_full, 0, it.head, it.q0 + BM, it.frame);
        for (int j = 0; j < p.n_kv; ++j) {
          mbar_wait(&k_empty[ks], kph ^ 1u);
          mbar_arrive_expect_tx(&k_full[ks], TILE_BYTES);
          tma_load_4d(smem_gen + offK + ks * TILE_BYTES, &tmQKV, &k_full[ks], 0, p.heads + it.head, j * BN, it.frame);
          if (++ks == KS) { ks = 0; kph ^= 1u; }
          mbar_wait(&v_empty[vs], vph ^ 1u);
          mbar_arrive_expect_tx(&v_full[vs], TILE_BYTES);
          tma_load_4d(smem_gen + offV + vs * TILE_BYTES, &tmQKV, &v_full[vs], 0, 2 * p.heads + it.head, j * BN,
                      it.frame);
          if (++vs == VS) { vs = 0; vph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, it_n = 0;
      uint32_t n_s[2] = {0, 0}, n_pv[2] = {0, 0};     // running counts of S / PV tiles issued per query tile
      const uint32_t idesc_pv = umma_idesc(H16<T>::kUmmaFmt, D) | kIdescBMnMajor;

      auto issue_s = [&](int t, int kslot, int cols) {
        mbar_wait(&s_free[t], (n_s[t] & 1u) ^ 1u);      // the warpgroup has copied the previous S to registers
        tc_fence_after();
        const uint32_t idesc = umma_idesc(H16<T>::kUmmaFmt, static_cast<uint32_t>(cols));
        const uint64_t da = umma_desc_sw128(smem_base + offQ + t * TILE_BYTES);
        const uint64_t db = umma_desc_sw128(smem_base + offK + kslot * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_f16(tmem_base + COL_S + t * BN, da + 2u * k, db + 2u * k, idesc, k);
        umma_commit(&s_ready[t]);
        ++n_s[t];
      };
      auto issue_pv = [&](int t, int vslot, int cols, bool first) {
        mbar_wait(&p_ready[t], n_pv[t] & 1u);
        tc_fence_after();
        const uint64_t db = umma_desc_sw128_mn(smem_base + offV + vslot * TILE_BYTES);
        const int nk = cols >> 4;
        for (int k = 0; k < nk; ++k)   // 16 keys per MMA: 8 TMEM columns of P, 16 rows (2048 B) of V
          umma_f16_ts(tmem_base + COL_O + t * D, tmem_base + COL_P + t * (BN / 2) + 8u * k, db + 128u * k, idesc_pv,
                      (!first || k > 0) ? 1u : 0u);
        umma_commit(&o_done[t]);
        ++n_pv[t];
      };

      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it_n) {
        const SaItem it = sa_decode(p, item);
        mbar_wait(&q_full, it_n & 1u);
        // S(0)
        mbar_wait(&k_full[ks], kph);
        issue_s(0, ks, sa_kv_cols(p, 0));
        if (it.has_b) issue_s(1, ks, sa_kv_cols(p, 0));
        umma_commit(&k_empty[ks]);
        if (++ks == KS) { ks = 0; kph ^= 1u; }
        for (int j = 0; j < p.n_kv; ++j) {
          if (j + 1 < p.n_kv) {
            const int cols = sa_kv_cols(p, j + 1);
            mbar_wait(&k_full[ks], kph);
            issue_s(0, ks, cols);
            if (it.has_b) issue_s(1, ks, cols);
            umma_commit(&k_empty[ks]);
            if (++ks == KS) { ks = 0; kph ^= 1u; }
          } else {
            umma_commit(&q_empty);        // every S of this item has been issued: Q may be overwritten when they retire
          }
          const int cols = sa_kv_cols(p, j);
          mbar_wait(&v_full[vs], vph);
          issue_pv(0, vs, cols, j == 0);
          if (it.has_b) issue_pv(1, vs, cols, j == 0);
          umma_commit(&v_empty[vs]);
          if (++vs == VS) { vs = 0; vph ^= 1u; }
        }
      }
    }
  } else {
    // ===================================== softmax warpgroups ===============================
    const int t = (warp - 2) >> 2;                    // query tile handled by this warpgroup
    const int quad = warp & 3;                        // TMEM lane quadrant of this warp
    const int row = quad * 32 + lane;                 // query row inside the tile
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + COL_S + t * BN;
    const uint32_t tO = tmem_base + lane_base + COL_O + t * D;
    const uint32_t tP = tmem_base + lane_base + COL_P + t * (BN / 2);
    const float sc = 0.125f * 1.4426950408889634f;    // d^-0.5 * log2(e)
    uint32_t cnt = 0;                                 // running key-tile counter of this query tile
    T* outp = reinterpret_cast<T*>(p.out);

    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const SaItem it = sa_decode(p, item);
      if (t == 1 && !it.has_b) continue;
      float m_run = 0.f, l_run = 0.f;
      for (int j = 0; j < p.n_kv; ++j, ++cnt) {
        const int cols = sa_kv_cols(p, j);
        uint32_t s[BN];
        mbar_wait(&s_ready[t], cnt & 1u);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < BN; c += 32)
          if (c < cols) tmem_ld32(tS + c, s + c);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[t]);
        // ---- mask keys beyond N (only the last key tile can be partial) ----
        const int valid = p.N - j * BN;
        if (valid < BN) {
#pragma unroll
          for (int c = 0; c < BN; ++c)
            if (c >= valid) s[c] = 0xff800000u;   // -inf
        }
        // ---- row maximum ----
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          if (c < cols) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              mx0 = fmaxf(mx0, __uint_as_float(s[c + i]));
              mx1 = fmaxf(mx1, __uint_as_float(s[c + i + 1]));
              mx2 = fmaxf(mx2, __uint_as_float(s[c + i + 2]));
              mx3 = fmaxf(mx3, __uint_as_float(s[c + i + 3]));
            }
          }
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        bool waited = false;
        if (j == 0) {
          m_run = mx;
        } else {
          const float m_new = fmaxf(m_run, mx);
          const bool need = (m_new - m_run) * sc > RESCALE_LOG2;
          if (__any_sync(0xffffffffu, need)) {
            // rare: rescale O (TMEM round trip by the owning warp) once the previous O += P V has retired
            mbar_wait(&o_done[t], (cnt - 1u) & 1u);
            tc_fence_after();
            waited = true;
            float alpha = 1.f;
            if (need) {
              alpha = exp2f((m_run - m_new) * sc);
              m_run = m_new;
              l_run *= alpha;
            }
#pragma unroll
            for (int c = 0; c < D; c += 16) {
              uint32_t o[16];
              tmem_ld16(tO + c, o);
              tmem_ld_wait16(o);
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(tO + c, o);
            }
            tmem_st_wait();
          }
        }
        // ---- P = exp2(s*sc - m*sc), row sum, pack to 16 bit, store to TMEM ----
        const float mb = m_run * sc;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          if (c < cols) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float p0 = exp2f(fmaf(__uint_as_float(s[c + i]), sc, -mb));
              const float p1 = exp2f(fmaf(__uint_as_float(s[c + i + 1]), sc, -mb));
              const float p2 = exp2f(fmaf(__uint_as_float(s[c + i + 2]), sc, -mb));
              const float p3 = exp2f(fmaf(__uint_as_float(s[c + i + 3]), sc, -mb));
              l0 += p0; l1 += p1; l2 += p2; l3 += p3;
              pk[i >> 1] = H16<T>::pack2(p0, p1);
              pk[(i >> 1) + 1] = H16<T>::pack2(p2, p3);
            }
            if (c == 0 && j > 0 && !waited) {   // P(j-1) must have been consumed by O += P V before it is overwritten
              mbar_wait(&o_done[t], (cnt - 1u) & 1u);
              tc_fence_after();
            }
            tmem_st16(tP + (c >> 1), pk);
          }
        }
        l_run += (l0 + l1) + (l2 + l3);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[t]);
      }
      // ---- epilogue: O / l -> 16 bit -> global ----
      mbar_wait(&o_done[t], (cnt - 1u) & 1u);
      tc_fence_after();
      const float inv = 1.f / l_run;
      const int q = it.q0 + t * BM + row;
      T* orow = outp + (static_cast<long long>(it.frame) * p.N + q) * (static_cast<long long>(p.heads) * D) +
                it.head * D;
#pragma unroll
      for (int c = 0; c < D; c += 32) {
        uint32_t o[32];
        tmem_ld32(tO + c, o);
        tmem_ld_wait();
        if (q < p.N) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 u;
            u.x = H16<T>::pack2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
            u.y = H16<T>::pack2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
            u.z = H16<T>::pack2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
            u.w = H16<T>::pack2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c + i) = u;
          }
        }
      }
      tc_fence_before();   // the O reads above are ordered before this thread's next p_ready arrive
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vda

using namespace vda;

extern "C" int vda_attention_spatial(const void* qkv, void* out, int frames, int N, int heads, int dtype,
                                     void* stream) {
  VDA_CHECK(frames > 0 && N > 0 && heads > 0, "bad attention shape");
  VDA_CHECK(dtype == VDA_BF16 || dtype == VDA_FP16, "bad dtype %d", dtype);
  VDA_CHECK((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
            "qkv/out must be 16-byte aligned");
  // qkv is [frames, N, 3, heads, 64]: a 4-D tensor (d, which*heads + head, token, frame); token rows beyond N are
  // zero-filled by TMA instead of running into the next frame
  CUtensorMap tm;
  const cuuint64_t row_bytes = static_cast<cuuint64_t>(3) * heads * sa::D * 2;
  cuuint64_t dims[4] = {sa::D, static_cast<cuuint64_t>(3 * heads), static_cast<cuuint64_t>(N),
                        static_cast<cuuint64_t>(frames)};
  cuuint64_t strides[3] = {sa::D * 2, row_bytes, row_bytes * N};
  cuuint32_t box[4] = {sa::D, 1, sa::BM, 1};
  if (make_tensor_map(&tm, dtype, qkv, 4, dims, strides, box)) return 1;

  SaParams p;
  p.N = N; p.heads = heads; p.frames = frames;
  p.n_qt = (N + sa::BM - 1) / sa::BM;
  p.n_pairs = p.n_qt / 2;
  p.n_items = frames * heads * p.n_pairs + ((p.n_qt & 1) ? frames * heads : 0);
  p.n_kv = (N + sa::BN - 1) / sa::BN;
  p.out = out;
  const int grid = p.n_items < sm_count() ? p.n_items : sm_count();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == VDA_BF16) {
    auto k = spatial_attention_tc_kernel<__nv_bfloat16>;
    static bool attr = false;
    if (!attr) {
      VDA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sa::SMEM_BYTES));
      attr = true;
    }
    k<<<grid, sa::THREADS, sa::SMEM_BYTES, st>>>(tm, p);
  } else {
    auto k = spatial_attention_tc_kernel<__half>;
    static bool attr = false;
    if (!attr) {
      VDA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sa::SMEM_BYTES));
      attr = true;
    }
    k<<<grid, sa::THREADS, sa::SMEM_BYTES, st>>>(tm, p);
  }
  VDA_CUDA(cudaGetLastError());
  return 0;
}
