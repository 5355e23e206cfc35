// Spatial ViT attention, four-stream version (dinov2_layers/attention.py:49-62): the same math as attention_spatial.cu,
// re-organised for thread-level parallelism.
//
// Why: the two-stream kernel (one 128-query x 128-key score tile per softmax warpgroup, two warpgroups per SM) leaves
// each SM sub-partition with TWO softmax warps.  ncu at the benchmarked shape (32 x 1370 x 16, profiles/r02*_attn*): issue
// slots 49 % busy, MUFU 48 %, tensor 30 %; a softmax warp issues 24 % of the time and sits in fixed-latency dependency
// stalls (`wait`) 30 %, MUFU queue / result stalls 15 %: with two warps the scheduler has nothing else to issue.
// Restructuring waits inside that design (own MMA issuer per stream, split-KV items) changed nothing (0.386 -> 0.390 ms).
//
// Here a CTA (640 threads, one per SM) runs FOUR streams, each = one softmax warpgroup owning a 128-query tile and
// issuing its own MMAs, on 64-key score tiles:
//   warp 0        TMA producer: per item the Q tiles of up to four neighbouring query tiles of one (frame, head), then the
//                 K / V tiles (64 keys x 64) of that (frame, head) through two 5-slot rings shared by all four streams
//   warp 1        TMEM allocation;  warps 2-3 idle (the control warpgroup donates its registers)
//   warps 4-19    softmax warpgroups of streams 0-3: one thread per query row, the 64 scores of a row go TMEM ->
//                 registers, online softmax (lazy rescale), P packed to 16 bit and written back INTO the S columns.
//                 The first warp of a warpgroup is also the stream's tcgen05.mma issuer (one elected lane):
//                 S = Q K^T (SS, 128 x 64 x 64), then per step  O += P V  (TS: P from TMEM, V from smem as MN-major
//                 operand)  immediately followed by the next S.  (A stream is strictly serial -- scores, softmax, P V, next
//                 scores -- so a separate issuer warp would only add a hand-over; and with 768 threads the register
//                 pool released by the control warps -- setmaxnreg can only draw on registers of the same CTA -- tops
//                 out at 104 per softmax thread, 640 threads allow 112.)
// TMEM (512 columns): stream t owns S_t = columns [64 t, 64 t + 64) and O_t = [256 + 64 t, ...); P_t aliases S_t[0, 32).
// Because the issuer queues O += P V and the next S back to back (the tensor pipe executes one thread's MMAs in order),
// a stream needs one mbarrier (s_ready: tensor pipe -> warpgroup) and one 128-thread named barrier (all P stores done ->
// issuer) per step: no s_free, no o_done, and "S(j+1) is complete" implies "O += P(j) V(j) is complete", which is all the
// lazy rescale needs.
// A stream idles while its own MMAs run (~2 x 128 tensor cycles + latency); the other three fill the issue slots: four
// softmax warps per scheduler instead of two, 64 + 32 live score registers per thread instead of 128 + 32.
// Ring slots / Q buffers expect four arrivals (one tcgen05.commit per consuming stream); for a query-tile group with
// fewer than four tiles the producer arrives for the absent streams.
#include "../../include/vda.h"
#include "common.cuh"

namespace vda {

namespace sa4 {
constexpr int BM = 128;            // queries per tile
constexpr int BN = 64;             // keys per tile
constexpr int D = 64;              // head dim
constexpr int NS = 4;              // streams
constexpr int KS = 5, VS = 5;      // K / V ring depth
constexpr int THREADS = 640;       // 4 control warps + 4 softmax warpgroups
constexpr int REGS_CTRL = 24, REGS_SOFTMAX = 112;   // 128*24 + 512*112 = 60416 <= 640 * 96 (the launch allocation)
constexpr uint32_t Q_BYTES = BM * D * 2;      // 16 KB
constexpr uint32_t KV_BYTES = BN * D * 2;     // 8 KB
constexpr uint32_t SMEM_BYTES = 2 * NS * Q_BYTES + (KS + VS) * KV_BYTES + 1024;
constexpr uint32_t COL_S = 0, COL_O = 256;
constexpr float RESCALE_LOG2 = 8.0f;
}  // namespace sa4

struct Sa4Params {
  int N, heads, frames;
  int n_qt;            // query tiles per (frame, head)
  int n_groups;        // groups of up to 4 query tiles per (frame, head)
  int n_items;         // frames * heads * n_groups
  int n_kv;            // 64-key tiles
  void* out;
};

__device__ __forceinline__ int sa4_kv_cols(const Sa4Params& p, int j) {   // columns of key tile j, rounded up to 32
  const int valid = p.N - j * sa4::BN;
  return valid >= sa4::BN ? sa4::BN : ((valid + 31) & ~31);
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}

#ifndef VDA_SA4_POLY_MASK
#define VDA_SA4_POLY_MASK 3     // pair p of a row uses the FMA-pipe polynomial when (p & MASK) == MASK: 3 = 25 %, 1 = 50 %, 255 = none
#endif
// exp2 on the FMA / ALU pipes for a pair of values (Cody-Waite + degree-3 minimax, max rel. error 7.7e-5; see
// attention_spatial.cu)
__device__ __forceinline__ float2 sa4_exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));
  const float2 j = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(j, make_float2(-1.f, -1.f), x);
  float2 q = __ffma2_rn(make_float2(5.508868381e-02f, 5.508868381e-02f), f, make_float2(2.426040515e-01f, 2.426040515e-01f));
  q = __ffma2_rn(q, f, make_float2(6.932762417e-01f, 6.932762417e-01f));
  q = __ffma2_rn(q, f, make_float2(9.999289404e-01f, 9.999289404e-01f));
  float2 r;
  r.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return r;
}

template <int N>
__device__ __forceinline__ void sa4_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void sa4_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

template <typename T>
__global__ void __launch_bounds__(sa4::THREADS, 1)
spatial_attention4_kernel(const __grid_constant__ CUtensorMap tmQKV, const Sa4Params p) {
  using namespace sa4;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full[2], q_empty[2];
  __shared__ __align__(8) uint64_t k_full[KS], k_empty[KS], v_full[VS], v_empty[VS];
  __shared__ __align__(8) uint64_t s_ready[NS], o_ready[NS];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // smem map: Q[buffer][stream] (2 x 4 x 16 KB) | K[KS] | V[VS]
  const uint32_t offQ = 0, offK = 2 * NS * Q_BYTES, offV = offK + KS * KV_BYTES;
  const int stride = static_cast<int>(gridDim.x);
  const int n_my = (p.n_items - static_cast<int>(blockIdx.x) + stride - 1) / stride;   // items blockIdx.x + i * gridDim.x

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], NS); }
    for (int s = 0; s < KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], NS); }
    for (int s = 0; s < VS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], NS); }
    for (int t = 0; t < NS; ++t) {
      mbar_init(&s_ready[t], 1);
      mbar_init(&o_ready[t], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // item -> (frame, head, first query tile, number of query tiles in the group)
  auto decode = [&](int item, int& frame, int& head, int& tile0, int& nact) {
    const int fh = item / p.n_groups;
    const int g = item - fh * p.n_groups;
    frame = fh / p.heads;
    head = fh - frame * p.heads;
    tile0 = g * NS;
    nact = min(NS, p.n_qt - tile0);
  };

  if (warp < 4) {
    sa4_reg_dec<REGS_CTRL>();
    if (warp == 0 && lane == 0) {
      // ===================================== TMA producer =====================================
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int i = 0; i < n_my; ++i) {
        int frame, head, tile0, nact;
        decode(static_cast<int>(blockIdx.x) + i * stride, frame, head, tile0, nact);
        const int qb = i & 1;
        mbar_wait(&q_empty[qb], ((i >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&q_full[qb], static_cast<uint32_t>(nact) * Q_BYTES);
        for (int t = 0; t < nact; ++t) {     // a 128-row Q tile = two 64-row boxes (same SWIZZLE_128B K-major layout)
          uint8_t* sQ = smem_gen + offQ + (qb * NS + t) * Q_BYTES;
          tma_load_4d(sQ, &tmQKV, &q_full[qb], 0, head, (tile0 + t) * BM, frame);
          tma_load_4d(sQ + KV_BYTES, &tmQKV, &q_full[qb], 0, head, (tile0 + t) * BM + BN, frame);
        }
        if (nact < NS) mbar_arrive_n(&q_empty[qb], static_cast<uint32_t>(NS - nact));   // absent streams never read Q
        for (int j = 0; j < p.n_kv; ++j) {
          mbar_wait(&k_empty[ks], kph ^ 1u);
          mbar_arrive_expect_tx(&k_full[ks], KV_BYTES);
          tma_load_4d(smem_gen + offK + ks * KV_BYTES, &tmQKV, &k_full[ks], 0, p.heads + head, j * BN, frame);
          if (nact < NS) mbar_arrive_n(&k_empty[ks], static_cast<uint32_t>(NS - nact));  // (counts for the phase that has just begun)
          if (++ks == KS) { ks = 0; kph ^= 1u; }
          mbar_wait(&v_empty[vs], vph ^ 1u);
          mbar_arrive_expect_tx(&v_full[vs], KV_BYTES);
          tma_load_4d(smem_gen + offV + vs * KV_BYTES, &tmQKV, &v_full[vs], 0, 2 * p.heads + head, j * BN, frame);
          if (nact < NS) mbar_arrive_n(&v_empty[vs], static_cast<uint32_t>(NS - nact));
          if (++vs == VS) { vs = 0; vph ^= 1u; }
        }
      }
    }
  } else {
    // ===================================== softmax warpgroups (+ MMA issue) ==================
    sa4_reg_inc<REGS_SOFTMAX>();
    const int t = (warp - 4) >> 2;                    // stream handled by this warpgroup
    const int quad = warp & 3;                        // TMEM lane quadrant of this warp
    const bool issuer = quad == 0;                    // first warp of the warpgroup issues the stream's MMAs
    const int row = quad * 32 + lane;                 // query row inside the tile
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + COL_S + t * BN;
    const uint32_t tO = tmem_base + lane_base + COL_O + t * D;
    const uint32_t tS0 = tmem_base + COL_S + t * BN, tO0 = tmem_base + COL_O + t * D;   // MMA operand addresses (lane 0)
    const float sc = 0.125f * 1.4426950408889634f;    // d^-0.5 * log2(e)
    const float2 sc2 = make_float2(sc, sc);
    uint32_t cnt = 0, n_done = 0;                     // key-tile steps / items finished by this stream
    T* outp = reinterpret_cast<T*>(p.out);
    // issuer state: ring cursors of the S (K ring) and P V (V ring) issue streams; k_acc = items whose K uses are counted
    int ks = 0, vs = 0, k_acc = 0, k_issued = -1;
    uint32_t kph = 0, vph = 0;
    const uint32_t idesc_pv = umma_idesc(H16<T>::kUmmaFmt, D) | kIdescBMnMajor;

    auto item_active = [&](int i) {
      int frame, head, tile0, nact;
      decode(static_cast<int>(blockIdx.x) + i * stride, frame, head, tile0, nact);
      return t < nact;
    };
    // S(i, j) of this stream (issuer warp only; warp-uniform, one elected lane issues)
    auto issue_s = [&](int i, int j) {
      const int qb = i & 1;
      if (j == 0) {
        for (; k_acc < i; ++k_acc) {                  // items in between have no tile for this stream: count their uses
          const int k2 = ks + p.n_kv;
          kph ^= static_cast<uint32_t>(k2 / KS) & 1u;
          ks = k2 % KS;
        }
        mbar_wait(&q_full[qb], (i >> 1) & 1u);
        k_issued = i;
      }
      const uint32_t idesc = umma_idesc(H16<T>::kUmmaFmt, static_cast<uint32_t>(sa4_kv_cols(p, j)));
      const uint64_t da = umma_desc_sw128(smem_base + offQ + (qb * NS + t) * Q_BYTES);
      mbar_wait(&k_full[ks], kph);
      tc_fence_after();
      const uint64_t db = umma_desc_sw128(smem_base + offK + ks * KV_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_f16(tS0, da + 2u * k, db + 2u * k, idesc, k);
        umma_commit(&s_ready[t]);
        umma_commit(&k_empty[ks]);
        if (j == p.n_kv - 1) umma_commit(&q_empty[qb]);     // last S of the item: Q is free once these retire
      }
      __syncwarp();
      if (++ks == KS) { ks = 0; kph ^= 1u; }
      if (j == p.n_kv - 1) k_acc = i + 1;
    };

    for (int i = 0; i < n_my; ++i) {
      int frame, head, tile0, nact;
      decode(static_cast<int>(blockIdx.x) + i * stride, frame, head, tile0, nact);
      if (t >= nact) {                                 // no query tile for this stream: only count the V ring uses
        const int v2 = vs + p.n_kv;
        vph ^= static_cast<uint32_t>(v2 / VS) & 1u;
        vs = v2 % VS;
        continue;
      }
      if (issuer && k_issued != i) issue_s(i, 0);      // (normally issued ahead, at the end of the previous item)
      float m_run = 0.f, l_run = 0.f;
      for (int j = 0; j < p.n_kv; ++j, ++cnt) {
        const int cols = sa4_kv_cols(p, j);
        const int valid = p.N - j * BN;
        uint32_t s[BN];
        mbar_wait(&s_ready[t], cnt & 1u);
        tc_fence_after();
        tmem_ld32(tS, s);
        if (cols > 32) tmem_ld32(tS + 32, s + 32);
        tmem_ld_wait();
        if (valid < BN) {                             // last, partial key tile: mask the keys beyond N
#pragma unroll
          for (int c = 0; c < BN; ++c)
            if (c >= valid) s[c] = 0xff800000u;       // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < BN; c += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[c]));
          mx1 = fmaxf(mx1, __uint_as_float(s[c + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[c + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(s[c + 3]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (j == 0) {
          m_run = mx;
        } else {
          // "S(j) is complete" implies "O += P(j-1) V(j-1) is complete" (issued earlier by the same thread): O is ours
          const float m_new = fmaxf(m_run, mx);
          const bool need = (m_new - m_run) * sc > RESCALE_LOG2;
          if (__any_sync(0xffffffffu, need)) {
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
              for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
              tmem_st16(tO + c, o);
            }
            tmem_st_wait();
          }
        }
        // ---- P = exp2(s*sc - m*sc), row sum, pack to 16 bit, store over the S columns ----
        const float nmb = -m_run * sc;
        const float2 nmb2 = make_float2(nmb, nmb);
        float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
        auto exp_chunk = [&](int c) {
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            float2 x0 = __ffma2_rn(make_float2(__uint_as_float(s[c + k]), __uint_as_float(s[c + k + 1])), sc2, nmb2);
            float2 x1 =
                __ffma2_rn(make_float2(__uint_as_float(s[c + k + 2]), __uint_as_float(s[c + k + 3])), sc2, nmb2);
            if ((((c + k) >> 1) & VDA_SA4_POLY_MASK) == VDA_SA4_POLY_MASK) x0 = sa4_exp2_poly2(x0);
            else { x0.x = exp2f(x0.x); x0.y = exp2f(x0.y); }
            if (((((c + k) >> 1) + 1) & VDA_SA4_POLY_MASK) == VDA_SA4_POLY_MASK) x1 = sa4_exp2_poly2(x1);
            else { x1.x = exp2f(x1.x); x1.y = exp2f(x1.y); }
            la = __fadd2_rn(la, x0);
            lb = __fadd2_rn(lb, x1);
            pk[k >> 1] = H16<T>::pack2(x0.x, x0.y);
            pk[(k >> 1) + 1] = H16<T>::pack2(x1.x, x1.y);
          }
          tmem_st16(tS + (c >> 1), pk);
        };
        exp_chunk(0);
        if (cols > 32) exp_chunk(32);
        l_run += (la.x + la.y) + (lb.x + lb.y);
        tmem_st_wait();
        tc_fence_before();
        named_bar_sync(1 + t, 128);                   // every row of P(j) is in TMEM
        if (issuer) {
          // O += P(j) V(j), then the next S right behind it (same issuing thread: executed in order, so the S / P
          // columns are overwritten only after P V has read them)
          tc_fence_after();
          const int nk = cols >> 4;
          mbar_wait(&v_full[vs], vph);
          const uint64_t db = umma_desc_sw128_mn(smem_base + offV + vs * KV_BYTES);
          if (elect_one()) {
            for (int k = 0; k < nk; ++k)    // 16 keys per MMA: 8 TMEM columns of P, 16 rows (2048 B) of V
              umma_f16_ts(tO0, tS0 + 8u * k, db + 128u * k, idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
            umma_commit(&v_empty[vs]);
            if (j == p.n_kv - 1) umma_commit(&o_ready[t]);
          }
          __syncwarp();
          if (++vs == VS) { vs = 0; vph ^= 1u; }
          if (j + 1 < p.n_kv) {
            issue_s(i, j + 1);
          } else {                                     // first S of this stream's next item (Q is double-buffered)
            int i2 = i + 1;
            while (i2 < n_my && !item_active(i2)) ++i2;
            if (i2 < n_my) issue_s(i2, 0);
          }
        }
      }
      // ---- epilogue: O / l -> 16 bit -> global ----
      mbar_wait(&o_ready[t], n_done & 1u);
      ++n_done;
      tc_fence_after();
      const float inv = 1.f / l_run;
      const int q = (tile0 + t) * BM + row;
      T* orow = outp + (static_cast<long long>(frame) * p.N + q) * (static_cast<long long>(p.heads) * D) + head * D;
#pragma unroll
      for (int c = 0; c < D; c += 32) {
        uint32_t o[32];
        tmem_ld32(tO + c, o);
        tmem_ld_wait();
        if (q < p.N) {
#pragma unroll
          for (int k = 0; k < 32; k += 8) {
            uint4 u;
            u.x = H16<T>::pack2(__uint_as_float(o[k]) * inv, __uint_as_float(o[k + 1]) * inv);
            u.y = H16<T>::pack2(__uint_as_float(o[k + 2]) * inv, __uint_as_float(o[k + 3]) * inv);
            u.z = H16<T>::pack2(__uint_as_float(o[k + 4]) * inv, __uint_as_float(o[k + 5]) * inv);
            u.w = H16<T>::pack2(__uint_as_float(o[k + 6]) * inv, __uint_as_float(o[k + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c + k) = u;
          }
        }
      }
      tc_fence_before();   // the O reads above are ordered before the next item's first P V (behind the next named barrier)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// host launcher (called from vda_attention_spatial)
int sa4_launch(const void* qkv, void* out, int frames, int N, int heads, int dtype, cudaStream_t st) {
  CUtensorMap tm;
  const cuuint64_t row_bytes = static_cast<cuuint64_t>(3) * heads * sa4::D * 2;
  cuuint64_t dims[4] = {sa4::D, static_cast<cuuint64_t>(3 * heads), static_cast<cuuint64_t>(N),
                        static_cast<cuuint64_t>(frames)};
  cuuint64_t strides[3] = {sa4::D * 2, row_bytes, row_bytes * N};
  cuuint32_t box[4] = {sa4::D, 1, sa4::BN, 1};      // 64-row boxes: K / V tiles, and Q tiles as two boxes
  if (make_tensor_map(&tm, dtype, qkv, 4, dims, strides, box)) return 1;
  Sa4Params p;
  p.N = N; p.heads = heads; p.frames = frames;
  p.n_qt = (N + sa4::BM - 1) / sa4::BM;
  p.n_groups = (p.n_qt + sa4::NS - 1) / sa4::NS;
  p.n_items = frames * heads * p.n_groups;
  p.n_kv = (N + sa4::BN - 1) / sa4::BN;
  p.out = out;
  const int grid = p.n_items < sm_count() ? p.n_items : sm_count();
  {
    // setmaxnreg.inc can only draw on registers this CTA was launched with: a smaller launch allocation than planned
    // would leave the softmax warpgroups waiting for registers forever -- refuse instead
    cudaFuncAttributes fa;
    VDA_CUDA(cudaFuncGetAttributes(&fa, dtype == VDA_BF16 ? reinterpret_cast<const void*>(spatial_attention4_kernel<__nv_bfloat16>)
                                                         : reinterpret_cast<const void*>(spatial_attention4_kernel<__half>)));
    VDA_CHECK(fa.numRegs * sa4::THREADS >= 128 * sa4::REGS_CTRL + 512 * sa4::REGS_SOFTMAX,
              "attention4: launch allocation of %d registers/thread cannot feed setmaxnreg %d/%d", fa.numRegs, sa4::REGS_CTRL,
              sa4::REGS_SOFTMAX);
  }
  if (dtype == VDA_BF16) {
    auto k = spatial_attention4_kernel<__nv_bfloat16>;
    VDA_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(k), sa4::SMEM_BYTES));
    k<<<grid, sa4::THREADS, sa4::SMEM_BYTES, st>>>(tm, p);
  } else {
    auto k = spatial_attention4_kernel<__half>;
    VDA_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(k), sa4::SMEM_BYTES));
    k<<<grid, sa4::THREADS, sa4::SMEM_BYTES, st>>>(tm, p);
  }
  VDA_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vda
