// Spatial ViT attention (dinov2_layers/attention.py:49-62) as a tcgen05 / TMEM / TMA flash-attention kernel.
//
//   out[f, n, h, :] = softmax_n'( q[f,n,h,:] . k[f,n',h,:] / sqrt(64) ) v[f,n',h,:]       d = 64, N ~ 1370..2443
//
// Persistent kernel, one CTA (384 threads = 3 warpgroups) per SM; a work item is (frame, head, pair of 128-query
// tiles); the per-CTA item list is processed as ONE flat stream of key-tile steps (S of the next item's first key
// tile is issued while the softmax of the current item's last tile runs; Q is double-buffered):
//   warp 0      TMA producer: Q tiles once per item, K / V tiles (128 keys x 64) through two 3-slot rings
//   warp 1      tcgen05.mma issuer + TMEM owner:  S = Q K^T  (SS form, 128 x kv x 64)  and  O += P V  (TS form: P is
//               read from TMEM, V from smem as an MN-major operand -- no transpose of V anywhere)
//   warps 2-3   idle (they only donate registers: setmaxnreg.dec on warpgroup 0, .inc on the softmax warpgroups)
//   warps 4-7   softmax warpgroup of query tile A      \  one thread per query row: the 128 scores of a row are read
//   warps 8-11  softmax warpgroup of query tile B      /  from TMEM straight into that thread's registers (no shuffles)
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
constexpr int THREADS = 384;
constexpr int REGS_CTRL = 56, REGS_SOFTMAX = 224;   // setmaxnreg budgets: 128*56 + 256*224 <= 64K
constexpr uint32_t TILE_BYTES = BM * D * 2;   // 16 KB: one Q, K or V tile
constexpr uint32_t SMEM_BYTES = (4 + KS + VS) * TILE_BYTES + 1024;   // Q: 2 buffers x 2 tiles
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

#ifdef VDA_SA_TIMING
// debug build only: per-phase cycle sums of the softmax warps of CTA 0 (tools/bench_attention.py timing)
__device__ unsigned long long g_sa_timing[3][8];
#define SA_T(i) do { const long long _t = clock64(); tacc[i] += _t - tprev; tprev = _t; } while (0)
#else
#define SA_T(i) do { } while (0)
#endif

// Exp-phase hand-over token between the two softmax warpgroups (-DVDA_SA_TOKEN): off by default.  Micro-benchmarks
// (tools/microbench/exp_phase.cu, mufu_issue.cu) show that a lone warp per scheduler issues one MUFU.EX2 per ~9.7
// cycles (1315 cycles per 128x128 tile for the kernel's instruction mix) while two warps per scheduler share the
// dispatch port and need ~2000 cycles for two tiles whatever the MUFU / polynomial split, so exclusive ownership of
// the MUFU buys nothing any more; letting the warpgroups drift freely measured 0.385 vs 0.396 ms per layer.
#ifdef VDA_SA_TOKEN
#define SA_TOKEN(x) x
#else
#define SA_TOKEN(x) do { } while (0)
#endif

// exp2 on the FMA / ALU pipes for a pair of values (Cody-Waite: x = j + f, j = round(x), f in [-0.5, 0.5];
// 2^f by a degree-3 minimax polynomial, max rel. error 7.7e-5 -- below the 16-bit rounding of P; 2^j by adding j
// to the exponent field).  A quarter of the exponentials of a score tile take this path so that the MUFU (16 ex2/clk/SM)
// is no longer the only unit that can produce them.
#ifndef VDA_SA_POLY_MASK
#define VDA_SA_POLY_MASK 3     // pair p of a row uses the polynomial when (p & MASK) == MASK; 1: 50%, 3: 25%, 255: none
                               // (measured per layer: none 0.416 ms, 25% 0.395 ms, 50% 0.445 ms -- issue-slot bound beyond 25%)
#endif
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));          // 1.5 * 2^23: integer part in the low bits
  const float2 j = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(j, make_float2(-1.f, -1.f), x);   // x - j in one packed op (a negated operand costs two scalar FADDs)
  float2 q = __ffma2_rn(make_float2(5.508868381e-02f, 5.508868381e-02f), f, make_float2(2.426040515e-01f, 2.426040515e-01f));
  q = __ffma2_rn(q, f, make_float2(6.932762417e-01f, 6.932762417e-01f));
  q = __ffma2_rn(q, f, make_float2(9.999289404e-01f, 9.999289404e-01f));
  float2 r;
  r.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return r;
}

template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

template <typename T>
__global__ void __launch_bounds__(sa::THREADS, 1)
spatial_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const SaParams p) {
  using namespace sa;
  pdl_trigger();   // the proj GEMM that follows (launched with the PDL attribute) may be scheduled as our CTAs retire
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full[2], q_empty[2];
  __shared__ __align__(8) uint64_t k_full[KS], k_empty[KS], v_full[VS], v_empty[VS];
  __shared__ __align__(8) uint64_t s_ready[2], s_free[2], p_ready[2], o_done[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // smem map: Q[buffer 0: A, B | buffer 1: A, B] | K[KS] | V[VS]
  const uint32_t offQ = 0, offK = 4 * TILE_BYTES, offV = (4 + KS) * TILE_BYTES;
  const int n_my = (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                   static_cast<int>(gridDim.x);   // items of this CTA: blockIdx.x + i * gridDim.x

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
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

  if (warp < 4) {
    reg_dec<REGS_CTRL>();
    if (warp == 0 && lane == 0) {
      // ===================================== TMA producer =====================================
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int i = 0; i < n_my; ++i) {
        const SaItem it = sa_decode(p, blockIdx.x + i * gridDim.x);
        const int qb = i & 1;
        uint8_t* sQ = smem_gen + offQ + qb * 2 * TILE_BYTES;
        mbar_wait(&q_empty[qb], ((i >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&q_full[qb], it.has_b ? 2 * TILE_BYTES : TILE_BYTES);
        tma_load_4d(sQ, &tmQKV, &q_full[qb], 0, it.head, it.q0, it.frame);
        if (it.has_b) tma_load_4d(sQ + TILE_BYTES, &tmQKV, &q_full[qb], 0, it.head, it.q0 + BM, it.frame);
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
    } else if (warp == 1) {
      // ===================================== MMA issuer =======================================
      // The whole warp runs the (warp-uniform) control flow; one elected lane issues the tcgen05 instructions.
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      uint32_t n_s[2] = {0, 0}, n_pv[2] = {0, 0};     // running counts of S / PV tiles issued per query tile
      const uint32_t idesc_pv = umma_idesc(H16<T>::kUmmaFmt, D) | kIdescBMnMajor;
      const int n_full = p.frames * p.heads * p.n_pairs;   // items below this index have two query tiles
#ifdef VDA_SA_TIMING
      long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif

      // S(i, j) for the query tiles of the i-th item of this CTA
      auto issue_s_step = [&](int i, int j) {
        const int nt = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) < n_full ? 2 : 1;
        const int qb = i & 1;
        SA_T(7);
        if (j == 0) mbar_wait(&q_full[qb], (i >> 1) & 1u);
        mbar_wait(&k_full[ks], kph);
        SA_T(0);
        const int cols = sa_kv_cols(p, j);
        const uint32_t idesc = umma_idesc(H16<T>::kUmmaFmt, static_cast<uint32_t>(cols));
        const uint64_t db = umma_desc_sw128(smem_base + offK + ks * TILE_BYTES);
        for (int t = 0; t < nt; ++t) {
          mbar_wait(&s_free[t], (n_s[t] & 1u) ^ 1u);    // the warpgroup has copied the previous S to registers
          tc_fence_after();
          SA_T(1 + t);
          const uint64_t da = umma_desc_sw128(smem_base + offQ + (qb * 2 + t) * TILE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < D / 16; ++k) umma_f16(tmem_base + COL_S + t * BN, da + 2u * k, db + 2u * k, idesc, k);
            umma_commit(&s_ready[t]);
          }
          __syncwarp();
          ++n_s[t];
        }
        if (elect_one()) {
          umma_commit(&k_empty[ks]);
          if (j == p.n_kv - 1) umma_commit(&q_empty[qb]);   // last S of the item: Q is free once these retire
        }
        __syncwarp();
        if (++ks == KS) { ks = 0; kph ^= 1u; }
      };

      if (n_my > 0) issue_s_step(0, 0);
      for (int i = 0; i < n_my; ++i) {
        const int nt = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) < n_full ? 2 : 1;
        for (int j = 0; j < p.n_kv; ++j) {
          // S of the next step of the flat stream (next key tile, or the first key tile of the next item)
          if (j + 1 < p.n_kv) issue_s_step(i, j + 1);
          else if (i + 1 < n_my) issue_s_step(i + 1, 0);
          // O += P(j) V(j)
          const int nk = sa_kv_cols(p, j) >> 4;
          SA_T(7);
          mbar_wait(&v_full[vs], vph);
          SA_T(3);
          const uint64_t db = umma_desc_sw128_mn(smem_base + offV + vs * TILE_BYTES);
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&p_ready[t], n_pv[t] & 1u);
            tc_fence_after();
            SA_T(4 + t);
            if (elect_one()) {
              for (int k = 0; k < nk; ++k)   // 16 keys per MMA: 8 TMEM columns of P, 16 rows (2048 B) of V
                umma_f16_ts(tmem_base + COL_O + t * D, tmem_base + COL_P + t * (BN / 2) + 8u * k, db + 128u * k,
                            idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
              umma_commit(&o_done[t]);
            }
            __syncwarp();
            ++n_pv[t];
          }
          if (elect_one()) umma_commit(&v_empty[vs]);
          __syncwarp();
          if (++vs == VS) { vs = 0; vph ^= 1u; }
        }
      }
#ifdef VDA_SA_TIMING
      SA_T(7);
      if (blockIdx.x == 0 && lane == 0)
        for (int k = 0; k < 8; ++k) g_sa_timing[2][k] = static_cast<unsigned long long>(tacc[k]);
#endif
    }
  } else {
    // ===================================== softmax warpgroups ===============================
    reg_inc<REGS_SOFTMAX>();
    const int t = (warp - 4) >> 2;                    // query tile handled by this warpgroup
    const int quad = warp & 3;                        // TMEM lane quadrant of this warp
    const int row = quad * 32 + lane;                 // query row inside the tile
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + COL_S + t * BN;
    const uint32_t tO = tmem_base + lane_base + COL_O + t * D;
    const uint32_t tP = tmem_base + lane_base + COL_P + t * (BN / 2);
    const float sc = 0.125f * 1.4426950408889634f;    // d^-0.5 * log2(e)
    const float2 sc2 = make_float2(sc, sc);
    uint32_t cnt = 0;                                 // running key-tile counter of this query tile
    T* outp = reinterpret_cast<T*>(p.out);
#ifdef VDA_SA_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif

    // The exponentials are MUFU-bound and one warp per scheduler saturates the MUFU, so the two warpgroups take
    // turns: a token (named barriers 1 / 2) is handed over after each exp phase, and a warpgroup runs everything
    // else of its step (S load, row max, TMEM stores, barriers) while the other one owns the MUFU.
    const int tok_mine = 1 + t, tok_other = 2 - t;
    (void)tok_mine; (void)tok_other;
    if (t == 1) SA_TOKEN(named_bar_arrive(tok_other, 256));     // warpgroup A owns the first exp phase

    for (int i = 0; i < n_my; ++i) {
      const SaItem it = sa_decode(p, blockIdx.x + i * gridDim.x);
      if (t == 1 && !it.has_b) {                      // single-tile item: just pass the token along
        for (int j = 0; j < p.n_kv; ++j) {
          SA_TOKEN(named_bar_sync(tok_mine, 256));
          if (i + 1 < n_my || j + 1 < p.n_kv) SA_TOKEN(named_bar_arrive(tok_other, 256));   // (B's very last hand-over has no taker)
        }
        continue;
      }
      float m_run = 0.f, l_run = 0.f;
      for (int j = 0; j < p.n_kv; ++j, ++cnt) {
        const int cols = sa_kv_cols(p, j);
        const int valid = p.N - j * BN;
        uint32_t s[BN];
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        SA_T(7);
        mbar_wait(&s_ready[t], cnt & 1u);
        tc_fence_after();
        SA_T(0);
        if (valid >= BN) {
          // full key tile: the row maximum of chunk c is computed while chunk c+1 is still in flight from TMEM
          tmem_ld32(tS, s);
          tmem_ld_wait32(s);
#pragma unroll
          for (int c = 0; c < BN; c += 32) {
            if (c + 32 < BN) tmem_ld32(tS + c + 32, s + c + 32);
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              mx0 = fmaxf(mx0, __uint_as_float(s[c + k]));
              mx1 = fmaxf(mx1, __uint_as_float(s[c + k + 1]));
              mx2 = fmaxf(mx2, __uint_as_float(s[c + k + 2]));
              mx3 = fmaxf(mx3, __uint_as_float(s[c + k + 3]));
            }
            if (c + 32 < BN) tmem_ld_wait32(s + c + 32);
          }
          tc_fence_before();
          mbar_arrive(&s_free[t]);
        } else {
          // last, partial key tile: mask the keys beyond N
#pragma unroll
          for (int c = 0; c < BN; c += 32)
            if (c < cols) tmem_ld32(tS + c, s + c);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&s_free[t]);
#pragma unroll
          for (int c = 0; c < BN; ++c)
            if (c >= valid) s[c] = 0xff800000u;   // -inf
#pragma unroll
          for (int c = 0; c < BN; c += 32) {
            if (c < cols) {
#pragma unroll
              for (int k = 0; k < 32; k += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(s[c + k]));
                mx1 = fmaxf(mx1, __uint_as_float(s[c + k + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(s[c + k + 2]));
                mx3 = fmaxf(mx3, __uint_as_float(s[c + k + 3]));
              }
            }
          }
        }
        SA_T(1);
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (j > 0) {   // O += P(j-1) V(j-1) was issued a whole exp phase ago: P and O are ours again
          mbar_wait(&o_done[t], (cnt - 1u) & 1u);
          tc_fence_after();
        }
        if (j == 0) {
          m_run = mx;
        } else {
          const float m_new = fmaxf(m_run, mx);
          const bool need = (m_new - m_run) * sc > RESCALE_LOG2;
          if (__any_sync(0xffffffffu, need)) {
            // rare: rescale O (TMEM round trip by the owning warp)
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
        SA_T(2);
        // ---- exp phase (owns the MUFU): P = exp2(s*sc - m*sc), row sum, pack to 16 bit, store to TMEM ----
        SA_TOKEN(named_bar_sync(tok_mine, 256));
        SA_T(4);
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
            if ((((c + k) >> 1) & VDA_SA_POLY_MASK) == VDA_SA_POLY_MASK) x0 = exp2_poly2(x0);
            else { x0.x = exp2f(x0.x); x0.y = exp2f(x0.y); }
            if (((((c + k) >> 1) + 1) & VDA_SA_POLY_MASK) == VDA_SA_POLY_MASK) x1 = exp2_poly2(x1);
            else { x1.x = exp2f(x1.x); x1.y = exp2f(x1.y); }
            la = __fadd2_rn(la, x0);
            lb = __fadd2_rn(lb, x1);
            pk[k >> 1] = H16<T>::pack2(x0.x, x0.y);
            pk[(k >> 1) + 1] = H16<T>::pack2(x1.x, x1.y);
          }
          tmem_st16(tP + (c >> 1), pk);
        };
        if (cols == BN) {   // straight-line code for full key tiles: the scheduler hides the FMAs behind the MUFU
          exp_chunk(0); exp_chunk(32); exp_chunk(64); exp_chunk(96);
        } else {
#pragma unroll
          for (int c = 0; c < BN; c += 32)
            if (c < cols) exp_chunk(c);
        }
        if (t == 0 || i + 1 < n_my || j + 1 < p.n_kv) SA_TOKEN(named_bar_arrive(tok_other, 256));
        SA_T(3);
        l_run += (la.x + la.y) + (lb.x + lb.y);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[t]);
        SA_T(5);
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
      tc_fence_before();   // the O reads above are ordered before this thread's next p_ready arrive
      SA_T(6);
    }
#ifdef VDA_SA_TIMING
    if (blockIdx.x == 0 && quad == 0 && lane == 0)
      for (int k = 0; k < 8; ++k) g_sa_timing[t][k] = static_cast<unsigned long long>(tacc[k]);
#endif
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

#ifdef VDA_SA_TIMING
extern "C" int vda_debug_sa_timing(unsigned long long* host_out) {
  VDA_CUDA(cudaMemcpyFromSymbol(host_out, g_sa_timing, sizeof(unsigned long long) * 24));
  return 0;
}
#endif

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
