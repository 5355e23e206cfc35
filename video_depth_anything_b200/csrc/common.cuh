// Shared device helpers for the sm_100a kernels: mbarrier / TMA / tcgen05 / TMEM PTX wrappers,
// 16-bit type traits, error handling.  Hand-written inline PTX; no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vda {

// ---------------------------------------------------------------------------------------------
// error plumbing (C ABI never throws; message retrievable with vda_last_error)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define VDA_CHECK(cond, ...)                      \
  do {                                            \
    if (!(cond)) {                                \
      vda::set_error(__VA_ARGS__);                \
      return 1;                                   \
    }                                             \
  } while (0)
#define VDA_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      vda::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                     __LINE__);                                                         \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

// ---------------------------------------------------------------------------------------------
// 16-bit operand types (bf16 or fp16: same tcgen05 kind::f16 rate, switchable at run time)
// ---------------------------------------------------------------------------------------------
template <typename T> struct H16;
template <> struct H16<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static constexpr uint32_t kUmmaFmt = 1;  // UMMA F16F32Format::BF16
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  __device__ static __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ static __forceinline__ float2 unpack2(uint32_t u) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(h);
  }
};
template <> struct H16<__half> {
  using T2 = __half2;
  static constexpr uint32_t kUmmaFmt = 0;  // UMMA F16F32Format::F16
  __device__ static __forceinline__ float to_f(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  __device__ static __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ static __forceinline__ float2 unpack2(uint32_t u) {
    __half2 h = *reinterpret_cast<__half2*>(&u);
    return __half22float2(h);
  }
};

// erf by Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7, far below one 16-bit output ulp): branch-free,
// ~10 FMA-pipe instructions + 2 MUFU, versus ~35 for erff -- the GELU epilogue of the fc1 GEMM is ALU-bound.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  float p = 1.061405429f;
  p = fmaf(p, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = __expf(-ax * ax);
  return copysignf(fmaf(-p, e, 1.0f), x);
}
__device__ __forceinline__ float gelu_erf(float x) {  // nn.GELU() default (erf form)
  return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f));
}
// Packed (2 x fp32 per instruction) erf-GELU: 0.5 x + 0.5 |x| erf(|x| / sqrt 2)  (x erf(x/sqrt2) is even, so no
// sign fix-up); erf by Abramowitz & Stegun 7.1.26 with the 1/sqrt2 folded into the constants.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// erf-GELU x Phi(x) for a pair, evaluated as x * sigmoid(2u(x)) with 2u(x) = 2 atanh(erf(x / sqrt 2)) = x q(x^2), q an
// even polynomial of degree 8 fitted on |x| <= 7.5 (beyond that the sigmoid is saturated and any growing q is exact to
// fp32): max abs error 3.7e-6, max relative error 2e-5 (a 16-bit output ulp is >= 4.9e-4 relative).  The constants
// carry the factor -log2(e), so Phi = 1 / (1 + 2^t): 8 packed FMA-pipe operations + 4 MUFU per element pair,
// versus 15 + 4 for the Abramowitz-Stegun erf used before (the fc1 epilogue is instruction-bound).
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 x2 = __fmul2_rn(x, x);
  float2 q = __ffma2_rn(make_float2(-3.291989515e-06f, -3.291989515e-06f), x2, make_float2(8.931364573e-05f, 8.931364573e-05f));
  q = __ffma2_rn(q, x2, make_float2(3.548166424e-04f, 3.548166424e-04f));
  q = __ffma2_rn(q, x2, make_float2(-1.052177921e-01f, -1.052177921e-01f));
  q = __ffma2_rn(q, x2, make_float2(-2.302048445e+00f, -2.302048445e+00f));
  const float2 t = __fmul2_rn(q, x);
  const float2 d = __fadd2_rn(make_float2(exp2f(t.x), exp2f(t.y)), make_float2(1.f, 1.f));
  return __fmul2_rn(x, make_float2(rcp_approx(d.x), rcp_approx(d.y)));
}
// Same GELU through one MUFU per element: x Phi(x) = 0.5 x (1 + tanh(u)), u = 0.5 x q(x^2) (sigmoid(2u) = (1 + tanh u) / 2).
// `tanh.approx.f32` carries a relative error of 2^-11, a quarter of a bf16 output ulp -- used for bf16 outputs only (it
// would double the rounding noise of fp16 outputs, which have 11 mantissa bits themselves).  6 packed ops + 2 MUFU per pair.
__device__ __forceinline__ float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float2 gelu_tanh2(float2 x) {
  const float2 x2 = __fmul2_rn(x, x);
  // coefficients of gelu_erf2's q times -1 / (2 log2 e): u = 0.5 x q_e(x^2) in natural units
  float2 q = __ffma2_rn(make_float2(1.140916610e-06f, 1.140916610e-06f), x2, make_float2(-3.095375195e-05f, -3.095375195e-05f));
  q = __ffma2_rn(q, x2, make_float2(-1.229700730e-04f, -1.229700730e-04f));
  q = __ffma2_rn(q, x2, make_float2(3.646570705e-02f, 3.646570705e-02f));
  q = __ffma2_rn(q, x2, make_float2(7.978291900e-01f, 7.978291900e-01f));
  const float2 u = __fmul2_rn(q, x);
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(h, make_float2(tanh_approx(u.x), tanh_approx(u.y)), h);
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, mbarrier completion
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA store (smem -> global, bulk async-group completion) and the fences around it
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups are still READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// generic-proxy writes to shared memory (st.shared) become visible to the async proxy (TMA) of this CTA
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a converged warp (ELECT): code under `if (elect_one())` in otherwise warp-uniform control flow keeps
// its descriptors in uniform registers, whereas a `lane == 0` branch makes ptxas wrap every UTCHMMA / UTMALDG in a
// per-active-lane loop (R2UR + BRA.U.ANY, ~50 cycles of issue overhead per instruction).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "@P mov.u32 %0, 1;\n\t"
      "}"
      : "+r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a cluster on one TPC issue one UMMA with M = 256.  Each CTA holds its own
// 128 rows of A, half of B's N rows and its 128 lanes of the accumulator; CTA 0 (the leader) issues.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same object in CTA 0 of the pair
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {  // whole warp, both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {  // whole warp, both CTAs
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA loads of a CTA pair: the transaction bytes are credited to the LEADER's mbarrier (same offset, CTA 0)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued pair-MMAs have completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// arrive on the barrier at this offset in the leader CTA (rank 0) of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint32_t umma_idesc_m(uint32_t fmt, uint32_t n, uint32_t m) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// K-major, 128-byte-swizzled operand tile: rows of 64 16-bit elements (128 B), 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);  // start address   [0,14)
  d |= static_cast<uint64_t>(1) << 16;                  // LBO (unused for swizzled K-major) [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;          // SBO = 1024 B    [32,46)
  d |= static_cast<uint64_t>(1) << 46;                  // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                  // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: fp32 accumulate, A/B K-major, M=128
__device__ __forceinline__ uint32_t umma_idesc(uint32_t fmt, uint32_t n) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait for outstanding tcgen05.ld; the registers are passed as in/out operands so the compiler cannot move a use
// of the (asynchronously written) destination registers above the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// 32 lanes x 32 columns of 32-bit cells: thread i of the warp owns TMEM lane (quadrant base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// wait for outstanding tcgen05.ld; the 32 destination registers are in/out operands so that no use of them can be
// scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait32(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                 "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                 "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
// named barriers (ids 1..15; 0 is __syncthreads): producer/consumer hand-off between warpgroups
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int threads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]: A is an M x K tile held in TMEM (lane = row, two 16-bit K elements per cell)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MN-major, 128-byte-swizzled operand tile: rows of 64 16-bit MN elements (128 B) per K index, 8-row (8 K) groups
// 1024 B apart (SBO); LBO (stride between 64-element MN chunks) is unused when the MN extent is 64.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
constexpr uint32_t kIdescBMnMajor = 1u << 16;   // instruction-descriptor bit: B operand is MN-major ("transposed")

// host helpers shared by the TMA-fed kernels (defined in gemm.cu)
int make_tensor_map(CUtensorMap* m, int dtype, const void* base, int rank, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box);
constexpr int kTmapF32 = 100;          // `dtype` of make_tensor_map for fp32 tensors (epilogue residual / output tiles)
constexpr int kTmapSwizzle64 = 1 << 10;  // OR-ed into `dtype`: SWIZZLE_64B instead of SWIZZLE_128B (boxes with 64-byte rows)
int sm_count();        // SM count of the CURRENT device (cached per device)
int current_device();  // cudaGetDevice, -1 on error

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device function attribute: the opt-in is cached per
// (kernel, device), so a second GPU in the same process gets its own cudaFuncSetAttribute call (defined in gemm.cu).
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes);

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may be
// scheduled while its predecessor still runs; pdl_wait() blocks until the predecessor grid has completed and its
// memory is visible (no-op for a normal launch), pdl_trigger() lets the successor be scheduled from now on.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
}  // namespace vda (host helper below needs <cstdlib>)
#include <cstdlib>
namespace vda {
// Host side: PDL is on unless VDA_PDL=0.  Only the GEMM launches carry the attribute (A/B on one box, 8 steps each,
// alternating: 612 -> 622 and 622 -> 624 frames/s); giving it to the LayerNorm and attention launches as well measured
// +0.1 / +0.3 % on another box, within noise, and was dropped.
inline bool pdl_enabled() {
  static const bool on = []() { const char* e = getenv("VDA_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
// Launch `kernel` with the programmatic-stream-serialization attribute (when enabled): the kernel MUST call pdl_wait()
// before it touches anything its predecessor wrote.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace vda
