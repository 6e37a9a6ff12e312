// tcgen05 (5th-generation tensor core) GEMM for the pointwise / strided convolutions:
//     out[M, N] = A[M, K] . Bw[N, K]^T   (+ fused epilogue, same EpiMode contract as gemm_simt.cuh)
// and the weight-gradient product dW[N, K] += X[R, N]^T . Y[R, K] (gemm_tn_tc_kernel, contraction over rows).
//
// sm_100a only, written directly against the PTX ISA (no CUTLASS):
//   * TMA (cp.async.bulk.tensor, 128-byte swizzle) streams [128 x 32] fp32 A tiles and [BN x 32] weight tiles into a
//     shared-memory ring of up to 8 stages; out-of-range rows / K-tail columns are zero-filled by the TMA unit, so
//     ragged M, N and K (e.g. K = 40) need no padding of the activations in HBM;
//   * ONE elected thread issues tcgen05.mma.cta_group::1 (M = 128, N = BN <= 256; kind::f16 with K = 16 per instruction
//     for the default 3xBF16 mode, kind::tf32 with K = 8 otherwise), accumulating in TMEM; two accumulator stages
//     (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1; tcgen05.commit releases shared-memory
//     slots / publishes accumulators through mbarriers;
//   * 4 - 16 epilogue warps read the accumulator with tcgen05.ld (a lane quarter per warp, 16 columns at a time), apply
//     the fused epilogue (bias, GELU, residual, GRN statistics, GELU / LayerNorm backward), stage the [32 x 16] result
//     chunk in shared memory and hand it to the TMA unit (cp.async.bulk.tensor store);
//   * 2 - 16 operand-splitter warps turn the fp32 A tile into its (bf16 hi | bf16 lo) or (tf32 hi, lo) form in place;
//   * persistent CTAs (one per SM) walk the tile list with N fastest so the A tile is re-used from L2; the tiles
//     beyond the last full wave are cut into column slices for the otherwise idle CTAs.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, then the epilogue warps, then the splitters.
//
// Precision: backend 3 (default) = "3xBF16", D += Ahi.Bhi + Alo.Bhi + Ahi.Blo with bf16 pairs (see gemm_tc_kernel);
// backend 1 = "3xTF32", the same with TF32 pairs (weights pre-split by the caller); backend 2 = ONE TF32 pass
// (kind::tf32 keeps 10 mantissa bits of each operand).
#pragma once
#include <cuda.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "gemm_simt.cuh"

// Timing knobs (tools/dbg_sweep.py): compiled in only with -DMPMAE_TC_KNOBS (build.py: MPMAE_BUILD_KNOBS=1); the
// production build sees a constant 0 and the guarded code disappears.
#ifdef MPMAE_TC_KNOBS
#define TC_DBG(p) ((p).dbg)
// event trace of CTA 0 (tools/tc_trace.py): role r appends clock64() to trace[r * 256 + n++]
#define TC_TRACE_DECL int trace_n = 0
#define TC_TRACE(p, role) \
  do { if ((p).trace && blockIdx.x == 0 && trace_n < 255) { (p).trace[(role) * 256 + 1 + trace_n] = (unsigned long long)clock64(); (p).trace[(role) * 256] = ++trace_n; } } while (0)
#else
#define TC_DBG(p) 0
#define TC_TRACE_DECL
#define TC_TRACE(p, role) do { } while (0)
#endif

namespace mpmae {
namespace tc {

constexpr int BM = 128, BK = 32, STAGES = 8, ACC_STAGES = 2;
constexpr uint64_t kSpinLimit = 4000000000ull;  // ~2 s of SM clocks: a lost barrier traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: the warp sleeps until the phase flips
      : "memory");                                         // (measured: 0 .. 10 ms hints and plain spinning time the same)
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if ((uint64_t)(clock64() - t0) > kSpinLimit) __trap();
  }
}
// Wait of a role that runs AHEAD of its consumer (the TMA producer on a full ring, the MMA issuer on a busy accumulator
// stage): its wake-up latency is off the critical path, so it sleeps between polls instead of spinning.  The ncu source page
// of round 1's pw1 kernel showed 15 % of all issued instructions in the spin loops of these two warps (try_wait + clock read
// + branch), issue slots taken from the epilogue warps that bound the kernel.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity, bool spin = false) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (!spin) __nanosleep(256);
    if ((uint64_t)(clock64() - t0) > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// DRAM -> L2 prefetch of a contiguous block (16-byte aligned, size a multiple of 16): no destination, no completion
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, both K-major, fp32 storage consumed as TF32
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same for bf16 operands (kind::f16: 16 elements = 32 bytes of K per instruction, twice the TF32 rate)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (=1),
// descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B).  cute/arch/mma_sm100_desc.hpp field layout.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t make_idesc(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// kind::f16 with bf16 A and B, fp32 accumulate
__device__ __forceinline__ uint32_t make_idesc_bf16(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// fp32 -> (bf16 hi, bf16 lo) with hi + lo = x up to 2^-17 |x|; two values per 32-bit word, low half = first element
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// issue only: the registers are defined after tmem_ld_wait16() (which names them, so no use can be hoisted above it)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 256-bit global store (sm_100+): one full 32-byte sector per lane per instruction
__device__ __forceinline__ void st_global_v8(float *ptr, const float4 &a, const float4 &b) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
               "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}

__device__ __forceinline__ void ld_global_v8(const float *ptr, float4 &a, float4 &b) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(ptr));
}

// Transposed warp reduction: every lane holds v[0..32); afterwards lane l returns sum over lanes of v[l].
// 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_colsum32(float *v, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

struct TcParams {
  GemmArgs g;
  int bn;        // N tile (multiple of 16, <= 256)
  int num_m, num_n, num_k;
  int stages;    // shared-memory ring depth (<= STAGES)
  int vec8;      // N % 8 == 0 and 32-byte aligned outputs: 256-bit stores
  int vec8_in;   // same for the prefetched per-element operand
  int tma_out;   // output tensor maps are valid: the fast path stores through TMA
  // wave-quantisation fix: the tiles beyond the last full wave of the persistent grid (tile index >= tail_start) are cut
  // into tail_S column slices of tail_ws columns each, one slice per otherwise idle CTA
  int tail_start, tail_items, tail_S, tail_ws;
  uint32_t num_n_magic, tail_S_magic;   // fast_div() constants of num_n and tail_S
  int l2_prefetch;   // the producer prefetches the epilogue's per-element operand of each item into L2
  int spin;          // the ahead-of-consumer roles spin instead of sleeping between polls (latency-bound shapes)
  unsigned long long *trace;   // knob builds only (TC_TRACE)
  uint32_t tmem_cols;
  int dbg;       // MPMAE_TC_DBG timing experiments (results invalid): 1 no stats, 2 no GELU, 4 no stores, 8 no tcgen05.ld,
                 // 16 no per-element operand loads, 32 no MMAs, 64 no operand split
};

__device__ __forceinline__ void tmem_ld16v(uint32_t taddr, float *v) { tmem_ld16(taddr, v); }

// Transposed warp reduction over 16 columns: lane l returns sum over all 32 lanes of v[l & 15].
__device__ __forceinline__ float warp_colsum16(float *v, int lane) {
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, then the epilogue warps, then (3xTF32 only) the
// operand-splitter warps.  GELU / statistics epilogues are latency bound and get 16 warps (4 per scheduler) with 2 splitter
// warps (their K is the narrow side); the plain store epilogue is light and its K is the wide side, so it runs 8 + 8.
// (Measured with the bf16-converting splitter: 16 + 2 is best for the forward GELU epilogue, 12 + 4 for the backward ones.)
// 18-20 warps = 5 per scheduler keeps 96 registers per thread.
// WIDE (K >= 256, e.g. the decoder block): measured best is 8 + 8 for the forward GELU epilogue and 12 + 4 for the backward
// ones (same as their narrow-K setting).
// AGELU (EPI_STORE only): the splitter warps also apply GELU and the GRN scale to every A element (pw2 of a sparse block
// reading the saved pre-activation); that is ~3x their work per element while the epilogue covers only N = C columns, so
// the roles become 4 epilogue + 16 splitter warps.
// (epilogue, splitter) warps per role set; -DTC_W_<set>_E= / _S= overrides for tuning runs (tools/role_sweep.sh)
#define TC_ROLE_DEFAULT(name, e, s) \
  constexpr int name##_E_default = e, name##_S_default = s;
TC_ROLE_DEFAULT(TC_W_AG, 4, 16)    // GELU + GRN scale on the A operand (pw2 of a sparse block)
TC_ROLE_DEFAULT(TC_W_ST, 8, 8)     // plain store / LayerNorm-backward epilogue
TC_ROLE_DEFAULT(TC_W_SQ, 16, 2)    // forward GELU + statistics, narrow K
TC_ROLE_DEFAULT(TC_W_SQW, 8, 8)    // ... wide K
TC_ROLE_DEFAULT(TC_W_BW, 12, 4)    // backward epilogues (DG, DH_GELU), narrow K
TC_ROLE_DEFAULT(TC_W_BWW, 12, 4)   // ... wide K
#ifndef TC_W_AG_E
#define TC_W_AG_E TC_W_AG_E_default
#endif
#ifndef TC_W_AG_S
#define TC_W_AG_S TC_W_AG_S_default
#endif
#ifndef TC_W_ST_E
#define TC_W_ST_E TC_W_ST_E_default
#endif
#ifndef TC_W_ST_S
#define TC_W_ST_S TC_W_ST_S_default
#endif
#ifndef TC_W_SQ_E
#define TC_W_SQ_E TC_W_SQ_E_default
#endif
#ifndef TC_W_SQ_S
#define TC_W_SQ_S TC_W_SQ_S_default
#endif
#ifndef TC_W_SQW_E
#define TC_W_SQW_E TC_W_SQW_E_default
#endif
#ifndef TC_W_SQW_S
#define TC_W_SQW_S TC_W_SQW_S_default
#endif
#ifndef TC_W_BW_E
#define TC_W_BW_E TC_W_BW_E_default
#endif
#ifndef TC_W_BW_S
#define TC_W_BW_S TC_W_BW_S_default
#endif
#ifndef TC_W_BWW_E
#define TC_W_BWW_E TC_W_BWW_E_default
#endif
#ifndef TC_W_BWW_S
#define TC_W_BWW_S TC_W_BWW_S_default
#endif
__host__ __device__ constexpr int pick2(int which, int e, int s) { return which == 0 ? e : s; }
__host__ __device__ constexpr int role_warps(int which, int mode, bool wide, bool agelu) {
  return agelu ? pick2(which, TC_W_AG_E, TC_W_AG_S)
         : mode == EPI_STORE ? pick2(which, TC_W_ST_E, TC_W_ST_S)
         : mode == EPI_GELU_SQ ? (wide ? pick2(which, TC_W_SQW_E, TC_W_SQW_S) : pick2(which, TC_W_SQ_E, TC_W_SQ_S))
                               : (wide ? pick2(which, TC_W_BWW_E, TC_W_BWW_S) : pick2(which, TC_W_BW_E, TC_W_BW_S));
}
__host__ __device__ constexpr int epi_warps(int mode, bool wide, bool agelu = false) { return role_warps(0, mode, wide, agelu); }
__host__ __device__ constexpr int split_warps(int mode, bool split, bool wide, bool agelu = false) {
  return !split ? 0 : role_warps(1, mode, wide, agelu);
}
__host__ __device__ constexpr int tc_threads(int mode, bool split, bool wide, bool agelu = false) {
  return 64 + 32 * epi_warps(mode, wide, agelu) + 32 * split_warps(mode, split, wide, agelu);
}
__host__ __device__ constexpr int out_arrays(int mode) { return mode == EPI_GELU_SQ ? 2 : 1; }
constexpr uint32_t kStageOutBytes = 32 * 16 * 4;   // one warp's [32 rows x 16 columns] output chunk

// 2D TMA store of a [32 x 16] fp32 chunk (64-byte swizzle) from shared memory; out-of-range rows / columns are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t addr, const float4 &v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// SPLIT = 3xTF32: A tiles are split in shared memory into a TF32-exact high part (in place) and the remainder
// (second buffer) by four splitter warps; the weight operand arrives pre-split (Bw = hi, Bw_lo = lo) as two TMA
// tiles; D += Ahi.Bhi + Alo.Bhi + Ahi.Blo.
struct WorkItem { int m_blk, n0, wn; };   // 128-row block, first column, column count of the MMA / epilogue
// x / d for 0 <= x * d < 2^32 with magic = floor(2^32 / d) + 1 (host side: div_magic): one IMAD.HI instead of the ~20
// instructions of an integer division -- every warp of the CTA runs get_work once per tile, and a stage-0 tile is only a
// few hundred instructions of epilogue work per warp
__device__ __forceinline__ int fast_div(int x, uint32_t magic) { return (int)__umulhi((uint32_t)x, magic); }
__device__ __forceinline__ bool get_work(const TcParams &p, int it, WorkItem &w) {
  const int idx = (int)blockIdx.x + it * (int)gridDim.x;
  if (idx < p.tail_start) {
    w.m_blk = p.num_n == 1 ? idx : fast_div(idx, p.num_n_magic);
    w.n0 = (idx - w.m_blk * p.num_n) * p.bn;
    w.wn = p.bn;
    return true;
  }
  const int t = idx - p.tail_start;
  if (t >= p.tail_items) return false;
  const int tq = p.tail_S > 1 ? fast_div(t, p.tail_S_magic) : t, tile = p.tail_start + tq;
  w.m_blk = p.num_n == 1 ? tile : fast_div(tile, p.num_n_magic);
  w.n0 = (tile - w.m_blk * p.num_n) * p.bn + (t - tq * p.tail_S) * p.tail_ws;
  w.wn = p.tail_ws;
  return true;
}

// BF16 (with SPLIT) = 3xBF16: a ring stage covers 32 elements of K.  ONE fp32 TMA box of A ([128 x 32], 128-byte rows,
// 128-byte swizzle) lands in the stage and the splitter warps convert every row IN PLACE into [32 bf16 high parts | 32 bf16
// remainders] -- still one 128-byte-swizzled row, so the same K-major descriptor addresses the high half (bytes 0..63) and
// the remainder half (bytes 64..127) of the tile by its start address alone.  The weight arrives pre-split in the same
// interleaved form ([N][K/32][hi 32 | lo 32] bf16, written by fold_kernel: one [bn x 128 B] box per stage).
// D += Ahi.Bhi + Alo.Bhi + Ahi.Blo with kind::f16 MMAs (two K = 16 instructions per product and stage); |error| <= 2^-16
// per product.  A row is converted by threads of ONE warp (no block-wide barrier), and a stage is 16 KB + bn * 128 B --
// half of what a 64-element stage took -- so the ring is 4..8 deep where it was 2: the TMA load, the conversion and the
// MMAs of consecutive k blocks overlap (profiles/r2_aa_trace.txt: with two stages they ran back to back).
template <int MODE, bool SPLIT, bool WIDE, bool BF16, bool AGELU = false>
__global__ void __launch_bounds__(tc_threads(MODE, SPLIT, WIDE, AGELU), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ CUtensorMap map_out,
               const __grid_constant__ CUtensorMap map_out2, const __grid_constant__ CUtensorMap map_bt,
               const __grid_constant__ CUtensorMap map_bt_lo, const TcParams p) {
  // programmatic dependent launch: this CTA may become resident while the previous grid drains; what does not read that
  // grid's results (barrier init, TMEM allocation, descriptor prefetch) runs before the wait
  pdl_trigger();
  constexpr int kEpiWarps = epi_warps(MODE, WIDE, AGELU), kEpiThreads = 32 * kEpiWarps;
  constexpr int kSplitWarps = split_warps(MODE, SPLIT, WIDE, AGELU), kSplitThreads = 32 * kSplitWarps;
  constexpr int kThreadsNoSplit = 64 + kEpiThreads;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a shared pointer
  const int bn = p.bn, nstage = p.stages;
  constexpr uint32_t a_bytes = BM * BK * 4;
  const uint32_t b_bytes = (uint32_t)bn * BK * 4;                    // also one [bn x (32 hi | 32 lo)] bf16 tile
  const uint32_t a_span = (SPLIT && !BF16) ? 2 * a_bytes : a_bytes;  // [A | Alo]; 3xBF16: one tile holds both halves
  const uint32_t stage_bytes = BF16 ? a_bytes + b_bytes : a_span + (SPLIT ? 2 : 1) * b_bytes;   // [A | Alo | B | Blo]
  uint8_t *stage_out = smem + (size_t)nstage * stage_bytes;          // [kEpiWarps][out_arrays][32 x 16] TMA-store staging
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(stage_out + (size_t)kEpiWarps * out_arrays(MODE) * kStageOutBytes);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *split_bar = empty_bar + STAGES;
  uint64_t *tfull_bar = split_bar + STAGES;
  uint64_t *tempty_bar = tfull_bar + ACC_STAGES;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tempty_bar + ACC_STAGES + 2);   // keeps the float4 vectors below 16-byte aligned
  float *colacc = reinterpret_cast<float *>(tmem_ptr + 4);           // [kTcGroups][bn]
  constexpr int kTcGroups = 4;                                       // a 128-row tile spans <= 4 groups of >= 32 rows
  float *colacc2 = colacc + kTcGroups * bn;                          // [bn]
  float *vec_bias_all = colacc2 + bn;                                // [num_n * bn] bias (zero padded)
  float *vec_kg_all = vec_bias_all + p.num_n * bn;                   // [num_n * bn] kg of group 0 (EPI_DH_GELU)
  float *statacc1 = vec_kg_all + p.num_n * bn;                       // [num_n * bn] kernel-long column sums (one group)
  float *statacc2 = statacc1 + p.num_n * bn;                         // [num_n * bn]
  float *vec_as_all = statacc2 + p.num_n * bn;                       // [num_n * bn] accumulator scale (EPI_DH_GELU)
  float *a_scale_s = vec_kg_all + p.num_n * bn;                      // AGELU (EPI_STORE): [32 | K] reduction scratch, A scale

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmArgs &g = p.g;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_out)) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&split_bar[s], kSplitWarps > 0 ? kSplitWarps : 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr, p.tmem_cols);
  pdl_wait();
  if (warp >= 2 && warp < 2 + kEpiWarps) {
    float grnb_cross = 0.f, grnb_den = 1.f;
    if (MODE == EPI_DH_GELU && g.grnb_ds) {   // backward of the batch-global GRN statistic: the cross term is a sum over all N
      grnb_den = __ldg(g.grnb_denom);
      float part = 0.f;
      for (int i = threadIdx.x - 64; i < g.N; i += kEpiThreads)
        part += __ldg(g.grnb_gamma + i) * __ldg(g.grnb_ds + i) * (__ldg(g.grnb_nx + i) * grnb_den);
      part = warp_sum(part);
      if (lane == 0) colacc[warp - 2] = part;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      float tot = 0.f;
      for (int i = 0; i < kEpiWarps; ++i) tot += colacc[i];
      grnb_cross = tot / ((float)g.N * grnb_den * grnb_den);
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    }
    if (MODE != EPI_STORE)
      for (int i = threadIdx.x - 64; i < (kTcGroups + 1) * bn; i += kEpiThreads) colacc[i] = 0.f;
    for (int i = threadIdx.x - 64; i < p.num_n * bn; i += kEpiThreads) {
      vec_bias_all[i] = (g.bias && i < g.N) ? __ldg(g.bias + i) : 0.f;
      if (MODE == EPI_DH_GELU) {
        if (g.grnb_ds) {
          float kgv = 0.f;
          if (i < g.N) {
            const float dsv = __ldg(g.grnb_ds + i), nx = __ldg(g.grnb_nx + i), gx = nx * grnb_den;
            const float dgx = __ldg(g.grnb_gamma + i) * dsv / grnb_den - grnb_cross;
            kgv = gx > 0.f ? dgx / gx : 0.f;
            if (blockIdx.x == 0) atomicAdd(g.grnb_dgamma + i, nx * dsv);
          }
          vec_kg_all[i] = kgv;
        } else {
          vec_kg_all[i] = (g.kg && i < g.N) ? __ldg(g.kg + i) : 0.f;
        }
        vec_as_all[i] = (g.acc_scale && i < g.N) ? __ldg(g.acc_scale + i) : 1.f;
      }
      if (MODE != EPI_STORE) { statacc1[i] = 0.f; statacc2[i] = 0.f; }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int total_tiles = p.num_m * p.num_n;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      TC_TRACE_DECL;
      // The epilogue reads one more [128 x wn] fp32 operand per item straight from global memory (residual / LayerNorm
      // input / saved pre-activation), one 64-byte piece per lane and chunk: ~24 KB in flight per SM, which at DRAM latency
      // is < 2 TB/s for the whole chip (the stage-0 da kernel reads 199 MB of `a` that way).  The producer pulls the item's
      // block into L2 when it issues the item's operand loads, a tile or two ahead of the epilogue.
      const float *pf_src = p.l2_prefetch != 1 ? nullptr
                            : (MODE == EPI_STORE) ? (g.ln_xhat ? g.ln_xhat : g.resid)
                            : (MODE == EPI_DG) ? g.aux : (MODE == EPI_DH_GELU) ? g.aux2 : nullptr;
      WorkItem w;
      for (int it = 0; get_work(p, it, w); ++it) {
        const int m_blk = w.m_blk;
        const bool tail = w.wn != bn;
        const uint32_t wb_bytes = (uint32_t)w.wn * BK * 4;   // bytes of one weight box of this item
        if (pf_src) {
          const int64_t r0 = (int64_t)m_blk * BM;
          const int rows = (int)(g.M - r0 < BM ? g.M - r0 : BM);
          const int cols = g.N - w.n0 < w.wn ? g.N - w.n0 : w.wn;
          if (cols == g.N) {
            bulk_prefetch_l2(pf_src + r0 * g.N, (uint32_t)rows * (uint32_t)g.N * 4u);
          } else {
            for (int r = 0; r < rows; ++r) bulk_prefetch_l2(pf_src + (r0 + r) * g.N + w.n0, (uint32_t)cols * 4u);
          }
        }
        for (int kb = 0; kb < p.num_k; ++kb) {
          mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, p.spin != 0);
          TC_TRACE(p, 0);
          uint8_t *sa = smem + (size_t)stage * stage_bytes;
          if (BF16) {
            mbar_expect_tx(&full_bar[stage], a_bytes + wb_bytes);
            tma_load_2d(sa, &map_a, &full_bar[stage], kb * BK, m_blk * BM);
            tma_load_2d(sa + a_bytes, tail ? &map_bt : &map_b, &full_bar[stage], kb * 64, w.n0);   // 64 bf16 = (hi | lo) of 32 k
          } else {
          mbar_expect_tx(&full_bar[stage], a_bytes + (SPLIT ? 2 : 1) * wb_bytes);
          tma_load_2d(sa, &map_a, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(sa + a_span, tail ? &map_bt : &map_b, &full_bar[stage], kb * BK, w.n0);
          if (SPLIT) tma_load_2d(sa + a_span + b_bytes, tail ? &map_bt_lo : &map_b_lo, &full_bar[stage], kb * BK, w.n0);
          }
          if (++stage == nstage) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      TC_TRACE_DECL;
      WorkItem w;
      for (int it = 0; get_work(p, it, w); ++it) {
        const uint32_t idesc = BF16 ? make_idesc_bf16(w.wn) : make_idesc(w.wn);
        mbar_wait_relaxed(&tempty_bar[as], aphase ^ 1, p.spin != 0);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * bn);
        for (int kb = 0; kb < p.num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          TC_TRACE(p, 1);
          if (SPLIT) mbar_wait(&split_bar[stage], phase);
          TC_TRACE(p, 2);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = BF16 ? sa + a_bytes : sa + a_span;
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
          if (BF16) {
            // a 128-byte row = [hi k0..31 | lo k0..31]: descriptor start + 0 / + 2 (x16 B) = the two K = 16 halves of the
            // high parts, + 4 / + 6 = of the remainders
            if (!(TC_DBG(p) & 32)) {
#pragma unroll
            for (int k = 0; k < 2; ++k) umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
#pragma unroll
            for (int k = 0; k < 2; ++k) umma_bf16(tmem_d, da + (uint64_t)(4 + 2 * k), db + (uint64_t)(2 * k), idesc, 1u);
#pragma unroll
            for (int k = 0; k < 2; ++k) umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(4 + 2 * k), idesc, 1u);
            }
          } else {
          if (!(TC_DBG(p) & 32)) {
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)   // 8 tf32 = 32 bytes per instruction: advance the start address by 2 (x16 B)
            umma_tf32(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          if (SPLIT && !(TC_DBG(p) & 32)) {
            const uint64_t dal = make_smem_desc(sa + a_bytes), dbl = make_smem_desc(sa + a_span + b_bytes);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) umma_tf32(tmem_d, dal + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) umma_tf32(tmem_d, da + (uint64_t)(2 * k), dbl + (uint64_t)(2 * k), idesc, 1u);
          }
          }
          umma_commit(&empty_bar[stage]);     // frees the smem slot once these MMAs have read it
          if (kb == p.num_k - 1) umma_commit(&tfull_bar[as]);
          if (++stage == nstage) { stage = 0; phase ^= 1; }
        }
        if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ================================================================ epilogue: kEpiWarps warps, TMEM lane quarter =
    // warp % 4, the warps of a quarter take alternate 16-column chunks.  Per-column vectors (bias, kg) sit in shared
    // memory; the per-element operand of the next chunk (residual / h / a) is prefetched into registers before the
    // accumulator chunk is read, so its HBM latency overlaps the math of the current chunk.
    const int q = warp & 3, part = (warp - 2) >> 2;
    constexpr int kParts = kEpiWarps / 4;
    const int et = threadIdx.x - 64;
    const float *pre_src = (MODE == EPI_STORE) ? g.resid : (MODE == EPI_DG) ? g.aux : (MODE == EPI_DH_GELU) ? g.aux2 : nullptr;
    const bool single_group = (int64_t)g.group_rows >= g.M;
    int as = 0;
    uint32_t aphase = 0;
    TC_TRACE_DECL;
    WorkItem w;
    for (int it = 0; get_work(p, it, w); ++it) {
      const int m_blk = w.m_blk;
      const int64_t m = (int64_t)m_blk * BM + q * 32 + lane;
      const bool row_ok = m < g.M;
      const int n_base = w.n0;
      const float *vec_bias = vec_bias_all + n_base, *vec_kg = vec_kg_all + n_base, *vec_as = vec_as_all + n_base;
      const bool has_out2 = MODE == EPI_GELU_SQ && g.out2 != nullptr;
      // statistics groups touched by this tile / warp / lane (all 0 with one group: the sparse blocks; the divisions are
      // skipped there -- they were a quarter of a stage-0 epilogue warp's instructions per tile)
      int64_t g_first = 0;
      int gw_lo = 0, gw_hi = 0, my_g = 0;
      if (MODE != EPI_STORE && !single_group) {
        const int64_t mw0 = (int64_t)m_blk * BM + q * 32;
        const int64_t mw_last = (mw0 + 31 < g.M ? mw0 + 31 : g.M - 1);
        g_first = ((int64_t)m_blk * BM) / g.group_rows;
        gw_lo = (int)(mw0 / g.group_rows - g_first);
        gw_hi = mw_last >= mw0 ? (int)(mw_last / g.group_rows - g_first) : gw_lo;
        my_g = row_ok ? (int)(m / g.group_rows - g_first) : gw_lo;
      }
      auto prefetch = [&](int c0, float4 *dst) {
        if (p.vec8_in) {
#pragma unroll
          for (int j = 0; j < 4; j += 2) {
            const int n = n_base + c0 + 4 * j;
            if (pre_src && row_ok && c0 < bn && n < g.N) ld_global_v8(pre_src + m * g.N + n, dst[j], dst[j + 1]);
            else dst[j] = dst[j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = n_base + c0 + 4 * j;
            dst[j] = (pre_src && row_ok && c0 < bn && n < g.N) ? __ldg(reinterpret_cast<const float4 *>(pre_src + m * g.N + n))
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * bn);
      const int ncols = (g.N - n_base < w.wn) ? g.N - n_base : w.wn;
      // ---- fast path: full 128-row tile, one statistics group.  No per-element predicates; every warp stages its
      // [32 rows x 16 columns] result chunk in shared memory (64-byte swizzle: conflict-free 16-byte stores) and one lane
      // hands it to the TMA unit (cp.async.bulk.tensor store: full-sector writes, no LSU work, columns beyond N clipped);
      // column statistics accumulate in shared memory for the whole kernel (flushed once at the end); the next chunk's
      // accumulator (tcgen05.ld) and per-element operand are in flight while this one is computed.
      if (p.tma_out && (single_group || MODE == EPI_GELU_SQ || MODE == EPI_DG) && (int64_t)(m_blk + 1) * BM <= g.M &&
          (!pre_src || ((ncols & 7) == 0 && p.vec8_in))) {
        const int64_t row_off = m * (int64_t)g.N + n_base;
        const float *prep = pre_src ? pre_src + row_off : nullptr;
        const uint32_t sbuf = smem_u32(stage_out) + (uint32_t)(warp - 2) * out_arrays(MODE) * kStageOutBytes;
        const uint32_t srow = sbuf + (uint32_t)lane * 64u, sx = (uint32_t)(lane >> 1) & 3u;
        const int row0 = m_blk * BM + q * 32;
        auto load_pre = [&](int c0, float4 *dst) {
          if (prep && !(TC_DBG(p) & 16)) {
            ld_global_v8(prep + c0, dst[0], dst[1]);
            if (c0 + 8 < ncols) ld_global_v8(prep + c0 + 8, dst[2], dst[3]);
          }
        };
        if (MODE == EPI_STORE && g.ln_rstd) {
          // fused LayerNorm backward (the product row is dvhat): pass 1 reads the whole row for the two row means (each of
          // the kParts warps of a lane quarter does it redundantly: no cross-warp exchange), pass 2 writes this warp's chunks
          const float *xh = g.ln_xhat + row_off;
          mbar_wait(&tfull_bar[as], aphase);
          tc_fence_after();
          float s1 = 0.f, s2 = 0.f;
          for (int cc = 0; cc < ncols; cc += 16) {
            float v[16];
            float4 hq[4];
            tmem_ld16(taddr + cc, v);
            ld_global_v8(xh + cc, hq[0], hq[1]);
            if (cc + 8 < ncols) ld_global_v8(xh + cc + 8, hq[2], hq[3]);
            else hq[2] = hq[3] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              s1 += (v[4 * j] + v[4 * j + 1]) + (v[4 * j + 2] + v[4 * j + 3]);
              s2 += v[4 * j] * hq[j].x + v[4 * j + 1] * hq[j].y + v[4 * j + 2] * hq[j].z + v[4 * j + 3] * hq[j].w;
            }
          }
          const float inv_n = 1.f / (float)g.N;
          s1 *= inv_n; s2 *= inv_n;
          const float rs = __ldg(g.ln_rstd + m);
          for (int cc = part * 16; cc < ncols; cc += 16 * kParts) {
            float v[16];
            float4 hq[4];
            tmem_ld16(taddr + cc, v);
            ld_global_v8(xh + cc, hq[0], hq[1]);
            if (cc + 8 < ncols) ld_global_v8(xh + cc + 8, hq[2], hq[3]);
            else hq[2] = hq[3] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 o;
              o.x = rs * (v[4 * j] - s1 - hq[j].x * s2); o.y = rs * (v[4 * j + 1] - s1 - hq[j].y * s2);
              o.z = rs * (v[4 * j + 2] - s1 - hq[j].z * s2); o.w = rs * (v[4 * j + 3] - s1 - hq[j].w * s2);
              sts_v4(srow + (((uint32_t)j ^ sx) << 4), o);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) { tma_store_2d(&map_out, sbuf, n_base + cc, row0); bulk_commit(); }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
          if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
          continue;
        }
        if (prep && p.l2_prefetch == 2) {   // experiment: every lane pulls its pieces of the NEXT item's operand into L2
          WorkItem wn;
          if (get_work(p, it + 1, wn)) {
            const float *nx = pre_src + ((int64_t)wn.m_blk * BM + q * 32 + lane) * (int64_t)g.N + wn.n0;
            const int nc = (g.N - wn.n0 < wn.wn) ? g.N - wn.n0 : wn.wn;
            if ((int64_t)wn.m_blk * BM + q * 32 + lane < g.M)
              for (int c = part * 16; c < nc; c += 16 * kParts) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + c));
          }
        }
        int c0 = part * 16;
        float4 pre[4];
        uint32_t vr[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) pre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 < ncols) load_pre(c0, pre);
        mbar_wait(&tfull_bar[as], aphase);
        if (warp == 2 && lane == 0) TC_TRACE(p, 3);
        tc_fence_after();
        if (c0 < ncols && !(TC_DBG(p) & 8)) tmem_ld16_issue(taddr + c0, vr);
        for (; c0 < ncols; c0 += 16 * kParts) {
          const int cn = c0 + 16 * kParts;
          float v[16];
          float4 cur[4];
          if (!(TC_DBG(p) & 8)) tmem_ld_wait16(vr);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(vr[j]);
#pragma unroll
          for (int j = 0; j < 4; ++j) cur[j] = pre[j];
          if (cn < ncols) {
            if (!(TC_DBG(p) & 8)) tmem_ld16_issue(taddr + cn, vr);
#pragma unroll
            for (int j = 0; j < 4; ++j) pre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            load_pre(cn, pre);
          }
          float s1[16], s2[16];
          float4 o[4], o2[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int j = q4 * 4;
            const float4 acc = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            const float4 pv = cur[q4];
            const float4 bv = *reinterpret_cast<const float4 *>(vec_bias + c0 + j);
            float4 r, r2 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE == EPI_STORE) {
              r = make_float4(acc.x + bv.x + pv.x, acc.y + bv.y + pv.y, acc.z + bv.z + pv.z, acc.w + bv.w + pv.w);
            } else if (MODE == EPI_GELU_SQ) {
              r = make_float4(acc.x + bv.x, acc.y + bv.y, acc.z + bv.z, acc.w + bv.w);
              r2 = (TC_DBG(p) & 2) ? r : gelu4_f(r);
              s1[j] = r2.x * r2.x; s1[j + 1] = r2.y * r2.y; s1[j + 2] = r2.z * r2.z; s1[j + 3] = r2.w * r2.w;
            } else if (MODE == EPI_DG) {
              r = acc;
              s1[j] = acc.x * pv.x; s1[j + 1] = acc.y * pv.y; s1[j + 2] = acc.z * pv.z; s1[j + 3] = acc.w * pv.w;
              s2[j] = acc.x; s2[j + 1] = acc.y; s2[j + 2] = acc.z; s2[j + 3] = acc.w;
            } else {  // EPI_DH_GELU
              const float4 kgv = *reinterpret_cast<const float4 *>(vec_kg + c0 + j);
              const float4 asv = *reinterpret_cast<const float4 *>(vec_as + c0 + j);
              float4 hh, dd;
              gelu_both4_f(pv, hh, dd);
              r.x = fmaf(kgv.x, hh.x, acc.x * asv.x) * dd.x;
              r.y = fmaf(kgv.y, hh.y, acc.y * asv.y) * dd.y;
              r.z = fmaf(kgv.z, hh.z, acc.z * asv.z) * dd.z;
              r.w = fmaf(kgv.w, hh.w, acc.w * asv.w) * dd.w;
              s2[j] = r.x; s2[j + 1] = r.y; s2[j + 2] = r.z; s2[j + 3] = r.w;
            }
            o[q4] = r; o2[q4] = r2;
          }
          if (!(TC_DBG(p) & 4)) {
            if (lane == 0) bulk_wait_read0();   // the previous chunk's TMA store has finished reading the staging buffer
            __syncwarp();
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              sts_v4(srow + (((uint32_t)q4 ^ sx) << 4), o[q4]);
              if (has_out2) sts_v4(srow + kStageOutBytes + (((uint32_t)q4 ^ sx) << 4), o2[q4]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_out, sbuf, n_base + c0, row0);
              if (has_out2) tma_store_2d(&map_out2, sbuf + kStageOutBytes, n_base + c0, row0);
              bulk_commit();
            }
          }
          if (TC_DBG(p) & 1) {
            if (s1[0] + s1[5] + s1[10] + s1[15] + s2[3] == 123.456f) atomicAdd(&statacc1[n_base + c0 + lane], s1[7]);
            continue;
          }
          if (MODE == EPI_GELU_SQ || MODE == EPI_DG) {
            if (single_group) {
              const float t = warp_colsum16(s1, lane);
              if (lane < 16) atomicAdd(&statacc1[n_base + c0 + lane], t);
            } else {
              // per-sample statistics (decoder): the warp's 32 rows touch at most two groups (group_rows >= 32); each
              // group's column sums go straight to global memory
              const bool colv = lane < 16 && n_base + c0 + lane < g.N;
              if (gw_hi == gw_lo) {
                const float t = warp_colsum16(s1, lane);
                if (colv) atomicAdd(&g.colsum[(g_first + gw_lo) * g.N + n_base + c0 + lane], t);
              } else {
                float lo[16], hi[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { lo[j] = (my_g == gw_lo) ? s1[j] : 0.f; hi[j] = (my_g == gw_lo) ? 0.f : s1[j]; }
                const float tl = warp_colsum16(lo, lane), th = warp_colsum16(hi, lane);
                if (colv) {
                  atomicAdd(&g.colsum[(g_first + gw_lo) * g.N + n_base + c0 + lane], tl);
                  atomicAdd(&g.colsum[(g_first + gw_hi) * g.N + n_base + c0 + lane], th);
                }
              }
            }
          }
          if (MODE == EPI_DG || MODE == EPI_DH_GELU) {
            const float t = warp_colsum16(s2, lane);
            if (lane < 16) atomicAdd(&statacc2[n_base + c0 + lane], t);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
        if (warp == 2 && lane == 0) TC_TRACE(p, 4);
        if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
        continue;
      }
      float4 pre[4];
      prefetch(part * 16, pre);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();

      for (int c0 = part * 16; c0 < bn; c0 += 16 * kParts) {
        float4 nxt[4];
        prefetch(c0 + 16 * kParts, nxt);
        float v[16];
        tmem_ld16(taddr + c0, v);
        float s1[16], s2[16];  // statistics contributions (dead code for EPI_STORE)
        float4 ov[4], ov2[4];
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          const int n = n_base + c0 + j4;
          const bool col_ok = n < g.N;   // N % 4 == 0: a float4 is entirely in or out
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f), o2 = o;
          float4 st1 = o, st2 = o;
          if (col_ok) {
            const float4 acc = make_float4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
            const float4 pv = pre[j4 >> 2];
            const float4 bv = *reinterpret_cast<const float4 *>(vec_bias + c0 + j4);
            if (MODE == EPI_STORE) {
              o = make_float4(acc.x + bv.x + pv.x, acc.y + bv.y + pv.y, acc.z + bv.z + pv.z, acc.w + bv.w + pv.w);
            } else if (MODE == EPI_GELU_SQ) {
              o = make_float4(acc.x + bv.x, acc.y + bv.y, acc.z + bv.z, acc.w + bv.w);
              o2 = make_float4(gelu_f(o.x), gelu_f(o.y), gelu_f(o.z), gelu_f(o.w));
              if (row_ok) st1 = make_float4(o2.x * o2.x, o2.y * o2.y, o2.z * o2.z, o2.w * o2.w);
            } else if (MODE == EPI_DG) {
              o = acc;
              if (row_ok) {
                st1 = make_float4(acc.x * pv.x, acc.y * pv.y, acc.z * pv.z, acc.w * pv.w);
                st2 = acc;
              }
            } else {  // EPI_DH_GELU: h = a * Phi(a) is recomputed from a (it shares the erf with gelu')
              if (row_ok) {
                const float4 kgv = *reinterpret_cast<const float4 *>(vec_kg + c0 + j4);
                const float4 asv = *reinterpret_cast<const float4 *>(vec_as + c0 + j4);
                float hx, hy, hz, hw, dx_, dy_, dz_, dw_;
                gelu_both_f(pv.x, hx, dx_); gelu_both_f(pv.y, hy, dy_); gelu_both_f(pv.z, hz, dz_); gelu_both_f(pv.w, hw, dw_);
                o.x = (acc.x * asv.x + kgv.x * hx) * dx_;
                o.y = (acc.y * asv.y + kgv.y * hy) * dy_;
                o.z = (acc.z * asv.z + kgv.z * hz) * dz_;
                o.w = (acc.w * asv.w + kgv.w * hw) * dw_;
                st2 = o;
              }
            }
          }
          ov[j4 >> 2] = o; ov2[j4 >> 2] = o2;
          if (MODE != EPI_STORE) {
            s1[j4] = st1.x; s1[j4 + 1] = st1.y; s1[j4 + 2] = st1.z; s1[j4 + 3] = st1.w;
            s2[j4] = st2.x; s2[j4 + 1] = st2.y; s2[j4 + 2] = st2.z; s2[j4 + 3] = st2.w;
          }
        }
        if (row_ok) {
          // rows are written in 32-byte (8-float) pieces when N allows: full sectors, half as many requests
          const int n0 = n_base + c0;
          if (p.vec8) {
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              if (n0 + 8 * h8 < g.N) {
                st_global_v8(g.out + m * g.N + n0 + 8 * h8, ov[2 * h8], ov[2 * h8 + 1]);
                if (has_out2) st_global_v8(g.out2 + m * g.N + n0 + 8 * h8, ov2[2 * h8], ov2[2 * h8 + 1]);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (n0 + 4 * j < g.N) {
                *reinterpret_cast<float4 *>(g.out + m * g.N + n0 + 4 * j) = ov[j];
                if (has_out2) *reinterpret_cast<float4 *>(g.out2 + m * g.N + n0 + 4 * j) = ov2[j];
              }
            }
          }
        }
        if (MODE != EPI_STORE) {
          // column sums over the warp's 32 rows (per statistics group), then one shared-memory atomic per column
          const int col = c0 + (lane & 15);
          if (MODE == EPI_GELU_SQ || MODE == EPI_DG) {
            if (gw_hi == gw_lo) {
              const float t = warp_colsum16(s1, lane);
              if (lane < 16) atomicAdd(&colacc[gw_lo * bn + col], t);
            } else {
              float lo[16], hi[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) { lo[j] = (my_g == gw_lo) ? s1[j] : 0.f; hi[j] = (my_g == gw_lo) ? 0.f : s1[j]; }
              const float tl = warp_colsum16(lo, lane), th = warp_colsum16(hi, lane);
              if (lane < 16) { atomicAdd(&colacc[gw_lo * bn + col], tl); atomicAdd(&colacc[gw_hi * bn + col], th); }
            }
          }
          if (MODE == EPI_DG || MODE == EPI_DH_GELU) {
            const float t = warp_colsum16(s2, lane);
            if (lane < 16) atomicAdd(&colacc2[col], t);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) pre[j] = nxt[j];
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }

      if (MODE != EPI_STORE) {
        // flush this tile's column statistics (epilogue warps only: named barrier 1)
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        const int64_t m_last = ((int64_t)(m_blk + 1) * BM < g.M ? (int64_t)(m_blk + 1) * BM : g.M) - 1;
        const int ng = single_group ? 1 : (int)(m_last / g.group_rows - g_first) + 1;
        if ((MODE == EPI_GELU_SQ || MODE == EPI_DG) && g.colsum) {
          for (int i = et; i < ng * bn; i += kEpiThreads) {
            const int gg = i / bn, cidx = i - gg * bn;
            if (n_base + cidx < g.N) atomicAdd(&g.colsum[(g_first + gg) * g.N + n_base + cidx], colacc[i]);
            colacc[i] = 0.f;
          }
        }
        if ((MODE == EPI_DG || MODE == EPI_DH_GELU) && g.colsum2) {
          for (int i = et; i < bn; i += kEpiThreads) {
            if (n_base + i < g.N) atomicAdd(&g.colsum2[n_base + i], colacc2[i]);
            colacc2[i] = 0.f;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
    }
    if (lane == 0) bulk_wait0();   // this warp's TMA stores have landed
    if (MODE != EPI_STORE) {   // fast-path statistics: one global atomic per column per CTA
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      for (int i = et; i < g.N; i += kEpiThreads) {
        if ((MODE == EPI_GELU_SQ || MODE == EPI_DG) && g.colsum) { const float t = statacc1[i]; if (t != 0.f) atomicAdd(&g.colsum[i], t); }
        if ((MODE == EPI_DG || MODE == EPI_DH_GELU) && g.colsum2) { const float t = statacc2[i]; if (t != 0.f) atomicAdd(&g.colsum2[i], t); }
      }
    }
  } else if (SPLIT) {
    // ================================================================ A splitter (4 warps): hi in place, lo next to it
    const int stid = threadIdx.x - kThreadsNoSplit;   // 0..kSplitThreads-1
    int stage = 0;
    uint32_t phase = 0;
    const float *a_sc = a_scale_s + 32;
    if (AGELU) {
      // per-column scale of the A operand in shared memory: the GRN scale derived here from the finished statistic
      // (every CTA redundantly: K <= a few thousand floats out of L2 while the first TMA loads are in flight), or given
      constexpr int kST = kSplitThreads > 0 ? kSplitThreads : 32;
      float *sc_w = a_scale_s + 32;
      if (g.grn_gsq) {
        float part = 0.f;
        for (int k = stid; k < g.K; k += kST) part += sqrtf(__ldg(g.grn_gsq + k));
        part = warp_sum(part);
        if (lane == 0) a_scale_s[stid >> 5] = part;
        asm volatile("bar.sync 2, %0;" ::"n"(kST) : "memory");
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < kST / 32; ++i) tot += a_scale_s[i];
        const float den = tot / (float)g.K + g.grn_eps;
        for (int k = stid; k < g.K; k += kST) {
          const float nx = sqrtf(__ldg(g.grn_gsq + k)) / den;
          const float sc = fmaf(__ldg(g.grn_gamma + k), nx, 1.f);
          sc_w[k] = sc;
          if (blockIdx.x == 0) { g.grn_nx[k] = nx; g.grn_scale[k] = sc; }
        }
        if (blockIdx.x == 0 && stid == 0) g.grn_denom[0] = den;
      } else {
        for (int k = stid; k < g.K; k += kST) sc_w[k] = g.a_scale ? __ldg(g.a_scale + k) : 1.f;
      }
      asm volatile("bar.sync 2, %0;" ::"n"(kST) : "memory");
    }
    WorkItem w;
    for (int it = 0; get_work(p, it, w); ++it) {
      for (int kb = 0; kb < p.num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        float4 *A = reinterpret_cast<float4 *>(smem + (size_t)stage * stage_bytes);
        float4 *Alo = A + a_bytes / 16;
        if (BF16) {
          // raw: one [128 rows x 128 B] fp32 box, physical 16-byte chunk pc of row r holds floats 4c .. 4c+3 with
          // c = pc ^ (r & 7).  Result in place: bf16 (hi, hi) of chunk c -> logical chunk c >> 1, 8-byte half c & 1;
          // (lo, lo) -> logical chunk 4 + (c >> 1).  A row belongs to kTPR consecutive threads of one warp: all of them read
          // their chunks, __syncwarp, then write (rows are independent, so there is no block-wide barrier).
          constexpr int kST = kSplitThreads > 0 ? kSplitThreads : 32;
          constexpr int kTPR = kST >= 128 ? kST / 128 : 1;          // threads per row
          constexpr int kPasses = kST >= 128 ? 1 : 128 / kST;       // rows per thread
          constexpr int kCh = 8 / kTPR;                             // chunks per thread and row
          uint8_t *base = smem + (size_t)stage * stage_bytes;
#pragma unroll 1
          for (int ps = 0; ps < ((TC_DBG(p) & 64) ? 0 : kPasses); ++ps) {
            const int r = kST >= 128 ? stid / kTPR : ps * kST + stid;
            const int sub = kST >= 128 ? stid % kTPR : 0;
            uint8_t *row = base + r * 128;
            float4 x[kCh];
#pragma unroll
            for (int j = 0; j < kCh; ++j) {
              const int c = j * kTPR + sub;                         // logical chunk: k = kb * 32 + 4 c .. + 3
              x[j] = *reinterpret_cast<const float4 *>(row + ((c ^ (r & 7)) << 4));
              if (AGELU) {   // A = saved pre-activation: h = gelu(a), times the GRN scale of its column
                const int k = kb * BK + c * 4;
                const float4 sc = k < g.K ? *reinterpret_cast<const float4 *>(a_sc + k) : make_float4(1.f, 1.f, 1.f, 1.f);
                const float4 hv = gelu4_f(x[j]);
                x[j] = make_float4(hv.x * sc.x, hv.y * sc.y, hv.z * sc.z, hv.w * sc.w);
              }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kCh; ++j) {
              const int c = j * kTPR + sub;
              uint32_t h0, l0, h1, l1;
              split_bf16x2(x[j].x, x[j].y, h0, l0);
              split_bf16x2(x[j].z, x[j].w, h1, l1);
              const uint32_t off = (uint32_t)((((c >> 1) ^ (r & 7)) << 4) + (c & 1) * 8);
              *reinterpret_cast<uint2 *>(row + off) = make_uint2(h0, h1);
              *reinterpret_cast<uint2 *>(row + (off ^ 64u)) = make_uint2(l0, l1);   // logical chunk + 4 = physical chunk ^ 4
            }
          }
        } else
        if (!(TC_DBG(p) & 64))
#pragma unroll
        for (int i = 0; i < (int)(a_bytes / 16) / (kSplitThreads > 0 ? kSplitThreads : 1); ++i) {
          float4 x = A[i * kSplitThreads + stid];
          if (AGELU) {
            const int idx = i * kSplitThreads + stid, r = idx >> 3, k = kb * BK + (((idx & 7) ^ (r & 7)) << 2);
            const float4 sc = k < g.K ? *reinterpret_cast<const float4 *>(a_sc + k) : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 hv = gelu4_f(x);
            x = make_float4(hv.x * sc.x, hv.y * sc.y, hv.z * sc.z, hv.w * sc.w);
          }
          float4 hi, lo;
          hi.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); lo.x = x.x - hi.x;
          hi.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); lo.y = x.y - hi.y;
          hi.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); lo.z = x.z - hi.z;
          hi.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); lo.w = x.w - hi.w;
          A[i * kSplitThreads + stid] = hi;
          Alo[i * kSplitThreads + stid] = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[stage]);
        if (++stage == nstage) { stage = 0; phase ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// [rows, cols] fp32 row-major, box = [box_rows, 32] with 128-byte swizzle, zero fill outside
inline bool make_map(CUtensorMap *map, const float *ptr, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [rows, cols] fp32 row-major, box = [32 rows, 16 floats] with 64-byte swizzle: the epilogue's TMA-store chunk
inline bool make_map_out(CUtensorMap *map, const float *ptr, int64_t rows, int64_t cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {16, 32};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [rows, cols] bf16 row-major, box = [box_rows, 64] with 128-byte swizzle (3xBF16 weight operand)
inline bool make_map_bf16(CUtensorMap *map, const void *ptr, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct MapCache {
  std::mutex mu;
  std::map<std::tuple<const void *, int64_t, int64_t, int>, CUtensorMap> maps;
  // box_rows > 0: fp32 operand map (K-major, 128-byte swizzle, box [box_rows x 32]); box_rows == -1: output map;
  // box_rows <= -16: bf16 operand map with box [-box_rows x 64]
  bool get(CUtensorMap *out, const float *ptr, int64_t rows, int64_t cols, int box_rows) {
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple((const void *)ptr, rows, cols, box_rows);
    auto it = maps.find(key);
    if (it == maps.end()) {
      CUtensorMap m;
      if (box_rows <= -16 ? !make_map_bf16(&m, ptr, rows, cols, -box_rows)
                          : (box_rows == -1 ? !make_map_out(&m, ptr, rows, cols) : !make_map(&m, ptr, rows, cols, box_rows)))
        return false;
      it = maps.emplace(key, m).first;
    }
    *out = it->second;
    return true;
  }
};
inline MapCache &map_cache() { static MapCache c; return c; }

#ifdef MPMAE_TC_KNOBS
inline unsigned long long *&tc_trace_buffer() { static unsigned long long *b = nullptr; return b; }
#endif
inline int pick_bn(int N) {
  if (N <= 256) return ((N + 15) / 16) * 16;
  for (int bn = 256; bn >= 128; bn -= 16)
    if (N % bn == 0) return bn;
  return 256;
}


// ------------------------------------------------------------------------------------------------ weight gradients
// dW[n, k] += rs[n] * sum_r X[r, n] * Y[r, k]: the contraction runs over ROWS, so both operands are "MN-major" for
// the tensor core (the feature dimension is contiguous in memory).  TMA boxes of [32 rows x 32 floats] (128-byte
// swizzle) land as the canonical MN-major atoms ([8 rows][128 B], 1024 B each); one tcgen05.mma (K = 8 rows)
// consumes one atom per 32-feature block, blocks LBO = 4096 B apart.  Work item = (m tile, n tile, row split);
// partial tiles are merged with fp32 atomics (the reference merges with atomics too,
// MinkowskiEngine/src/convolution_kernel.cu:198-290).  Single-pass TF32.
struct TnParams {
  const float *rs;   // [Nw] scale of dW rows or null
  float *dW;         // [Nw, Kw]
  int64_t R;
  int Nw, Kw;
  int swap;          // 0: MMA M side = X features (n), N side = Y features (k);  1: M side = Y (k), N side = X (n)
  int Msz, Nsz;      // feature counts on the M / N side
  int bn;            // N tile, multiple of 32, <= 256
  int num_m, num_n, splits, chunks_per_split;   // chunk = 32 rows
  int stages;
  int vec4;          // dW 16-byte aligned and Kw % 4 == 0: vector atomics when the M side is the dW row (swap == 0)
  int gelu_m, gelu_n;   // SPLIT only: the M- / N-side operand is consumed as gelu(operand) (dW2f = dy^T . gelu(a))
  uint32_t tmem_cols;
  unsigned long long *trace;   // knob builds only (TC_TRACE)
  int m3d, n3d;      // the operand's feature count is a multiple of 32: ONE 3-D TMA box per chunk brings all its 32-feature blocks
};

// MN-major TF32 operands have exactly one legal shared-memory layout: 128-byte swizzle with 32-byte atomicity
// (UMMA LayoutType SWIZZLE_128B_BASE32B = 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of [4 rows][128 B].
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | ((4096u >> 4) << 16);   // LBO: next 32-feature block
  const uint32_t hi = (512u >> 4) | (1u << 14) | (1u << 29);             // SBO: next 4-row atom ; v1 ; SWIZZLE_128B_BASE32B
  return (uint64_t)lo | ((uint64_t)hi << 32);
}

constexpr int kTnStages = 8;   // ring depth cap: a stage is only 32 rows (<= 48 KB of operands), and the single-pass kernel is
                               // bound by bytes in flight (ncu r2_s: 9 % warps active, tensor pipe 30 %, DRAM 28 %)
constexpr int kTnThreads = 192, kTnSplitThreads = 512;   // 16 splitter warps: both operand tiles are split (one also through GELU) every chunk

// SPLIT = 3xTF32: both operand tiles are split in shared memory (hi in place, remainder in a second buffer) by four
// splitter warps; D += Mhi.Nhi + Mlo.Nhi + Mhi.Nlo.  Used where the product feeds back into the data path (the GRN
// statistic gradient is derived from dW2f by the chain rule of the weight fold).
template <bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? kTnThreads + kTnSplitThreads : kTnThreads, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap map_m, const __grid_constant__ CUtensorMap map_n,
                  const __grid_constant__ CUtensorMap map_m3, const __grid_constant__ CUtensorMap map_n3, const TnParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a shared pointer
  const int bn = p.bn, nstage = p.stages;
  constexpr uint32_t m_bytes = BM * 32 * 4;                 // 4 blocks x [32 rows][128 B]
  const uint32_t n_bytes = (uint32_t)bn * 32 * 4;           // bn/32 blocks
  const uint32_t m_span = SPLIT ? 2 * m_bytes : m_bytes;
  const uint32_t stage_bytes = m_span + (SPLIT ? 2 : 1) * n_bytes;   // [M | Mlo | N | Nlo]
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)nstage * stage_bytes);
  uint64_t *empty_bar = full_bar + kTnStages;
  uint64_t *split_bar = empty_bar + kTnStages;
  uint64_t *tfull_bar = split_bar + kTnStages;
  uint64_t *tempty_bar = tfull_bar + ACC_STAGES;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tempty_bar + ACC_STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_m)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_n)) : "memory");
    for (int s = 0; s < kTnStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&split_bar[s], kTnSplitThreads / 32); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr, p.tmem_cols);
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int total = p.num_m * p.num_n * p.splits;
  const int total_chunks = (int)((p.R + 31) / 32);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      TC_TRACE_DECL;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int sp = item / (p.num_m * p.num_n), t = item - sp * (p.num_m * p.num_n);
        const int m_blk = t / p.num_n, n_blk = t - m_blk * p.num_n;
        const int c_begin = sp * p.chunks_per_split;
        const int c_end = min(c_begin + p.chunks_per_split, total_chunks);
        for (int ch = c_begin; ch < c_end; ++ch) {
          mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
          TC_TRACE(p, 0);
          uint8_t *sm = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage], m_bytes + n_bytes);
          // (the single producer thread was the bottleneck of the single-pass kernel: 9 box loads per 32-row chunk took
          // 0.41 us to issue, profiles/r2_aj_trace_tn.txt; a 3-D box [32 features][32 rows][blocks] lands in the same layout)
          if (p.m3d) {
            tma_load_3d(sm, &map_m3, &full_bar[stage], 0, ch * 32, m_blk * (BM / 32));
          } else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) tma_load_2d(sm + j * 4096, &map_m, &full_bar[stage], m_blk * BM + j * 32, ch * 32);
          }
          if (p.n3d) {
            tma_load_3d(sm + m_span, &map_n3, &full_bar[stage], 0, ch * 32, n_blk * (bn / 32));
          } else {
            for (int j = 0; j < bn / 32; ++j)
              tma_load_2d(sm + m_span + j * 4096, &map_n, &full_bar[stage], n_blk * bn + j * 32, ch * 32);
          }
          if (++stage == nstage) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(bn) | (1u << 15) | (1u << 16);   // A and B MN-major
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      TC_TRACE_DECL;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int sp = item / (p.num_m * p.num_n);
        const int c_begin = sp * p.chunks_per_split;
        const int c_end = min(c_begin + p.chunks_per_split, total_chunks);
        mbar_wait_relaxed(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * bn);
        for (int ch = c_begin; ch < c_end; ++ch) {
          mbar_wait(&full_bar[stage], phase);
          TC_TRACE(p, 1);
          if (SPLIT) mbar_wait(&split_bar[stage], phase);
          TC_TRACE(p, 2);
          tc_fence_after();
          const uint32_t sm = smem_u32(smem + (size_t)stage * stage_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 8 rows per instruction = one 1024-byte atom per feature block
            umma_tf32(tmem_d, make_smem_desc_mn(sm + k * 1024), make_smem_desc_mn(sm + m_span + k * 1024), idesc,
                      (ch > c_begin || k > 0) ? 1u : 0u);
          if (SPLIT) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32(tmem_d, make_smem_desc_mn(sm + m_bytes + k * 1024), make_smem_desc_mn(sm + m_span + k * 1024), idesc, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32(tmem_d, make_smem_desc_mn(sm + k * 1024), make_smem_desc_mn(sm + m_span + n_bytes + k * 1024), idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
          if (ch == c_end - 1) umma_commit(&tfull_bar[as]);
          if (++stage == nstage) { stage = 0; phase ^= 1; }
        }
        if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    TC_TRACE_DECL;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const int t = item % (p.num_m * p.num_n);
      const int m_blk = t / p.num_n, n_blk = t - m_blk * p.num_n;
      const int mi = m_blk * BM + q * 32 + lane;          // index on the M side
      mbar_wait(&tfull_bar[as], aphase);
      if (warp == 2 && lane == 0) TC_TRACE(p, 3);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * bn);
      for (int c0 = 0; c0 < bn; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (mi < p.Msz) {
          const int ni0 = n_blk * bn + c0;
          if (!p.swap && p.vec4 && ni0 + 16 <= p.Nsz) {
            // a lane owns 16 consecutive k of one dW row: four 16-byte vector atomics instead of 16 scalar ones
            const float sc = p.rs ? __ldg(p.rs + mi) : 1.f;
            float4 *dst = reinterpret_cast<float4 *>(p.dW + (int64_t)mi * p.Kw + ni0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              atomicAdd(dst + j, make_float4(sc * v[4 * j], sc * v[4 * j + 1], sc * v[4 * j + 2], sc * v[4 * j + 3]));
          } else if (!p.swap || !p.vec4) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int ni = ni0 + j;
              if (ni < p.Nsz) {
                const int n = p.swap ? ni : mi, k = p.swap ? mi : ni;
                const float sc = p.rs ? __ldg(p.rs + n) : 1.f;
                atomicAdd(&p.dW[(int64_t)n * p.Kw + k], sc * v[j]);
              }
            }
          }
        }
        if (p.swap && p.vec4) {
          // swapped roles: a lane holds ONE k (its TMEM lane) and 16 consecutive n, but dW rows run along k.  4 x 4 transposes
          // inside lane quads (4 shuffles + 8 selects each) give every lane 4 consecutive k of one n: four 16-byte vector
          // atomics per chunk instead of sixteen 4-byte ones (the scalar epilogue took 11 - 18 us of a 30 us item at stages
          // 2 / 3, profiles/r2_aj_trace_tn.txt).  All lanes shuffle; only the atomics are guarded.
          const int t = lane & 3;
          const bool odd = (t & 1) != 0, hi = (t & 2) != 0;
          const int kq = mi & ~3;                       // Msz % 4 == 0: a quad is inside or outside as a whole
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            float x0 = v[4 * b], x1 = v[4 * b + 1], x2 = v[4 * b + 2], x3 = v[4 * b + 3];
            {
              const float s0 = odd ? x0 : x1, s1 = odd ? x2 : x3;
              const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
              if (odd) { x0 = r0; x2 = r1; } else { x1 = r0; x3 = r1; }
            }
            {
              const float s0 = hi ? x0 : x2, s1 = hi ? x1 : x3;
              const float r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
              if (hi) { x0 = r0; x1 = r1; } else { x2 = r0; x3 = r1; }
            }
            const int n = n_blk * bn + c0 + 4 * b + t;   // x0..x3 = (k = kq .. kq + 3, n)
            if (kq < p.Msz && n < p.Nsz) {
              const float sc = p.rs ? __ldg(p.rs + n) : 1.f;
              atomicAdd(reinterpret_cast<float4 *>(p.dW + (int64_t)n * p.Kw + kq), make_float4(sc * x0, sc * x1, sc * x2, sc * x3));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (warp == 2 && lane == 0) TC_TRACE(p, 4);
      if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
    }
  } else if (SPLIT) {
    const int stid = threadIdx.x - kTnThreads;   // 0..kTnSplitThreads-1
    int stage = 0;
    uint32_t phase = 0;
    auto split_region = [&](float4 *src, float4 *lo, int n4, bool act) {
      for (int i = stid; i < n4; i += kTnSplitThreads) {
        float4 x = src[i];
        if (act) x = gelu4_f(x);
        float4 hi, l;
        hi.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - hi.x;
        hi.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - hi.y;
        hi.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - hi.z;
        hi.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - hi.w;
        src[i] = hi;
        lo[i] = l;
      }
    };
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const int sp = item / (p.num_m * p.num_n);
      const int c_begin = sp * p.chunks_per_split;
      const int c_end = min(c_begin + p.chunks_per_split, total_chunks);
      for (int ch = c_begin; ch < c_end; ++ch) {
        mbar_wait(&full_bar[stage], phase);
        uint8_t *sm = smem + (size_t)stage * stage_bytes;
        split_region(reinterpret_cast<float4 *>(sm), reinterpret_cast<float4 *>(sm + m_bytes), m_bytes / 16, p.gelu_m != 0);
        split_region(reinterpret_cast<float4 *>(sm + m_span), reinterpret_cast<float4 *>(sm + m_span + n_bytes), n_bytes / 16,
                     p.gelu_n != 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[stage]);
        if (++stage == nstage) { stage = 0; phase ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// [rows, cols] fp32 row-major, box = [32 rows, 32 floats], 128-byte swizzle with 32-byte atoms
inline bool make_map_box32(CUtensorMap *map, const float *ptr, int64_t rows, int64_t cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the same [rows, cols] fp32 matrix (cols % 32 == 0) as a 3-D tensor (32 features, rows, cols / 32 blocks): one box
// [32][32 rows][nblocks] = nblocks MN-major operand blocks of 4096 bytes, block indices beyond cols / 32 zero-filled
inline bool make_map_box32_3d(CUtensorMap *map, const float *ptr, int64_t rows, int64_t cols, int nblocks) {
  EncodeTiledFn fn = encode_fn();
  if (!fn || cols % 32 != 0) return false;
  const cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
  const cuuint64_t strides[2] = {(cuuint64_t)cols * 4, 128};
  const cuuint32_t box[3] = {32, 32, (cuuint32_t)nblocks};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc

// the fused LayerNorm-backward epilogue needs every tile on the fast path with the whole row in one tile
inline bool tc_ln_bwd_ok(const GemmArgs &a) {
  return a.M % tc::BM == 0 && a.N <= 256 && a.N % 8 == 0 && a.ln_xhat && a.ln_rstd && !a.resid && !a.bias &&
         (((uintptr_t)a.ln_xhat | (uintptr_t)a.out) & 31) == 0;
}
inline bool tc_gemm_supported(int mode, const GemmArgs &a) {
  if (a.M < 1 || a.N < 8 || a.K < 8) return false;
  if (a.K % 8 != 0 || a.N % 4 != 0) return false;
  if (((uintptr_t)a.A | (uintptr_t)a.Bw | (uintptr_t)a.out) & 15) return false;
  if (mode != EPI_STORE && a.group_rows < 32) return false;
  if (a.M < 64) return false;  // a 128-row MMA tile would be mostly padding: tiny products stay on the SIMT path
  return tc::encode_fn() != nullptr;
}

template <int MODE, bool SPLIT, bool WIDE, bool BF16 = false, bool AGELU = false>
inline cudaError_t launch_gemm_rows_tc_impl(const GemmArgs &a, cudaStream_t st) {
  using namespace tc;
  TcParams p{};
  p.g = a;
  static const int env_bn = getenv("MPMAE_TC_BN") ? atoi(getenv("MPMAE_TC_BN")) : 0;   // environment knobs are read once per process
  static const bool env_dbg = getenv("MPMAE_TC_DBG") != nullptr;        // timing experiments re-read their mask per launch
  auto smem_for = [&](int bn, int stages) {
    // 3xBF16: one tile per operand holds the high parts and the remainders of 32 k; 3xTF32: separate hi / lo tiles
    const size_t a_stage = (size_t)((SPLIT && !BF16) ? 2 : 1) * BM * BK * 4, b_stage = (size_t)((SPLIT && !BF16) ? 2 : 1) * bn * BK * 4;
    const size_t ring = (size_t)stages * (a_stage + b_stage);
    const int num_n = cdiv(a.N, bn);
    return 1024 + ring + (size_t)epi_warps(MODE, WIDE, AGELU) * out_arrays(MODE) * kStageOutBytes + 320 +
           (size_t)(5 * bn + (MODE == EPI_STORE ? 2 : 5) * num_n * bn + (AGELU ? a.K + 32 : 0)) * 4;
  };
  // N tile: the widest multiple of 16 (<= 256) that divides N and fits the shared-memory budget with a 2-stage ring,
  // else a ragged tail; then as many ring stages as still fit
  int bn = 0;
  const int cap = a.N <= 256 ? ((a.N + 15) / 16) * 16 : 256;
  for (int c = cap; c >= 64 && !bn; c -= 16)
    if ((a.N % c == 0 || c >= a.N) && smem_for(c, 2) <= 226 * 1024) bn = c;
  for (int c = cap; c >= 16 && !bn; c -= 16)
    if (smem_for(c, 2) <= 226 * 1024) bn = c;
  if (env_bn > 0 && env_bn % 16 == 0 && smem_for(env_bn, 2) <= 226 * 1024) bn = env_bn;
  if (!bn) return cudaErrorInvalidConfiguration;
  p.bn = bn;
  p.stages = 2;
  static const int env_stages = getenv("MPMAE_TC_STAGES") ? atoi(getenv("MPMAE_TC_STAGES")) : STAGES;
  // the GELU-backward epilogue streams a second [M, N] operand through the LSU; a deep operand ring running ahead of it made
  // that kernel SLOWER (stage 0: 126 us with 3 stages, 141 with 5 -- profiles/r2_ad_sweep.txt)
  const int max_stages = MODE == EPI_DH_GELU ? 3 : STAGES;
  while (p.stages < max_stages && p.stages < env_stages && smem_for(bn, p.stages + 1) <= 226 * 1024) ++p.stages;
  p.num_m = (int)cdiv64(a.M, BM);
  p.num_n = cdiv(a.N, p.bn);
  p.num_k = cdiv(a.K, BK);
  p.vec8 = (a.N % 8 == 0) && (((uintptr_t)a.out | (uintptr_t)a.out2) & 31) == 0;
  p.vec8_in = (a.N % 8 == 0) && (((uintptr_t)a.resid | (uintptr_t)a.aux | (uintptr_t)a.aux2) & 31) == 0;
  uint32_t cols = 32;
  while (cols < (uint32_t)(ACC_STAGES * p.bn)) cols <<= 1;
  p.tmem_cols = cols;
  p.dbg = 0;
  if (env_dbg) { const char *e = getenv("MPMAE_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
  CUtensorMap ma, mb, mbl, mo, mo2;
  if (!map_cache().get(&ma, a.A, a.M, a.K, BM)) return cudaErrorInvalidValue;
  if (BF16) {   // Bw_lo holds the interleaved bf16 pair [N][ceil(K / 32)][hi 32 | lo 32] (bf16_pair_cols(K) columns per row)
    if (!map_cache().get(&mb, a.Bw_lo, a.N, bf16_pair_cols(a.K), -p.bn)) return cudaErrorInvalidValue;
    mbl = mb;
  } else {
    if (!map_cache().get(&mb, a.Bw, a.N, a.K, p.bn)) return cudaErrorInvalidValue;
    mbl = mb;
    if (SPLIT && !map_cache().get(&mbl, a.Bw_lo, a.N, a.K, p.bn)) return cudaErrorInvalidValue;
  }
  p.tma_out = map_cache().get(&mo, a.out, a.M, a.N, -1) ? 1 : 0;
  mo2 = mo;
  if (MODE == EPI_GELU_SQ && a.out2 && p.tma_out && !map_cache().get(&mo2, a.out2, a.M, a.N, -1)) p.tma_out = 0;
  if (!p.tma_out) mo = mo2 = ma;
  if (a.ln_rstd && (MODE != EPI_STORE || !p.tma_out || p.num_n != 1 || !tc_ln_bwd_ok(a))) return cudaErrorInvalidConfiguration;
  int grid = p.num_m * p.num_n;
  if (grid > 148) grid = 148;
  // tail splitting (see TcParams): only when every tile takes the predicate-free epilogue path and N tiles are whole
  const int total = p.num_m * p.num_n;
  p.tail_start = total; p.tail_items = 0; p.tail_S = 1; p.tail_ws = p.bn;
  auto div_magic = [](int d) { return (uint32_t)((1ull << 32) / (uint64_t)d + 1ull); };   // fast_div(): d >= 2 (d == 1 is branched)
  CUtensorMap mbt = mb, mbtl = mbl;
  {
    const bool single_group = (int64_t)a.group_rows >= a.M;
    const float *pre = MODE == EPI_STORE ? a.resid : MODE == EPI_DG ? a.aux : MODE == EPI_DH_GELU ? a.aux2 : nullptr;
    const bool all_fast = p.tma_out && a.M % BM == 0 && (single_group || MODE == EPI_GELU_SQ || MODE == EPI_DG) &&
                          (!pre || (a.N % 8 == 0 && p.vec8_in)) && !a.ln_rstd && a.N % p.bn == 0;
    const int full = (total / grid) * grid, rem = total - full;
    static const bool no_tail = getenv("MPMAE_TC_NO_TAIL") != nullptr;
    if (all_fast && !no_tail && total > grid && rem > 0 && rem * 2 <= grid) {
      for (int ws = 16; ws < p.bn; ws += 16) {
        if (p.bn % ws != 0 || rem * (p.bn / ws) > grid) continue;
        bool ok;
        if (BF16) {
          ok = map_cache().get(&mbt, a.Bw_lo, a.N, bf16_pair_cols(a.K), -ws);
          mbtl = mbt;
        } else {
          ok = map_cache().get(&mbt, a.Bw, a.N, a.K, ws);
          mbtl = mbt;
          if (ok && SPLIT) ok = map_cache().get(&mbtl, a.Bw_lo, a.N, a.K, ws);
        }
        if (ok) { p.tail_start = full; p.tail_S = p.bn / ws; p.tail_ws = ws; p.tail_items = rem * p.tail_S; }
        break;
      }
    }
  }
  // measured (profiles/r2_x_sweep.txt): no gain, the stage-0 da kernel 5 % slower -- opt-in only
  static const int env_pf = getenv("MPMAE_TC_L2PF") ? atoi(getenv("MPMAE_TC_L2PF")) : 0;   // 1: bulk by the producer, 2: per lane
  p.l2_prefetch = (env_pf > 0 && a.N % 4 == 0 && (((uintptr_t)a.resid | (uintptr_t)a.aux | (uintptr_t)a.aux2 | (uintptr_t)a.ln_xhat) & 15) == 0) ? env_pf : 0;
  p.trace = nullptr;
#ifdef MPMAE_TC_KNOBS
  if (getenv("MPMAE_TC_TRACE")) {
    static unsigned long long *buf = nullptr;
    if (!buf) cudaMalloc(&buf, 8 * 256 * sizeof(unsigned long long));
    cudaMemsetAsync(buf, 0, 8 * 256 * sizeof(unsigned long long), st);
    p.trace = buf;
    tc_trace_buffer() = buf;
  }
#endif
  static const int env_spin = getenv("MPMAE_TC_SPIN") ? atoi(getenv("MPMAE_TC_SPIN")) : -1;
  p.spin = env_spin >= 0 ? env_spin : 0;
  p.num_n_magic = div_magic(p.num_n > 1 ? p.num_n : 2);
  p.tail_S_magic = p.tail_S > 1 ? div_magic(p.tail_S) : 0u;
  const size_t smem = smem_for(bn, p.stages);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<MODE, SPLIT, WIDE, BF16, AGELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  pdl(gemm_tc_kernel<MODE, SPLIT, WIDE, BF16, AGELU>, grid, tc_threads(MODE, SPLIT, WIDE, AGELU), smem, st)(ma, mb, mbl, mo, mo2, mbt, mbtl, p);
  return cudaGetLastError();
}

template <int MODE>
inline cudaError_t launch_gemm_rows_tc(const GemmArgs &a, int backend, cudaStream_t st) {
  // GELU on the A operand is applied by the operand-splitter warps: split backends only
  if (a.a_gelu) {
    if constexpr (MODE == EPI_STORE) {
      if (backend == 3 && a.b16 && a.Bw_lo) return launch_gemm_rows_tc_impl<EPI_STORE, true, true, true, true>(a, st);
      if (backend == 1 && a.Bw_lo) return launch_gemm_rows_tc_impl<EPI_STORE, true, true, false, true>(a, st);
    }
    return cudaErrorInvalidConfiguration;
  }
  const bool wide = MODE == EPI_STORE || a.K >= 256;   // WIDE only changes the roles of the statistics / GELU epilogues
  if (backend == 3 && a.b16 && a.Bw_lo)
    return wide ? launch_gemm_rows_tc_impl<MODE, true, true, true>(a, st)
                : launch_gemm_rows_tc_impl<MODE, true, MODE == EPI_STORE, true>(a, st);
  const bool split = backend == 1 && a.Bw_lo;
  if (split) return wide ? launch_gemm_rows_tc_impl<MODE, true, true>(a, st)
                         : launch_gemm_rows_tc_impl<MODE, true, MODE == EPI_STORE>(a, st);
  return wide ? launch_gemm_rows_tc_impl<MODE, false, true>(a, st) : launch_gemm_rows_tc_impl<MODE, false, MODE == EPI_STORE>(a, st);
}


inline bool tc_wgrad_supported(const WgradArgs &a) {
  if (a.R < 256 || a.N < 8 || a.K < 8 || a.N % 4 != 0 || a.K % 4 != 0) return false;
  if (((uintptr_t)a.X | (uintptr_t)a.Y) & 15) return false;
  return tc::encode_fn() != nullptr;
}

// dW += rs * X^T . Y on the tensor cores (db is NOT produced here)
template <bool SPLIT>
inline cudaError_t launch_gemm_wgrad_tc_impl(const WgradArgs &a, cudaStream_t st) {
  using namespace tc;
  TnParams p{};
  p.rs = a.rs; p.dW = a.dW; p.R = a.R; p.Nw = a.N; p.Kw = a.K;
  auto cost = [](int msz, int nsz) {  // operand re-reads: M-side operand once per n tile, N-side once per m tile
    const int nm = cdiv(msz, BM), nn = cdiv(nsz, 256);
    return (double)msz * nn + (double)nsz * nm;
  };
  p.swap = cost(a.K, a.N) < cost(a.N, a.K) ? 1 : 0;
  static const int env_swap = getenv("MPMAE_TN_SWAP") ? atoi(getenv("MPMAE_TN_SWAP")) : -1;   // experiment: force 0 / 1
  if (env_swap >= 0) p.swap = env_swap;
  p.Msz = p.swap ? a.K : a.N;
  p.Nsz = p.swap ? a.N : a.K;
  if (a.y_gelu && !SPLIT) return cudaErrorInvalidConfiguration;   // the activation is applied by the splitter warps
  p.gelu_m = (a.y_gelu && p.swap) ? 1 : 0;     // Y sits on the M side when swapped
  p.gelu_n = (a.y_gelu && !p.swap) ? 1 : 0;
  int bn;
  if (p.Nsz <= 256) bn = ((p.Nsz + 31) / 32) * 32;
  else {
    bn = 256;
    for (int c = 256; c >= 128; c -= 32)
      if (p.Nsz % c == 0) { bn = c; break; }
  }
  p.bn = bn;
  p.num_m = cdiv(p.Msz, BM);
  p.num_n = cdiv(p.Nsz, bn);
  const int total_chunks = (int)cdiv64(a.R, 32);
  const int tiles = p.num_m * p.num_n;
  // one item per CTA when the output tile is big (the epilogue's atomics dominate), two when it is small (the second
  // item's MMAs hide the first one's epilogue)
  const int items_target = ((int64_t)p.Msz * p.Nsz >= 32768) ? 148 : 296;
  // floor, not ceil: 160 items on 148 persistent CTAs take two item times (stage 3: 20 tiles x 8 splits; decoder:
  // 32 tiles x 5), 140 / 128 items take one
  int splits = tiles >= items_target ? 1 : items_target / tiles;
  p.vec4 = (((uintptr_t)a.dW & 15) == 0 && a.K % 4 == 0) ? 1 : 0;
  static const int env_items = getenv("MPMAE_TN_ITEMS") ? atoi(getenv("MPMAE_TN_ITEMS")) : 0;
  if (env_items > 0) splits = cdiv(env_items, tiles);
  const int max_splits = cdiv(total_chunks, 8);   // >= 256 rows per item
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.chunks_per_split = cdiv(total_chunks, splits);
  p.splits = cdiv(total_chunks, p.chunks_per_split);
  uint32_t cols = 32;
  while (cols < (uint32_t)(ACC_STAGES * bn)) cols <<= 1;
  p.tmem_cols = cols;
  const float *msrc = p.swap ? a.Y : a.X, *nsrc = p.swap ? a.X : a.Y;
  static std::map<std::tuple<const void *, int64_t, int64_t>, CUtensorMap> cache;
  static std::mutex mu;
  auto get = [&](CUtensorMap *out, const float *ptr, int64_t rows, int64_t cols_) {
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple((const void *)ptr, rows, cols_);
    auto it = cache.find(key);
    if (it == cache.end()) {
      CUtensorMap m;
      if (!make_map_box32(&m, ptr, rows, cols_)) return false;
      it = cache.emplace(key, m).first;
    }
    *out = it->second;
    return true;
  };
  CUtensorMap mm, mn;
  if (!get(&mm, msrc, a.R, p.Msz) || !get(&mn, nsrc, a.R, p.Nsz)) return cudaErrorInvalidValue;
  static std::map<std::tuple<const void *, int64_t, int64_t, int>, CUtensorMap> cache3;
  auto get3 = [&](CUtensorMap *out, const float *ptr, int64_t rows, int64_t cols_, int nblocks) {
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple((const void *)ptr, rows, cols_, nblocks);
    auto it = cache3.find(key);
    if (it == cache3.end()) {
      CUtensorMap m;
      if (!make_map_box32_3d(&m, ptr, rows, cols_, nblocks)) return false;
      it = cache3.emplace(key, m).first;
    }
    *out = it->second;
    return true;
  };
  static const bool no3d = getenv("MPMAE_TN_NO3D") != nullptr;
  CUtensorMap mm3 = mm, mn3 = mn;
  p.m3d = (!no3d && p.Msz % 32 == 0 && get3(&mm3, msrc, a.R, p.Msz, BM / 32)) ? 1 : 0;
  p.n3d = (!no3d && p.Nsz % 32 == 0 && get3(&mn3, nsrc, a.R, p.Nsz, bn / 32)) ? 1 : 0;
  p.trace = nullptr;
#ifdef MPMAE_TC_KNOBS
  if (getenv("MPMAE_TC_TRACE")) {
    static unsigned long long *buf = nullptr;
    if (!buf) cudaMalloc(&buf, 8 * 256 * sizeof(unsigned long long));
    cudaMemsetAsync(buf, 0, 8 * 256 * sizeof(unsigned long long), st);
    p.trace = buf;
    tc_trace_buffer() = buf;
  }
#endif
  const size_t stage_bytes = (size_t)(SPLIT ? 2 : 1) * (BM * 32 * 4 + (size_t)bn * 32 * 4);
  int stages = (int)((224 * 1024 - 2048) / stage_bytes);
  static const int env_tn_stages = getenv("MPMAE_TN_STAGES") ? atoi(getenv("MPMAE_TN_STAGES")) : 0;
  const int cap = env_tn_stages >= 2 && env_tn_stages <= kTnStages ? env_tn_stages : kTnStages;
  if (stages > cap) stages = cap;
  if (stages < 2) return cudaErrorInvalidConfiguration;
  p.stages = stages;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_tc_kernel<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int grid = tiles * p.splits;
  if (grid > 148) grid = 148;
  pdl(gemm_tn_tc_kernel<SPLIT>, grid, SPLIT ? kTnThreads + kTnSplitThreads : kTnThreads, smem, st)(mm, mn, mm3, mn3, p);
  return cudaGetLastError();
}

// exact = 3xTF32 (fp32-faithful); otherwise single-pass TF32
inline cudaError_t launch_gemm_wgrad_tc(const WgradArgs &a, bool exact, cudaStream_t st) {
  return exact ? launch_gemm_wgrad_tc_impl<true>(a, st) : launch_gemm_wgrad_tc_impl<false>(a, st);
}

}  // namespace mpmae
