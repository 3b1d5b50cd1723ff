// tcgen05 GEMM for the token-wise linear layers (QKV / out-proj / fc1 / fc2: 75 % of the denoiser
// FLOPs):   out = epilogue( A[M,K] · W[N,K]^T + bias )      fp32 accumulate in TMEM.
// Operands are bf16 (kind::f16, default) or TF32-in-fp32 (kind::tf32); both use the same code path.
//
// Blackwell-native structure (sm_100a only):
//   * persistent CTAs (grid = #SMs), static round-robin tile scheduler, tiles 128 x 192, K-block =
//     one 128-byte swizzle row (64 bf16 or 32 fp32 elements)
//   * warp 0: TMA producer (cp.async.bulk.tensor.2d, SWIZZLE_128B, 4-stage mbarrier ring)
//   * warp 1: single-thread tcgen05.mma.cta_group::1 issuer, A and B from shared memory through UMMA
//     descriptors, accumulators in TMEM (2 x 192 columns, double buffered so the epilogue of tile i
//     overlaps the MMAs of tile i+1)
//   * warp 2: TMEM allocator;  warps 4-11 (residual epilogues) / 4-15 (store, GELU): epilogue. Each warp
//     owns a TMEM lane quarter and a 96- / 64-column slice of the tile: tcgen05.ld 32x32b (thread = row) -> per-warp shared-memory transpose
//     -> fused bias / GELU / gate*y+residual on *row-contiguous* float4s -> fully coalesced
//     128-bit global loads/stores (4 x 128-byte lines per warp instruction)
// TF32 mode: operands are fp32 bit patterns already rounded to TF32 (round-to-nearest) by their
// producers (weights at pack time, activations by the LN / attention / GELU epilogues), so the tensor
// core's truncation of the low 13 mantissa bits is exact. bf16 mode: the producers store bf16.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "gemm_simt.cuh"

namespace mdgen {

constexpr int TC_BM = 128;
constexpr int TC_BN = 192;
constexpr int TC_BK = 32;                       // fp32 elements per K-block (128 bytes)
constexpr int TC_STAGES = 4;
constexpr int TC_EPI_PITCH = 36;                // floats per staged row (144 B: conflict-free float4)
// epilogue warps: 8 (two 96-column halves per TMEM lane quarter) for the residual epilogues, which
// need ~166 registers per thread; 12 (three 64-column thirds) for the store / GELU epilogues, which
// are instruction-issue bound and fit 128 registers
constexpr int TC_EPI_WARPS_MAX = 12;
constexpr int TC_EPI_BYTES = TC_EPI_WARPS_MAX * 32 * TC_EPI_PITCH * 4;
__host__ __device__ constexpr int tc_epi_warps(int mode) { return 12; }
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;   // 24 KB
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + TC_EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_TMEM_COLS = 512;               // 2 accumulator buffers x 192 columns (pow2 alloc)

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] · B[smem desc]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (1: unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B between 8-row groups)
//   [46,48) version = 1 (Blackwell) | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate, both K-major:
//   [4,6) c_format = 1 (F32) | [7,10) a_format = 2 (TF32) | [10,13) b_format = 2 (TF32)
//   [15] a_major = 0, [16] b_major = 0 (K) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same for kind::f16 with bf16 operands: a_format = b_format = 1 (BF16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f16 with fp16 operands: a_format = b_format = 0 (F16), fp32 accumulate
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// two floats -> packed bf16x2 (round-to-nearest-even), low half = first argument
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// erf-GELU with erf from Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7): two MUFU ops + ~10 FMA
// instead of erff()'s branchy ~25 instructions (the fc1 epilogue is instruction-issue bound).
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.0f);
  const float erf_v = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_v);
}

// Packed fp32x2 helpers (Blackwell: one issue slot per pair for every add / multiply / FMA).
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__host__ __device__ constexpr uint64_t splat2(uint32_t bits) { return ((uint64_t)bits << 32) | bits; }
// gelu(x) = max(x, 0) - |x|/2 * erfc(|x| / sqrt 2), with erfc(z) = 2^(z * P(z)) on z in [0, 4] (clamped: erfc(4) = 1.5e-8):
// P = degree-6 weighted-minimax fit of log2(erfc(z)) / z (tools/fit_erfc_poly.py; |gelu error| <= 4e-7 absolute, the same class
// as the Abramowitz-Stegun form above) -> ONE MUFU (ex2) per element instead of two (rcp + ex2): the fc1 epilogue sits on the
// MUFU pipe (16 / clk / SM; 393 M outputs per launch), everything else is packed fp32x2 arithmetic.
__device__ __forceinline__ void gelu_fast2(float& x0, float& x1) {
  constexpr uint64_t kC0 = splat2(0xBFD05F7Au) /*-1.6279137*/, kC1 = splat2(0xBF6B1796u) /*-0.91832864*/,
                     kC2 = splat2(0xBE1889EEu) /*-0.14896366*/, kC3 = splat2(0x3CF14663u) /*0.029452508*/,
                     kC4 = splat2(0xBB16E11Bu) /*-0.0023022357*/, kC5 = splat2(0xB9F1FF41u) /*-0.00046157281*/,
                     kC6 = splat2(0x38D22DABu) /*0.00010022087*/, kNegRsqrt2 = splat2(0xBF3504F3u) /*-0.70710678*/;
  const float z0 = fminf(fabsf(x0) * 0.70710678118654752440f, 4.0f), z1 = fminf(fabsf(x1) * 0.70710678118654752440f, 4.0f);
  const uint64_t z = pk2(z0, z1);
  uint64_t p = fma2(kC6, z, kC5);
  p = fma2(p, z, kC4);
  p = fma2(p, z, kC3);
  p = fma2(p, z, kC2);
  p = fma2(p, z, kC1);
  p = fma2(p, z, kC0);
  float q0, q1;
  upk2(mul2(p, z), q0, q1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
  upk2(fma2(mul2(z, kNegRsqrt2), pk2(e0, e1), pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), x0, x1);
}

// ---- vectorised epilogue on 4 consecutive columns ------------------------------------------------
template <int MODE>
__device__ __forceinline__ void tc_epilogue4(const Epilogue& ep, const float* gate_row, long long m, int n,
                                             float a0, float a1, float a2, float a3) {
  float4 v = make_float4(a0, a1, a2, a3);
  if (ep.bias) {
    float4 b = *reinterpret_cast<const float4*>(ep.bias + n);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (MODE == EPI_GELU) { v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w); }
  if (MODE == EPI_RESID_GATE) {
    float4 g = *reinterpret_cast<const float4*>(gate_row + n);
    float4 r = *reinterpret_cast<const float4*>(ep.resid + (size_t)m * ep.ldo + n);
    v.x = r.x + g.x * v.x; v.y = r.y + g.y * v.y; v.z = r.z + g.z * v.z; v.w = r.w + g.w * v.w;
  }
  if (MODE == EPI_RESID) {
    float4 r = *reinterpret_cast<const float4*>(ep.resid + (size_t)m * ep.ldo + n);
    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
  }
  if (ep.round_out) { v.x = round_operand(v.x, ep.round_out); v.y = round_operand(v.y, ep.round_out); v.z = round_operand(v.z, ep.round_out); v.w = round_operand(v.w, ep.round_out); }
  *reinterpret_cast<float4*>(ep.out + (size_t)m * ep.ldo + n) = v;
}

// BF16IN : operands are 16-bit - fp16 (default) or bf16, selected at run time by ep.half_fmt - (kind::f16 MMA,
//          K-block = 64 elements = 128 bytes, UMMA_K = 16) instead of
//          TF32-in-fp32 (kind::tf32, K-block = 32 elements, UMMA_K = 8). Same 128-byte swizzled rows,
//          same descriptors and 32-byte K advance, twice the MMA rate and half the shared-memory/L2
//          operand traffic per FLOP (which is what bounds the fp32-operand variant).
// BF16OUT: the epilogue stores the same 16-bit format (q|k|v, the fc1 hidden activations).
// TMAOUT:  (16-bit outputs, store / GELU epilogues) thread = accumulator row: bias / GELU / pack on the registers that
//          tcgen05.ld delivers, rows written into a 128-byte-swizzled 32 x 64 staging tile per warp and sent to global
//          memory by ONE bulk-tensor (TMA) store per warp and tile. No per-row address arithmetic, bounds predicates,
//          shared-memory read-back or st.global in the warps (12 -> ~2 instructions per output element); rows >= M are
//          clipped by the tensor map.
template <int MODE, bool BF16IN, bool BF16OUT, bool TMAOUT = false>
__global__ void __launch_bounds__(128 + 32 * tc_epi_warps(MODE), 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmB,
                                                         const __grid_constant__ CUtensorMap tmO, long long M,
                                                         int N, int K, Epilogue ep) {
  static_assert(!TMAOUT || (BF16OUT && (MODE == EPI_STORE || MODE == EPI_GELU)), "TMA-store epilogue: 16-bit store / GELU");
  constexpr int BKE = BF16IN ? 64 : 32;                   // elements per 128-byte K-block row
  constexpr int NEPI = tc_epi_warps(MODE);                // epilogue warps
  constexpr int ECOLS = TC_BN / (NEPI / 4);               // columns of the tile owned by one epilogue warp
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024-B alignment
  uint8_t* sgen = smem_raw + (sbase - raw);
  const uint32_t bar_base = sbase + TC_STAGES * TC_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * TC_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * TC_STAGES + 2 + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + TC_STAGES * TC_STAGE_BYTES + 8 * (2 * TC_STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = N / TC_BN;
  const long long m_blocks = (M + TC_BM - 1) / TC_BM;
  const long long tiles = m_blocks * n_blocks;
  const int kblocks = K / BKE;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), NEPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"((uint32_t)TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int m0 = (int)(tile / n_blocks) * TC_BM;
        const int n0 = (int)(tile % n_blocks) * TC_BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (ep.dbg & 1) {
            mbar_arrive(full_bar(stage));
          } else {
            mbar_expect_tx(full_bar(stage), TC_STAGE_BYTES);
            const uint32_t sa = sbase + stage * TC_STAGE_BYTES;
            tma_load_2d(sa, &tmA, full_bar(stage), kb * BKE, m0);
            tma_load_2d(sa + TC_A_BYTES, &tmB, full_bar(stage), kb * BKE, n0);
          }
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer (one thread) =====================
      const uint32_t idesc = BF16IN ? (ep.half_fmt == kFmtF16 ? umma_idesc_f16(TC_BM, TC_BN) : umma_idesc_bf16(TC_BM, TC_BN))
                                    : umma_idesc_tf32(TC_BM, TC_BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1u, bphase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(buf), bphase ^ 1u);           // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * TC_BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);               // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = sbase + stage * TC_STAGE_BYTES;
          const uint64_t adesc = umma_desc_k128(sa);
          const uint64_t bdesc = umma_desc_k128(sa + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            if (ep.dbg & 2) break;
            // advance 8 tf32 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-B units
            if (BF16IN)
              tc_mma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                          (uint32_t)((kb | k) != 0));
            else
              tc_mma_tf32(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                          (uint32_t)((kb | k) != 0));
          }
          tc_commit(empty_bar(stage));                      // smem slot free once these MMAs retire
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(buf));                          // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4 && TMAOUT) {
    // ===================== epilogue warps, TMA-store form (TMEM -> regs -> swizzled smem tile -> bulk store) ==========
    static_assert(!TMAOUT || ECOLS == 64, "one 64-column (128-byte) box per warp");
    const int q = warp & 3;                                 // TMEM lane quarter of this warp
    const int slice = (warp - 4) >> 2;                      // which 64-column slice of the tile
    const uint32_t stg = sbase + TC_STAGES * TC_STAGE_BYTES + 1024 + (warp - 4) * 4096;   // 32 rows x 128 B, 1 KB aligned
    const uint32_t row_addr = stg + lane * 128;
    const uint32_t sw = (uint32_t)(lane & 7);
    const bool f16 = ep.half_fmt == kFmtF16;
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u, bphase = (it >> 1) & 1u;
      const int m0 = (int)(tile / n_blocks) * TC_BM + q * 32;
      const int n0 = (int)(tile % n_blocks) * TC_BN + slice * 64;
      mbar_wait(tfull_bar(buf), bphase);
      tc_fence_after();
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // last tile's store has read the staging tile
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float4 b4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          b4[j] = ep.bias ? *reinterpret_cast<const float4*>(ep.bias + n0 + c * 32 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * TC_BN + slice * 64 + c * 32, v);
        tc_ld_wait();
        if (c == 1) {                                       // accumulator is in registers: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
        auto finish = [&](auto is_f16) {                    // bias (+ GELU) + pack, 16-bit format fixed at compile time
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a0 = __uint_as_float(v[4 * j]) + b4[j].x, a1 = __uint_as_float(v[4 * j + 1]) + b4[j].y;
            float a2 = __uint_as_float(v[4 * j + 2]) + b4[j].z, a3 = __uint_as_float(v[4 * j + 3]) + b4[j].w;
            if (MODE == EPI_GELU) { gelu_fast2(a0, a1); gelu_fast2(a2, a3); }
            if (decltype(is_f16)::value) { v[2 * j] = pack_f16x2_rn(a0, a1); v[2 * j + 1] = pack_f16x2_rn(a2, a3); }
            else { v[2 * j] = pack_bf16x2(a0, a1); v[2 * j + 1] = pack_bf16x2(a2, a3); }
          }
        };
        if (f16) finish(std::true_type{}); else finish(std::false_type{});
        // this row's 64 bytes of the chunk = 16-byte units 4c .. 4c+3 of the 128-byte row, XOR-swizzled by (row & 7)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((((uint32_t)(4 * c + u)) ^ sw) << 4)),
                       "r"(v[4 * u]), "r"(v[4 * u + 1]), "r"(v[4 * u + 2]), "r"(v[4 * u + 3]) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                     ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(n0), "r"(m0), "r"(stg) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp >= 4) {
    // ===================== epilogue warps (TMEM -> regs -> smem transpose -> global) =====================
    const int q = warp & 3;                                 // TMEM lane quarter of this warp
    const int half = (warp - 4) >> 2;                       // which ECOLS-column slice of the tile
    float* stg = reinterpret_cast<float*>(sgen + TC_STAGES * TC_STAGE_BYTES + 256) + (warp - 4) * 32 * TC_EPI_PITCH;
    const int rr0 = lane >> 3, c4 = lane & 7;
    // gate rows: base of the current step's modulation row (read once), + b * bstride rows per sample
    const float* gate_base = nullptr;
    if (MODE == EPI_RESID_GATE || MODE == EPI_GATE)
      gate_base = ep.mod.base + (size_t)(ep.mod.step_ptr ? *ep.mod.step_ptr : 0) * ep.mod.width + ep.gate_off;
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u, bphase = (it >> 1) & 1u;
      const long long mbase = (tile / n_blocks) * TC_BM + q * 32;
      const int n0 = (int)(tile % n_blocks) * TC_BN + half * ECOLS;
      if (MODE == EPI_RESID_GATE || MODE == EPI_RESID) {
        // pull the residual rows of this CTA's NEXT tile into L2 now (3 x 128-byte lines per lane cover
        // this warp's 32 rows x 96 columns), so the epilogue's residual loads two tiles from now hit L2
        // instead of exposing DRAM latency in every 32-column chunk
        const long long tnext = tile + gridDim.x;
        if (tnext < tiles) {
          const long long mn = (tnext / n_blocks) * TC_BM + q * 32 + lane;
          const int nn = (int)(tnext % n_blocks) * TC_BN + half * ECOLS;
          if (mn < M) {
            const float* pr = ep.resid + (size_t)mn * ep.ldo + nn;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pr));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + 32));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + 64));
          }
        }
      }
      // gate row of this warp's 32 token rows: one sample index per tile when the rows do not straddle a sample
      // boundary (two integer divisions per tile instead of two per row and chunk; the per-row form was ~2/3 of
      // the epilogue's instructions)
      bool gate_uniform = false;
      const float* gate_row0 = nullptr;
      if (MODE == EPI_RESID_GATE || MODE == EPI_GATE) {
        const long long mlast = (mbase + 31 < M ? mbase + 31 : M - 1);
        const long long b_first = mbase / ep.mod.tokens_per_b;
        gate_uniform = (mlast / ep.mod.tokens_per_b) == b_first;
        gate_row0 = gate_base + (size_t)(((int)b_first % ep.mod.bmod) * ep.mod.bstride) * ep.mod.width;
      }
      mbar_wait(tfull_bar(buf), bphase);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < ECOLS; c0 += 32) {
        // residual rows of this chunk are prefetched first (8 independent 128-bit loads per lane) so
        // their DRAM/L2 latency overlaps the TMEM load and the shared-memory transpose
        float4 res[8];
        if (MODE == EPI_RESID_GATE || MODE == EPI_RESID) {
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            const long long m = mbase + itr * 4 + rr0;
            res[itr] = (m < M) ? *reinterpret_cast<const float4*>(ep.resid + (size_t)m * ep.ldo + n0 + c0 + 4 * c4)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * TC_BN + half * ECOLS + c0, v);
        tc_ld_wait();
        // thread = row: write the 32 columns of this row into the warp's staging tile
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stg + lane * TC_EPI_PITCH + 4 * j) =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                          __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        const int n = n0 + c0 + 4 * c4;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias) bias4 = *reinterpret_cast<const float4*>(ep.bias + n);
        float4 gate4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((MODE == EPI_RESID_GATE || MODE == EPI_GATE) && gate_uniform)
          gate4 = *reinterpret_cast<const float4*>(gate_row0 + n);
        // 8 lanes cover one 128-byte row segment; a warp instruction covers 4 rows
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {
          const int rr = itr * 4 + rr0;
          const long long m = mbase + rr;
          float4 a = *reinterpret_cast<const float4*>(stg + rr * TC_EPI_PITCH + 4 * c4);
          if (m < M) {
            a.x += bias4.x; a.y += bias4.y; a.z += bias4.z; a.w += bias4.w;
            if (MODE == EPI_GELU) { gelu_fast2(a.x, a.y); gelu_fast2(a.z, a.w); }
            if (MODE == EPI_RESID_GATE) {
              float4 g = gate4;
              if (!gate_uniform) {
                const int bb = (int)(m / ep.mod.tokens_per_b) % ep.mod.bmod;
                g = *reinterpret_cast<const float4*>(gate_base + (size_t)(bb * ep.mod.bstride) * ep.mod.width + n);
              }
              a.x = res[itr].x + g.x * a.x; a.y = res[itr].y + g.y * a.y;
              a.z = res[itr].z + g.z * a.z; a.w = res[itr].w + g.w * a.w;
            }
            if (MODE == EPI_RESID) { a.x += res[itr].x; a.y += res[itr].y; a.z += res[itr].z; a.w += res[itr].w; }
            if (MODE == EPI_GATE) {
              float4 g = gate4;
              if (!gate_uniform) {
                const int bb = (int)(m / ep.mod.tokens_per_b) % ep.mod.bmod;
                g = *reinterpret_cast<const float4*>(gate_base + (size_t)(bb * ep.mod.bstride) * ep.mod.width + n);
              }
              a.x *= g.x; a.y *= g.y; a.z *= g.z; a.w *= g.w;
            }
            if (ep.round_out) { a.x = round_operand(a.x, ep.round_out); a.y = round_operand(a.y, ep.round_out); a.z = round_operand(a.z, ep.round_out); a.w = round_operand(a.w, ep.round_out); }
            if (BF16OUT) {
              uint2 pk;
              if (ep.half_fmt == kFmtF16) { pk.x = pack_f16x2_rn(a.x, a.y); pk.y = pack_f16x2_rn(a.z, a.w); }
              else { pk.x = pack_bf16x2(a.x, a.y); pk.y = pack_bf16x2(a.z, a.w); }
              *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(ep.out) + (size_t)m * ep.ldo + n) = pk;
            } else {
              *reinterpret_cast<float4*>(ep.out + (size_t)m * ep.ldo + n) = a;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Weight-stationary form of the 16-bit GEMM for K <= 384 (QKV, fc1: M = all tokens, K = 384).
// Measured on the kernel above with its operand loads / MMAs switched off (Epilogue::dbg): QKV runs 0.171 ms with MMAs
// only, 0.197 ms with operand fills only and 0.255 ms with both -- each 128 x 192 tile pulls 98 KB of A AND 147 KB of W
// from L2 for 18.9 MFLOP. Here a CTA owns ONE 192-row block of W for its whole life (147 KB resident in shared memory,
// loaded once) and walks down the token rows of that column block, so a tile costs 98 KB of operand fill instead of 245.
//   warp 0: TMA producer (W block once, then a 4-stage ring of 128 x 64 A blocks)   warp 1: MMA issuer
//   warp 2: TMEM allocator        warps 4-15: epilogue, bulk-tensor stores of 32 x 16 chunks (1 KB staging per warp)
// Measured (profiles/r2_gemm_epilogue.md): the ring's depth IN TIME is what bounds these K = 384 GEMMs, not bytes: QKV takes
// 0.51 / 0.33 / 0.24 / 0.23 ms with a 1 / 2 / 3 / 4-deep A ring (round trip MMA retire -> refill -> data landed ~0.7 us against 0.23 us
// of MMA work per K-block), an L2 prefetch of A ahead of the loads and a single load+MMA thread were both slower.
// CTA c works on column block c % n_blocks; the CTAs of one column block interleave its row tiles, so the CTAs that read
// the same A tile do so at about the same time (one DRAM read, the other column blocks hit L2).
constexpr int WS_MAXKB = 6;                                  // K <= 6 x 64
constexpr int WS_ASTAGES = 4;
constexpr int WS_A_BYTES = TC_BM * 128;                      // 16 KB: 128 rows x 64 16-bit elements
constexpr int WS_B_BYTES = TC_BN * 128;                      // 24 KB per K-block of the resident W block
constexpr int WS_STG_BYTES = 12 * 1024;                      // 32 rows x 32 B per epilogue warp
constexpr int WS_SMEM_BYTES = WS_MAXKB * WS_B_BYTES + WS_ASTAGES * WS_A_BYTES + 1024 /*barriers*/ + WS_STG_BYTES + 1024 /*align*/;

template <int MODE>
__global__ void __launch_bounds__(512, 1) gemm_tc_ws_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            const __grid_constant__ CUtensorMap tmO, long long M, int N,
                                                            int K, Epilogue ep) {
  static_assert(MODE == EPI_STORE || MODE == EPI_GELU, "weight-stationary GEMM: 16-bit store / GELU epilogues");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sB = sbase;                                               // [kblocks][192 rows][128 B] swizzle-128B
  const uint32_t sA = sbase + WS_MAXKB * WS_B_BYTES;                       // [stage][128 rows][128 B]
  const uint32_t bar_base = sA + WS_ASTAGES * WS_A_BYTES;
  auto full_bar = [&](int st) { return bar_base + 8u * st; };
  auto empty_bar = [&](int st) { return bar_base + 8u * (WS_ASTAGES + st); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * WS_ASTAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * WS_ASTAGES + 2 + b); };
  const uint32_t bfull_bar = bar_base + 8u * (2 * WS_ASTAGES + 4);
  const uint32_t tmem_slot_addr = bar_base + 8u * (2 * WS_ASTAGES + 5);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot_addr - smem_u32(smem_raw)));
  const uint32_t stg_base = bar_base + 1024;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = N / TC_BN;
  const long long m_blocks = (M + TC_BM - 1) / TC_BM;
  const int kblocks = K / 64;
  const int nb = (int)(blockIdx.x % n_blocks);                             // this CTA's column block of W
  const int rank = (int)(blockIdx.x / n_blocks);                           // its rank among the CTAs of that block
  const int group = (int)(gridDim.x / n_blocks) + (nb < (int)(gridDim.x % n_blocks) ? 1 : 0);
  const int n0 = nb * TC_BN;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int st = 0; st < WS_ASTAGES; ++st) { mbar_init(full_bar(st), 1); mbar_init(empty_bar(st), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 12); }
    mbar_init(bfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(tmem_slot_addr), "r"((uint32_t)TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      if (!(ep.dbg & 1)) {
        mbar_expect_tx(bfull_bar, (uint32_t)kblocks * WS_B_BYTES);
        for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(sB + kb * WS_B_BYTES, &tmB, bfull_bar, kb * 64, n0);
      } else {
        mbar_arrive(bfull_bar);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (long long mt = rank; mt < m_blocks; mt += group) {
        const int m0 = (int)mt * TC_BM;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (ep.dbg & 1) {
            mbar_arrive(full_bar(stage));
          } else {
            mbar_expect_tx(full_bar(stage), WS_A_BYTES);
            tma_load_2d(sA + stage * WS_A_BYTES, &tmA, full_bar(stage), kb * 64, m0);
          }
          if (++stage == WS_ASTAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer (one thread) =====================
      const uint32_t idesc = ep.half_fmt == kFmtF16 ? umma_idesc_f16(TC_BM, TC_BN) : umma_idesc_bf16(TC_BM, TC_BN);
      mbar_wait(bfull_bar, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (long long mt = rank; mt < m_blocks; mt += group, ++it) {
        const uint32_t buf = it & 1u, bphase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(buf), bphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * TC_BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_k128(sA + stage * WS_A_BYTES);
          const uint64_t bdesc = umma_desc_k128(sB + kb * WS_B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (ep.dbg & 2) break;
            tc_mma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          }
          tc_commit(empty_bar(stage));
          if (++stage == WS_ASTAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(buf));
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (TMEM -> regs -> swizzled 32 x 16 chunk -> bulk-tensor store) ==========
    // 16-column chunks: 1 KB of staging per warp, which is what lets the A ring be 4 deep next to the resident W block
    const int q = warp & 3;
    const int slice = (warp - 4) >> 2;
    const uint32_t stg = stg_base + (warp - 4) * 1024;                     // 32 rows x 32 B, swizzle-32B
    const uint32_t row_addr = stg + lane * 32;
    const uint32_t sw = (uint32_t)((lane >> 2) & 1);
    const bool f16 = ep.half_fmt == kFmtF16;
    uint32_t it = 0;
    for (long long mt = rank; mt < m_blocks; mt += group, ++it) {
      const uint32_t buf = it & 1u, bphase = (it >> 1) & 1u;
      const int m0 = (int)mt * TC_BM + q * 32;
      const int nc = n0 + slice * 64;
      mbar_wait(tfull_bar(buf), bphase);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 b4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          b4[j] = ep.bias ? *reinterpret_cast<const float4*>(ep.bias + nc + c * 16 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t v[16];
        tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * TC_BN + slice * 64 + c * 16, v);
        tc_ld_wait();
        if (c == 3) {                                       // accumulator is in registers: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
        auto finish = [&](auto is_f16) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float a0 = __uint_as_float(v[4 * j]) + b4[j].x, a1 = __uint_as_float(v[4 * j + 1]) + b4[j].y;
            float a2 = __uint_as_float(v[4 * j + 2]) + b4[j].z, a3 = __uint_as_float(v[4 * j + 3]) + b4[j].w;
            if (MODE == EPI_GELU) { gelu_fast2(a0, a1); gelu_fast2(a2, a3); }
            if (decltype(is_f16)::value) { v[2 * j] = pack_f16x2_rn(a0, a1); v[2 * j + 1] = pack_f16x2_rn(a2, a3); }
            else { v[2 * j] = pack_bf16x2(a0, a1); v[2 * j + 1] = pack_bf16x2(a2, a3); }
          }
        };
        if (f16) finish(std::true_type{}); else finish(std::false_type{});
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous chunk's store has read the staging tile
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 2; ++u)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((((uint32_t)u) ^ sw) << 4)),
                       "r"(v[4 * u]), "r"(v[4 * u + 1]), "r"(v[4 * u + 2]), "r"(v[4 * u + 3]) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(nc + c * 16), "r"(m0), "r"(stg) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS) : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------
// Launch attributes (cudaFuncSetAttribute) and the SM count are per device: a process may own
// handles on several GPUs, so the "already configured" state is kept per device ordinal.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}
inline int device_sm_count() {
  static int n[kMaxDevices] = {0};
  const int d = current_device();
  if (!n[d]) cudaDeviceGetAttribute(&n[d], cudaDevAttrMultiProcessorCount, d);
  return n[d];
}
// debugging / A-B switch: MDGEN_NO_TMA_OUT=1 keeps the 16-bit-output GEMMs on the shared-memory-transpose epilogue
inline bool tc_no_tma_out() {
  static const bool v = [] { const char* e = getenv("MDGEN_NO_TMA_OUT"); return e && e[0] == '1'; }();
  return v;
}
// MDGEN_NO_WS=1 keeps the K <= 384 GEMMs on the tile-streaming kernel
inline bool tc_no_ws() {
  static const bool v = [] { const char* e = getenv("MDGEN_NO_WS"); return e && e[0] == '1'; }();
  return v;
}
inline bool tc_gemm_supported(int N, int K, bool bf16 = false) {
  const int bk = bf16 ? 64 : TC_BK;
  return (N % TC_BN == 0) && (K % bk == 0) && K >= bk;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn(std::string* err) {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    if (err) *err = std::string("cuTensorMapEncodeTiled unavailable: ") + cudaGetErrorString(e);
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// 2-D fp32 row-major [rows, cols] (leading dimension ld elements), box = [box_rows, 32 cols], SWIZZLE_128B
// (box_cols = 16 with 16-bit elements: 32-byte rows, SWIZZLE_32B -- the store chunks of the weight-stationary kernel)
inline int make_tmap_2d(CUtensorMap* map, const void* ptr, long long rows, int cols, int ld, int box_rows,
                        bool bf16, std::string* err, int box_cols = 0) {
  PFN_encodeTiled fn = get_encode_fn(err);
  if (!fn) return -2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (bf16 ? 2 : 4)};
  if (!box_cols) box_cols = bf16 ? 64 : TC_BK;                                   // 128-byte rows either way
  const int box_bytes = box_cols * (bf16 ? 2 : 4);
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return -2;
  }
  return 0;
}

struct TmapCacheEntry {
  const void* ptr; long long rows; int cols, ld, box_rows; bool bf16; int box_cols;
  CUtensorMap map;
};

inline int get_tmap(const void* ptr, long long rows, int cols, int ld, int box_rows, bool bf16, CUtensorMap* out,
                    std::string* err, int box_cols = 0) {
  static std::vector<TmapCacheEntry> cache;
  static std::mutex mu;                      // handles on different host threads share this cache
  std::lock_guard<std::mutex> lock(mu);
  for (auto& e : cache)
    if (e.ptr == ptr && e.rows == rows && e.cols == cols && e.ld == ld && e.box_rows == box_rows &&
        e.bf16 == bf16 && e.box_cols == box_cols) {
      *out = e.map;
      return 0;
    }
  TmapCacheEntry e{ptr, rows, cols, ld, box_rows, bf16, box_cols, {}};
  int rc = make_tmap_2d(&e.map, ptr, rows, cols, ld, box_rows, bf16, err, box_cols);
  if (rc) return rc;
  if (cache.size() > 4096) cache.clear();
  cache.push_back(e);
  *out = e.map;
  return 0;
}

template <int MODE, bool BF16IN, bool BF16OUT, bool TMAOUT = false>
inline int tc_launch_mode(const CUtensorMap& ta, const CUtensorMap& tb, long long M, int N, int K,
                          const Epilogue& ep, cudaStream_t s, int grid, std::string* err,
                          const CUtensorMap* to = nullptr) {
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<MODE, BF16IN, BF16OUT, TMAOUT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(gemm_tc): ") + cudaGetErrorString(e);
      return -2;
    }
    configured[dev] = true;
  }
  gemm_tc_kernel<MODE, BF16IN, BF16OUT, TMAOUT><<<grid, 128 + 32 * tc_epi_warps(MODE), TC_SMEM_BYTES, s>>>(
      ta, tb, to ? *to : ta, M, N, K, ep);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("gemm_tc launch: ") + cudaGetErrorString(e);
    return -2;
  }
  return 0;
}

template <int MODE>
inline int tc_launch_ws(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, long long M, int N, int K,
                        const Epilogue& ep, cudaStream_t s, int grid, std::string* err) {
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_ws_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES);
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(gemm_tc_ws): ") + cudaGetErrorString(e);
      return -2;
    }
    configured[dev] = true;
  }
  gemm_tc_ws_kernel<MODE><<<grid, 512, WS_SMEM_BYTES, s>>>(ta, tb, to, M, N, K, ep);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("gemm_tc_ws launch: ") + cudaGetErrorString(e);
    return -2;
  }
  return 0;
}

// A / W: fp32 (TF32-rounded) or bf16 buffers according to `in_bf16`; lda / ldw / ep.ldo in elements.
inline int tc_gemm_launch(int mode, const void* A, int lda, const void* W, int ldw, long long M, int N, int K,
                          const Epilogue& ep, cudaStream_t s, std::string* err, bool in_bf16 = false,
                          bool out_bf16 = false) {
  const int num_sms = device_sm_count();
  CUtensorMap ta, tb;
  if (get_tmap(A, M, K, lda, TC_BM, in_bf16, &ta, err)) return -2;
  if (get_tmap(W, N, K, ldw, TC_BN, in_bf16, &tb, err)) return -2;
  long long tiles = ((M + TC_BM - 1) / TC_BM) * (N / TC_BN);
  int grid = (int)std::min<long long>(tiles, num_sms);
  if (!in_bf16) {
    switch (mode) {
      case EPI_STORE: return tc_launch_mode<EPI_STORE, false, false>(ta, tb, M, N, K, ep, s, grid, err);
      case EPI_GELU: return tc_launch_mode<EPI_GELU, false, false>(ta, tb, M, N, K, ep, s, grid, err);
      case EPI_RESID_GATE: return tc_launch_mode<EPI_RESID_GATE, false, false>(ta, tb, M, N, K, ep, s, grid, err);
      case EPI_RESID: return tc_launch_mode<EPI_RESID, false, false>(ta, tb, M, N, K, ep, s, grid, err);
      case EPI_GATE: return tc_launch_mode<EPI_GATE, false, false>(ta, tb, M, N, K, ep, s, grid, err);
    }
  } else if (!out_bf16) {
    switch (mode) {
      case EPI_STORE: return tc_launch_mode<EPI_STORE, true, false>(ta, tb, M, N, K, ep, s, grid, err);
      case EPI_GELU: return tc_launch_mode<EPI_GELU, true, false>(ta, tb, M, N, K, ep, s, grid, err);
      case EPI_RESID_GATE: return tc_launch_mode<EPI_RESID_GATE, true, false>(ta, tb, M, N, K, ep, s, grid, err);
      case EPI_RESID: return tc_launch_mode<EPI_RESID, true, false>(ta, tb, M, N, K, ep, s, grid, err);
      case EPI_GATE: return tc_launch_mode<EPI_GATE, true, false>(ta, tb, M, N, K, ep, s, grid, err);
    }
  } else {
    // 16-bit output: bulk-tensor stores when the output rows satisfy TMA's 16-byte address / stride rules
    const bool tma_out = !tc_no_tma_out() && ep.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0 &&
                         ep.round_out == 0 && M < (1ll << 31);
    if (tma_out && !tc_no_ws() && K <= WS_MAXKB * 64 && tiles >= 4ll * num_sms && N / TC_BN <= grid && (mode == EPI_STORE || mode == EPI_GELU)) {
      CUtensorMap to;
      if (get_tmap(ep.out, M, N, ep.ldo, 32, true, &to, err, 16)) return -2;
      return mode == EPI_STORE ? tc_launch_ws<EPI_STORE>(ta, tb, to, M, N, K, ep, s, grid, err)
                               : tc_launch_ws<EPI_GELU>(ta, tb, to, M, N, K, ep, s, grid, err);
    }
    if (tma_out) {
      CUtensorMap to;
      if (get_tmap(ep.out, M, N, ep.ldo, 32, true, &to, err)) return -2;
      switch (mode) {
        case EPI_STORE: return tc_launch_mode<EPI_STORE, true, true, true>(ta, tb, M, N, K, ep, s, grid, err, &to);
        case EPI_GELU: return tc_launch_mode<EPI_GELU, true, true, true>(ta, tb, M, N, K, ep, s, grid, err, &to);
      }
    }
    switch (mode) {
      case EPI_STORE: return tc_launch_mode<EPI_STORE, true, true>(ta, tb, M, N, K, ep, s, grid, err);
      case EPI_GELU: return tc_launch_mode<EPI_GELU, true, true>(ta, tb, M, N, K, ep, s, grid, err);
    }
  }
  if (err) *err = "bad epilogue mode / dtype combination";
  return -1;
}

}  // namespace mdgen
