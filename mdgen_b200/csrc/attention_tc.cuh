// tcgen05 fused attention for long sequences (mha_t over T frames, mha_l over L >= 65 residues):
// softmax(Q K^T + key-padding) V with bias-KV token and RoPE, scores never leave the SM
// (restates mdgen/model/mha.py:260-397; K8-K13 of SURVEY.md §2c collapse into this kernel).
//
// One CTA = one (sequence, head, 128-query tile); 320 threads, two CTAs per SM (256 TMEM columns and
// ~95 KB of shared memory each). Inside a CTA the tensor pipe runs ahead of / behind the softmax:
// S is double buffered in TMEM (QK^T of tile g+1 is issued before the softmax of tile g starts) and
// O accumulates in TMEM across key tiles (P·V of tile g runs while tile g+1 is exponentiated), so
// no MMA latency sits on the softmax critical path:
//   (attn_prep_kernel)   : pre-pass that builds, once per (sequence, head), the UMMA-ready images of
//                          every 96-key tile (K rotated by RoPE, V transposed, TF32-rounded,
//                          K-major SWIZZLE_128B, plus additive key mask and |k| bound) in a global
//                          scratch; the CTAs of all query tiles of that (sequence, head) are adjacent
//                          in the grid, so they share those images through L2.
//   warp 8 lane 0        : producer - one cp.async.bulk (TMA 1-D) per key tile refills a 3-stage
//                          shared-memory ring from that scratch.
//   warp 9 lane 0        : the single MMA-issuing thread (tcgen05.mma + tcgen05.commit).
//   warps 0-7  "softmax" : warp w owns the 32 query rows of TMEM lane quarter (w & 3) and the column
//                          half (w >> 2) of every S tile: reads S from TMEM, fp32 online softmax,
//                          writes P back over S in TMEM. Warps run decoupled: they meet the MMA warp
//                          through mbarriers (s_full / p_ready per S buffer); the two warps sharing
//                          a row only exchange data on the rare exact-max path and in the epilogue.
// Per key tile:  S[128x96] = Q·K^T    (3 x tcgen05.mma kind::tf32, K = 24 = 3 x 8, A/B from smem)
//                P = exp2(S - m)      (softmax warps, TMEM -> regs -> TMEM, in place)
//                O[128x32] += P·V     (12 x tcgen05.mma kind::tf32, A = P from TMEM, B = V^T smem);
//                                     O lives in TMEM and is rescaled in place only when a row's softmax
//                                     reference moves
// Online softmax reference: the exact two-pass (row max, then exp) is used for a tile only when the
// row has no reference yet or when the Cauchy-Schwarz bound |q_i|*max_j|k_j| could exceed the
// reference by 2^100; otherwise the tile is exponentiated in a single TMEM pass against the
// existing reference (no overflow is possible, and softmax is shift invariant).
// head_dim 24 is padded to 32 only in shared-memory row pitch (128-byte rows); the QK^T MMAs read
// just the three valid 32-byte K-chunks. MUFU ex2 is the roofline of this kernel, not the tensor
// pipe: 96 MMA flops per score element vs. one exp.
#pragma once
#include "attention_simt.cuh"
#include "gemm_tc.cuh"

namespace mdgen {

constexpr int AT_QT = 128;                      // queries per tile (UMMA M)
constexpr int AT_KT = 96;                       // keys per tile (UMMA N of QK^T, K of PV)
constexpr int AT_THREADS = 320;                 // 8 softmax warps + TMA producer warp + MMA issuer warp
constexpr int AT_HALF = AT_KT / 2;              // columns of an S tile handled by one softmax warp
constexpr int AT_Q_BYTES = 128 * 128;           // Q tile: 128 rows x 128-byte pitch
constexpr int AT_K_BYTES = AT_KT * 128;         // K tile
constexpr int AT_KM_FLOATS = 128 + 4 + 4;       // key mask | per-slice "has masked key" flags | per-slice max |k|
// V^T image: k-atoms of [32 d-rows x 128 bytes] (SWIZZLE_128B): 32 TF32 keys per atom, or 64 bf16 keys per
// atom when the P·V product runs on kind::f16 (PV16: P and V in bf16, half the MMAs and half the P stores)
__host__ __device__ constexpr int at_vt_bytes(bool pv16) { return pv16 ? 2 * 4096 : (AT_KT / 32) * 4096; }
__host__ __device__ constexpr int at_img_bytes(bool pv16) { return AT_K_BYTES + at_vt_bytes(pv16) + AT_KM_FLOATS * 4; }
constexpr int AT_STAGE_BYTES = 26 * 1024;       // smem stage pitch (keeps K / V^T 1024-byte aligned)
constexpr int AT_STAGES = 3;
constexpr int AT_SMEM_BYTES = 1024 /*align*/ + AT_Q_BYTES + AT_STAGES * AT_STAGE_BYTES + 128 /*barriers*/ + 4 * 512 /*row exchange*/;
constexpr int AT_TMEM_COLS = 256;               // S/P double buffer: cols [0,96) [96,192); O_even / O_odd: cols [192,224) [224,256)

__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// D[tmem] (+)= A[tmem] · B[smem desc]^T, kind::tf32
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// mbarrier wait with a nanosleep back-off: used by the loader warps so their polling does not
// steal issue slots from the softmax warps sharing the SM sub-partitions.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(256);
  }
}
// D[tmem] (+)= A[tmem, bf16 pairs] · B[smem desc]^T, kind::f16
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// byte offset of 16-byte chunk `c` of row `r` inside a [rows x 128 B] K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// Pre-pass: builds, once per (sequence, head), the UMMA-ready images of every 96-key tile in a
// global scratch: K rotated by RoPE (TF32), V transposed (TF32, or bf16 when PV16), K-major
// SWIZZLE_128B layout, plus the additive key mask, per-slice "has masked key" flags and per-slice
// max |k| (overflow bound).
// The attention CTAs of all query tiles of that (sequence, head) then refill their shared-memory
// ring from these images with plain 1-D TMA bulk copies (L2 hits: they run concurrently).
template <bool PV16>
__global__ void __launch_bounds__(256) attn_prep_kernel(AttnParams p, uint8_t* __restrict__ scratch) {
  constexpr int IMG = at_img_bytes(PV16);
  const SeqMap& sm = p.sm;
  const int tid = threadIdx.x, lane = tid & 31;
  const int h = blockIdx.x % kH;
  const long long s = blockIdx.x / kH;
  const int S = sm.S, nkeys = S + 1;
  const int nkt = (nkeys + AT_KT - 1) / AT_KT;
  uint8_t* img = scratch + (size_t)blockIdx.x * nkt * IMG;
  // (rows 24..31 of every V^T atom are never written: the scratch is zero-filled at allocation;
  //  stale data there only reaches the unused accumulator columns 24..31)
  for (int j = tid; j < nkt * AT_KT; j += 256) {
    const int kt = j / AT_KT, r = j - kt * AT_KT;
    uint8_t* kbase = img + (size_t)kt * IMG;
    uint8_t* vbase = kbase + AT_K_BYTES;
    float* kmask = reinterpret_cast<float*>(vbase + at_vt_bytes(PV16));
    float k[kHD], v[kHD];
    float mval = 0.f;
    if (j < S) {
      long long tk = seq_token(sm, s, j);
      load24(p.qkv, (size_t)tk * kQKV + kC + h * kHD, p.qkv_fmt, k);
      load24(p.qkv, (size_t)tk * kQKV + 2 * kC + h * kHD, p.qkv_fmt, v);
      if (p.mask && p.mask[tk] == 0.f) mval = -INFINITY;
    } else if (j == S) {
#pragma unroll
      for (int i = 0; i < kHD; ++i) { k[i] = p.bias_k[h * kHD + i]; v[i] = p.bias_v[h * kHD + i]; }
    } else {
#pragma unroll
      for (int i = 0; i < kHD; ++i) { k[i] = 0.f; v[i] = 0.f; }
      mval = -INFINITY;
    }
    if (j <= S) rope24(k, p.cosT + j * kHalf, p.sinT + j * kHalf);
    float kn2 = 0.f;
#pragma unroll
    for (int i = 0; i < kHD; ++i) { k[i] = round_tf32(k[i]); kn2 = fmaf(k[i], k[i], kn2); }
#pragma unroll
    for (int c = 0; c < 6; ++c)
      *reinterpret_cast<float4*>(kbase + sw128_off(r, c)) = make_float4(k[4*c], k[4*c+1], k[4*c+2], k[4*c+3]);
    if (PV16) {
      uint8_t* ab = vbase + (r >> 6) * 4096;           // 64 bf16 keys per 128-byte row
      const int kc = (r & 63) >> 3, kw = r & 7;
#pragma unroll
      for (int d = 0; d < kHD; ++d)
        reinterpret_cast<uint16_t*>(ab + sw128_off(d, kc))[kw] = (uint16_t)(__float_as_uint(round_bf16_rn(v[d])) >> 16);
    } else {
      uint8_t* ab = vbase + (r >> 5) * 4096;
      const int kc = (r & 31) >> 2, kw = r & 3;
#pragma unroll
      for (int d = 0; d < kHD; ++d)
        reinterpret_cast<float*>(ab + sw128_off(d, kc))[kw] = round_tf32_fast(v[d]);
    }
    kmask[r] = mval;
    // per 32-key slice: "has a masked / out-of-range key" flag and max |k| (for the overflow bound)
    const unsigned any = __ballot_sync(0xffffffffu, mval != 0.f);
    const float kn = warp_max(sqrtf(kn2));
    if (lane == 0) {
      reinterpret_cast<int*>(kmask + 128)[r >> 5] = any ? 1 : 0;
      kmask[132 + (r >> 5)] = kn;
    }
  }
}

// Staged pre-pass (bf16 q|k|v only): one block per (sequence, key tile, head octet). The 96 k|v row
// segments of the octet (2 x 384 contiguous bytes per token) are fetched with 16-byte cp.async into shared
// memory - every global sector is requested once and all loads of the block are in flight together -
// then each warp builds the image slice of one (head, 32-key slice) exactly like attn_prep_kernel.
constexpr int AP2_PITCH = 784;                          // bytes per staged row: 768 + 16 (odd multiple of 16 B)
constexpr int AP2_SMEM_BYTES = AT_KT * AP2_PITCH;
template <bool PV16>
__global__ void __launch_bounds__(256) attn_prep2_kernel(AttnParams p, uint8_t* __restrict__ scratch) {
  constexpr int IMG = at_img_bytes(PV16);
  extern __shared__ __align__(16) uint8_t ap2_smem[];
  const SeqMap& sm = p.sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = sm.S, nkeys = S + 1;
  const int nkt = (nkeys + AT_KT - 1) / AT_KT;
  const int hg = blockIdx.x & 1;
  const int kt = (int)((blockIdx.x >> 1) % nkt);
  const long long s = (blockIdx.x >> 1) / nkt;
  const uint16_t* qkv = reinterpret_cast<const uint16_t*>(p.qkv);
  // ---- stage 1: rows of this key tile -> shared memory ([k octet | v octet], 48 chunks of 16 bytes)
  for (int row = warp; row < AT_KT; row += 8) {
    const int j = kt * AT_KT + row;
    if (j < S) {
      const long long tk = seq_token(sm, s, j);
      const uint16_t* src = qkv + (size_t)tk * kQKV + kC + hg * 192;
      uint8_t* dst = ap2_smem + row * AP2_PITCH;
      cp_async16(dst + lane * 16, src + (lane < 24 ? lane * 8 : kC + (lane - 24) * 8));
      if (lane < 16) cp_async16(dst + (32 + lane) * 16, src + kC + (8 + lane) * 8);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- stage 2: 8 heads x 96 keys = 768 items, a warp = one (head, 32-key slice)
#pragma unroll 1
  for (int it = 0; it < 3; ++it) {
    const int item = it * 256 + tid;
    const int hl = item / AT_KT, r = item - hl * AT_KT;
    const int h = hg * 8 + hl;
    const int j = kt * AT_KT + r;
    uint8_t* kbase = scratch + ((size_t)(s * kH + h) * nkt + kt) * IMG;
    uint8_t* vbase = kbase + AT_K_BYTES;
    float* kmask = reinterpret_cast<float*>(vbase + at_vt_bytes(PV16));
    float k[kHD], v[kHD];
    float mval = 0.f;
    if (j < S) {
      const uint8_t* rowp = ap2_smem + r * AP2_PITCH;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const uint4 uk = *reinterpret_cast<const uint4*>(rowp + hl * 48 + i * 16);
        const uint4 uv = *reinterpret_cast<const uint4*>(rowp + 384 + hl * 48 + i * 16);
        const uint32_t wk[4] = {uk.x, uk.y, uk.z, uk.w}, wv[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float2 fk = unpack_half2(wk[c], p.qkv_fmt), fv = unpack_half2(wv[c], p.qkv_fmt);
          k[8 * i + 2 * c] = fk.x;
          k[8 * i + 2 * c + 1] = fk.y;
          v[8 * i + 2 * c] = fv.x;
          v[8 * i + 2 * c + 1] = fv.y;
        }
      }
      if (p.mask && p.mask[seq_token(sm, s, j)] == 0.f) mval = -INFINITY;
    } else if (j == S) {
#pragma unroll
      for (int i = 0; i < kHD; ++i) { k[i] = p.bias_k[h * kHD + i]; v[i] = p.bias_v[h * kHD + i]; }
    } else {
#pragma unroll
      for (int i = 0; i < kHD; ++i) { k[i] = 0.f; v[i] = 0.f; }
      mval = -INFINITY;
    }
    if (j <= S) rope24(k, p.cosT + j * kHalf, p.sinT + j * kHalf);
    float kn2 = 0.f;
#pragma unroll
    for (int i = 0; i < kHD; ++i) { k[i] = round_tf32(k[i]); kn2 = fmaf(k[i], k[i], kn2); }
#pragma unroll
    for (int c = 0; c < 6; ++c)
      *reinterpret_cast<float4*>(kbase + sw128_off(r, c)) = make_float4(k[4*c], k[4*c+1], k[4*c+2], k[4*c+3]);
    if (PV16) {
      uint8_t* ab = vbase + (r >> 6) * 4096;
      const int kc = (r & 63) >> 3, kw = r & 7;
#pragma unroll
      for (int d = 0; d < kHD; ++d)
        reinterpret_cast<uint16_t*>(ab + sw128_off(d, kc))[kw] = (uint16_t)(__float_as_uint(round_bf16_rn(v[d])) >> 16);
    } else {
      uint8_t* ab = vbase + (r >> 5) * 4096;
      const int kc = (r & 31) >> 2, kw = r & 3;
#pragma unroll
      for (int d = 0; d < kHD; ++d)
        reinterpret_cast<float*>(ab + sw128_off(d, kc))[kw] = round_tf32_fast(v[d]);
    }
    kmask[r] = mval;
    const unsigned any = __ballot_sync(0xffffffffu, mval != 0.f);
    const float kn = warp_max(sqrtf(kn2));
    if (lane == 0) {
      reinterpret_cast<int*>(kmask + 128)[r >> 5] = any ? 1 : 0;
      kmask[132 + (r >> 5)] = kn;
    }
  }
}

// PV16: P and V^T in bf16, P·V on kind::f16 (6 MMAs per key tile instead of 12, half the P stores); the
//       default. PV16 = false keeps P and V in TF32 (kind::tf32) as the higher-precision reference variant.
// BREF: experiment (not validated on hardware yet, off by default): a row that has no softmax reference yet adopts
//       the Cauchy-Schwarz bound |q_i| max_j |k_j| of its first key tile as reference (when that bound is small
//       enough that nothing can underflow) instead of the exact row maximum, which removes the second TMEM read of
//       the first key tile and the row-max exchange between the two column-half warps.
template <bool PV16, bool BREF>
__global__ void __launch_bounds__(AT_THREADS, 2) attn_tc_kernel(AttnParams p, const uint8_t* __restrict__ scratch) {
  constexpr int IMG = at_img_bytes(PV16);
  constexpr int VT = at_vt_bytes(PV16);
  extern __shared__ uint8_t smem_raw[];
  const SeqMap& sm = p.sm;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  // carve-up
  const uint32_t q_off = 0;
  const uint32_t st_off = AT_Q_BYTES;                          // stage s: [K | V^T | mask block]
  const uint32_t bar_off = AT_Q_BYTES + AT_STAGES * AT_STAGE_BYTES;
  auto b_sfull = [&](int b) { return sbase + bar_off + 8 * b; };             // [2] MMA -> softmax
  auto b_kvfull = [&](int s) { return sbase + bar_off + 16 + 8 * s; };       // [3] TMA -> MMA
  auto b_kvfree = [&](int s) { return sbase + bar_off + 40 + 8 * s; };       // [3] MMA -> TMA
  // O-valid barriers, one per key-tile parity: a waiter is then never more than one phase away from the phase
  // it waits for (with a single barrier a softmax warp one tile ahead of the slowest warp could pass a parity
  // wait for P·V(g) while P·V(g-1) is still pending)
  auto b_odone = [&](int b) { return sbase + bar_off + (b ? 96 : 64); };     // [2] MMA -> softmax (O valid)
  auto b_pready = [&](int b) { return sbase + bar_off + 72 + 8 * b; };       // [2] softmax -> MMA
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + bar_off + 88);
  float* xch = reinterpret_cast<float*>(sgen + bar_off + 128);   // [4][128]: qnorm | max half0 | max half1 | l half1

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = sm.S, nkeys = S + 1;
  const int nqt = (S + AT_QT - 1) / AT_QT, nkt = (nkeys + AT_KT - 1) / AT_KT;
  const int qt = blockIdx.x % nqt;                      // query tile of this CTA (fastest: the CTAs
  const long long sh = blockIdx.x / nqt;                // sharing one set of key images run together)
  const int h = (int)(sh % kH);
  const long long s = sh / kH;
  const uint8_t* img = scratch + (size_t)sh * nkt * IMG;

  // one cp.async.bulk (TMA 1-D) per key tile into stage g % AT_STAGES
  auto produce = [&](int g) {
    const int st = g % AT_STAGES, use = g / AT_STAGES;
    if (use > 0) mbar_wait_backoff(b_kvfree(st), (uint32_t)((use & 1) ^ 1));   // PV of the previous user retired
    mbar_expect_tx(b_kvfull(st), IMG);
    const uint32_t dst = sbase + st_off + st * AT_STAGE_BYTES;
    const uint8_t* src = img + (size_t)g * IMG;
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"((uint32_t)IMG), "r"(b_kvfull(st))
        : "memory");
  };
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(b_sfull(b), 1); mbar_init(b_pready(b), 8); }
    mbar_init(b_odone(0), 1); mbar_init(b_odone(1), 1);
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(b_kvfull(i), 1); mbar_init(b_kvfree(i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"((uint32_t)AT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- stage the Q tile (softmax threads: one row each), RoPE'd, scaled by log2(e), TF32-rounded
  const int r = tid;
  const int e = qt * AT_QT + r;

  long long tq = 0;
  float qnorm = 0.f;
  if (tid < 128) {
    tq = seq_token(sm, s, e < S ? e : S - 1);
    float q[kHD];
    load24(p.qkv, (size_t)tq * kQKV + h * kHD, p.qkv_fmt, q);
    const int pe = e < S ? e : S - 1;
    rope24(q, p.cosT + pe * kHalf, p.sinT + pe * kHalf);
    float qn2 = 0.f;
#pragma unroll
    for (int i = 0; i < kHD; ++i) { q[i] = round_tf32(q[i] * 1.4426950408889634f); qn2 = fmaf(q[i], q[i], qn2); }
    qnorm = sqrtf(qn2) * 1.001f;
    xch[r] = qnorm;
#pragma unroll
    for (int c = 0; c < 6; ++c)
      *reinterpret_cast<float4*>(sgen + q_off + sw128_off(r, c)) = make_float4(q[4*c], q[4*c+1], q[4*c+2], q[4*c+3]);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 2 * AT_KT;

  if (warp == 8) {
    if (lane == 0) {
      // =============================== producer (TMA 1-D bulk copies) ===============================
      for (int g = 0; g < nkt; ++g) produce(g);
    }
  } else if (warp == 9) {
    if (lane == 0) {
      // =============================== MMA issuer (one thread) ===============================
      constexpr uint32_t idesc_qk = umma_idesc_tf32(AT_QT, AT_KT);
      constexpr uint32_t idesc_pv = PV16 ? umma_idesc_bf16(AT_QT, 32) : umma_idesc_tf32(AT_QT, 32);
      constexpr int NKS = PV16 ? AT_KT / 16 : AT_KT / 8;     // k-steps of the P·V product
      const uint64_t adesc = umma_desc_k128(sbase + q_off);
      // S[gg & 1] = Q · K(gg)^T
      auto issue_qk = [&](int gg) {
        const int st = gg % AT_STAGES, use = gg / AT_STAGES;
        mbar_wait(b_kvfull(st), (uint32_t)(use & 1));
        tc_fence_after();
        const uint64_t bdesc = umma_desc_k128(sbase + st_off + st * AT_STAGE_BYTES);
        const uint32_t tS = tmem_base + (gg & 1) * AT_KT;
#pragma unroll
        for (int k = 0; k < 3; ++k)
          tc_mma_tf32(tS, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_qk, (uint32_t)(k != 0));
        tc_commit(b_sfull(gg & 1));
      };
      issue_qk(0);
      if (nkt > 1) issue_qk(1);
      for (int g = 0; g < nkt; ++g) {
        const int st = g % AT_STAGES, buf = g & 1;
        mbar_wait(b_pready(buf), (uint32_t)((g >> 1) & 1));   // all eight softmax warps wrote P(g)
        tc_fence_after();
        const uint32_t tmem_S = tmem_base + buf * AT_KT;
        const uint64_t vdesc = umma_desc_k128(sbase + st_off + st * AT_STAGE_BYTES + AT_K_BYTES);
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
          // 32-byte k-steps inside a 128-byte swizzle row, 4096 bytes between k-atoms
          const uint64_t bdesc = vdesc + (uint64_t)(((ks >> 2) * 4096 + (ks & 3) * 32) >> 4);
          // two independent accumulation chains (even / odd k-steps -> O_even / O_odd) halve the length of
          // the dependent-MMA chain of this small (N = 32) product; the epilogue adds the two halves
          const uint32_t acc = (uint32_t)((g | (ks >> 1)) != 0);
          if (PV16)   // P of column half hf sits in the first 24 columns of that half's S region
            tc_mma_bf16_ts(tmem_O + (ks & 1) * 32, tmem_S + (ks / 3) * AT_HALF + (ks % 3) * 8, bdesc, idesc_pv, acc);
          else
            tc_mma_tf32_ts(tmem_O + (ks & 1) * 32, tmem_S + 8 * ks, bdesc, idesc_pv, acc);
        }
        tc_commit(b_kvfree(st));                   // K/V stage reusable once the PV MMAs retire
        tc_commit(b_odone(g & 1));
        if (g + 2 < nkt) issue_qk(g + 2);          // refill this S buffer (in order after P·V(g))
      }
    }
  } else if (warp < 8) {
    // =============================== softmax ===============================
    const int qq = warp & 3, hf = warp >> 2;         // TMEM lane quarter / column half
    const int row = qq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qq * 32) << 16;
    const int c_lo = hf * AT_HALF;                   // first S column of this warp
    const float qn = xch[row];
    const int e2 = qt * AT_QT + row;
    const bool qok2 = e2 < S;
    float m_run = -INFINITY, l_run = 0.f;            // l_run: this warp's column half only
    for (int g = 0; g < nkt; ++g) {
      const int st = g % AT_STAGES, buf = g & 1;
      const uint32_t tmem_S = tmem_base + buf * AT_KT + c_lo;
      mbar_wait(b_sfull(buf), (uint32_t)((g >> 1) & 1));
      tc_fence_after();
      uint32_t va[16], vb[16];
      tc_ld16(tmem_S + lane_addr, va);                // first score chunk in flight while the tile flags are read
      const float* kmask = reinterpret_cast<const float*>(sgen + st_off + st * AT_STAGE_BYTES + AT_K_BYTES + VT);
      const int4 fl = *reinterpret_cast<const int4*>(kmask + 128);      // one flag per 32-key slice
      const bool masked = (fl.x | fl.y | fl.z) != 0;
      const float4 kn = *reinterpret_cast<const float4*>(kmask + 132);  // max |k| per slice
      const float bound = qn * fmaxf(kn.x, fmaxf(kn.y, kn.z));          // >= every score of this row & tile
      // exact two-pass only when this row has no reference yet or the bound could overflow exp2
      // (identical decision in both warps of a row pair: same m_run, same bound)
      const bool fresh = m_run == -INFINITY;
      const bool adopt = BREF && fresh && bound <= 40.f;      // same decision in both warps of a row
      const bool need_exact = __any_sync(0xffffffffu, (fresh && !adopt) || (!fresh && bound - m_run > 100.f));
      float m_new = (adopt && !need_exact) ? bound : m_run;
      if (need_exact) {
        // ---- pass 1: exact row max over this warp's columns, then exchange with the partner warp
        float tmax = -INFINITY;
        tc_ld_wait();
#pragma unroll
        for (int cc = 0; cc < AT_HALF / 16; ++cc) {
          uint32_t (&cur)[16] = (cc & 1) ? vb : va;
          uint32_t (&nxt)[16] = (cc & 1) ? va : vb;
          if (cc < AT_HALF / 16 - 1) tc_ld16(tmem_S + lane_addr + (cc + 1) * 16, nxt);
          if (masked) {
#pragma unroll
            for (int i = 0; i < 16; ++i) tmax = fmaxf(tmax, __uint_as_float(cur[i]) + kmask[c_lo + cc * 16 + i]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) tmax = fmaxf(tmax, __uint_as_float(cur[i]));
          }
          if (cc < AT_HALF / 16 - 1) tc_ld_wait();
        }
        tc_ld16(tmem_S + lane_addr, va);              // chunk 0 again for the exponentiation pass
        xch[128 * (1 + hf) + row] = tmax;
        named_bar_sync(3 + qq, 64);                  // the two warps of this lane quarter
        tmax = fmaxf(tmax, xch[128 * (2 - hf) + row]);
        named_bar_sync(3 + qq, 64);                  // partner has read before the slot is reused
        m_new = fmaxf(m_run, tmax);
      }
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2f(m_run - m_use);      // 1 when the reference is unchanged
      // ---- P = exp2(S - m), written back over S. The denominator sums the exact fp32 p; the
      // numerator operand is rounded to TF32 by adding half an ulp (the MMA truncates the low 13
      // mantissa bits), so its rounding error is zero-mean. PV16: P is packed to bf16 pairs (rn) and
      // lands in the first 24 columns of this warp's 48-column region (already consumed scores).
      float lsum = 0.f;
      tc_ld_wait();
#pragma unroll
      for (int cc = 0; cc < AT_HALF / 16; ++cc) {
        uint32_t (&cur)[16] = (cc & 1) ? vb : va;
        uint32_t (&nxt)[16] = (cc & 1) ? va : vb;
        if (cc < AT_HALF / 16 - 1) tc_ld16(tmem_S + lane_addr + (cc + 1) * 16, nxt);
        if (PV16) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x0 = __uint_as_float(cur[2 * i]), x1 = __uint_as_float(cur[2 * i + 1]);
            if (masked) { x0 += kmask[c_lo + cc * 16 + 2 * i]; x1 += kmask[c_lo + cc * 16 + 2 * i + 1]; }
            const float p0 = ex2f(x0 - m_use), p1 = ex2f(x1 - m_use);
            lsum += p0;
            lsum += p1;
            pk[i] = pack_bf16x2_rn(p0, p1);
          }
          if (cc < AT_HALF / 16 - 1) tc_ld_wait();
          tc_st8(tmem_S + lane_addr + cc * 8, pk);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x = __uint_as_float(cur[i]);
            if (masked) x += kmask[c_lo + cc * 16 + i];
            float pv = ex2f(x - m_use);
            lsum += pv;
            cur[i] = __float_as_uint(pv) + 0x1000u;
          }
          if (cc < AT_HALF / 16 - 1) tc_ld_wait();
          tc_st16(tmem_S + lane_addr + cc * 16, cur);
        }
      }
      // ---- the softmax reference of some row moved: the half-0 warp rescales the row's O accumulator
      // in TMEM (rare: the first tile has nothing to rescale, later tiles keep the reference unless
      // the overflow bound trips)
      if (hf == 0 && g > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        mbar_wait(b_odone((g - 1) & 1), (uint32_t)(((g - 1) >> 1) & 1));   // P·V of every earlier tile has retired
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 6; ++c) {          // 3 x 8 columns of O_even, then of O_odd
          const uint32_t col = (uint32_t)((c / 3) * 32 + (c % 3) * 8);
          uint32_t o[8];
          tc_ld8(tmem_O + lane_addr + col, o);
          tc_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tc_st8(tmem_O + lane_addr + col, o);
        }
      }
      tc_st_wait();
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_pready(buf));   // this warp's share of P (and rescaled O) is in TMEM
    }
    // ---- epilogue: half-1 hands its partial denominator to half-0, which writes O / l
    if (hf == 1) xch[128 * 3 + row] = l_run;
    named_bar_sync(3 + qq, 64);
    if (hf == 0) {
      const float l_tot = l_run + xch[128 * 3 + row];
      mbar_wait(b_odone((nkt - 1) & 1), (uint32_t)(((nkt - 1) >> 1) & 1));   // the last P·V (hence all) retired
      tc_fence_after();
      uint32_t o0[8], o1[8], o2[8], p0[8], p1[8], p2[8];
      tc_ld8(tmem_O + lane_addr + 0, o0);
      tc_ld8(tmem_O + lane_addr + 8, o1);
      tc_ld8(tmem_O + lane_addr + 16, o2);
      tc_ld8(tmem_O + lane_addr + 32, p0);
      tc_ld8(tmem_O + lane_addr + 40, p1);
      tc_ld8(tmem_O + lane_addr + 48, p2);
      tc_ld_wait();
      if (qok2) {
        const long long tq2 = seq_token(sm, s, e2);
        const float inv = 1.0f / l_tot;
        float acc[kHD];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i] = (__uint_as_float(o0[i]) + __uint_as_float(p0[i])) * inv;
          acc[8 + i] = (__uint_as_float(o1[i]) + __uint_as_float(p1[i])) * inv;
          acc[16 + i] = (__uint_as_float(o2[i]) + __uint_as_float(p2[i])) * inv;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i)
          store_operand4(p.out, (size_t)tq2 * kC + h * kHD + 4 * i,
                         make_float4(acc[4*i], acc[4*i+1], acc[4*i+2], acc[4*i+3]), p.round_out);
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)AT_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent variant (PV16 operands): 2 CTAs per SM, each loops over (sequence, head, query tile) work
// items with a static stride. The non-persistent kernel spends ~1/3 of every CTA's life in start-up
// (TMEM allocation, Q staging through dependent global loads, first ring fill, the two-pass first key
// tile) and tail (epilogue, barrier, dealloc), during which its 8 softmax warps issue no exponentials.
// Here every role streams across item boundaries:
//   warp 10      : Q stager - rotates / scales / TF32-rounds the NEXT item's 128 query rows into the other
//                  half of a double-buffered Q tile while the current item is being processed
//   warp 8 lane 0: TMA producer - the K/V ring is indexed by a running key-tile counter, so the first
//                  tiles of item n+1 are already in flight while item n finishes
//   warp 9 lane 0: MMA issuer - QK^T of tile c+2 is issued right after P·V of tile c, across items; the first
//                  P·V of an item (which overwrites O) waits until the previous item's epilogue has read O
//   warps 0-7    : softmax (as in attn_tc_kernel); the half-0 warps also run the per-item epilogue
// All mbarrier phases are derived from running counters (c = key tiles processed, i = items processed).
constexpr int ATP_STAGE_BYTES = 21 * 1024;      // >= at_img_bytes(true) = 21024, keeps K / V^T 1024-byte aligned
constexpr int ATP_SMEM_BYTES = 1024 /*align*/ + 2 * AT_Q_BYTES + AT_STAGES * ATP_STAGE_BYTES + 256 /*barriers*/ + 8 * 512 /*row exchange*/;
__host__ __device__ constexpr int atp_threads(int np) { return (4 * np + 3) * 32; }

// NP  = column parts of an S tile (softmax warps per TMEM lane quarter): 2 -> 8 softmax warps x 48 columns,
//       3 -> 12 softmax warps x 32 columns (more resident warps per scheduler to hide the per-tile latencies)
// DBG (timing experiments only, results undefined): 1 = no MUFU in the probability loop, 2 = P·V MMAs skipped,
// 3 = no TMEM traffic in the probability loop
template <int DBG, int NP>
__global__ void __launch_bounds__(atp_threads(NP), 2) attn_tcp_kernel(AttnParams p, const uint8_t* __restrict__ scratch,
                                                                     int total_items) {
  constexpr int IMG = at_img_bytes(true);
  constexpr int VT = at_vt_bytes(true);
  constexpr int NSW = 4 * NP;                    // softmax warps
  constexpr int COLS = AT_KT / NP;               // S columns per softmax warp
  constexpr int NCH = COLS / 16;                 // 16-column chunks per warp and tile
  extern __shared__ uint8_t smem_raw[];
  const SeqMap& sm = p.sm;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const uint32_t q_off = 0;                                    // two Q tiles
  const uint32_t st_off = 2 * AT_Q_BYTES;                      // stage s: [K | V^T | mask block]
  const uint32_t bar_off = 2 * AT_Q_BYTES + AT_STAGES * ATP_STAGE_BYTES;
  auto b_sfull = [&](int b) { return sbase + bar_off + 8 * b; };             // [2] MMA -> softmax
  auto b_kvfull = [&](int s) { return sbase + bar_off + 16 + 8 * s; };       // [3] TMA -> MMA
  auto b_kvfree = [&](int s) { return sbase + bar_off + 40 + 8 * s; };       // [3] MMA -> TMA
  auto b_odone = [&](int b) { return sbase + bar_off + (b ? 136 : 64); };    // [2] MMA -> softmax (O valid), by tile parity
  auto b_pready = [&](int b) { return sbase + bar_off + 72 + 8 * b; };       // [2] softmax -> MMA
  auto b_qfull = [&](int b) { return sbase + bar_off + 88 + 8 * b; };        // [2] stager -> MMA, softmax
  auto b_qfree = [&](int b) { return sbase + bar_off + 104 + 8 * b; };       // [2] MMA + softmax -> stager
  const uint32_t b_ofree = sbase + bar_off + 120;                            //     epilogue -> MMA (O consumed)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + bar_off + 128);
  // [2 + 2 NP][128]: qnorm of the two Q tiles | row max of each column part | denominator of each column part
  float* xch = reinterpret_cast<float*>(sgen + bar_off + 256);
  float* xmax = xch + 2 * 128;
  float* xl = xch + (2 + NP) * 128;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = sm.S, nkeys = S + 1;
  const int nqt = (S + AT_QT - 1) / AT_QT, nkt = (nkeys + AT_KT - 1) / AT_KT;
  const int bid = blockIdx.x, nblk = gridDim.x;
  const int n_my = (total_items - bid + nblk - 1) / nblk;       // work items of this CTA (grid <= total_items)

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(b_sfull(b), 1); mbar_init(b_pready(b), NSW);
      mbar_init(b_qfull(b), 1); mbar_init(b_qfree(b), NSW + 1);
      mbar_init(b_odone(b), 1);
    }
    mbar_init(b_ofree, 4);
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(b_kvfull(i), 1); mbar_init(b_kvfree(i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"((uint32_t)AT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 2 * AT_KT;

  if (warp == NSW + 2) {
    // =============================== Q stager ===============================
    for (int i = 0; i < n_my; ++i) {
      const int item = bid + i * nblk;
      const int qt = item % nqt;
      const long long sh = item / nqt;
      const int h = (int)(sh % kH);
      const long long s = sh / kH;
      const int qb = i & 1, u = i >> 1;
      if (u > 0) mbar_wait_backoff(b_qfree(qb), (uint32_t)((u & 1) ^ 1));   // QK^T of the previous user retired
      uint8_t* qdst = sgen + q_off + qb * AT_Q_BYTES;
#pragma unroll 1
      for (int rr = 0; rr < 4; ++rr) {
        const int r = rr * 32 + lane;
        const int e = qt * AT_QT + r;
        const int pe = e < S ? e : S - 1;
        const long long tq = seq_token(sm, s, pe);
        float q[kHD];
        load24(p.qkv, (size_t)tq * kQKV + h * kHD, p.qkv_fmt, q);
        rope24(q, p.cosT + pe * kHalf, p.sinT + pe * kHalf);
        float qn2 = 0.f;
#pragma unroll
        for (int k = 0; k < kHD; ++k) { q[k] = round_tf32(q[k] * 1.4426950408889634f); qn2 = fmaf(q[k], q[k], qn2); }
        xch[qb * 128 + r] = sqrtf(qn2) * 1.001f;
#pragma unroll
        for (int c = 0; c < 6; ++c)
          *reinterpret_cast<float4*>(qdst + sw128_off(r, c)) = make_float4(q[4*c], q[4*c+1], q[4*c+2], q[4*c+3]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_qfull(qb));
    }
  } else if (warp == NSW) {
    if (lane == 0) {
      // =============================== producer (TMA 1-D bulk copies) ===============================
      int c = 0;
      for (int i = 0; i < n_my; ++i) {
        const long long sh = (bid + i * nblk) / nqt;
        const uint8_t* img = scratch + (size_t)sh * nkt * IMG;
        for (int g = 0; g < nkt; ++g, ++c) {
          const int st = c % AT_STAGES, use = c / AT_STAGES;
          if (use > 0) mbar_wait_backoff(b_kvfree(st), (uint32_t)((use & 1) ^ 1));
          mbar_expect_tx(b_kvfull(st), IMG);
          const uint32_t dst = sbase + st_off + st * ATP_STAGE_BYTES;
          const uint8_t* src = img + (size_t)g * IMG;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
              ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"((uint32_t)IMG), "r"(b_kvfull(st))
              : "memory");
        }
      }
    }
  } else if (warp == NSW + 1) {
    if (lane == 0) {
      // =============================== MMA issuer (one thread) ===============================
      constexpr uint32_t idesc_qk = umma_idesc_tf32(AT_QT, AT_KT);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(AT_QT, 32);
      constexpr int NKS = AT_KT / 16;
      const int ttot = n_my * nkt;
      int qk_c = 0, qk_i = 0, qk_g = 0;              // cursor of the next QK^T product (tile, item, tile-in-item)
      auto issue_qk = [&]() {
        const int qb = qk_i & 1;
        if (qk_g == 0) mbar_wait(b_qfull(qb), (uint32_t)((qk_i >> 1) & 1));
        const int st = qk_c % AT_STAGES, use = qk_c / AT_STAGES;
        mbar_wait(b_kvfull(st), (uint32_t)(use & 1));
        tc_fence_after();
        const uint64_t adesc = umma_desc_k128(sbase + q_off + qb * AT_Q_BYTES);
        const uint64_t bdesc = umma_desc_k128(sbase + st_off + st * ATP_STAGE_BYTES);
        const uint32_t tS = tmem_base + (qk_c & 1) * AT_KT;
#pragma unroll
        for (int k = 0; k < 3; ++k)
          tc_mma_tf32(tS, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_qk, (uint32_t)(k != 0));
        tc_commit(b_sfull(qk_c & 1));
        if (qk_g == nkt - 1) tc_commit(b_qfree(qb));   // every QK^T of this item has been issued
        ++qk_c;
        if (++qk_g == nkt) { qk_g = 0; ++qk_i; }
      };
      issue_qk();
      if (ttot > 1) issue_qk();
      int i = 0, g = 0;
      for (int c = 0; c < ttot; ++c) {
        const int st = c % AT_STAGES, buf = c & 1;
        mbar_wait(b_pready(buf), (uint32_t)((c >> 1) & 1));   // every softmax warp wrote its share of P(c)
        if (g == 0 && i > 0) mbar_wait(b_ofree, (uint32_t)((i - 1) & 1));   // previous item's O has been read
        tc_fence_after();
        const uint32_t tmem_S = tmem_base + buf * AT_KT;
        const uint64_t vdesc = umma_desc_k128(sbase + st_off + st * ATP_STAGE_BYTES + AT_K_BYTES);
#pragma unroll
        for (int ks = 0; ks < (DBG == 2 ? 0 : NKS); ++ks) {
          const uint64_t bdesc = vdesc + (uint64_t)(((ks >> 2) * 4096 + (ks & 3) * 32) >> 4);
          // P of column part j sits (bf16 pairs) in the first COLS/2 columns of that part's S region
          const uint32_t a_addr = tmem_S + ((16 * ks) / COLS) * COLS + (((16 * ks) % COLS) / 16) * 8;
          tc_mma_bf16_ts(tmem_O + (ks & 1) * 32, a_addr, bdesc, idesc_pv, (uint32_t)((g | (ks >> 1)) != 0));
        }
        tc_commit(b_kvfree(st));
        tc_commit(b_odone(c & 1));
        if (c + 2 < ttot) issue_qk();
        if (++g == nkt) { g = 0; ++i; }
      }
    }
  } else if (warp < NSW) {
    // =============================== softmax ===============================
    const int qq = warp & 3, part = warp >> 2;       // TMEM lane quarter / column part
    const int row = qq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qq * 32) << 16;
    const int c_lo = part * COLS;
    int c = 0;
    for (int i = 0; i < n_my; ++i) {
      const int item = bid + i * nblk;
      const int qt = item % nqt;
      const long long sh = item / nqt;
      const int h = (int)(sh % kH);
      const long long s = sh / kH;
      const int qb = i & 1;
      mbar_wait(b_qfull(qb), (uint32_t)((i >> 1) & 1));
      const float qn = xch[qb * 128 + row];
      __syncwarp();
      if (lane == 0) mbar_arrive(b_qfree(qb));
      const int e2 = qt * AT_QT + row;
      const bool qok2 = e2 < S;
      float m_run = -INFINITY, l_run = 0.f;          // l_run: this warp's column part only
      for (int g = 0; g < nkt; ++g, ++c) {
        const int st = c % AT_STAGES, buf = c & 1;
        const uint32_t tmem_S = tmem_base + buf * AT_KT + c_lo;
        mbar_wait(b_sfull(buf), (uint32_t)((c >> 1) & 1));
        tc_fence_after();
        uint32_t va[16], vb[16];
        if (DBG != 3) tc_ld16(tmem_S + lane_addr, va);
        else {
#pragma unroll
          for (int k = 0; k < 16; ++k) { va[k] = __float_as_uint(-1.f - k); vb[k] = __float_as_uint(-2.f - k); }
        }
        const float* kmask = reinterpret_cast<const float*>(sgen + st_off + st * ATP_STAGE_BYTES + AT_K_BYTES + VT);
        const int4 fl = *reinterpret_cast<const int4*>(kmask + 128);
        const bool masked = (fl.x | fl.y | fl.z) != 0;
        const float4 kn = *reinterpret_cast<const float4*>(kmask + 132);
        const float bound = qn * fmaxf(kn.x, fmaxf(kn.y, kn.z));
        // exact two-pass only when this row has no reference yet or the bound could overflow exp2
        // (identical decision in all warps of a row: same m_run, same bound)
        const bool need_exact = __any_sync(0xffffffffu, (m_run == -INFINITY) || (bound - m_run > 100.f));
        float m_new = m_run;
        if (need_exact) {
          float tmax = -INFINITY;
          tc_ld_wait();
#pragma unroll
          for (int cc = 0; cc < NCH; ++cc) {
            uint32_t (&cur)[16] = (cc & 1) ? vb : va;
            uint32_t (&nxt)[16] = (cc & 1) ? va : vb;
            if (DBG != 3 && cc < NCH - 1) tc_ld16(tmem_S + lane_addr + (cc + 1) * 16, nxt);
            if (masked) {
#pragma unroll
              for (int k = 0; k < 16; ++k) tmax = fmaxf(tmax, __uint_as_float(cur[k]) + kmask[c_lo + cc * 16 + k]);
            } else {
#pragma unroll
              for (int k = 0; k < 16; ++k) tmax = fmaxf(tmax, __uint_as_float(cur[k]));
            }
            if (cc < NCH - 1) tc_ld_wait();
          }
          if (DBG != 3) tc_ld16(tmem_S + lane_addr, va);
          xmax[part * 128 + row] = tmax;
          named_bar_sync(3 + qq, 32 * NP);           // the NP warps of this lane quarter
#pragma unroll
          for (int j = 0; j < NP; ++j) tmax = fmaxf(tmax, xmax[j * 128 + row]);
          named_bar_sync(3 + qq, 32 * NP);           // partners have read before the slots are reused
          m_new = fmaxf(m_run, tmax);
        }
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = ex2f(m_run - m_use);
        float lsum = 0.f;
        tc_ld_wait();
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          uint32_t (&cur)[16] = (cc & 1) ? vb : va;
          uint32_t (&nxt)[16] = (cc & 1) ? va : vb;
          if (DBG != 3 && cc < NCH - 1) tc_ld16(tmem_S + lane_addr + (cc + 1) * 16, nxt);
          uint32_t pk[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float x0 = __uint_as_float(cur[2 * k]), x1 = __uint_as_float(cur[2 * k + 1]);
            if (masked) { x0 += kmask[c_lo + cc * 16 + 2 * k]; x1 += kmask[c_lo + cc * 16 + 2 * k + 1]; }
            const float p0 = DBG == 1 ? x0 - m_use : ex2f(x0 - m_use), p1 = DBG == 1 ? x1 - m_use : ex2f(x1 - m_use);
            lsum += p0;
            lsum += p1;
            pk[k] = pack_bf16x2_rn(p0, p1);
          }
          if (DBG != 3 && cc < NCH - 1) tc_ld_wait();
          if (DBG != 3) tc_st8(tmem_S + lane_addr + cc * 8, pk);
          else if (pk[0] == 0x12345678u && pk[7] == 0x9abcdef0u) xch[row] = 1.f;   // keep the math alive
        }
        if (part == 0 && g > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
          mbar_wait(b_odone((c - 1) & 1), (uint32_t)(((c - 1) >> 1) & 1));   // P·V of every earlier tile has retired
          tc_fence_after();
#pragma unroll
          for (int cb = 0; cb < 6; ++cb) {
            const uint32_t col = (uint32_t)((cb / 3) * 32 + (cb % 3) * 8);
            uint32_t o[8];
            tc_ld8(tmem_O + lane_addr + col, o);
            tc_ld_wait();
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tc_st8(tmem_O + lane_addr + col, o);
          }
        }
        tc_st_wait();
        l_run = l_run * alpha + lsum;
        m_run = m_new;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_pready(buf));
      }
      // ---- per-item epilogue: the other parts hand their partial denominators to part 0, which writes O / l
      if (part != 0) xl[part * 128 + row] = l_run;
      named_bar_sync(3 + qq, 32 * NP);
      if (part == 0) {
        float l_tot = l_run;
#pragma unroll
        for (int j = 1; j < NP; ++j) l_tot += xl[j * 128 + row];
        mbar_wait(b_odone((c - 1) & 1), (uint32_t)(((c - 1) >> 1) & 1));   // the item's last P·V (hence all) retired
        tc_fence_after();
        uint32_t o0[8], o1[8], o2[8], p0[8], p1[8], p2[8];
        tc_ld8(tmem_O + lane_addr + 0, o0);
        tc_ld8(tmem_O + lane_addr + 8, o1);
        tc_ld8(tmem_O + lane_addr + 16, o2);
        tc_ld8(tmem_O + lane_addr + 32, p0);
        tc_ld8(tmem_O + lane_addr + 40, p1);
        tc_ld8(tmem_O + lane_addr + 48, p2);
        tc_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_ofree);           // O is in registers: the next item may overwrite it
        if (qok2) {
          const long long tq2 = seq_token(sm, s, e2);
          const float inv = 1.0f / l_tot;
          float acc[kHD];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            acc[k] = (__uint_as_float(o0[k]) + __uint_as_float(p0[k])) * inv;
            acc[8 + k] = (__uint_as_float(o1[k]) + __uint_as_float(p1[k])) * inv;
            acc[16 + k] = (__uint_as_float(o2[k]) + __uint_as_float(p2[k])) * inv;
          }
#pragma unroll
          for (int k = 0; k < 6; ++k)
            store_operand4(p.out, (size_t)tq2 * kC + h * kHD + 4 * k,
                           make_float4(acc[4*k], acc[4*k+1], acc[4*k+2], acc[4*k+3]), p.round_out);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)AT_TMEM_COLS) : "memory");
  }
}

template <int DBG, int NP>
inline void attn_tcp_kernel_launch(const AttnParams& p, const uint8_t* scratch, int items, int grid, cudaStream_t s) {
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaFuncSetAttribute(attn_tcp_kernel<DBG, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATP_SMEM_BYTES);
    configured[dev] = true;
  }
  attn_tcp_kernel<DBG, NP><<<grid, atp_threads(NP), ATP_SMEM_BYTES, s>>>(p, scratch, items);
}

// mode: bits 0-1 = DBG, bit 2 = 12 softmax warps (NP = 3)
inline int attn_tcp_launch(const AttnParams& p, uint8_t* scratch, int prep2, int mode, cudaStream_t s, std::string* err) {
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device();
  const int num_sms = device_sm_count();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attn_prep2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AP2_SMEM_BYTES);
    if (e != cudaSuccess || num_sms <= 0) {
      if (err) *err = std::string("attn_tcp setup: ") + cudaGetErrorString(e);
      return -2;
    }
    configured[dev] = true;
  }
  const int nqt = (p.sm.S + AT_QT - 1) / AT_QT;
  const int nkt = (p.sm.S + 1 + AT_KT - 1) / AT_KT;
  const long long blocks = p.sm.num_seq * kH;
  const long long items = blocks * nqt;
  if (items > 0x7fffffffLL) { if (err) *err = "attn_tcp: too many work items"; return -2; }
  if (prep2 && p.qkv_fmt)
    attn_prep2_kernel<true><<<(unsigned)(p.sm.num_seq * nkt * 2), 256, AP2_SMEM_BYTES, s>>>(p, scratch);
  else
    attn_prep_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(p, scratch);
  const int grid = (int)(items < 2LL * num_sms ? items : 2LL * num_sms);
  switch (mode & 7) {
    case 1: attn_tcp_kernel_launch<1, 2>(p, scratch, (int)items, grid, s); break;
    case 2: attn_tcp_kernel_launch<2, 2>(p, scratch, (int)items, grid, s); break;
    case 3: attn_tcp_kernel_launch<3, 2>(p, scratch, (int)items, grid, s); break;
    case 4: attn_tcp_kernel_launch<0, 3>(p, scratch, (int)items, grid, s); break;
    case 5: attn_tcp_kernel_launch<1, 3>(p, scratch, (int)items, grid, s); break;
    default: attn_tcp_kernel_launch<0, 2>(p, scratch, (int)items, grid, s); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("attn_tcp launch: ") + cudaGetErrorString(e);
    return -2;
  }
  return 0;
}

// bytes of global scratch the kernel needs for this launch (zero-filled once at allocation; sized for
// the larger of the two image layouts so the variant can be switched on a live handle)
inline size_t attn_tc_scratch_bytes(const SeqMap& sm) {
  const int nkt = (sm.S + 1 + AT_KT - 1) / AT_KT;
  return (size_t)sm.num_seq * kH * nkt * at_img_bytes(false);
}

template <bool PV16, bool BREF>
inline int attn_tc_launch_t(const AttnParams& p, uint8_t* scratch, int prep2, bool prep_only, cudaStream_t s, std::string* err) {
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<PV16, BREF>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_prep2_kernel<PV16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AP2_SMEM_BYTES);
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(attn_tc): ") + cudaGetErrorString(e);
      return -2;
    }
    configured[dev] = true;
  }
  const int nqt = (p.sm.S + AT_QT - 1) / AT_QT;
  const int nkt = (p.sm.S + 1 + AT_KT - 1) / AT_KT;
  long long blocks = p.sm.num_seq * kH;
  if (prep2 && p.qkv_fmt)
    attn_prep2_kernel<PV16><<<(unsigned)(p.sm.num_seq * nkt * 2), 256, AP2_SMEM_BYTES, s>>>(p, scratch);
  else
    attn_prep_kernel<PV16><<<(unsigned)blocks, 256, 0, s>>>(p, scratch);
  if (!prep_only)
    attn_tc_kernel<PV16, BREF><<<(unsigned)(blocks * nqt), AT_THREADS, AT_SMEM_BYTES, s>>>(p, scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("attn_tc launch: ") + cudaGetErrorString(e);
    return -2;
  }
  return 0;
}

// `variant` (option attn_variant) selects the build variant of the tcgen05 attention:
//   bit 0 (1)  : bf16 P·V (PV16)                         bit 1 (2) : staged pre-pass (attn_prep2_kernel)
//   bit 2 (4)  : persistent kernel (implies PV16)        bit 3 (8) : ... with 12 instead of 8 softmax warps
//   bit 7 (128): bound-adopted first softmax reference (BREF, experiment pending hardware validation)
//   timing experiments, results undefined: bit 4 (16) pre-pass only; bits 5-6 DBG mode of the persistent kernel
// Default 3. Measured on B200 at B=64, T=1000, L=4 (profiles/r1_attention_ncu.md): 0 -> 2.20 ms per mha_t
// launch, 1 -> 2.05, 3 -> 1.95, 7 -> 1.96-2.05, 15 -> 2.08.
constexpr int kAttnVariantDefault = 256;   // generation 8 (attention_v8.cuh); 3 = best generation-7 variant
inline int attn_tc_launch(const AttnParams& p, uint8_t* scratch, int variant, cudaStream_t s, std::string* err) {
  const int p2 = (variant >> 1) & 1;
  const bool po = (variant & 16) != 0;
  if (variant & 4) return attn_tcp_launch(p, scratch, p2, ((variant >> 5) & 3) | ((variant & 8) ? 4 : 0), s, err);
  if (variant & 128)   // experiment: bound-adopted softmax reference (BREF)
    return (variant & 1) ? attn_tc_launch_t<true, true>(p, scratch, p2, po, s, err)
                         : attn_tc_launch_t<false, true>(p, scratch, p2, po, s, err);
  return (variant & 1) ? attn_tc_launch_t<true, false>(p, scratch, p2, po, s, err)
                       : attn_tc_launch_t<false, false>(p, scratch, p2, po, s, err);
}

}  // namespace mdgen
