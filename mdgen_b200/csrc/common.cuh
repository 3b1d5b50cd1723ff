// Common definitions for the mdgen_b200 device library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace mdgen {

// Fixed architecture (mdgen/parsing.py:87-93 defaults; every published checkpoint uses them).
constexpr int kC = 384;        // embed_dim
constexpr int kH = 16;         // mha_heads
constexpr int kHD = 24;        // head_dim
constexpr int kHalf = 12;      // rotary half
constexpr int kFF = 1536;      // ffn dim
constexpr int kQKV = 3 * kC;   // packed q|k|v projection width
constexpr int kIpaH = 4, kIpaC = 32, kIpaPq = 8, kIpaPv = 8;
constexpr int kIpaProjUsed = 128 + 256 + 96 + 192;  // q | kv | q_pts | kv_pts = 672
constexpr int kIpaProj = 768;  // row pitch of the fused projection (padded to a tensor-core N tile multiple)
constexpr int kIpaCat = kIpaH * (kIpaC + 4 * kIpaPv);  // 256
constexpr int kTFreq = 256;

// Selects the adaLN modulation row for a token: the table has one row per (step | sample).
//   sampling: row = *step_ptr (same t for the whole batch, mdgen/transport/integrators.py:98-101)
//   forward : row = b (per-sample t)
struct ModRef {
  const float* base;    // [rows, width]
  const int* step_ptr;  // device scalar or nullptr
  int width;            // floats per row
  int bstride;          // 0 (shared row) or 1 (row per sample)
  int tokens_per_b;     // tokens per sample in the tensor being processed
  int bmod;             // batch modulus (two-trunk IPA stacks 2B sequences over B samples)
};

__device__ __forceinline__ const float* mod_row(const ModRef& m, long long token) {
  int step = m.step_ptr ? *m.step_ptr : 0;
  int b = (int)(token / m.tokens_per_b) % m.bmod;
  return m.base + (size_t)(step + b * m.bstride) * m.width;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Round-to-nearest fp32 -> tf32 (10-bit mantissa), result kept in an fp32 container. Tensor-core
// kind::tf32 ignores the low 13 mantissa bits, so producers round once here (unbiased) instead
// of letting the MMA truncate.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));   // (emulated with ~4 integer ops on sm_100a)
  return __uint_as_float(u);
}
// One-instruction variant for activations that only feed a tensor-core MMA: add half a TF32 ulp to
// the magnitude; the MMA's truncation of the low 13 bits completes the round-to-nearest.
__device__ __forceinline__ float round_tf32_fast(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

// bf16 round-to-nearest-even kept in an fp32 container (precision experiments: mode 2 below)
__device__ __forceinline__ float round_bf16_rn(float x) {
  uint32_t u = __float_as_uint(x);
  u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
  return __uint_as_float(u);
}
// rounding applied by producers of GEMM operands: 0 none, 1 TF32 (tensor-core path), 2 bf16 (emulation)
__device__ __forceinline__ float round_operand(float x, int mode) {
  return mode == 1 ? round_tf32_fast(x) : (mode == 2 ? round_bf16_rn(x) : x);
}

__device__ __forceinline__ uint32_t pack_bf16x2_rn(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2_rn(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// packed pair of fp16 -> two floats (x = low half)
__device__ __forceinline__ float2 unpack_f16x2(uint32_t u) {
  float2 r;
  asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "r"(u));
  return r;
}
// 16-bit storage formats of GEMM-operand activations / weights
constexpr int kFmtF32 = 0, kFmtBF16 = 1, kFmtF16 = 2;
__device__ __forceinline__ float2 unpack_half2(uint32_t u, int fmt) {
  return fmt == kFmtF16 ? unpack_f16x2(u) : make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}
// Store 4 consecutive GEMM-operand values starting at element index `idx` of `base`:
//   mode 0/1/2: fp32 storage (unrounded / TF32-rounded / bf16-rounded-in-fp32)
//   mode 3    : true bf16 storage, 8 bytes       mode 4: true fp16 storage (the default tensor-core GEMM path:
//               TF32's 11-bit significand at bf16's size and MMA rate)
__device__ __forceinline__ void store_operand4(void* base, size_t idx, float4 v, int mode) {
  if (mode == 4) {
    uint2 pk;
    pk.x = pack_f16x2_rn(v.x, v.y);
    pk.y = pack_f16x2_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + idx) = pk;
  } else if (mode == 3) {
    uint2 pk;
    pk.x = pack_bf16x2_rn(v.x, v.y);
    pk.y = pack_bf16x2_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + idx) = pk;
  } else {
    if (mode) { v.x = round_operand(v.x, mode); v.y = round_operand(v.y, mode); v.z = round_operand(v.z, mode); v.w = round_operand(v.w, mode); }
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = v;
  }
}

// 16-byte asynchronous global -> shared copy (LDGSTS, L2 only)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ float gelu_erf(float x) {  // mdgen/model/layers.py:77-84
  return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

}  // namespace mdgen
