// fp32 SIMT GEMM  out = epilogue(A[M,K] · W[N,K]^T + bias).
// This is the *validation / small-shape* GEMM of the library: exact fp32 FMA arithmetic, any
// M/N/K/ld. The production path for the large token GEMMs is gemm_tc.cuh (tcgen05 TF32); both
// share the epilogue definitions below so they can be A/B-checked on the device.
#pragma once
#include "common.cuh"

namespace mdgen {

enum EpiMode : int {
  EPI_STORE = 0,       // out = acc + bias
  EPI_GELU = 1,        // out = gelu(acc + bias)                   (fc1, layers.py:84)
  EPI_RESID_GATE = 2,  // out = resid + gate[b] * (acc + bias)     (latent_model.py:462,476,481)
  EPI_RESID = 3,       // out = resid + (acc + bias)               (x + ipa(...), latent_model.py:372)
  EPI_GATE = 4,        // out = gate[b] * (acc + bias): the gated branch output; the residual add is fused into the
                       // LayerNorm kernel that reads the residual stream next (ln_mod_kernel, y_add)
};

struct Epilogue {
  const float* bias;   // [N] or nullptr
  const float* resid;  // [M, ldo] (may alias out)
  ModRef mod;          // gate row source
  int gate_off;        // column offset of the gate chunk inside the mod row
  float* out;
  int ldo;
  int round_out;       // round result to TF32 (output only feeds another tensor-core GEMM)
  int half_fmt;        // tensor-core kernel with 16-bit operands / output: kFmtBF16 or kFmtF16
  int dbg;             // measurement switches of the tensor-core kernel (results are garbage): bit 0 = no TMA operand
                       // loads (MMA-issue-bound rate), bit 1 = no MMAs (operand-fill-bound rate)
};

template <int MODE>
__device__ __forceinline__ float apply_epilogue(const Epilogue& ep, float acc, long long m, int n) {
  float v = acc + (ep.bias ? ep.bias[n] : 0.f);
  if (MODE == EPI_GELU) v = gelu_erf(v);
  if (MODE == EPI_RESID_GATE) {
    const float* mr = mod_row(ep.mod, m);
    v = ep.resid[(size_t)m * ep.ldo + n] + mr[ep.gate_off + n] * v;
  }
  if (MODE == EPI_RESID) v = ep.resid[(size_t)m * ep.ldo + n] + v;
  if (MODE == EPI_GATE) v = mod_row(ep.mod, m)[ep.gate_off + n] * v;
  if (ep.round_out) v = round_operand(v, ep.round_out);
  return v;
}

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8;

template <int MODE>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, int lda,
                                                        const float* __restrict__ W, int ldw,
                                                        long long M, int N, int K, Epilogue ep) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  // loader mapping: each thread fetches 4 (row, k) elements of A and of W per k-tile
  const int lrow = tid >> 1;         // 0..127
  const int lk = (tid & 1) * 4;      // 0 or 4
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int k = k0 + lk + i;
      long long m = m0 + lrow;
      int n = n0 + lrow;
      ra[i] = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.f;
      rb[i] = (n < N && k < K) ? W[(size_t)n * ldw + k] : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[buf][lk + i][lrow] = ra[i];
      Bs[buf][lk + i][lrow] = rb[i];
    }
  };
  const int nk = (K + SG_BK - 1) / SG_BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * SG_BK);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n < N) ep.out[(size_t)m * ep.ldo + n] = apply_epilogue<MODE>(ep, acc[i][j], m, n);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// "Skinny" fp32 GEMM for the key-frame (IPA) trunk: M = B*L rows only (256 at BASELINE configs 2/5),
// so the kernel is latency- not throughput-bound. 32 x 64 tiles (many CTAs even for M = 256) and a
// 3-stage cp.async pipeline keep several K-tiles in flight per CTA. Exact fp32 FMA arithmetic: the
// trunk output is broadcast-added to every frame, so its rounding error is coherent across the whole
// trajectory (measured: TF32 here costs 40x more final-state error than TF32 in the token GEMMs).
// Requirements: N % 64 == 0, K % 32 == 0, lda/ldw % 4 == 0 (16-byte aligned rows).
constexpr int SK_BM = 32, SK_BN = 64, SK_BK = 32, SK_PITCH = 36, SK_STAGES = 3;
constexpr int SK_SMEM_BYTES = SK_STAGES * (SK_BM + SK_BN) * SK_PITCH * 4;

template <int MODE>
__global__ void __launch_bounds__(128) gemm_skinny_kernel(const float* __restrict__ A, int lda,
                                                          const float* __restrict__ W, int ldw,
                                                          long long M, int N, int K, Epilogue ep) {
  extern __shared__ __align__(16) float sk_smem[];
  float* As = sk_smem;                                       // [stage][32][36]
  float* Ws = sk_smem + SK_STAGES * SK_BM * SK_PITCH;        // [stage][64][36]
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * SK_BM;
  const int n0 = blockIdx.y * SK_BN;
  const int ty = tid >> 4, tx = tid & 15;
  const int nk = K / SK_BK;
  auto load_tile = [&](int kt, int stage) {
    const int k0 = kt * SK_BK;
#pragma unroll
    for (int i = 0; i < 2; ++i) {                            // A: 32 rows x 8 chunks
      int c = tid + i * 128, row = c >> 3, ch = c & 7;
      long long m = min(m0 + row, M - 1);
      cp_async16(As + (stage * SK_BM + row) * SK_PITCH + ch * 4, A + (size_t)m * lda + k0 + ch * 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {                            // W: 64 rows x 8 chunks
      int c = tid + i * 128, row = c >> 3, ch = c & 7;
      cp_async16(Ws + (stage * SK_BN + row) * SK_PITCH + ch * 4, W + (size_t)(n0 + row) * ldw + k0 + ch * 4);
    }
  };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int s = 0; s < SK_STAGES - 1; ++s) {
    if (s < nk) load_tile(s, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;" ::"n"(SK_STAGES - 2) : "memory");
    __syncthreads();
    if (kt + SK_STAGES - 1 < nk) load_tile(kt + SK_STAGES - 1, (kt + SK_STAGES - 1) % SK_STAGES);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float* as = As + (kt % SK_STAGES) * SK_BM * SK_PITCH;
    const float* ws = Ws + (kt % SK_STAGES) * SK_BN * SK_PITCH;
#pragma unroll
    for (int k4 = 0; k4 < SK_BK / 4; ++k4) {
      float4 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(as + (ty * 4 + i) * SK_PITCH + k4 * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(ws + (tx + 16 * j) * SK_PITCH + k4 * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx + 16 * j;
      ep.out[(size_t)m * ep.ldo + n] = apply_epilogue<MODE>(ep, acc[i][j], m, n);
    }
  }
}

}  // namespace mdgen
