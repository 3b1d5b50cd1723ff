// Per-step token embedding (mdgen/model/latent_model.py:233-246):
//     h[n, c] = W_lat[c, :] . x[n, :] + cond[n, c] + ipa[b(n), l(n), c]
// cond (the step-invariant conditioning embedding, built once per sampling call by embed_kernel<0>) is the only large
// operand: N x 384 fp32 read every step, N x 384 written. The first version of this kernel (embed_kernel<1>) issued
// 114 instructions per output (28 useful FMAs; the rest per-token index arithmetic: `% L`, 64-bit offsets) and kept only
// 4 dependent global loads in flight per thread: 0.54 ms per launch at 17 % of DRAM bandwidth (ncu, profiles/
// r2_side_kernels.md). Here: the block's 128 cond rows arrive as 16-row (24 KB) bulk copies into a double-buffered
// shared-memory stage (72 KB in flight per SM with 3 resident blocks, no registers, no address arithmetic), the trunk
// row offset of every token is computed once per block, and the thread keeps its weight row in registers as before.
#pragma once
#include "gemm_tc.cuh"   // mbarrier / shared-address helpers

namespace mdgen {

constexpr int kEsTok = 128;      // tokens per block
constexpr int kEsChunk = 16;     // cond rows per bulk copy
constexpr int kEsCondBytes = 2 * kEsChunk * kC * 4;
constexpr int kEsSmemBytes = kEsCondBytes + kEsTok * 32 * 4 + kEsTok * 4 + 16;

__global__ void __launch_bounds__(kC) embed_step_kernel(
    const float* __restrict__ xin, int D, const float* __restrict__ W /*[C,D]*/, const float* __restrict__ cond /*[N,C]*/,
    const float* __restrict__ ipa /*[steps][B,L,C]*/, const int* __restrict__ step_ptr, long long ipa_step_stride,
    float* __restrict__ out, long long N, int T, int L) {
  extern __shared__ __align__(128) uint8_t es_smem[];
  float* conds = reinterpret_cast<float*>(es_smem);                                   // [2][16][C]
  float (*xs)[32] = reinterpret_cast<float (*)[32]>(es_smem + kEsCondBytes);          // [128][32], zero padded
  int* ioff = reinterpret_cast<int*>(es_smem + kEsCondBytes + kEsTok * 32 * 4);       // trunk row offset per token
  const uint32_t bar0 = smem_u32(es_smem + kEsCondBytes + kEsTok * 32 * 4 + kEsTok * 4);
  const int c = threadIdx.x;
  const long long n0 = (long long)blockIdx.x * kEsTok;
  const int nt = (int)min((long long)kEsTok, N - n0);
  const int nchunks = (nt + kEsChunk - 1) / kEsChunk;
  auto issue = [&](int ch) {                                 // one thread: bulk copy of chunk `ch` into stage ch & 1
    const uint32_t bytes = (uint32_t)min(kEsChunk, nt - ch * kEsChunk) * kC * 4;
    const uint32_t bar = bar0 + 8u * (ch & 1);
    mbar_expect_tx(bar, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(conds + (ch & 1) * kEsChunk * kC)),
                   "l"(reinterpret_cast<uint64_t>(cond + (size_t)(n0 + ch * kEsChunk) * kC)), "r"(bytes), "r"(bar)
                 : "memory");
  };
  if (c == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    issue(0);
    if (nchunks > 1) issue(1);
  }
  for (int i = c; i < kEsTok * 32; i += kC) (&xs[0][0])[i] = 0.f;
  __syncthreads();
  for (int i = c; i < nt * D; i += kC) xs[i / D][i % D] = xin[(size_t)n0 * D + i];
  if (c < nt) {
    const long long n = n0 + c;
    const long long b = n / ((long long)T * L);
    ioff[c] = (int)((b * L + n % L) * kC);
  }
  float w[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) w[k] = (k < D) ? W[(size_t)c * D + k] : 0.f;
  const int nk4 = (D + 3) >> 2;
  if (step_ptr) ipa += (size_t)(*step_ptr) * ipa_step_stride;   // trunk output of this Euler step
  ipa += c;
  float* outp = out + (size_t)n0 * kC + c;
  __syncthreads();
  for (int ch = 0; ch < nchunks; ++ch) {
    mbar_wait(bar0 + 8u * (ch & 1), (uint32_t)((ch >> 1) & 1));
    const float* cs = conds + (ch & 1) * kEsChunk * kC + c;
    const int cnt = min(kEsChunk, nt - ch * kEsChunk);
#pragma unroll 4
    for (int u = 0; u < cnt; ++u) {
      const int i = ch * kEsChunk + u;
      float s = cs[u * kC] + ipa[ioff[i]];
#pragma unroll
      for (int k4 = 0; k4 < 7; ++k4) {       // 128-bit broadcast reads of the token's latent
        if (k4 < nk4) {
          const float4 xv = *reinterpret_cast<const float4*>(&xs[i][4 * k4]);
          s = fmaf(w[4 * k4], xv.x, s); s = fmaf(w[4 * k4 + 1], xv.y, s);
          s = fmaf(w[4 * k4 + 2], xv.z, s); s = fmaf(w[4 * k4 + 3], xv.w, s);
        }
      }
      outp[(size_t)i * kC] = s;
    }
    __syncthreads();                                          // every thread has read this stage
    if (c == 0 && ch + 2 < nchunks) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(ch + 2);
    }
  }
}

inline cudaError_t embed_step_configure() {
  static bool done[kMaxDevices] = {false};
  const int dev = current_device();
  if (done[dev]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(embed_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEsSmemBytes);
  if (e == cudaSuccess) done[dev] = true;
  return e;
}

}  // namespace mdgen
