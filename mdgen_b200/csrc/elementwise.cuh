// Token-wise / small kernels of the MDGen denoiser: timestep embedding, adaLN table,
// LayerNorm+modulate, token embeddings, final layer + Euler update. All HBM-bound:
// one warp per token row (384 floats = 3 x float4 per lane), coalesced 128-bit accesses.
#pragma once
#include "common.cuh"

namespace mdgen {

// ---------------------------------------------------------------------------------------------
// Sinusoidal timestep features: emb[r] = [cos(tau f_i) | sin(tau f_i)], tau = t*time_multiplier,
// f_i = exp(-ln(1e4) i/128)   (mdgen/model/layers.py:30-50, mdgen/model/latent_model.py:243)
__global__ void sinus_kernel(const float* __restrict__ t, float mult, const float* __restrict__ freqs,
                             float* __restrict__ emb, int R) {
  int r = blockIdx.x;
  int i = threadIdx.x;  // 0..127
  if (r >= R) return;
  float a = (t[r] * mult) * freqs[i];
  emb[(size_t)r * kTFreq + i] = cosf(a);
  emb[(size_t)r * kTFreq + 128 + i] = sinf(a);
}

// out[r, j] = act( b[j] + sum_k W[j,k] * in[r,k] ) — one warp per output column j, the weight row
// lives in registers and is reused for every row r. Used for the t-embedder MLP
// (mdgen/model/layers.py:23-27) and the concatenated adaLN table (latent_model.py:346-349,405-408;
// layers.py:65-68). ACT: 0 = none, 1 = SiLU on the output.
template <int KPL, int ACT>
__global__ void rowdot_kernel(const float* __restrict__ W, const float* __restrict__ b,
                              const float* __restrict__ in, float* __restrict__ out, int J, int R) {
  constexpr int K = KPL * 32;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= J) return;
  float w[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) w[i] = W[(size_t)warp * K + i * 32 + lane];
  float bias = b[warp];
  for (int r = 0; r < R; ++r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) s = fmaf(w[i], in[(size_t)r * K + i * 32 + lane], s);
    s = warp_sum(s) + bias;
    if (ACT == 1) s = silu(s);
    if (lane == 0) out[(size_t)r * J + warp] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// y = LN0(x) * (1 + scale) + shift  — no-affine LayerNorm eps 1e-6 + adaLN modulate
// (mdgen/model/layers.py:14-15; latent_model.py:375,380,457,465,479). Optionally rounds the output
// to TF32 (it only feeds a tensor-core GEMM).
// y_add != nullptr: the residual add of the previous branch is fused in: x <- x + y_add (written back) before the
// LayerNorm (x = residual + gate * branch of latent_model.py:462,476,481; the GEMM that produced the branch stored
// gate * branch with EPI_GATE). This moves 3 KB per token of residual traffic out of the GEMM epilogues (which reach
// about half of the HBM copy rate) into this streaming kernel (which reaches ~90 % of it).
__global__ void ln_mod_kernel(float* __restrict__ x, const float* __restrict__ y_add, void* __restrict__ y, ModRef mod,
                              int shift_off, int scale_off, long long N, int rmode) {
  long long tok = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (tok >= N) return;
  float4* xr = reinterpret_cast<float4*>(x + (size_t)tok * kC);
  float4 v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = xr[i * 32 + lane];
  if (y_add) {
    const float4* yr = reinterpret_cast<const float4*>(y_add + (size_t)tok * kC);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float4 a = yr[i * 32 + lane];
      v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
      xr[i * 32 + lane] = v[i];
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  float mean = warp_sum(s) * (1.0f / kC);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kC) + 1e-6f);
  const float* mr = mod_row(mod, tok);
  const float4* sh = reinterpret_cast<const float4*>(mr + shift_off);
  const float4* sc = reinterpret_cast<const float4*>(mr + scale_off);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float4 a = sh[i * 32 + lane], b = sc[i * 32 + lane], o;
    o.x = (v[i].x - mean) * rstd * (1.0f + b.x) + a.x;
    o.y = (v[i].y - mean) * rstd * (1.0f + b.y) + a.y;
    o.z = (v[i].z - mean) * rstd * (1.0f + b.z) + a.z;
    o.w = (v[i].w - mean) * rstd * (1.0f + b.w) + a.w;
    store_operand4(y, (size_t)tok * kC + (size_t)(i * 32 + lane) * 4, o, rmode);
  }
}

// Affine LayerNorm eps 1e-5 (IPALayer.ipa_norm, mdgen/model/latent_model.py:351,372).
template <bool ROUND>
__global__ void ln_affine_kernel(const float* __restrict__ x, float* __restrict__ y,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 long long N) {
  long long tok = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (tok >= N) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)tok * kC);
  float4 v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = xr[i * 32 + lane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  float mean = warp_sum(s) * (1.0f / kC);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kC) + 1e-5f);
  const float4* g = reinterpret_cast<const float4*>(gamma);
  const float4* be = reinterpret_cast<const float4*>(beta);
  float4* yr = reinterpret_cast<float4*>(y + (size_t)tok * kC);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float4 a = g[i * 32 + lane], b = be[i * 32 + lane], o;
    o.x = (v[i].x - mean) * rstd * a.x + b.x;
    o.y = (v[i].y - mean) * rstd * a.y + b.y;
    o.z = (v[i].z - mean) * rstd * a.z + b.z;
    o.w = (v[i].w - mean) * rstd * a.w + b.w;
    if (ROUND) { o.x = round_tf32_fast(o.x); o.y = round_tf32_fast(o.y); o.z = round_tf32_fast(o.z); o.w = round_tf32_fast(o.w); }
    yr[i * 32 + lane] = o;
  }
}

// ---------------------------------------------------------------------------------------------
// Token embeddings (mdgen/model/latent_model.py:233-241). One thread per channel c keeps its
// weight row (D <= 28 floats) in registers; a block walks a strip of tokens whose D-vectors are
// staged in shared memory, so global reads/writes are coalesced along c.
//   MODE 0 (once per sampling call): cond[n,c] = b_lat[c] + pos[l,c] + Wc[c,:]·x_cond[n,:] + b_c[c]
//                                                + E_mask[x_cond_mask[n]][c]
//   MODE 1 (every step; builds without the tensor-core kernels only - the default path runs embed_step_kernel,
//           csrc/embed_step.cuh): h[n,c] = Wl[c,:]·x[n,:] + cond[n,c] + ipa[b,l,c]
constexpr int kEmbedTok = 128;  // tokens per block (amortises the per-thread weight-row load)
template <int MODE>
__global__ void __launch_bounds__(kC) embed_kernel(
    const float* __restrict__ xin, int D, const float* __restrict__ W /*[C,D]*/,
    const float* __restrict__ bias0, const float* __restrict__ bias1,
    const float* __restrict__ pos /*[crop,C] or null*/, const float* __restrict__ emask /*[2,C]*/,
    const int64_t* __restrict__ cmask, const float* __restrict__ cond /*[N,C]*/,
    const float* __restrict__ ipa /*[steps][B,L,C]*/, const int* __restrict__ step_ptr, long long ipa_step_stride,
    float* __restrict__ out, long long N, int T, int L) {
  __shared__ __align__(16) float xs[kEmbedTok][32];
  __shared__ int ms[kEmbedTok];
  int c = threadIdx.x;
  long long n0 = (long long)blockIdx.x * kEmbedTok;
  int nt = (int)min((long long)kEmbedTok, N - n0);
  for (int i = c; i < kEmbedTok * 32; i += kC) (&xs[0][0])[i] = 0.f;
  __syncthreads();
  for (int i = c; i < nt * D; i += kC) xs[i / D][i % D] = xin[(size_t)n0 * D + i];
  if (MODE == 0 && c < nt) ms[c] = (int)cmask[n0 + c];
  float w[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) w[k] = (k < D) ? W[(size_t)c * D + k] : 0.f;
  float bsum = (MODE == 0) ? (bias0[c] + bias1[c]) : 0.f;
  if (MODE == 1 && step_ptr) ipa += (size_t)(*step_ptr) * ipa_step_stride;   // trunk output of this Euler step
  __syncthreads();
  // sample index / residue index of the block's first token; per-token values follow incrementally
  // (a 64-bit division per token and thread used to dominate this kernel)
  const long long TL = (long long)T * L;
  long long b0 = n0 / TL;
  long long rem0 = n0 - b0 * TL;
  const int l0 = (int)(n0 % L);
  constexpr int U = 4;                 // tokens in flight per thread (independent global loads)
  for (int i = 0; i < nt; i += U) {
    float add[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      add[u] = 0.f;
      if (i + u < nt) {
        const long long n = n0 + i + u;
        const int l = (l0 + i + u) % L;
        if (MODE == 0) {
          add[u] = bsum + emask[(size_t)ms[i + u] * kC + c];
          if (pos) add[u] += pos[(size_t)l * kC + c];
        } else {
          long long b = b0, rem = rem0 + i + u;
          while (rem >= TL) { rem -= TL; ++b; }
          add[u] = cond[(size_t)n * kC + c] + ipa[((size_t)b * L + l) * kC + c];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u < nt) {
        float s = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < 7; ++k4) {     // 128-bit broadcast reads of the token's latent (w[k] = 0 for k >= D)
          const float4 xv = *reinterpret_cast<const float4*>(&xs[i + u][4 * k4]);
          s = fmaf(w[4 * k4], xv.x, s); s = fmaf(w[4 * k4 + 1], xv.y, s);
          s = fmaf(w[4 * k4 + 2], xv.z, s); s = fmaf(w[4 * k4 + 3], xv.w, s);
        }
        out[(size_t)(n0 + i + u) * kC + c] = s + add[u];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FinalLayer (mdgen/model/layers.py:70-74) fused with the Euler update
// (mdgen/transport/integrators.py:98-113):  v = Linear_{C->D}( LN0(h)*(1+scale)+shift )
//   EULER = true : x_out[n,:] = x_in[n,:] + dt[*step] * v      (state stays fp32)
//   EULER = false: x_out[n,:] = v
// One warp per token; W [D,C] (<= 43 KB) is staged in shared memory once per block.
// reduces p[j] over the 32 lanes for all j at once: afterwards lane j holds sum_lanes p[j] (31 shuffles
// instead of 32 x 5 for one warp_sum per output)
__device__ __forceinline__ float warp_transpose_reduce32(float (&p)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = hi ? p[k] : p[k + off];
      const float keep = hi ? p[k + off] : p[k];
      p[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return p[0];
}

template <bool EULER>
__global__ void __launch_bounds__(256) final_kernel(
    const float* __restrict__ h, ModRef mod, int shift_off, int scale_off,
    const float* __restrict__ W /*[D,C]*/, const float* __restrict__ bias, int D,
    const float* __restrict__ x_in, const float* __restrict__ dt, float* __restrict__ x_out,
    long long N) {
  extern __shared__ float ws[];  // [D][C]
  for (int i = threadIdx.x; i < D * kC; i += blockDim.x) ws[i] = W[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  float step_dt = 0.f;
  if (EULER) step_dt = dt[mod.step_ptr ? *mod.step_ptr : 0];
  const float my_bias = lane < D ? bias[lane] : 0.f;
  constexpr int TK = 2;   // tokens per warp pass: every weight row read from shared memory serves both
  for (long long tok0 = ((long long)blockIdx.x * nw + wid) * TK; tok0 < N; tok0 += (long long)gridDim.x * nw * TK) {
    float4 v[TK][3];
#pragma unroll
    for (int t = 0; t < TK; ++t) {
      const long long tok = tok0 + t < N ? tok0 + t : tok0;       // (odd tail: recompute the first token)
      const float4* xr = reinterpret_cast<const float4*>(h + (size_t)tok * kC);
#pragma unroll
      for (int i = 0; i < 3; ++i) v[t][i] = xr[i * 32 + lane];
    }
#pragma unroll
    for (int t = 0; t < TK; ++t) {
      const long long tok = tok0 + t < N ? tok0 + t : tok0;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) s += (v[t][i].x + v[t][i].y) + (v[t][i].z + v[t][i].w);
      const float mean = warp_sum(s) * (1.0f / kC);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float a = v[t][i].x - mean, b = v[t][i].y - mean, c = v[t][i].z - mean, d = v[t][i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kC) + 1e-6f);
      const float* mr = mod_row(mod, tok);
      const float4* sh = reinterpret_cast<const float4*>(mr + shift_off);
      const float4* sc = reinterpret_cast<const float4*>(mr + scale_off);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float4 a = sh[i * 32 + lane], b = sc[i * 32 + lane];
        v[t][i].x = (v[t][i].x - mean) * rstd * (1.0f + b.x) + a.x;
        v[t][i].y = (v[t][i].y - mean) * rstd * (1.0f + b.y) + a.y;
        v[t][i].z = (v[t][i].z - mean) * rstd * (1.0f + b.z) + a.z;
        v[t][i].w = (v[t][i].w - mean) * rstd * (1.0f + b.w) + a.w;
      }
    }
    float p[TK][32];
#pragma unroll
    for (int d = 0; d < 32; ++d) {
#pragma unroll
      for (int t = 0; t < TK; ++t) p[t][d] = 0.f;
      if (d < 28 && d < D) {
        const float4* wr = reinterpret_cast<const float4*>(ws + d * kC);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float4 w4 = wr[i * 32 + lane];
#pragma unroll
          for (int t = 0; t < TK; ++t) {
            p[t][d] = fmaf(v[t][i].x, w4.x, p[t][d]); p[t][d] = fmaf(v[t][i].y, w4.y, p[t][d]);
            p[t][d] = fmaf(v[t][i].z, w4.z, p[t][d]); p[t][d] = fmaf(v[t][i].w, w4.w, p[t][d]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < TK; ++t) {
      const float mine = warp_transpose_reduce32(p[t], lane);     // lane d holds output d
      if (lane < D && tok0 + t < N) {
        const float vout = mine + my_bias;
        const size_t o = (size_t)(tok0 + t) * D + lane;
        x_out[o] = EULER ? fmaf(step_dt, vout, x_in[o]) : vout;
      }
    }
  }
}

__global__ void step_advance_kernel(int* step) { *step += 1; }
__global__ void step_set_kernel(int* step, int v) { *step = v; }

// dst[r0 + r, c] = round?(scale * src[r, c]) — weight packing (q/k/v concat, head_dim^-0.5 fold).
__global__ void pack_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                 long long rows, int cols, int dst_ld, long long dst_row0, float scale,
                                 int do_round) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  long long r = i / cols;
  int c = (int)(i % cols);
  float v = src[i] * scale;
  if (do_round == 1) v = round_tf32(v);
  if (do_round == 2) v = round_bf16_rn(v);
  dst[(size_t)(dst_row0 + r) * dst_ld + c] = v;
}

// fp32 -> bf16 / fp16 (round-to-nearest-even) copy: 16-bit weight copies for the kind::f16 GEMM path
__global__ void to_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = (uint16_t)(pack_bf16x2_rn(src[i], 0.f) & 0xFFFFu);
}
__global__ void to_f16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = (uint16_t)(pack_f16x2_rn(src[i], 0.f) & 0xFFFFu);
}

// 16-bit (bf16 / fp16 per `fmt`) -> fp32 widening copy (debug entry point: 16-bit GEMM outputs handed back as fp32)
__global__ void from_half_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, long long n, int fmt) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = unpack_half2((uint32_t)src[i], fmt).x;
}

// ---------------------------------------------------------------------------------------------
// Training / validation loss pieces (mdgen/transport/transport.py:138-223)
// Interpolant plan (mdgen/transport/path.py:118-135): xt = alpha(t) x1 + sigma(t) x0, ut = alpha'(t) x1 + sigma'(t) x0
//   path 0 (GVP, path.py:173-191): alpha = sin(pi t / 2), sigma = cos(pi t / 2)      path 1 (Linear, :68-96 ICPlan): alpha = t, sigma = 1 - t
// t is per sample; `per` = elements per sample. 128-bit accesses when per % 4 == 0.
__global__ void flow_plan_kernel(const float* __restrict__ x1, const float* __restrict__ x0, const float* __restrict__ t,
                                 float* __restrict__ xt, float* __restrict__ ut, long long per, long long n, int path) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float tb = t[i / per];
  float a, s, da, ds;
  if (path == 0) {
    const float ang = tb * 1.57079632679489661923f;
    a = sinf(ang); s = cosf(ang);
    da = 1.57079632679489661923f * s; ds = -1.57079632679489661923f * a;
  } else {
    a = tb; s = 1.0f - tb; da = 1.0f; ds = -1.0f;
  }
  if ((per & 3) == 0 && i + 3 < n) {
    const float4 v1 = *reinterpret_cast<const float4*>(x1 + i), v0 = *reinterpret_cast<const float4*>(x0 + i);
    *reinterpret_cast<float4*>(xt + i) = make_float4(a * v1.x + s * v0.x, a * v1.y + s * v0.y, a * v1.z + s * v0.z, a * v1.w + s * v0.w);
    *reinterpret_cast<float4*>(ut + i) = make_float4(da * v1.x + ds * v0.x, da * v1.y + ds * v0.y, da * v1.z + ds * v0.z, da * v1.w + ds * v0.w);
  } else {
    for (long long j = i; j < i + 4 && j < n; ++j) {
      const float tj = t[j / per];
      float aj = a, sj = s, daj = da, dsj = ds;
      if (tj != tb) {   // element group straddles two samples
        if (path == 0) { const float g = tj * 1.57079632679489661923f; aj = sinf(g); sj = cosf(g); daj = 1.57079632679489661923f * sj; dsj = -1.57079632679489661923f * aj; }
        else { aj = tj; sj = 1.0f - tj; }
      }
      xt[j] = aj * x1[j] + sj * x0[j];
      ut[j] = daj * x1[j] + dsj * x0[j];
    }
  }
}

// mean_flat (transport.py:13-17) of the squared error (transport.py:189): one block per sample,
//   loss[b] = sum_i (pred - target)^2 mask / sum_i mask     (fixed summation order: deterministic)
__global__ void __launch_bounds__(1024) masked_mse_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                          const float* __restrict__ mask, float* __restrict__ loss, long long per) {
  __shared__ float snum[32], sden[32];
  const long long base = (long long)blockIdx.x * per;
  float num = 0.f, den = 0.f;
  for (long long i = threadIdx.x; i < per; i += blockDim.x) {
    const float d = pred[base + i] - target[base + i], m = mask[base + i];
    num = fmaf(d * d, m, num);
    den += m;
  }
  num = warp_sum(num); den = warp_sum(den);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { snum[w] = num; sden[w] = den; }
  __syncthreads();
  if (w == 0) {
    num = lane < (int)(blockDim.x >> 5) ? snum[lane] : 0.f;
    den = lane < (int)(blockDim.x >> 5) ? sden[lane] : 0.f;
    num = warp_sum(num); den = warp_sum(den);
    if (lane == 0) loss[blockIdx.x] = num / den;
  }
}

// Exponential moving average of the parameters (mdgen/ema.py:41-50): stored -= (stored - param) * (1 - decay)
__global__ void ema_update_kernel(float* __restrict__ stored, const float* __restrict__ param, long long n, float one_minus_decay) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float diff = stored[i] - param[i];
  diff *= one_minus_decay;
  stored[i] -= diff;
}

// ---------------------------------------------------------------------------------------------
// Runge-Kutta plumbing of the adaptive dopri5 sampler (mdgen_b200/ode.py; torchdiffeq's rk_common): stage / solution /
// error / interpolant combinations  out = (y ? y : 0) + scale * sum_i c_i k_i  over up to 8 tensors in ONE pass, and the
// mixed-tolerance RMS error ratio  sqrt(mean((err / (atol + rtol max(|y0|, |y1|)))^2))  as a fixed-order two-stage
// reduction (deterministic).
struct LinComb {
  const float* k[8];
  float c[8];
  int nk;
};
__global__ void lincomb_kernel(const float* __restrict__ y, float scale, LinComb lc, float* __restrict__ out, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < lc.nk) {
        const float4 v = *reinterpret_cast<const float4*>(lc.k[j] + i);
        acc.x = fmaf(lc.c[j], v.x, acc.x); acc.y = fmaf(lc.c[j], v.y, acc.y);
        acc.z = fmaf(lc.c[j], v.z, acc.z); acc.w = fmaf(lc.c[j], v.w, acc.w);
      }
    float4 b = y ? *reinterpret_cast<const float4*>(y + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(out + i) = make_float4(fmaf(scale, acc.x, b.x), fmaf(scale, acc.y, b.y),
                                                      fmaf(scale, acc.z, b.z), fmaf(scale, acc.w, b.w));
  } else {
    for (long long e = i; e < n; ++e) {
      float acc = 0.f;
      for (int j = 0; j < lc.nk; ++j) acc = fmaf(lc.c[j], lc.k[j][e], acc);
      out[e] = fmaf(scale, acc, y ? y[e] : 0.f);
    }
  }
}
constexpr int kErrBlocks = 512;
__global__ void __launch_bounds__(256) rk_error_partial_kernel(const float* __restrict__ err, const float* __restrict__ y0,
                                                               const float* __restrict__ y1, float rtol, float atol,
                                                               double* __restrict__ partial, long long n) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float tol = atol + rtol * fmaxf(fabsf(y0[i]), fabsf(y1[i]));
    const float r = err[i] / tol;
    acc += (double)r * (double)r;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void rk_error_final_kernel(const double* __restrict__ partial, int nblocks, long long n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += partial[b];
    out[0] = (float)sqrt(t / (double)n);
  }
}

// RoPE tables for positions 0..n-1: cos/sin(pos * inv_freq[i]), i < 12
// (fair-esm RotaryEmbedding; see oracle/ref_shims/esm/rotary_embedding.py).
__global__ void rope_table_kernel(const float* __restrict__ inv_freq, float* __restrict__ cosT,
                                  float* __restrict__ sinT, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * kHalf) return;
  int pos = i / kHalf, f = i % kHalf;
  float a = (float)pos * inv_freq[f];
  cosT[i] = cosf(a);
  sinT[i] = sinf(a);
}

}  // namespace mdgen
