// Fused multi-head attention with bias-KV token, RoPE and key-padding mask — the restated
// MultiheadAttention.forward of mdgen/model/mha.py:260-397 *without* materialising the
// [B', H, S, S+1] score tensor (K8-K13 in SURVEY.md §2c). fp32 SIMT version: online softmax,
// scores never leave registers. (The tensor-core variant lives in attention_tc.cuh.)
//
// Input is the packed projection qkv [N, 1152] = [q*24^-1/2 | k | v] (the scale is folded into
// W_q/b_q at pack time, mha.py:263). Head h owns channels [24h, 24h+24). RoPE (fair-esm) is applied
// at load time: position = index inside the sequence; the bias key sits at position S
// (mha.py:265-268 appends before :356-357 rotates).
//
// Sequence addressing (factorised attention, latent_model.py:458-461,472-475):
//   token(s, e) = (s / inner) * outer_stride + (s % inner) * inner_stride + e * elem_stride
//   mha_l: inner=1, outer_stride=L, elem_stride=1        (sequence = residues of one frame)
//   mha_t: inner=L, outer_stride=T*L, inner_stride=1, elem_stride=L  (frames of one residue)
#pragma once
#include "common.cuh"

namespace mdgen {

struct SeqMap {
  int S;              // sequence length (queries); keys = S + 1
  long long num_seq;
  int inner;
  long long outer_stride;
  int inner_stride;
  int elem_stride;
};
__device__ __forceinline__ long long seq_token(const SeqMap& sm, long long s, int e) {
  return (s / sm.inner) * sm.outer_stride + (s % sm.inner) * (long long)sm.inner_stride +
         (long long)e * sm.elem_stride;
}

struct AttnParams {
  const void* qkv;      // [N, 1152] fp32, bf16 or fp16 (written in that format by the QKV GEMM epilogue)
  int qkv_fmt;          // kFmtF32 / kFmtBF16 / kFmtF16
  const float* mask;    // [N] 1 = real token (key padding = 1 - mask), may be nullptr
  const float* bias_k;  // [384] raw (rotated at position S inside the kernel)
  const float* bias_v;  // [384]
  const float* cosT;    // [>= S+1, 12]
  const float* sinT;
  void* out;            // [N, 384] fp32, bf16 (round_out == 3) or fp16 (round_out == 4)
  int round_out;        // operand rounding / storage mode of the output (see store_operand4)
  SeqMap sm;
};

// 24 consecutive projection values (one head of q, k or v) starting at element `idx` of qkv
__device__ __forceinline__ void load24(const void* qkv, size_t idx, int fmt, float* out) {
  if (fmt != kFmtF32) {
    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(qkv) + idx);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint4 u = p[i];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_half2(w[j], fmt);
        out[8 * i + 2 * j] = f.x;
        out[8 * i + 2 * j + 1] = f.y;
      }
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(qkv) + idx);
#pragma unroll
    for (int i = 0; i < 6; ++i) { const float4 t = p[i]; out[4*i] = t.x; out[4*i+1] = t.y; out[4*i+2] = t.z; out[4*i+3] = t.w; }
  }
}

__device__ __forceinline__ void rope24(float* x, const float* c, const float* s) {
#pragma unroll
  for (int i = 0; i < kHalf; ++i) {
    float a = x[i], b = x[i + kHalf];
    x[i] = a * c[i] - b * s[i];          // x*cos + rotate_half(x)*sin, rotate_half = [-x2, x1]
    x[i + kHalf] = b * c[i] + a * s[i];
  }
}

// ---- short sequences (S <= 64): one thread per (sequence, query, head), keys read through L1.
__global__ void __launch_bounds__(128) attn_small_kernel(AttnParams p) {
  const SeqMap& sm = p.sm;
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = sm.num_seq * sm.S * kH;
  if (gid >= total) return;
  int h = (int)(gid % kH);
  long long r = gid / kH;
  int e = (int)(r % sm.S);
  long long s = r / sm.S;
  long long tq = seq_token(sm, s, e);
  float q[kHD], acc[kHD];
  load24(p.qkv, (size_t)tq * kQKV + h * kHD, p.qkv_fmt, q);
  rope24(q, p.cosT + e * kHalf, p.sinT + e * kHalf);
#pragma unroll
  for (int i = 0; i < kHD; ++i) acc[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j <= sm.S; ++j) {
    float k[kHD], v[kHD];
    if (j < sm.S) {
      long long tk = seq_token(sm, s, j);
      if (p.mask && p.mask[tk] == 0.f) continue;
      load24(p.qkv, (size_t)tk * kQKV + kC + h * kHD, p.qkv_fmt, k);
      load24(p.qkv, (size_t)tk * kQKV + 2 * kC + h * kHD, p.qkv_fmt, v);
    } else {
#pragma unroll
      for (int i = 0; i < kHD; ++i) { k[i] = p.bias_k[h * kHD + i]; v[i] = p.bias_v[h * kHD + i]; }
    }
    rope24(k, p.cosT + j * kHalf, p.sinT + j * kHalf);
    float sc = 0.f;
#pragma unroll
    for (int i = 0; i < kHD; ++i) sc = fmaf(q[i], k[i], sc);
    float mn = fmaxf(m, sc);
    float corr = __expf(m - mn);      // m = -inf on the first key -> 0
    float pj = __expf(sc - mn);
    l = l * corr + pj;
#pragma unroll
    for (int i = 0; i < kHD; ++i) acc[i] = fmaf(acc[i], corr, pj * v[i]);
    m = mn;
  }
  float inv = 1.0f / l;
#pragma unroll
  for (int i = 0; i < 6; ++i)
    store_operand4(p.out, (size_t)tq * kC + h * kHD + 4 * i,
                   make_float4(acc[4*i] * inv, acc[4*i+1] * inv, acc[4*i+2] * inv, acc[4*i+3] * inv), p.round_out);
}

// ---- S == 4 (tetrapeptide residue attention, mha_l at crop 4): every q/k/v row is loaded from HBM
// exactly once. Lane = (token-in-sequence tl, head-in-octet hh): a warp owns one sequence x 8 heads;
// each lane rotates its own q and k once and the 4x5 attention exchanges k/v through warp shuffles.
// (118 registers, 2 blocks per SM; capping at 80 registers for 3 blocks per SM spills 22 floats and
// measured 1.5 % faster only - not worth it.)
__global__ void __launch_bounds__(256) attn_l4_kernel(AttnParams p) {
  const SeqMap& sm = p.sm;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // warp = (sequence, head octet)
  const int lane = threadIdx.x & 31;
  if (w >= sm.num_seq * 2) return;
  const long long s = w >> 1;
  const int tl = lane >> 3, h = (int)(w & 1) * 8 + (lane & 7);
  const long long tok = seq_token(sm, s, tl);
  float q[kHD], k[kHD], v[kHD];
  load24(p.qkv, (size_t)tok * kQKV + h * kHD, p.qkv_fmt, q);
  load24(p.qkv, (size_t)tok * kQKV + kC + h * kHD, p.qkv_fmt, k);
  load24(p.qkv, (size_t)tok * kQKV + 2 * kC + h * kHD, p.qkv_fmt, v);
  rope24(q, p.cosT + tl * kHalf, p.sinT + tl * kHalf);
  rope24(k, p.cosT + tl * kHalf, p.sinT + tl * kHalf);
  const float valid = (p.mask == nullptr || p.mask[tok] != 0.f) ? 1.f : 0.f;
  // scores against the 4 sibling keys + the bias key (position 4)
  float sc[5];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < kHD; ++i) d = fmaf(q[i], __shfl_sync(0xffffffffu, k[i], j * 8 + (lane & 7)), d);
    float ok = __shfl_sync(0xffffffffu, valid, j * 8 + (lane & 7));
    sc[j] = ok != 0.f ? d : -INFINITY;
  }
  float bk[kHD];
#pragma unroll
  for (int i = 0; i < kHD; ++i) bk[i] = p.bias_k[h * kHD + i];
  rope24(bk, p.cosT + 4 * kHalf, p.sinT + 4 * kHalf);
  {
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < kHD; ++i) d = fmaf(q[i], bk[i], d);
    sc[4] = d;
  }
  float m = sc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) m = fmaxf(m, sc[j]);
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < 5; ++j) { sc[j] = __expf(sc[j] - m); l += sc[j]; }
  const float inv = 1.0f / l;
  float o[kHD];
#pragma unroll
  for (int i = 0; i < kHD; ++i) o[i] = sc[4] * p.bias_v[h * kHD + i];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < kHD; ++i) o[i] = fmaf(sc[j], __shfl_sync(0xffffffffu, v[i], j * 8 + (lane & 7)), o[i]);
#pragma unroll
  for (int i = 0; i < 6; ++i)
    store_operand4(p.out, (size_t)tok * kC + h * kHD + 4 * i,
                   make_float4(o[4*i] * inv, o[4*i+1] * inv, o[4*i+2] * inv, o[4*i+3] * inv), p.round_out);
}

// ---- S == 4, shared-memory exchange (default, option l4_variant = 1; measured on B200 at 256,000 tokens: 0.29 ms
// per launch against 0.37 ms for the shuffle kernel; a block-staged cp.async variant measured 0.34 ms and was dropped):
// same lane mapping and the same arithmetic order as attn_l4_kernel (results are bit-identical), but the
// rotated keys and the values of the four sibling tokens are exchanged through a warp-private shared-memory
// tile (48 multicast 128-bit reads per lane) instead of 192 warp shuffles. Head stride padded to 28 floats:
// the 8 heads of a multicast read fall into 8 distinct 4-bank groups.
constexpr int L4_HS = 28;
__global__ void __launch_bounds__(128) attn_l4s_kernel(AttnParams p) {
  __shared__ __align__(16) float ks[4][4][8][L4_HS];
  __shared__ __align__(16) float vs[4][4][8][L4_HS];
  const SeqMap& sm = p.sm;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * 4 + wib;        // warp = (sequence, head octet)
  if (w >= sm.num_seq * 2) return;                            // whole warps leave; no block-wide barrier below
  const long long s = w >> 1;
  const int tl = lane >> 3, hh = lane & 7, h = (int)(w & 1) * 8 + hh;
  const long long tok = seq_token(sm, s, tl);
  float q[kHD], t[kHD];
  load24(p.qkv, (size_t)tok * kQKV + h * kHD, p.qkv_fmt, q);
  load24(p.qkv, (size_t)tok * kQKV + kC + h * kHD, p.qkv_fmt, t);
  rope24(q, p.cosT + tl * kHalf, p.sinT + tl * kHalf);
  rope24(t, p.cosT + tl * kHalf, p.sinT + tl * kHalf);
#pragma unroll
  for (int i = 0; i < 6; ++i)
    *reinterpret_cast<float4*>(&ks[wib][tl][hh][4 * i]) = make_float4(t[4*i], t[4*i+1], t[4*i+2], t[4*i+3]);
  load24(p.qkv, (size_t)tok * kQKV + 2 * kC + h * kHD, p.qkv_fmt, t);
#pragma unroll
  for (int i = 0; i < 6; ++i)
    *reinterpret_cast<float4*>(&vs[wib][tl][hh][4 * i]) = make_float4(t[4*i], t[4*i+1], t[4*i+2], t[4*i+3]);
  const bool valid = p.mask == nullptr || p.mask[tok] != 0.f;
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);   // bit 8 j = token j of this sequence is real
  __syncwarp();
  float sc[5];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float4 k4 = *reinterpret_cast<const float4*>(&ks[wib][j][hh][4 * i]);
      d = fmaf(q[4*i], k4.x, d); d = fmaf(q[4*i+1], k4.y, d); d = fmaf(q[4*i+2], k4.z, d); d = fmaf(q[4*i+3], k4.w, d);
    }
    sc[j] = ((vmask >> (8 * j)) & 1u) ? d : -INFINITY;
  }
#pragma unroll
  for (int i = 0; i < kHD; ++i) t[i] = p.bias_k[h * kHD + i];
  rope24(t, p.cosT + 4 * kHalf, p.sinT + 4 * kHalf);
  {
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < kHD; ++i) d = fmaf(q[i], t[i], d);
    sc[4] = d;
  }
  float m = sc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) m = fmaxf(m, sc[j]);
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < 5; ++j) { sc[j] = __expf(sc[j] - m); l += sc[j]; }
  const float inv = 1.0f / l;
  float o[kHD];
#pragma unroll
  for (int i = 0; i < kHD; ++i) o[i] = sc[4] * p.bias_v[h * kHD + i];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float4 v4 = *reinterpret_cast<const float4*>(&vs[wib][j][hh][4 * i]);
      o[4*i] = fmaf(sc[j], v4.x, o[4*i]); o[4*i+1] = fmaf(sc[j], v4.y, o[4*i+1]);
      o[4*i+2] = fmaf(sc[j], v4.z, o[4*i+2]); o[4*i+3] = fmaf(sc[j], v4.w, o[4*i+3]);
    }
#pragma unroll
  for (int i = 0; i < 6; ++i)
    store_operand4(p.out, (size_t)tok * kC + h * kHD + 4 * i,
                   make_float4(o[4*i] * inv, o[4*i+1] * inv, o[4*i+2] * inv, o[4*i+3] * inv), p.round_out);
}

// ---- long sequences: block = (query tile of 256, head, sequence); 128 threads x 2 queries each;
// K/V tiles of 32 keys staged (and rotated) in shared memory and broadcast to all threads.
constexpr int AF_QT = 256;   // queries per block
constexpr int AF_KT = 32;    // keys per tile
__global__ void __launch_bounds__(128) attn_flash_simt_kernel(AttnParams p) {
  const SeqMap& sm = p.sm;
  __shared__ __align__(16) float Ks[AF_KT][kHD];
  __shared__ __align__(16) float Vs[AF_KT][kHD];
  __shared__ float valid[AF_KT];
  const int tid = threadIdx.x;
  const int h = blockIdx.y;
  const long long s = blockIdx.z;
  const int q0 = blockIdx.x * AF_QT;
  const float LOG2E = 1.4426950408889634f;

  float q[2][kHD], acc[2][kHD], m[2], l[2];
  long long tq[2];
  bool qok[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    int e = q0 + tid + u * 128;
    qok[u] = e < sm.S;
    int ee = qok[u] ? e : sm.S - 1;
    tq[u] = seq_token(sm, s, ee);
    load24(p.qkv, (size_t)tq[u] * kQKV + h * kHD, p.qkv_fmt, q[u]);
    rope24(q[u], p.cosT + ee * kHalf, p.sinT + ee * kHalf);
#pragma unroll
    for (int i = 0; i < kHD; ++i) { q[u][i] *= LOG2E; acc[u][i] = 0.f; }   // scores in log2 units
    m[u] = -INFINITY; l[u] = 0.f;
  }

  const int nkeys = sm.S + 1;
  for (int k0 = 0; k0 < nkeys; k0 += AF_KT) {
    __syncthreads();
    // stage + rotate K: 32 keys x 12 pairs = 384 items; V: 32 x 6 float4 = 192 items
    for (int it = tid; it < AF_KT * kHalf; it += 128) {
      int jj = it / kHalf, i = it % kHalf;
      int j = k0 + jj;
      float a = 0.f, b = 0.f;
      if (j < sm.S) {
        long long tk = seq_token(sm, s, j);
        const float* kp = reinterpret_cast<const float*>(p.qkv) + (size_t)tk * kQKV + kC + h * kHD;   // fp32 only
        a = kp[i]; b = kp[i + kHalf];
      } else if (j == sm.S) {
        a = p.bias_k[h * kHD + i]; b = p.bias_k[h * kHD + i + kHalf];
      }
      int jp = j <= sm.S ? j : 0;
      float c = p.cosT[jp * kHalf + i], sn = p.sinT[jp * kHalf + i];
      Ks[jj][i] = a * c - b * sn;
      Ks[jj][i + kHalf] = b * c + a * sn;
    }
    for (int it = tid; it < AF_KT * 6; it += 128) {
      int jj = it / 6, i = it % 6;
      int j = k0 + jj;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < sm.S) {
        long long tk = seq_token(sm, s, j);
        v = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.qkv) + (size_t)tk * kQKV + 2 * kC + h * kHD)[i];
      } else if (j == sm.S) {
        v = reinterpret_cast<const float4*>(p.bias_v + h * kHD)[i];
      }
      reinterpret_cast<float4*>(&Vs[jj][0])[i] = v;
    }
    if (tid < AF_KT) {
      int j = k0 + tid;
      float ok = 0.f;
      if (j < sm.S) ok = (p.mask == nullptr || p.mask[seq_token(sm, s, j)] != 0.f) ? 1.f : 0.f;
      else if (j == sm.S) ok = 1.f;
      valid[tid] = ok;
    }
    __syncthreads();
#pragma unroll
    for (int c0 = 0; c0 < AF_KT; c0 += 8) {
      float sc[2][8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float4* kr = reinterpret_cast<const float4*>(&Ks[c0 + jj][0]);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          float4 kk = kr[i];
          s0 = fmaf(q[0][4*i], kk.x, s0); s0 = fmaf(q[0][4*i+1], kk.y, s0);
          s0 = fmaf(q[0][4*i+2], kk.z, s0); s0 = fmaf(q[0][4*i+3], kk.w, s0);
          s1 = fmaf(q[1][4*i], kk.x, s1); s1 = fmaf(q[1][4*i+1], kk.y, s1);
          s1 = fmaf(q[1][4*i+2], kk.z, s1); s1 = fmaf(q[1][4*i+3], kk.w, s1);
        }
        bool ok = valid[c0 + jj] != 0.f;
        sc[0][jj] = ok ? s0 : -INFINITY;
        sc[1][jj] = ok ? s1 : -INFINITY;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float mx = m[u];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) mx = fmaxf(mx, sc[u][jj]);
        float ms = (mx == -INFINITY) ? 0.f : mx;
        float corr = exp2f(m[u] - ms);
        l[u] *= corr;
#pragma unroll
        for (int i = 0; i < kHD; ++i) acc[u][i] *= corr;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) sc[u][jj] = exp2f(sc[u][jj] - ms);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) l[u] += sc[u][jj];
        m[u] = mx;
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float4* vr = reinterpret_cast<const float4*>(&Vs[c0 + jj][0]);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          float4 vv = vr[i];
          acc[0][4*i] = fmaf(sc[0][jj], vv.x, acc[0][4*i]); acc[0][4*i+1] = fmaf(sc[0][jj], vv.y, acc[0][4*i+1]);
          acc[0][4*i+2] = fmaf(sc[0][jj], vv.z, acc[0][4*i+2]); acc[0][4*i+3] = fmaf(sc[0][jj], vv.w, acc[0][4*i+3]);
          acc[1][4*i] = fmaf(sc[1][jj], vv.x, acc[1][4*i]); acc[1][4*i+1] = fmaf(sc[1][jj], vv.y, acc[1][4*i+1]);
          acc[1][4*i+2] = fmaf(sc[1][jj], vv.z, acc[1][4*i+2]); acc[1][4*i+3] = fmaf(sc[1][jj], vv.w, acc[1][4*i+3]);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (!qok[u]) continue;
    float inv = 1.0f / l[u];
#pragma unroll
    for (int i = 0; i < 6; ++i)
      store_operand4(p.out, (size_t)tq[u] * kC + h * kHD + 4 * i,
                     make_float4(acc[u][4*i] * inv, acc[u][4*i+1] * inv, acc[u][4*i+2] * inv, acc[u][4*i+3] * inv),
                     p.round_out);
  }
}

}  // namespace mdgen
