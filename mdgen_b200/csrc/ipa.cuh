// Invariant Point Attention trunk kernels (key-frame only: B*L or 2*B*L rows).
// Restates InvariantPointAttention.forward with c_z = 0 (mdgen/model/ipa.py:113-255) and the
// SE(3) frame algebra it uses (Rigid.apply / invert_apply, mdgen/rigid_utils.py:1047-1073,64-86).
//
// The four projections q | kv | q_pts | kv_pts are one GEMM into proj[rows, 672]
// (weights concatenated at pack time); layouts inside a row follow the reference's views:
//   q      [0,128)    h*32 + c
//   kv     [128,384)  h*64 + (k: c | v: 32 + c)                          (ipa.py:117-123)
//   q_pts  [384,480)  coordinate-major: x*32 + (h*8 + p)                  (ipa.py:130-135)
//   kv_pts [480,672)  coordinate-major: x*64 + (h*16 + (k: p | v: 8+p))   (ipa.py:141-151)
#pragma once
#include "common.cuh"

namespace mdgen {

// Rigid.apply on every projected point, in place: p <- R_row p + t_row  (ipa.py:132,143).
__global__ void ipa_points_kernel(float* __restrict__ proj, const float* __restrict__ rot /*[rows,9]*/,
                                  const float* __restrict__ trans /*[rows,3]*/, long long rows) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * 96) return;
  long long r = gid / 96;
  int pt = (int)(gid % 96);
  float* base = proj + (size_t)r * kIpaProj;
  int ox, stride;
  if (pt < 32) { ox = 384 + pt; stride = 32; } else { ox = 480 + (pt - 32); stride = 64; }
  float x = base[ox], y = base[ox + stride], z = base[ox + 2 * stride];
  const float* R = rot + (size_t)r * 9;
  const float* t = trans + (size_t)r * 3;
  base[ox] = R[0] * x + R[1] * y + R[2] * z + t[0];
  base[ox + stride] = R[3] * x + R[4] * y + R[5] * z + t[1];
  base[ox + 2 * stride] = R[6] * x + R[7] * y + R[8] * z + t[2];
}

// One block per (sequence b', query residue i); warp h = head h.
//   logit(i,j,h) = q_i·k_j * sqrt(1/96) - 0.5 * softplus(w_h) * sqrt(1/108) * sum_p |T_i q_p - T_j k_p|^2
//                  + 1e5 * (m_i m_j - 1)                                         (ipa.py:161-198)
//   a = softmax_j; o = a v; o_pt = T_i^-1 (a v_pts); cat = [o | o_pt.x | o_pt.y | o_pt.z | |o_pt|]
//   (ipa.py:203-251). Dynamic shared memory: 4 * L floats of logits.
__global__ void __launch_bounds__(128) ipa_attn_kernel(
    const float* __restrict__ proj, const float* __restrict__ rot, const float* __restrict__ trans,
    const float* __restrict__ fmask /*[rows]*/, const float* __restrict__ head_w,
    float* __restrict__ cat, int L, int round_out) {
  extern __shared__ float lg[];  // [4][L]
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = blockIdx.x;
  const long long b = row / L;
  const float* pi = proj + (size_t)row * kIpaProj;
  float q[kIpaC];
#pragma unroll
  for (int c = 0; c < kIpaC; ++c) q[c] = pi[h * 32 + c];
  float qp[kIpaPq][3];
#pragma unroll
  for (int p = 0; p < kIpaPq; ++p)
#pragma unroll
    for (int x = 0; x < 3; ++x) qp[p][x] = pi[384 + x * 32 + h * 8 + p];
  const float s1 = 0.10206207261596577f;                 // sqrt(1/(3*32))
  float hw = head_w[h];
  float sp = (hw > 20.f) ? hw : log1pf(expf(hw));         // softplus (torch threshold 20)
  const float s2 = sp * 0.09622504486493763f * 0.5f;     // sqrt(1/(3*8*9/2)) * 0.5
  float mi = fmask[row];
  float* lgh = lg + h * L;
  float mx = -INFINITY;
  for (int j = lane; j < L; j += 32) {
    const float* pj = proj + (size_t)(b * L + j) * kIpaProj;
    const float4* kr = reinterpret_cast<const float4*>(pj + 128 + h * 64);
    float dot = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      float4 kk = kr[c4];
      dot = fmaf(q[4*c4], kk.x, dot); dot = fmaf(q[4*c4+1], kk.y, dot);
      dot = fmaf(q[4*c4+2], kk.z, dot); dot = fmaf(q[4*c4+3], kk.w, dot);
    }
    float d2 = 0.f;
#pragma unroll
    for (int p = 0; p < kIpaPq; ++p)
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        float d = qp[p][x] - pj[480 + x * 64 + h * 16 + p];
        d2 = fmaf(d, d, d2);
      }
    float mj = fmask[b * L + j];
    float v = dot * s1 - s2 * d2 + 1e5f * (mi * mj - 1.0f);
    lgh[j] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) {
    float e = expf(lgh[j] - mx);
    lgh[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float inv = 1.0f / sum;
  // o[c = lane] and o_pt[(x, p)] for lanes < 24 (x = lane / 8, p = lane % 8)
  float o = 0.f, op = 0.f;
  int px = lane >> 3, pp = lane & 7;
  for (int j = 0; j < L; ++j) {
    const float* pj = proj + (size_t)(b * L + j) * kIpaProj;
    float a = lgh[j];
    o = fmaf(a, pj[128 + h * 64 + 32 + lane], o);
    if (lane < 24) op = fmaf(a, pj[480 + px * 64 + h * 16 + 8 + pp], op);
  }
  o *= inv; op *= inv;
  // invert_apply: local = R_i^T (o_pt - t_i)   (rigid_utils.py:1061-1073)
  float gx = __shfl_sync(0xffffffffu, op, pp), gy = __shfl_sync(0xffffffffu, op, 8 + pp),
        gz = __shfl_sync(0xffffffffu, op, 16 + pp);
  const float* R = rot + (size_t)row * 9;
  const float* t = trans + (size_t)row * 3;
  float dx = gx - t[0], dy = gy - t[1], dz = gz - t[2];
  float lx = R[0] * dx + R[3] * dy + R[6] * dz;
  float ly = R[1] * dx + R[4] * dy + R[7] * dz;
  float lz = R[2] * dx + R[5] * dy + R[8] * dz;
  float nrm = sqrtf(lx * lx + ly * ly + lz * lz + 1e-8f);
  float* co = cat + (size_t)row * kIpaCat;
  auto rnd = [&](float v) { return round_out ? round_tf32_fast(v) : v; };
  co[h * 32 + lane] = rnd(o);
  if (lane < 8) {
    co[128 + h * 8 + lane] = rnd(lx);
    co[160 + h * 8 + lane] = rnd(ly);
    co[192 + h * 8 + lane] = rnd(lz);
    co[224 + h * 8 + lane] = rnd(nrm);
  }
}

// ---- quaternion helpers --------------------------------------------------------------------
// Closed-form rotation -> unit quaternion (w,x,y,z), canonical sign w >= 0. Replaces the
// batched 4x4 eigh of mdgen/rigid_utils.py:191-210 (K21 in SURVEY.md §2c): for a rotation matrix
// the top eigenvector of the K matrix *is* this quaternion up to sign.
__device__ __forceinline__ void rot_to_quat_dev(const float* R, float* q) {
  float xx = R[0], xy = R[1], xz = R[2], yx = R[3], yy = R[4], yz = R[5], zx = R[6], zy = R[7], zz = R[8];
  float tr = xx + yy + zz;
  float w, x, y, z;
  if (tr > 0.f) {
    float s = sqrtf(tr + 1.0f) * 2.0f;
    w = 0.25f * s; x = (zy - yz) / s; y = (xz - zx) / s; z = (yx - xy) / s;
  } else if (xx > yy && xx > zz) {
    float s = sqrtf(1.0f + xx - yy - zz) * 2.0f;
    w = (zy - yz) / s; x = 0.25f * s; y = (xy + yx) / s; z = (xz + zx) / s;
  } else if (yy > zz) {
    float s = sqrtf(1.0f + yy - xx - zz) * 2.0f;
    w = (xz - zx) / s; x = (xy + yx) / s; y = 0.25f * s; z = (yz + zy) / s;
  } else {
    float s = sqrtf(1.0f + zz - xx - yy) * 2.0f;
    w = (yx - xy) / s; x = (xz + zx) / s; y = (yz + zy) / s; z = 0.25f * s;
  }
  float n = 1.0f / sqrtf(w * w + x * x + y * y + z * z);
  if (w < 0.f) n = -n;
  q[0] = w * n; q[1] = x * n; q[2] = y * n; q[3] = z * n;
}

// o = a^-1 ∘ b as [quat | trans]: (Ra^T Rb, Ra^T (tb - ta))   (mdgen/utils.py:7-14)
__device__ __forceinline__ void relative_tensor7(const float* Ra, const float* ta, const float* Rb,
                                                 const float* tb, float* o7) {
  float Ro[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Ro[i * 3 + j] = Ra[0 * 3 + i] * Rb[0 * 3 + j] + Ra[1 * 3 + i] * Rb[1 * 3 + j] + Ra[2 * 3 + i] * Rb[2 * 3 + j];
  rot_to_quat_dev(Ro, o7);
  float dx = tb[0] - ta[0], dy = tb[1] - ta[1], dz = tb[2] - ta[2];
  o7[4] = Ra[0] * dx + Ra[3] * dy + Ra[6] * dz;
  o7[5] = Ra[1] * dx + Ra[4] * dy + Ra[7] * dz;
  o7[6] = Ra[2] * dx + Ra[5] * dy + Ra[8] * dz;
}

// Initial trunk activations (LatentMDGenModel.run_ipa, mdgen/model/latent_model.py:184-203).
//   single trunk: x[b,l,:] = aatype_emb[aatype]  (or 0)
//   two trunks  : rows [0,BL): x_r = W_r·tensor7(end^-1∘start) + b_r (+aa);   frames = start
//                 rows [BL,2BL): x_f = W_f·tensor7(start^-1∘end) + b_f (+aa); frames = end
// Also writes the per-row frame arrays (rot [rows,9], trans [rows,3]) used by the IPA kernels.
__global__ void __launch_bounds__(kC) ipa_init_kernel(
    int two, const float* __restrict__ srot, const float* __restrict__ strans,
    const float* __restrict__ erot, const float* __restrict__ etrans,
    const int64_t* __restrict__ aatype, const float* __restrict__ aa_emb /*[21,C] or null*/,
    const float* __restrict__ Wf, const float* __restrict__ bf, const float* __restrict__ Wr,
    const float* __restrict__ br, const float* __restrict__ mask /*[B,T,L]*/, int T, int L,
    float* __restrict__ x, float* __restrict__ frot, float* __restrict__ ftrans,
    float* __restrict__ fmask /*[rows] = mask[:,0]*/, long long BL,
    const float* __restrict__ quat_sign /*[2,BL] or null: eigh's eigenvector sign (latent_model.py:194-195)*/) {
  __shared__ float o7[7];
  // rows are stacked as [replica (Euler step)][trunk (1 or 2)][B*L]
  long long row = blockIdx.x;
  long long bl = row % BL;
  int which = two ? (int)((row / BL) & 1) : 0;   // 0: frames=start (x_r), 1: frames=end (x_f)
  int c = threadIdx.x;
  const float* Rme = which == 0 ? srot + bl * 9 : erot + bl * 9;
  const float* tme = which == 0 ? strans + bl * 3 : etrans + bl * 3;
  if (c < 9) frot[row * 9 + c] = Rme[c];
  if (c < 3) ftrans[row * 3 + c] = tme[c];
  if (c == 32) fmask[row] = mask[(size_t)(bl / L) * T * L + (bl % L)];
  float v = 0.f;
  if (two) {
    if (c == 0) {
      // which==0 -> x_r = end.invert().compose(start); which==1 -> x_f = start.invert().compose(end)
      const float* Ra = which == 0 ? erot + bl * 9 : srot + bl * 9;
      const float* ta = which == 0 ? etrans + bl * 3 : strans + bl * 3;
      relative_tensor7(Ra, ta, Rme, tme, o7);
      if (quat_sign && quat_sign[which * BL + bl] < 0.f) { o7[0] = -o7[0]; o7[1] = -o7[1]; o7[2] = -o7[2]; o7[3] = -o7[3]; }
    }
    __syncthreads();
    const float* W = which == 0 ? Wr : Wf;
    v = (which == 0 ? br : bf)[c];
#pragma unroll
    for (int k = 0; k < 7; ++k) v = fmaf(W[c * 7 + k], o7[k], v);
  }
  if (aa_emb) v += aa_emb[(size_t)aatype[bl] * kC + c];
  x[(size_t)row * kC + c] = v;
}

// ipa_out[rep][b,l,:] = x_r + x_f (two trunks) or a plain copy — latent_model.py:207.
// x rows are stacked [rep][trunk][B*L]; per = B*L*C elements of one trunk.
__global__ void ipa_sum_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, long long per,
                               int two) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long rep = i / per, rem = i - rep * per;
  const float* src = x + rep * per * (two ? 2 : 1) + rem;
  out[i] = two ? src[0] + src[per] : src[0];
}

}  // namespace mdgen
