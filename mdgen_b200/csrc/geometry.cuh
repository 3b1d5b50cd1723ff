// SE(3) featurisation / decode kernels around the sampler (HBM-bound, one thread per residue).
//   prep_kernel   == NewMDGenWrapper.prep_batch offsets/latents/cond (mdgen/wrapper.py:283-365)
//   decode_kernel == inference() tail + frames_torsions_to_atom14
//                    (mdgen/wrapper.py:456-478, mdgen/geometry.py:61-79,236-334)
#pragma once
#include "common.cuh"
#include "ipa.cuh"

namespace mdgen {

struct PrepFlags {
  int D;              // 21 | 28
  int two;            // tps / inpainting: second offset set relative to frame T-1
  int sim_condition, tps_condition, inpainting, cond_interval, no_torsion;
};

// Algorithmic bytes per residue: read 36 (rot) + 12 (trans) + 56 (torsions) [+48 ref frames, cached],
// write 2*4*D (latents, x_cond) + 8 (mask).
__global__ void prep_kernel(const float* __restrict__ rots, const float* __restrict__ trans,
                            const float* __restrict__ tors, float* __restrict__ latents,
                            float* __restrict__ x_cond, int64_t* __restrict__ cmask, int B, int T,
                            int L, PrepFlags f) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long N = (long long)B * T * L;
  if (n >= N) return;
  int l = (int)(n % L);
  int t = (int)((n / L) % T);
  long long b = n / ((long long)T * L);
  float R[9], tr[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = rots[(size_t)n * 9 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) tr[i] = trans[(size_t)n * 3 + i];
  float lat[28];
  {
    long long n0 = (b * T + 0) * L + l;
    float R0[9], t0[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R0[i] = rots[(size_t)n0 * 9 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t0[i] = trans[(size_t)n0 * 3 + i];
    relative_tensor7(R0, t0, R, tr, lat);                      // wrapper.py:307-309
  }
  int o = 7;
  if (f.two) {
    long long n1 = (b * T + (T - 1)) * L + l;
    float R1[9], t1[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R1[i] = rots[(size_t)n1 * 9 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t1[i] = trans[(size_t)n1 * 3 + i];
    relative_tensor7(R1, t1, R, tr, lat + 7);                  // wrapper.py:315-317
    o = 14;
  }
#pragma unroll
  for (int i = 0; i < 14; ++i) lat[o + i] = f.no_torsion ? 0.f : tors[(size_t)n * 14 + i];
  int cm = 0;                                                   // wrapper.py:338-346
  if (f.sim_condition && t == 0) cm = 1;
  if (f.tps_condition && (t == 0 || t == T - 1)) cm = 1;
  if (f.cond_interval && (t % f.cond_interval) == 0) cm = 1;
  if (f.inpainting && (l == 0 || l == 3)) cm = 1;
  cmask[n] = cm;
  for (int i = 0; i < f.D; ++i) {
    latents[(size_t)n * f.D + i] = lat[i];
    x_cond[(size_t)n * f.D + i] = cm ? lat[i] : 0.f;
  }
}

struct ResidueTables {
  const float* default_frame;   // [21,8,4,4]
  const float* group_pos;       // [21,14,3]
  const int* atom_group;        // [21,14]
  const float* atom_mask;       // [21,14]
};

__device__ __forceinline__ void rmul(const float* A, const float* B, float* Cc) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Cc[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void rvec(const float* A, const float* v, float* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}

// Algorithmic bytes per residue: read 4*D (sample) [+ frame-0 rigid and tables, cached],
// write 168 (14 atoms x 3 x fp32).
__global__ void decode_kernel(const float* __restrict__ samples, int D, int tors_off,
                              const float* __restrict__ srot, const float* __restrict__ strans,
                              const int64_t* __restrict__ seqres, ResidueTables tb,
                              float* __restrict__ atom14, int B, int T, int L) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long N = (long long)B * T * L;
  if (n >= N) return;
  int l = (int)(n % L);
  long long b = n / ((long long)T * L);
  const float* s = samples + (size_t)n * D;
  // Rigid.from_tensor_7(normalize_quats=True): q / |q| (no eps)   rigid_utils.py:324-325,1158
  float w = s[0], x = s[1], y = s[2], z = s[3];
  float qn = 1.0f / sqrtf(w * w + x * x + y * y + z * z);
  w *= qn; x *= qn; y *= qn; z *= qn;
  float Rq[9] = {w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
                 2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
                 2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z};
  const float* R0 = srot + ((size_t)b * L + l) * 9;
  const float* t0 = strans + ((size_t)b * L + l) * 3;
  float Rb[9], tb3[3];
  rmul(R0, Rq, Rb);                                            // rigids[:,0:1].compose(...)  :469
  rvec(R0, s + 4, tb3);
  tb3[0] += t0[0]; tb3[1] += t0[1]; tb3[2] += t0[2];
  int aa = (int)seqres[(size_t)b * L + l];
  // 8 rigid groups: default frame ∘ rot_x(torsion)  (geometry.py:273-334); groups 5..7 chained
  float gR[8][9], gt[8][3];
  float cR[9], ct[3];   // running chi chain (frame -> backbone)
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float* d44 = tb.default_frame + ((size_t)aa * 8 + g) * 16;
    float dR[9] = {d44[0], d44[1], d44[2], d44[4], d44[5], d44[6], d44[8], d44[9], d44[10]};
    float dt[3] = {d44[3], d44[7], d44[11]};
    float a0 = 0.f, a1 = 1.f;                                  // backbone group: (sin, cos) = (0, 1)
    if (g > 0) {
      float u = s[tors_off + (g - 1) * 2], v = s[tors_off + (g - 1) * 2 + 1];
      float nn = 1.0f / sqrtf(u * u + v * v);                  // wrapper.py:476
      a0 = u * nn; a1 = v * nn;
    }
    float rx[9] = {1.f, 0.f, 0.f, 0.f, a1, -a0, 0.f, a0, a1};
    float fR[9];
    rmul(dR, rx, fR);
    if (g <= 4) {
#pragma unroll
      for (int i = 0; i < 9; ++i) gR[g][i] = fR[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) gt[g][i] = dt[i];
      if (g == 4) {
#pragma unroll
        for (int i = 0; i < 9; ++i) cR[i] = fR[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) ct[i] = dt[i];
      }
    } else {
      float nR[9], nt[3];
      rmul(cR, fR, nR);
      rvec(cR, dt, nt);
#pragma unroll
      for (int i = 0; i < 3; ++i) nt[i] += ct[i];
#pragma unroll
      for (int i = 0; i < 9; ++i) { gR[g][i] = nR[i]; cR[i] = nR[i]; }
#pragma unroll
      for (int i = 0; i < 3; ++i) { gt[g][i] = nt[i]; ct[i] = nt[i]; }
    }
  }
  // to global: backbone ∘ group, then place the 14 atoms (geometry.py:236-270)
  float* outp = atom14 + (size_t)n * 42;
#pragma unroll 1
  for (int a = 0; a < 14; ++a) {
    int g = tb.atom_group[aa * 14 + a];
    const float* lp = tb.group_pos + ((size_t)aa * 14 + a) * 3;
    float p1[3], p2[3];
    rvec(gR[g], lp, p1);
    p1[0] += gt[g][0]; p1[1] += gt[g][1]; p1[2] += gt[g][2];
    rvec(Rb, p1, p2);
    float mk = tb.atom_mask[aa * 14 + a];
    outp[a * 3 + 0] = (p2[0] + tb3[0]) * mk;
    outp[a * 3 + 1] = (p2[1] + tb3[1]) * mk;
    outp[a * 3 + 2] = (p2[2] + tb3[2]) * mk;
  }
}

// ---------------------------------------------------------------------------------------------
// Rollout re-featurisation: what sim_inference.py:91-96 does on the host between
// rollouts — atom14_to_frames (mdgen/geometry.py:218-231) and atom37_to_torsions(atom14_to_atom37(.))
// (mdgen/geometry.py:9-27, 82-202) — as one device kernel, one thread per residue, so chained rollouts
// never leave the GPU. Algorithmic bytes per residue: 168 B in (+ previous residue's CA, C), 104 B out.
struct FeatTables {
  const int* chi_idx;       // [21,4,4] atom14 index of each chi atom
  const float* chi_amask;   // [21,4,4] atom exists (RESTYPE_ATOM37_MASK)
  const float* chi_mask;    // [21,4]   chi defined for this residue type
  const float* bb_mask;     // [21,4]   N, CA, C, O exist
};

// Rigid.from_3_points (mdgen/rigid_utils.py:1176-1216): Gram-Schmidt frame, columns e0|e1|e2.
// Written with explicit round-to-nearest intrinsics in the reference's exact operation order (no FMA
// contraction): several torsions of the reference are *degenerate* (first residue's pre-omega / phi
// use zero-padded atoms, undefined chis use four identical atoms), where the result is decided by
// rounding, so only an identical operation sequence reproduces the reference's values.
__device__ __forceinline__ float sumsq3(const float* v) {
  return __fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2]));
}
__device__ __forceinline__ float dot3(const float* a, const float* b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
}
__device__ __forceinline__ void frame_from_3_points(const float* pnx, const float* org, const float* pxy, float* R) {
  float e0[3] = {__fsub_rn(org[0], pnx[0]), __fsub_rn(org[1], pnx[1]), __fsub_rn(org[2], pnx[2])};
  float e1[3] = {__fsub_rn(pxy[0], org[0]), __fsub_rn(pxy[1], org[1]), __fsub_rn(pxy[2], org[2])};
  float d = __fsqrt_rn(__fadd_rn(sumsq3(e0), 1e-8f));
  e0[0] = __fdiv_rn(e0[0], d); e0[1] = __fdiv_rn(e0[1], d); e0[2] = __fdiv_rn(e0[2], d);
  const float dot = dot3(e0, e1);
  e1[0] = __fsub_rn(e1[0], __fmul_rn(e0[0], dot));
  e1[1] = __fsub_rn(e1[1], __fmul_rn(e0[1], dot));
  e1[2] = __fsub_rn(e1[2], __fmul_rn(e0[2], dot));
  d = __fsqrt_rn(__fadd_rn(sumsq3(e1), 1e-8f));
  e1[0] = __fdiv_rn(e1[0], d); e1[1] = __fdiv_rn(e1[1], d); e1[2] = __fdiv_rn(e1[2], d);
  const float e2[3] = {__fsub_rn(__fmul_rn(e0[1], e1[2]), __fmul_rn(e0[2], e1[1])),
                       __fsub_rn(__fmul_rn(e0[2], e1[0]), __fmul_rn(e0[0], e1[2])),
                       __fsub_rn(__fmul_rn(e0[0], e1[1]), __fmul_rn(e0[1], e1[0]))};
#pragma unroll
  for (int i = 0; i < 3; ++i) { R[i * 3 + 0] = e0[i]; R[i * 3 + 1] = e1[i]; R[i * 3 + 2] = e2[i]; }
}

// (sin, cos) of the torsion defined by 4 points a0..a3   (geometry.py:172-194):
// frame = from_3_points(a1, a2, a0); rel = frame.invert().apply(a3) = R^T a3 + (-(R^T a2))
// (rigid_utils.py:1047-1085); sin = rel.z, cos = rel.y, normalised with +1e-8 under the root.
__device__ __forceinline__ void torsion_sincos(const float* a0, const float* a1, const float* a2, const float* a3,
                                               float* out2) {
  float R[9];
  frame_from_3_points(a1, a2, a0, R);
  const float c1[3] = {R[1], R[4], R[7]}, c2[3] = {R[2], R[5], R[8]};     // rows 1, 2 of R^T
  const float ry = __fadd_rn(dot3(c1, a3), -dot3(c1, a2));
  const float rz = __fadd_rn(dot3(c2, a3), -dot3(c2, a2));
  const float den = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rz, rz), __fmul_rn(ry, ry)), 1e-8f));
  out2[0] = __fdiv_rn(rz, den);
  out2[1] = __fdiv_rn(ry, den);
}

__global__ void featurize_kernel(const float* __restrict__ atom14, const int64_t* __restrict__ seqres,
                                 FeatTables tb, float* __restrict__ rots, float* __restrict__ trans,
                                 float* __restrict__ tors, float* __restrict__ tmask, int B, int L) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= (long long)B * L) return;
  const int l = (int)(n % L);
  const int aa = (int)seqres[n];
  const float* a = atom14 + (size_t)n * 42;
  // frames: from_3_points(C, CA, N) composed with diag(-1, 1, -1); translation = CA
  {
    float R[9];
    frame_from_3_points(a + 6, a + 3, a + 0, R);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      rots[(size_t)n * 9 + i * 3 + 0] = -R[i * 3 + 0];
      rots[(size_t)n * 9 + i * 3 + 1] = R[i * 3 + 1];
      rots[(size_t)n * 9 + i * 3 + 2] = -R[i * 3 + 2];
      trans[(size_t)n * 3 + i] = a[3 + i];
    }
  }
  // backbone atoms as atom37 sees them (absent atoms zeroed), previous residue padded with zeros
  float bb[4][3], pv[3][3], bm[4], pm[3];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    bm[k] = tb.bb_mask[aa * 4 + k];
#pragma unroll
    for (int i = 0; i < 3; ++i) bb[k][i] = a[k * 3 + i] * bm[k];
  }
  if (l > 0) {
    const int aap = (int)seqres[n - 1];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      pm[k] = tb.bb_mask[aap * 4 + k];
#pragma unroll
      for (int i = 0; i < 3; ++i) pv[k][i] = a[-42 + k * 3 + i] * pm[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 3; ++k) { pm[k] = 0.f; pv[k][0] = pv[k][1] = pv[k][2] = 0.f; }
  }
  float sc[7][2], m[7];
  torsion_sincos(pv[1], pv[2], bb[0], bb[1], sc[0]);   // pre-omega: CA-, C-, N, CA
  torsion_sincos(pv[2], bb[0], bb[1], bb[2], sc[1]);   // phi:       C-, N, CA, C
  torsion_sincos(bb[0], bb[1], bb[2], bb[3], sc[2]);   // psi:       N, CA, C, O
  sc[2][0] = -sc[2][0]; sc[2][1] = -sc[2][1];          // geometry.py:196-201
  m[0] = pm[1] * pm[2] * bm[0] * bm[1];
  m[1] = pm[2] * bm[0] * bm[1] * bm[2];
  m[2] = bm[0] * bm[1] * bm[2] * bm[3];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float p[4][3];
    float am = 1.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = tb.chi_idx[(aa * 4 + c) * 4 + k];
      const float mk = tb.chi_amask[(aa * 4 + c) * 4 + k];
      am *= mk;
#pragma unroll
      for (int i = 0; i < 3; ++i) p[k][i] = a[idx * 3 + i] * mk;
    }
    torsion_sincos(p[0], p[1], p[2], p[3], sc[3 + c]);
    m[3 + c] = tb.chi_mask[aa * 4 + c] * am;
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    tors[(size_t)n * 14 + 2 * k] = sc[k][0];
    tors[(size_t)n * 14 + 2 * k + 1] = sc[k][1];
    if (tmask) tmask[(size_t)n * 7 + k] = m[k];
  }
}

}  // namespace mdgen
