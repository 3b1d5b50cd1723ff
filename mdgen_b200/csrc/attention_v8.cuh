// tcgen05 fused attention, generation 8 (default for sequences longer than 64): softmax(Q K^T + key padding) V
// with bias-KV token and RoPE (mdgen/model/mha.py:260-397), fp16 operands / fp32 accumulation throughout.
//
// What changed against attention_tc.cuh (generation 7) and why (profiles/r2_attention_v8.md):
//   * measured on B200 (tools/tmem_probe.cu): tcgen05.ld sustains 350-700 B/clk/SM, so the old kernel (45 B/clk/SM)
//     was never TMEM-bandwidth bound; it was bound by issue slots (6.8 instructions per exponential: two FADDs,
//     mbarrier spin loops, per-CTA start-up) and by the 7.6 GB of key-tile images each launch pulled out of L2.
//   * persistent CTAs, ONE per SM, each working on NQ = 4 query tiles (512 queries) of one (sequence, head) at
//     a time against one shared K/V ring: a key-tile image is fetched from L2 once per 512 queries, not per 128.
//   * one softmax thread per query row (FA4 style): no cross-warp exchange. Per 48-key tile a thread reads its
//     48 fp32 scores from TMEM, takes the row maximum (FMNMX3), keeps its softmax reference unless the maximum
//     grew by more than 2^8 (lazy rescale of the O row in TMEM), exponentiates (MUFU ex2) and writes P back over
//     the scores as fp16 pairs. Row sums come out of the P·V MMA itself (a row of ones in V^T), the key mask goes
//     into the QK^T MMA (a mask column: masked keys score -30000), so the inner loop is 1 FADD + 1 MUFU + 1/2 F2FP
//     + 1/2 FMNMX3 per score.
//   * S is double buffered in TMEM and every query tile has its own MMA-issuing thread: QK^T of tile c+2 is issued
//     right after P·V of tile c, so a softmax thread finds its next scores ready and no MMA / mbarrier round trip
//     (measured: 500-1100 cycles each) sits on its critical path.
//   * fp16 instead of TF32 / bf16 operands: Q, K (RoPE applied in fp32, then rounded once) and P, V carry TF32's
//     11-bit significand at half the bytes; QK^T is 2 kind::f16 MMAs (K = 32 = 24 + mask column + padding)
//     instead of 3 kind::tf32 ones, P·V 3 instead of 6.
//
// The pre-pass writes, per (sequence, head, key tile of 48 keys), a 7 KB image
//   K   [48 keys x 32 halfs]  K-major SWIZZLE_64B   cols 0..23 = RoPE(k), col 25 = 1 for masked / out-of-range keys
//   V^T [32 rows x 64 keys]   K-major SWIZZLE_128B  rows 0..23 = v^T (keys 0..47), row 24 = 1 (row sums)
// and, per 128 tokens of every (sequence, head), an 8 KB query-tile image
//   Q   [128 rows x 32 halfs] K-major SWIZZLE_64B   cols 0..23 = RoPE(q) log2(e), col 25 = -30000 (masked-key offset)
// Roles inside a CTA (NQ = 4: 21 warps; NQ = 2, used when the sequence has at most 256 queries: 11 warps, 2 CTAs/SM):
//   warps [0, 4 NQ)      softmax + epilogue: warp w owns rows 32 (w & 3) .. +31 of query tile w >> 2
//   1 warp               TMA producer (lane 0): cp.async.bulk of the item's query tiles into a double buffer and of
//                        one K/V image per key tile into an 8-stage ring
//   NQ warps             MMA issuers (lane 0 each), one per query tile: S = Q K^T, O += P V
// TMEM per query tile (128 columns): S0 [0, 48), S1 [48, 96) (P overwrites the first 24 columns of its buffer),
// O [96, 128) (N = 32: 24 head dims + row sum + padding).
#pragma once
#include "attention_tc.cuh"

namespace mdgen {

constexpr int A8_KT = 48;                        // keys per tile
constexpr int A8_K_BYTES = A8_KT * 64;           // 3072: K block, 64-byte rows
constexpr int A8_V_BYTES = 32 * 128;             // 4096: V^T block, one SWIZZLE_128B atom (keys 48..63 unused)
constexpr int A8_IMG = A8_K_BYTES + A8_V_BYTES;  // 7168
constexpr int A8_STAGES = 8;
constexpr int A8_QT_BYTES = 128 * 64;            // one staged query tile (128 rows x 64 B)
constexpr float A8_MASKED = -30000.0f;           // score offset of a masked key (exp2 -> exactly 0)
constexpr int A8_TSTRIDE = 128, A8_S1 = A8_KT, A8_OCOL = 96;   // TMEM columns: per query tile, S buffer 1, O
constexpr float A8_LAZY = 8.0f;
#ifndef A8_POLY_EVERY
#define A8_POLY_EVERY 3                           // one score pair in 3 is exponentiated on the FMA pipe (measured best of 2, 3, 4)
#endif                  // the softmax reference moves only when the row maximum grows by > 2^8

__host__ __device__ constexpr int a8_threads(int nq) { return (5 * nq + 1) * 32; }
__host__ __device__ constexpr int a8_smem_bytes(int nq) {
  return 1024 /*align*/ + 2 * nq * A8_QT_BYTES + A8_STAGES * A8_IMG + 512 /*barriers*/;
}
__host__ __device__ inline int a8_nkt(int S) { return (S + 1 + A8_KT - 1) / A8_KT; }
__host__ __device__ inline int a8_nqt_pad(int S) { return ((S + 511) / 512) * 4; }   // query tiles, padded to whole items

// byte offset of 16-byte chunk `c` (0..3) of row `r` inside a [rows x 64 B] K-major SWIZZLE_64B tile
// (Swizzle<2,4,3>: address bits [4,6) ^= bits [7,9), 512-byte atoms of 8 rows)
__device__ __forceinline__ uint32_t sw64_off(int r, int c) {
  return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_64B: stride between 8-row groups = 512 B, layout type 4
__device__ __forceinline__ uint64_t umma_desc_k64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// 2^x for two scores on the FMA pipe instead of the MUFU pipe (which bounds this kernel: ncu XU pipe 77 %), packed to
// an fp16 pair. Cody-Waite split with the 1.5 * 2^23 magic number (t = x + magic holds round(x) in its low mantissa
// bits), degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, below fp16's half ulp of
// 2.4e-4), exponent re-inserted with an integer shift-add. The three adds and three FMAs are Blackwell's packed
// fp32x2 instructions (add.f32x2 / fma.f32x2: one issue slot for both values). Valid for x <= 127; x is clamped at
// -125 (a masked key's -30000 becomes 2^-125, which rounds to 0 in fp16).
__device__ __forceinline__ uint32_t exp2_poly_pack_f16x2(float x0, float x1) {
  constexpr uint64_t kMagic = 0x4B4000004B400000ull;      // {12582912.f, 12582912.f}
  constexpr uint64_t kNegMagic = 0xCB400000CB400000ull;
  constexpr uint64_t kNegOne = 0xBF800000BF800000ull;
  constexpr uint64_t kC0 = 0x3F7FFB493F7FFB49ull;         // 0.99992806
  constexpr uint64_t kC1 = 0x3F31798D3F31798Dull;         // 0.69326097
  constexpr uint64_t kC2 = 0x3E786F0D3E786F0Dull;         // 0.24261113
  constexpr uint64_t kC3 = 0x3D61FBB03D61FBB0ull;         // 0.05517167
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  uint64_t X, T, N, F, P;
  asm("mov.b64 %0, {%1, %2};" : "=l"(X) : "f"(x0), "f"(x1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(T) : "l"(X), "l"(kMagic));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(N) : "l"(T), "l"(kNegMagic));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(F) : "l"(N), "l"(kNegOne), "l"(X));      // f = x - round(x)
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(P) : "l"(F), "l"(kC3), "l"(kC2));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(P) : "l"(P), "l"(F), "l"(kC1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(P) : "l"(P), "l"(F), "l"(kC0));
  uint32_t p0, p1, t0, t1;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(P));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(t0), "=r"(t1) : "l"(T));
  return pack_f16x2_rn(__uint_as_float(p0 + (t0 << 23)), __uint_as_float(p1 + (t1 << 23)));
}

// ---------------------------------------------------------------------------------------------------------
// Pre-pass: one block per (sequence, 48-token tile, head octet), 384 threads = one (head, token) pair each. The
// q|k|v row segments of the octet (3 x 384 contiguous bytes per token when q|k|v is 16-bit) are staged through
// shared memory with 16-byte cp.async; every thread then builds the K row and V^T column of its token as a key,
// and its row of the Q image as a query (RoPE, x log2 e, fp16; col 25 = the masked-key score offset).
// Scratch layout: [num_seq * 16 * nkt key-tile images (7 KB)] [num_seq * 16 * nqt_pad query-tile images (8 KB)].
constexpr int A8P_PITCH = 1168;                  // staged row: 3 x 384 B + 16 (bank spread)
constexpr int A8P_SMEM = A8_KT * A8P_PITCH;
constexpr int A8P_THREADS = 8 * A8_KT;           // 384
__global__ void __launch_bounds__(A8P_THREADS, 2) attn8_prep_kernel(AttnParams p, uint8_t* __restrict__ scratch) {
  extern __shared__ __align__(16) uint8_t a8p_smem[];
  const SeqMap& sm = p.sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = sm.S;
  const int nkt = a8_nkt(S), nqt = a8_nqt_pad(S);
  const int hg = blockIdx.x & 1;
  const int kt = (int)((blockIdx.x >> 1) % nkt);
  const long long s = (blockIdx.x >> 1) / nkt;
  uint8_t* qimg0 = scratch + (size_t)sm.num_seq * kH * nkt * A8_IMG;
  const bool half_in = p.qkv_fmt != kFmtF32;
  // token of key / query j of this sequence: seq_token is affine in j (one 64-bit division per block, not per row)
  const long long tok_base = seq_token(sm, s, 0);
  const long long tok_step = sm.elem_stride;
  if (half_in) {
    const uint16_t* qkv = reinterpret_cast<const uint16_t*>(p.qkv) + hg * 192;
    // 72 chunks of 16 bytes per row: q octet (24) | k octet (24) | v octet (24); a lane owns chunks lane, lane + 32, lane + 64
    int soff[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int ch = lane + 32 * u;
      soff[u] = (ch / 24) * kC + (ch % 24) * 8;
    }
    for (int row = warp; row < A8_KT; row += A8P_THREADS / 32) {
      const int j = kt * A8_KT + row;
      if (j < S) {
        const uint16_t* src = qkv + (size_t)(tok_base + j * tok_step) * kQKV;
        uint8_t* dst = a8p_smem + row * A8P_PITCH + lane * 16;
        cp_async16(dst, src + soff[0]);
        cp_async16(dst + 512, src + soff[1]);
        if (lane < 8) cp_async16(dst + 1024, src + soff[2]);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  const int hl = tid / A8_KT, r = tid - hl * A8_KT;
  const int h = hg * 8 + hl;
  const int j = kt * A8_KT + r;
  uint8_t* kbase = scratch + ((size_t)(s * kH + h) * nkt + kt) * A8_IMG;
  uint8_t* vbase = kbase + A8_K_BYTES;
  float k[kHD], v[kHD];
  float masked = 0.f;
  if (j < S) {
    const long long tk = tok_base + j * tok_step;
    float q[kHD];
    if (half_in) {
      const uint8_t* rowp = a8p_smem + r * A8P_PITCH + hl * 48;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const uint4 uq = *reinterpret_cast<const uint4*>(rowp + i * 16);
        const uint4 uk = *reinterpret_cast<const uint4*>(rowp + 384 + i * 16);
        const uint4 uv = *reinterpret_cast<const uint4*>(rowp + 768 + i * 16);
        const uint32_t wq[4] = {uq.x, uq.y, uq.z, uq.w}, wk[4] = {uk.x, uk.y, uk.z, uk.w}, wv[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float2 fq = unpack_half2(wq[c], p.qkv_fmt), fk = unpack_half2(wk[c], p.qkv_fmt), fv = unpack_half2(wv[c], p.qkv_fmt);
          q[8 * i + 2 * c] = fq.x; q[8 * i + 2 * c + 1] = fq.y;
          k[8 * i + 2 * c] = fk.x; k[8 * i + 2 * c + 1] = fk.y;
          v[8 * i + 2 * c] = fv.x; v[8 * i + 2 * c + 1] = fv.y;
        }
      }
    } else {
      load24(p.qkv, (size_t)tk * kQKV + h * kHD, kFmtF32, q);
      load24(p.qkv, (size_t)tk * kQKV + kC + h * kHD, kFmtF32, k);
      load24(p.qkv, (size_t)tk * kQKV + 2 * kC + h * kHD, kFmtF32, v);
    }
    if (p.mask && p.mask[tk] == 0.f) masked = 1.f;
    // ---- this token as a query: row (j & 127) of query tile (j >> 7)
    rope24(q, p.cosT + j * kHalf, p.sinT + j * kHalf);
    uint8_t* qdst = qimg0 + ((size_t)(s * kH + h) * nqt + (j >> 7)) * A8_QT_BYTES;
    const int qr = j & 127;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint4 w;
      w.x = pack_f16x2_rn(q[8 * c] * 1.4426950408889634f, q[8 * c + 1] * 1.4426950408889634f);
      w.y = pack_f16x2_rn(q[8 * c + 2] * 1.4426950408889634f, q[8 * c + 3] * 1.4426950408889634f);
      w.z = pack_f16x2_rn(q[8 * c + 4] * 1.4426950408889634f, q[8 * c + 5] * 1.4426950408889634f);
      w.w = pack_f16x2_rn(q[8 * c + 6] * 1.4426950408889634f, q[8 * c + 7] * 1.4426950408889634f);
      *reinterpret_cast<uint4*>(qdst + sw64_off(qr, c)) = w;
    }
    *reinterpret_cast<uint4*>(qdst + sw64_off(qr, 3)) = make_uint4(pack_f16x2_rn(0.f, A8_MASKED), 0u, 0u, 0u);
  } else if (j == S) {
#pragma unroll
    for (int i = 0; i < kHD; ++i) { k[i] = p.bias_k[h * kHD + i]; v[i] = p.bias_v[h * kHD + i]; }
  } else {
#pragma unroll
    for (int i = 0; i < kHD; ++i) { k[i] = 0.f; v[i] = 0.f; }
    masked = 1.f;
  }
  if (j <= S) rope24(k, p.cosT + j * kHalf, p.sinT + j * kHalf);
  // K row: 24 halfs | col 24 = 0 | col 25 = mask flag | zeros
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint4 w;
    w.x = pack_f16x2_rn(k[8 * c], k[8 * c + 1]); w.y = pack_f16x2_rn(k[8 * c + 2], k[8 * c + 3]);
    w.z = pack_f16x2_rn(k[8 * c + 4], k[8 * c + 5]); w.w = pack_f16x2_rn(k[8 * c + 6], k[8 * c + 7]);
    *reinterpret_cast<uint4*>(kbase + sw64_off(r, c)) = w;
  }
  *reinterpret_cast<uint4*>(kbase + sw64_off(r, 3)) = make_uint4(pack_f16x2_rn(0.f, masked), 0u, 0u, 0u);
  // V^T: element (d, key r) at row d, 16-byte chunk r >> 3, half r & 7; row 24 = 1
  const int kc = r >> 3, kw = r & 7;
#pragma unroll
  for (int d = 0; d < kHD; ++d)
    reinterpret_cast<uint16_t*>(vbase + sw128_off(d, kc))[kw] = (uint16_t)(pack_f16x2_rn(v[d], 0.f) & 0xFFFFu);
  reinterpret_cast<uint16_t*>(vbase + sw128_off(24, kc))[kw] = 0x3C00u;   // fp16 1.0
}

// ---------------------------------------------------------------------------------------------------------
// mbarriers per query tile t: s_full[t][b] (QK^T into S buffer b done) -> softmax; p_ready[t][b] (4 warps wrote
// P over buffer b) -> P·V -> pv_done[t][b] (O includes that tile);
// o_free[t] once per item (epilogue has read O). Shared: kv_full / kv_free per ring stage (kv_free counts the NQ
// issuers), q_full / q_free per query double buffer.
template <int NQ, bool POLY>
__global__ void __launch_bounds__(a8_threads(NQ), NQ == 4 ? 1 : 2) attn8_kernel(AttnParams p, const uint8_t* __restrict__ scratch,
                                                                            int total_items, int nqi /*items per (seq, head)*/) {
  constexpr int NSW = 4 * NQ;                    // softmax warps
  constexpr int W_TMA = NSW, W_MMA = NSW + 1;    // producer warp, first of the NQ MMA-issuer warps
  constexpr int QROWS = NQ * 128;                // queries per item
  constexpr int TMEM_COLS = NQ * A8_TSTRIDE;
  extern __shared__ uint8_t smem_raw[];
  const SeqMap& sm = p.sm;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const uint32_t q_off = 0;                                        // [2 buffers][NQ tiles][128 rows x 64 B]
  const uint32_t st_off = 2 * NQ * A8_QT_BYTES;                    // stage: [K 3072 | V^T 4096]
  const uint32_t bar_off = st_off + A8_STAGES * A8_IMG;
  auto b_sfull = [&](int t, int b) { return sbase + bar_off + 8 * (2 * t + b); };            // [NQ][2]
  auto b_pready = [&](int t, int b) { return sbase + bar_off + 64 + 8 * (2 * t + b); };      // [NQ][2]
  // P·V-done barriers, one per key-tile parity: with S double buffered a softmax warp may finish tile c before
  // P·V(c-1) retires, so a single barrier advancing once per tile would let a parity wait for tile c pass while
  // tile c-1 is still pending (the waiter must never be more than one phase behind)
  auto b_pvdone = [&](int t, int b) { return sbase + bar_off + 128 + 8 * (2 * t + b); };     // [NQ][2]
  auto b_ofree = [&](int t) { return sbase + bar_off + 192 + 8 * t; };                       // [NQ]
  auto b_kvfull = [&](int s) { return sbase + bar_off + 224 + 8 * s; };                      // [8]
  auto b_kvfree = [&](int s) { return sbase + bar_off + 288 + 8 * s; };                      // [8]
  auto b_qfull = [&](int b) { return sbase + bar_off + 352 + 8 * b; };                       // [2]
  auto b_qfree = [&](int b) { return sbase + bar_off + 368 + 8 * b; };                       // [2]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + bar_off + 384);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = sm.S;
  const int nkt = a8_nkt(S), nqt = a8_nqt_pad(S);
  const int bid = blockIdx.x, nblk = gridDim.x;
  const int n_my = (total_items - bid + nblk - 1) / nblk;          // items of this CTA (grid <= total_items)
  const int ttot = n_my * nkt;

  if (tid == 0) {
    for (int t = 0; t < NQ; ++t) {
      for (int b = 0; b < 2; ++b) { mbar_init(b_sfull(t, b), 1); mbar_init(b_pready(t, b), 4); mbar_init(b_pvdone(t, b), 1); }
      mbar_init(b_ofree(t), 4);
    }
    for (int i = 0; i < A8_STAGES; ++i) { mbar_init(b_kvfull(i), 1); mbar_init(b_kvfree(i), NQ); }
    for (int b = 0; b < 2; ++b) { mbar_init(b_qfull(b), 1); mbar_init(b_qfree(b), NQ); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    if (lane == 0) {
      // =============================== producer (TMA 1-D bulk copies) ===============================
      // per item: the NQ staged query tiles (one 8 KB image each, contiguous) into the Q double buffer, then one
      // K/V image per key tile into the ring
      const uint8_t* qimg0 = scratch + (size_t)sm.num_seq * kH * nkt * A8_IMG;
      int c = 0;
      for (int i = 0; i < n_my; ++i) {
        const int item = bid + i * nblk;
        const long long sh = item / nqi;
        const int qi = item % nqi;
        const int qb = i & 1, u = i >> 1;
        if (u > 0) mbar_wait_backoff(b_qfree(qb), (uint32_t)((u & 1) ^ 1));    // every QK^T of the previous user retired
        mbar_expect_tx(b_qfull(qb), NQ * A8_QT_BYTES);
        {
          const uint32_t dst = sbase + q_off + qb * NQ * A8_QT_BYTES;
          const uint8_t* src = qimg0 + ((size_t)sh * nqt + (size_t)qi * NQ) * A8_QT_BYTES;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
              ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"((uint32_t)(NQ * A8_QT_BYTES)), "r"(b_qfull(qb))
              : "memory");
        }
        const uint8_t* img = scratch + (size_t)sh * nkt * A8_IMG;
        for (int g = 0; g < nkt; ++g, ++c) {
          const int st = c % A8_STAGES, use = c / A8_STAGES;
          if (use > 0) mbar_wait_backoff(b_kvfree(st), (uint32_t)((use & 1) ^ 1));
          mbar_expect_tx(b_kvfull(st), A8_IMG);
          const uint32_t dst = sbase + st_off + st * A8_IMG;
          const uint8_t* src = img + (size_t)g * A8_IMG;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
              ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"((uint32_t)A8_IMG), "r"(b_kvfull(st))
              : "memory");
        }
      }
    }
  } else if (warp >= W_MMA) {
    if (lane == 0) {
      // =============================== MMA issuer of query tile t ===============================
      // program order: QK^T(0), QK^T(1); for every tile c: P·V(c), then QK^T(c+2) into the buffer P(c) occupied
      const int t = warp - W_MMA;
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, A8_KT);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 32);
      const uint32_t tT = tmem_base + t * A8_TSTRIDE;
      int qk_c = 0, qk_i = 0, qk_g = 0;             // cursor of the next QK^T (tile, item, tile-in-item)
      auto issue_qk = [&]() {
        const int qb = qk_i & 1, st = qk_c % A8_STAGES;
        if (qk_g == 0) mbar_wait(b_qfull(qb), (uint32_t)((qk_i >> 1) & 1));
        mbar_wait(b_kvfull(st), (uint32_t)((qk_c / A8_STAGES) & 1));
        tc_fence_after();
        const uint64_t kdesc = umma_desc_k64(sbase + st_off + st * A8_IMG);
        const uint64_t qdesc = umma_desc_k64(sbase + q_off + (qb * NQ + t) * A8_QT_BYTES);
        const uint32_t tS = tT + (qk_c & 1) * A8_S1;
        tc_mma_bf16(tS, qdesc, kdesc, idesc_qk, 0u);
        tc_mma_bf16(tS, qdesc + 2, kdesc + 2, idesc_qk, 1u);
        tc_commit(b_sfull(t, qk_c & 1));
        if (qk_g == nkt - 1) tc_commit(b_qfree(qb));                     // this tile's last QK^T of the item
        ++qk_c;
        if (++qk_g == nkt) { qk_g = 0; ++qk_i; }
      };
      if (ttot > 0) issue_qk();
      if (ttot > 1) issue_qk();
      int i = 0, g = 0;
      for (int c = 0; c < ttot; ++c) {
        const int st = c % A8_STAGES, buf = c & 1;
        mbar_wait(b_pready(t, buf), (uint32_t)((c >> 1) & 1));                   // the 4 warps of this tile wrote P(c)
        if (g == 0 && i > 0) mbar_wait(b_ofree(t), (uint32_t)((i - 1) & 1));     // previous item's O has been read
        tc_fence_after();
        const uint64_t vdesc = umma_desc_k128(sbase + st_off + st * A8_IMG + A8_K_BYTES);
        const uint32_t tP = tT + buf * A8_S1;
#pragma unroll
        for (int ks = 0; ks < A8_KT / 16; ++ks)
          tc_mma_bf16_ts(tT + A8_OCOL, tP + 8 * ks, vdesc + (uint64_t)(2 * ks), idesc_pv, (uint32_t)((g | ks) != 0));
        tc_commit(b_pvdone(t, buf));
        tc_commit(b_kvfree(st));                                                 // this tile is done with the stage
        if (c + 2 < ttot) issue_qk();                                            // refill this S buffer (in order after P·V(c))
        if (++g == nkt) { g = 0; ++i; }
      }
    }
  } else {
    // =============================== softmax + epilogue: one thread per query row ===============================
    const int t = warp >> 2, qq = warp & 3;
    const int row = qq * 32 + lane;
    const uint32_t tT = tmem_base + t * A8_TSTRIDE + ((uint32_t)(qq * 32) << 16);
    const uint32_t tO = tT + A8_OCOL;
    int c = 0;
    for (int i = 0; i < n_my; ++i) {
      float m_ref = -INFINITY;
      for (int g = 0; g < nkt; ++g, ++c) {
        const int buf = c & 1;
        const uint32_t tS = tT + buf * A8_S1;
        mbar_wait(b_sfull(t, buf), (uint32_t)((c >> 1) & 1));
        tc_fence_after();
        uint32_t va[32], vb[16];
        tc_ld32(tS, va);
        tc_ld16(tS + 32, vb);
        tc_ld_wait();
        // ---- row maximum of the 48 scores (two independent chains)
        float mx0 = fmaxf(__uint_as_float(va[0]), __uint_as_float(va[1])), mx1 = fmaxf(__uint_as_float(vb[0]), __uint_as_float(vb[1]));
#pragma unroll
        for (int k = 2; k < 32; k += 2) mx0 = fmaxf(mx0, fmaxf(__uint_as_float(va[k]), __uint_as_float(va[k + 1])));
#pragma unroll
        for (int k = 2; k < 16; k += 2) mx1 = fmaxf(mx1, fmaxf(__uint_as_float(vb[k]), __uint_as_float(vb[k + 1])));
        const float tmax = fmaxf(mx0, mx1);
        // ---- lazy reference: keep it unless the maximum outgrew it by more than 2^A8_LAZY (P stays <= 2^8: fp16 safe)
        const bool move = tmax > m_ref + A8_LAZY;          // always true on the first tile (m_ref = -inf)
        const float m_new = move ? tmax : m_ref;
        if (g > 0 && __any_sync(0xffffffffu, move)) {
          // the reference of some row moved: rescale this thread's O row (all 32 columns: head dims + row sum),
          // once P·V of the previous tile has retired
          mbar_wait(b_pvdone(t, (c - 1) & 1), (uint32_t)(((c - 1) >> 1) & 1));
          tc_fence_after();
          const float alpha = move ? ex2f(m_ref - m_new) : 1.0f;
#pragma unroll 1
          for (int cb = 0; cb < 32; cb += 8) {
            uint32_t o[8];
            tc_ld8(tO + cb, o);
            tc_ld_wait();
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tc_st8(tO + cb, o);
          }
        }
        m_ref = m_new;
        // ---- P = exp2(S - m_ref) as fp16 pairs (in the score registers), stored over the first 24 score columns
        // (every A8_POLY_EVERY-th pair is exponentiated on the FMA pipe, the others on the MUFU pipe)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float x0 = __uint_as_float(va[2 * k]) - m_new, x1 = __uint_as_float(va[2 * k + 1]) - m_new;
          va[k] = (POLY && (k % A8_POLY_EVERY) == A8_POLY_EVERY - 1) ? exp2_poly_pack_f16x2(x0, x1)
                                                                      : pack_f16x2_rn(ex2f(x0), ex2f(x1));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float x0 = __uint_as_float(vb[2 * k]) - m_new, x1 = __uint_as_float(vb[2 * k + 1]) - m_new;
          vb[k] = (POLY && (k % A8_POLY_EVERY) == A8_POLY_EVERY - 1) ? exp2_poly_pack_f16x2(x0, x1)
                                                                      : pack_f16x2_rn(ex2f(x0), ex2f(x1));
        }
        {
          uint32_t p0[16], p1[8];
#pragma unroll
          for (int k = 0; k < 16; ++k) p0[k] = va[k];
#pragma unroll
          for (int k = 0; k < 8; ++k) p1[k] = vb[k];
          tc_st16(tS, p0);
          tc_st8(tS + 16, p1);
        }
        tc_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_pready(t, buf));
      }
      // ---- epilogue of this item: O(t) row / row sum -> out
      mbar_wait(b_pvdone(t, (c - 1) & 1), (uint32_t)(((c - 1) >> 1) & 1));
      tc_fence_after();
      uint32_t o[32];
      tc_ld32(tO, o);
      tc_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_ofree(t));            // O is in registers: the next item may overwrite it
      const int item = bid + i * nblk;
      const int e = (item % nqi) * QROWS + t * 128 + row;
      if (e < S) {
        const long long sh = item / nqi;
        const int h = (int)(sh % kH);
        const long long tq = seq_token(sm, sh / kH, e);
        const float inv = 1.0f / __uint_as_float(o[24]);
#pragma unroll
        for (int k = 0; k < 6; ++k)
          store_operand4(p.out, (size_t)tq * kC + h * kHD + 4 * k,
                         make_float4(__uint_as_float(o[4 * k]) * inv, __uint_as_float(o[4 * k + 1]) * inv,
                                     __uint_as_float(o[4 * k + 2]) * inv, __uint_as_float(o[4 * k + 3]) * inv),
                         p.round_out);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

inline size_t attn8_scratch_bytes(const SeqMap& sm) {
  return (size_t)sm.num_seq * kH * ((size_t)a8_nkt(sm.S) * A8_IMG + (size_t)a8_nqt_pad(sm.S) * A8_QT_BYTES);
}

template <int NQ, bool POLY>
inline int attn8_launch_t(const AttnParams& p, const uint8_t* scratch, cudaStream_t s, std::string* err) {
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attn8_kernel<NQ, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, a8_smem_bytes(NQ));
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(attn8): ") + cudaGetErrorString(e);
      return -2;
    }
    configured[dev] = true;
  }
  const int nqi = (p.sm.S + NQ * 128 - 1) / (NQ * 128);
  const long long items = p.sm.num_seq * kH * nqi;
  if (items > 0x7fffffffLL) { if (err) *err = "attn8: too many work items"; return -2; }
  const int per_sm = NQ == 4 ? 1 : 2;
  const int grid = (int)std::min<long long>(items, (long long)per_sm * device_sm_count());
  attn8_kernel<NQ, POLY><<<grid, a8_threads(NQ), a8_smem_bytes(NQ), s>>>(p, scratch, (int)items, nqi);
  return 0;
}

// flags: bit 0 = force the 2-query-tile kernel (testing), bit 1 = all exponentials on the MUFU pipe (A/B of the
//        FMA-pipe polynomial that otherwise takes every A8_POLY_EVERY-th score pair)
inline int attn8_launch(const AttnParams& p, uint8_t* scratch, int flags, cudaStream_t s, std::string* err) {
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attn8_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A8P_SMEM);
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(attn8_prep): ") + cudaGetErrorString(e);
      return -2;
    }
    configured[dev] = true;
  }
  attn8_prep_kernel<<<(unsigned)(p.sm.num_seq * a8_nkt(p.sm.S) * 2), A8P_THREADS, A8P_SMEM, s>>>(p, scratch);
  const bool small = p.sm.S <= 256 || (flags & 1);
  const bool poly = (flags & 2) == 0;
  int rc = small ? (poly ? attn8_launch_t<2, true>(p, scratch, s, err) : attn8_launch_t<2, false>(p, scratch, s, err))
                 : (poly ? attn8_launch_t<4, true>(p, scratch, s, err) : attn8_launch_t<4, false>(p, scratch, s, err));
  if (rc) return rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("attn8 launch: ") + cudaGetErrorString(e);
    return -2;
  }
  return 0;
}

}  // namespace mdgen
