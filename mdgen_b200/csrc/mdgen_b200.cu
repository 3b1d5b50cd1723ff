// libmdgen_b200.so — C-ABI implementation (see include/mdgen_b200.h).
// Host-side orchestration of the sm_100a kernels: weight packing, workspace, the denoiser step
// (IPA trunk -> 5 factorised-attention layers -> final layer) and the native Euler loop.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/mdgen_b200.h"
#include "attention_simt.cuh"
#include "common.cuh"
#include "elementwise.cuh"
#include "gemm_simt.cuh"
#include "geometry.cuh"
#include "ipa.cuh"
#ifndef MDGEN_NO_TC
#include "gemm_tc.cuh"
#include "attention_tc.cuh"
#include "attention_v8.cuh"
#include "embed_step.cuh"
#endif

using namespace mdgen;

namespace {

std::string g_create_error;

struct RawTensor {
  float* ptr = nullptr;
  int64_t numel = 0;
};

// `*_tc` = TF32-rounded (round-to-nearest) copy consumed by the tensor-core GEMMs; the fp32
// master is kept for the SIMT validation path.
struct MhaW {
  float *wqkv, *bqkv, *wo, *bo, *bias_k, *bias_v;
  float *wqkv_tc, *wo_tc;
  float *wqkv_bf, *wo_bf;   // bf16-rounded fp32 copies (precision experiments only)
  uint16_t *wqkv_b16, *wo_b16;   // true bf16 copies for the kind::f16 GEMM path
  uint16_t *wqkv_f16, *wo_f16;   // fp16 copies (default operand format)
};
struct IpaLayerW {
  float *ln_g, *ln_b, *head_w, *wproj, *bproj, *wout, *bout, *wproj_tc, *wout_tc;
  MhaW mha;
  float *w1, *b1, *w2, *b2, *w1_tc, *w2_tc;
};
struct MainLayerW {
  MhaW mha_l, mha_t;
  float *w1, *b1, *w2, *b2, *w1_tc, *w2_tc, *w1_bf, *w2_bf;
  uint16_t *w1_b16, *w2_b16, *w1_f16, *w2_f16;
};

struct ProfEntry {
  std::string name;
  cudaEvent_t e0, e1;
};

}  // namespace

struct mdgen_handle {
  mdgen_config cfg;
  std::string err;
  std::map<std::string, RawTensor> raw;
  std::vector<void*> allocs;  // everything cudaMalloc'ed by the handle
  std::vector<void*> weight_allocs;   // the subset holding packed weights (freed when weights are re-finalised)
  bool alloc_is_weight = false;
  bool finalized = false;
  int64_t launches = 0;
#ifdef MDGEN_NO_TC
  int use_tc = 0;
#else
  int use_tc = 1;      // tcgen05 TF32 GEMMs for the token GEMMs (0 = fp32 SIMT validation path)
#endif
  // token GEMMs (QKV / out / fc1 / fc2): 2 = fp16 operands (kind::f16; TF32's 11-bit significand at half the
  // bytes and twice the MMA rate - the default), 1 = bf16 operands, 0 = TF32 operands (kind::tf32)
  int gemm_bf16 = 2;
  int emu_bf16 = 0;    // precision experiments: bit 0 MLP, bit 1 attention projections see bf16-rounded operands
  int use_graph = 1;                 // replay steps from a CUDA graph when the workload is launch-bound
  long long graph_max_tokens = 65536;
  cudaStream_t gstream = nullptr;
  cudaEvent_t gev_in = nullptr, gev_out = nullptr;
  int64_t graph_launches_per_pair = 0, graph_replays = 0;
  bool trunk_precomputed = false;   // set by mdgen_sample_euler while the step loop runs
  int use_tc_attn = 1; // tcgen05 attention for sequences longer than 64 (needs use_tc)
  int l4_variant = 1;   // S = 4 residue attention: 1 = warp-private shared-memory exchange (default), 0 = warp shuffles
#ifdef MDGEN_NO_TC
  int attn_variant = 0;
#else
  int attn_variant = kAttnVariantDefault;   // build variant of the tcgen05 attention (attention_tc.cuh: attn_tc_launch)
#endif
  int tc_min_rows = 1024;   // fewer rows (the IPA key-frame trunk) stay on the exact-fp32 skinny GEMM  // below this many rows the SIMT GEMM is used (latency-bound shapes)
  int profile = 0;
  std::vector<ProfEntry> prof;

  // packed weights
  std::vector<IpaLayerW> ipa;
  std::vector<MainLayerW> layers;
  float *w_lat = nullptr, *b_lat = nullptr, *w_cond = nullptr, *b_cond = nullptr, *e_mask = nullptr,
        *pos = nullptr, *aa_emb = nullptr, *wf = nullptr, *bf = nullptr, *wr = nullptr, *br = nullptr;
  float *w_t0 = nullptr, *b_t0 = nullptr, *w_t2 = nullptr, *b_t2 = nullptr, *w_ada = nullptr,
        *b_ada = nullptr, *w_fin = nullptr, *b_fin = nullptr, *freqs = nullptr, *inv_freq = nullptr;
  int modw = 0;

  // RoPE tables
  float *cosT = nullptr, *sinT = nullptr;
  int rope_n = 0;

  // residue tables
  float *tb_frame = nullptr, *tb_pos = nullptr, *tb_mask = nullptr;
  int* tb_group = nullptr;
  int* ft_chi_idx = nullptr;
  float *ft_chi_amask = nullptr, *ft_chi_mask = nullptr, *ft_bb_mask = nullptr;

  // workspace
  long long cap_tokens = 0, cap_rows = 0;
  long long cap_op_bytes = 0;   // (tokens + 128) x element size the GEMM-operand activation buffers hold (2 = bf16, 4 = fp32 / TF32)
  int cap_modrows = 0;
  float *h = nullptr, *xn = nullptr, *qkv = nullptr, *att = nullptr, *hid = nullptr, *cond = nullptr;
  float* ybuf = nullptr;   // gate * branch output of the out-proj / fc2 GEMMs when the residual add is fused into ln_mod
  int fuse_resid_ln = 0;   // 1: residual add in ln_mod_kernel (EPI_GATE GEMM epilogues; measured neutral: the bytes only move); 0 (default): in the GEMM epilogue
  int gemm_dbg = 0;          // measurement switches forwarded to the tensor-core GEMM (Epilogue::dbg)
  float *xi = nullptr, *xni = nullptr, *proj = nullptr, *cat = nullptr, *qkvi = nullptr, *atti = nullptr,
        *hidi = nullptr, *frot = nullptr, *ftrans = nullptr, *fmask = nullptr, *ipa_out = nullptr;
  float *tvals = nullptr, *sinus = nullptr, *h1 = nullptr, *st = nullptr, *mod = nullptr, *dt = nullptr;
  float *xbuf = nullptr, *xbuf2 = nullptr;  // ping-pong Euler state [N, D<=28]
  uint8_t* attn_scratch = nullptr;           // UMMA-ready key-tile images of the tcgen05 attention
  size_t attn_scratch_bytes = 0;
  int* step = nullptr;
  int reuse_cond = 0;          // mdgen_forward: keep h->cond of the previous call (stages of one adaptive ODE solve)
  long long cond_tokens = 0;   // tokens h->cond was built for
  double* err_partial = nullptr;
  float* err_out = nullptr;
};

namespace {

#define CUDA_TRY(h, expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                        \
      return MDGEN_E_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define CHECK_LAUNCH(h)                                                                     \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      (h)->err = std::string("kernel launch failed at ") + __FILE__ + ":" +                 \
                 std::to_string(__LINE__) + ": " + cudaGetErrorString(_e);                  \
      return MDGEN_E_CUDA;                                                                  \
    }                                                                                       \
    (h)->launches++;                                                                        \
  } while (0)

#define TRY(expr)                 \
  do {                            \
    int _rc = (expr);             \
    if (_rc != MDGEN_OK) return _rc; \
  } while (0)

struct ProfScope {
  mdgen_handle* h;
  cudaStream_t s;
  int idx = -1;
  ProfScope(mdgen_handle* h_, cudaStream_t s_, const char* name) : h(h_), s(s_) {
    if (!h->profile) return;
    ProfEntry e;
    e.name = name;
    cudaEventCreate(&e.e0);
    cudaEventCreate(&e.e1);
    cudaEventRecord(e.e0, s);
    h->prof.push_back(e);
    idx = (int)h->prof.size() - 1;
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(h->prof[idx].e1, s);
  }
};

int dev_alloc(mdgen_handle* h, void** p, size_t bytes) {
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {
    h->err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
    return MDGEN_E_NOMEM;
  }
  h->allocs.push_back(*p);
  if (h->alloc_is_weight) h->weight_allocs.push_back(*p);
  return MDGEN_OK;
}
template <typename T>
int dev_alloc_t(mdgen_handle* h, T** p, size_t n) {
  return dev_alloc(h, reinterpret_cast<void**>(p), n * sizeof(T));
}
void dev_free(mdgen_handle* h, void* p) {
  if (!p) return;
  for (size_t i = 0; i < h->allocs.size(); ++i)
    if (h->allocs[i] == p) {
      h->allocs.erase(h->allocs.begin() + i);
      break;
    }
  cudaFree(p);
}

int get_raw(mdgen_handle* h, const std::string& name, int64_t numel, const float** out) {
  auto it = h->raw.find(name);
  if (it == h->raw.end()) {
    h->err = "missing tensor '" + name + "'";
    return MDGEN_E_WEIGHTS;
  }
  if (it->second.numel != numel) {
    h->err = "tensor '" + name + "' has " + std::to_string(it->second.numel) + " elements, expected " +
             std::to_string(numel);
    return MDGEN_E_WEIGHTS;
  }
  *out = it->second.ptr;
  return MDGEN_OK;
}

// dst[row0 + r, :] = scale * src[r, :]  (optionally TF32-rounded)
int pack(mdgen_handle* h, cudaStream_t s, const std::string& name, long long rows, int cols, float* dst,
         int dst_ld, long long row0, float scale, int do_round) {
  const float* src;
  TRY(get_raw(h, name, rows * cols, &src));
  long long n = rows * cols;
  pack_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, rows, cols, dst_ld, row0, scale,
                                                                do_round);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

int pack_new(mdgen_handle* h, cudaStream_t s, const std::string& name, long long rows, int cols, float** dst,
             float scale = 1.f, int do_round = 0) {
  TRY(dev_alloc_t(h, dst, (size_t)rows * cols));
  return pack(h, s, name, rows, cols, *dst, cols, 0, scale, do_round);
}

// TF32-rounded copy of an already packed fp32 matrix.
int tc_copy(mdgen_handle* h, cudaStream_t s, const float* src, size_t n, float** dst, int mode = 1) {
  TRY(dev_alloc_t(h, dst, n));
  pack_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, *dst, 1, (int)n, (int)n, 0, 1.f, mode);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

// true bf16 / fp16 copy of an fp32 matrix
int b16_copy(mdgen_handle* h, cudaStream_t s, const float* src, size_t n, uint16_t** dst) {
  TRY(dev_alloc_t(h, dst, n));
  to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, *dst, (long long)n);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}
int f16_copy(mdgen_handle* h, cudaStream_t s, const float* src, size_t n, uint16_t** dst) {
  TRY(dev_alloc_t(h, dst, n));
  to_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, *dst, (long long)n);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

int pack_mha(mdgen_handle* h, cudaStream_t s, const std::string& p, MhaW* w) {
  const float scale = 0.20412414523193154f;  // 24^-0.5 folded into W_q, b_q (mha.py:102,263)
  TRY(dev_alloc_t(h, &w->wqkv, (size_t)kQKV * kC));
  TRY(dev_alloc_t(h, &w->bqkv, (size_t)kQKV));
  TRY(pack(h, s, p + "attn.q_proj.weight", kC, kC, w->wqkv, kC, 0, scale, 0));
  TRY(pack(h, s, p + "attn.k_proj.weight", kC, kC, w->wqkv, kC, kC, 1.f, 0));
  TRY(pack(h, s, p + "attn.v_proj.weight", kC, kC, w->wqkv, kC, 2 * kC, 1.f, 0));
  TRY(tc_copy(h, s, w->wqkv, (size_t)kQKV * kC, &w->wqkv_tc));
  TRY(tc_copy(h, s, w->wqkv, (size_t)kQKV * kC, &w->wqkv_bf, 2));
  TRY(b16_copy(h, s, w->wqkv, (size_t)kQKV * kC, &w->wqkv_b16));
  TRY(f16_copy(h, s, w->wqkv, (size_t)kQKV * kC, &w->wqkv_f16));
  TRY(pack(h, s, p + "attn.q_proj.bias", 1, kC, w->bqkv, kQKV, 0, scale, 0));
  TRY(pack(h, s, p + "attn.k_proj.bias", 1, kC, w->bqkv + kC, kQKV, 0, 1.f, 0));
  TRY(pack(h, s, p + "attn.v_proj.bias", 1, kC, w->bqkv + 2 * kC, kQKV, 0, 1.f, 0));
  TRY(pack_new(h, s, p + "attn.out_proj.weight", kC, kC, &w->wo));
  TRY(tc_copy(h, s, w->wo, (size_t)kC * kC, &w->wo_tc));
  TRY(tc_copy(h, s, w->wo, (size_t)kC * kC, &w->wo_bf, 2));
  TRY(b16_copy(h, s, w->wo, (size_t)kC * kC, &w->wo_b16));
  TRY(f16_copy(h, s, w->wo, (size_t)kC * kC, &w->wo_f16));
  TRY(pack_new(h, s, p + "attn.out_proj.bias", 1, kC, &w->bo));
  TRY(pack_new(h, s, p + "attn.bias_k", 1, kC, &w->bias_k));
  TRY(pack_new(h, s, p + "attn.bias_v", 1, kC, &w->bias_v));
  return MDGEN_OK;
}

int ensure_rope(mdgen_handle* h, cudaStream_t s, int n) {
  if (n <= h->rope_n) return MDGEN_OK;
  int cap = n + 64;
  if (h->cosT) { dev_free(h, h->cosT); dev_free(h, h->sinT); }
  TRY(dev_alloc_t(h, &h->cosT, (size_t)cap * kHalf));
  TRY(dev_alloc_t(h, &h->sinT, (size_t)cap * kHalf));
  rope_table_kernel<<<(cap * kHalf + 255) / 256, 256, 0, s>>>(h->inv_freq, h->cosT, h->sinT, cap);
  CHECK_LAUNCH(h);
  h->rope_n = cap;
  return MDGEN_OK;
}

int ensure_workspace(mdgen_handle* h, long long N, long long rows, int modrows) {
  // the GEMM-operand activations (xn, q|k|v, att, hid) are bf16 on the default path and sized accordingly
  const bool b16 = h->use_tc && h->gemm_bf16 && h->use_tc_attn && N >= h->tc_min_rows;
  const int elem = b16 ? 2 : 4;
  if (N > h->cap_tokens) {
    float** bufs[] = {&h->h, &h->cond, &h->xbuf, &h->xbuf2, &h->ybuf};
    for (auto b : bufs) { dev_free(h, *b); *b = nullptr; }
    h->cap_tokens = 0;
    size_t n = (size_t)N + 128;
    TRY(dev_alloc_t(h, &h->h, n * kC));
    TRY(dev_alloc_t(h, &h->ybuf, n * kC));
    TRY(dev_alloc_t(h, &h->cond, n * kC));
    TRY(dev_alloc_t(h, &h->xbuf, n * 28));
    TRY(dev_alloc_t(h, &h->xbuf2, n * 28));
    h->cap_tokens = N;
  }
  if ((N + 128) * elem > h->cap_op_bytes) {
    float** bufs[] = {&h->xn, &h->qkv, &h->att, &h->hid};
    for (auto b : bufs) { dev_free(h, *b); *b = nullptr; }
    h->cap_op_bytes = 0;
    // +128 rows of slack so tensor-core tiles may over-read the last partial tile
    const size_t nb = (size_t)(N + 128) * elem;
    TRY(dev_alloc(h, reinterpret_cast<void**>(&h->xn), nb * kC));
    TRY(dev_alloc(h, reinterpret_cast<void**>(&h->qkv), nb * kQKV));
    TRY(dev_alloc(h, reinterpret_cast<void**>(&h->att), nb * kC));
    TRY(dev_alloc(h, reinterpret_cast<void**>(&h->hid), nb * kFF));
    h->cap_op_bytes = (long long)nb;
  }
  if (rows > h->cap_rows) {
    float** bufs[] = {&h->xi, &h->xni, &h->proj, &h->cat, &h->qkvi, &h->atti, &h->hidi,
                      &h->frot, &h->ftrans, &h->fmask, &h->ipa_out};
    for (auto b : bufs) { dev_free(h, *b); *b = nullptr; }
    size_t n = (size_t)rows + 128;
    TRY(dev_alloc_t(h, &h->xi, n * kC));
    TRY(dev_alloc_t(h, &h->xni, n * kC));
    TRY(dev_alloc_t(h, &h->proj, n * kIpaProj));
    TRY(dev_alloc_t(h, &h->cat, n * kIpaCat));
    TRY(dev_alloc_t(h, &h->qkvi, n * kQKV));
    TRY(dev_alloc_t(h, &h->atti, n * kC));
    TRY(dev_alloc_t(h, &h->hidi, n * kFF));
    TRY(dev_alloc_t(h, &h->frot, n * 9));
    TRY(dev_alloc_t(h, &h->ftrans, n * 3));
    TRY(dev_alloc_t(h, &h->fmask, n));
    TRY(dev_alloc_t(h, &h->ipa_out, n * kC));
    h->cap_rows = rows;
  }
  if (modrows > h->cap_modrows) {
    float** bufs[] = {&h->tvals, &h->sinus, &h->h1, &h->st, &h->mod, &h->dt};
    for (auto b : bufs) { dev_free(h, *b); *b = nullptr; }
    size_t r = (size_t)modrows;
    TRY(dev_alloc_t(h, &h->tvals, r));
    TRY(dev_alloc_t(h, &h->sinus, r * kTFreq));
    TRY(dev_alloc_t(h, &h->h1, r * kC));
    TRY(dev_alloc_t(h, &h->st, r * kC));
    TRY(dev_alloc_t(h, &h->mod, r * h->modw));
    TRY(dev_alloc_t(h, &h->dt, r));
    h->cap_modrows = modrows;
  }
  if (!h->step) TRY(dev_alloc_t(h, &h->step, 4));
  return MDGEN_OK;
}

// ---- GEMM dispatch ---------------------------------------------------------------------------
// W = fp32 master (SIMT paths), W_tc = tensor-core operand copy: TF32-rounded fp32, or true bf16 when
// in_bf16 (then A is a bf16 buffer too); out_bf16: the epilogue stores bf16.
int gemm(mdgen_handle* h, cudaStream_t s, int mode, const float* A, int lda, const float* W,
         const void* W_tc, int ldw, long long M, int N, int K, const Epilogue& ep_in, const char* tag,
         int half_fmt = 0 /*kFmtBF16 / kFmtF16: 16-bit A and W*/, bool out_bf16 = false /*16-bit output*/) {
  ProfScope ps(h, s, tag);
  const bool in_bf16 = half_fmt != 0;
  Epilogue ep = ep_in;
  ep.half_fmt = half_fmt;
  ep.dbg = h->gemm_dbg;
#ifndef MDGEN_NO_TC
  if (h->use_tc && W_tc && M >= h->tc_min_rows && tc_gemm_supported(N, K, in_bf16)) {
    int rc = tc_gemm_launch(mode, A, lda, W_tc, ldw, M, N, K, ep, s, &h->err, in_bf16, out_bf16);
    if (rc != MDGEN_OK) return rc;
    h->launches++;
    return MDGEN_OK;
  }
#endif
  if (in_bf16 || out_bf16) { h->err = "internal: bf16 operands need the tensor-core GEMM"; return MDGEN_E_INVALID; }
  if (M <= 4096 && N % SK_BN == 0 && K % SK_BK == 0 && lda % 4 == 0 && ldw % 4 == 0) {
    dim3 g((unsigned)((M + SK_BM - 1) / SK_BM), (unsigned)(N / SK_BN));
    switch (mode) {
      case EPI_STORE: gemm_skinny_kernel<EPI_STORE><<<g, 128, SK_SMEM_BYTES, s>>>(A, lda, W, ldw, M, N, K, ep); break;
      case EPI_GELU: gemm_skinny_kernel<EPI_GELU><<<g, 128, SK_SMEM_BYTES, s>>>(A, lda, W, ldw, M, N, K, ep); break;
      case EPI_RESID_GATE: gemm_skinny_kernel<EPI_RESID_GATE><<<g, 128, SK_SMEM_BYTES, s>>>(A, lda, W, ldw, M, N, K, ep); break;
      case EPI_RESID: gemm_skinny_kernel<EPI_RESID><<<g, 128, SK_SMEM_BYTES, s>>>(A, lda, W, ldw, M, N, K, ep); break;
      default: h->err = "bad epilogue mode"; return MDGEN_E_INVALID;
    }
    CHECK_LAUNCH(h);
    return MDGEN_OK;
  }
  dim3 grid((unsigned)((M + SG_BM - 1) / SG_BM), (unsigned)((N + SG_BN - 1) / SG_BN));
  switch (mode) {
    case EPI_STORE: gemm_simt_kernel<EPI_STORE><<<grid, 256, 0, s>>>(A, lda, W, ldw, M, N, K, ep); break;
    case EPI_GELU: gemm_simt_kernel<EPI_GELU><<<grid, 256, 0, s>>>(A, lda, W, ldw, M, N, K, ep); break;
    case EPI_RESID_GATE: gemm_simt_kernel<EPI_RESID_GATE><<<grid, 256, 0, s>>>(A, lda, W, ldw, M, N, K, ep); break;
    case EPI_RESID: gemm_simt_kernel<EPI_RESID><<<grid, 256, 0, s>>>(A, lda, W, ldw, M, N, K, ep); break;
    default: h->err = "bad epilogue mode"; return MDGEN_E_INVALID;
  }
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

Epilogue make_epi(const float* bias, float* out, int ldo, int round_out = 0) {
  Epilogue e;
  memset(&e, 0, sizeof(e));
  e.bias = bias; e.out = out; e.ldo = ldo; e.round_out = round_out;
  return e;
}
Epilogue make_epi_gate(const float* bias, float* x, int ldo, const ModRef& mod, int gate_off) {
  Epilogue e = make_epi(bias, x, ldo);
  e.resid = x; e.mod = mod; e.gate_off = gate_off;
  return e;
}

int attention(mdgen_handle* h, cudaStream_t s, const float* qkv, const float* mask, const MhaW& w,
              float* out, const SeqMap& sm, int round_out, const char* tag, bool allow_tc = true,
              int qkv_fmt = 0) {
  ProfScope ps(h, s, tag);
  AttnParams p;
  p.qkv = qkv; p.qkv_fmt = qkv_fmt; p.mask = mask; p.bias_k = w.bias_k; p.bias_v = w.bias_v;
  p.cosT = h->cosT; p.sinT = h->sinT; p.out = out; p.round_out = round_out; p.sm = sm;
#ifndef MDGEN_NO_TC
  if (allow_tc && h->use_tc && h->use_tc_attn && sm.S > 64) {
    size_t need = std::max(attn_tc_scratch_bytes(sm), attn8_scratch_bytes(sm));
    if (need > h->attn_scratch_bytes) {
      if (h->attn_scratch) dev_free(h, h->attn_scratch);
      h->attn_scratch = nullptr; h->attn_scratch_bytes = 0;
      TRY(dev_alloc(h, reinterpret_cast<void**>(&h->attn_scratch), need));
      CUDA_TRY(h, cudaMemsetAsync(h->attn_scratch, 0, need, s));   // V^T pad rows must read as zeros
      h->attn_scratch_bytes = need;
    }
    if (h->attn_variant & 256) {     // generation 8 (default): fp16 operands, one softmax thread per query row
      if (attn8_launch(p, h->attn_scratch, h->attn_variant & 255, s, &h->err)) return MDGEN_E_CUDA;
    } else if (attn_tc_launch(p, h->attn_scratch, h->attn_variant, s, &h->err)) return MDGEN_E_CUDA;
    h->launches += 2;
    return MDGEN_OK;
  }
#endif
  if (sm.S == 4) {
    long long threads = sm.num_seq * 2 * 32;
    if (h->l4_variant == 1) attn_l4s_kernel<<<(unsigned)((sm.num_seq * 2 + 3) / 4), 128, 0, s>>>(p);
    else attn_l4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(p);
  } else if (sm.S <= 64) {
    long long total = sm.num_seq * sm.S * kH;
    attn_small_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(p);
  } else {
    dim3 grid((sm.S + AF_QT - 1) / AF_QT, kH, (unsigned)sm.num_seq);
    attn_flash_simt_kernel<<<grid, 128, 0, s>>>(p);
  }
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

int ln_mod(mdgen_handle* h, cudaStream_t s, float* x, float* y, const ModRef& mod, int shift_off,
           int scale_off, long long N, int rmode, const float* y_add = nullptr) {
  ProfScope ps(h, s, "ln_mod");
  unsigned blocks = (unsigned)((N * 32 + 255) / 256);
  ln_mod_kernel<<<blocks, 256, 0, s>>>(x, y_add, y, mod, shift_off, scale_off, N, rmode);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

int check_cond(mdgen_handle* h, const mdgen_cond* c) {
  if (!h->finalized) { h->err = "weights not finalised"; return MDGEN_E_WEIGHTS; }
  if (!c || c->B <= 0 || c->T <= 0 || c->L <= 0) { h->err = "bad cond dims"; return MDGEN_E_INVALID; }
  if (!c->mask || !c->start_rot || !c->start_trans || !c->x_cond || !c->x_cond_mask) {
    h->err = "cond: mask/start frames/x_cond/x_cond_mask are required"; return MDGEN_E_INVALID;
  }
  bool two = !h->cfg.sim_condition && (h->cfg.tps_condition || h->cfg.inpainting);
  if (two && (!c->end_rot || !c->end_trans)) { h->err = "cond: end frames required"; return MDGEN_E_INVALID; }
  if (h->cfg.use_aa_emb && !c->aatype) { h->err = "cond: aatype required"; return MDGEN_E_INVALID; }
  if (h->cfg.abs_pos_emb && c->L != h->cfg.crop) {
    h->err = "abs_pos_emb requires L == crop"; return MDGEN_E_INVALID;
  }
  return MDGEN_OK;
}

// Timestep embedding + every adaLN modulation vector for `R` time rows (tvals already on device).
int build_mod_table(mdgen_handle* h, cudaStream_t s, int R) {
  ProfScope ps(h, s, "adaln_table");
  sinus_kernel<<<R, 128, 0, s>>>(h->tvals, h->cfg.time_multiplier, h->freqs, h->sinus, R);
  CHECK_LAUNCH(h);
  rowdot_kernel<8, 1><<<(kC * 32 + 255) / 256, 256, 0, s>>>(h->w_t0, h->b_t0, h->sinus, h->h1, kC, R);
  CHECK_LAUNCH(h);
  // st = SiLU(temb): every consumer applies SiLU first (Sequential(SiLU, Linear))
  rowdot_kernel<12, 1><<<(kC * 32 + 255) / 256, 256, 0, s>>>(h->w_t2, h->b_t2, h->h1, h->st, kC, R);
  CHECK_LAUNCH(h);
  rowdot_kernel<12, 0><<<(unsigned)(((long long)h->modw * 32 + 255) / 256), 256, 0, s>>>(
      h->w_ada, h->b_ada, h->st, h->mod, h->modw, R);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

// Step-invariant conditioning embedding (latent_model.py:233-241 minus latent_to_emb(x)).
int build_cond(mdgen_handle* h, cudaStream_t s, const mdgen_cond* c) {
  ProfScope ps(h, s, "cond_embed");
  long long N = (long long)c->B * c->T * c->L;
  embed_kernel<0><<<(unsigned)((N + kEmbedTok - 1) / kEmbedTok), kC, 0, s>>>(
      c->x_cond, h->cfg.latent_dim, h->w_cond, h->b_lat, h->b_cond, h->pos, h->e_mask, c->x_cond_mask,
      nullptr, nullptr, nullptr, 0, h->cond, N, c->T, c->L);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

// IPA key-frame trunk (latent_model.py:175-210, 369-384), exact fp32 end to end: its output is
// broadcast-added to every frame, so TF32 rounding here is coherent over the whole trajectory (measured:
// 40x the final-state error of TF32 in the token GEMMs). Evaluates `nrep` replicas stacked on the row
// axis: nrep = 1 with the usual modulation-row selection (forward / per-step), or nrep = K Euler steps
// at once (replica k uses modulation row k) — the trunk never sees x, so the sampler hoists it out of
// the step loop.   Output: h->ipa_out [nrep][B*L][C].
int run_ipa_trunk(mdgen_handle* h, cudaStream_t s, const mdgen_cond* c, int nrep, const int* step_ptr, int bstride) {
  ProfScope ps(h, s, "ipa_trunk");
  const int B = c->B, T = c->T, L = c->L, n = h->cfg.num_layers;
  const bool two = !h->cfg.sim_condition && (h->cfg.tps_condition || h->cfg.inpainting);
  const long long BL = (long long)B * L, rows1 = (two ? 2 : 1) * BL, rows = rows1 * nrep;
  ModRef modi = (nrep > 1) ? ModRef{h->mod, nullptr, h->modw, 1, (int)rows1, nrep}
                           : ModRef{h->mod, step_ptr, h->modw, bstride, L, B};
  ipa_init_kernel<<<(unsigned)rows, kC, 0, s>>>(two ? 1 : 0, c->start_rot, c->start_trans, c->end_rot,
                                               c->end_trans, c->aatype, h->cfg.use_aa_emb ? h->aa_emb : nullptr,
                                               h->wf, h->bf, h->wr, h->br, c->mask, T, L, h->xi, h->frot,
                                               h->ftrans, h->fmask, BL, two ? c->quat_sign : nullptr);
  CHECK_LAUNCH(h);
  SeqMap smi{L, rows / L, 1, (long long)L, 0, 1};
  for (int i = 0; i < n; ++i) {
    const IpaLayerW& w = h->ipa[i];
    int off = i * 6 * kC;
    ln_affine_kernel<false><<<(unsigned)((rows * 32 + 255) / 256), 256, 0, s>>>(h->xi, h->xni, w.ln_g, w.ln_b, rows);
    CHECK_LAUNCH(h);
    TRY(gemm(h, s, EPI_STORE, h->xni, kC, w.wproj, nullptr, kC, rows, kIpaProj, kC,
             make_epi(w.bproj, h->proj, kIpaProj), "ipa_gemm"));
    ipa_points_kernel<<<(unsigned)((rows * 96 + 255) / 256), 256, 0, s>>>(h->proj, h->frot, h->ftrans, rows);
    CHECK_LAUNCH(h);
    ipa_attn_kernel<<<(unsigned)rows, 128, 4 * L * sizeof(float), s>>>(h->proj, h->frot, h->ftrans, h->fmask,
                                                                       w.head_w, h->cat, L, 0);
    CHECK_LAUNCH(h);
    Epilogue eo = make_epi(w.bout, h->xi, kC);
    eo.resid = h->xi;
    TRY(gemm(h, s, EPI_RESID, h->cat, kIpaCat, w.wout, nullptr, kIpaCat, rows, kC, kIpaCat, eo, "ipa_gemm"));
    TRY(ln_mod(h, s, h->xi, h->xni, modi, off + 0, off + kC, rows, 0));
    TRY(gemm(h, s, EPI_STORE, h->xni, kC, w.mha.wqkv, nullptr, kC, rows, kQKV, kC,
             make_epi(w.mha.bqkv, h->qkvi, kQKV), "ipa_gemm"));
    TRY(attention(h, s, h->qkvi, h->fmask, w.mha, h->atti, smi, 0, "ipa_mha", /*allow_tc=*/false));
    TRY(gemm(h, s, EPI_RESID_GATE, h->atti, kC, w.mha.wo, nullptr, kC, rows, kC, kC,
             make_epi_gate(w.mha.bo, h->xi, kC, modi, off + 2 * kC), "ipa_gemm"));
    TRY(ln_mod(h, s, h->xi, h->xni, modi, off + 3 * kC, off + 4 * kC, rows, 0));
    TRY(gemm(h, s, EPI_GELU, h->xni, kC, w.w1, nullptr, kC, rows, kFF, kC, make_epi(w.b1, h->hidi, kFF), "ipa_gemm"));
    TRY(gemm(h, s, EPI_RESID_GATE, h->hidi, kFF, w.w2, nullptr, kFF, rows, kC, kFF,
             make_epi_gate(w.b2, h->xi, kC, modi, off + 5 * kC), "ipa_gemm"));
  }
  long long ne = BL * kC * nrep;
  ipa_sum_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, s>>>(h->xi, h->ipa_out, ne, BL * kC, two ? 1 : 0);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

// One denoiser evaluation on state x_in. euler: x_out = x_in + dt[step] * v ; else x_out = v.
int run_step(mdgen_handle* h, cudaStream_t s, const mdgen_cond* c, const float* x_in, float* x_out,
             bool euler, const int* step_ptr, int bstride) {
  const int B = c->B, T = c->T, L = c->L, n = h->cfg.num_layers, D = h->cfg.latent_dim;
  const long long N = (long long)B * T * L;
  const int rt = (h->use_tc && N >= h->tc_min_rows) ? 1 : 0;  // round GEMM-operand activations to TF32

  ModRef modm{h->mod, step_ptr, h->modw, bstride, T * L, B};

  // ---------------- IPA trunk on the key frame(s): per step here, unless the sampler precomputed it
  // for all steps at once (the trunk depends on t, the frames and aatype only — never on x)
  if (!h->trunk_precomputed) TRY(run_ipa_trunk(h, s, c, 1, step_ptr, bstride));

  // ---------------- token embedding  (latent_model.py:233-246)
  {
    ProfScope ps(h, s, "embed");
#ifndef MDGEN_NO_TC
    if (embed_step_configure() != cudaSuccess) { h->err = "cudaFuncSetAttribute(embed_step_kernel) failed"; return MDGEN_E_CUDA; }
    embed_step_kernel<<<(unsigned)((N + kEsTok - 1) / kEsTok), kC, kEsSmemBytes, s>>>(
        x_in, D, h->w_lat, h->cond, h->ipa_out, h->trunk_precomputed ? step_ptr : nullptr, (long long)B * L * kC,
        h->h, N, T, L);
#else
    embed_kernel<1><<<(unsigned)((N + kEmbedTok - 1) / kEmbedTok), kC, 0, s>>>(
        x_in, D, h->w_lat, nullptr, nullptr, nullptr, nullptr, nullptr, h->cond, h->ipa_out,
        h->trunk_precomputed ? step_ptr : nullptr, (long long)B * L * kC, h->h, N, T, L);
#endif
    CHECK_LAUNCH(h);
  }

  // ---------------- main layers  (latent_model.py:446-483)
  SeqMap sml{L, (long long)B * T, 1, (long long)L, 0, 1};
  SeqMap smt{T, (long long)B * L, L, (long long)T * L, 1, L};
  for (int i = 0; i < n; ++i) {
    const MainLayerW& w = h->layers[i];
    int off = n * 6 * kC + i * 9 * kC;
    // GEMM operand modes of the token GEMMs: 3 = true bf16 storage + kind::f16 MMA (default),
    // 1 = TF32-rounded fp32 + kind::tf32 MMA, 2 = bf16-rounded values through the TF32 MMA (precision
    // experiments, option "emu_bf16": bit 0 MLP, bit 1 attention projections), 0 = fp32 SIMT path.
    const int hf = (rt && h->gemm_bf16 && tc_gemm_supported(kQKV, kC, true)) ? (h->gemm_bf16 == 2 ? kFmtF16 : kFmtBF16) : 0;
    const bool bf = hf != 0;
    const int hfq = (bf && h->use_tc_attn) ? hf : 0;   // q|k|v stored 16-bit (the SIMT flash kernel reads fp32 only)
    const bool bfq = hfq != 0;
    const int rm16 = hf == kFmtF16 ? 4 : 3;            // store_operand4 mode of the 16-bit format
    const int rm_attn = !rt ? 0 : (bf ? rm16 : ((h->emu_bf16 & 2) ? 2 : 1));
    const int rm_mlp = !rt ? 0 : (bf ? rm16 : ((h->emu_bf16 & 1) ? 2 : 1));
    auto pick = [&](const uint16_t* f16, const uint16_t* b16, const float* emu, const float* tc, int emu_bit) -> const void* {
      return bf ? (const void*)(hf == kFmtF16 ? f16 : b16) : ((h->emu_bf16 & emu_bit) ? (const void*)emu : (const void*)tc);
    };
    const void* wqkv_l = pick(w.mha_l.wqkv_f16, w.mha_l.wqkv_b16, w.mha_l.wqkv_bf, w.mha_l.wqkv_tc, 2);
    const void* wo_l = pick(w.mha_l.wo_f16, w.mha_l.wo_b16, w.mha_l.wo_bf, w.mha_l.wo_tc, 2);
    const void* wqkv_t = pick(w.mha_t.wqkv_f16, w.mha_t.wqkv_b16, w.mha_t.wqkv_bf, w.mha_t.wqkv_tc, 2);
    const void* wo_t = pick(w.mha_t.wo_f16, w.mha_t.wo_b16, w.mha_t.wo_bf, w.mha_t.wo_tc, 2);
    const void* w1x = pick(w.w1_f16, w.w1_b16, w.w1_bf, w.w1_tc, 1);
    const void* w2x = pick(w.w2_f16, w.w2_b16, w.w2_bf, w.w2_tc, 1);
    // The gated branch outputs (out-proj, fc2) either add into the residual stream in their GEMM epilogue
    // (EPI_RESID_GATE) or - default on the tensor-core path - are stored as gate * branch (EPI_GATE) and added by the
    // next ln_mod_kernel, which reads the residual stream anyway. The last fc2 always adds in its epilogue (the
    // final layer's LayerNorm lives in final_kernel).
    const bool fuse = rt && h->fuse_resid_ln;
    auto branch_out = [&](const float* A, int lda, const float* W, const void* Wx, int ldw, int K, const float* bias,
                          int gate_off, const char* tag, bool defer) -> int {
      if (defer) {
        Epilogue e = make_epi(bias, h->ybuf, kC);
        e.mod = modm; e.gate_off = gate_off;
        return gemm(h, s, EPI_GATE, A, lda, W, Wx, ldw, N, kC, K, e, tag, hf);
      }
      return gemm(h, s, EPI_RESID_GATE, A, lda, W, Wx, ldw, N, kC, K, make_epi_gate(bias, h->h, kC, modm, gate_off), tag, hf);
    };
    // residue attention (over L)
    TRY(ln_mod(h, s, h->h, h->xn, modm, off + 0, off + kC, N, rm_attn, (fuse && i > 0) ? h->ybuf : nullptr));
    TRY(gemm(h, s, EPI_STORE, h->xn, kC, w.mha_l.wqkv, wqkv_l, kC, N, kQKV, kC,
             make_epi(w.mha_l.bqkv, h->qkv, kQKV, (h->emu_bf16 & 4) ? 2 : 0), "gemm_qkv", hf, bfq));
    TRY(attention(h, s, h->qkv, c->mask, w.mha_l, h->att, sml, rm_attn, "mha_l", true, hfq));
    TRY(branch_out(h->att, kC, w.mha_l.wo, wo_l, kC, kC, w.mha_l.bo, off + 2 * kC, "gemm_out", fuse));
    // time attention (over T)
    TRY(ln_mod(h, s, h->h, h->xn, modm, off + 3 * kC, off + 4 * kC, N, rm_attn, fuse ? h->ybuf : nullptr));
    TRY(gemm(h, s, EPI_STORE, h->xn, kC, w.mha_t.wqkv, wqkv_t, kC, N, kQKV, kC,
             make_epi(w.mha_t.bqkv, h->qkv, kQKV, (h->emu_bf16 & 4) ? 2 : 0), "gemm_qkv", hf, bfq));
    TRY(attention(h, s, h->qkv, c->mask, w.mha_t, h->att, smt, rm_attn, "mha_t", true, hfq));
    TRY(branch_out(h->att, kC, w.mha_t.wo, wo_t, kC, kC, w.mha_t.bo, off + 5 * kC, "gemm_out", fuse));
    // MLP (hidden activations are 16-bit on the default path: written by fc1, read only by fc2)
    TRY(ln_mod(h, s, h->h, h->xn, modm, off + 6 * kC, off + 7 * kC, N, rm_mlp, fuse ? h->ybuf : nullptr));
    TRY(gemm(h, s, EPI_GELU, h->xn, kC, w.w1, w1x, kC, N, kFF, kC, make_epi(w.b1, h->hid, kFF, bf ? 0 : rm_mlp),
             "gemm_fc1", hf, bf));
    TRY(branch_out(h->hid, kFF, w.w2, w2x, kFF, kFF, w.b2, off + 8 * kC, "gemm_fc2", fuse && i + 1 < n));
  }

  // ---------------- final layer (+ Euler update)
  {
    ProfScope ps(h, s, "final");
    int off = n * 15 * kC;
    int blocks = (int)std::min<long long>((N + 15) / 16, 148 * 2);   // 2 resident blocks per SM (124 registers): W staged once each
    size_t smem = (size_t)D * kC * sizeof(float);
    if (euler)
      final_kernel<true><<<blocks, 256, smem, s>>>(h->h, modm, off, off + kC, h->w_fin, h->b_fin, D, x_in,
                                                  h->dt, x_out, N);
    else
      final_kernel<false><<<blocks, 256, smem, s>>>(h->h, modm, off, off + kC, h->w_fin, h->b_fin, D, x_in,
                                                   h->dt, x_out, N);
    CHECK_LAUNCH(h);
  }
  return MDGEN_OK;
}

int prepare_call(mdgen_handle* h, cudaStream_t s, const mdgen_cond* c, int modrows, int trunk_reps = 1) {
  TRY(check_cond(h, c));
  long long N = (long long)c->B * c->T * c->L;
  bool two = !h->cfg.sim_condition && (h->cfg.tps_condition || h->cfg.inpainting);
  TRY(ensure_workspace(h, N, (two ? 2 : 1) * (long long)c->B * c->L * trunk_reps, modrows));
  TRY(ensure_rope(h, s, std::max(c->T, c->L) + 1));
  return MDGEN_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int mdgen_abi_version(void) { return MDGEN_ABI_VERSION; }

const char* mdgen_last_error(const mdgen_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mdgen_create(const mdgen_config* cfg, mdgen_handle** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return MDGEN_E_INVALID; }
  if (cfg->abi_version != MDGEN_ABI_VERSION) { g_create_error = "ABI version mismatch"; return MDGEN_E_INVALID; }
  if (cfg->latent_dim != 21 && cfg->latent_dim != 28) { g_create_error = "latent_dim must be 21 or 28"; return MDGEN_E_INVALID; }
  if (cfg->num_layers < 1 || cfg->num_layers > 16) { g_create_error = "num_layers out of range"; return MDGEN_E_INVALID; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                     " (mdgen_b200 has no CPU fallback)";
    return MDGEN_E_CUDA;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  if (prop.major != 10) {
    g_create_error = std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor) + "; this library is built for sm_100a (B200) only";
    return MDGEN_E_CUDA;
  }
  mdgen_handle* h = new mdgen_handle();
  h->cfg = *cfg;
  h->modw = cfg->num_layers * 15 * kC + 2 * kC;
  *out = h;
  return MDGEN_OK;
}

void mdgen_destroy(mdgen_handle* h) {
  if (!h) return;
  if (h->gstream) { cudaStreamDestroy(h->gstream); cudaEventDestroy(h->gev_in); cudaEventDestroy(h->gev_out); }
  for (void* p : h->allocs) cudaFree(p);
  for (auto& e : h->prof) { cudaEventDestroy(e.e0); cudaEventDestroy(e.e1); }
  delete h;
}

int mdgen_set_tensor(mdgen_handle* h, const char* name, const float* data, int64_t numel) {
  if (!h || !name || !data || numel <= 0) { if (h) h->err = "mdgen_set_tensor: bad argument"; return MDGEN_E_INVALID; }
  RawTensor& r = h->raw[name];
  if (r.ptr && r.numel != numel) { dev_free(h, r.ptr); r.ptr = nullptr; }
  if (!r.ptr) TRY(dev_alloc_t(h, &r.ptr, (size_t)numel));
  r.numel = numel;
  CUDA_TRY(h, cudaMemcpy(r.ptr, data, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice));
  h->finalized = false;
  return MDGEN_OK;
}

static int finalize_weights_impl(mdgen_handle* h, cudaStream_t s);

int mdgen_finalize_weights(mdgen_handle* h, void* stream) {
  if (!h) return MDGEN_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  // a reload replaces the previous generation of packed weights: wait for work that may still read them,
  // then release them (every packed copy is re-created below)
  if (!h->weight_allocs.empty()) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    std::vector<void*> old;
    old.swap(h->weight_allocs);
    for (void* p : old) dev_free(h, p);
  }
  h->finalized = false;
  h->alloc_is_weight = true;
  int rc = finalize_weights_impl(h, s);
  h->alloc_is_weight = false;
  return rc;
}

static int finalize_weights_impl(mdgen_handle* h, cudaStream_t s) {
  const mdgen_config& c = h->cfg;
  const int n = c.num_layers, D = c.latent_dim;
  const bool two_any = c.tps_condition || c.inpainting;
  TRY(pack_new(h, s, "latent_to_emb.weight", kC, D, &h->w_lat));
  TRY(pack_new(h, s, "latent_to_emb.bias", 1, kC, &h->b_lat));
  TRY(pack_new(h, s, "cond_to_emb.weight", kC, D, &h->w_cond));
  TRY(pack_new(h, s, "cond_to_emb.bias", 1, kC, &h->b_cond));
  TRY(pack_new(h, s, "mask_to_emb.weight", 2, kC, &h->e_mask));
  if (c.abs_pos_emb) TRY(pack_new(h, s, "pos_embed", c.crop, kC, &h->pos));
  if (c.use_aa_emb) TRY(pack_new(h, s, "aatype_to_emb.weight", 21, kC, &h->aa_emb));
  if (two_any) {
    TRY(pack_new(h, s, "latent_to_emb_f.weight", kC, 7, &h->wf));
    TRY(pack_new(h, s, "latent_to_emb_f.bias", 1, kC, &h->bf));
    TRY(pack_new(h, s, "latent_to_emb_r.weight", kC, 7, &h->wr));
    TRY(pack_new(h, s, "latent_to_emb_r.bias", 1, kC, &h->br));
  }
  TRY(pack_new(h, s, "t_embedder.mlp.0.weight", kC, kTFreq, &h->w_t0));
  TRY(pack_new(h, s, "t_embedder.mlp.0.bias", 1, kC, &h->b_t0));
  TRY(pack_new(h, s, "t_embedder.mlp.2.weight", kC, kC, &h->w_t2));
  TRY(pack_new(h, s, "t_embedder.mlp.2.bias", 1, kC, &h->b_t2));
  TRY(pack_new(h, s, "emb_to_latent.linear.weight", D, kC, &h->w_fin));
  TRY(pack_new(h, s, "emb_to_latent.linear.bias", 1, D, &h->b_fin));
  // concatenated adaLN table weights: [ipa 0..n-1 (6C) | main 0..n-1 (9C) | final (2C)]
  TRY(dev_alloc_t(h, &h->w_ada, (size_t)h->modw * kC));
  TRY(dev_alloc_t(h, &h->b_ada, (size_t)h->modw));
  h->ipa.resize(n);
  h->layers.resize(n);
  for (int i = 0; i < n; ++i) {
    std::string p = "ipa_layers." + std::to_string(i) + ".";
    IpaLayerW& w = h->ipa[i];
    TRY(pack(h, s, p + "adaLN_modulation.1.weight", 6 * kC, kC, h->w_ada, kC, (long long)i * 6 * kC, 1.f, 0));
    TRY(pack(h, s, p + "adaLN_modulation.1.bias", 1, 6 * kC, h->b_ada + (size_t)i * 6 * kC, 6 * kC, 0, 1.f, 0));
    TRY(pack_new(h, s, p + "ipa_norm.weight", 1, kC, &w.ln_g));
    TRY(pack_new(h, s, p + "ipa_norm.bias", 1, kC, &w.ln_b));
    TRY(pack_new(h, s, p + "ipa.head_weights", 1, kIpaH, &w.head_w));
    TRY(dev_alloc_t(h, &w.wproj, (size_t)kIpaProj * kC));
    TRY(dev_alloc_t(h, &w.bproj, (size_t)kIpaProj));
    CUDA_TRY(h, cudaMemsetAsync(w.wproj, 0, (size_t)kIpaProj * kC * sizeof(float), s));   // rows 672..767 = pad
    CUDA_TRY(h, cudaMemsetAsync(w.bproj, 0, (size_t)kIpaProj * sizeof(float), s));
    TRY(pack(h, s, p + "ipa.linear_q.weight", 128, kC, w.wproj, kC, 0, 1.f, 0));
    TRY(pack(h, s, p + "ipa.linear_kv.weight", 256, kC, w.wproj, kC, 128, 1.f, 0));
    TRY(pack(h, s, p + "ipa.linear_q_points.weight", 96, kC, w.wproj, kC, 384, 1.f, 0));
    TRY(pack(h, s, p + "ipa.linear_kv_points.weight", 192, kC, w.wproj, kC, 480, 1.f, 0));
    TRY(pack(h, s, p + "ipa.linear_q.bias", 1, 128, w.bproj, kIpaProj, 0, 1.f, 0));
    TRY(pack(h, s, p + "ipa.linear_kv.bias", 1, 256, w.bproj + 128, kIpaProj, 0, 1.f, 0));
    TRY(pack(h, s, p + "ipa.linear_q_points.bias", 1, 96, w.bproj + 384, kIpaProj, 0, 1.f, 0));
    TRY(pack(h, s, p + "ipa.linear_kv_points.bias", 1, 192, w.bproj + 480, kIpaProj, 0, 1.f, 0));
    TRY(tc_copy(h, s, w.wproj, (size_t)kIpaProj * kC, &w.wproj_tc));
    TRY(pack_new(h, s, p + "ipa.linear_out.weight", kC, kIpaCat, &w.wout));
    TRY(tc_copy(h, s, w.wout, (size_t)kC * kIpaCat, &w.wout_tc));
    TRY(pack_new(h, s, p + "ipa.linear_out.bias", 1, kC, &w.bout));
    TRY(pack_mha(h, s, p + "mha_l.", &w.mha));
    TRY(pack_new(h, s, p + "fc1.weight", kFF, kC, &w.w1));
    TRY(tc_copy(h, s, w.w1, (size_t)kFF * kC, &w.w1_tc));
    TRY(pack_new(h, s, p + "fc1.bias", 1, kFF, &w.b1));
    TRY(pack_new(h, s, p + "fc2.weight", kC, kFF, &w.w2));
    TRY(tc_copy(h, s, w.w2, (size_t)kC * kFF, &w.w2_tc));
    TRY(pack_new(h, s, p + "fc2.bias", 1, kC, &w.b2));
  }
  for (int i = 0; i < n; ++i) {
    std::string p = "layers." + std::to_string(i) + ".";
    MainLayerW& w = h->layers[i];
    long long row0 = (long long)n * 6 * kC + (long long)i * 9 * kC;
    TRY(pack(h, s, p + "adaLN_modulation.1.weight", 9 * kC, kC, h->w_ada, kC, row0, 1.f, 0));
    TRY(pack(h, s, p + "adaLN_modulation.1.bias", 1, 9 * kC, h->b_ada + row0, 9 * kC, 0, 1.f, 0));
    TRY(pack_mha(h, s, p + "mha_l.", &w.mha_l));
    TRY(pack_mha(h, s, p + "mha_t.", &w.mha_t));
    TRY(pack_new(h, s, p + "fc1.weight", kFF, kC, &w.w1));
    TRY(tc_copy(h, s, w.w1, (size_t)kFF * kC, &w.w1_tc));
    TRY(tc_copy(h, s, w.w1, (size_t)kFF * kC, &w.w1_bf, 2));
    TRY(b16_copy(h, s, w.w1, (size_t)kFF * kC, &w.w1_b16));
    TRY(f16_copy(h, s, w.w1, (size_t)kFF * kC, &w.w1_f16));
    TRY(pack_new(h, s, p + "fc1.bias", 1, kFF, &w.b1));
    TRY(pack_new(h, s, p + "fc2.weight", kC, kFF, &w.w2));
    TRY(tc_copy(h, s, w.w2, (size_t)kC * kFF, &w.w2_tc));
    TRY(tc_copy(h, s, w.w2, (size_t)kC * kFF, &w.w2_bf, 2));
    TRY(b16_copy(h, s, w.w2, (size_t)kC * kFF, &w.w2_b16));
    TRY(f16_copy(h, s, w.w2, (size_t)kC * kFF, &w.w2_f16));
    TRY(pack_new(h, s, p + "fc2.bias", 1, kC, &w.b2));
  }
  {
    long long row0 = (long long)n * 15 * kC;
    TRY(pack(h, s, "emb_to_latent.adaLN_modulation.1.weight", 2 * kC, kC, h->w_ada, kC, row0, 1.f, 0));
    TRY(pack(h, s, "emb_to_latent.adaLN_modulation.1.bias", 1, 2 * kC, h->b_ada + row0, 2 * kC, 0, 1.f, 0));
  }
  // rotary inverse frequencies: all attention modules share the same buffer values
  TRY(pack_new(h, s, "layers.0.mha_t.attn.rot_emb.inv_freq", 1, kHalf, &h->inv_freq));
  // timestep-embedding frequencies f_i = exp(-ln(1e4) i / 128)  (layers.py:43-45)
  {
    float f[128];
    for (int i = 0; i < 128; ++i) f[i] = expf(-logf(10000.0f) * (float)i / 128.0f);
    TRY(dev_alloc_t(h, &h->freqs, 128));
    CUDA_TRY(h, cudaMemcpyAsync(h->freqs, f, sizeof(f), cudaMemcpyHostToDevice, s));
  }
  CUDA_TRY(h, cudaStreamSynchronize(s));
  // raw copies are no longer needed
  for (auto& kv : h->raw) dev_free(h, kv.second.ptr);
  h->raw.clear();
  h->rope_n = 0;
  h->finalized = true;
  return MDGEN_OK;
}

int mdgen_set_residue_tables(mdgen_handle* h, const float* default_frame, const float* atom14_group_pos,
                             const int32_t* atom14_to_group, const float* atom14_mask) {
  if (!h || !default_frame || !atom14_group_pos || !atom14_to_group || !atom14_mask) return MDGEN_E_INVALID;
  if (!h->tb_frame) {
    TRY(dev_alloc_t(h, &h->tb_frame, 21 * 8 * 16));
    TRY(dev_alloc_t(h, &h->tb_pos, 21 * 14 * 3));
    TRY(dev_alloc_t(h, &h->tb_group, 21 * 14));
    TRY(dev_alloc_t(h, &h->tb_mask, 21 * 14));
  }
  CUDA_TRY(h, cudaMemcpy(h->tb_frame, default_frame, 21 * 8 * 16 * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->tb_pos, atom14_group_pos, 21 * 14 * 3 * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->tb_group, atom14_to_group, 21 * 14 * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->tb_mask, atom14_mask, 21 * 14 * 4, cudaMemcpyHostToDevice));
  return MDGEN_OK;
}

int mdgen_set_featurize_tables(mdgen_handle* h, const int32_t* chi_atom14_idx, const float* chi_atom_mask,
                               const float* chi_mask, const float* bb_mask) {
  if (!h || !chi_atom14_idx || !chi_atom_mask || !chi_mask || !bb_mask) return MDGEN_E_INVALID;
  if (!h->ft_chi_idx) {
    TRY(dev_alloc_t(h, &h->ft_chi_idx, 21 * 16));
    TRY(dev_alloc_t(h, &h->ft_chi_amask, 21 * 16));
    TRY(dev_alloc_t(h, &h->ft_chi_mask, 21 * 4));
    TRY(dev_alloc_t(h, &h->ft_bb_mask, 21 * 4));
  }
  CUDA_TRY(h, cudaMemcpy(h->ft_chi_idx, chi_atom14_idx, 21 * 16 * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->ft_chi_amask, chi_atom_mask, 21 * 16 * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->ft_chi_mask, chi_mask, 21 * 4 * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->ft_bb_mask, bb_mask, 21 * 4 * 4, cudaMemcpyHostToDevice));
  return MDGEN_OK;
}

int mdgen_featurize_atom14(mdgen_handle* h, int32_t B, int32_t L, const float* atom14, const int64_t* seqres,
                           float* rots, float* trans, float* torsions, float* torsion_mask, void* stream) {
  if (!h || !atom14 || !seqres || !rots || !trans || !torsions || B <= 0 || L <= 0) {
    if (h) h->err = "mdgen_featurize_atom14: bad argument";
    return MDGEN_E_INVALID;
  }
  if (!h->ft_chi_idx) { h->err = "featurisation tables not set"; return MDGEN_E_WEIGHTS; }
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope ps(h, s, "featurize");
  FeatTables tb{h->ft_chi_idx, h->ft_chi_amask, h->ft_chi_mask, h->ft_bb_mask};
  long long N = (long long)B * L;
  featurize_kernel<<<(unsigned)((N + 127) / 128), 128, 0, s>>>(atom14, seqres, tb, rots, trans, torsions,
                                                              torsion_mask, B, L);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

int mdgen_forward(mdgen_handle* h, const float* x, const float* t, const mdgen_cond* cond, float* out,
                  void* stream) {
  if (!h || !x || !t || !out || !cond) { if (h) h->err = "mdgen_forward: null argument"; return MDGEN_E_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  h->trunk_precomputed = false;
  TRY(prepare_call(h, s, cond, cond ? cond->B : 0));
  CUDA_TRY(h, cudaMemcpyAsync(h->tvals, t, (size_t)cond->B * sizeof(float), cudaMemcpyDeviceToDevice, s));
  TRY(build_mod_table(h, s, cond->B));
  const long long ntok = (long long)cond->B * cond->T * cond->L;
  if (!(h->reuse_cond && h->cond_tokens == ntok)) TRY(build_cond(h, s, cond));
  h->cond_tokens = ntok;
  return run_step(h, s, cond, x, out, /*euler=*/false, nullptr, /*bstride=*/1);
}

int mdgen_sample_euler(mdgen_handle* h, const float* zs, const float* t_grid, int32_t K,
                       const mdgen_cond* cond, float* x_out, void* stream) {
  if (!h || !zs || !t_grid || !x_out || !cond || K < 1) { if (h) h->err = "mdgen_sample_euler: bad argument"; return MDGEN_E_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  // the IPA trunk of all K steps is evaluated at once when its stacked rows stay modest
  const bool two_t = !h->cfg.sim_condition && (h->cfg.tps_condition || h->cfg.inpainting);
  const long long trunk_rows = (long long)K * (two_t ? 2 : 1) * cond->B * cond->L;
  const bool hoist = trunk_rows <= 262144;
  TRY(prepare_call(h, s, cond, K, hoist ? K : 1));
  // time rows t_k (k < K) and fp32 step sizes dt_k = t_{k+1} - t_k (integrators.py:90; torchdiffeq)
  std::vector<float> dt(K);
  for (int k = 0; k < K; ++k) dt[k] = t_grid[k + 1] - t_grid[k];
  CUDA_TRY(h, cudaMemcpyAsync(h->tvals, t_grid, (size_t)K * sizeof(float), cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->dt, dt.data(), (size_t)K * sizeof(float), cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));  // dt (host vector) must be consumed before it goes out of scope
  TRY(build_mod_table(h, s, K));
  TRY(build_cond(h, s, cond));
  h->cond_tokens = 0;
  if (hoist) TRY(run_ipa_trunk(h, s, cond, K, nullptr, 0));
  h->trunk_precomputed = hoist;
  step_set_kernel<<<1, 1, 0, s>>>(h->step, 0);
  CHECK_LAUNCH(h);
  // ping-pong Euler state: x_k in bufA/bufB alternately; the last step writes x_out
  float* bufA = h->xbuf;
  float* bufB = h->xbuf2;
  auto one_step = [&](cudaStream_t st, const float* src, float* dst) -> int {
    int rc = run_step(h, st, cond, src, dst, /*euler=*/true, h->step, /*bstride=*/0);
    if (rc != MDGEN_OK) return rc;
    step_advance_kernel<<<1, 1, 0, st>>>(h->step);
    CHECK_LAUNCH(h);
    return MDGEN_OK;
  };
  const long long Ntok = (long long)cond->B * cond->T * cond->L;
  // Small workloads (e.g. sim_inference.py's one trajectory per call) are launch-bound: ~125 launches
  // per step. Every step issues the same launch sequence (the step index lives in device memory), so
  // steps 1..K-2 are replayed from a CUDA graph that holds one even/odd pair of steps. The graph runs on
  // a private capturable stream (PyTorch's current stream is usually the un-capturable legacy stream),
  // fenced against the caller's stream with events. Any capture failure falls back to eager launches.
  const bool want_graph = h->use_graph && !h->profile && K >= 6 && Ntok <= h->graph_max_tokens;
  int k = 0;
  const float* cur = zs;
  auto finish = [&](int rc) { h->trunk_precomputed = false; return rc; };
  if (want_graph) {
    if (!h->gstream) {
      if (cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->gev_in, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->gev_out, cudaEventDisableTiming) != cudaSuccess) {
        h->gstream = nullptr;
        (void)cudaGetLastError();
      }
    }
  }
  if (want_graph && h->gstream) {
    cudaStream_t g = h->gstream;
    CUDA_TRY(h, cudaEventRecord(h->gev_in, s));
    CUDA_TRY(h, cudaStreamWaitEvent(g, h->gev_in, 0));
    // step 0 eagerly (zs -> bufA): also performs every lazy allocation / attribute set of the step
    int rc = one_step(g, zs, K == 1 ? x_out : bufA);
    if (rc != MDGEN_OK) return finish(rc);
    k = 1; cur = bufA;
    const int pairs = (K - 2) / 2;              // steps 1 .. 2*pairs run from the graph, in (A->B, B->A) pairs
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool ok = cudaStreamBeginCapture(g, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      const int64_t launches0 = h->launches;
      int r1 = one_step(g, bufA, bufB);
      int r2 = (r1 == MDGEN_OK) ? one_step(g, bufB, bufA) : r1;
      cudaError_t ce = cudaStreamEndCapture(g, &graph);
      ok = (r1 == MDGEN_OK && r2 == MDGEN_OK && ce == cudaSuccess && graph != nullptr);
      h->graph_launches_per_pair = h->launches - launches0;
      h->launches = launches0;                   // captured, not yet executed
      if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    }
    (void)cudaGetLastError();
    if (ok) {
      for (int p = 0; p < pairs && ok; ++p) {
        ok = cudaGraphLaunch(exec, g) == cudaSuccess;
        if (ok) { k += 2; h->launches += h->graph_launches_per_pair; }
      }
      h->graph_replays += pairs;
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    (void)cudaGetLastError();
    // remaining steps (and everything, if the graph could not be built) eagerly on the same stream
    for (; k < K; ++k) {
      float* nxt = (k == K - 1) ? x_out : ((k & 1) == 0 ? bufA : bufB);
      rc = one_step(g, cur, nxt);
      if (rc != MDGEN_OK) return finish(rc);
      cur = nxt;
    }
    CUDA_TRY(h, cudaEventRecord(h->gev_out, g));
    CUDA_TRY(h, cudaStreamWaitEvent(s, h->gev_out, 0));
    return finish(MDGEN_OK);
  }
  for (; k < K; ++k) {
    float* nxt = (k == K - 1) ? x_out : ((k & 1) == 0 ? bufA : bufB);
    // (K == 1 with x_out == zs updates in place: each state element is read and written by the same thread)
    int rc = one_step(s, cur, nxt);
    if (rc != MDGEN_OK) return finish(rc);
    cur = nxt;
  }
  h->trunk_precomputed = false;
  return MDGEN_OK;
}

int mdgen_prep_batch(mdgen_handle* h, int32_t B, int32_t T, int32_t L, const float* rots, const float* trans,
                     const float* torsions, float* latents, float* x_cond, int64_t* x_cond_mask,
                     void* stream) {
  if (!h || !rots || !trans || !torsions || !latents || !x_cond || !x_cond_mask || B <= 0 || T <= 0 || L <= 0) {
    if (h) h->err = "mdgen_prep_batch: bad argument";
    return MDGEN_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope ps(h, s, "prep");
  PrepFlags f;
  f.D = h->cfg.latent_dim;
  f.two = (h->cfg.tps_condition || h->cfg.inpainting) ? 1 : 0;
  f.sim_condition = h->cfg.sim_condition; f.tps_condition = h->cfg.tps_condition;
  f.inpainting = h->cfg.inpainting; f.cond_interval = h->cfg.cond_interval; f.no_torsion = h->cfg.no_torsion;
  long long N = (long long)B * T * L;
  prep_kernel<<<(unsigned)((N + 127) / 128), 128, 0, s>>>(rots, trans, torsions, latents, x_cond, x_cond_mask,
                                                         B, T, L, f);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

int mdgen_decode_atom14(mdgen_handle* h, int32_t B, int32_t T, int32_t L, const float* samples,
                        const float* start_rot, const float* start_trans, const int64_t* seqres,
                        float* atom14, void* stream) {
  if (!h || !samples || !start_rot || !start_trans || !seqres || !atom14 || B <= 0 || T <= 0 || L <= 0) {
    if (h) h->err = "mdgen_decode_atom14: bad argument";
    return MDGEN_E_INVALID;
  }
  if (!h->tb_frame) { h->err = "residue tables not set"; return MDGEN_E_WEIGHTS; }
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope ps(h, s, "decode");
  ResidueTables tb{h->tb_frame, h->tb_pos, h->tb_group, h->tb_mask};
  int tors_off = (h->cfg.tps_condition || h->cfg.inpainting) ? 14 : 7;
  long long N = (long long)B * T * L;
  decode_kernel<<<(unsigned)((N + 127) / 128), 128, 0, s>>>(samples, h->cfg.latent_dim, tors_off, start_rot,
                                                           start_trans, seqres, tb, atom14, B, T, L);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

int mdgen_flow_plan(mdgen_handle* h, int32_t B, int64_t per, int32_t path, const float* x1, const float* x0,
                    const float* t, float* xt, float* ut, void* stream) {
  if (!h || !x1 || !x0 || !t || !xt || !ut || B <= 0 || per <= 0 || path < 0 || path > 1) {
    if (h) h->err = "mdgen_flow_plan: bad argument";
    return MDGEN_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope ps(h, s, "flow_plan");
  const long long n = (long long)B * per;
  flow_plan_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, s>>>(x1, x0, t, xt, ut, per, n, path);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

int mdgen_masked_mse(mdgen_handle* h, int32_t B, int64_t per, const float* pred, const float* target,
                     const float* mask, float* loss, void* stream) {
  if (!h || !pred || !target || !mask || !loss || B <= 0 || per <= 0) {
    if (h) h->err = "mdgen_masked_mse: bad argument";
    return MDGEN_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope ps(h, s, "masked_mse");
  masked_mse_kernel<<<(unsigned)B, 1024, 0, s>>>(pred, target, mask, loss, per);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

int mdgen_ema_update(mdgen_handle* h, float* stored, const float* param, int64_t n, float decay, void* stream) {
  if (!h || !stored || !param || n <= 0) { if (h) h->err = "mdgen_ema_update: bad argument"; return MDGEN_E_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  ema_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(stored, param, n, 1.0f - decay);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

int mdgen_lincomb(mdgen_handle* h, int64_t n, const float* y, float scale, const float* coeffs,
                  const float* const* ks, int32_t nk, float* out, void* stream) {
  if (!h || !out || n <= 0 || nk < 0 || nk > 8 || (nk > 0 && (!coeffs || !ks))) {
    if (h) h->err = "mdgen_lincomb: bad argument";
    return MDGEN_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  LinComb lc;
  lc.nk = 0;
  for (int i = 0; i < nk; ++i)
    if (coeffs[i] != 0.f) { lc.k[lc.nk] = ks[i]; lc.c[lc.nk] = coeffs[i]; ++lc.nk; }
  for (int i = lc.nk; i < 8; ++i) { lc.k[i] = nullptr; lc.c[i] = 0.f; }
  lincomb_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, s>>>(y, scale, lc, out, n);
  CHECK_LAUNCH(h);
  return MDGEN_OK;
}

int mdgen_rk_error_ratio(mdgen_handle* h, int64_t n, const float* err, const float* y0, const float* y1,
                         float rtol, float atol, float* ratio, void* stream) {
  if (!h || !err || !y0 || !y1 || !ratio || n <= 0) { if (h) h->err = "mdgen_rk_error_ratio: bad argument"; return MDGEN_E_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  if (!h->err_partial) {
    TRY(dev_alloc_t(h, &h->err_partial, kErrBlocks));
    TRY(dev_alloc_t(h, &h->err_out, 4));
  }
  rk_error_partial_kernel<<<kErrBlocks, 256, 0, s>>>(err, y0, y1, rtol, atol, h->err_partial, n);
  CHECK_LAUNCH(h);
  rk_error_final_kernel<<<1, 32, 0, s>>>(h->err_partial, kErrBlocks, n, h->err_out);
  CHECK_LAUNCH(h);
  CUDA_TRY(h, cudaMemcpyAsync(ratio, h->err_out, sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  return MDGEN_OK;
}

int mdgen_debug_linear(mdgen_handle* h, const float* A, const float* W, const float* bias, int64_t M, int32_t N,
                       int32_t K, int32_t act, int32_t use_tc, float* out, void* stream) {
  if (!h || !A || !W || !out || M <= 0 || N <= 0 || K <= 0) return MDGEN_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  Epilogue ep = make_epi(bias, out, N);
  int mode = act ? EPI_GELU : EPI_STORE;
  if (use_tc) {
#ifndef MDGEN_NO_TC
    // 1: TF32 operands; 2 / 3: bf16 / fp16 operands, fp32 output; 4 / 5: bf16 / fp16 operands AND output (the QKV / fc1
    // form, bulk-tensor-store epilogue), widened to fp32 into `out` afterwards
    const bool out16 = use_tc >= 4;
    if (out16) use_tc -= 2;
    const bool bf = use_tc >= 2;
    ep.half_fmt = use_tc == 3 ? kFmtF16 : kFmtBF16;
    uint16_t* out_h = nullptr;
    if (out16) { TRY(dev_alloc_t(h, &out_h, (size_t)M * N)); ep.out = reinterpret_cast<float*>(out_h); }
    if (!tc_gemm_supported(N, K, bf)) { h->err = "shape unsupported by the tensor-core GEMM"; return MDGEN_E_INVALID; }
    float *Ar = nullptr, *Wr = nullptr;   // rounded operand copies (fp32 containers or bf16 arrays)
    TRY(dev_alloc_t(h, &Ar, (size_t)M * K));
    TRY(dev_alloc_t(h, &Wr, (size_t)N * K));
    if (use_tc == 3) {
      to_f16_kernel<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, s>>>(A, (uint16_t*)Ar, (long long)M * K);
      to_f16_kernel<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, s>>>(W, (uint16_t*)Wr, (long long)N * K);
    } else if (bf) {
      to_bf16_kernel<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, s>>>(A, (uint16_t*)Ar, (long long)M * K);
      to_bf16_kernel<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, s>>>(W, (uint16_t*)Wr, (long long)N * K);
    } else {
      pack_rows_kernel<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, s>>>(A, Ar, M, K, K, 0, 1.f, 1);
      pack_rows_kernel<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, s>>>(W, Wr, N, K, K, 0, 1.f, 1);
    }
    int rc = tc_gemm_launch(mode, Ar, K, Wr, K, M, N, K, ep, s, &h->err, bf, out16);
    if (out16 && rc == 0)
      from_half_kernel<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, s>>>(out_h, out, (long long)M * N, ep.half_fmt);
    cudaStreamSynchronize(s);
    dev_free(h, Ar);
    dev_free(h, Wr);
    if (out_h) dev_free(h, out_h);
    return rc == 0 ? MDGEN_OK : MDGEN_E_CUDA;
#else
    h->err = "library built without tensor-core kernels";
    return MDGEN_E_INVALID;
#endif
  }
  int save = h->use_tc;
  h->use_tc = 0;
  int rc = gemm(h, s, mode, A, K, W, (const void*)nullptr, K, M, N, K, ep, "debug");
  h->use_tc = save;
  return rc;
}

int64_t mdgen_launch_count(const mdgen_handle* h) { return h ? h->launches : -1; }

int mdgen_set_option(mdgen_handle* h, const char* key, int64_t value) {
  if (!h || !key) return MDGEN_E_INVALID;
  std::string k(key);
  if (k == "use_tc") {
#ifdef MDGEN_NO_TC
    if (value) { h->err = "library built without tensor-core kernels"; return MDGEN_E_INVALID; }
#endif
    h->use_tc = (int)value;
  } else if (k == "tc_min_rows") h->tc_min_rows = (int)value;
  else if (k == "use_tc_attn") h->use_tc_attn = (int)value;
  else if (k == "attn_variant") h->attn_variant = (int)value & 511;
  else if (k == "l4_variant") h->l4_variant = (int)value & 1;
  else if (k == "emu_bf16") h->emu_bf16 = (int)value;
  else if (k == "gemm_bf16") h->gemm_bf16 = (int)value;
  else if (k == "use_graph") h->use_graph = (int)value;
  else if (k == "fuse_resid_ln") h->fuse_resid_ln = (int)value;
  else if (k == "gemm_dbg") h->gemm_dbg = (int)value;
  else if (k == "reuse_cond") { h->reuse_cond = (int)value; if (!value) h->cond_tokens = 0; }
  else if (k == "graph_max_tokens") h->graph_max_tokens = value;
  else if (k == "profile") {
    h->profile = (int)value;
    if (!value) {
      for (auto& e : h->prof) { cudaEventDestroy(e.e0); cudaEventDestroy(e.e1); }
      h->prof.clear();
    }
  } else { h->err = "unknown option " + k; return MDGEN_E_INVALID; }
  return MDGEN_OK;
}

int64_t mdgen_get_option(const mdgen_handle* h, const char* key) {
  if (!h || !key) return -1;
  std::string k(key);
  if (k == "use_tc") return h->use_tc;
  if (k == "tc_min_rows") return h->tc_min_rows;
  if (k == "use_tc_attn") return h->use_tc_attn;
  if (k == "attn_variant") return h->attn_variant;
  if (k == "l4_variant") return h->l4_variant;
  if (k == "gemm_bf16") return h->gemm_bf16;
  if (k == "use_graph") return h->use_graph;
  if (k == "fuse_resid_ln") return h->fuse_resid_ln;
  if (k == "gemm_dbg") return h->gemm_dbg;
  if (k == "graph_replays") return h->graph_replays;
  if (k == "profile") return h->profile;
  if (k == "modw") return h->modw;
  return -1;
}

int mdgen_profile_dump(mdgen_handle* h, char* buf, int64_t cap) {
  if (!h || !buf || cap <= 0) return MDGEN_E_INVALID;
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<double, long long>> acc;
  for (auto& e : h->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e.e0, e.e1) == cudaSuccess) {
      acc[e.name].first += ms;
      acc[e.name].second += 1;
    }
    cudaEventDestroy(e.e0);
    cudaEventDestroy(e.e1);
  }
  h->prof.clear();
  std::string out;
  for (auto& kv : acc) {
    char line[256];
    snprintf(line, sizeof(line), "%s %.4f %lld\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  snprintf(buf, (size_t)cap, "%s", out.c_str());
  return MDGEN_OK;
}

}  // extern "C"
