/*
 * mdgen_b200.h — C ABI of libmdgen_b200.so: the B200-native (sm_100a) MDGen denoiser hot path.
 *
 * The reference (bjing2016/mdgen) is pure Python/PyTorch and defines no FFI; its boundary for this
 * path is the Python surface of `mdgen.wrapper.NewMDGenWrapper`. Each entry point below names the
 * reference routine it replaces (file:line relative to the reference repo root). The host-side
 * mirror of that surface (same class/method names) lives in mdgen_b200/wrapper.py and binds this
 * library through ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - return 0 on success, a negative MDGEN_E_* code on failure (no exceptions cross the ABI);
 *     mdgen_last_error() returns a human readable message for the last failure on the handle
 *     (or, with a NULL handle, for the last failed mdgen_create).
 *   - every `const float*` / `float*` / `const int64_t*` argument is caller-owned CUDA *device*
 *     memory, contiguous row-major in the reference's logical layout, unless marked (host).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream). Calls enqueue work on
 *     it and return without synchronising the device (workspace growth may call cudaMalloc).
 *   - one handle per (device, stream); a handle is not thread-safe across concurrent calls.
 *   - all floating point is fp32 at the boundary. Internally the token GEMMs run on the tensor cores
 *     with bf16 operands (option "gemm_bf16" = 0: TF32 operands), the attention contractions with
 *     fp16 / bf16 operands, all with fp32 accumulation; the IPA key-frame trunk, softmax
 *     denominators, LayerNorm statistics, the residual stream and the Euler state stay fp32.
 */
#ifndef MDGEN_B200_H_
#define MDGEN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDGEN_ABI_VERSION 2

enum {
  MDGEN_OK = 0,
  MDGEN_E_INVALID = -1,     /* bad argument / unsupported configuration */
  MDGEN_E_CUDA = -2,        /* a CUDA runtime/driver call failed */
  MDGEN_E_WEIGHTS = -3,     /* missing / mis-shaped tensor, or weights not finalised */
  MDGEN_E_NOMEM = -4
};

typedef struct mdgen_handle mdgen_handle;

/* Architecture switches, derived from the reference's argparse namespace
 * (mdgen/parsing.py:5-125; latent_dim rule mdgen/wrapper.py:196-200).
 * Fixed by the kernels: embed_dim 384, mha_heads 16, ffn 1536, IPA 4 heads x 32, 8+8 points. */
typedef struct mdgen_config {
  int32_t abi_version;    /* MDGEN_ABI_VERSION */
  int32_t latent_dim;     /* 21 (sim / upsampling / ATLAS) or 28 (tps / inpainting) */
  int32_t num_layers;     /* IPA layers == main layers (args.num_layers, default 5) */
  int32_t crop;           /* rows of pos_embed (args.crop) */
  int32_t abs_pos_emb;    /* --abs_pos_emb */
  int32_t use_aa_emb;     /* !--no_aa_emb */
  int32_t sim_condition;  /* --sim_condition  (single IPA trunk from the start frame) */
  int32_t tps_condition;  /* --tps_condition  (two IPA trunks, both end frames conditioned) */
  int32_t inpainting;     /* --inpainting     (two IPA trunks, residues 0 and 3 conditioned) */
  int32_t cond_interval;  /* --cond_interval  (0 = off) */
  int32_t no_torsion;     /* --no_torsion */
  float time_multiplier;  /* --time_multiplier (100) */
} mdgen_config;

/* Conditioning of one denoiser call == the `model_kwargs` dict built by
 * NewMDGenWrapper.prep_batch (mdgen/wrapper.py:353-365). */
typedef struct mdgen_cond {
  int32_t B, T, L;
  const float* mask;          /* [B,T,L]   1 = real residue, 0 = padding            */
  const float* start_rot;     /* [B,L,3,3] rigids[:,0]  rotation matrices            */
  const float* start_trans;   /* [B,L,3]                                              */
  const float* end_rot;       /* [B,L,3,3] rigids[:,-1] (two-trunk configs; else NULL) */
  const float* end_trans;     /* [B,L,3]                                              */
  const float* x_cond;        /* [B,T,L,latent_dim]                                   */
  const int64_t* x_cond_mask; /* [B,T,L]   0/1                                        */
  const int64_t* aatype;      /* [B,L]     0..20                                      */
  /* Two-trunk configs only, optional (NULL = canonical w >= 0): sign (+1 / -1) applied to the relative
   * quaternions fed to latent_to_emb_r ([0,:,:] = end^-1 o start) and latent_to_emb_f ([1,:,:] =
   * start^-1 o end). The reference keeps whatever sign torch.linalg.eigh returns there
   * (mdgen/model/latent_model.py:194-195 -> mdgen/rigid_utils.py:191-210); a caller that must
   * reproduce it bit-for-sign passes sign(eigh(...)[..., -1][..., 0]) computed on its own backend. */
  const float* quat_sign;     /* [2,B,L]   or NULL                                    */
} mdgen_cond;

/* Lifetime. Replaces LatentMDGenModel.__init__ (mdgen/model/latent_model.py:44-128). */
int mdgen_create(const mdgen_config* cfg, mdgen_handle** out);
void mdgen_destroy(mdgen_handle* h);
const char* mdgen_last_error(const mdgen_handle* h);

/* Weights. `name` is the reference state-dict key without the wrapper's "model." prefix
 * (SURVEY.md Appendix A), `data` a device pointer to its fp32 values, `numel` its element count.
 * The library copies/packs privately (q/k/v concatenated, head_dim^-0.5 folded into W_q/b_q,
 * TF32 round-to-nearest of GEMM operands), so the caller may free `data` after
 * mdgen_finalize_weights. Replaces nn.Module.load_state_dict for this model. */
int mdgen_set_tensor(mdgen_handle* h, const char* name, const float* data, int64_t numel);
int mdgen_finalize_weights(mdgen_handle* h, void* stream);

/* Residue geometry tables used by mdgen_decode_atom14 (host pointers; copied to the device):
 * restype_rigid_group_default_frame [21,8,4,4], restype_atom14_rigid_group_positions [21,14,3],
 * restype_atom14_to_rigid_group [21,14], restype_atom14_mask [21,14]
 * (mdgen/residue_constants.py:1124-1130). */
int mdgen_set_residue_tables(mdgen_handle* h, const float* default_frame /*host*/,
                             const float* atom14_group_pos /*host*/,
                             const int32_t* atom14_to_group /*host*/,
                             const float* atom14_mask /*host*/);
/* Tables of the rollout re-featurisation (host pointers): chi atoms as atom14 indices [21,4,4], their
 * existence mask [21,4,4] (RESTYPE_ATOM37_MASK), chi_angles_mask [21,4], backbone N/CA/C/O mask [21,4]
 * (mdgen/residue_constants.py:33-102,1475-1478; mdgen/geometry.py:337-358). */
int mdgen_set_featurize_tables(mdgen_handle* h, const int32_t* chi_atom14_idx /*host*/,
                               const float* chi_atom_mask /*host*/, const float* chi_mask /*host*/,
                               const float* bb_mask /*host*/);

/* One denoiser evaluation: out[B,T,L,D] = model.forward_inference(x, t, **cond)
 * (mdgen/model/latent_model.py:263-269 -> :212-260). `t` is a device vector [B]. */
int mdgen_forward(mdgen_handle* h, const float* x, const float* t, const mdgen_cond* cond,
                  float* out, void* stream);

/* Fixed-grid Euler sampling, all K steps driven from native code without returning to Python:
 *   x_0 = zs;  x_{k+1} = x_k + (t_grid[k+1]-t_grid[k]) * forward(x_k, t_grid[k]*1_B);  x_out = x_K
 * == transport_sampler.sample_ode('euler', num_steps=K+1)(zs, forward_inference)[-1]
 * (mdgen/transport/transport.py:408-451, mdgen/transport/integrators.py:90-113, torchdiffeq
 * fixed-grid Euler). t_grid is a HOST array of K+1 floats (the float32 linspace the reference
 * builds on the host). x_out may alias zs. */
int mdgen_sample_euler(mdgen_handle* h, const float* zs, const float* t_grid /*host*/, int32_t K,
                       const mdgen_cond* cond, float* x_out, void* stream);

/* Featurisation == NewMDGenWrapper.prep_batch (mdgen/wrapper.py:283-365) + get_offsets
 * (mdgen/utils.py:7-14) + Rigid.invert/compose/to_tensor_7 (mdgen/rigid_utils.py:1075,1031,1143):
 * frame-0 (and frame T-1 for two-trunk configs) relative offsets as [quat(w>=0) | trans], torsions
 * appended, conditioning mask and x_cond = where(mask, latents, 0).
 *   rots [B,T,L,3,3], trans [B,T,L,3], torsions [B,T,L,7,2]
 *   -> latents [B,T,L,D], x_cond [B,T,L,D], x_cond_mask [B,T,L] (int64) */
int mdgen_prep_batch(mdgen_handle* h, int32_t B, int32_t T, int32_t L, const float* rots,
                     const float* trans, const float* torsions, float* latents, float* x_cond,
                     int64_t* x_cond_mask, void* stream);

/* Decode tail of NewMDGenWrapper.inference (mdgen/wrapper.py:456-478) + frames_torsions_to_atom14
 * (mdgen/geometry.py:61-79,236-334): samples [B,T,L,D] + frame-0 rigids + seqres [B,L]
 * -> atom14 [B,T,L,14,3]. */
int mdgen_decode_atom14(mdgen_handle* h, int32_t B, int32_t T, int32_t L, const float* samples,
                        const float* start_rot, const float* start_trans, const int64_t* seqres,
                        float* atom14, void* stream);

/* Rollout re-featurisation on the device == what sim_inference.py:91-96 does on the
 * host between rollouts: frames = atom14_to_frames(atom14) (mdgen/geometry.py:218-231) and torsions =
 * atom37_to_torsions(atom14_to_atom37(atom14, seqres), seqres) (mdgen/geometry.py:9-27, 82-202).
 *   atom14 [B,L,14,3] (one frame), seqres [B,L]
 *   -> rots [B,L,3,3], trans [B,L,3], torsions [B,L,7,2], torsion_mask [B,L,7] (may be NULL) */
int mdgen_featurize_atom14(mdgen_handle* h, int32_t B, int32_t L, const float* atom14, const int64_t* seqres,
                           float* rots, float* trans, float* torsions, float* torsion_mask, void* stream);

/* Training / validation loss pieces (the forward half of Transport.training_losses,
 * mdgen/transport/transport.py:138-223). All device pointers; `per` = elements per sample (T*L*D).
 * Interpolant plan (mdgen/transport/path.py:118-135): xt = alpha(t) x1 + sigma(t) x0, ut = alpha'(t) x1 + sigma'(t) x0
 * with per-sample t [B]; path 0 = GVP (path.py:173-191: sin / cos of pi t / 2), 1 = Linear (alpha = t, sigma = 1 - t). */
int mdgen_flow_plan(mdgen_handle* h, int32_t B, int64_t per, int32_t path, const float* x1, const float* x0,
                    const float* t, float* xt, float* ut, void* stream);
/* loss[b] = mean_flat((pred - target)^2, mask) = sum((pred-target)^2 * mask) / sum(mask) over the non-batch dims
 * (transport.py:13-17, :189). */
int mdgen_masked_mse(mdgen_handle* h, int32_t B, int64_t per, const float* pred, const float* target,
                     const float* mask, float* loss, void* stream);
/* ExponentialMovingAverage.update for one tensor (mdgen/ema.py:41-50): stored -= (stored - param) * (1 - decay). */
int mdgen_ema_update(mdgen_handle* h, float* stored, const float* param, int64_t n, float decay, void* stream);

/* Runge-Kutta plumbing of the adaptive dopri5 sampler (the reference's default sampling_method: torchdiffeq.odeint
 * behind mdgen/transport/integrators.py:106-113). out[n] = (y ? y : 0) + scale * sum_{i<nk} coeffs[i] * ks[i]
 * (nk <= 8; coeffs host floats, ks host array of device pointers): stage, solution, error and dense-output
 * combinations in one pass over the state. */
int mdgen_lincomb(mdgen_handle* h, int64_t n, const float* y, float scale, const float* coeffs /*host*/,
                  const float* const* ks /*host array of device pointers*/, int32_t nk, float* out, void* stream);
/* ratio (host float out) = sqrt(mean((err / (atol + rtol * max(|y0|, |y1|)))^2)): torchdiffeq's mixed-tolerance RMS
 * error norm over the whole state (fixed summation order). Synchronises the stream (the step controller needs it). */
int mdgen_rk_error_ratio(mdgen_handle* h, int64_t n, const float* err, const float* y0, const float* y1,
                         float rtol, float atol, float* ratio /*host*/, void* stream);

/* Introspection for tests / bench. */
int mdgen_abi_version(void);
/* number of kernel launches issued by this handle since creation (bench.py's gpu_launches) */
int64_t mdgen_launch_count(const mdgen_handle* h);
/* Tuning / validation knobs (none of them is needed for normal use; the defaults are the measured-best path):
 *   "use_tc"        1 (default) tcgen05 tensor-core GEMMs and attention, 0 = exact-fp32 SIMT validation kernels
 *   "gemm_bf16"     operand type of the token GEMMs: 2 (default) fp16 (kind::f16), 1 = bf16 (kind::f16), 0 = TF32
 *                   (kind::tf32). bf16 does not meet the 1e-3 parity tolerance on trained-like weights.
 *   "use_tc_attn"   1 (default) tcgen05 fused attention for sequences longer than 64
 *   "attn_variant"  build variant of that attention: 256 (default) generation 8 (csrc/attention_v8.cuh; +1 forces its
 *                   2-query-tile kernel, +2 keeps every exponential on the MUFU pipe); 0-15 (+128) generation 7
 *                   (csrc/attention_tc.cuh: bit 0 bf16 P.V, bit 1 staged pre-pass, bit 2 persistent kernel, bit 3 with
 *                   12 softmax warps, bit 7 bound-adopted softmax reference; bits 4-6 are timing aids, output undefined)
 *   "l4_variant"    1 (default) shared-memory exchange S = 4 residue attention, 0 = shuffle-based kernel
 *   "fuse_resid_ln" 1: the out-proj / fc2 GEMMs store gate * branch and the next LayerNorm kernel adds it to the residual
 *                   stream (measured neutral), 0 (default): residual add in the GEMM epilogue
 *   "tc_min_rows"   GEMMs with fewer rows (the IPA key-frame trunk) stay on the exact-fp32 skinny GEMM (1024)
 *   "use_graph", "graph_max_tokens"   CUDA-graph replay of the Euler steps for launch-bound workloads (1, 65536)
 *   "reuse_cond"    1: mdgen_forward keeps the conditioning embedding of its previous call (same cond object and
 *                   shapes: the stages of one adaptive ODE solve); the caller resets it to 0 afterwards
 *   "emu_bf16"      precision experiments (tools/diag_precision.py);  "profile" 1 = per-family CUDA-event timing
 *   "gemm_dbg"      measurement switches of the tcgen05 GEMMs, OUTPUT UNDEFINED: bit 0 no operand loads, bit 1 no MMAs
 *                   (profiles/r2_gemm_epilogue.md: which of fill / MMA / epilogue bounds a GEMM)
 * Environment (read once per process): MDGEN_NO_TMA_OUT=1 keeps the 16-bit-output GEMMs on the shared-memory-transpose
 * epilogue, MDGEN_NO_WS=1 keeps the K <= 384 GEMMs on the tile-streaming kernel (A/B switches for the same file).
 * Unknown keys return MDGEN_E_INVALID. */
int mdgen_set_option(mdgen_handle* h, const char* key, int64_t value);
int64_t mdgen_get_option(const mdgen_handle* h, const char* key);
/* Accumulated device time (ms) per kernel family since the last reset; fills up to `cap`
 * (name,ms,calls) triples when profiling is enabled with mdgen_set_option(h,"profile",1). */
int mdgen_profile_dump(mdgen_handle* h, char* buf, int64_t cap);

/* Test hook: out[M,N] = act(A[M,K] · W[N,K]^T + bias) through one of the library's GEMM kernels
 * (use_tc = 0: fp32 SIMT kernel; 1: tcgen05 kernel, operands rounded to TF32 first; 2 / 3: bf16 / fp16 operands, fp32 output;
 * 4 / 5: bf16 / fp16 operands and 16-bit output -- the form of the QKV / fc1 GEMMs, bulk-tensor-store epilogue --
 * widened to fp32 into `out`).
 * act: 0 = none, 1 = erf-GELU. Lets the parity tests A/B the tensor-core kernel in isolation. */
int mdgen_debug_linear(mdgen_handle* h, const float* A, const float* W, const float* bias, int64_t M,
                       int32_t N, int32_t K, int32_t act, int32_t use_tc, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDGEN_B200_H_ */
