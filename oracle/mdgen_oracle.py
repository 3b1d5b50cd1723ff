"""CPU restatement of the MDGen hot path — TEST INFRASTRUCTURE (the parity oracle).

Plain PyTorch fp32 on the host, functional over a state dict (keys of
`LatentMDGenModel.state_dict()`), one function per reference routine with its file:line.
It materialises the attention scores exactly like the reference does, so it is also the
honest "port" CPU baseline that bench.py times beside the B200 numbers. The functions are
device-agnostic (tensors stay on the device of their inputs), which lets bench.py also time the
same stock-ATen formulation on the B200 (the north-star's "reference single-GPU PyTorch" figure).

Pinned by tests/golden/*.npz, which were produced by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_loader.py) on the same seeded inputs; see
tests/golden/gen_golden.py and tests/test_oracle_golden.py.

Third-party arithmetic the reference delegates to un-vendored packages (fair-esm rotary
embedding, torchdiffeq fixed-grid Euler and adaptive dopri5) is restated from the packages' published
algorithms: "parity unpinned by the reference" for those pieces. The dopri5 restatement
(`sample_dopri5`, `dopri5_replay`) additionally has no live torchdiffeq to be checked against in this image: its
Butcher tableau is pinned against scipy's independent RK45 tableau and its dense output against the defining
interpolation conditions (tests/test_ode_cpu.py); the step controller follows torchdiffeq 0.2.x from its source.

Never imported by mdgen_b200/ (the product).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

C = 384
H = 16
HD = 24
IPA_H, IPA_C, IPA_PQ, IPA_PV = 4, 32, 8, 8

_TABLES = None


def residue_tables() -> Dict[str, torch.Tensor]:
    global _TABLES
    if _TABLES is None:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                         "mdgen_b200", "data", "residue_tables.npz")
        z = np.load(p)
        _TABLES = {k: torch.from_numpy(z[k]) for k in z.files}
    return _TABLES


# --------------------------------------------------------------------------- rigid algebra
def rot_matmul(a, b):
    """mdgen/rigid_utils.py:24-61 (written out products, fp32)."""
    return torch.einsum("...ij,...jk->...ik", a, b)


def rot_vec_mul(r, v):
    """mdgen/rigid_utils.py:64-86."""
    return torch.einsum("...ij,...j->...i", r, v)


def quat_to_rot(q):
    """mdgen/rigid_utils.py:141-188: standard (w,x,y,z) -> matrix, *no* normalisation."""
    w, x, y, z = q.unbind(-1)
    r = torch.stack([
        w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z,
    ], -1)
    return r.reshape(q.shape[:-1] + (3, 3))


def rot_to_quat(rot):
    """mdgen/rigid_utils.py:191-210: top eigenvector of the symmetric 4x4 K matrix / 3
    (torch.linalg.eigh; eigenvector sign is arbitrary, callers canonicalise w >= 0)."""
    xx, xy, xz = rot[..., 0, 0], rot[..., 0, 1], rot[..., 0, 2]
    yx, yy, yz = rot[..., 1, 0], rot[..., 1, 1], rot[..., 1, 2]
    zx, zy, zz = rot[..., 2, 0], rot[..., 2, 1], rot[..., 2, 2]
    k = torch.stack([
        torch.stack([xx + yy + zz, zy - yz, xz - zx, yx - xy], -1),
        torch.stack([zy - yz, xx - yy - zz, xy + yx, xz + zx], -1),
        torch.stack([xz - zx, xy + yx, yy - xx - zz, yz + zy], -1),
        torch.stack([yx - xy, xz + zx, yz + zy, zz - xx - yy], -1),
    ], -2) * (1.0 / 3.0)
    _, vec = torch.linalg.eigh(k)
    return vec[..., -1]


def rigid_invert(R, t):
    """mdgen/rigid_utils.py:1075-1085: (R^T, -R^T t)."""
    Rt = R.transpose(-1, -2)
    return Rt, -rot_vec_mul(Rt, t)


def rigid_compose(Ra, ta, Rb, tb):
    """mdgen/rigid_utils.py:1031-1045: (Ra Rb, Ra tb + ta)."""
    return rot_matmul(Ra, Rb), rot_vec_mul(Ra, tb) + ta


def to_tensor_7(R, t, canonical: bool = True):
    """mdgen/rigid_utils.py:1143-1156 + the w>=0 fix of mdgen/wrapper.py:309."""
    q = rot_to_quat(R)
    if canonical:
        q = q * torch.where(q[..., 0:1] < 0, -1.0, 1.0)
    return torch.cat([q, t], -1)


def get_offsets(R0, t0, R, t):
    """mdgen/utils.py:7-14: ref.invert().compose(rigids).to_tensor_7(), w>=0
    (mdgen/wrapper.py:307-309)."""
    Ri, ti = rigid_invert(R0, t0)
    Ro, to = rigid_compose(Ri, ti, R, t)
    return to_tensor_7(Ro, to)


# --------------------------------------------------------------------------- small pieces
def timestep_embedding(t, dim=256, max_period=10000):
    """mdgen/model/layers.py:30-50."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], -1)


def t_embedder(sd, t):
    """mdgen/model/layers.py:52-55 (Linear 256->C, SiLU, Linear C->C)."""
    h = F.linear(timestep_embedding(t), sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    return F.linear(F.silu(h), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])


def adaln(sd, prefix, temb):
    """Sequential(SiLU, Linear) — mdgen/model/latent_model.py:346-349,405-408; layers.py:65-68."""
    return F.linear(F.silu(temb), sd[prefix + "adaLN_modulation.1.weight"],
                    sd[prefix + "adaLN_modulation.1.bias"])


def layer_norm0(x):
    """no-affine LayerNorm, eps 1e-6 (mdgen/model/latent_model.py:362,367,439,444)."""
    return F.layer_norm(x, (x.shape[-1],), None, None, 1e-6)


def gelu(x):
    """erf GELU — mdgen/model/layers.py:77-84."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def rope_tables(n: int, inv_freq: torch.Tensor):
    """fair-esm RotaryEmbedding tables for n positions (see oracle/ref_shims/esm)."""
    t = torch.arange(n, dtype=torch.float32, device=inv_freq.device)
    freqs = torch.outer(t, inv_freq.float())
    emb = torch.cat([freqs, freqs], -1)
    return emb.cos(), emb.sin()


def rotate_half(x):
    x1, x2 = x.chunk(2, -1)
    return torch.cat([-x2, x1], -1)


def mha(sd, prefix, x, mask):
    """AttentionWithRoPE + MultiheadAttention.forward, batch-first restatement of
    mdgen/model/latent_model.py:325-329 and mdgen/model/mha.py:260-397.
    x [Bp,S,C], mask [Bp,S] (1 = real token)."""
    Bp, S, _ = x.shape
    p = prefix + "attn."
    q = F.linear(x, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"]) * (HD ** -0.5)   # :260-263
    k = F.linear(x, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(x, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    k = torch.cat([k, sd[p + "bias_k"].reshape(1, 1, C).expand(Bp, 1, C)], 1)        # :265-268
    v = torch.cat([v, sd[p + "bias_v"].reshape(1, 1, C).expand(Bp, 1, C)], 1)
    pad = torch.cat([1 - mask, mask.new_zeros(Bp, 1)], 1).to(torch.bool)             # :273-280
    q = q.reshape(Bp, S, H, HD).transpose(1, 2)                                       # :282-286
    k = k.reshape(Bp, S + 1, H, HD).transpose(1, 2)
    v = v.reshape(Bp, S + 1, H, HD).transpose(1, 2)
    cos, sin = rope_tables(S + 1, sd[p + "rot_emb.inv_freq"])                         # :356-357
    q = q * cos[:S] + rotate_half(q) * sin[:S]
    k = k * cos + rotate_half(k) * sin
    w = torch.matmul(q, k.transpose(-1, -2))                                          # :359
    w = w.masked_fill(pad[:, None, None, :], float("-inf"))                           # :370-376
    w = F.softmax(w, dim=-1, dtype=torch.float32)                                     # :381
    o = torch.matmul(w, v).transpose(1, 2).reshape(Bp, S, C)                          # :389-396
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])           # :397


def ipa(sd, prefix, s, R, t, mask):
    """InvariantPointAttention.forward with c_z = 0 — mdgen/model/ipa.py:113-255.
    s [B,L,C]; frames R [B,L,3,3], t [B,L,3]; mask [B,L]."""
    B, L, _ = s.shape
    p = prefix + "ipa."
    q = F.linear(s, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"]).view(B, L, IPA_H, IPA_C)
    kv = F.linear(s, sd[p + "linear_kv.weight"], sd[p + "linear_kv.bias"]).view(B, L, IPA_H, 2 * IPA_C)
    k, v = kv[..., :IPA_C], kv[..., IPA_C:]                                           # :123
    qp = F.linear(s, sd[p + "linear_q_points.weight"], sd[p + "linear_q_points.bias"])
    qp = torch.stack(torch.split(qp, qp.shape[-1] // 3, -1), -1)                      # :130-131 coord-major
    qp = rot_vec_mul(R[:, :, None], qp) + t[:, :, None]                               # :132 r.apply
    qp = qp.view(B, L, IPA_H, IPA_PQ, 3)
    kvp = F.linear(s, sd[p + "linear_kv_points.weight"], sd[p + "linear_kv_points.bias"])
    kvp = torch.stack(torch.split(kvp, kvp.shape[-1] // 3, -1), -1)                   # :141-142
    kvp = rot_vec_mul(R[:, :, None], kvp) + t[:, :, None]
    kvp = kvp.view(B, L, IPA_H, IPA_PQ + IPA_PV, 3)
    kp, vp = kvp[..., :IPA_PQ, :], kvp[..., IPA_PQ:, :]                               # :149-151
    a = torch.einsum("bihc,bjhc->bhij", q, k) * math.sqrt(1.0 / (3 * IPA_C))          # :161-166
    d2 = ((qp[:, :, None] - kp[:, None]) ** 2).sum(-1)                                # :171-175 [B,i,j,H,P]
    hw = F.softplus(sd[p + "head_weights"]) * math.sqrt(1.0 / (3 * (IPA_PQ * 9.0 / 2)))  # :176-181
    pt = (d2 * hw.view(1, 1, 1, IPA_H, 1)).sum(-1) * (-0.5)                           # :182-185
    sq = 1e5 * (mask[:, :, None] * mask[:, None, :] - 1)                              # :188-190
    a = a + pt.permute(0, 3, 1, 2) + sq[:, None]                                      # :193-198
    a = F.softmax(a, -1)                                                              # :203
    o = torch.einsum("bhij,bjhc->bihc", a, v).reshape(B, L, IPA_H * IPA_C)            # :209-212
    op = torch.einsum("bhij,bjhpx->bihpx", a, vp)                                     # :216-225
    op = rot_vec_mul(R.transpose(-1, -2)[:, :, None, None], op - t[:, :, None, None])  # :226 invert_apply
    opn = torch.sqrt((op ** 2).sum(-1) + 1e-8).reshape(B, L, IPA_H * IPA_PV)          # :229-231
    op = op.reshape(B, L, IPA_H * IPA_PV, 3)
    cat = torch.cat([o, op[..., 0], op[..., 1], op[..., 2], opn], -1)                 # :250-251
    return F.linear(cat, sd[p + "linear_out.weight"], sd[p + "linear_out.bias"])


def ipa_layer(sd, prefix, x, temb, mask, R, t):
    """IPALayer.forward — mdgen/model/latent_model.py:369-384. x [B,L,C], temb [B,C]."""
    sh_l, sc_l, g_l, sh_m, sc_m, g_m = adaln(sd, prefix, temb).chunk(6, -1)
    xn = F.layer_norm(x, (C,), sd[prefix + "ipa_norm.weight"], sd[prefix + "ipa_norm.bias"], 1e-5)
    x = x + ipa(sd, prefix, xn, R, t, mask)                                           # :372
    res = x
    h = layer_norm0(x) * (1 + sc_l[:, None]) + sh_l[:, None]                          # :375
    x = res + g_l[:, None] * mha(sd, prefix + "mha_l.", h, mask)                      # :376-377
    res = x
    h = layer_norm0(x) * (1 + sc_m[:, None]) + sh_m[:, None]                          # :380
    h = F.linear(gelu(F.linear(h, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"])),
                 sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"])                  # :381
    return res + g_m[:, None] * h                                                     # :382


def mdgen_layer(sd, prefix, x, temb, mask):
    """LatentMDGenLayer.forward — mdgen/model/latent_model.py:446-483.
    x [B,T,L,C], temb [B,C], mask [B,T,L]."""
    B, T, L, _ = x.shape
    m = adaln(sd, prefix, temb)[:, None, None, :].chunk(9, -1)                        # :449-451
    sh_l, sc_l, g_l, sh_t, sc_t, g_t, sh_m, sc_m, g_m = m
    res = x
    h = layer_norm0(x) * (1 + sc_l) + sh_l                                            # :457
    h = mha(sd, prefix + "mha_l.", h.reshape(B * T, L, C), mask.reshape(B * T, L)).reshape(B, T, L, C)
    x = res + g_l * h                                                                 # :462
    res = x
    h = layer_norm0(x) * (1 + sc_t) + sh_t                                            # :465
    h = mha(sd, prefix + "mha_t.", h.transpose(1, 2).reshape(B * L, T, C),
            mask.transpose(1, 2).reshape(B * L, T)).reshape(B, L, T, C).transpose(1, 2)  # :472-475
    x = res + g_t * h                                                                 # :476
    res = x
    h = layer_norm0(x) * (1 + sc_m) + sh_m                                            # :479
    h = F.linear(gelu(F.linear(h, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"])),
                 sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"])                  # :480
    return res + g_m * h                                                              # :481


# The reference leaves torch.linalg.eigh's arbitrary eigenvector sign on the relative quaternions of the
# two-trunk branch (mdgen/model/latent_model.py:194-195). True: canonicalise w >= 0 (deterministic across
# LAPACK / cuSOLVER builds; the product's default "canonical" mode). False: keep eigh's sign, i.e. the
# unmodified reference on this backend (the product's quat_sign_mode "eigh" / "eigh_cpu").
CANONICAL_TPS_QUAT = True


def run_ipa(sd, cfg, temb, mask_bl, start, end, aatype):
    """LatentMDGenModel.run_ipa — mdgen/model/latent_model.py:175-210.
    start/end = (R [B,L,3,3], t [B,L,3]). See CANONICAL_TPS_QUAT for the sign of the tps-branch quaternions."""
    B, L = mask_bl.shape
    n = cfg.num_layers
    if cfg.sim_condition:
        x = torch.zeros(B, L, C, device=mask_bl.device)
        if aatype is not None and cfg.use_aa_emb:
            x = x + sd["aatype_to_emb.weight"][aatype]                                # :188
        for i in range(n):
            x = ipa_layer(sd, f"ipa_layers.{i}.", x, temb, mask_bl, *start)           # :191-192
        return x
    Rs, ts = start
    Re, te = end
    x_f = to_tensor_7(*rigid_compose(*rigid_invert(Rs, ts), Re, te), canonical=CANONICAL_TPS_QUAT)   # :194
    x_r = to_tensor_7(*rigid_compose(*rigid_invert(Re, te), Rs, ts), canonical=CANONICAL_TPS_QUAT)   # :195
    x_f = F.linear(x_f, sd["latent_to_emb_f.weight"], sd["latent_to_emb_f.bias"])
    x_r = F.linear(x_r, sd["latent_to_emb_r.weight"], sd["latent_to_emb_r.bias"])
    if aatype is not None and cfg.use_aa_emb:
        x_f = x_f + sd["aatype_to_emb.weight"][aatype]
        x_r = x_r + sd["aatype_to_emb.weight"][aatype]
    for i in range(n):
        x_r = ipa_layer(sd, f"ipa_layers.{i}.", x_r, temb, mask_bl, Rs, ts)           # :205
        x_f = ipa_layer(sd, f"ipa_layers.{i}.", x_f, temb, mask_bl, Re, te)           # :206
    return x_r + x_f


def forward(sd, cfg, x, t, mask, start, end, x_cond, x_cond_mask, aatype):
    """LatentMDGenModel.forward (non-design) — mdgen/model/latent_model.py:212-260.
    x [B,T,L,D], t [B], mask [B,T,L], x_cond [B,T,L,D], x_cond_mask [B,T,L] int64."""
    h = F.linear(x, sd["latent_to_emb.weight"], sd["latent_to_emb.bias"])             # :233
    if cfg.abs_pos_emb:
        h = h + sd["pos_embed"]                                                       # :235
    if x_cond is not None:
        h = h + F.linear(x_cond, sd["cond_to_emb.weight"], sd["cond_to_emb.bias"]) \
            + sd["mask_to_emb.weight"][x_cond_mask]                                   # :241
    temb = t_embedder(sd, t * cfg.time_multiplier)                                    # :243
    h = h + run_ipa(sd, cfg, temb, mask[:, 0], start, end, aatype)[:, None]           # :246
    for i in range(cfg.num_layers):
        h = mdgen_layer(sd, f"layers.{i}.", h, temb, mask)                            # :248-249
    sh, sc = adaln(sd, "emb_to_latent.", temb)[:, None, None, :].chunk(2, -1)         # layers.py:71
    h = layer_norm0(h) * (1 + sc) + sh
    return F.linear(h, sd["emb_to_latent.linear.weight"], sd["emb_to_latent.linear.bias"])


def sample_euler(sd, cfg, zs, t_grid, **kw):
    """Sampler.sample_ode('euler') -> ode.sample -> torchdiffeq fixed-grid Euler, last state only
    (mdgen/transport/transport.py:408-451, mdgen/transport/integrators.py:90-113)."""
    x = zs
    B = zs.shape[0]
    for i in range(len(t_grid) - 1):
        t0, t1 = t_grid[i], t_grid[i + 1]
        tv = torch.ones(B, device=zs.device) * t0.to(zs.device)                       # integrators.py:99
        x = x + (t1 - t0).to(zs.device) * forward(sd, cfg, x, tv, **kw)
    return x


# --------------------------------------------------------------------------- adaptive sampler (default of the reference)
# torchdiffeq.odeint(..., method='dopri5') as called by mdgen/transport/integrators.py:106-113 with
# rtol 1e-3 / atol 1e-6 (mdgen/transport/transport.py:411-414). torchdiffeq 0.2.x, _impl/dopri5.py:
_DP_ALPHA = torch.tensor([1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0], dtype=torch.float64)
_DP_BETA = torch.zeros(6, 7, dtype=torch.float64)
for _i, _row in enumerate([[1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9],
                           [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
                           [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
                           [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]]):
    _DP_BETA[_i, :len(_row)] = torch.tensor(_row, dtype=torch.float64)
_DP_CSOL = torch.tensor([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0], dtype=torch.float64)
_DP_CERR = torch.tensor([35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
                         -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1 / 60], dtype=torch.float64)
_DP_CMID = torch.tensor([6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2,
                         -2691868925 / 45128329728 / 2, 187940372067 / 1594534317056 / 2,
                         -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2], dtype=torch.float64)


def _dp_step(func, t0, dt, y0, f0):
    """torchdiffeq _impl/rk_common.py:_runge_kutta_step on the Dormand-Prince-Shampine tableau: the stage
    matrix k[..., i] is filled left to right, every stage state is y0 + k @ (beta_i * dt)."""
    k = torch.zeros(*y0.shape, 7, dtype=y0.dtype, device=y0.device)
    k[..., 0] = f0
    yi = y0
    for i in range(6):
        yi = y0 + k @ (_DP_BETA[i] * dt).to(y0.dtype)
        k[..., i + 1] = func(t0 + float(_DP_ALPHA[i]) * dt, yi)
    y1 = yi                                             # c_sol[:-1] == beta[-1], c_sol[-1] == 0 (FSAL)
    f1 = k[..., -1]
    err = k @ (_DP_CERR * dt).to(y0.dtype)
    y_mid = y0 + k @ (_DP_CMID * dt).to(y0.dtype)
    return y1, f1, err, y_mid


def _dp_interp(y0, y1, y_mid, f0, f1, dt, x):
    """torchdiffeq _impl/interp.py: _interp_fit + _interp_evaluate at x = (t - t0) / dt."""
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    return y0 + x * (d + x * (c + x * (b + x * a)))


def _rms(x):
    return float(x.abs().pow(2).mean().sqrt())         # torchdiffeq _impl/misc.py:_rms_norm


def dopri5_solve(func, y0, t_end, rtol=1e-3, atol=1e-6, t_start=0.0):
    """Adaptive solve from t_start to t_end (torchdiffeq _impl/rk_common.py:RKAdaptiveStepsizeODESolver:
    _before_integrate, _advance, _adaptive_step; misc.py:_select_initial_step, _compute_error_ratio,
    _optimal_step_size). Returns (y(t_end), accepted steps [(t0, dt)], nfe)."""
    f0 = func(t_start, y0)
    nfe = 1
    scale = atol + y0.abs() * rtol                                                    # _select_initial_step, order = 4
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f_probe = func(t_start + h0, y0 + h0 * f0)
    nfe += 1
    d2 = _rms((f_probe - f0) / scale) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / 5.0)
    dt = min(100 * h0, h1)
    t, y, f = t_start, y0, f0
    last = None
    steps = []
    while t_end > t:                                                                  # _advance: no clipping to t_end
        y1, f1, err, y_mid = _dp_step(func, t, dt, y, f)
        nfe += 6
        ratio = _rms(err / (atol + rtol * torch.max(y.abs(), y1.abs())))              # _compute_error_ratio
        if ratio <= 1:                                                                # accept
            last = (t, dt, y, y1, y_mid, f, f1)
            steps.append((t, dt))
            t, y, f = t + dt, y1, f1
        if ratio == 0:                                                                # _optimal_step_size
            dt = dt * 10.0
        else:
            dt = dt * min(10.0, max(0.9 / ratio ** 0.2, 1.0 if ratio < 1 else 0.2))
    t0, h, ya, yb, ym, fa, fb = last
    return _dp_interp(ya, yb, ym, fa, fb, h, (t_end - t0) / h), steps, nfe


def dopri5_replay(func, y0, steps, t_end):
    """The same Dormand-Prince steps with a GIVEN accepted step sequence [(t0, dt)] (no controller) and the
    dense output at t_end inside the last step: used to compare two right-hand sides (oracle vs CUDA forward)
    without the accept/reject decisions amplifying rounding differences."""
    y = y0
    f = func(steps[0][0], y0)
    last = None
    for t0, dt in steps:
        y1, f1, _, y_mid = _dp_step(func, t0, dt, y, f)
        last = (t0, dt, y, y1, y_mid, f, f1)
        y, f = y1, f1
    t0, h, ya, yb, ym, fa, fb = last
    return _dp_interp(ya, yb, ym, fa, fb, h, (t_end - t0) / h)


def _rhs(sd, cfg, kw):
    def func(t, x):
        tv = torch.ones(x.shape[0], device=x.device) * float(t)                       # integrators.py:98-101
        return forward(sd, cfg, x, tv, **kw)
    return func


def sample_dopri5(sd, cfg, zs, rtol=1e-3, atol=1e-6, **kw):
    """Sampler.sample_ode('dopri5') (the reference's default sampling_method, mdgen/parsing.py:102): last state,
    accepted steps and number of model evaluations."""
    return dopri5_solve(_rhs(sd, cfg, kw), zs, 1.0, rtol=rtol, atol=atol)


def sample_dopri5_replay(sd, cfg, zs, steps, **kw):
    return dopri5_replay(_rhs(sd, cfg, kw), zs, steps, 1.0)


# --------------------------------------------------------------------------- wrapper pieces
def prep_batch(cfg, batch):
    """NewMDGenWrapper.prep_batch — mdgen/wrapper.py:283-365 (supported flag subset)."""
    R, tr = batch["rots"], batch["trans"]
    B, T, L = tr.shape[:3]
    off = get_offsets(R[:, 0:1], tr[:, 0:1], R, tr)                                   # :307-309
    if cfg.tps_condition or cfg.inpainting:
        off_r = get_offsets(R[:, -1:], tr[:, -1:], R, tr)                             # :315-317
        off = torch.cat([off, off_r], -1)
    tors = batch["torsions"].reshape(B, T, L, 14)
    if getattr(cfg, "no_torsion", False):
        tors = torch.zeros_like(tors)
    latents = torch.cat([off, tors], -1)                                              # :327
    cond_mask = torch.zeros(B, T, L, dtype=torch.int64, device=tr.device)
    if cfg.sim_condition:
        cond_mask[:, 0] = 1
    if cfg.tps_condition:
        cond_mask[:, 0] = 1
        cond_mask[:, -1] = 1
    if cfg.cond_interval:
        cond_mask[:, ::cfg.cond_interval] = 1
    if cfg.inpainting:
        cond_mask[:, :, [0, 3]] = 1                                                   # COND_IDX :42,346
    nfr = 14 if (cfg.tps_condition or cfg.inpainting) else 7
    frame_lm = batch["mask"][..., None].expand(-1, -1, nfr)                          # :311, :318
    tors_lm = batch["torsion_mask"][..., None].expand(-1, -1, -1, 2).reshape(B, L, 14)   # :312
    loss_mask = torch.cat([frame_lm, tors_lm], -1)[:, None].expand(-1, T, -1, -1)     # :334-335
    return {
        "latents": latents,
        "loss_mask": loss_mask,
        "start": (R[:, 0], tr[:, 0]),
        "end": (R[:, -1], tr[:, -1]),
        "mask": batch["mask"][:, None].expand(-1, T, -1),
        "aatype": batch["seqres"],
        "x_cond": torch.where(cond_mask[..., None].bool(), latents, torch.zeros((), device=tr.device)),
        "x_cond_mask": cond_mask,
    }


def torsion_angles_to_frames(R, t, alpha, aatype, tb):
    """mdgen/geometry.py:273-334. R [...,3,3], t [...,3], alpha [...,7,2], aatype [...]."""
    d44 = tb["default_frame"][aatype]                                                 # [...,8,4,4]
    dR, dt = d44[..., :3, :3], d44[..., :3, 3]
    bb = alpha.new_zeros(alpha.shape[:-2] + (1, 2))
    bb[..., 1] = 1
    a = torch.cat([bb, alpha], -2)                                                    # [...,8,2]
    rot = a.new_zeros(a.shape[:-1] + (3, 3))
    rot[..., 0, 0] = 1
    rot[..., 1, 1] = a[..., 1]
    rot[..., 1, 2] = -a[..., 0]
    rot[..., 2, 1] = a[..., 0]
    rot[..., 2, 2] = a[..., 1]
    fR = rot_matmul(dR, rot)                                                          # default_r.compose(all_rots)
    ft = dt
    R5, t5 = fR[..., 5, :, :], ft[..., 5, :]
    R6, t6 = fR[..., 6, :, :], ft[..., 6, :]
    R7, t7 = fR[..., 7, :, :], ft[..., 7, :]
    c1 = (fR[..., 4, :, :], ft[..., 4, :])
    c2 = rigid_compose(*c1, R5, t5)
    c3 = rigid_compose(*c2, R6, t6)
    c4 = rigid_compose(*c3, R7, t7)
    aR = torch.cat([fR[..., :5, :, :], c2[0][..., None, :, :], c3[0][..., None, :, :],
                    c4[0][..., None, :, :]], -3)
    at = torch.cat([ft[..., :5, :], c2[1][..., None, :], c3[1][..., None, :], c4[1][..., None, :]], -2)
    return rigid_compose(R[..., None, :, :], t[..., None, :], aR, at)                 # r[...,None].compose


def frames_to_atom14(gR, gt, aatype, tb):
    """mdgen/geometry.py:236-270."""
    gi = tb["atom14_to_group"][aatype].long()                                         # [...,14]
    oh = F.one_hot(gi, 8).to(gR.dtype)                                                # [...,14,8]
    aR = torch.einsum("...ag,...gij->...aij", oh, gR)
    at = torch.einsum("...ag,...gi->...ai", oh, gt)
    lit = tb["atom14_group_pos"][aatype]
    pos = rot_vec_mul(aR, lit) + at
    return pos * tb["atom14_mask"][aatype][..., None]


def decode_atom14(cfg, samples, R0, t0, seqres):
    """Tail of NewMDGenWrapper.inference — mdgen/wrapper.py:456-478 + mdgen/geometry.py:61-79.
    samples [B,T,L,D]; R0 [B,L,3,3], t0 [B,L,3] = frame-0 rigids; seqres [B,L]."""
    B, T, L, _ = samples.shape
    off = samples[..., :7]
    tors = samples[..., 14:28] if (cfg.tps_condition or cfg.inpainting) else samples[..., 7:21]
    q = off[..., :4]
    q = q / torch.sqrt((q ** 2).sum(-1, keepdim=True))                               # rigid_utils.py:324-325
    fR, ft = rigid_compose(R0[:, None], t0[:, None], quat_to_rot(q), off[..., 4:7])   # :469
    tors = tors.reshape(B, T, L, 7, 2)
    tors = tors / torch.linalg.norm(tors, dim=-1, keepdim=True)                       # :476
    aat = seqres[:, None].expand(B, T, L)
    tb = residue_tables()
    gR, gt = torsion_angles_to_frames(fR, ft, tors, aat, tb)
    return frames_to_atom14(gR, gt, aat, tb)


# --------------------------------------------------------------------------- rollout re-featurisation
def from_3_points(p_neg_x, origin, p_xy, eps=1e-8):
    """Rigid.from_3_points — mdgen/rigid_utils.py:1176-1216 (Gram-Schmidt; columns e0,e1,e2)."""
    e0 = origin - p_neg_x
    e1 = p_xy - origin
    e0 = e0 / torch.sqrt((e0 * e0).sum(-1, keepdim=True) + eps)
    dot = (e0 * e1).sum(-1, keepdim=True)
    e1 = e1 - e0 * dot
    e1 = e1 / torch.sqrt((e1 * e1).sum(-1, keepdim=True) + eps)
    e2 = torch.cross(e0, e1, dim=-1)
    return torch.stack([e0, e1, e2], -1), origin          # R[..., i, j] = e_j[i]


def featurize_atom14(atom14, seqres):
    """What sim_inference.py:91-96 does between rollouts, for every batch row:
      frames   = atom14_to_frames(atom14)                       (mdgen/geometry.py:218-231)
      torsions = atom37_to_torsions(atom14_to_atom37(atom14))   (mdgen/geometry.py:9-27, 82-202)
    atom14 [B,L,14,3], seqres [B,L] -> rots [B,L,3,3], trans [B,L,3], torsions [B,L,7,2], mask [B,L,7]."""
    tb = residue_tables()
    B, L = seqres.shape
    # frames from (C, CA, N), then compose with diag(-1, 1, -1)   geometry.py:219-231
    R, t = from_3_points(atom14[..., 2, :], atom14[..., 1, :], atom14[..., 0, :])
    R = R * torch.tensor([-1.0, 1.0, -1.0])[None, None, None, :]
    # backbone atoms as atom37 sees them (absent atoms are zeroed by RESTYPE_ATOM37_MASK)
    bbm = tb["bb_mask"][seqres]                                     # [B,L,4]  N, CA, C, O
    bb = atom14[..., :4, :] * bbm[..., None]
    prev = torch.cat([torch.zeros_like(bb[:, :1]), bb[:, :-1]], 1)  # geometry.py:95-98
    prevm = torch.cat([torch.zeros_like(bbm[:, :1]), bbm[:, :-1]], 1)
    pre_omega = torch.stack([prev[..., 1, :], prev[..., 2, :], bb[..., 0, :], bb[..., 1, :]], -2)   # :103-106
    phi = torch.stack([prev[..., 2, :], bb[..., 0, :], bb[..., 1, :], bb[..., 2, :]], -2)            # :107-110
    psi = torch.stack([bb[..., 0, :], bb[..., 1, :], bb[..., 2, :], bb[..., 3, :]], -2)              # :111-114
    pre_omega_m = prevm[..., 1] * prevm[..., 2] * bbm[..., 0] * bbm[..., 1]                          # :116-118
    phi_m = prevm[..., 2] * bbm[..., 0] * bbm[..., 1] * bbm[..., 2]
    psi_m = bbm[..., 0] * bbm[..., 1] * bbm[..., 2] * bbm[..., 3]
    idx = tb["chi_atom14_idx"][seqres].long()                                                        # [B,L,4,4]
    am = tb["chi_atom_mask"][seqres]
    chis = torch.gather(atom14[:, :, None].expand(B, L, 4, 14, 3), 3,
                        idx[..., None].expand(B, L, 4, 4, 3)) * am[..., None]                       # :128-133
    chis_m = tb["chi_mask"][seqres] * am.prod(-1)                                                    # :135-150
    pos = torch.cat([pre_omega[:, :, None], phi[:, :, None], psi[:, :, None], chis], 2)              # [B,L,7,4,3]
    tmask = torch.cat([pre_omega_m[..., None], phi_m[..., None], psi_m[..., None], chis_m], -1)
    fR, ft = from_3_points(pos[..., 1, :], pos[..., 2, :], pos[..., 0, :])                           # :172-177
    fRt = fR.transpose(-1, -2)                                                                       # :179 invert().apply():
    rel = rot_vec_mul(fRt, pos[..., 3, :]) + (-rot_vec_mul(fRt, ft))                                 # R^T p + (-(R^T o))
    sc = torch.stack([rel[..., 2], rel[..., 1]], -1)                                                 # :181-183
    sc = sc / torch.sqrt((sc * sc).sum(-1, keepdim=True) + 1e-8)                                     # :185-194
    sc = sc * torch.tensor([1.0, 1.0, -1.0, 1.0, 1.0, 1.0, 1.0])[None, None, :, None]                # :196-201
    return R, t, sc, tmask


# ---------------------------------------------------------------------------------------------------------
# Training / validation loss (forward half of the training step)
def flow_plan(t, x0, x1, path_type="GVP"):
    """ICPlan.plan — mdgen/transport/path.py:118-135 with GVPCPlan (:173-191) or the linear ICPlan (:68-96)."""
    import math
    tb = t.reshape(-1, *([1] * (x1.dim() - 1)))
    if path_type == "GVP":
        a, s = torch.sin(tb * math.pi / 2), torch.cos(tb * math.pi / 2)
        da, ds = math.pi / 2 * torch.cos(tb * math.pi / 2), -math.pi / 2 * torch.sin(tb * math.pi / 2)
    else:
        a, s, da, ds = tb, 1 - tb, torch.ones_like(tb), -torch.ones_like(tb)
    return a * x1 + s * x0, da * x1 + ds * x0


def mean_flat(x, mask):
    """mdgen/transport/transport.py:13-17."""
    dims = list(range(1, x.dim()))
    return torch.sum(x * mask, dim=dims) / torch.sum(mask, dim=dims)


def training_losses(sd, cfg, x1, loss_mask, t, x0, path_type="GVP", **kw):
    """Transport.training_losses (non-design) — mdgen/transport/transport.py:138-189 with the random draws (t, x0) of
    Transport.sample (:126-136) passed in. Returns (loss [B], pred)."""
    xt, ut = flow_plan(t, x0, x1, path_type)
    pred = forward(sd, cfg, xt, t, **kw)
    return mean_flat((pred - ut) ** 2, loss_mask), pred
