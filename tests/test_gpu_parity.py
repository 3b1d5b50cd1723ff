"""GPU parity tests proper: the CUDA path (through the C ABI, via the Python mirror of the
reference surface) against (a) golden vectors produced by the unmodified reference and (b) the
CPU oracle at other shapes. Tolerance: BASELINE.json north_star — 1e-3 relative fp32."""
import pytest
import torch

from mdgen_b200.synthetic import (euler_time_grid, synthetic_batch, synthetic_noise,
                                  synthetic_state_dict)
from tests.golden.cases import CASES
from tests.helpers import load_case, max_rel, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-3          # north_star: "within 1e-3 relative fp32"
TOL_SIMT = 1e-4     # fp32 SIMT validation kernels: re-association noise only
TOL_GEOM = 2e-5     # prep / decode kernels are exact fp32 geometry


# fp32 SIMT validation kernels | tcgen05 TF32 operands | tcgen05 bf16 operands | tcgen05 fp16 operands (default)
PATHS = ["simt", "tf32", "bf16", "fp16"]
GEMM_DTYPE = {"tf32": 0, "bf16": 1, "fp16": 2}
# "Trained-like" stress weights (sharp softmax rows, O(1) adaLN gates: tests/golden/cases.py) amplify operand
# rounding ~50x compared with the benign bench weights. Measured on B200 against the reference's golden velocity
# (tools/diag_stress.py, max-rel): simt 2e-6 | fp16 operands (default) 7.9e-4 | TF32 operands 0.7-1.2e-3 |
# bf16 operands 5.6e-3. Only the default fp16 path has to meet the north-star 1e-3 there; bf16 cannot (8-bit
# significand), which is why it is no longer the default; the bounds below pin the measured behaviour.
STRESS_TOL = {"simt": TOL_SIMT, "fp16": TOL, "tf32": 2e-3, "bf16": 1e-2}


def _wrapper(args, sd, path):
    from mdgen_b200.wrapper import NewMDGenWrapper
    m = NewMDGenWrapper(args)
    m.model.load_state_dict(sd)
    m = m.eval().to("cuda")
    from mdgen_b200._lib import MDGenError
    eng = m.model.engine()
    try:
        eng.set_option("use_tc", 0 if path == "simt" else 1)
    except MDGenError as e:
        pytest.skip(f"tensor-core kernels unavailable in this build: {e}")
    if path != "simt":
        eng.set_option("gemm_bf16", GEMM_DTYPE[path])
        # run the token GEMMs on the tensor cores even for the tiny test shapes; the IPA key-frame
        # trunk (<= 64 rows here) stays on its production fp32 path
        eng.set_option("tc_min_rows", 65)
    return m


def _dev(batch):
    return {k: v.cuda() for k, v in batch.items()}


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", list(CASES))
def test_golden_cases(name, path):
    case, args, cfg, sd, batch, zs, g = load_case(name)
    args.sampling_method = "euler"
    m = _wrapper(args, sd, path)
    tol = TOL_SIMT if path == "simt" else TOL
    if case.get("stress"):
        tol = STRESS_TOL[path]
    db = _dev(batch)
    prep = m.prep_batch(db)
    kw = prep["model_kwargs"]
    assert max_rel(prep["latents"].cpu(), g["latents"]) < TOL_GEOM
    assert max_rel(kw["x_cond"].cpu(), g["x_cond"]) < TOL_GEOM
    assert (kw["x_cond_mask"].cpu().numpy() == g["x_cond_mask"]).all()
    v = m.model.forward_inference(zs.cuda(), torch.tensor(case["t_fwd"]).cuda(), **kw)
    assert max_rel(v.cpu(), g["v"]) < tol, ("forward", max_rel(v.cpu(), g["v"]))
    xk = m.model.sample_euler(zs.cuda(), euler_time_grid(case["K"]), **kw)
    assert max_rel(xk.cpu(), g["x_euler"]) < tol, ("euler", max_rel(xk.cpu(), g["x_euler"]))
    assert rel_l2(xk.cpu(), g["x_euler"]) < tol
    # decode tail on the reference's 49-step state
    eng = m.model.engine()
    a = eng.decode_atom14(torch.from_numpy(g["x49"]).cuda(), db["rots"][:, 0], db["trans"][:, 0],
                          db["seqres"])
    assert max_rel(a.cpu(), g["atom14"]) < TOL_GEOM
    # public API end to end (49 Euler steps, seeded noise) == reference inference()
    # The decode tail divides torsion (sin,cos) pairs by their norm with no epsilon
    # (mdgen/wrapper.py:476); with random weights some sampled pairs have tiny norms, which amplifies
    # the (in-tolerance) state error for the few side-chain atoms they place. So the end-to-end atom14
    # check is in rel-L2 at the north-star tolerance, with a looser bound on the single worst atom;
    # the ODE state (above) and the decode kernel on identical inputs (above) are checked tightly.
    atom14, aa = m.inference(db, zs=zs.cuda())
    # (stress weights on the non-default bf16 / TF32 operand paths: only the ODE state is pinned, the decode tail's
    #  unguarded torsion normalisation amplifies their larger state error beyond any meaningful bound)
    if not (case.get("stress") and path in ("bf16", "tf32")):
        assert rel_l2(atom14.cpu(), g["atom14"]) < tol, ("inference", rel_l2(atom14.cpu(), g["atom14"]))
        assert max_rel(atom14.cpu(), g["atom14"]) < 10 * tol, ("inference", max_rel(atom14.cpu(), g["atom14"]))
    x49 = m.model.sample_euler(zs.cuda(), euler_time_grid(49), **kw)
    assert max_rel(x49.cpu(), g["x49"]) < tol, ("x49", max_rel(x49.cpu(), g["x49"]))
    assert (aa.cpu().numpy() == g["aa_out"]).all()


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c.get("canonical_quat")])
def test_two_trunk_follows_reference_eigh_sign(name, path):
    """tps / inpainting against the UNPATCHED reference: with the eigenvector signs the reference's own
    torch.linalg.eigh produced (golden `quat_sign`), the CUDA path reproduces its velocities and Euler state;
    `quat_sign_mode="eigh_cpu"` derives the same signs by making the same LAPACK call."""
    case, args, cfg, sd, batch, zs, g = load_case(name)
    args.sampling_method = "euler"
    m = _wrapper(args, sd, path)
    tol = TOL_SIMT if path == "simt" else TOL
    kw = m.prep_batch(_dev(batch))["model_kwargs"]
    t = torch.tensor(case["t_fwd"]).cuda()
    for mode in (torch.from_numpy(g["quat_sign"]), "eigh_cpu"):
        m.model.quat_sign_mode = mode
        v = m.model.forward_inference(zs.cuda(), t, **kw)
        assert max_rel(v.cpu(), g["v_eigh"]) < tol, (mode, max_rel(v.cpu(), g["v_eigh"]))
        xk = m.model.sample_euler(zs.cuda(), euler_time_grid(case["K"]), **kw)
        assert max_rel(xk.cpu(), g["x_euler_eigh"]) < tol, (mode, max_rel(xk.cpu(), g["x_euler_eigh"]))
    m.model.quat_sign_mode = "canonical"
    v = m.model.forward_inference(zs.cuda(), t, **kw)
    assert max_rel(v.cpu(), g["v"]) < tol


_ORACLE_CACHE = {}


def _oracle_shape_case(shape):
    """Synthetic case at `shape` = (B, T, L, K) and the CPU oracle's Euler state for it (cached per shape)."""
    from mdgen_b200.config import config_from_args, default_args
    from oracle import mdgen_oracle as O
    B, T, L, K = shape
    args = default_args(sim_condition=True, prepend_ipa=True, crop=L, num_frames=T,
                        abs_pos_emb=(L == 4), sampling_method="euler")
    cfg = config_from_args(args)
    sd = synthetic_state_dict(cfg, seed=0)
    batch = synthetic_batch(B, T, L, seed=3, pad_last=(5 if L > 8 else 0))
    zs = synthetic_noise(B, T, L, cfg.latent_dim, seed=4)
    if shape not in _ORACLE_CACHE:
        op = O.prep_batch(cfg, batch)
        kw = dict(mask=op["mask"], start=op["start"], end=op["end"], x_cond=op["x_cond"],
                  x_cond_mask=op["x_cond_mask"], aatype=op["aatype"])
        with torch.no_grad():
            xo = O.sample_euler(sd, cfg, zs, euler_time_grid(K), **kw)
        _ORACLE_CACHE[shape] = (op["latents"], xo)
    return args, sd, batch, zs, _ORACLE_CACHE[shape]


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("shape", [(1, 300, 4, 4), (2, 40, 70, 3), (1, 130, 33, 2)])
def test_oracle_other_shapes(shape, path):
    """Long time axis (flash path, ragged tiles), long residue axis, odd sizes, padding."""
    args, sd, batch, zs, (lat_o, xo) = _oracle_shape_case(shape)
    m = _wrapper(args, sd, path)
    prep = m.prep_batch(_dev(batch))
    xk = m.model.sample_euler(zs.cuda(), euler_time_grid(shape[3]), **prep["model_kwargs"])
    assert max_rel(prep["latents"].cpu(), lat_o) < TOL_GEOM
    tol = TOL_SIMT if path == "simt" else TOL
    assert max_rel(xk.cpu(), xo) < tol, max_rel(xk.cpu(), xo)
    assert rel_l2(xk.cpu(), xo) < tol


@pytest.mark.parametrize("variant", [256, 257, 258, 0, 1, 2, 3, 4, 6, 14])
@pytest.mark.parametrize("shape", [(1, 300, 4, 4), (2, 40, 70, 3), (5, 300, 4, 2)])
def test_attention_variants_match_oracle(shape, variant):
    """Every build variant of the tcgen05 attention kernels (option `attn_variant`: 256 = generation 8, the
    default, 257 = its 2-query-tile kernel forced, 258 = all exponentials on the MUFU pipe; generation 7: bit 0 bf16 P.V, bit 1 staged
    pre-pass, bit 2 persistent kernel, bit 3 persistent with 12 softmax warps) against the
    oracle: frame attention over 300 frames (3 query tiles, ragged key tiles; with B = 5 there are 960 work
    items, so every persistent CTA walks through several of them) and residue attention over 70 residues with
    padded (masked) keys."""
    args, sd, batch, zs, (_, xo) = _oracle_shape_case(shape)
    m = _wrapper(args, sd, "fp16" if variant >= 256 else "bf16")
    m.model.engine().set_option("attn_variant", variant)
    assert m.model.engine().get_option("attn_variant") == variant
    prep = m.prep_batch(_dev(batch))
    xk = m.model.sample_euler(zs.cuda(), euler_time_grid(shape[3]), **prep["model_kwargs"])
    assert max_rel(xk.cpu(), xo) < TOL, max_rel(xk.cpu(), xo)
    assert rel_l2(xk.cpu(), xo) < TOL


def test_dopri5_sampler_matches_oracle_replay():
    """the reference's default adaptive sampler (`--sampling_method dopri5`,
    integrators.py:106-113 -> torchdiffeq) around the CUDA forward. The accepted step sequence of the GPU run is
    replayed through the oracle (same Dormand-Prince steps and dense output, CPU forward), so only the two
    right-hand sides differ; the oracle's own adaptive run must land on the same solution within the
    integrator's tolerance (its accept / reject decisions may differ at rounding level)."""
    from mdgen_b200.config import config_from_args, default_args
    from oracle import mdgen_oracle as O
    B, T, L = 2, 24, 4
    args = default_args(sim_condition=True, prepend_ipa=True, crop=L, num_frames=T, abs_pos_emb=True,
                        sampling_method="dopri5")
    cfg = config_from_args(args)
    sd = synthetic_state_dict(cfg, seed=0)
    batch = synthetic_batch(B, T, L, seed=5)
    zs = synthetic_noise(B, T, L, cfg.latent_dim, seed=6)
    m = _wrapper(args, sd, "bf16")
    prep = m.prep_batch(_dev(batch))
    sample_fn = m.transport_sampler.sample_ode(sampling_method=args.sampling_method)   # == wrapper.py:441
    x = sample_fn(zs.cuda(), m.model.forward_inference, **prep["model_kwargs"])[-1]
    st = m.transport_sampler.last_stats
    assert st["accepted"] >= 2 and st["nfe"] == 2 + 6 * (st["accepted"] + st["rejected"])
    assert st["steps"][-1][0] < 1.0 <= st["steps"][-1][0] + st["steps"][-1][1]
    op = O.prep_batch(cfg, batch)
    kw = dict(mask=op["mask"], start=op["start"], end=op["end"], x_cond=op["x_cond"],
              x_cond_mask=op["x_cond_mask"], aatype=op["aatype"])
    with torch.no_grad():
        xr = O.sample_dopri5_replay(sd, cfg, zs, st["steps"], **kw)
        xa, steps, nfe = O.sample_dopri5(sd, cfg, zs, **kw)
    assert max_rel(x.cpu(), xr) < TOL, max_rel(x.cpu(), xr)
    assert rel_l2(x.cpu(), xa) < 5e-3, (rel_l2(x.cpu(), xa), len(steps), st["accepted"])
    # public API: inference() with the default sampling_method decodes that state
    atom14, _ = m.inference(_dev(batch), zs=zs.cuda())
    a_ref = m.model.engine().decode_atom14(x, _dev(batch)["rots"][:, 0], _dev(batch)["trans"][:, 0],
                                           _dev(batch)["seqres"])
    assert torch.isfinite(atom14).all() and rel_l2(atom14.cpu(), a_ref.cpu()) < 1e-4


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of silently falling back."""
    from mdgen_b200._lib import MDGenError
    from mdgen_b200.config import default_args
    from mdgen_b200.wrapper import NewMDGenWrapper
    args = default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=4)
    m = NewMDGenWrapper(args)
    with pytest.raises(MDGenError):
        m.model.engine()


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _tf32(x):
    """round-to-nearest fp32 -> tf32 (10-bit mantissa), like cvt.rna.tf32.f32"""
    i = x.view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


@pytest.mark.parametrize("shape", [(128, 192, 32), (1000, 1152, 384), (4096, 384, 1536),
                                   (50000, 1536, 384), (333, 384, 384)])
@pytest.mark.parametrize("dtype", ["tf32", "bf16"])
@pytest.mark.parametrize("act", [0, 1])
def test_tc_gemm_matches_fp64_of_rounded_operands(shape, act, dtype):
    """The tcgen05 TF32 GEMM in isolation: exact products of TF32-rounded operands with fp32
    accumulation -> compare with an fp64 matmul of the same rounded operands (tight tolerance),
    and with the fp32 SIMT kernel on the unrounded operands (TF32 rounding tolerance)."""
    from mdgen_b200._lib import Engine, MDGenError
    from mdgen_b200.config import config_from_args, default_args
    eng = Engine(config_from_args(default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=4)))
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    if dtype == "bf16" and K % 64:
        pytest.skip("bf16 K-block is 64 elements")
    rnd = _tf32 if dtype == "tf32" else _bf16
    try:
        y = eng.debug_linear(A, W, b, act=act, use_tc=1 if dtype == "tf32" else 2)
    except MDGenError as e:
        pytest.skip(str(e))
    ref = rnd(A).double() @ rnd(W).double().T + b.double()
    if act:
        ref = torch.nn.functional.gelu(ref)
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err   # fp32 accumulation-order noise only (K up to 1536)
    y32 = eng.debug_linear(A, W, b, act=act, use_tc=0)
    ref32 = A.double() @ W.double().T + b.double()
    if act:
        ref32 = torch.nn.functional.gelu(ref32)
    assert float((y32.double() - ref32).abs().max() / ref32.abs().max()) < 2e-6
    assert float((y.double() - ref32).abs().max() / ref32.abs().max()) < (2e-3 if dtype == "tf32" else 2e-2)


@pytest.mark.parametrize("shape", [(128, 192, 64), (1000, 1152, 384), (50000, 1536, 384), (333, 384, 384), (31, 192, 128)])
@pytest.mark.parametrize("fmt", ["bf16", "f16"])
@pytest.mark.parametrize("act", [0, 1])
def test_tc_gemm_16bit_output_epilogue(shape, act, fmt):
    """The QKV / fc1 form of the tcgen05 GEMM (16-bit operands AND 16-bit output, written by bulk-tensor stores from a
    swizzled staging tile, rows >= M clipped by the tensor map): every output is the correctly-rounded-to-16-bit value of
    the fp64 product of the rounded operands up to fp32 accumulation noise, including ragged M."""
    from mdgen_b200._lib import Engine
    from mdgen_b200.config import config_from_args, default_args
    eng = Engine(config_from_args(default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=4)))
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K + 1)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    rnd = _bf16 if fmt == "bf16" else (lambda x: x.half().float())
    y = eng.debug_linear(A, W, b, act=act, use_tc=4 if fmt == "bf16" else 5)
    ref = rnd(A).double() @ rnd(W).double().T + b.double()
    if act:
        ref = torch.nn.functional.gelu(ref)
    ulp = 2.0 ** -8 if fmt == "bf16" else 2.0 ** -11
    err = (y.double() - ref).abs()
    bound = ref.abs() * ulp * 1.02 + 2e-5 * float(ref.abs().max())
    assert bool((err <= bound).all()), float((err - bound).max())
    assert torch.equal(y, rnd(y))                         # values are representable in the 16-bit format


# ---------------------------------------------------------------------------------------------
# Size-independent properties at BASELINE.json's full sequence shape (T = 1000 frames, crop 4),
# where the CPU oracle is too slow to be the checker.
def _full_size_setup(B=6, T=1000, L=4, path="bf16"):
    from mdgen_b200.config import config_from_args, default_args
    args = default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=L, num_frames=T,
                        sampling_method="euler")
    cfg = config_from_args(args)
    sd = synthetic_state_dict(cfg, seed=0)
    m = _wrapper(args, sd, path)
    batch = synthetic_batch(B, T, L, seed=7, vary_frames=False)
    zs = synthetic_noise(B, T, L, cfg.latent_dim, seed=8)
    prep = m.prep_batch(_dev(batch))
    return m, batch, zs.cuda(), prep["model_kwargs"]


def _sub_kwargs(kw, idx):
    out = dict(kw)
    out["mask"] = kw["mask"][idx]
    out["start_frames"] = kw["start_frames"][idx]
    out["end_frames"] = kw["end_frames"][idx]
    out["aatype"] = kw["aatype"][idx]
    out["x_cond"] = kw["x_cond"][idx]
    out["x_cond_mask"] = kw["x_cond_mask"][idx]
    return out


def test_full_size_euler_composition_and_batch_independence():
    """(a) K Euler steps == K/2 steps followed by the remaining K/2 on the same grid (exercises the
    device step counter, the hoisted IPA trunk and the adaLN table indexing);
    (b) trajectories are independent: sampling a permuted sub-batch gives the permuted rows —
    the property the multi-GPU sharding relies on."""
    m, batch, zs, kw = _full_size_setup()
    K = 4
    grid = euler_time_grid(100)[: K + 1]
    x_all = m.model.sample_euler(zs, grid, **kw)
    assert torch.isfinite(x_all).all()
    x_half = m.model.sample_euler(zs, grid[: K // 2 + 1], **kw)
    x_two = m.model.sample_euler(x_half, grid[K // 2:], **kw)
    assert max_rel(x_two.cpu(), x_all.cpu()) < 1e-5
    idx = torch.tensor([4, 1, 3], device="cuda")
    x_sub = m.model.sample_euler(zs[idx], grid, **_sub_kwargs(kw, idx))
    assert max_rel(x_sub.cpu(), x_all[idx].cpu()) < 1e-5


def test_decode_is_equivariant_under_rigid_motion():
    """atom14(decode(x; g∘T0)) == g · atom14(decode(x; T0)) for a global rigid motion g of the
    frame-0 rigids (mdgen/wrapper.py:469 composes the offsets onto them)."""
    from mdgen_b200._lib import Engine
    from mdgen_b200.config import config_from_args, default_args
    cfg = config_from_args(default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=4))
    eng = Engine(cfg)
    B, T, L = 3, 50, 4
    batch = synthetic_batch(B, T, L, seed=11)
    x = synthetic_noise(B, T, L, 21, seed=12).cuda()
    R0, t0, sq = batch["rots"][:, 0].cuda(), batch["trans"][:, 0].cuda(), batch["seqres"].cuda()
    q = torch.tensor([0.3, -0.5, 0.2, 0.78]); q = q / q.norm()
    w, a, b, c = q.tolist()
    G = torch.tensor([[w*w+a*a-b*b-c*c, 2*(a*b-w*c), 2*(a*c+w*b)], [2*(a*b+w*c), w*w-a*a+b*b-c*c, 2*(b*c-w*a)],
                      [2*(a*c-w*b), 2*(b*c+w*a), w*w-a*a-b*b+c*c]]).cuda()
    shift = torch.tensor([1.5, -2.0, 0.7]).cuda()
    a1 = eng.decode_atom14(x, R0, t0, sq)
    a2 = eng.decode_atom14(x, G @ R0, (t0 @ G.T) + shift, sq)
    present = (a1.abs().sum(-1, keepdim=True) > 0).float()      # absent atoms are exact zeros in both
    expect = (a1 @ G.T + shift) * present
    assert max_rel(a2.cpu(), expect.cpu()) < 2e-5


def test_padded_residues_do_not_influence_real_ones():
    """ATLAS-style padding (mask = 0): changing the padded residues' latents / noise must not change
    the velocity of real residues (key-padding masks in both attentions and in IPA)."""
    from mdgen_b200.config import config_from_args, default_args
    B, T, L, pad = 2, 70, 20, 6
    args = default_args(sim_condition=True, prepend_ipa=True, crop=L, num_frames=T, sampling_method="euler")
    cfg = config_from_args(args)
    sd = synthetic_state_dict(cfg, seed=0)
    m = _wrapper(args, sd, "bf16")
    batch = synthetic_batch(B, T, L, seed=5, pad_last=pad)
    zs = synthetic_noise(B, T, L, cfg.latent_dim, seed=6).cuda()
    kw = m.prep_batch(_dev(batch))["model_kwargs"]
    t = torch.tensor([0.4, 0.6]).cuda()
    v1 = m.model.forward_inference(zs, t, **kw)
    zs2 = zs.clone()
    zs2[:, :, L - pad:] = 5.0 * torch.randn_like(zs2[:, :, L - pad:])
    v2 = m.model.forward_inference(zs2, t, **kw)
    assert max_rel(v2[:, :, : L - pad].cpu(), v1[:, :, : L - pad].cpu()) < 1e-5


def test_featurize_atom14_matches_reference_golden(golden_dir):
    """Device re-featurisation kernel (mdgen_featurize_atom14) vs the reference's host functions
    (golden: tests/golden/gen_featurize_golden.py), and a chained two-rollout run stays finite and
    equals featurising on the host oracle."""
    import os

    import numpy as np
    from mdgen_b200._lib import Engine
    from mdgen_b200.config import config_from_args, default_args
    from oracle import mdgen_oracle as O
    f = dict(np.load(os.path.join(golden_dir, "featurize.npz")))
    eng = Engine(config_from_args(default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=4)))
    for name in ("sim_c1", "atlas_small", "tps"):
        case = CASES[name]
        g = np.load(os.path.join(golden_dir, f"{name}.npz"))
        batch = synthetic_batch(case["B"], case["T"], case["L"], seed=1, **case.get("batch", {}))
        a14 = torch.from_numpy(g["atom14"])[:, -1].cuda()
        R, t, sc, m = eng.featurize_atom14(a14, batch["seqres"].cuda())
        assert (R.cpu() - torch.from_numpy(f[f"{name}/rots"])).abs().max() < 1e-5
        assert (t.cpu() - torch.from_numpy(f[f"{name}/trans"])).abs().max() < 1e-6
        ref_t, ref_m = torch.from_numpy(f[f"{name}/torsions"]), torch.from_numpy(f[f"{name}/torsion_mask"])
        assert (m.cpu() == ref_m).all()
        d = (sc.cpu() - ref_t).abs()
        assert (d * ref_m[..., None]).max() < 5e-5           # defined torsions
        # undefined torsions (mask 0: zero-padded / repeated atoms) are decided by rounding in the
        # reference; the kernel reproduces its operation order, so they agree too, but less tightly
        assert d.max() < 2e-3, float(d.max())
    # chained rollouts through the public wrapper (no host round trip between them)
    case, args, cfg, sd, batch, zs, g = load_case("sim_c1")
    args.sampling_method = "euler"
    m = _wrapper(args, sd, "bf16")
    one = {k: (v[:, :1] if k in ("torsions", "trans", "rots") else v) for k, v in _dev(batch).items()}
    a1, nb = m.rollout(one, zs=zs.cuda(), num_steps=4)
    Ro, to, so, _ = O.featurize_atom14(a1[:, -1].cpu(), batch["seqres"])
    assert (nb["rots"][:, 0].cpu() - Ro).abs().max() < 1e-5 and (nb["torsions"][:, 0].cpu() - so).abs().max() < 2e-3
    a2, _ = m.rollout(nb, zs=zs.cuda(), num_steps=4)
    assert torch.isfinite(a2).all() and a2.shape == a1.shape


def test_atlas_shape_forward_matches_oracle():
    """ATLAS-shaped chain (crop 256, no abs_pos_emb, 16 padded residues): residue attention over
    L = 256 runs on the tcgen05 attention kernel, the IPA trunk sees 256 residues per sample."""
    from mdgen_b200.config import config_from_args, default_args
    from oracle import mdgen_oracle as O
    B, T, L = 1, 66, 256
    args = default_args(sim_condition=True, prepend_ipa=True, crop=L, num_frames=T, sampling_method="euler")
    cfg = config_from_args(args)
    sd = synthetic_state_dict(cfg, seed=0)
    batch = synthetic_batch(B, T, L, seed=9, pad_last=16)
    zs = synthetic_noise(B, T, L, cfg.latent_dim, seed=10)
    m = _wrapper(args, sd, "bf16")
    kw = m.prep_batch(_dev(batch))["model_kwargs"]
    t = torch.tensor([0.35])
    v = m.model.forward_inference(zs.cuda(), t.cuda(), **kw)
    op = O.prep_batch(cfg, batch)
    okw = dict(mask=op["mask"], start=op["start"], end=op["end"], x_cond=op["x_cond"],
               x_cond_mask=op["x_cond_mask"], aatype=op["aatype"])
    with torch.no_grad():
        vo = O.forward(sd, cfg, zs, t, **okw)
    assert max_rel(v.cpu(), vo) < TOL, max_rel(v.cpu(), vo)
    assert rel_l2(v.cpu(), vo) < TOL


# ---------------------------------------------------------------------------------------------
# Forward half of the training step
TRAIN_CASES = ["sim_c1", "atlas_small", "upsampling", "tps", "inpaint", "stress"]


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_training_loss_matches_reference_general_step(name, golden_dir):
    """Transport.training_losses on the CUDA path (mdgen_flow_plan -> mdgen_forward -> mdgen_masked_mse) against
    the per-sample losses of the reference's own general_step(stage='val') with the (t, x0) it drew; then the public
    hooks: general_step / validation_step return a finite scalar with their own draws, training_step refuses."""
    import os

    import numpy as np
    g = np.load(os.path.join(golden_dir, "train_loss.npz"))
    case, args, cfg, sd, batch, zs, _ = load_case(name)
    m = _wrapper(args, sd, "fp16")
    db = _dev(batch)
    prep = m.prep_batch(db)
    out = m.transport.training_losses(model=m.model, x1=prep["latents"], mask=prep["loss_mask"],
                                      model_kwargs=prep["model_kwargs"], t=torch.from_numpy(g[f"{name}/t"]).cuda(),
                                      x0=torch.from_numpy(g[f"{name}/x0"]).cuda())
    assert out["pred"].shape == prep["latents"].shape
    assert max_rel(out["loss"].cpu(), g[f"{name}/loss"]) < TOL, (out["loss"].cpu(), g[f"{name}/loss"])
    torch.manual_seed(3)
    l1 = m.general_step(db, stage="val")
    l2 = m.validation_step(db, 0)
    assert l1.dim() == 0 and torch.isfinite(l1) and torch.isfinite(l2)
    assert "val_loss" in m._log and len(m._log["val_loss"]) == 2
    with pytest.raises(NotImplementedError):
        m.training_step(db, 0)


def test_flow_plan_and_masked_mse_kernels_match_torch():
    """The two loss kernels in isolation (odd sizes, zero mask rows, both interpolants)."""
    from mdgen_b200._lib import Engine
    from mdgen_b200.config import config_from_args, default_args
    from oracle import mdgen_oracle as O
    eng = Engine(config_from_args(default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=4)))
    g = torch.Generator(device="cuda").manual_seed(5)
    for shape in [(3, 7, 5, 21), (2, 64, 4, 28), (5, 1, 3, 21)]:
        x1 = torch.randn(shape, device="cuda", generator=g)
        x0 = torch.randn(shape, device="cuda", generator=g)
        t = torch.rand(shape[0], device="cuda", generator=g)
        for path in ("GVP", "Linear"):
            xt, ut = eng.flow_plan(x1, x0, t, path)
            xr, ur = O.flow_plan(t, x0, x1, path)
            assert max_rel(xt.cpu(), xr.cpu()) < 1e-6 and max_rel(ut.cpu(), ur.cpu()) < 1e-6
        mask = (torch.rand(shape, device="cuda", generator=g) > 0.3).float()
        loss = eng.masked_mse(xt, ut, mask)
        ref = O.mean_flat((xt - ut).double() ** 2, mask.double())
        assert max_rel(loss.cpu(), ref.cpu()) < 1e-5


def test_ema_update_kernel_and_checkpoint_hooks():
    """ExponentialMovingAverage (mdgen/ema.py) on the device + the wrapper's EMA hooks (wrapper.py:65-130)."""
    case, args, cfg, sd, batch, zs, _ = load_case("sim_c1")
    args.ema, args.ema_decay = True, 0.9
    m = _wrapper(args, sd, "fp16")
    m.ema.to(m.device)
    before = {k: v.clone() for k, v in m.ema.params.items()}
    with torch.no_grad():
        for p in m.model.parameters():
            p.add_(1.0)
    m.on_before_zero_grad()
    for k, v in m.model.state_dict().items():
        if v.dtype == torch.float32:
            exp = before[k] - (before[k] - v) * (1 - 0.9)
            assert torch.allclose(m.ema.params[k], exp, rtol=1e-6, atol=1e-6), k
    ck = {}
    m.on_save_checkpoint(ck)
    assert set(ck["ema"]) == {"params", "decay"} and ck["ema"]["decay"] == 0.9
    m.on_load_checkpoint(ck)
    db = _dev(batch)
    m.validation_step(db, 0)                      # loads the EMA weights for validation ...
    assert m.cached_weights is not None
    m.on_validation_epoch_end()                   # ... and restores the trained ones
    assert m.cached_weights is None


def test_rk_kernels_match_torch_and_residual_fusion_option():
    """mdgen_lincomb / mdgen_rk_error_ratio (the adaptive sampler's stage combinations and error norm) against torch,
    and the `fuse_resid_ln` option (residual add inside ln_mod_kernel, EPI_GATE GEMM epilogues) against the default."""
    from mdgen_b200._lib import Engine
    from mdgen_b200.config import config_from_args, default_args
    eng = Engine(config_from_args(default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=4)))
    g = torch.Generator(device="cuda").manual_seed(9)
    n = 2 * 37 * 4 * 21 + 3
    ks = [torch.randn(n, device="cuda", generator=g) for _ in range(7)]
    y = torch.randn(n, device="cuda", generator=g)
    cf = [0.3, 0.0, -1.25, 2.0, 0.5, -0.125, 1.0 / 60]
    out = eng.lincomb(y, 0.37, cf, ks)
    ref = y.double() + 0.37 * sum(c * k.double() for c, k in zip(cf, ks))
    assert max_rel(out.cpu(), ref.cpu()) < 1e-6
    out0 = eng.lincomb(None, 1.0, cf[:3], ks[:3])
    assert max_rel(out0.cpu(), sum(c * k.double() for c, k in zip(cf[:3], ks[:3])).cpu()) < 1e-6
    r = eng.rk_error_ratio(ks[0], ks[1], ks[2], 1e-3, 1e-6)
    tol = 1e-6 + 1e-3 * torch.maximum(ks[1].abs(), ks[2].abs()).double()
    assert abs(r - float(((ks[0].double() / tol) ** 2).mean().sqrt())) < 1e-5 * r
    # residual-add fusion option
    case, args, cfg, sd, batch, zs, gold = load_case("stress")
    m = _wrapper(args, sd, "fp16")
    kw = m.prep_batch(_dev(batch))["model_kwargs"]
    t = torch.tensor(case["t_fwd"]).cuda()
    v0 = m.model.forward_inference(zs.cuda(), t, **kw)
    m.model.engine().set_option("fuse_resid_ln", 1)
    v1 = m.model.forward_inference(zs.cuda(), t, **kw)
    # one extra fp32 rounding per branch, amplified by the stress weights: both stay inside TOL of the golden velocity,
    # so they are within 2*TOL of each other; on ordinary weights the two orders agree to fp32 round-off.
    assert max_rel(v1.cpu(), gold["v"]) < TOL
    assert max_rel(v1.cpu(), v0.cpu()) < 2 * TOL
    case, args, cfg, sd, batch, zs, gold = load_case("sim_c1")
    m = _wrapper(args, sd, "fp16")
    kw = m.prep_batch(_dev(batch))["model_kwargs"]
    t = torch.tensor(case["t_fwd"]).cuda()
    v0 = m.model.forward_inference(zs.cuda(), t, **kw)
    m.model.engine().set_option("fuse_resid_ln", 1)
    v1 = m.model.forward_inference(zs.cuda(), t, **kw)
    assert max_rel(v1.cpu(), v0.cpu()) < 2e-5
    assert max_rel(v1.cpu(), gold["v"]) < TOL
