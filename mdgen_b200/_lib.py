"""ctypes binding of libmdgen_b200.so (the C ABI declared in include/mdgen_b200.h).

There is deliberately no fallback: if the CUDA library is missing or fails, the product path
raises. PyTorch is used only for device memory (tensor.data_ptr()) and the current stream.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from .config import MDGenConfig

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmdgen_b200.so")
_lib = None

EXPORTS = [
    "mdgen_create", "mdgen_destroy", "mdgen_last_error", "mdgen_set_tensor",
    "mdgen_finalize_weights", "mdgen_set_residue_tables", "mdgen_forward", "mdgen_sample_euler",
    "mdgen_prep_batch", "mdgen_decode_atom14", "mdgen_abi_version", "mdgen_launch_count",
    "mdgen_set_option", "mdgen_get_option", "mdgen_profile_dump", "mdgen_debug_linear",
    "mdgen_set_featurize_tables", "mdgen_featurize_atom14",
    "mdgen_flow_plan", "mdgen_masked_mse", "mdgen_ema_update", "mdgen_lincomb", "mdgen_rk_error_ratio",
]


class MDGenError(RuntimeError):
    pass


class CConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "latent_dim", "num_layers", "crop", "abs_pos_emb", "use_aa_emb",
        "sim_condition", "tps_condition", "inpainting", "cond_interval", "no_torsion")] + [
        ("time_multiplier", C.c_float)]


class CCond(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("L", C.c_int32)] + [
        (n, C.c_void_p) for n in ("mask", "start_rot", "start_trans", "end_rot", "end_trans",
                                  "x_cond", "x_cond_mask", "aatype", "quat_sign")]


def lib_path() -> str:
    return _LIB_PATH


def load_library():
    """Loads the shared library (building it is the job of mdgen_b200.build / __graft_entry__)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(_LIB_PATH):
        raise MDGenError(
            f"{_LIB_PATH} not found: build it with `python -m mdgen_b200.build` "
            "(mdgen_b200 has no CPU / PyTorch fallback)")
    lib = C.CDLL(_LIB_PATH)
    lib.mdgen_last_error.restype = C.c_char_p
    lib.mdgen_last_error.argtypes = [C.c_void_p]
    lib.mdgen_create.argtypes = [C.POINTER(CConfig), C.POINTER(C.c_void_p)]
    lib.mdgen_destroy.argtypes = [C.c_void_p]
    lib.mdgen_destroy.restype = None
    lib.mdgen_set_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    lib.mdgen_finalize_weights.argtypes = [C.c_void_p, C.c_void_p]
    lib.mdgen_set_residue_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    lib.mdgen_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(CCond), C.c_void_p,
                                  C.c_void_p]
    lib.mdgen_sample_euler.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.POINTER(CCond), C.c_void_p, C.c_void_p]
    lib.mdgen_prep_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 7
    lib.mdgen_decode_atom14.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 6
    lib.mdgen_set_featurize_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    lib.mdgen_featurize_atom14.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 7
    lib.mdgen_flow_plan.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int32] + [C.c_void_p] * 6
    lib.mdgen_masked_mse.argtypes = [C.c_void_p, C.c_int32, C.c_int64] + [C.c_void_p] * 5
    lib.mdgen_ema_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
    lib.mdgen_lincomb.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_void_p]
    lib.mdgen_rk_error_ratio.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                         C.c_float, C.c_void_p, C.c_void_p]
    lib.mdgen_launch_count.restype = C.c_int64
    lib.mdgen_launch_count.argtypes = [C.c_void_p]
    lib.mdgen_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    lib.mdgen_get_option.argtypes = [C.c_void_p, C.c_char_p]
    lib.mdgen_get_option.restype = C.c_int64
    lib.mdgen_profile_dump.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    lib.mdgen_debug_linear.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p]
    _lib = lib
    return lib


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise MDGenError(f"{name}: expected a CUDA tensor (mdgen_b200 has no CPU path)")
    return t.detach().to(torch.float32).contiguous()


def _i64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise MDGenError(f"{name}: expected a CUDA tensor (mdgen_b200 has no CPU path)")
    return t.detach().to(torch.int64).contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class Engine:
    """Owns one `mdgen_handle` (one device, current stream at call time)."""

    def __init__(self, cfg: MDGenConfig):
        self.lib = load_library()
        self.cfg = cfg
        cc = CConfig(
            abi_version=2, latent_dim=cfg.latent_dim, num_layers=cfg.num_layers, crop=cfg.crop,
            abs_pos_emb=int(cfg.abs_pos_emb), use_aa_emb=int(cfg.use_aa_emb),
            sim_condition=int(cfg.sim_condition), tps_condition=int(cfg.tps_condition),
            inpainting=int(cfg.inpainting), cond_interval=int(cfg.cond_interval),
            no_torsion=int(cfg.no_torsion), time_multiplier=float(cfg.time_multiplier))
        h = C.c_void_p()
        rc = self.lib.mdgen_create(C.byref(cc), C.byref(h))
        if rc != 0:
            raise MDGenError(f"mdgen_create failed ({rc}): {self.lib.mdgen_last_error(None).decode()}")
        self.h = h
        self._weights_version = None
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data",
                                 "residue_tables.npz"))
        arrs = [np.ascontiguousarray(z["default_frame"], np.float32),
                np.ascontiguousarray(z["atom14_group_pos"], np.float32),
                np.ascontiguousarray(z["atom14_to_group"], np.int32),
                np.ascontiguousarray(z["atom14_mask"], np.float32)]
        self._check(self.lib.mdgen_set_residue_tables(self.h, *[a.ctypes.data for a in arrs]))
        farrs = [np.ascontiguousarray(z["chi_atom14_idx"], np.int32),
                 np.ascontiguousarray(z["chi_atom_mask"], np.float32),
                 np.ascontiguousarray(z["chi_mask"], np.float32),
                 np.ascontiguousarray(z["bb_mask"], np.float32)]
        self._check(self.lib.mdgen_set_featurize_tables(self.h, *[a.ctypes.data for a in farrs]))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.mdgen_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise MDGenError(f"libmdgen_b200 error {rc}: {self.lib.mdgen_last_error(self.h).decode()}")

    # -- weights -----------------------------------------------------------------------------
    def load_state_dict(self, sd):
        """sd: mapping name -> CUDA tensor (keys of LatentMDGenModel.state_dict())."""
        keep = []
        for name, t in sd.items():
            tt = _f32(t, name)
            keep.append(tt)
            self._check(self.lib.mdgen_set_tensor(self.h, name.encode(), tt.data_ptr(), tt.numel()))
        torch.cuda.current_stream().synchronize()
        self._check(self.lib.mdgen_finalize_weights(self.h, _stream()))

    # -- options -----------------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        self._check(self.lib.mdgen_set_option(self.h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        return int(self.lib.mdgen_get_option(self.h, key.encode()))

    @property
    def launch_count(self) -> int:
        return int(self.lib.mdgen_launch_count(self.h))

    def profile_dump(self):
        buf = C.create_string_buffer(1 << 16)
        self._check(self.lib.mdgen_profile_dump(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, calls = line.split()
            out[name] = (float(ms), int(calls))
        return out

    def debug_linear(self, A, W, bias, act=0, use_tc=1):
        """Test hook: act(A @ W.T + bias) through the library's GEMM kernels."""
        A, W = _f32(A, "A"), _f32(W, "W")
        b = _f32(bias, "bias") if bias is not None else None
        M, K = A.shape
        N = W.shape[0]
        out = torch.empty(M, N, device=A.device, dtype=torch.float32)
        self._check(self.lib.mdgen_debug_linear(self.h, A.data_ptr(), W.data_ptr(),
                                                b.data_ptr() if b is not None else None, M, N, K,
                                                int(act), int(use_tc), out.data_ptr(), _stream()))
        return out

    # -- calls -------------------------------------------------------------------------------
    def _cond(self, B, T, L, mask, start, end, x_cond, x_cond_mask, aatype, quat_sign=None):
        keep = {}
        keep["mask"] = _f32(mask.expand(B, T, L) if mask.dim() == 3 else mask, "mask")
        keep["start_rot"] = _f32(start[0], "start_rot")
        keep["start_trans"] = _f32(start[1], "start_trans")
        if end is not None:
            keep["end_rot"] = _f32(end[0], "end_rot")
            keep["end_trans"] = _f32(end[1], "end_trans")
        keep["x_cond"] = _f32(x_cond, "x_cond")
        keep["x_cond_mask"] = _i64(x_cond_mask, "x_cond_mask")
        if aatype is not None:
            keep["aatype"] = _i64(aatype, "aatype")
        if quat_sign is not None:
            keep["quat_sign"] = _f32(quat_sign, "quat_sign")
            assert keep["quat_sign"].shape == (2, B, L)
        D = self.cfg.latent_dim
        assert keep["mask"].shape == (B, T, L)
        assert keep["start_rot"].shape == (B, L, 3, 3) and keep["start_trans"].shape == (B, L, 3)
        assert keep["x_cond"].shape == (B, T, L, D) and keep["x_cond_mask"].shape == (B, T, L)
        c = CCond(B=B, T=T, L=L)
        for k, v in keep.items():
            setattr(c, k, v.data_ptr())
        return c, keep

    def forward(self, x, t, mask, start, end, x_cond, x_cond_mask, aatype, quat_sign=None):
        B, T, L, D = x.shape
        x = _f32(x, "x")
        t = _f32(t, "t").reshape(B)
        c, keep = self._cond(B, T, L, mask, start, end, x_cond, x_cond_mask, aatype, quat_sign)
        out = torch.empty_like(x)
        self._check(self.lib.mdgen_forward(self.h, x.data_ptr(), t.data_ptr(), C.byref(c),
                                           out.data_ptr(), _stream()))
        return out

    def sample_euler(self, zs, t_grid, mask, start, end, x_cond, x_cond_mask, aatype, quat_sign=None):
        B, T, L, D = zs.shape
        zs = _f32(zs, "zs")
        tg = np.ascontiguousarray(t_grid.detach().cpu().numpy() if torch.is_tensor(t_grid)
                                  else np.asarray(t_grid), np.float32)
        K = int(tg.shape[0]) - 1
        c, keep = self._cond(B, T, L, mask, start, end, x_cond, x_cond_mask, aatype, quat_sign)
        out = torch.empty_like(zs)
        self._check(self.lib.mdgen_sample_euler(self.h, zs.data_ptr(), tg.ctypes.data, K, C.byref(c),
                                                out.data_ptr(), _stream()))
        return out

    def prep_batch(self, rots, trans, torsions):
        B, T, L = trans.shape[:3]
        D = self.cfg.latent_dim
        rots, trans, torsions = _f32(rots, "rots"), _f32(trans, "trans"), _f32(torsions, "torsions")
        lat = torch.empty(B, T, L, D, device=rots.device, dtype=torch.float32)
        xc = torch.empty_like(lat)
        cm = torch.empty(B, T, L, device=rots.device, dtype=torch.int64)
        self._check(self.lib.mdgen_prep_batch(self.h, B, T, L, rots.data_ptr(), trans.data_ptr(),
                                              torsions.data_ptr(), lat.data_ptr(), xc.data_ptr(),
                                              cm.data_ptr(), _stream()))
        return lat, xc, cm

    def featurize_atom14(self, atom14, seqres):
        """atom14 [B,L,14,3] (one frame), seqres [B,L] -> rots, trans, torsions, torsion_mask."""
        B, L = seqres.shape
        a = _f32(atom14, "atom14")
        sq = _i64(seqres, "seqres")
        assert a.shape == (B, L, 14, 3)
        rots = torch.empty(B, L, 3, 3, device=a.device, dtype=torch.float32)
        trans = torch.empty(B, L, 3, device=a.device, dtype=torch.float32)
        tors = torch.empty(B, L, 7, 2, device=a.device, dtype=torch.float32)
        tmask = torch.empty(B, L, 7, device=a.device, dtype=torch.float32)
        self._check(self.lib.mdgen_featurize_atom14(self.h, B, L, a.data_ptr(), sq.data_ptr(), rots.data_ptr(),
                                                    trans.data_ptr(), tors.data_ptr(), tmask.data_ptr(),
                                                    _stream()))
        return rots, trans, tors, tmask

    # -- training / validation loss pieces ------------------------------------------------------------
    def flow_plan(self, x1, x0, t, path_type="GVP"):
        """xt, ut of the interpolant (mdgen/transport/path.py:118-135) with per-sample t [B]."""
        x1, x0 = _f32(x1, "x1"), _f32(x0, "x0")
        B = x1.shape[0]
        t = _f32(t, "t").reshape(B)
        per = x1.numel() // B
        xt, ut = torch.empty_like(x1), torch.empty_like(x1)
        self._check(self.lib.mdgen_flow_plan(self.h, B, per, {"GVP": 0, "Linear": 1}[path_type], x1.data_ptr(),
                                             x0.data_ptr(), t.data_ptr(), xt.data_ptr(), ut.data_ptr(), _stream()))
        return xt, ut

    def masked_mse(self, pred, target, mask):
        """mean_flat((pred - target)^2, mask) per sample (mdgen/transport/transport.py:13-17,189)."""
        pred, target = _f32(pred, "pred"), _f32(target, "target")
        mask = _f32(mask.expand_as(pred), "mask")
        B = pred.shape[0]
        loss = torch.empty(B, device=pred.device, dtype=torch.float32)
        self._check(self.lib.mdgen_masked_mse(self.h, B, pred.numel() // B, pred.data_ptr(), target.data_ptr(),
                                              mask.data_ptr(), loss.data_ptr(), _stream()))
        return loss

    def ema_update(self, stored, param, decay):
        """stored <- stored - (stored - param) (1 - decay), in place (mdgen/ema.py:41-50)."""
        assert stored.is_cuda and stored.is_contiguous() and stored.dtype == torch.float32
        p = _f32(param, "param")
        self._check(self.lib.mdgen_ema_update(self.h, stored.data_ptr(), p.data_ptr(), stored.numel(), float(decay),
                                              _stream()))

    # -- Runge-Kutta plumbing of the adaptive sampler ------------------------------------------------
    def lincomb(self, y, scale, coeffs, ks):
        """(y or 0) + scale * sum_i coeffs[i] * ks[i]  in one pass (<= 8 terms)."""
        ks = [_f32(k, "k") for k in ks]
        n = ks[0].numel()
        yy = _f32(y, "y") if y is not None else None
        out = torch.empty_like(ks[0])
        cf = (C.c_float * len(ks))(*[float(c) for c in coeffs])
        ptrs = (C.c_void_p * len(ks))(*[k.data_ptr() for k in ks])
        self._check(self.lib.mdgen_lincomb(self.h, n, yy.data_ptr() if yy is not None else None, float(scale), cf, ptrs,
                                           len(ks), out.data_ptr(), _stream()))
        return out

    def rk_error_ratio(self, err, y0, y1, rtol, atol) -> float:
        err, y0, y1 = _f32(err, "err"), _f32(y0, "y0"), _f32(y1, "y1")
        r = C.c_float(0.0)
        self._check(self.lib.mdgen_rk_error_ratio(self.h, err.numel(), err.data_ptr(), y0.data_ptr(), y1.data_ptr(),
                                                  float(rtol), float(atol), C.byref(r), _stream()))
        return float(r.value)

    def decode_atom14(self, samples, start_rot, start_trans, seqres):
        B, T, L, D = samples.shape
        samples = _f32(samples, "samples")
        sr, st = _f32(start_rot, "start_rot"), _f32(start_trans, "start_trans")
        sq = _i64(seqres, "seqres")
        out = torch.empty(B, T, L, 14, 3, device=samples.device, dtype=torch.float32)
        self._check(self.lib.mdgen_decode_atom14(self.h, B, T, L, samples.data_ptr(), sr.data_ptr(),
                                                 st.data_ptr(), sq.data_ptr(), out.data_ptr(),
                                                 _stream()))
        return out
