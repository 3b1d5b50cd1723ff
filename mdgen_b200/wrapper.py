"""`NewMDGenWrapper` — the drop-in boundary (mdgen/wrapper.py:175-507).

Same class name, constructor (`args` namespace), attributes (`.args`, `.model`, `.latent_dim`,
`.transport`, `.transport_sampler`) and `prep_batch` / `inference` signatures as the reference,
so `sim_inference.py` / `upsampling_inference.py` / `tps_inference.py` call it unchanged
(`from mdgen_b200.wrapper import NewMDGenWrapper`). Everything numeric is executed by
libmdgen_b200 through the C ABI; this file is glue.
"""
from __future__ import annotations

import time
from collections import defaultdict

import torch
import torch.nn as nn

from .config import backfill_args, config_from_args
from .model import LatentMDGenModel
from .rigid import Rigid, Rotation
from .ema import ExponentialMovingAverage
from .transport import Sampler, create_transport

DESIGN_IDX = [1, 2]      # mdgen/wrapper.py:30-32
COND_IDX = [0, 3]

try:  # the reference subclasses pl.LightningModule (mdgen/wrapper.py:46)
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # pytorch_lightning is not installed in this image

    class _Base(nn.Module):
        def __init__(self):
            super().__init__()
            self.trainer = None
            self.current_epoch = 0

        def save_hyperparameters(self, *a, **k):
            return None

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")

        @classmethod
        def load_from_checkpoint(cls, path, map_location=None, **kw):
            """Lightning checkpoint layout: {'state_dict', 'hyper_parameters': {'args': ...}}
            (mdgen/wrapper.py:50,120-130)."""
            ckpt = torch.load(path, map_location=map_location or "cpu", weights_only=False)
            obj = cls(ckpt["hyper_parameters"]["args"])
            obj.load_state_dict(ckpt["state_dict"], strict=True)
            return obj


class NewMDGenWrapper(_Base):
    def __init__(self, args):
        super().__init__()
        self.save_hyperparameters()
        self.args = args
        backfill_args(args)                                   # wrapper.py:178-194,215-216
        self.cfg = config_from_args(args)
        self.latent_dim = self.cfg.latent_dim                 # wrapper.py:196-202
        self.model = LatentMDGenModel(args, self.latent_dim)
        self.transport = create_transport(args, args.path_type, args.prediction, None)
        self.transport_sampler = Sampler(self.transport)
        self.stage = "val"
        self._log = defaultdict(list)                         # wrapper.py:52-55
        self.last_log_time = time.time()
        self.iter_step = 0
        if args.ema:                                          # wrapper.py:204-208
            self.ema = ExponentialMovingAverage(model=self.model, decay=args.ema_decay)
            self.cached_weights = None

    # ------------------------------------------------------------------------------------------
    def prep_batch(self, batch):
        """== mdgen/wrapper.py:283-365 (featurisation kernel: mdgen_prep_batch)."""
        eng = self.model.engine()
        rots, trans = batch["rots"], batch["trans"]
        B, T, L = trans.shape[:3]
        torsions = batch["torsions"]
        if self.args.no_design_torsion and not self.args.no_torsion:       # wrapper.py:322-325
            torsions = torsions.clone()
            torsions[:, :, DESIGN_IDX] = 0
        with torch.cuda.device(trans.device):
            latents, x_cond, cond_mask = eng.prep_batch(rots, trans, torsions)
        rigids = Rigid(Rotation(rots), trans)
        D = self.latent_dim
        frame_loss_mask = batch["mask"].unsqueeze(-1).expand(-1, -1, D - 14)
        torsion_loss_mask = batch["torsion_mask"].unsqueeze(-1).expand(-1, -1, -1, 2).reshape(B, L, 14)
        if self.args.supervise_all_torsions:                                # wrapper.py:329-332
            torsion_loss_mask = torch.ones_like(torsion_loss_mask)
        elif self.args.supervise_no_torsions:
            torsion_loss_mask = torch.zeros_like(torsion_loss_mask)
        loss_mask = torch.cat([frame_loss_mask, torsion_loss_mask], -1).unsqueeze(1).expand(-1, T, -1, -1)
        return {
            "rigids": rigids,
            "latents": latents,
            "loss_mask": loss_mask,
            "model_kwargs": {
                "start_frames": rigids[:, 0],
                "end_frames": rigids[:, -1],
                "mask": batch["mask"].unsqueeze(1).expand(-1, T, -1),
                "aatype": batch["seqres"],
                "x_cond": x_cond,
                "x_cond_mask": cond_mask,
            },
        }

    @torch.no_grad()
    def inference(self, batch, zs=None, num_steps=None):
        """== mdgen/wrapper.py:405-484: prep -> noise -> ODE sample -> decode to atom14.

        `zs` / `num_steps` are optional extensions (the reference draws zs with torch.randn on the
        model device and hard-wires sample_ode's default num_steps=50, i.e. 49 Euler steps —
        wrapper.py:439-447; both defaults are preserved)."""
        prep = self.prep_batch(batch)
        rigids = prep["rigids"]
        B, T, L = rigids.shape
        if zs is None:
            zs = torch.randn(B, T, L, self.latent_dim, device=self.device)     # wrapper.py:439
        kw = {} if num_steps is None else {"num_steps": num_steps}
        sample_fn = self.transport_sampler.sample_ode(sampling_method=self.args.sampling_method, **kw)
        samples = sample_fn(zs, self.model.forward_inference, **prep["model_kwargs"])[-1]
        eng = self.model.engine()
        with torch.cuda.device(samples.device):
            atom14 = eng.decode_atom14(samples, batch["rots"][:, 0], batch["trans"][:, 0],
                                       batch["seqres"])              # wrapper.py:456-478
        aa_out = batch["seqres"][:, None].expand(B, T, L)            # wrapper.py:483
        return atom14, aa_out

    @torch.no_grad()
    def rollout(self, batch, zs=None, num_steps=None):
        """One forward-simulation rollout with on-device re-featurisation of its last frame ==
        `rollout(model, batch)` of sim_inference.py:61-98 without the device->CPU->device round trip
        (`atom14_to_frames` / `atom37_to_torsions` run in `mdgen_featurize_atom14`).
        `batch` holds ONE conditioning frame per trajectory (T = 1); returns (atom14, new_batch) where
        `new_batch` seeds the next rollout."""
        T = self.args.num_frames
        expanded = {
            "torsions": batch["torsions"].expand(-1, T, -1, -1, -1),
            "torsion_mask": batch["torsion_mask"],
            "trans": batch["trans"].expand(-1, T, -1, -1),
            "rots": batch["rots"].expand(-1, T, -1, -1, -1),
            "seqres": batch["seqres"],
            "mask": batch["mask"],
        }
        atom14, _ = self.inference(expanded, zs=zs, num_steps=num_steps)
        eng = self.model.engine()
        with torch.cuda.device(atom14.device):
            rots, trans, tors, _ = eng.featurize_atom14(atom14[:, -1], batch["seqres"])
        new_batch = dict(batch)
        new_batch["rots"] = rots[:, None]
        new_batch["trans"] = trans[:, None]
        new_batch["torsions"] = tors[:, None]
        return atom14, new_batch

    # -- training / validation surface (mdgen/wrapper.py:56-172, 367-403) ---------------------------------------
    def log(self, key, data):
        """== wrapper.py:56-63."""
        if isinstance(data, torch.Tensor):
            data = data.mean().item()
        if self.stage == "train" or self.args.validate:
            self._log["iter_" + key].append(data)
        self._log[self.stage + "_" + key].append(data)

    def general_step(self, batch, stage="train"):
        """== wrapper.py:367-403 (non-design): featurise, draw (t, x0), interpolate, run the denoiser, masked MSE.
        The loss VALUE is computed by the CUDA path for both stages; it carries no autograd graph (backward kernels are a
        'next' row), so `training_step` refuses to hand it to an optimiser."""
        self.iter_step += 1
        self.stage = stage
        start1 = time.time()
        prep = self.prep_batch(batch)
        start = time.time()
        out_dict = self.transport.training_losses(
            model=self.model, x1=prep["latents"], aatype1=None, mask=prep["loss_mask"],
            model_kwargs=prep["model_kwargs"])
        self.log("model_dur", time.time() - start)
        loss = out_dict["loss"]
        self.log("loss", loss)
        self.log("time", out_dict["t"])
        self.log("dur", time.time() - self.last_log_time)
        if "name" in batch:
            self.log("name", ",".join(batch["name"]))
        self.log("general_step_dur", time.time() - start1)
        self.last_log_time = time.time()
        return loss.mean()

    def training_step(self, batch, batch_idx):
        raise NotImplementedError(
            "mdgen_b200: the backward pass of the training step is not implemented (SURVEY.md §8f-3); the forward half "
            "- general_step(batch, stage='val') / validation_step - runs on the CUDA path")

    @torch.no_grad()
    def validation_step(self, batch, batch_idx):
        """== wrapper.py:88-99."""
        if self.args.ema:
            if self.ema.device != self.device:
                self.ema.to(self.device)
            if self.cached_weights is None:
                self.load_ema_weights()
        loss = self.general_step(batch, stage="val")
        self.validation_step_extra(batch, batch_idx)
        return loss

    def validation_step_extra(self, batch, batch_idx):
        pass

    def load_ema_weights(self):
        """== wrapper.py:65-72."""
        self.cached_weights = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        self.model.load_state_dict(self.ema.state_dict()["params"])

    def restore_cached_weights(self):
        """== wrapper.py:74-77."""
        self.model.load_state_dict(self.cached_weights)
        self.cached_weights = None

    def on_before_zero_grad(self, *args, **kwargs):
        if self.args.ema:
            self.ema.update(self.model)                       # wrapper.py:78-80

    def on_validation_epoch_end(self):
        if self.args.ema:
            self.restore_cached_weights()                     # wrapper.py:107-110

    def on_load_checkpoint(self, checkpoint):
        if self.args.ema:
            self.ema.load_state_dict(checkpoint["ema"])       # wrapper.py:120-124

    def on_save_checkpoint(self, checkpoint):
        if self.args.ema:                                     # wrapper.py:126-130
            if self.cached_weights is not None:
                self.restore_cached_weights()
            checkpoint["ema"] = self.ema.state_dict()

    def configure_optimizers(self):
        cls = torch.optim.AdamW if self.args.adamW else torch.optim.Adam
        return cls(filter(lambda p: p.requires_grad, self.model.parameters()), lr=self.args.lr)
