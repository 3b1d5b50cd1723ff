"""`NewMDGenWrapper` — the drop-in boundary (mdgen/wrapper.py:175-507).

Same class name, constructor (`args` namespace), attributes (`.args`, `.model`, `.latent_dim`,
`.transport`, `.transport_sampler`) and `prep_batch` / `inference` signatures as the reference,
so `sim_inference.py` / `upsampling_inference.py` / `tps_inference.py` call it unchanged
(`from mdgen_b200.wrapper import NewMDGenWrapper`). Everything numeric is executed by
libmdgen_b200 through the C ABI; this file is glue.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .config import backfill_args, config_from_args
from .model import LatentMDGenModel
from .rigid import Rigid, Rotation
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

    # -- training hooks: kept as names; the backward path is a 'next' row -----
    def general_step(self, batch, stage="train"):
        raise NotImplementedError("mdgen_b200: training step not implemented (SURVEY.md §8f-3)")

    def training_step(self, batch, batch_idx):
        return self.general_step(batch, stage="train")

    def validation_step(self, batch, batch_idx):
        return self.general_step(batch, stage="val")

    def configure_optimizers(self):
        cls = torch.optim.AdamW if self.args.adamW else torch.optim.Adam
        return cls(filter(lambda p: p.requires_grad, self.model.parameters()), lr=self.args.lr)
