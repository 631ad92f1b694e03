"""ARCVAETrainerWithLoss — the STEP of trainer.py of the reference (trainer.py:267-333, :489-522) plus the two
schedules; epoch bookkeeping, plotting and checkpoint I/O of the reference are host glue and out of scope."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from .complete_vae_loss import _run
from .parallel import GradSync


class _Adam:
    """mlx.optimizers.Adam as configured at trainer.py:75-76: betas (0.9, 0.999), eps 1e-8, NO bias correction.
    State lives in two flat buffers shaped like the module's flat parameter buffer."""

    def __init__(self, module, learning_rate: float):
        self.module = module
        self.learning_rate = float(learning_rate)
        self.m = torch.zeros_like(module.params.flat)
        self.v = torch.zeros_like(module.params.flat)

    def update(self, grad_scale: float = 1.0):
        p, g = self.module.params.flat, self.module.grads.flat
        _lib.check(_lib.load().arcvae_adam_step(p.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                                p.numel(), self.learning_rate, 0.9, 0.999, 1e-8, float(grad_scale),
                                                _lib.stream_ptr()))

    @property
    def state(self):
        return {"m": self.m, "v": self.v}


class ARCVAETrainerWithLoss:
    """Constructor signature of trainer.py:20-37.  ``train_step`` is the body of the batch loop (trainer.py:303-333)."""

    def __init__(self, encoder, decoder, property_predictor, dataset, learning_rate: float = 1e-4,
                 batch_size: int = 32, beta_start: float = 0.0, beta_end: float = 0.4, beta_warmup_epochs: int = 100,
                 lambda_prop: float = 0.1, lambda_collapse: float = 0.01, free_bits: float = 0.5,
                 lambda_mi: float = 0.01, grad_clip: float = 1.0, checkpoint_dir: str = "./checkpoints", *,
                 clip_mode: str = "reference_noop", pad_mask: bool = False, process_group=None,
                 broadcast_parameters: bool = True, error_check_every: int = 64):
        if clip_mode not in ("reference_noop", "global_norm"):
            raise ValueError("clip_mode must be 'reference_noop' or 'global_norm'")
        self.encoder, self.decoder, self.property_predictor, self.dataset = encoder, decoder, property_predictor, dataset
        self.batch_size, self.grad_clip = batch_size, grad_clip
        self.lambda_prop, self.lambda_collapse = lambda_prop, lambda_collapse
        self.free_bits, self.lambda_mi = free_bits, lambda_mi
        self.beta_start, self.beta_end, self.beta_warmup_epochs = beta_start, beta_end, beta_warmup_epochs
        self.learning_rate = learning_rate
        self.encoder_optimizer = _Adam(encoder, learning_rate)   # two optimizers, as trainer.py:75-76
        self.decoder_optimizer = _Adam(decoder, learning_rate)
        self.clip_mode, self.pad_mask = clip_mode, pad_mask
        self.sync = GradSync(process_group)
        self.step_count = 0
        self.checkpoint_dir = checkpoint_dir                      # created on the first save (trainer.py:80-81 creates it eagerly)
        self.history = {k: [] for k in ("epoch", "train_loss", "train_recon", "train_kl", "train_collapse", "train_prop",
                                        "val_loss", "val_recon", "val_kl", "val_collapse", "val_prop", "beta",
                                        "teacher_forcing", "learning_rate", "mutual_info")}       # trainer.py:84-100
        self.error_check_every = int(error_check_every)
        if self.sync.enabled and broadcast_parameters:
            # replicas must start from ONE model: rank 0's parameters and Adam moments win (a seed=None construction
            # initialises every rank differently)
            for t in (encoder.params.flat, decoder.params.flat, self.encoder_optimizer.m, self.encoder_optimizer.v,
                      self.decoder_optimizer.m, self.decoder_optimizer.v):
                self.sync.broadcast(t)

    # ---- schedules (trainer.py:102-114) -----------------------------------------------------------------
    def compute_beta(self, epoch: int) -> float:
        if epoch < self.beta_warmup_epochs:
            return float(self.beta_start + (self.beta_end - self.beta_start) * (epoch / self.beta_warmup_epochs))
        return float(self.beta_end)

    def compute_teacher_forcing_ratio(self, epoch: int, total_epochs: int) -> float:
        return float(max(0.5, 0.9 - 0.4 * (epoch / total_epochs)))

    # ---- one optimizer step (trainer.py:305-333) ----------------------------------------------------------
    def train_step(self, molecules: torch.Tensor, conditions: torch.Tensor, beta: float, teacher_forcing_ratio: float,
                   *, eps: Optional[torch.Tensor] = None, tf_mask=None, seed: Optional[int] = None,
                   global_row_start: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """value_and_grad -> clip -> two Adam updates.  Returns the loss dict (0-d CUDA tensors; nothing is synced).

        Clipping: the reference's ``_clip_gradients`` sums only arrays at the top level of each gradient dict, every
        leaf is one level down, so the norm is 0 and nothing is ever scaled (trainer.py:502-514).
        ``clip_mode='reference_noop'`` reproduces that; ``'global_norm'`` is a real global-norm clip (one host sync).

        Data parallelism: ``molecules`` is this rank's shard.  Every scalar of the returned dict and every gradient is
        that of the GLOBAL batch (what a single device would compute on the concatenated shards):
          * KL / MI / penalties come from the all-reduced batch statistics inside the loss kernel;
          * the kernel reports recon = (this rank's CE sum) / (global token count); the partial rides in the aux tail of
            the decoder's flat gradient buffer through the gradient all-reduce and is summed there (on the fused bf16
            path the CE sum is part of the all-reduced statistics instead; shards must then be EQUAL in size, since the
            d logits are scaled by 1 / (world x B x T) before the statistics are exchanged);
          * the Philox draw of eps is indexed by the GLOBAL row (``global_row_start``, default rank x local batch: even
            shards), so replicas do not repeat each other's noise and N ranks reproduce the single-device draw."""
        enc, dec, sync = self.encoder, self.decoder, self.sync
        enc.zero_grad()
        dec.zero_grad()
        if seed is None:
            seed = self.step_count
        dp = sync.enabled
        if global_row_start is None:
            global_row_start = sync.rank * int(molecules.shape[0])
        # forward + loss + decoder reverse pass; under DP the decoder gradients start their all-reduce while the
        # encoder BPTT is still being computed
        hooks = _DPHooks(sync) if dp else None
        d = _run(enc, dec, self.property_predictor, molecules, conditions, beta, self.lambda_prop, self.lambda_collapse,
                 teacher_forcing_ratio, self.free_bits, self.lambda_mi, 4.85, eps, tf_mask, seed, self.pad_mask, True,
                 sync.allreduce_stats if dp else None, backward_hooks=hooks,
                 eps_offset=int(global_row_start) * enc.latent_dim, ce_world=sync.world)
        if dp:
            sync.allreduce_async(enc.grads.flat)
            sync.wait()
            recon = dec.grads.aux[0].clone()                       # sum over ranks of local_CE / global_tokens
            d["recon_loss"] = recon
            d["total_loss"] = recon + d["weighted_kl"] + d["collapse_penalty"] + d["weighted_prop_loss"] + d["mi_penalty"]
        scale = 1.0
        if self.clip_mode == "global_norm" and self.grad_clip > 0:
            acc = torch.zeros(1, dtype=torch.float64, device=enc.device)
            lib = _lib.load()
            for m in (enc, dec):
                _lib.check(lib.arcvae_sumsq(m.grads.flat.data_ptr(), m.grads.flat.numel(), acc.data_ptr(), _lib.stream_ptr()))
            norm = float(acc.sqrt().item())
            if norm > self.grad_clip:
                scale = self.grad_clip / (norm + 1e-8)
        self.encoder_optimizer.update(scale)     # trainer.py:320
        self.decoder_optimizer.update(scale)     # trainer.py:324
        self.step_count += 1
        if self.error_check_every > 0 and self.step_count % self.error_check_every == 0:
            self.check_device_error()
        return d

    def check_device_error(self):
        """The persistent cluster kernels wait with bounded spins; a protocol time-out raises a STICKY per-device flag
        (never cleared by the library) and the Adam kernel refuses to apply gradients while it is set, so the weights
        are not corrupted.  This reads the flag (one 4-byte D2H, synchronises the stream) and raises."""
        _lib.check_device_error()

    # ---- forward-only evaluation without teacher forcing (trainer.py:116-175, :418-487) --------
    def _eval_batches(self, dataset, beta: float, max_batches: Optional[int]) -> Dict[str, float]:
        from .complete_vae_loss import complete_vae_loss
        keys = (("loss", "total_loss"), ("recon", "recon_loss"), ("kl", "kl_loss"), ("collapse", "collapse_penalty"),
                ("prop", "prop_loss"))
        acc = None
        n = 0
        for batch_idx, (molecules, conditions) in enumerate(dataset.to_batches(self.batch_size, shuffle=False)):
            if max_batches is not None and batch_idx >= max_batches:
                break
            d = complete_vae_loss(self.encoder, self.decoder, self.property_predictor, molecules, conditions, beta=beta,
                                  lambda_prop=self.lambda_prop, lambda_collapse=self.lambda_collapse,
                                  teacher_forcing_ratio=0.0, free_bits=self.free_bits, lambda_mi=self.lambda_mi,
                                  target_mi=4.85, seed=self.step_count + batch_idx, pad_mask=self.pad_mask)
            v = torch.stack([d[k].detach().double() for _, k in keys])   # stays on the device: one sync at the end
            acc = v if acc is None else acc + v
            n += 1
        if n == 0:
            return {name: 0.0 for name, _ in keys}
        host = (acc / n).cpu().tolist()
        return {name: float(host[i]) for i, (name, _) in enumerate(keys)}

    def _compute_true_train_loss(self, epoch: int, num_batches: int = 10) -> Dict[str, float]:
        """trainer.py:116-175: the first ``num_batches`` unshuffled training batches, teacher_forcing_ratio = 0.0."""
        return self._eval_batches(self.dataset, self.compute_beta(epoch), num_batches)

    def _validate(self, val_dataset, beta: float) -> Dict[str, float]:
        """trainer.py:418-487: every validation batch, teacher_forcing_ratio = 0.0 (greedy feedback chain), means of the
        per-batch loss terms."""
        return self._eval_batches(val_dataset, beta, None)

    # ---- checkpoint interop (trainer.py:577-603, :685-736) ---------------------------------------
    def save_checkpoint(self, epoch: int, is_best: bool = False) -> str:
        """trainer.py:577-597 call surface: writes ``<checkpoint_dir>/checkpoint_epoch_{epoch:03d}.npz`` and, with
        ``is_best``, ``<checkpoint_dir>/checkpoint_best.npz``; returns the epoch file's path.  The FILE FORMAT is the flat
        one of ``save_checkpoint_to`` (the reference pickles nested dicts of ``mx.array``)."""
        import os
        os.makedirs(str(self.checkpoint_dir), exist_ok=True)
        if is_best:
            self.save_checkpoint_to(os.path.join(str(self.checkpoint_dir), "checkpoint_best.npz"), epoch)
        return self.save_checkpoint_to(os.path.join(str(self.checkpoint_dir), f"checkpoint_epoch_{int(epoch):03d}.npz"), epoch)

    def save_checkpoint_to(self, path, epoch: int = 0) -> str:
        """Flat ``.npz`` with the reference's parameter names: ``encoder/<module>.<leaf>``,
        ``decoder/<module>.<leaf>``, Adam moments ``encoder_opt/m|v/<module>.<leaf>`` ..., ``epoch``.  The reference
        pickles nested dicts of ``mx.array`` (unreadable without MLX); ``tools/convert_mlx_checkpoint.py`` re-saves such
        a file in this flat format on a machine that has MLX."""
        import numpy as np
        import json
        out = {"epoch": np.int64(epoch), "format": np.array("arcvae-flat-v1"),
               "history_json": np.array(json.dumps(self.history))}
        for tag, mod, opt in (("encoder", self.encoder, self.encoder_optimizer), ("decoder", self.decoder, self.decoder_optimizer)):
            for name, view in mod.params.views.items():
                out[f"{tag}/{name}"] = view.detach().cpu().numpy()
            off = 0
            for name, view in mod.params.views.items():
                n = view.numel()
                beg = view.data_ptr() - mod.params.flat.data_ptr()
                beg //= 4
                out[f"{tag}_opt/m/{name}"] = opt.m[beg:beg + n].reshape(view.shape).cpu().numpy()
                out[f"{tag}_opt/v/{name}"] = opt.v[beg:beg + n].reshape(view.shape).cpu().numpy()
                off += n
        path = str(path)
        np.savez(path, **out)
        return path if path.endswith(".npz") else path + ".npz"

    def load_checkpoint(self, path) -> int:
        """Inverse of ``save_checkpoint``; returns the stored epoch (trainer.py:685-712).  Optimizer moments are optional
        (a converted reference checkpoint may carry weights only)."""
        import numpy as np
        ck = np.load(str(path), allow_pickle=False)
        for tag, mod, opt in (("encoder", self.encoder, self.encoder_optimizer), ("decoder", self.decoder, self.decoder_optimizer)):
            mod.load_parameters({name: ck[f"{tag}/{name}"] for name in mod.params.views})
            for name, view in mod.params.views.items():
                n = view.numel()
                beg = (view.data_ptr() - mod.params.flat.data_ptr()) // 4
                for mom, buf in (("m", opt.m), ("v", opt.v)):
                    key = f"{tag}_opt/{mom}/{name}"
                    if key in ck.files:
                        buf[beg:beg + n].copy_(torch.as_tensor(ck[key]).reshape(-1).to(buf.device))
        if "history_json" in ck.files:
            import json
            self.history = json.loads(str(ck["history_json"]))
        return int(ck["epoch"]) if "epoch" in ck.files else 0

    def _train_epoch_batches(self, beta: float, teacher_forcing_ratio: float) -> Dict[str, float]:
        """trainer.py:242-416 without the tqdm / logging side paths: iterate ``dataset.to_batches`` and step."""
        total, n = 0.0, 0
        for molecules, conditions in self.dataset.to_batches(self.batch_size, shuffle=True):
            d = self.train_step(molecules, conditions, beta, teacher_forcing_ratio)
            total += float(d["total_loss"])
            n += 1
        self.check_device_error()
        return {"loss": total / max(1, n)}


class _DPHooks:
    """Called by the step between the decoder and encoder reverse passes."""

    def __init__(self, sync: GradSync):
        self.sync = sync

    def after_decoder_backward(self, decoder, losses, recon_is_global=False):
        # this rank's partial reconstruction loss travels with the gradients (FlatParams.aux).  On the fused-CE path the
        # CE sum went through the statistics all-reduce, so recon is already the global value on every rank: each rank
        # contributes 1/world of it and the sum restores it
        r = losses["recon_loss"].reshape(1)
        decoder.grads.aux[0:1].copy_(r / self.sync.world if recon_is_global else r)
        self.sync.allreduce_async(decoder.grads.storage)
