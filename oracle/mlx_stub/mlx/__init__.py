"""TEST INFRASTRUCTURE — a minimal stand-in for Apple MLX, backed by torch CPU tensors.

Purpose: let the UNMODIFIED reference sources under /root/reference (models/*.py, losses/*.py, complete_vae_loss.py,
trainer.py, mlx_data/dataloader.py) be imported and executed in this image, where the real ``mlx`` wheel is not
installable, so that the CPU oracle (oracle/arcvae_oracle.py) can be pinned to the reference's OWN code running
(oracle/make_ref_golden.py, tests/test_ref_pin.py).

What this pins and what it does not: every line of the reference's Python (model composition, loss arithmetic, the
no-op clip, the trainer batch loop, the dataset) runs as written.  The MLX PRIMITIVES it calls (``nn.LSTM``,
``nn.Linear``, ``nn.Embedding``, ``optimizers.Adam``, ``mx.maximum`` / ``mx.clip`` VJPs, ``mx.argmax`` tie-break,
``mx.value_and_grad`` over ``nn.Module`` trees) are restated here from MLX's published behaviour,
not taken from an MLX binary; that residual risk stays until the reference is run under real MLX.

Only ``oracle/`` scripts and ``tests/`` may import this package; it is never on the product path."""
__version__ = "0.0-stub"
