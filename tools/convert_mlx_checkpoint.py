#!/usr/bin/env python3
"""Run ON A MACHINE THAT HAS MLX (the reference's environment): re-save a checkpoint written by the reference's
``ARCVAETrainerWithLoss.save_checkpoint`` (trainer.py:577-603: ``np.savez`` of nested dicts of ``mx.array``, i.e. pickled
0-d object arrays that cannot be read without MLX) into the flat ``arcvae-flat-v1`` ``.npz`` that
``mlx_vae_b200.ARCVAETrainerWithLoss.load_checkpoint`` reads.

    python tools/convert_mlx_checkpoint.py checkpoints/checkpoint_epoch_030.npz flat_epoch_030.npz

Key mapping: checkpoint['encoder_weights'][module][leaf] -> 'encoder/<module>.<leaf>' (same for the
decoder).  MLX Adam state is keyed like the parameter tree with leaves {'m', 'v'} -> '<tag>_opt/m|v/<module>.<leaf>'.
It only uses np.load / np.savez and np.array() on mx.array values.  Exercised in this repository on a checkpoint
written by the reference's own ``save_checkpoint`` running under the ``mlx`` stand-in (tests/golden/ref_checkpoint_tiny.npz,
tests/test_ref_pin.py::test_reference_checkpoint_converts_to_the_flat_format); real MLX arrays convert through the same
``np.array(v)`` call."""
import sys

import numpy as np


def flatten(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        name = f"{prefix}.{k}" if prefix else str(k)
        if isinstance(v, dict):
            out.update(flatten(v, name))
        else:
            out[name] = np.array(v)          # mx.array -> numpy
    return out


def main(src, dst):
    ck = np.load(src, allow_pickle=True)
    out = {"epoch": np.int64(int(ck["epoch"])) if "epoch" in ck.files else np.int64(0), "format": np.array("arcvae-flat-v1")}
    for tag in ("encoder", "decoder"):
        for name, arr in flatten(ck[f"{tag}_weights"].item()).items():
            out[f"{tag}/{name}"] = arr.astype(np.float32)
        key = f"{tag}_optimizer_state"
        if key in ck.files:
            state = ck[key].item()
            for name, arr in flatten(state).items():     # e.g. 'fc_mu.weight.m' / 'fc_mu.weight.v' ; 'step', 'learning_rate' are skipped
                if name.endswith(".m") or name.endswith(".v"):
                    out[f"{tag}_opt/{name[-1]}/{name[:-2]}"] = arr.astype(np.float32)
    if "history" in ck.files:                        # trainer.py:585 — the 15-list history dict
        import json
        hist = ck["history"].item()
        out["history_json"] = np.array(json.dumps({k: [float(x) for x in v] for k, v in hist.items()}))
    np.savez(dst, **out)
    print(f"wrote {dst}: {len(out)} arrays")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
