#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — writes tests/golden/ref_<case>.npz: outputs of the UNMODIFIED reference sources
(/root/reference) executed under the ``mlx`` stand-in (oracle/mlx_stub), in fp64.

Run in the authoring container (the reference tree is mounted read-only there; it does not exist on the GPU box):
    python oracle/make_ref_golden.py
The fixtures pin oracle/arcvae_oracle.py and oracle/dataset_oracle.py to the reference's own code
(tests/test_ref_pin.py); the GPU parity tests then compare the CUDA path with that oracle."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_runner as R  # noqa: E402


def main():
    if not R.reference_available():
        sys.exit("/root/reference is not mounted here: the fixtures can only be regenerated in the authoring container")
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    for name in R.CASES:
        ck = os.path.join(out_dir, "ref_checkpoint_tiny.npz") if name == "tiny" else None
        res = R.run_reference(name, checkpoint_to=ck)
        path = os.path.join(out_dir, f"ref_{name}.npz")
        np.savez_compressed(path, **res)
        print(f"{path}: {len(res)} entries, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
