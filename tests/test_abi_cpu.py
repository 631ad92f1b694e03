"""CPU tests of the C-ABI boundary: the shared library builds for sm_100a, loads without a GPU, exports every symbol
include/arcvae_b200.h declares, the ctypes binding covers them all, and the product refuses to run without a GPU
(no CPU fallback)."""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HEADER = os.path.join(ROOT, "include", "arcvae_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(arcvae_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    import mlx_vae_b200
    return mlx_vae_b200._lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    import mlx_vae_b200
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/arcvae_b200.h but not exported"
        assert n in mlx_vae_b200._lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.arcvae_abi_version() == 2


def test_sizing_entry_points_run_without_a_gpu(lib):
    import mlx_vae_b200
    d = mlx_vae_b200._lib.Dims(V=80, E=128, H=256, L=128, C=1, NL=2, pad_token=0, end_token=2)
    tape = lib.arcvae_encoder_tape_bytes(d, 4096, 128)
    assert 6e9 < tape < 12e9                       # 2 layers x (gates 2.1 GB + c + h) + bf16 copies
    assert lib.arcvae_decoder_tape_bytes(d, 4096, 128) > 2e9
    assert lib.arcvae_sampler_workspace_bytes(d, 1024, 128) > 0
    assert lib.arcvae_launch_count() == 0


def test_no_cpu_fallback(lib):
    import torch
    import mlx_vae_b200 as M
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(M._lib.ArcvaeError):
        M.MLXEncoder(vocab_size=11, embedding_dim=8, hidden_dim=16, latent_dim=8, num_conditions=1, num_layers=2)
    with pytest.raises(M._lib.ArcvaeError):
        M.reconstruction_loss(torch.zeros(2, 3, 4), torch.zeros(2, 3, dtype=torch.int32))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "mlx-vae_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                for banned in ("arcvae_oracle", "philox_ref", "ref_runner", "mlx_stub", "dataset_oracle", "import mlx"):
                    assert banned not in src, (f, banned)
