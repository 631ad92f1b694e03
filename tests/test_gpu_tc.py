"""GPU tests of the tcgen05 + TMA bf16 GEMM (csrc/gemm_tc.cu) against torch fp64 on the same bf16-rounded operands.
The products of bf16 values are exact in fp32, so the only difference is fp32 accumulation order: tolerance 2e-4
relative to the largest output (a stated bf16-path tolerance is only needed where operands get ROUNDED to bf16)."""
import numpy as np
import pytest
import torch

from _util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import mlx_vae_b200
    return mlx_vae_b200


def run(M, a_mn, b_mn, m, n, k, bias=False, accumulate=False, splitk=1, bf16_out=False, seed=0):
    lib = M._lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed + m + 3 * n + 7 * k)
    A = torch.randn((k, m) if a_mn else (m, k), device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn((k, n) if b_mn else (n, k), device="cuda", generator=g).to(torch.bfloat16)
    bvec = torch.randn(n, device="cuda", generator=g) if bias else None
    C0 = torch.randn(m, n, device="cuda", generator=g)
    C = C0.clone()
    Cb = torch.zeros(m, n, device="cuda", dtype=torch.bfloat16) if bf16_out else None
    Ad = A.double().T if a_mn else A.double()
    Bd = B.double() if b_mn else B.double().T
    ref = Ad @ Bd
    if bias:
        ref = ref + bvec.double()
    if accumulate:
        ref = ref + C0.double()
    M._lib.check(lib.arcvae_gemm_bf16(int(a_mn), int(b_mn), m, n, k, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1],
                                      C.data_ptr(), n, Cb.data_ptr() if bf16_out else None, n,
                                      bvec.data_ptr() if bias else None, int(accumulate), splitk, 0))
    torch.cuda.synchronize()
    assert rel_err(C.cpu(), ref.cpu()) < 2e-4, (a_mn, b_mn, m, n, k)
    if bf16_out:
        assert rel_err(Cb.float().cpu(), ref.cpu()) < 1e-2


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (128, 16, 64), (256, 256, 256), (4096, 1024, 256), (300, 768, 256),
                                   (1000, 80, 256), (128, 256, 1024), (4096, 256, 1024), (77, 48, 80), (128, 512, 128)])
def test_nt_k_major(M, m, n, k):
    run(M, False, False, m, n, k)
    run(M, False, False, m, n, k, bias=True, bf16_out=True)
    run(M, False, False, m, n, k, accumulate=True)


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (1024, 256, 4096), (768, 256, 2048), (80, 256, 1000), (128, 64, 8192),
                                   (256, 256, 130)])
def test_tn_mn_major(M, m, n, k):
    run(M, True, True, m, n, k)
    run(M, True, True, m, n, k, accumulate=True)
    run(M, True, True, m, n, k, accumulate=True, splitk=4)


def test_splitk_large_k(M):
    run(M, True, True, 1024, 256, 128 * 512, accumulate=True, splitk=37)
    run(M, False, False, 256, 256, 8192, accumulate=True, splitk=8)


def test_many_tiles_persistent(M):
    run(M, False, False, 128 * 300 + 5, 1024, 256, bias=True)


@pytest.mark.parametrize("m,n,k,bias", [(4096, 1024, 256, True), (128 * 9 + 37, 256, 64, False), (5000, 512, 128, True),
                                        (128 * 300, 1024, 256, True), (1024, 768, 192, True)])
def test_weight_stationary_bf16_out(M, m, n, k, bias):
    """K <= 256 with a bf16-only output dispatches to gemm_ws_kernel (csrc/gemm_ws.cu): weight panel resident in shared
    memory, TMA-store epilogue.  Must equal the generic kernel (ARCVAE_NO_WS=1) bit for bit — same MMA order, same
    rounding — and torch fp64 within bf16 output rounding."""
    import os
    lib = M._lib.load()
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    A = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(n, k, device="cuda", generator=g).to(torch.bfloat16)
    bvec = torch.randn(n, device="cuda", generator=g) if bias else None
    outs = []
    for env in (False, True):
        if env:
            os.environ["ARCVAE_NO_WS"] = "1"
        try:
            Cb = torch.full((m, n), 7.0, device="cuda", dtype=torch.bfloat16)
            M._lib.check(lib.arcvae_gemm_bf16(0, 0, m, n, k, A.data_ptr(), k, B.data_ptr(), k, None, 0, Cb.data_ptr(), n,
                                              bvec.data_ptr() if bias else None, 0, 1, 0))
            torch.cuda.synchronize()
            outs.append(Cb)
        finally:
            os.environ.pop("ARCVAE_NO_WS", None)
    ref = A.double() @ B.double().T + (bvec.double() if bias else 0.0)
    assert rel_err(outs[0].float().cpu(), ref.cpu()) < 1e-2
    assert torch.equal(outs[0], outs[1]), float((outs[0].float() - outs[1].float()).abs().max())
