"""GPU tests (-m gpu) of the tcgen05 building blocks of the training backward (train_tc.cu) against torch matmuls on the same
bf16 operands (fp32 accumulate): the row GEMM with its three epilogues and the MN-major weight-gradient GEMM."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _report(name, text):
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", name), "a") as f:
        f.write(text + "\n")


@pytest.mark.parametrize("N", [128, 256, 64])
def test_rows_gemm_tc_matches_torch(engine, N):
    g = torch.Generator(device="cuda").manual_seed(N)
    M = 128 * 37
    A = (torch.randn(M, 128, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, 128, generator=g, device="cuda") * 0.1).to(torch.bfloat16)
    b = torch.randn(N, generator=g, device="cuda") * 0.1
    ref = A.float() @ W.float().t()
    out = engine.train_tc_gemm(A, W, b, epi=0).float()
    exp = torch.relu(ref + b)
    err0 = float((out - exp).abs().max())
    out1 = engine.train_tc_gemm(A, W, b, epi=1).float()
    err1 = float((out1 - (ref + b)).abs().max())
    mask = (torch.randn(M, N, generator=g, device="cuda")).to(torch.bfloat16)
    mask[::7] = 0
    out2 = engine.train_tc_gemm(A, W, None, epi=2, mask=mask).float()
    exp2 = ref * (mask.float() > 0)
    err2 = float((out2 - exp2).abs().max())
    # in place over the mask (the backward overwrites H1 with dZ1)
    m2 = mask.clone()
    _lib = engine.lib
    from mpinets_b200 import _lib as L
    import ctypes as C
    L.check(_lib.mpn_train_tc_gemm(engine._ctx, engine.stream, 2, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), None,
                                   C.c_void_p(m2.data_ptr()), M, N, C.c_void_p(m2.data_ptr())))
    err3 = float((m2.float() - exp2).abs().max())
    torch.cuda.synchronize()
    _report("train_tc.txt", f"rows_gemm N={N}: relu {err0:.3e} plain {err1:.3e} mask {err2:.3e} inplace {err3:.3e} tc_error {engine.tc_error()}")
    tol = 2e-2 * float(ref.abs().max())       # bf16 rounding of the outputs
    assert not engine.tc_error()
    assert err0 < tol and err1 < tol and err2 < tol and err3 < tol


def test_wgrad_tc_matches_torch(engine):
    g = torch.Generator(device="cuda").manual_seed(7)
    res = {}
    for R in (64, 128 * 5 + 64, 128 * 1000 + 8):
        dY = (torch.randn(R, 128, generator=g, device="cuda") * 0.3).to(torch.bfloat16)
        X = (torch.randn(R, 128, generator=g, device="cuda") * 0.3).to(torch.bfloat16)
        ref = dY.float().t() @ X.float()
        for variant in (0, 1):
            out = engine.train_tc_wgrad(dY, X, variant)
            torch.cuda.synchronize()
            res[(R, variant)] = float((out - ref).abs().max()) / float(ref.abs().max())
            _report("train_tc.txt", f"wgrad R={R} variant={variant}: rel err {res[(R, variant)]:.3e} tc_error {engine.tc_error()}")
    assert all(res[(R, 0)] < 1e-4 for R in (64, 128 * 5 + 64, 128 * 1000 + 8)), res
