"""Pins the tcgen05 pointwise engine (csrc/tc.cuh): UMMA descriptor encodings, TMEM addressing, the
A-from-TMEM / B-from-shared MMA and the 3xTF32 error compensation, against an fp64-accumulated reference."""
import numpy as np
import pytest
import torch

import yfv2  # noqa: F401
import yfv2_engine as eng

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(24, 24), (48, 48), (96, 96), (72, 72), (72, 83)])
@pytest.mark.parametrize("P", [128, 1000, 128 * 37])
def test_pw_tc_matches_fp32(K, N, P):
    g = torch.Generator().manual_seed(K * 1000 + N + P)
    x = torch.randn(K, P, generator=g)
    w = torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)
    ref = (w.double() @ x.double())
    out = eng.debug_pw_tc(x.cuda(), w.cuda()).cpu()
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    # fp32 FFMA accumulation of K terms gives ~K * 6e-8 * scale; 3xTF32 must be in that class (single-pass TF32 is ~1e-3)
    assert err <= 2e-6 * scale * max(1.0, K / 24), (err, scale)


def test_pw_tc_exact_on_tf32_representable_inputs():
    """Integers < 2^10 are exact in TF32 and their dot products exact in fp32: any layout/addressing mistake shows
    up as a wrong integer, not as rounding noise."""
    K, N, P = 48, 48, 256
    g = torch.Generator().manual_seed(3)
    x = torch.randint(-8, 9, (K, P), generator=g).float()
    w = torch.randint(-8, 9, (N, K), generator=g).float()
    out = eng.debug_pw_tc(x.cuda(), w.cuda()).cpu()
    assert torch.equal(out, w @ x)
