"""GPU tests (-m gpu) of the calibration gradients (llama-fft_b200/csrc/bcl_backward.cu + autograd.py) against torch autograd
through the restated reference forward (oracle.bc_forward_torch, i.e. what src/calibrate_bc.py:281-294 differentiates) and
against the FFT-free fp64 oracle."""
import numpy as np
import pytest
import torch

import llama_fft_b200 as bc
from oracle import bc_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("p,q,B,N", [(2, 3, 64, 5), (3, 2, 128, 7), (2, 2, 256, 3), (5, 7, 64, 33), (32, 32, 128, 16),
                                     (5, 7, 64, 300), (8, 6, 128, 517), (3, 4, 256, 260)])      # >= 256 tokens: split tensor-core chain
@pytest.mark.parametrize("with_bias", [False, True])
def test_gradients_match_autograd_of_reference_forward_fp32(p, q, B, N, with_bias):
    g = torch.Generator().manual_seed(p * 7 + q + B + N)
    c = torch.randn(p, q, B, generator=g) * 0.1
    x = torch.randn(N, q * B, generator=g)
    gy = torch.randn(N, p * B, generator=g)
    bias = torch.randn(p * B, generator=g) if with_bias else None
    # reference side: autograd through the torch.fft formulation, CPU fp32
    cr, xr = c.clone().requires_grad_(True), x.clone().requires_grad_(True)
    br = None if bias is None else bias.clone().requires_grad_(True)
    O.bc_forward_torch(cr, xr, br).backward(gy)
    # this repo
    m = bc.BlockCirculantLinear(q * B, p * B, block_size=B, bias=with_bias)
    with torch.no_grad():
        m.c.copy_(c)
        if with_bias:
            m.bias.copy_(bias)
    m = m.to(DEV).train()
    xd = x.to(DEV).requires_grad_(True)
    y = m(xd)
    assert y.requires_grad
    y.backward(gy.to(DEV))
    assert rel(y.detach(), O.bc_forward_torch(c, x, bias)) <= 1e-4
    assert rel(xd.grad, xr.grad) <= 1e-4
    assert rel(m.c.grad, cr.grad) <= 1e-4
    if with_bias:
        assert rel(m.bias.grad, br.grad) <= 1e-5
    if p * q * B * B <= 1 << 22:
        gx64, gc64, gb64 = O.bc_backward_dense64(c.numpy(), x.numpy(), gy.numpy())
        assert rel(xd.grad, torch.from_numpy(gx64)) <= 1e-4 and rel(m.c.grad, torch.from_numpy(gc64)) <= 1e-4


def test_gradients_3d_fp16_activations_and_frozen_parts():
    """fp16 activations with fp32 c (what calibrate_bc trains: src/calibrate_bc.py:55-70 unfreezes only c / bias)."""
    g = torch.Generator().manual_seed(0)
    p, q, B = 4, 3, 128
    c = torch.randn(p, q, B, generator=g) * 0.1
    x = torch.randn(2, 9, q * B, generator=g).half()
    m = bc.BlockCirculantLinear(q * B, p * B, block_size=B, bias=False)
    with torch.no_grad():
        m.c.copy_(c)
    m = m.to(DEV).train()
    y = m(x.to(DEV))                                        # x does not require grad: only dc is computed
    assert y.dtype == torch.float16 and y.shape == (2, 9, p * B)
    loss = (y.float() ** 2).mean()
    loss.backward()
    cr = c.clone().requires_grad_(True)
    yr = O.bc_forward_torch(cr, x)
    (yr.float() ** 2).mean().backward()
    assert rel(m.c.grad, cr.grad) <= 5e-3                   # fp16 y on both sides: gradients flow through a 16-bit rounding
    m.c.requires_grad_(False)
    with torch.no_grad():
        y2 = m.eval()(x.to(DEV))
    assert not y2.requires_grad


def test_optimizer_step_invalidates_cached_spectrum():
    """AdamW on c (src/calibrate_bc.py:257-294) then an eval forward must see the NEW c (spectrum cache is version-keyed)."""
    g = torch.Generator().manual_seed(1)
    m = bc.BlockCirculantLinear(256, 128, block_size=64, bias=True).to(DEV)
    x = torch.randn(6, 256, generator=g).to(DEV)
    m.eval()
    with torch.no_grad():
        y0 = m(x).clone()
    m.train()
    opt = torch.optim.AdamW([m.c, m.bias], lr=1e-2)
    for _ in range(3):
        opt.zero_grad()
        (m(x) ** 2).mean().backward()
        opt.step()
    m.eval()
    with torch.no_grad():
        y1 = m(x)
    ref = O.bc_forward_torch(m.c.detach().cpu(), x.cpu(), m.bias.detach().cpu())
    assert rel(y1, ref) <= 1e-4 and rel(y1, y0) > 1e-3


def test_backward_weights_accumulate_and_validation():
    import ctypes
    from llama_fft_b200 import _cabi
    L = _cabi.lib()
    g = torch.Generator().manual_seed(2)
    p, q, B, N = 3, 3, 64, 10
    x = torch.randn(N, q * B, generator=g).to(DEV)
    gy = torch.randn(N, p * B, generator=g).to(DEV)
    ws = torch.empty(L.bcl_backward_workspace_bytes(N, p, q, B), device=DEV, dtype=torch.uint8)
    dc = torch.zeros(p, q, B, device=DEV)
    st = _cabi.stream_ptr(DEV)
    for half in (slice(0, 4), slice(4, N)):                 # two token chunks accumulated == one call
        xs, gs = x[half], gy[half]
        rc = L.bcl_backward_weights(xs.data_ptr(), _cabi.F32, xs.stride(0), gs.data_ptr(), _cabi.F32, gs.stride(0), xs.shape[0], p, q, B,
                                    dc.data_ptr(), 1, ws.data_ptr(), ws.numel(), st)
        _cabi.check(rc, "bcl_backward_weights")
    _, gc64, _ = O.bc_backward_dense64(np.zeros((p, q, B)), x.cpu().numpy(), gy.cpu().numpy())
    assert rel(dc, torch.from_numpy(gc64)) <= 1e-4
    assert L.bcl_backward_weights(x.data_ptr(), _cabi.F32, x.stride(0), gy.data_ptr(), _cabi.F32, gy.stride(0), N, p, q, B,
                                  dc.data_ptr(), 0, ws.data_ptr(), 16, st) == -6
