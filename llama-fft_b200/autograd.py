"""Differentiable BC linear for calibration.

reference: torch autograd through ``BlockCirculantLinear.forward`` (src/patch_llama_fft.py:226-275), which
``src/calibrate_bc.py:281-294`` uses to fine-tune ``c`` (and optionally ``bias``) against teacher logits.  Here the forward
and both gradients run on the same hand-written CUDA kernels (include/bclinear.h "gradients for calibration"):

    dx    = BC forward of gy with ct[i,j,t] = c[j,i,(-t) mod B]                 (bcl_conj_transpose + bcl_forward)
    dc    = sum over tokens of the circular cross-correlation of gy and x       (bcl_backward_weights)
    dbias = gy.sum(0)

In grad mode the spectrum is rebuilt on every call (reference rule, src/bench_all_in_one.py:212-219) and always in the fp32
accuracy class, whatever ``spectrum_dtype`` the layer serves inference with: the CUDA-core stream kernel for a few tokens, the
split-bf16 tensor-core chain (GEMM layout, spec_dtype f32) from ``prefill_token_floor`` tokens -- calibration batches are
hundreds to thousands of tokens.  No CPU path: CPU tensors raise.
"""
from __future__ import annotations

import torch

from . import _cabi


def _spectrum_f32(c32: torch.Tensor) -> torch.Tensor:
    p, q, B = c32.shape
    spec = torch.empty_like(c32)
    with torch.cuda.device(c32.device):
        rc = _cabi.lib().bcl_precompute_spectrum(c32.data_ptr(), p, q, B, spec.data_ptr(), _cabi.F32, _cabi.LAYOUT_STREAM,
                                                 _cabi.stream_ptr(c32.device))
    _cabi.check(rc, "bcl_precompute_spectrum")
    return spec


def _forward_f32(x_flat: torch.Tensor, c32: torch.Tensor, bias) -> torch.Tensor:
    """One fp32-class BC forward with a freshly built spectrum of c32 (p, q, B)."""
    from .patch_llama_fft import bc_forward, prefill_token_floor
    p, q, B = c32.shape
    if x_flat.shape[0] >= prefill_token_floor(p, q, torch.float32):
        L = _cabi.lib()
        spec = torch.empty(L.bcl_spectrum_bytes(p, q, B, _cabi.F32, _cabi.LAYOUT_GEMM), device=c32.device, dtype=torch.uint8)
        with torch.cuda.device(c32.device):
            rc = L.bcl_precompute_spectrum(c32.data_ptr(), p, q, B, spec.data_ptr(), _cabi.F32, _cabi.LAYOUT_GEMM, _cabi.stream_ptr(c32.device))
        _cabi.check(rc, "bcl_precompute_spectrum")
        return bc_forward(x_flat, spec, bias, p, q, B, layout=_cabi.LAYOUT_GEMM, spectrum_dtype=torch.float32)
    return bc_forward(x_flat, _spectrum_f32(c32), bias, p, q, B)


def _aligned(t: torch.Tensor) -> torch.Tensor:
    if t.stride(-1) != 1 or (t.stride(0) * t.element_size()) % 16 or t.data_ptr() % 16:
        return t.contiguous()
    return t


class _BCLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_flat, c, bias):
        _cabi.require_cuda(c, "BlockCirculantLinear.c")
        c32 = c.detach().to(torch.float32).contiguous()
        y = _forward_f32(x_flat.detach(), c32, None if bias is None else bias.detach())
        ctx.save_for_backward(x_flat.detach(), c32)
        ctx.meta = (c.dtype, None if bias is None else bias.dtype, x_flat.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        x_flat, c32 = ctx.saved_tensors
        c_dtype, bias_dtype, x_dtype = ctx.meta
        p, q, B = c32.shape
        L = _cabi.lib()
        dev = gy.device
        N = gy.shape[0]
        # y = y32.to(x.dtype) (+ bias): the gradient that reaches y32 is gy cast to x.dtype first (autograd of .to / +)
        g_act = _aligned(gy.to(x_dtype))
        gx = gc = gb = None
        if ctx.needs_input_grad[0]:
            ct = torch.empty(q, p, B, device=dev, dtype=torch.float32)
            with torch.cuda.device(dev):
                _cabi.check(L.bcl_conj_transpose(c32.data_ptr(), p, q, B, ct.data_ptr(), _cabi.stream_ptr(dev)), "bcl_conj_transpose")
            gx = _forward_f32(g_act, ct, None)                               # (N, in) in x.dtype
        if ctx.needs_input_grad[1]:
            xa = _aligned(x_flat)
            dc = torch.empty(p, q, B, device=dev, dtype=torch.float32)
            ws = torch.empty(L.bcl_backward_workspace_bytes(max(N, 1), p, q, B), device=dev, dtype=torch.uint8)
            with torch.cuda.device(dev):
                rc = L.bcl_backward_weights(xa.data_ptr(), _cabi.DTYPE[xa.dtype], xa.stride(0), g_act.data_ptr(), _cabi.DTYPE[g_act.dtype],
                                            g_act.stride(0), N, p, q, B, dc.data_ptr(), 0, ws.data_ptr(), ws.numel(), _cabi.stream_ptr(dev))
            _cabi.check(rc, "bcl_backward_weights")
            gc = dc.to(c_dtype)
        if bias_dtype is not None and ctx.needs_input_grad[2]:
            gb = gy.sum(0).to(bias_dtype)
        return gx, gc, gb


def bc_linear_autograd(layer, x_flat: torch.Tensor) -> torch.Tensor:
    _cabi.require_cuda(x_flat, "input x")
    return _BCLinearFn.apply(x_flat, layer.c, layer.bias)
