"""Chain forward: DEPENDENT block-circulant linears of one token in ONE kernel launch (``bcl_forward_chain``), and the whole-MLP
fusion built on it.

The reference swaps ``mlp.gate_proj / up_proj / down_proj`` (src/patch_llama_fft.py:379) and lets HF's ``LlamaMLP.forward`` chain them:
``down_proj(act_fn(gate_proj(x)) * up_proj(x))`` -- three module calls, two elementwise kernels.  At batch-1 decode that is five
launches of ~5 us latency each for ~0.5 us of HBM time.  ``fuse_mlp(model)`` keeps the three ``BlockCirculantLinear`` modules where
they are (same state-dict keys, ``isinstance`` checks, checkpoint I/O, calibration) and replaces the *container's* forward for
single-token eval / no-grad calls by one launch:

    step 0   gate and up stacked along the output-block axis (one rfft(x)), SWIGLU epilogue: h = silu(round(gate)) * round(up)
    ------   grid-wide barrier (the spectrum of down_proj is already streaming into shared memory)
    step 1   y = down_proj(h)

Calls with MANY tokens (prefill: enough work for the tensor-core path, 16-bit activations, bf16 spectra, B <= 128) take
``bcl_forward_swiglu``: gate and up stacked, one forward-DFT pass and one frequency-GEMM pass for both, and an inverse stage whose
epilogue writes h = silu(gate) * up directly -- the two 11008-wide projection outputs never reach HBM and HF's act_fn / multiply
kernels disappear -- followed by down_proj's own forward.

Every other call (grad mode, CPU tensors, a bias, another activation function, fp32 activations) takes the original forward.
"""
from __future__ import annotations

import ctypes
import types
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _cabi
from .patch_llama_fft import BlockCirculantLinear, prefill_workspace
from .shared_input import SharedInputGroup

_counters: Dict[Tuple[int, int], torch.Tensor] = {}


def barrier_counter(device) -> torch.Tensor:
    """The grid-barrier counter of the chain launches on the CURRENT stream of `device` (8 bytes, zeroed once, then owned by the
    library: include/bclinear.h).  One per (device, stream): launches of one stream never overlap, launches of different streams might."""
    dev = device if isinstance(device, torch.device) else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, _cabi.stream_ptr(dev))
    c = _counters.get(key)
    if c is None:
        c = torch.zeros(1, device=dev, dtype=torch.int64)
        _counters[key] = c
    return c


def _problem(x: torch.Tensor, spec: torch.Tensor, y: torch.Tensor, epilogue: int, bias: Optional[torch.Tensor] = None,
             y_spectrum: Optional[torch.Tensor] = None, x_is_spectrum: bool = False) -> _cabi.Problem:
    """x_is_spectrum: x is the (1, in) float32 packed spectrum an earlier step of the chain emitted (EPILOGUE_EMIT_SPECTRUM)."""
    p, q = int(spec.shape[0]), int(spec.shape[1])
    if y_spectrum is not None:
        epilogue |= _cabi.EPILOGUE_EMIT_SPECTRUM
    if x_is_spectrum and x.dtype != torch.float32:
        raise TypeError("an emitted spectrum is float32")
    return _cabi.Problem(x.data_ptr(), spec.data_ptr(), None if bias is None else bias.data_ptr(), y.data_ptr(),
                         1, x.stride(0), y.stride(0), p, q, _cabi.X_SPECTRUM if x_is_spectrum else _cabi.DTYPE[x.dtype], _cabi.DTYPE[y.dtype],
                         0 if bias is None else _cabi.DTYPE[bias.dtype], epilogue, None if y_spectrum is None else y_spectrum.data_ptr())


class BCChain:
    """A fixed chain: plan once (buffers, spectra and step structure fixed), ``run()`` many times (CUDA-graph capturable).

    steps: list of steps; a step is a list of (x, spec, y, epilogue[, bias[, y_spectrum[, x_is_spectrum]]]) with x (1, q*B), STREAM-layout
    spec (p, q, B) and y (1, p*B) -- or (1, p*B/2) with ``_cabi.EPILOGUE_SWIGLU``.  A later step may read what an earlier one writes;
    with y_spectrum (float32, same width as y) the step also emits rfft of its rounded output, which a later step can take as x
    (x_is_spectrum=True) instead of transforming y again."""

    def __init__(self, steps: Sequence[Sequence[tuple]]):
        flat = [pr for st in steps for pr in st]
        if not 1 <= len(steps) <= _cabi.CHAIN_MAX_STEPS or not 1 <= len(flat) <= _cabi.CHAIN_MAX_PROBLEMS:
            raise ValueError(f"a chain holds 1..{_cabi.CHAIN_MAX_STEPS} steps and 1..{_cabi.CHAIN_MAX_PROBLEMS} problems")
        x0, spec0 = flat[0][0], flat[0][1]
        _cabi.require_cuda(x0, "input x")
        self.device = x0.device
        self.B = int(spec0.shape[-1])
        self.spec_dtype = spec0.dtype
        self._keep = flat                                           # the tensors behind the raw pointers
        self.problems = (_cabi.Problem * len(flat))()
        for k, pr in enumerate(flat):
            x, spec, y, epi = pr[:4]
            bias = pr[4] if len(pr) > 4 else None
            ysp = pr[5] if len(pr) > 5 else None
            xsp = bool(pr[6]) if len(pr) > 6 else False
            if spec.dim() != 3 or spec.shape[-1] != self.B or spec.dtype != self.spec_dtype or not spec.is_contiguous():
                raise ValueError("all problems of a chain share the block size and the spectrum dtype (STREAM layout, contiguous)")
            if x.shape != (1, spec.shape[1] * self.B):
                raise ValueError("chain problems are single tokens: x must be (1, in_features)")
            self.problems[k] = _problem(x, spec, y, epi, bias, ysp, xsp)
        first = [0]
        for st in steps:
            first.append(first[-1] + len(st))
        self.step_first = (ctypes.c_int * len(first))(*first)
        self.nsteps = len(steps)
        self.ys = [pr[2] for pr in flat]

    def run(self):
        ctr = barrier_counter(self.device) if self.nsteps > 1 else None
        with _cabi.on_device(self.device):
            rc = _cabi.lib().bcl_forward_chain(self.problems, self.step_first, self.nsteps, self.B, _cabi.DTYPE[self.spec_dtype],
                                               None if ctr is None else ctr.data_ptr(), _cabi.stream_ptr(self.device))
        _cabi.check(rc, "bcl_forward_chain")
        return self.ys


# ---------------------------------------------------------------------------------------------------------
# whole-MLP fusion
# ---------------------------------------------------------------------------------------------------------
def _is_silu(act) -> bool:
    if isinstance(act, nn.SiLU) or act is torch.nn.functional.silu:
        return True
    return type(act).__name__ in ("SiLUActivation", "SiLU")


class FusedBCMLP:
    """down(silu(gate(x)) * up(x)) for one token as ONE launch.  Holds the stacked (gate | up) spectrum through a SharedInputGroup
    (so parameter updates invalidate it exactly like the members' own spectra) and a persistent intermediate h."""

    def __init__(self, gate: BlockCirculantLinear, up: BlockCirculantLinear, down: BlockCirculantLinear):
        if not (gate.in_features == up.in_features and gate.out_features == up.out_features == down.in_features):
            raise ValueError("gate/up/down shapes do not form an MLP")
        if len({gate.block_size, up.block_size, down.block_size}) != 1 or len({gate.spectrum_dtype, up.spectrum_dtype, down.spectrum_dtype}) != 1:
            raise ValueError("the three projections must share the block size and the spectrum dtype")
        if gate.bias is not None or up.bias is not None or down.bias is not None:
            raise ValueError("the fused MLP covers bias-free projections (Llama); others take the per-module path")
        self.gate, self.up, self.down = gate, up, down
        self.group = gate._shared_group if (gate._shared_group is not None and gate._shared_group is up._shared_group
                                            and gate._shared_group.members == [gate, up]) else SharedInputGroup([gate, up])
        self._h: Optional[torch.Tensor] = None
        self.fused_calls = 0
        self.fused_prefill_calls = 0

    def eligible(self, x: torch.Tensor) -> bool:
        if torch.is_grad_enabled() or self.gate.training or self.up.training or self.down.training:
            return False
        if not (x.is_cuda and x.dtype in _cabi.DTYPE and x.shape[-1] == self.gate.in_features and x.numel() == self.gate.in_features):
            return False
        return self.gate._cache_spectrum and self.up._cache_spectrum and self.down._cache_spectrum and self.gate.c.is_cuda

    def _common_ok(self, x: torch.Tensor) -> bool:
        if torch.is_grad_enabled() or self.gate.training or self.up.training or self.down.training:
            return False
        if not (x.is_cuda and x.dtype in _cabi.DTYPE and x.shape[-1] == self.gate.in_features and self.gate.c.is_cuda):
            return False
        return self.gate._cache_spectrum and self.up._cache_spectrum and self.down._cache_spectrum

    def eligible_prefill(self, x: torch.Tensor) -> bool:
        """Many tokens: the stacked gate | up layer would take the tensor-core path and bcl_forward_swiglu covers the configuration."""
        if not self._common_ok(x) or x.dtype == torch.float32 or self.gate.block_size > 128 or self.gate.spectrum_dtype not in (torch.bfloat16, torch.float32):
            return False
        n = x.numel() // self.gate.in_features
        return n > 1 and self._stacked_work(n)

    def _stacked_work(self, n: int) -> bool:
        from .patch_llama_fft import prefill_token_floor
        return n >= prefill_token_floor(self.group.out_blocks, self.group.in_blocks, self.gate.spectrum_dtype)

    def forward_prefill(self, x: torch.Tensor) -> torch.Tensor:
        g = self.group
        x_flat = x.reshape(-1, self.gate.in_features)
        if x_flat.stride(-1) != 1 or (x_flat.stride(0) * x_flat.element_size()) % 16 or x_flat.data_ptr() % 16:
            x_flat = x_flat.contiguous()
        n = x_flat.shape[0]
        spec = g.build_spectrum(_cabi.LAYOUT_GEMM)
        h = torch.empty(n, self.down.in_features, device=x_flat.device, dtype=x_flat.dtype)
        L = _cabi.lib()
        sd = _cabi.DTYPE[self.gate.spectrum_dtype]                       # bf16, or f32 = the (hi, lo) split stages
        ws_bytes = L.bcl_workspace_bytes(n, g.out_blocks, g.in_blocks, g.block_size, sd, _cabi.LAYOUT_GEMM)
        ws = prefill_workspace(x_flat.device, ws_bytes)
        with _cabi.on_device(x_flat.device):
            rc = L.bcl_forward_swiglu(x_flat.data_ptr(), _cabi.DTYPE[x_flat.dtype], n, x_flat.stride(0), spec.data_ptr(), sd,
                                      g.out_blocks, g.in_blocks, g.block_size, h.data_ptr(), _cabi.DTYPE[h.dtype], h.stride(0),
                                      ws.data_ptr(), ws.numel(), _cabi.stream_ptr(x_flat.device))
        _cabi.check(rc, "bcl_forward_swiglu")
        self.fused_prefill_calls += 1
        y = self.down(h)
        return y.view(*x.shape[:-1], self.down.out_features)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x_flat = x.reshape(1, -1)
        if x_flat.stride(-1) != 1 or x_flat.data_ptr() % 16:
            x_flat = x_flat.contiguous()
        spec_gu = self.group.build_spectrum(_cabi.LAYOUT_STREAM)
        spec_d = self.down.build_spectrum(layout=_cabi.LAYOUT_STREAM)
        if self._h is None or self._h.dtype != x_flat.dtype or self._h.device != x_flat.device:
            self._h = torch.empty(1, self.down.in_features, device=x_flat.device, dtype=x_flat.dtype)
            self._hf = torch.empty(1, self.down.in_features, device=x_flat.device, dtype=torch.float32)
        y = torch.empty(1, self.down.out_features, device=x_flat.device, dtype=x_flat.dtype)
        # step 0 also emits rfft(h) (of the rounded h, bit for bit what down_proj would compute): step 1 starts from the spectrum
        probs = (_cabi.Problem * 2)(_problem(x_flat, spec_gu, self._h, _cabi.EPILOGUE_SWIGLU, None, self._hf),
                                    _problem(self._hf, spec_d, y, _cabi.EPILOGUE_LINEAR, None, None, True))
        first = (ctypes.c_int * 3)(0, 1, 2)
        ctr = barrier_counter(x_flat.device)
        with _cabi.on_device(x_flat.device):
            rc = _cabi.lib().bcl_forward_chain(probs, first, 2, self.gate.block_size, _cabi.DTYPE[self.gate.spectrum_dtype], ctr.data_ptr(),
                                               _cabi.stream_ptr(x_flat.device))
        _cabi.check(rc, "bcl_forward_chain")
        self.fused_calls += 1
        return y.view(*x.shape[:-1], self.down.out_features)


def fuse_mlp(model: nn.Module) -> int:
    """Give every module that owns bias-free BC ``gate_proj`` / ``up_proj`` / ``down_proj`` and a SiLU ``act_fn`` (HF LlamaMLP) a
    forward that serves single-token eval / no-grad calls with one chain launch.  Returns the number of fused MLPs."""
    n = 0
    for mod in model.modules():
        g, u, d = (getattr(mod, nm, None) for nm in ("gate_proj", "up_proj", "down_proj"))
        if not all(isinstance(m, BlockCirculantLinear) for m in (g, u, d)) or hasattr(mod, "_bcl_fused_mlp"):
            continue
        if not _is_silu(getattr(mod, "act_fn", None)):
            continue
        try:
            fused = FusedBCMLP(g, u, d)
        except ValueError:
            continue
        orig = mod.forward

        def forward(self, x, _fused=fused, _orig=orig):
            if _fused.eligible(x):
                return _fused.forward(x)
            if _fused.eligible_prefill(x):
                return _fused.forward_prefill(x)
            return _orig(x)
        mod._bcl_fused_mlp = fused
        mod._bcl_orig_forward = orig
        mod.forward = types.MethodType(forward, mod)
        n += 1
    return n


def unfuse_mlp(model: nn.Module) -> int:
    n = 0
    for mod in model.modules():
        if hasattr(mod, "_bcl_fused_mlp"):
            orig = mod._bcl_orig_forward
            mod.__dict__.pop("forward", None)                       # back to the class's forward ...
            if getattr(mod.forward, "__func__", None) is not getattr(orig, "__func__", orig):
                mod.forward = orig                                  # ... unless the original was itself an instance-level override
            del mod._bcl_fused_mlp, mod._bcl_orig_forward
            n += 1
    return n
