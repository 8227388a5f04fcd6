"""Shared-input fusion: BC layers that read the SAME activation -- q/k/v of an attention block, gate/up of an
MLP -- run as ONE block-circulant forward over their concatenated output block-rows.

Why: rfft(x) (reference src/patch_llama_fft.py:249-252) is identical for every member, so the fused call reads x and writes /
re-reads its spectrum once instead of once per member (prefill: one forward-DFT pass instead of two or three; decode: one launch
instead of two or three).  Block-row j of a BC linear depends on c[j] only, so stacking members along the output-block axis
    c_fused = cat([c_q, c_k, c_v], dim=0),   y_fused[:, rows of member m] == y_m
is exact: the fused outputs are bit-identical to the members' own outputs.

Drop-in mechanics: the members stay ordinary ``BlockCirculantLinear`` modules in the model (same state-dict keys, same
``isinstance`` checks, checkpoint I/O and calibration untouched).  ``fuse_shared_inputs(model)`` only attaches a group object;
in eval / no-grad mode the first member called with an activation computes the fused result and every member returns its
column slice (a view, like the reference's reshape results).  Anything else -- grad mode, a member called with a different
tensor -- takes the member's own path.
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _cabi
from .patch_llama_fft import BlockCirculantLinear, bc_forward

SHARED_INPUT_SETS: Tuple[Tuple[str, ...], ...] = (("q_proj", "k_proj", "v_proj"), ("gate_proj", "up_proj"))


def _signature(m: BlockCirculantLinear):
    """What members of one group must agree on (the epilogue's rounding / bias promotion is per call, reference :266-271)."""
    return (m.in_features, m.block_size, m.spectrum_dtype, None if m.bias is None else m.bias.dtype)


class SharedInputGroup:
    """Members read the same (N, in) activation; one fused forward serves all of them."""

    def __init__(self, members: Sequence[BlockCirculantLinear]):
        members = list(members)
        if len(members) < 2:
            raise ValueError("a shared-input group needs at least two layers")
        m0 = members[0]
        for m in members:
            if not isinstance(m, BlockCirculantLinear):
                raise TypeError("members must be BlockCirculantLinear modules")
            if _signature(m) != _signature(m0):
                raise ValueError("members must share in_features, block_size, spectrum_dtype and bias presence / dtype")
        self.members: List[BlockCirculantLinear] = members
        self.in_blocks = m0.in_blocks
        self.block_size = m0.block_size
        self.out_blocks = sum(m.out_blocks for m in members)
        self.offsets = [0]
        for m in members:
            self.offsets.append(self.offsets[-1] + m.out_features)
        self._spec = {}            # layout -> fused spectrum
        self._bias = None
        self._key = None
        self._pending = None       # (x kept alive, key, fused y, members still to read it)
        self.fused_calls = 0       # forward_all launches (tests / the bench assert that the fused path really ran)

    # -- fused parameters -----------------------------------------------------------------------------------
    def _members_key(self):
        return tuple(m._spectrum_key() + ((None,) if m.bias is None else (m.bias.data_ptr(), m.bias._version, m.bias.dtype))
                     for m in self.members)

    def _refresh(self):
        key = self._members_key()
        if key != self._key:
            self._spec, self._bias, self._key = {}, None, key
            if self.members[0].bias is not None:
                self._bias = torch.cat([m.bias.detach() for m in self.members])

    def build_spectrum(self, layout: int = _cabi.LAYOUT_STREAM) -> torch.Tensor:
        """rfft of the stacked first columns, in the layout the kernel wants (same entry point as a single layer)."""
        self._refresh()
        spec = self._spec.get(layout)
        if spec is not None:
            return spec
        m0 = self.members[0]
        c = torch.cat([m.c.detach().to(torch.float32) for m in self.members], dim=0).contiguous()
        _cabi.require_cuda(c, "BlockCirculantLinear.c")
        L = _cabi.lib()
        if layout == _cabi.LAYOUT_STREAM:
            spec = torch.empty(self.out_blocks, self.in_blocks, self.block_size, device=c.device, dtype=m0.spectrum_dtype)
        else:
            nbytes = L.bcl_spectrum_bytes(self.out_blocks, self.in_blocks, self.block_size, _cabi.DTYPE[m0.spectrum_dtype], layout)
            spec = torch.empty(nbytes, device=c.device, dtype=torch.uint8)
        with torch.cuda.device(c.device):
            rc = L.bcl_precompute_spectrum(c.data_ptr(), self.out_blocks, self.in_blocks, self.block_size, spec.data_ptr(),
                                           _cabi.DTYPE[m0.spectrum_dtype], layout, _cabi.stream_ptr(c.device))
        _cabi.check(rc, "bcl_precompute_spectrum")
        self._spec[layout] = spec
        return spec

    # -- forward --------------------------------------------------------------------------------------------
    def forward_all(self, x_flat: torch.Tensor) -> torch.Tensor:
        """(N, in) -> (N, sum of the members' out_features); member m owns columns offsets[m]:offsets[m+1]."""
        m0 = self.members[0]
        self.fused_calls += 1
        layout = _cabi.LAYOUT_GEMM if m0._use_prefill_path(x_flat.shape[0]) else _cabi.LAYOUT_STREAM
        spec = self.build_spectrum(layout)
        return bc_forward(x_flat, spec, self._bias, self.out_blocks, self.in_blocks, self.block_size, layout=layout,
                          spectrum_dtype=m0.spectrum_dtype)

    @staticmethod
    def _x_key(x_flat: torch.Tensor):
        # inference tensors (torch.inference_mode, which every eval function of the reference's bench uses) carry no version counter
        version = None if x_flat.is_inference() else x_flat._version
        return (x_flat.data_ptr(), version, tuple(x_flat.shape), tuple(x_flat.stride()), x_flat.dtype, str(x_flat.device),
                torch.cuda.current_stream(x_flat.device).cuda_stream)

    def member_forward(self, member: BlockCirculantLinear, x_flat: torch.Tensor) -> torch.Tensor:
        """Called from BlockCirculantLinear.forward (eval / no-grad only).  The activation is kept referenced while a fused
        result is pending, so its address cannot be recycled for another tensor in between."""
        idx = next(i for i, m in enumerate(self.members) if m is member)
        key = self._x_key(x_flat)
        if self._pending is None or self._pending[1] != key or idx not in self._pending[3]:
            self._pending = (x_flat, key, self.forward_all(x_flat), set(range(len(self.members))))
        _, _, y, waiting = self._pending
        waiting.discard(idx)
        if not waiting:
            self._pending = None
        return y[:, self.offsets[idx]:self.offsets[idx + 1]]


def fuse_shared_inputs(model: nn.Module, name_sets: Sequence[Sequence[str]] = SHARED_INPUT_SETS, *, drop_member_spectra: bool = True) -> int:
    """Attach a SharedInputGroup to every module that owns all names of a set as BlockCirculantLinear children with equal
    in_features / block size / spectrum dtype (LlamaAttention: q/k/v, LlamaMLP: gate/up).  Returns the number of groups."""
    n = 0
    for parent in model.modules():
        for names in name_sets:
            members = [getattr(parent, nm, None) for nm in names]
            if not all(isinstance(m, BlockCirculantLinear) for m in members):
                continue
            if any(_signature(m) != _signature(members[0]) for m in members):
                continue
            group = SharedInputGroup(members)
            for m in members:
                m._shared_group = group
                if drop_member_spectra:
                    m.drop_spectrum()
            n += 1
    return n


def unfuse_shared_inputs(model: nn.Module) -> int:
    n = 0
    seen = set()
    for m in model.modules():
        g = getattr(m, "_shared_group", None)
        if isinstance(m, BlockCirculantLinear) and g is not None:
            m._shared_group = None
            if id(g) not in seen:
                seen.add(id(g))
                n += 1
    return n
