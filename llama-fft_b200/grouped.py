"""Grouped forward: several independent BC linears (q/k/v of a layer, gate/up, a sweep over layers) in ONE kernel
launch -- ``bcl_forward_grouped`` of include/bclinear.h.  The reference calls one module at a time
(src/patch_llama_fft.py:226); at batch-1 decode one projection is 0.5-1.4 MB of spectrum, far below launch
latency, so independent problems are submitted together and the launch streams all their spectra at HBM speed."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _cabi


class GroupedBCForward:
    """Plan once, replay many times (CUDA-graph capturable).

    xs[g]: (N_g, in_g) activations (N_g <= 8); specs[g]: STREAM-layout spectrum (p_g, q_g, B); biases[g]: optional.
    Output buffers are allocated here and written in place by every ``run()``."""

    def __init__(self, xs: Sequence[torch.Tensor], specs: Sequence[torch.Tensor],
                 biases: Optional[Sequence[Optional[torch.Tensor]]] = None, *, num_ctas: Optional[int] = None,
                 _y_packed=None, layout: int = _cabi.LAYOUT_STREAM, shapes: Optional[Sequence[Tuple[int, int]]] = None,
                 block_size: Optional[int] = None):
        """layout = LAYOUT_DECODE_TC: `specs` are the opaque bf16 operand images of ``decode_tc_spectrum`` (uint8 buffers), `shapes`
        = [(p, q)] and `block_size` say what they hold; the plan then runs the tensor-core decode kernel (one token per problem)."""
        assert len(xs) == len(specs) and len(xs) > 0
        biases = list(biases) if biases is not None else [None] * len(xs)
        dev = xs[0].device
        _cabi.require_cuda(xs[0], "input x")
        tc = layout == _cabi.LAYOUT_DECODE_TC
        if tc and (shapes is None or block_size is None or len(shapes) != len(specs)):
            raise ValueError("LAYOUT_DECODE_TC needs shapes=[(p, q)] and block_size")
        B = int(block_size) if tc else specs[0].shape[-1]
        spec_dtype = torch.bfloat16 if tc else specs[0].dtype
        self.device, self.B = dev, B
        self.xs, self.specs, self.biases = list(xs), list(specs), biases
        self.ys = []
        self.y_all = None
        if _y_packed is not None:
            self.y_all = torch.empty(_y_packed[0], device=dev, dtype=xs[0].dtype)
        probs = (_cabi.Problem * len(xs))()
        for g, (x, s, b) in enumerate(zip(xs, specs, biases)):
            assert x.device == dev and s.device == dev and x.dim() == 2 and x.is_contiguous() and s.is_contiguous()
            if tc:
                p, q = int(shapes[g][0]), int(shapes[g][1])
            else:
                assert s.dim() == 3 and s.shape[-1] == B and s.dtype == spec_dtype
                p, q = int(s.shape[0]), int(s.shape[1])
            if x.shape[1] != q * B:
                raise ValueError("in_features mismatch")
            y_dtype = x.dtype if b is None else torch.promote_types(x.dtype, b.dtype)
            if self.y_all is not None:
                o = _y_packed[1][g]
                y = self.y_all[o:o + x.shape[0] * p * B].view(x.shape[0], p * B)
            else:
                y = torch.empty(x.shape[0], p * B, device=dev, dtype=y_dtype)
            self.ys.append(y)
            probs[g] = _cabi.Problem(x.data_ptr(), s.data_ptr(), None if b is None else b.data_ptr(), y.data_ptr(),
                                     x.shape[0], x.stride(0), y.stride(0), p, q,
                                     _cabi.DTYPE[x.dtype], _cabi.DTYPE[y_dtype], 0 if b is None else _cabi.DTYPE[b.dtype], 0)
        L = _cabi.lib()
        if num_ctas is None:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            num_ctas = sms           # the warp-specialised kernel runs one persistent CTA per SM
        self.num_ctas = int(num_ctas)
        nbytes = L.bcl_group_plan_bytes_for(probs, len(xs), B, _cabi.DTYPE[spec_dtype], self.num_ctas)   # with room for the dynamic tail
        self.host_plan = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        rc = L.bcl_group_plan_encode_layout(probs, len(xs), B, _cabi.DTYPE[spec_dtype], layout, self.num_ctas,
                                            self.host_plan.data_ptr(), nbytes)
        _cabi.check(rc, "bcl_group_plan_encode_layout")
        self.dev_plan = self.host_plan.to(dev)
        # (DECODE_TC images pad p to 8 and 2q to 16: the algorithmic figure stays p * q * B reals)
        self.algorithmic_bytes = sum(x.numel() * x.element_size() + (y.shape[1] * x.shape[1] // B * 2 if tc else s.numel() * s.element_size())
                                     + y.numel() * y.element_size()
                                     + (0 if b is None else b.numel() * b.element_size())
                                     for x, s, y, b in zip(self.xs, self.specs, self.ys, self.biases))

    @classmethod
    def from_packed(cls, shapes, specs, N, x_dtype, device, block_size=None, **kw):
        """Problems whose x (and y) live back to back in ONE staging buffer each, so a host round trip is one copy
        per direction (``run_host``).  shapes[g] = (p, q); returns the instance with .x_all / .y_all set."""
        B = int(block_size) if block_size is not None else specs[0].shape[-1]
        if kw.get("layout", _cabi.LAYOUT_STREAM) == _cabi.LAYOUT_DECODE_TC:
            kw.update(shapes=list(shapes), block_size=B)
        esz = torch.empty(0, dtype=x_dtype).element_size()
        align = 16 // esz
        x_off, y_off, xo, yo = [], [], 0, 0
        for (p, q) in shapes:
            x_off.append(xo); y_off.append(yo)
            xo += -(-N * q * B // align) * align
            yo += -(-N * p * B // align) * align
        x_all = torch.zeros(xo, device=device, dtype=x_dtype)
        xs = [x_all[o:o + N * q * B].view(N, q * B) for o, (p, q) in zip(x_off, shapes)]
        self = cls(xs, specs, None, _y_packed=(yo, y_off), **kw)
        self.x_all = x_all
        return self

    def run_host(self, host_x: torch.Tensor, host_y: torch.Tensor):
        """End-to-end step with HOST buffers: H2D(x_all) -> one grouped launch -> D2H(y_all) -> synchronise."""
        with _cabi.on_device(self.device):
            rc = _cabi.lib().bcl_forward_grouped_host(
                self.dev_plan.data_ptr(), self.host_plan.data_ptr(),
                host_x.data_ptr(), self.x_all.data_ptr(), self.x_all.numel() * self.x_all.element_size(),
                host_y.data_ptr(), self.y_all.data_ptr(), self.y_all.numel() * self.y_all.element_size(),
                _cabi.stream_ptr(self.device))
        _cabi.check(rc, "bcl_forward_grouped_host")
        return host_y

    def submit_host(self, host_x: torch.Tensor, host_y: torch.Tensor, pipe: "HostPipe") -> None:
        """Pipelined end-to-end step (``bcl_forward_grouped_host_pipelined``): enqueue H2D(x) / kernel / D2H(y) on the pipe's three
        streams with this group as the slot; returns without waiting.  Call ``wait_host()`` before reading host_y or
        resubmitting this group."""
        if not hasattr(self, "_slot_events"):
            self._slot_events = [torch.cuda.Event() for _ in range(3)]
            for ev in self._slot_events:
                ev.record(pipe.compute)                       # materialise the cudaEvent_t handles
            self._slot_used = False
        ei, ec, eo = self._slot_events
        with _cabi.on_device(self.device):
            rc = _cabi.lib().bcl_forward_grouped_host_pipelined(
                self.dev_plan.data_ptr(), self.host_plan.data_ptr(),
                host_x.data_ptr(), self.x_all.data_ptr(), self.x_all.numel() * self.x_all.element_size(),
                host_y.data_ptr(), self.y_all.data_ptr(), self.y_all.numel() * self.y_all.element_size(),
                pipe.h2d.cuda_stream, pipe.compute.cuda_stream, pipe.d2h.cuda_stream,
                ei.cuda_event, ec.cuda_event, eo.cuda_event, 1 if self._slot_used else 0)
        _cabi.check(rc, "bcl_forward_grouped_host_pipelined")
        self._slot_used = True

    def submit_host_stream(self, host_x: torch.Tensor, host_y: torch.Tensor, stream: "torch.cuda.Stream") -> None:
        """One end-to-end step on `stream`, asynchronously (``bcl_forward_grouped_host_async``): H2D(x), the grouped launch and D2H(y) in
        stream order; returns at once.  Give every staging slot its own stream and neighbouring steps overlap by themselves.
        ``stream.synchronize()`` before reading host_y or resubmitting this group."""
        with _cabi.on_device(self.device):
            rc = _cabi.lib().bcl_forward_grouped_host_async(
                self.dev_plan.data_ptr(), self.host_plan.data_ptr(),
                host_x.data_ptr(), self.x_all.data_ptr(), self.x_all.numel() * self.x_all.element_size(),
                host_y.data_ptr(), self.y_all.data_ptr(), self.y_all.numel() * self.y_all.element_size(), stream.cuda_stream)
        _cabi.check(rc, "bcl_forward_grouped_host_async")

    def wait_host(self) -> None:
        """Block until this group's last submitted step has delivered host_y."""
        if getattr(self, "_slot_used", False):
            self._slot_events[2].synchronize()

    def run(self):
        with _cabi.on_device(self.device):
            rc = _cabi.lib().bcl_forward_grouped(self.dev_plan.data_ptr(), self.host_plan.data_ptr(),
                                                 _cabi.stream_ptr(self.device))
        _cabi.check(rc, "bcl_forward_grouped")
        return self.ys


def decode_tc_spectrum(c: torch.Tensor) -> torch.Tensor:
    """c (p, q, B) on a CUDA device -> the bf16 operand image of the tensor-core decode kernel (BCL_LAYOUT_DECODE_TC), a uint8 buffer."""
    _cabi.require_cuda(c, "c")
    p, q, B = (int(v) for v in c.shape)
    L = _cabi.lib()
    nbytes = L.bcl_spectrum_bytes(p, q, B, _cabi.BF16, _cabi.LAYOUT_DECODE_TC)
    if nbytes == 0:
        raise ValueError("LAYOUT_DECODE_TC covers block sizes 64 and 128")
    c32 = c.detach().to(torch.float32).contiguous()
    spec = torch.empty(nbytes, device=c.device, dtype=torch.uint8)
    with _cabi.on_device(c.device):
        rc = L.bcl_precompute_spectrum(c32.data_ptr(), p, q, B, spec.data_ptr(), _cabi.BF16, _cabi.LAYOUT_DECODE_TC, _cabi.stream_ptr(c.device))
    _cabi.check(rc, "bcl_precompute_spectrum")
    return spec


class HostPipe:
    """The three streams of the pipelined host entry point (H2D copies, kernels, D2H copies)."""

    def __init__(self, device):
        self.device = device
        self.h2d = torch.cuda.Stream(device)
        self.compute = torch.cuda.Stream(device)
        self.d2h = torch.cuda.Stream(device)

    def synchronize(self) -> None:
        for s in (self.h2d, self.compute, self.d2h):
            s.synchronize()
