"""Grouped forward: several independent BC linears (q/k/v of a layer, gate/up, a sweep over layers) in ONE kernel
launch -- ``bcl_forward_grouped`` of include/bclinear.h.  The reference calls one module at a time
(src/patch_llama_fft.py:226); at batch-1 decode one projection is 0.5-1.4 MB of spectrum, far below launch
latency, so independent problems are submitted together and the launch streams all their spectra at HBM speed."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _cabi


class GroupedBCForward:
    """Plan once, replay many times (CUDA-graph capturable).

    xs[g]: (N_g, in_g) activations (N_g <= 8); specs[g]: STREAM-layout spectrum (p_g, q_g, B); biases[g]: optional.
    Output buffers are allocated here and written in place by every ``run()``."""

    def __init__(self, xs: Sequence[torch.Tensor], specs: Sequence[torch.Tensor],
                 biases: Optional[Sequence[Optional[torch.Tensor]]] = None, *, num_ctas: Optional[int] = None,
                 _y_packed=None):
        assert len(xs) == len(specs) and len(xs) > 0
        biases = list(biases) if biases is not None else [None] * len(xs)
        dev = xs[0].device
        _cabi.require_cuda(xs[0], "input x")
        B = specs[0].shape[-1]
        spec_dtype = specs[0].dtype
        self.device, self.B = dev, B
        self.xs, self.specs, self.biases = list(xs), list(specs), biases
        self.ys = []
        self.y_all = None
        if _y_packed is not None:
            self.y_all = torch.empty(_y_packed[0], device=dev, dtype=xs[0].dtype)
        probs = (_cabi.Problem * len(xs))()
        for g, (x, s, b) in enumerate(zip(xs, specs, biases)):
            assert x.device == dev and s.device == dev and x.dim() == 2 and s.dim() == 3
            assert s.shape[-1] == B and s.dtype == spec_dtype and x.is_contiguous() and s.is_contiguous()
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
        nbytes = L.bcl_group_plan_bytes(len(xs), self.num_ctas)
        self.host_plan = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        rc = L.bcl_group_plan_encode(probs, len(xs), B, _cabi.DTYPE[spec_dtype], self.num_ctas,
                                     self.host_plan.data_ptr(), nbytes)
        _cabi.check(rc, "bcl_group_plan_encode")
        self.dev_plan = self.host_plan.to(dev)
        self.algorithmic_bytes = sum(x.numel() * x.element_size() + s.numel() * s.element_size() + y.numel() * y.element_size()
                                     + (0 if b is None else b.numel() * b.element_size())
                                     for x, s, y, b in zip(self.xs, self.specs, self.ys, self.biases))

    @classmethod
    def from_packed(cls, shapes, specs, N, x_dtype, device, **kw):
        """Problems whose x (and y) live back to back in ONE staging buffer each, so a host round trip is one copy
        per direction (``run_host``).  shapes[g] = (p, q); returns the instance with .x_all / .y_all set."""
        B = specs[0].shape[-1]
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
