"""Output-block-row sharding of a BC linear across the GPUs of one box (BASELINE.json configs[4]).

The reference has no distributed code.  The BC forward shards naturally along output block-rows: rank g holds
c[j0_g:j1_g] (and its spectrum), sees the full x (N, in), computes y[:, j0_g*B : j1_g*B] with the same sm_100a kernels as
the unsharded layer, and the slices meet in ONE exchange step:

    exchange="nccl"   torch.distributed all_gather_into_tensor over NCCL / NVLink.  One token with an even split gathers straight into
                      y (slices of a single row ARE the final layout); several tokens need one interleave copy (G, N, out/G) -> (N, out)
    exchange="push"   the rank's kernels write its slice IN PLACE into its own (N, out) symmetric buffer (y_row_stride = out), then one
                      bcl_push_columns launch stores the slice into the same columns of every peer's buffer: the gather lands in the
                      final layout, bytes cross NVLink once, no collective and no interleave pass
    exchange="p2p"    fully fused: the producing kernels' epilogues store every finished tile into EVERY peer's y as well (stream kernel:
                      peer stores; tensor-core inverse DFT: one TMA store per destination); nothing runs afterwards

Block-rows are split as evenly as possible; when p is not divisible by the world size the NCCL path pads the slices to the
largest one (70B k/v: p = 4 on 8 ranks leaves ranks idle; ``shard_rows`` reports empty ranges for them).
One process per GPU (torchrun); data-parallel replicas need none of this.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from .patch_llama_fft import BlockCirculantLinear


def shard_rows(p: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced split of the p output block-rows: the first p % world ranks get one extra row."""
    base, extra = divmod(p, world)
    out, j = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((j, j + n))
        j += n
    return out


def gather_columns(y_local: torch.Tensor, ranges: List[Tuple[int, int]], B: int, group=None) -> torch.Tensor:
    """All-gather column slices: y_local is (N, (j1-j0)*B) on every rank -> (N, p*B) everywhere.
    Even split, one token: the collective writes y directly (row 0's slices back to back are the final layout).  Otherwise the
    slices (padded to the widest when the split is ragged) are gathered as (G, N, w) and ONE permuting copy lays them out."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    N = y_local.shape[0]
    widths = [(j1 - j0) * B for j0, j1 in ranges]
    wmax = max(widths)
    p = ranges[-1][1]
    assert widths[rank] == y_local.shape[1]
    even = all(w == wmax for w in widths)
    if even and N == 1:
        y = y_local.new_empty(1, p * B)
        dist.all_gather_into_tensor(y.view(world, wmax), y_local.contiguous(), group=group)
        return y
    send = y_local
    if y_local.shape[1] != wmax:
        send = y_local.new_zeros(N, wmax)
        send[:, :y_local.shape[1]] = y_local
    recv = y_local.new_empty(world, N, wmax)
    dist.all_gather_into_tensor(recv.view(world * N, wmax), send.contiguous(), group=group)
    if even:
        return recv.permute(1, 0, 2).reshape(N, p * B)               # one copy kernel: (G, N, w) -> (N, G * w)
    y = y_local.new_empty(N, p * B)
    for r, (j0, j1) in enumerate(ranges):
        if j1 > j0:
            y[:, j0 * B:j1 * B] = recv[r, :, :(j1 - j0) * B]
    return y


class ShardedBlockCirculantLinear(nn.Module):
    """Rank-local shard of a BlockCirculantLinear (same parameters c / bias, restricted to this rank's block-rows)."""

    def __init__(self, full: BlockCirculantLinear, group=None, *, exchange: str = "nccl"):
        super().__init__()
        if exchange not in ("nccl", "push", "p2p"):
            raise ValueError("exchange must be 'nccl', 'push' or 'p2p'")
        self.group = group
        self.exchange = exchange
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.in_features, self.out_features, self.block_size = full.in_features, full.out_features, full.block_size
        self.ranges = shard_rows(full.out_blocks, self.world)
        j0, j1 = self.ranges[self.rank]
        self.j0, self.j1 = j0, j1
        B = full.block_size
        self.local: Optional[BlockCirculantLinear] = None
        if j1 > j0:
            loc = BlockCirculantLinear(full.in_features, (j1 - j0) * B, block_size=B, bias=full.bias is not None,
                                       spectrum_dtype=full.spectrum_dtype).to(full.c.device)
            with torch.no_grad():
                loc.c.copy_(full.c[j0:j1])
                if full.bias is not None:
                    loc.bias.data = full.bias.detach()[j0 * B:j1 * B].clone()
            self.local = loc.eval() if not full.training else loc
        self._symm = None       # (N, dtype) -> (symmetric y buffer, peer views) for exchange="p2p"
        self._bias_dtype = None if full.bias is None else full.bias.dtype     # known on EVERY rank, with or without local rows

    def y_dtype(self, x_dtype: torch.dtype) -> torch.dtype:
        """Output dtype of the full layer (reference :266-271: x.dtype, promoted with the bias dtype), the same on every rank."""
        return x_dtype if self._bias_dtype is None else torch.promote_types(x_dtype, self._bias_dtype)

    # -- p2p: y lives in symmetric memory; each rank's kernels write their slice into every peer's copy ------------
    def _symmetric_y(self, N: int, dtype: torch.dtype, device):
        key = (N, dtype)
        if self._symm is not None and self._symm[0] == key:
            return self._symm[1], self._symm[2]
        import torch.distributed._symmetric_memory as symm_mem
        y = symm_mem.empty(N, self.out_features, dtype=dtype, device=device)
        hdl = symm_mem.rendezvous(y, self.group if self.group is not None else dist.group.WORLD)
        peers = [hdl.get_buffer(r, (N, self.out_features), dtype) for r in range(self.world)]
        self._symm = (key, y, peers, hdl)
        return y, peers

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lead = None
        if x.dim() == 3:
            lead = x.shape[:2]
            x = x.reshape(-1, x.shape[-1])
        elif x.dim() != 2:
            raise ValueError("Unsupported input dim")
        B = self.block_size
        N = x.shape[0]
        if self.exchange == "nccl":
            if self.local is not None:
                y_loc = self.local(x)
            else:                                           # a rank without rows still joins the collective, in the dtype the others produce
                y_loc = torch.empty(N, 0, device=x.device, dtype=self.y_dtype(x.dtype))
            y = gather_columns(y_loc, self.ranges, B, self.group)
        elif self.exchange == "push":
            y_dtype = self.y_dtype(x.dtype)
            y, peers = self._symmetric_y(N, y_dtype, x.device)
            hdl = self._symm[3]
            hdl.barrier()                                   # peers finished reading the previous contents
            if self.local is not None:
                bc_forward_bcast(self.local, x, [y[:, self.j0 * B:self.j1 * B]])        # in place, final layout (row stride = out)
                push_columns(y, [pv for r, pv in enumerate(peers) if r != self.rank], self.j0 * B, (self.j1 - self.j0) * B)
            hdl.barrier()                                   # every slice of every peer has landed
        else:
            y_dtype = self.y_dtype(x.dtype)
            y, peers = self._symmetric_y(N, y_dtype, x.device)
            hdl = self._symm[3]
            hdl.barrier()                                   # peers finished reading the previous contents
            if self.local is not None:                      # ONE forward whose epilogue stores into every rank's y
                order = [peers[self.rank]] + [pv for r, pv in enumerate(peers) if r != self.rank]   # own copy first (plain address)
                bc_forward_bcast(self.local, x, [pv[:, self.j0 * B:self.j1 * B] for pv in order])
            hdl.barrier()                                   # every slice of every peer has landed
        if lead is not None:
            return y.view(lead[0], lead[1], self.out_features)
        return y


def bc_forward_bcast(layer: BlockCirculantLinear, x: torch.Tensor, dsts: List[torch.Tensor]) -> None:
    """``bcl_forward_bcast``: one forward of `layer`, result stored to every tensor of `dsts` (same shape/strides/dtype;
    views into peer-mapped symmetric buffers for the sharded layer)."""
    import ctypes
    from . import _cabi
    from .patch_llama_fft import prefill_workspace
    if x.stride(-1) != 1 or (x.stride(0) * x.element_size()) % 16 or x.data_ptr() % 16:
        x = x.contiguous()
    N = x.shape[0]
    p, q, B = layer.out_blocks, layer.in_blocks, layer.block_size
    layout = _cabi.LAYOUT_GEMM if layer._use_prefill_path(N) else _cabi.LAYOUT_STREAM
    spec = layer.build_spectrum(layout=layout)
    spec_dt = _cabi.DTYPE[layer.spectrum_dtype] if layout == _cabi.LAYOUT_GEMM else _cabi.DTYPE[spec.dtype]
    d0 = dsts[0]
    assert all(d.shape == (N, p * B) and d.dtype == d0.dtype and d.stride() == d0.stride() for d in dsts)
    L = _cabi.lib()
    ws_ptr, ws_bytes = None, 0
    if layout == _cabi.LAYOUT_GEMM and N > 0:
        ws = prefill_workspace(x.device, L.bcl_workspace_bytes(N, p, q, B, spec_dt, layout))
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
    ptrs = (ctypes.c_void_p * len(dsts))(*[d.data_ptr() for d in dsts])
    bias = layer.bias
    with _cabi.on_device(x.device):
        rc = L.bcl_forward_bcast(x.data_ptr(), _cabi.DTYPE[x.dtype], N, x.stride(0), spec.data_ptr(), spec_dt, layout, p, q, B,
                                 None if bias is None else bias.data_ptr(), 0 if bias is None else _cabi.DTYPE[bias.dtype],
                                 ptrs, len(dsts), _cabi.DTYPE[d0.dtype], d0.stride(0), ws_ptr, ws_bytes, _cabi.stream_ptr(x.device))
    _cabi.check(rc, "bcl_forward_bcast")


def push_columns(y: torch.Tensor, dsts: List[torch.Tensor], col_offset: int, width: int) -> None:
    """``bcl_push_columns``: copy y[:, col_offset : col_offset + width] into the same columns of every tensor of `dsts` (peer-mapped
    buffers of the same shape / strides / dtype)."""
    import ctypes
    from . import _cabi
    if not dsts or width == 0 or y.shape[0] == 0:
        return
    assert all(d.shape == y.shape and d.stride() == y.stride() and d.dtype == y.dtype for d in dsts) and y.stride(1) == 1
    ptrs = (ctypes.c_void_p * len(dsts))(*[d.data_ptr() for d in dsts])
    with _cabi.on_device(y.device):
        rc = _cabi.lib().bcl_push_columns(y.data_ptr(), ptrs, len(dsts), y.shape[0], y.stride(0), col_offset, width, _cabi.DTYPE[y.dtype],
                                          _cabi.stream_ptr(y.device))
    _cabi.check(rc, "bcl_push_columns")


def shard_bc_linear(layer: BlockCirculantLinear, group=None, *, exchange: str = "nccl") -> ShardedBlockCirculantLinear:
    return ShardedBlockCirculantLinear(layer, group, exchange=exchange)
