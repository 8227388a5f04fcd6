"""Host logic of output-block-row sharding (llama-fft_b200/sharded.py) on CPU: world_size-2 gloo processes gather their
column slices; the local slices come from the oracle (the CUDA forward needs a GPU, and there is no CPU fallback)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import llama_fft_b200 as bc
from oracle import bc_oracle as O


def test_shard_rows_partitions_every_row_once():
    for p in (1, 4, 32, 86, 112):
        for world in (1, 2, 4, 8):
            r = bc.shard_rows(p, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == p
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [j1 - j0 for j0, j1 in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, p, q, B, N, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)                      # same c, x, bias on every rank
        c = torch.randn(p, q, B, generator=g) * 0.1
        x = torch.randn(N, q * B, generator=g)
        bias = torch.randn(p * B, generator=g)
        full = bc.BlockCirculantLinear(q * B, p * B, block_size=B, bias=True)
        with torch.no_grad():
            full.c.copy_(c)
            full.bias.copy_(bias)
        sh = bc.shard_bc_linear(full)                             # slices c / bias for this rank
        j0, j1 = sh.ranges[rank]
        ok = sh.world == world and sh.rank == rank
        if j1 > j0:
            ok = ok and torch.equal(sh.local.c, c[j0:j1]) and torch.equal(sh.local.bias, bias[j0 * B:j1 * B])
            y_loc = O.bc_forward_torch(c[j0:j1], x, bias[j0 * B:j1 * B])
        else:
            ok = ok and sh.local is None
            y_loc = x.new_empty(N, 0)
        y = bc.gather_columns(y_loc, sh.ranges, B)
        ok = ok and torch.equal(y, O.bc_forward_torch(c, x, bias))
        raised = False
        try:
            sh(x)                                                 # CPU tensors: the product path must refuse, not fall back
        except RuntimeError as e:
            raised = "no CPU" in str(e)
        except Exception:
            raised = False
        if j1 == j0:                                              # an idle rank only takes part in the collective
            raised = True
        out[rank] = bool(ok and raised)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p,q,B,N", [(6, 3, 64, 5), (3, 2, 128, 4), (1, 2, 64, 3)])
def test_world2_gloo_gather_matches_unsharded(p, q, B, N):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, p, q, B, N, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
    assert all(pr.exitcode == 0 for pr in procs)
    assert dict(out) == {0: True, 1: True}
