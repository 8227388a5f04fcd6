"""Multi-rank GPU parity of output-block-row sharding (llama-fft_b200/sharded.py; BASELINE.json configs[4]): one process per
GPU over NCCL, every exchange mode (NCCL all-gather, push, fused peer stores), against the reference's forward on the oracle
(src/patch_llama_fft.py:226-275 restated in oracle/bc_oracle.py) and against the unsharded layer on the same GPU.
Skipped when the box has fewer than 2 GPUs (the CPU twin, tests/test_sharded_cpu.py, covers the host logic with gloo)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


CASES = [
    # (in, out, B, N, activation dtype, spectrum dtype, bias, tol)
    (512, 768, 64, 1, torch.float16, torch.float32, False, 1e-4),          # decode, p = 12 over 2 ranks
    (1024, 640, 128, 5, torch.float32, torch.float32, True, 1e-4),         # p = 5: ragged split (3 + 2), bias, small N
    (1024, 2048, 128, 300, torch.float16, torch.bfloat16, False, 1e-2),    # tensor-core path, ragged token tile
    (1024, 1024, 128, 300, torch.float16, torch.float32, False, 1e-4),     # split (fp32-spectrum) tensor-core path
    (8192, 1024, 256, 1, torch.float16, torch.bfloat16, False, 1e-2),      # 70B k/v decode: p = 4
    (8192, 8192, 256, 384, torch.bfloat16, torch.bfloat16, False, 1e-2),   # 70B q/o prefill, B = 256 half-N DFT launches
]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ.setdefault("BCL_PREFILL_MIN_TOKENS", "256")
    os.environ.setdefault("BCL_PREFILL_MIN_TOKENS_F32", "256")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import llama_fft_b200 as bc
    from llama_fft_b200.patch_llama_fft import prefill_status
    from oracle import bc_oracle as O
    from tests import _golden as G
    res = {}
    try:
        for ci, (inf, outf, B, N, dt, sdt, has_bias, tol) in enumerate(CASES):
            g = torch.Generator().manual_seed(100 + ci)              # same c, x, bias on every rank
            c = torch.randn(outf // B, inf // B, B, generator=g) * 0.02
            x = torch.randn(N, inf, generator=g).to(dt)
            b = (torch.randn(outf, generator=g) * 0.1).to(dt) if has_bias else None
            full = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=has_bias, spectrum_dtype=sdt)
            with torch.no_grad():
                full.c.copy_(c)
                if has_bias:
                    full.bias.data = b.clone()
            full = full.to(dev).eval()
            y_ref = O.bc_forward_torch(c, x, b).double()
            scale = y_ref.abs().max().item()
            with torch.no_grad():
                y_full = full(x.to(dev))
            for mode in ("nccl", "push", "p2p"):
                sh = bc.shard_bc_linear(full, exchange=mode)
                with torch.no_grad():
                    y = sh(x.to(dev))
                    y2 = sh(x.to(dev))                               # second call: the barriers / symmetric buffers are reusable
                torch.cuda.synchronize()
                bound = tol * scale + (2 * G.ulp(dt) * y_ref.abs() if dt != torch.float32 else 0.0)
                ok = bool(((y.double().cpu() - y_ref).abs() <= bound).all()) and torch.equal(y, y2) and y.shape == (N, outf)
                # the sharded kernels are the unsharded ones on a row subset: same bits as the unsharded layer whenever both take the same kernel family
                same = torch.equal(y, y_full)
                res[(ci, mode)] = (ok, same, float((y.double().cpu() - y_ref).abs().max().item() / scale))
            res[(ci, "status")] = prefill_status(dev)
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_sharded_forward_matches_oracle_in_every_exchange_mode():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(600)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    res = dict(out)
    assert sorted(res) == [0, 1]
    for rank in (0, 1):
        for ci in range(len(CASES)):
            assert res[rank][(ci, "status")] == 0
            for mode in ("nccl", "push", "p2p"):
                ok, same, err = res[rank][(ci, mode)]
                assert ok, f"rank {rank} case {ci} {mode}: rel err {err:.3e}"
            assert res[rank][(ci, "nccl")][1] and res[rank][(ci, "push")][1], f"case {ci}: nccl / push must equal the unsharded layer bit for bit"
