#!/usr/bin/env python3
"""Output-block-row sharding on G GPUs of one box (torchrun, one process per GPU): BASELINE.json configs[4],
Llama-2-70B shapes (8192 / 28672), B=256.  Checks the sharded result against the unsharded layer on the same GPU, then
times local compute, the NCCL all-gather exchange and the fused peer-store exchange (max over ranks, CUDA events).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 tools/sharded_bench.py
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_fft_b200 as bc  # noqa: E402

SHAPES_70B = [("q/o", 8192, 8192), ("k/v", 8192, 1024), ("gate/up", 8192, 28672), ("down", 28672, 8192)]


def ev_time(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--tokens", default="1,2048")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--spectrum", default="bfloat16")
    ap.add_argument("--no-p2p", action="store_true")
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    B = a.block
    sdt = torch.bfloat16 if a.spectrum == "bfloat16" else torch.float32
    for name, inf, outf in SHAPES_70B:
        torch.manual_seed(0)                                        # same weights on every rank
        full = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False, spectrum_dtype=sdt).to(dev).eval()
        modes = ["nccl"] + ([] if a.no_p2p else ["p2p"])
        shards = {m: bc.shard_bc_linear(full, exchange=m) for m in modes}
        for N in (int(v) for v in a.tokens.split(",")):
            x = torch.randn(N, inf, device=dev, generator=torch.Generator(device=dev).manual_seed(1)).half()
            row = {"shape": f"{inf}x{outf}", "name": name, "B": B, "tokens": N, "world": world, "spectrum": a.spectrum}
            with torch.no_grad():
                y_full = full(x)
                row["unsharded_us"] = ev_time(lambda: full(x), a.reps)
                for m, sh in shards.items():
                    try:
                        y = sh(x)
                        err = (y.float() - y_full.float()).abs().max().item() / max(y_full.float().abs().max().item(), 1e-30)
                        row[f"{m}_rel_err_vs_unsharded"] = err
                        row[f"{m}_total_us"] = ev_time(lambda: sh(x), a.reps)
                    except Exception as e:                          # noqa: BLE001
                        row[f"{m}_error"] = f"{type(e).__name__}: {e}"[:200]
                sh = shards["nccl"]
                if sh.local is not None:
                    row["local_compute_us"] = ev_time(lambda: sh.local(x), a.reps)
                    y_loc = sh.local(x)
                else:
                    row["local_compute_us"] = ev_time(lambda: None, a.reps)
                    y_loc = x.new_empty(N, 0)
                row["nccl_gather_only_us"] = ev_time(lambda: bc.gather_columns(y_loc, sh.ranges, B), a.reps)
            if rank == 0:
                print(json.dumps({k: (round(v, 2) if isinstance(v, float) and v > 1 else v) for k, v in row.items()}), flush=True)
        del full, shards
        torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
