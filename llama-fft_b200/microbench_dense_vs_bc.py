"""The reference's layer microbenchmark (``src/microbench_dense_vs_bc.py``) on the B200 path: ``nn.Linear`` against
``BlockCirculantLinear`` for the three Llama-2-7B MLP shapes (4096 -> 11008, 4096 -> 11008, 11008 -> 4096), optional SiLU-MLP pattern,
same flags (``--device --dtype --batch --seq --block_sizes --warmup --iters --threads --bench_mlp``), same timing discipline
(``bench_fn``, :26-39: warm-up, synchronise, ``iters`` calls, synchronise, wall clock) and the same report lines.

Differences: the BC modules run the sm_100a kernels behind include/bclinear.h, so ``--device cuda`` is the only device (there is no CPU
path; time the reference itself for a CPU column -- bench.py's cpu_baseline does); the ``down`` input is drawn once, outside the
timed closure (the reference regenerates it with ``torch.randn`` inside it, :136-138); ``--config1`` adds BASELINE.json configs[0]
(one 4096 x 4096 layer, B = 64, batch 1, fp32 I/O, one module call) and ``--json`` prints the numbers as one JSON object.
"""
import argparse
import json
import time
from typing import Callable

import torch
import torch.nn as nn
import torch.nn.functional as F

from .patch_llama_fft import BlockCirculantLinear


def _sync(device: torch.device) -> None:
    if device.type == "cuda":
        torch.cuda.synchronize(device)


@torch.inference_mode()
def bench_fn(fn: Callable[[], torch.Tensor], *, device: torch.device, warmup: int, iters: int) -> float:
    """ms per call (reference :26-39)."""
    for _ in range(warmup):
        fn()
    _sync(device)
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    _sync(device)
    return (time.perf_counter() - t0) * 1000.0 / float(iters)


def make_dense_linear(in_f: int, out_f: int, *, device, dtype) -> nn.Linear:
    return nn.Linear(in_f, out_f, bias=True).to(device=device, dtype=dtype).eval()


def make_bc_linear(in_f: int, out_f: int, B: int, *, device, dtype, spectrum_dtype=torch.float32) -> BlockCirculantLinear:
    """Reference :48-51: bias=True, parameters cast to the activation dtype."""
    return BlockCirculantLinear(in_f, out_f, block_size=B, bias=True, spectrum_dtype=spectrum_dtype).to(device=device, dtype=dtype).eval()


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--device", type=str, default="cuda", choices=["cpu", "cuda"])
    p.add_argument("--dtype", type=str, default="", help="float32|float16|bfloat16 (default float16, the reference's CUDA default)")
    p.add_argument("--batch", type=int, default=2)
    p.add_argument("--seq", type=int, default=128)
    p.add_argument("--block_sizes", type=str, default="64,128,256")
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--iters", type=int, default=50)
    p.add_argument("--threads", type=int, default=0, help="accepted for compatibility (CPU threads of the reference's CPU mode)")
    p.add_argument("--bench_mlp", action="store_true", help="also the LLaMA-style MLP: down(silu(gate(x)) * up(x))")
    p.add_argument("--spectrum", type=str, default="float32", choices=["float32", "bfloat16"])
    p.add_argument("--config1", action="store_true", help="BASELINE.json configs[0]: 4096x4096, B=64, batch 1, fp32")
    p.add_argument("--json", action="store_true")
    args = p.parse_args(argv)
    if args.device != "cuda":
        raise SystemExit("the block-circulant forward runs as sm_100a CUDA kernels only (no CPU path): use --device cuda; for a CPU column time "
                         "the reference itself (bench.py's cpu_baseline / --impl reference)")
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA requested but not available")
    device = torch.device("cuda")
    dtype = {"": torch.float16, "float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}[args.dtype.lower().strip()]
    sdt = torch.float32 if args.spectrum == "float32" else torch.bfloat16
    Bs = [int(v) for v in args.block_sizes.split(",") if v.strip()]
    in_f, mid_f, out_f = 4096, 11008, 4096
    x = torch.randn(args.batch, args.seq, in_f, device=device, dtype=dtype)
    tmp = torch.randn(args.batch, args.seq, mid_f, device=device, dtype=dtype)
    kw = dict(device=device, warmup=args.warmup, iters=args.iters)
    res = {"device": str(device), "dtype": str(dtype), "batch": args.batch, "seq": args.seq, "warmup": args.warmup, "iters": args.iters, "dense": {}, "bc": {}}
    say = (lambda *a: None) if args.json else print
    say(f"Device: {device}, dtype: {dtype}, batch={args.batch}, seq={args.seq}, warmup={args.warmup}, iters={args.iters}")
    dg, du, dd = (make_dense_linear(in_f, mid_f, device=device, dtype=dtype), make_dense_linear(in_f, mid_f, device=device, dtype=dtype),
                  make_dense_linear(mid_f, out_f, device=device, dtype=dtype))
    say("\n=== Dense baselines ===")
    for name, fn in (("gate 4096->11008", lambda: dg(x)), ("up   4096->11008", lambda: du(x)), ("down 11008->4096", lambda: dd(tmp))):
        res["dense"][name] = bench_fn(fn, **kw)
        say(f"{name:20s}: {res['dense'][name]:9.3f} ms/iter")
    dense_mlp = lambda: dd(F.silu(dg(x)) * du(x))                     # noqa: E731
    if args.bench_mlp:
        res["dense"]["MLP"] = bench_fn(dense_mlp, **kw)
        say(f"{'MLP (dense)':20s}: {res['dense']['MLP']:9.3f} ms/iter")
    for B in Bs:
        say(f"\n=== Block-circulant B={B} ===")
        bg, bu, bd = (make_bc_linear(in_f, mid_f, B, device=device, dtype=dtype, spectrum_dtype=sdt),
                      make_bc_linear(in_f, mid_f, B, device=device, dtype=dtype, spectrum_dtype=sdt),
                      make_bc_linear(mid_f, out_f, B, device=device, dtype=dtype, spectrum_dtype=sdt))
        row = {}
        for name, fn in (("gate 4096->11008", lambda: bg(x)), ("up   4096->11008", lambda: bu(x)), ("down 11008->4096", lambda: bd(tmp))):
            row[name] = bench_fn(fn, **kw)
            say(f"{name + ' | dense':28s}: {res['dense'][name]:9.3f} ms/iter")
            say(f"{name + ' | bc':28s}: {row[name]:9.3f} ms/iter")
        if args.bench_mlp:
            row["MLP"] = bench_fn(lambda: bd(F.silu(bg(x)) * bu(x)), **kw)
            say(f"{'MLP (dense)':28s}: {res['dense']['MLP']:9.3f} ms/iter")
            say(f"{'MLP (bc)':28s}: {row['MLP']:9.3f} ms/iter")
            say(f"{'BC/Dense ratio':28s}: {row['MLP'] / res['dense']['MLP']:9.3f}  ( <1.0 means BC faster )")
        res["bc"][str(B)] = row
    if args.config1:                                                 # BASELINE.json configs[0]
        lin = make_dense_linear(4096, 4096, device=device, dtype=torch.float32)
        bcl = BlockCirculantLinear(4096, 4096, block_size=64, bias=True, spectrum_dtype=sdt).to(device).eval()
        x1 = torch.randn(1, 1, 4096, device=device)
        res["config1"] = {"what": "4096x4096, B=64, batch 1, fp32 I/O, one module call", "dense_ms": bench_fn(lambda: lin(x1), **kw),
                          "bc_ms": bench_fn(lambda: bcl(x1), **kw)}
        say(f"\n=== BASELINE config 1 ===\n{'4096x4096 B=64 | dense':28s}: {res['config1']['dense_ms']:9.4f} ms/iter\n{'4096x4096 B=64 | bc':28s}: {res['config1']['bc_ms']:9.4f} ms/iter")
    if args.json:
        print(json.dumps(res))
    return res


if __name__ == "__main__":
    main()
