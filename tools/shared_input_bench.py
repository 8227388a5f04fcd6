#!/usr/bin/env python3
"""Shared-input fusion timing (llama-fft_b200/shared_input.py): q/k/v and gate/up of one Llama-2-7B layer, B=128, as separate BC
forwards vs one fused forward, at prefill (16384 tokens, bf16 spectra, tcgen05 path) and decode (1 token, fp32 spectra).
CUDA events; prefill inputs rotate over two buffer sets larger than L2; decode numbers are per-module launches (no grouping)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_fft_b200 as bc  # noqa: E402

DEV = torch.device("cuda:0")


def timeit(fn, reps, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    B = 128
    for name, inf, outs in (("qkv", 4096, (4096, 4096, 4096)), ("gate_up", 4096, (11008, 11008))):
        for N, sdt, reps in ((16384, torch.bfloat16, 10), (1, torch.float32, 200)):
            torch.manual_seed(0)
            ls = [bc.BlockCirculantLinear(inf, o, block_size=B, bias=False, spectrum_dtype=sdt).to(DEV).eval() for o in outs]
            xs = [torch.randn(N, inf, device=DEV).half() for _ in range(2)]
            with torch.no_grad():
                sep = timeit(lambda i: [m(xs[i % 2]) for m in ls], reps)
                g = bc.SharedInputGroup(ls)
                for m in ls:
                    m._shared_group = g
                fused = timeit(lambda i: [m(xs[i % 2]) for m in ls], reps)
            print(json.dumps({"group": name, "tokens": N, "spectrum": str(sdt), "separate_us": round(sep, 1), "fused_us": round(fused, 1),
                              "speedup": round(sep / fused, 3)}), flush=True)
            del ls, xs, g
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
