#!/usr/bin/env python3
"""BC linear (B=128, bf16 and fp32 spectra) vs the dense fp16 nn.Linear of the same shape (cuBLAS) over token counts: where the
block-circulant layer wins on the same GPU.  CUDA events, inputs rotated over two buffers."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_fft_b200 as bc
DEV = torch.device("cuda:0")

def ev(fn, reps):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for inf, outf in ((4096, 4096), (4096, 11008), (11008, 4096)):
    torch.manual_seed(0)
    dense = torch.nn.Linear(inf, outf, bias=False, device=DEV, dtype=torch.float16)
    lb = bc.BlockCirculantLinear(inf, outf, block_size=128, bias=False, spectrum_dtype=torch.bfloat16).to(DEV).eval()
    lf = bc.BlockCirculantLinear(inf, outf, block_size=128, bias=False, spectrum_dtype=torch.float32).to(DEV).eval()
    for N in (1, 16, 128, 256, 512, 1024, 2048, 4096, 16384):
        xs = [torch.randn(N, inf, device=DEV).half() for _ in range(2)]
        reps = 50 if N <= 4096 else 10
        with torch.no_grad():
            row = {"shape": f"{inf}x{outf}", "tokens": N, "dense_fp16_us": round(ev(lambda i: dense(xs[i & 1]), reps), 1),
                   "bc_bf16_us": round(ev(lambda i: lb(xs[i & 1]), reps), 1), "bc_fp32_us": round(ev(lambda i: lf(xs[i & 1]), reps), 1)}
        row["speedup_bf16"] = round(row["dense_fp16_us"] / row["bc_bf16_us"], 2)
        row["speedup_fp32"] = round(row["dense_fp16_us"] / row["bc_fp32_us"], 2)
        print(json.dumps(row), flush=True)
