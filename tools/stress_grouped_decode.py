#!/usr/bin/env python3
"""Randomised stress of the grouped decode kernels (split-q fp32 kernel, row-per-warp bf16 kernel): random block sizes, shapes, problem
counts and CTA counts; every problem is compared with its own per-module forward.  Run under `timeout`: a protocol bug would hang.
    timeout 300 python tools/stress_grouped_decode.py [configs] [seed]"""
import os, sys, random
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_fft_b200 as bc
from llama_fft_b200.patch_llama_fft import bc_forward
DEV = torch.device("cuda:0")
n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
g = torch.Generator().manual_seed(1)
worst = 0.0
for it in range(n_cfg):
    B = rng.choice([64, 128, 256])
    sdt = rng.choice([torch.float32, torch.bfloat16])
    qcap = {64: 172, 128: 86, 256: 43}[B]
    nprob = rng.randint(1, 40)
    shapes = [(rng.randint(1, 48), rng.randint(1, qcap)) for _ in range(nprob)]
    num_ctas = rng.choice([1, 2, 7, 37, 148, 148, 148])
    layers, xs = [], []
    for (p, q) in shapes:
        m = bc.BlockCirculantLinear(q * B, p * B, block_size=B, bias=False, spectrum_dtype=sdt)
        with torch.no_grad():
            m.c.copy_(torch.randn(p, q, B, generator=g) * 0.02)
        layers.append(m.to(DEV).eval())
        xs.append(torch.randn(1, q * B, generator=g).half().to(DEV))
    specs = [m.build_spectrum() for m in layers]
    grp = bc.GroupedBCForward(xs, specs, num_ctas=num_ctas)
    ys = grp.run()
    torch.cuda.synchronize()
    for m, x, s, y, (p, q) in zip(layers, xs, specs, ys, shapes):
        ref = bc_forward(x, s, None, p, q, B)
        err = (y.float() - ref.float()).abs().max().item() / max(ref.float().abs().max().item(), 1e-30)
        worst = max(worst, err)
        assert err <= 2e-3, (it, B, sdt, p, q, num_ctas, err)          # same math, other summation order: <= 2 fp16 ulp
    print(f"cfg {it}: B={B} {str(sdt)[6:]} problems={nprob} ctas={num_ctas} ok", flush=True)
print("stress OK, worst rel diff", worst)
