#!/usr/bin/env python3
"""Debug: per-item clock64 stamps of the first CTA of each role inside the fused prefill launch (BCL_TRACE_PTR / BCL_TRACE_STAGE).
   Needs a tracing build of the library: make -C llama-fft_b200/csrc clean && make -C llama-fft_b200/csrc TRACE=1
   python tools/trace_fused.py [tokens] [first_item_to_print]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_fft_b200 as bc
from llama_fft_b200 import _cabi
from llama_fft_b200.patch_llama_fft import bc_forward, prefill_status
DEV = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
i0 = int(sys.argv[2]) if len(sys.argv) > 2 else 8
inf, outf, B = 4096, 4096, 128
p, q = outf // B, inf // B
layer = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False, spectrum_dtype=torch.bfloat16).to(DEV).eval()
spec_g = layer.build_spectrum(layout=_cabi.LAYOUT_GEMM)
x = torch.randn(n, inf, device=DEV).half()
y = torch.empty(n, outf, device=DEV, dtype=torch.float16)
names = ["ld_begin", "ld_issued", "mma_tmem_free", "mma_data0", "mma_commit", "epi_full", "epi_release", "epi_end"]
for stage in (1, 2, 3):
    trace = torch.zeros(64 * 8, device=DEV, dtype=torch.int64)
    os.environ.pop("BCL_TRACE_PTR", None)
    for rep in range(3):
        if rep == 2:
            os.environ["BCL_TRACE_PTR"] = hex(trace.data_ptr())
            os.environ["BCL_TRACE_STAGE"] = str(stage)
        bc_forward(x, spec_g, None, p, q, B, out=y, layout=_cabi.LAYOUT_GEMM)
        torch.cuda.synchronize()
    t = trace.cpu().view(64, 8)
    t0 = int(t[0][t[0] > 0].min())
    print(f"== fused role {stage} (status {prefill_status(DEV)})")
    print("item " + " ".join(f"{k:>13s}" for k in names))
    last = max(i for i in range(64) if int(t[i].max()) != 0)
    for i in sorted(set(list(range(0, 3)) + list(range(i0, i0 + 6)) + list(range(max(last - 5, 0), last + 1)))):
        if i > last:
            break
        print(f"{i:4d} " + " ".join(f"{(int(v) - t0) if int(v) else -1:13d}" for v in t[i]))
