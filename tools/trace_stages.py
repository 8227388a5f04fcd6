#!/usr/bin/env python3
"""Debug: per-item clock64 stamps of CTA 0 of each tcgen05 prefill stage (BCL_TRACE_PTR / BCL_TRACE_STAGE), as cycle offsets.
   Needs a tracing build of the library: make -C llama-fft_b200/csrc clean && make -C llama-fft_b200/csrc TRACE=1
   python tools/trace_stages.py [tokens]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llama_fft_b200 import _cabi
DEV = torch.device("cuda:0")
n, inf, outf, B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 4096, 4096, 128
p, q = outf // B, inf // B
L = _cabi.lib()
F, Kp, Np = _cabi.prefill_geometry(p, q, B)
st = _cabi.stream_ptr(DEV)
c = torch.randn(p, q, B, device=DEV) * 0.01
spec = torch.empty(L.bcl_spectrum_bytes(p, q, B, _cabi.BF16, _cabi.LAYOUT_GEMM), device=DEV, dtype=torch.uint8)
_cabi.check(L.bcl_precompute_spectrum(c.data_ptr(), p, q, B, spec.data_ptr(), _cabi.BF16, _cabi.LAYOUT_GEMM, st), "spec")
status = torch.zeros(1, device=DEV, dtype=torch.int32)
x = torch.randn(n, inf, device=DEV).half()
y = torch.empty(n, outf, device=DEV, dtype=torch.float16)
Xf = torch.empty(F, Kp // 8, (n + 31) // 32 * 32, 8, device=DEV, dtype=torch.bfloat16)
Yf = torch.empty(F // 4, p, (n + 31) // 32 * 32, 8, device=DEV, dtype=torch.bfloat16)
calls = {
    1: lambda: L.bcl_prefill_rfft(x.data_ptr(), _cabi.F16, n, inf, p, q, B, spec.data_ptr(), 0, Xf.data_ptr(), status.data_ptr(), st),
    2: lambda: L.bcl_prefill_gemm(Xf.data_ptr(), spec.data_ptr(), Yf.data_ptr(), n, p, q, B, status.data_ptr(), st),
    3: lambda: L.bcl_prefill_irfft(Yf.data_ptr(), n, p, q, B, spec.data_ptr(), 0, _cabi.F16, None, 0, y.data_ptr(), _cabi.F16, outf, status.data_ptr(), st),
}
names = ["ld_begin", "ld_issued", "mma_tmem_free", "mma_data0", "mma_commit", "epi_full", "epi_release", "epi_end"]
for stage in (1, 2, 3):
    trace = torch.zeros(64 * 8, device=DEV, dtype=torch.int64)
    os.environ.pop("BCL_TRACE_PTR", None)
    for rep in range(3):
        if rep == 2:
            os.environ["BCL_TRACE_PTR"] = hex(trace.data_ptr())
            os.environ["BCL_TRACE_STAGE"] = str(stage)
        calls[stage]()
        torch.cuda.synchronize()
    t = trace.cpu().view(64, 8)
    t0 = int(t[0][t[0] > 0].min())
    print(f"== stage {stage} (status {int(status.item())})")
    print("item " + " ".join(f"{k:>13s}" for k in names))
    for i in range(12):
        if int(t[i].max()) == 0:
            break
        print(f"{i:4d} " + " ".join(f"{(int(v) - t0) if int(v) else -1:13d}" for v in t[i]))
