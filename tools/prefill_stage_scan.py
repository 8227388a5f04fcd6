#!/usr/bin/env python3
"""Stage timing scan of the prefill path (one shape): each stage alone at several token counts, to separate fixed per-launch
cost from throughput.  CUDA events; host launch cost measured separately (time.perf_counter without sync)."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llama_fft_b200 import _cabi  # noqa: E402

DEV = torch.device("cuda:0")


def ev_time(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / reps * 1e6
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, host


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="4096x4096")
    ap.add_argument("--block", type=int, default=128)
    ap.add_argument("--tokens", default="256,2048,4096,16384")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--impl", type=int, default=0)
    a = ap.parse_args()
    inf, outf = (int(v) for v in a.shape.split("x"))
    B = a.block
    p, q = outf // B, inf // B
    L = _cabi.lib()
    F, Kp, Np = _cabi.prefill_geometry(p, q, B)
    st = _cabi.stream_ptr(DEV)
    c = torch.randn(p, q, B, device=DEV) * 0.01
    spec = torch.empty(L.bcl_spectrum_bytes(p, q, B, _cabi.BF16, _cabi.LAYOUT_GEMM), device=DEV, dtype=torch.uint8)
    _cabi.check(L.bcl_precompute_spectrum(c.data_ptr(), p, q, B, spec.data_ptr(), _cabi.BF16, _cabi.LAYOUT_GEMM, st), "spec")
    status = torch.zeros(1, device=DEV, dtype=torch.int32)
    for n in (int(v) for v in a.tokens.split(",")):
        x = torch.randn(n, inf, device=DEV).half()
        y = torch.empty(n, outf, device=DEV, dtype=torch.float16)
        Xf = torch.empty(F, Kp // 8, (n + 31) // 32 * 32, 8, device=DEV, dtype=torch.bfloat16)
        Yf = torch.empty(F // 4, p, (n + 31) // 32 * 32, 8, device=DEV, dtype=torch.bfloat16)
        r = {"shape": a.shape, "tokens": n}
        r["rfft_us"], r["rfft_host_us"] = ev_time(lambda: L.bcl_prefill_rfft(x.data_ptr(), _cabi.F16, n, inf, p, q, B, spec.data_ptr(), a.impl, Xf.data_ptr(), status.data_ptr(), st), a.reps)
        r["gemm_us"], r["gemm_host_us"] = ev_time(lambda: L.bcl_prefill_gemm(Xf.data_ptr(), spec.data_ptr(), Yf.data_ptr(), n, p, q, B, status.data_ptr(), st), a.reps)
        r["irfft_us"], r["irfft_host_us"] = ev_time(lambda: L.bcl_prefill_irfft(Yf.data_ptr(), n, p, q, B, spec.data_ptr(), a.impl, _cabi.F16, None, 0, y.data_ptr(), _cabi.F16, outf, status.data_ptr(), st), a.reps)
        r["status"] = int(status.item())
        print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)


if __name__ == "__main__":
    main()
