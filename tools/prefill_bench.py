#!/usr/bin/env python3
"""Prefill (N = 8 x 2048 tokens) timing of one BC linear per Llama-2-7B shape, B=128, fp16 activations: the tcgen05 GEMM-layout
path, its three stages alone, the CUDA-core stream kernel, and -- as the kernel-level bar on the same GPU -- the reference's
torch.fft/einsum formulation (cuFFT + cuBLAS, rfft(c) cached).  CUDA events, inputs rotated over buffers larger than L2."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_fft_b200 as bc  # noqa: E402
from llama_fft_b200 import _cabi  # noqa: E402
from llama_fft_b200.patch_llama_fft import bc_forward, prefill_status  # noqa: E402

DEV = torch.device("cuda:0")


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--block", type=int, default=128)
    ap.add_argument("--shapes", default="4096x4096,4096x11008,11008x4096")
    ap.add_argument("--stage-tokens", type=int, default=2048)
    ap.add_argument("--spectrum", default="bfloat16", choices=["bfloat16", "float32"], help="float32: the split-bf16 (fp32-accurate) GEMM chain")
    ap.add_argument("--skip-torch", action="store_true")
    ap.add_argument("--skip-stream", action="store_true")
    args = ap.parse_args()
    N, B = args.tokens, args.block
    peak = 6552.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    L = _cabi.lib()
    out = []
    for shp in args.shapes.split(","):
        inf, outf = (int(v) for v in shp.split("x"))
        p, q = outf // B, inf // B
        torch.manual_seed(0)
        sdt = torch.bfloat16 if args.spectrum == "bfloat16" else torch.float32
        layer = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False, spectrum_dtype=sdt).to(DEV).eval()
        nbuf = 2
        xs = [torch.randn(N, inf, device=DEV).half() for _ in range(nbuf)]
        ys = [torch.empty(N, outf, device=DEV, dtype=torch.float16) for _ in range(nbuf)]
        alg = N * inf * 2 + p * q * B * (2 if sdt == torch.bfloat16 else 4) + N * outf * 2
        row = {"shape": shp, "B": B, "tokens": N, "algorithmic_MB": alg / 1e6, "t_hbm_us": alg / peak / 1e3}
        spec_g = layer.build_spectrum(layout=_cabi.LAYOUT_GEMM)
        spec_s = layer.build_spectrum()

        def gemm_path(i):
            bc_forward(xs[i % nbuf], spec_g, None, p, q, B, out=ys[i % nbuf], layout=_cabi.LAYOUT_GEMM, spectrum_dtype=sdt)
        row["gemm_path_us"] = timeit(gemm_path, args.reps)
        row["gemm_path_status"] = prefill_status(DEV)
        row["gemm_path_hbm_frac"] = row["t_hbm_us"] / row["gemm_path_us"]
        row["gemm_path_Mtok_s"] = N / row["gemm_path_us"]

        row["spectrum"] = args.spectrum
        if sdt != torch.bfloat16:
            if not args.skip_stream:
                row["stream_kernel_f32spec_us"] = timeit(lambda i: bc_forward(xs[i % nbuf], spec_s, None, p, q, B, out=ys[i % nbuf]), 3, warm=1)
            print(json.dumps(row), flush=True)
            out.append(row)
            del xs, ys, layer
            torch.cuda.empty_cache()
            continue
        # stages alone on one 2048-token chunk (what the chunk loop runs)
        nc = min(N, args.stage_tokens)
        F, Kp, Np = _cabi.prefill_geometry(p, q, B)
        Xf = torch.empty(F, Kp // 8, (nc + 31) // 32 * 32, 8, device=DEV, dtype=torch.bfloat16)
        Yf = torch.empty(F // 4, p, (nc + 31) // 32 * 32, 8, device=DEV, dtype=torch.bfloat16)
        status = torch.zeros(1, device=DEV, dtype=torch.int32)
        st = _cabi.stream_ptr(DEV)
        row["stage_tokens"] = nc
        row["stage_rfft_us"] = timeit(lambda i: L.bcl_prefill_rfft(xs[i % nbuf].data_ptr(), _cabi.F16, nc, inf, p, q, B, spec_g.data_ptr(), 0, Xf.data_ptr(), status.data_ptr(), st), args.reps)
        row["stage_gemm_us"] = timeit(lambda i: L.bcl_prefill_gemm(Xf.data_ptr(), spec_g.data_ptr(), Yf.data_ptr(), nc, p, q, B, status.data_ptr(), st), args.reps)
        row["stage_irfft_us"] = timeit(lambda i: L.bcl_prefill_irfft(Yf.data_ptr(), nc, p, q, B, spec_g.data_ptr(), 0, _cabi.F16, None, 0, ys[i % nbuf].data_ptr(), _cabi.F16, outf, status.data_ptr(), st), args.reps)

        if not args.skip_stream:
            def stream_path(i):
                bc_forward(xs[i % nbuf], spec_s, None, p, q, B, out=ys[i % nbuf])
            row["stream_kernel_bf16spec_us"] = timeit(stream_path, max(2, args.reps // 3), warm=1)
        if not args.skip_torch:
            c32 = layer.c.detach().float()
            Cf = torch.fft.rfft(c32, dim=-1)

            def torch_path(i):                                   # the reference's formulation, src/patch_llama_fft.py:246-266
                x32 = xs[i % nbuf].reshape(N, q, B).to(torch.float32)
                Xf_ = torch.fft.rfft(x32, dim=-1)
                Yf_ = torch.einsum("oif,nif->nof", Cf, Xf_)
                ys[i % nbuf].copy_(torch.fft.irfft(Yf_, n=B, dim=-1).reshape(N, outf))
            try:
                row["torch_cufft_cublas_us"] = timeit(torch_path, max(2, args.reps // 3), warm=1)
            except Exception as e:                               # noqa: BLE001
                row["torch_cufft_cublas_us"] = "failed: %s" % type(e).__name__
        print(json.dumps(row), flush=True)
        out.append(row)
        del xs, ys, layer
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
