#!/usr/bin/env python3
"""Prefill path under CUDA-graph replay (no host launch cost in the timed region): the whole chunked chain at the chunk budget
given by BCL_PREFILL_CHUNK_MB (read once per process), and each stage alone on L2-resident token counts.  Separates what
the GPU does from what the host's ~8 us per launch costs (profiles/r1_prefill_notes.md)."""
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


def graph_time(body, reps):
    """body() enqueues the work once; returns us per body() under graph replay"""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            body()
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=16384)
    ap.add_argument("--shape", default="4096x4096")
    ap.add_argument("--block", type=int, default=128)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--stage-tokens", default="")
    a = ap.parse_args()
    N, B = a.tokens, a.block
    inf, outf = (int(v) for v in a.shape.split("x"))
    p, q = outf // B, inf // B
    L = _cabi.lib()
    torch.manual_seed(0)
    layer = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False, spectrum_dtype=torch.bfloat16).to(DEV).eval()
    spec_g = layer.build_spectrum(layout=_cabi.LAYOUT_GEMM)
    xs = [torch.randn(N, inf, device=DEV).half() for _ in range(2)]
    ys = [torch.empty(N, outf, device=DEV, dtype=torch.float16) for _ in range(2)]
    row = {"shape": a.shape, "tokens": N, "chunk_mb": os.environ.get("BCL_PREFILL_CHUNK_MB", "default")}

    def chain():
        for i in range(2):
            bc_forward(xs[i], spec_g, None, p, q, B, out=ys[i], layout=_cabi.LAYOUT_GEMM)
    row["chain_graph_us"] = round(graph_time(chain, a.reps) / 2, 2)
    row["status"] = prefill_status(DEV)
    print(json.dumps(row), flush=True)

    st_tokens = [int(v) for v in a.stage_tokens.split(",") if v]
    F, Kp, Np = _cabi.prefill_geometry(p, q, B)
    status = torch.zeros(1, device=DEV, dtype=torch.int32)
    R = 16
    for nc in st_tokens:
        Xf = torch.empty(F, Kp // 8, (nc + 31) // 32 * 32, 8, device=DEV, dtype=torch.bfloat16)
        Yf = torch.empty(F // 4, p, (nc + 31) // 32 * 32, 8, device=DEV, dtype=torch.bfloat16)
        x, y = xs[0][:nc], ys[0][:nc]
        r = {"shape": a.shape, "stage_tokens": nc, "MB_in_plus_out": round(nc * (inf + outf) * 2 / 1e6, 1)}

        def run(fn):
            with torch.cuda.stream(torch.cuda.current_stream()):
                pass
            def body():
                st = _cabi.stream_ptr(DEV)
                for _ in range(R):
                    fn(st)
            return round(graph_time(body, a.reps) / R, 2)
        r["rfft_us"] = run(lambda st: L.bcl_prefill_rfft(x.data_ptr(), _cabi.F16, nc, inf, p, q, B, spec_g.data_ptr(), 0, Xf.data_ptr(), status.data_ptr(), st))
        r["gemm_us"] = run(lambda st: L.bcl_prefill_gemm(Xf.data_ptr(), spec_g.data_ptr(), Yf.data_ptr(), nc, p, q, B, status.data_ptr(), st))
        r["irfft_us"] = run(lambda st: L.bcl_prefill_irfft(Yf.data_ptr(), nc, p, q, B, spec_g.data_ptr(), 0, _cabi.F16, None, 0, y.data_ptr(), _cabi.F16, outf, status.data_ptr(), st))
        r["status"] = int(status.item())
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
