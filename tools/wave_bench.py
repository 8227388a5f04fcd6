#!/usr/bin/env python3
"""Wave-schedule prefill kernel (bcl_prefill_wave.cu) against the three-launch chain: bit-identity of the outputs and time per
forward, one BC linear per shape, 16-bit activations, bf16 spectra.  CUDA events, inputs rotated over two buffer sets.
  python tools/wave_bench.py --shapes 4096x4096 --tokens 16384 --lags 16,23,32 --rings 0 --stages 3"""
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--block", type=int, default=128)
    ap.add_argument("--shapes", default="4096x4096,4096x11008,11008x4096")
    ap.add_argument("--lags", default="0")          # 0 = the library's default
    ap.add_argument("--rings", default="0")
    ap.add_argument("--stages", default="0")
    ap.add_argument("--cpt", default="0")           # chunks per tile
    ap.add_argument("--dtype", default="f16")
    a = ap.parse_args()
    N, B = a.tokens, a.block
    dt = torch.float16 if a.dtype == "f16" else torch.bfloat16
    peak = 6552.6
    for shp in a.shapes.split(","):
        inf, outf = (int(v) for v in shp.split("x"))
        p, q = outf // B, inf // B
        torch.manual_seed(0)
        layer = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False, spectrum_dtype=torch.bfloat16).to(DEV).eval()
        spec_g = layer.build_spectrum(layout=_cabi.LAYOUT_GEMM)
        xs = [torch.randn(N, inf, device=DEV).to(dt) for _ in range(2)]
        ys = [torch.empty(N, outf, device=DEV, dtype=dt) for _ in range(2)]
        alg = N * inf * 2 + p * q * B * 2 + N * outf * 2

        def fwd(i):
            bc_forward(xs[i % 2], spec_g, None, p, q, B, out=ys[i % 2], layout=_cabi.LAYOUT_GEMM, spectrum_dtype=torch.bfloat16)

        os.environ["BCL_PREFILL_WAVE"] = "0"
        t_chain = timeit(fwd, a.reps)
        fwd(0)
        torch.cuda.synchronize()
        ref = ys[0].clone()
        print(json.dumps({"shape": shp, "tokens": N, "impl": "chain", "us": round(t_chain, 2), "hbm_frac": round(alg / peak / 1e3 / t_chain, 4),
                          "status": prefill_status(DEV)}), flush=True)
        os.environ["BCL_PREFILL_WAVE"] = "1"
        os.environ["BCL_PREFILL_WAVE_MIN_TILES"] = "1"
        for lag in a.lags.split(","):
            for ring in a.rings.split(","):
                for stg in a.stages.split(","):
                    for cpt in a.cpt.split(","):
                        for k, v in (("BCL_WAVE_LAG", lag), ("BCL_WAVE_RING", ring), ("BCL_WAVE_STAGES", stg), ("BCL_WAVE_CHUNKS_PER_TILE", cpt)):
                            if v != "0":
                                os.environ[k] = v
                            else:
                                os.environ.pop(k, None)
                        ys[0].zero_()
                        fwd(0)
                        torch.cuda.synchronize()
                        st = prefill_status(DEV)
                        same = bool(torch.equal(ys[0], ref))
                        maxd = float((ys[0].float() - ref.float()).abs().max())
                        t = timeit(fwd, a.reps) if st == 0 else float("nan")
                        print(json.dumps({"shape": shp, "tokens": N, "impl": "wave", "lag": lag, "ring": ring, "stages": stg, "cpt": cpt, "us": round(t, 2),
                                          "hbm_frac": round(alg / peak / 1e3 / t, 4) if st == 0 else None, "bit_identical": same, "max_abs_diff": maxd,
                                          "status": st, "status_after": prefill_status(DEV)}), flush=True)
        del xs, ys, layer
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
