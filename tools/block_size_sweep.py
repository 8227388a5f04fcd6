#!/usr/bin/env python3
"""Block-size sweep (BASELINE.json configs[3]; mirrors the reference's bench_block_sizes / big-bench block list 64,128,256):
one square BC linear, in = out in {4096, 11008}, B in {64, 128, 256}, at decode (1 token, fp32 and bf16 spectra; 32 distinct
layers in one grouped launch, two weight sets alternating so nothing is served from L2) and at prefill (16384 tokens, fp16
activations, bf16 spectra).  Per point: time, algorithmic bytes and the fraction of the measured HBM peak.
One JSON line per point; CUDA events."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_fft_b200 as bc  # noqa: E402
from llama_fft_b200 import _cabi  # noqa: E402
from llama_fft_b200.patch_llama_fft import bc_forward, prefill_status  # noqa: E402

DEV = torch.device("cuda:0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6552.6


def ev(fn, reps, warm=3):
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
    peak = peak_gbs()
    L = _cabi.lib()
    sms = torch.cuda.get_device_properties(DEV).multi_processor_count
    n_layers = 32
    for dim in (4096, 11008):
        for B in (64, 128, 256):
            p = q = dim // B
            if p * B != dim:                       # 11008 = 43 * 256 = 86 * 128 = 172 * 64: always divisible
                continue
            # ---- decode: 32 layers' worth of this shape in one grouped launch ----
            for sdt in (torch.float32, torch.bfloat16):
                g = torch.Generator(device=DEV).manual_seed(B + dim)
                groups = []
                for _ in range(2):
                    specs = []
                    for _ in range(n_layers):
                        c = torch.randn(p, q, B, device=DEV, generator=g) * 0.01
                        spec = torch.empty(p, q, B, device=DEV, dtype=sdt)
                        _cabi.check(L.bcl_precompute_spectrum(c.data_ptr(), p, q, B, spec.data_ptr(), _cabi.DTYPE[sdt], _cabi.LAYOUT_STREAM,
                                                              _cabi.stream_ptr(DEV)), "spec")
                        specs.append(spec)
                    grp = bc.GroupedBCForward.from_packed([(p, q)] * n_layers, specs, 1, torch.float16, DEV, num_ctas=sms)
                    grp.x_all.copy_(torch.randn(grp.x_all.numel(), device=DEV, generator=g).to(torch.float16))
                    groups.append(grp)
                us = ev(lambda i: groups[i & 1].run(), 200, warm=10)
                alg = n_layers * (dim * 2 + p * q * B * (4 if sdt == torch.float32 else 2) + dim * 2)
                print(json.dumps({"regime": "decode", "dim": dim, "B": B, "spectrum": str(sdt).replace("torch.", ""), "layers_per_launch": n_layers,
                                  "us_per_launch": round(us, 2), "us_per_layer": round(us / n_layers, 3), "algorithmic_MB": round(alg / 1e6, 2),
                                  "GBps": round(alg / us / 1e3, 1), "hbm_frac": round(alg / us / 1e3 / peak, 3)}), flush=True)
                del groups
                torch.cuda.empty_cache()
            # ---- prefill: 16384 tokens, bf16 spectra ----
            N = 16384
            torch.manual_seed(0)
            layer = bc.BlockCirculantLinear(dim, dim, block_size=B, bias=False, spectrum_dtype=torch.bfloat16).to(DEV).eval()
            spec = layer.build_spectrum(layout=_cabi.LAYOUT_GEMM)
            xs = [torch.randn(N, dim, device=DEV).half() for _ in range(2)]
            ys = [torch.empty(N, dim, device=DEV, dtype=torch.float16) for _ in range(2)]
            us = ev(lambda i: bc_forward(xs[i & 1], spec, None, p, q, B, out=ys[i & 1], layout=_cabi.LAYOUT_GEMM), 5)
            alg = N * dim * 2 * 2 + p * q * B * 2
            flops = 4 * N * p * q * (B - 1)
            print(json.dumps({"regime": "prefill", "dim": dim, "B": B, "tokens": N, "spectrum": "bfloat16", "us": round(us, 1),
                              "Mtok_s": round(N / us, 2), "algorithmic_MB": round(alg / 1e6, 1), "hbm_frac": round(alg / us / 1e3 / peak, 3),
                              "algorithmic_TFLOPs": round(flops / us / 1e6, 1), "dft_stages": "tcgen05" if B <= 128 else "tcgen05, two half-N launches per stage",
                              "status": prefill_status(DEV)}), flush=True)
            del layer, spec, xs, ys
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
