#!/usr/bin/env python3
"""bf16-spectrum batch-1 decode of all 224 projections (BASELINE configs[1]) per launch: the row-per-warp CUDA-core kernel (STREAM
layout) against the tensor-core decode kernel (BCL_LAYOUT_DECODE_TC, bcl_stream_tc.cu).  CUDA events, two alternating weight sets."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_fft_b200 as bc  # noqa: E402
from llama_fft_b200 import _cabi  # noqa: E402

DEV = torch.device("cuda:0")
B = 128
LAYER = [(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)]


def build(layout, sets=2, layers=32):
    groups = []
    for s in range(sets):
        torch.manual_seed(s)
        xs, specs, shapes = [], [], []
        for _ in range(layers):
            for (p, q) in LAYER:
                c = torch.randn(p, q, B, device=DEV) * 0.01
                if layout == _cabi.LAYOUT_DECODE_TC:
                    specs.append(bc.decode_tc_spectrum(c))
                else:
                    spec = torch.empty(p, q, B, device=DEV, dtype=torch.bfloat16)
                    _cabi.check(_cabi.lib().bcl_precompute_spectrum(c.data_ptr(), p, q, B, spec.data_ptr(), _cabi.BF16, _cabi.LAYOUT_STREAM, _cabi.stream_ptr(DEV)), "spec")
                    specs.append(spec)
                xs.append(torch.randn(1, q * B, device=DEV).half())
                shapes.append((p, q))
        kw = dict(layout=layout, shapes=shapes, block_size=B) if layout == _cabi.LAYOUT_DECODE_TC else {}
        groups.append(bc.GroupedBCForward(xs, specs, **kw))
    return groups


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    peak = 6552.6
    for name, layout in (("stream_rw (CUDA cores)", _cabi.LAYOUT_STREAM), ("stream_tc (tcgen05)", _cabi.LAYOUT_DECODE_TC)):
        groups = build(layout)
        for g in groups:
            g.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(3):
            e0.record()
            for i in range(steps):
                groups[i & 1].run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / steps * 1e3)
        us = sorted(ts)[1]
        alg = groups[0].algorithmic_bytes
        print(json.dumps({"kernel": name, "us_per_step": round(us, 2), "tokens_per_s": round(1e6 / us), "algorithmic_MB": round(alg / 1e6, 2),
                          "hbm_frac_of_measured_peak": round(alg / (us * 1e-6) / 1e9 / peak, 4)}), flush=True)
        del groups
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
