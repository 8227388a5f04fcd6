#!/usr/bin/env python3
"""Per-role wait / busy cycles of the wave kernel's first 8 CTAs (library built with `make WAVE_PROF=1`).
producer: dep waits (ring slot free, Xf ready, Yf ready), empty-stage waits per item type; MMA issuer: TMEM-slot waits per type,
full-stage waits per type; epilogue warps 2 and 9: accumulator waits per type, busy cycles per type.  Slot 7 of every role = total."""
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
NAMES = {0: ["dep_ringfree", "dep_xf", "dep_yf", "emptyA", "emptyB", "emptyC", "-", "total"],
         8: ["tslotA", "tslotB", "tslotC", "fullA", "fullB", "fullC", "-", "total"],
         16: ["accwaitA", "accwaitB", "accwaitC", "busyA", "busyB", "busyC", "-", "total"],
         24: ["accwaitA", "accwaitB", "accwaitC", "busyA", "busyB", "busyC", "-", "total"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=16384)
    ap.add_argument("--shape", default="4096x4096")
    ap.add_argument("--block", type=int, default=128)
    a = ap.parse_args()
    N, B = a.tokens, a.block
    inf, outf = (int(v) for v in a.shape.split("x"))
    p, q = outf // B, inf // B
    torch.manual_seed(0)
    layer = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False, spectrum_dtype=torch.bfloat16).to(DEV).eval()
    spec_g = layer.build_spectrum(layout=_cabi.LAYOUT_GEMM)
    x = torch.randn(N, inf, device=DEV).half()
    y = torch.empty(N, outf, device=DEV, dtype=torch.float16)
    prof = torch.zeros(8 * 32, device=DEV, dtype=torch.int64)
    os.environ["BCL_PREFILL_WAVE"] = "1"
    os.environ["BCL_PREFILL_WAVE_MIN_TILES"] = "1"
    os.environ["BCL_WAVE_PROF_PTR"] = str(prof.data_ptr())
    for _ in range(3):
        bc_forward(x, spec_g, None, p, q, B, out=y, layout=_cabi.LAYOUT_GEMM, spectrum_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    pr = prof.cpu().view(8, 32).tolist()
    print(json.dumps({"status": prefill_status(DEV), "lag": os.environ.get("BCL_WAVE_LAG"), "ring": os.environ.get("BCL_WAVE_RING"), "stages": os.environ.get("BCL_WAVE_STAGES")}))
    for cta in (0, 3, 7):
        for base, role in ((0, "producer"), (8, "mma"), (16, "epi_w2"), (24, "epi_w9")):
            print("cta %d %-8s " % (cta, role) + "  ".join("%s=%d" % (n, v) for n, v in zip(NAMES[base], pr[cta][base:base + 8]) if n != "-"))


if __name__ == "__main__":
    main()
