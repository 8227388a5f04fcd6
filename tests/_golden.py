"""Loader for tests/golden/reference_vectors.npz (written by tests/golden/make_golden.py)."""
import os
from functools import lru_cache

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "reference_vectors.npz")
GOLDEN_PT = os.path.join(HERE, "golden", "reference_bc_params.pt")
DT = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}


@lru_cache(maxsize=1)
def vectors():
    return np.load(GOLDEN)


def t(arr, dtype=None):
    x = torch.from_numpy(np.array(arr))
    return x if dtype is None else x.to(dtype)


def forward_cases():
    """Yield dicts: name, in, out, B, dtype, bias_mode, y_dtype, c (fp32), x, bias|None, y, y_cached."""
    v = vectors()
    for row in v["forward_meta"]:
        name, inf, outf, B, dt, bias_mode, y_dt, seed, scale, store_c = str(row).split("|")
        inf, outf, B, seed, scale = int(inf), int(outf), int(B), int(seed), float(scale)
        yield dict(name=name, inf=inf, outf=outf, B=B, dtype=DT[dt], bias_mode=bias_mode, y_dtype=DT[y_dt],
                   seed=seed, scale=scale, store_c=bool(int(store_c)))


def load_case(case):
    v = vectors()
    name, B = case["name"], case["B"]
    p, q = case["outf"] // B, case["inf"] // B
    if case["store_c"]:
        c = t(v[f"{name}/c"], torch.float32)
    else:
        g = torch.Generator().manual_seed(case["seed"])
        c = torch.randn(p, q, B, generator=g) * case["scale"]
        chk = v[f"{name}/c_checksum"]
        got = np.array([c.double().sum().item(), c.double().abs().sum().item()])
        assert np.allclose(got, chk, rtol=0, atol=1e-9 * max(1.0, abs(chk[1]))), \
            "torch CPU RNG drifted from the version the golden vectors were drawn with"
    x = t(v[f"{name}/x"], case["dtype"])
    bias = None
    if case["bias_mode"] != "none":
        bias = t(v[f"{name}/bias"], case["dtype"] if case["bias_mode"] == "same" else torch.float32)
    y = t(v[f"{name}/y"], case["y_dtype"])
    y_cached = t(v[f"{name}/y_cached"], case["y_dtype"])
    return c, x, bias, y, y_cached


def ulp(dtype):
    return {torch.float32: 2.0 ** -23, torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}[dtype]
