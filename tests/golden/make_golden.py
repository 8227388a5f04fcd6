#!/usr/bin/env python3
"""Generate tests/golden/*.npz|*.pt by running the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

It imports /root/reference/src/{patch_llama_fft,fft_utils,bench_all_in_one}.py
and records seeded input/output pairs of the hot path
(BlockCirculantLinear.forward, src/patch_llama_fft.py:226-275; forward_cached,
src/bench_all_in_one.py:179-231), the single-block primitive
(src/fft_utils.py:3-34), the projections (src/patch_llama_fft.py:47-128),
from_linear (:180-224) and a save_bc_params checkpoint (:287-300).

float16 tensors are stored as float16, bfloat16 up-cast to float32 (exact), with the dtype name
beside them.  Large parameters (Llama-shaped c) are not stored: they are
re-drawn from the recorded seed with the recorded recipe and checked against a
recorded float64 checksum, so RNG drift is detected rather than trusted.
"""
import io
import os
import sys
import contextlib

import numpy as np
import torch

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

import patch_llama_fft as plf            # noqa: E402
import fft_utils as rfu                  # noqa: E402
import bench_all_in_one as bai           # noqa: E402

DT = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}


def f32(t):
    """float16 stays float16 (numpy has it); bfloat16 is up-cast to float32 (exact)."""
    t = t.detach().cpu()
    return t.numpy() if t.dtype == torch.float16 else t.to(torch.float32).numpy()


def draw_c(seed, p, q, B, scale):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(p, q, B, generator=g) * scale


def draw_x(seed, shape, dtype):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(dtype)


def forward_cases():
    """(name, in, out, B, x_shape, x_dtype, bias_mode, c_scale, store_c)"""
    cases = []
    k = 0
    # the reference's own test-5 shape, and small shapes over every B the path supports
    for (inf, outf, B) in [(128, 64, 32), (256, 128, 64), (256, 384, 128), (512, 256, 256), (64, 64, 64)]:
        for shape in [(4,), (1,), (2, 3)]:
            for dt in ["float32", "float16", "bfloat16"]:
                for bias in ["none", "same", "fp32"]:
                    if bias == "fp32" and dt == "float32":
                        continue
                    if B == 32 and (dt != "float32"):
                        continue
                    cases.append((f"small{k}", inf, outf, B, shape + (inf,), dt, bias, 1.0, True))
                    k += 1
    # Llama-2-7B / sweep / 70B shapes at decode-sized N (c re-drawn from seed, reference init scale 0.01)
    for (inf, outf, B) in [(4096, 4096, 64), (4096, 4096, 128), (4096, 11008, 128), (11008, 4096, 128),
                           (4096, 4096, 256), (8192, 1024, 256)]:
        for shape, dt in [((1,), "float32"), ((1, 1), "float16"), ((2,), "bfloat16")]:
            cases.append((f"llama{k}", inf, outf, B, shape + (inf,), dt, "none", 0.01, False))
            k += 1
    return cases


def main():
    out = {}
    meta = []
    torch.manual_seed(1234)
    for idx, (name, inf, outf, B, xshape, dt, bias_mode, scale, store_c) in enumerate(forward_cases()):
        seed = 1000 + idx
        p, q = outf // B, inf // B
        c = draw_c(seed, p, q, B, scale)
        x = draw_x(seed + 500000, xshape, DT[dt])
        layer = plf.BlockCirculantLinear(inf, outf, block_size=B, bias=(bias_mode != "none"))
        with torch.no_grad():
            layer.c.copy_(c)
            if bias_mode != "none":
                g = torch.Generator().manual_seed(seed + 900000)
                layer.bias.copy_(torch.randn(outf, generator=g) * 0.1)
        if bias_mode == "same":
            layer.bias.data = layer.bias.data.to(DT[dt])
        layer.eval()
        with torch.no_grad():
            y = layer(x)
            # cached variant must agree bit-for-bit with the shipped forward on CPU
            bai.enable_cfft_cache(layer)
            y_cached = layer(x)
            y_cached2 = layer(x)
            bai.disable_cfft_cache(layer)
        assert torch.equal(y_cached, y_cached2)
        out[f"{name}/x"] = f32(x)
        out[f"{name}/y"] = f32(y)
        out[f"{name}/y_cached"] = f32(y_cached)
        if bias_mode != "none":
            out[f"{name}/bias"] = f32(layer.bias)
        if store_c:
            out[f"{name}/c"] = f32(c)
        csum = np.array([c.double().sum().item(), c.double().abs().sum().item()], dtype=np.float64)
        out[f"{name}/c_checksum"] = csum
        meta.append("|".join(map(str, [name, inf, outf, B, dt, bias_mode, str(y.dtype).replace("torch.", ""),
                                       seed, scale, int(store_c)])))
    out["forward_meta"] = np.array(meta)

    # ---- single-block primitive, src/fft_utils.py ---------------------------
    prim = []
    for i, n in enumerate([8, 64, 128, 256]):
        for dt in ["float32", "float16"]:
            g = torch.Generator().manual_seed(77 + i)
            c = torch.randn(n, generator=g)
            x = torch.randn(n, generator=g).to(DT[dt])
            y = rfu.circulant_matvec_fft(c, x)
            key = f"prim{i}_{dt}"
            out[f"{key}/c"], out[f"{key}/x"], out[f"{key}/y"] = f32(c), f32(x), f32(y)
            prim.append(f"{key}|{n}|{dt}")
    out["prim_meta"] = np.array(prim)

    # ---- projections, src/patch_llama_fft.py:47-128 -------------------------
    proj = []
    for i, (B, dt) in enumerate([(8, "float32"), (32, "float32"), (64, "float32"), (64, "float16"), (128, "bfloat16")]):
        g = torch.Generator().manual_seed(300 + i)
        W = torch.randn(B, B, generator=g).to(DT[dt])
        key = f"proj{i}"
        out[f"{key}/W"] = f32(W)
        out[f"{key}/loss_aware"] = f32(plf.dense_block_to_circulant_column_loss_aware(W))
        out[f"{key}/diag"] = f32(plf.dense_block_to_circulant_column(W, convention="diag"))
        out[f"{key}/diag_inv"] = f32(plf.dense_block_to_circulant_column(W, convention="diag_inv"))
        proj.append(f"{key}|{B}|{dt}")
    # the reference's one known-answer vector (src/validate_implementation.py:58)
    c_true = torch.arange(1.0, 9.0)
    idx = torch.arange(8)
    C = c_true[(idx[:, None] - idx[None, :]) % 8]
    out["known/c_true"] = f32(c_true)
    out["known/C"] = f32(C)
    out["known/c_recovered"] = f32(plf.dense_block_to_circulant_column_loss_aware(C))
    out["known/circ_from_first_col"] = f32(plf.circulant_from_first_col(c_true))
    out["proj_meta"] = np.array(proj)

    # ---- from_linear, src/patch_llama_fft.py:180-224 ------------------------
    fl = []
    for i, (inf, outf, B, dt, has_bias) in enumerate([(256, 128, 64, "float32", True), (256, 256, 128, "float16", False),
                                                      (512, 256, 256, "bfloat16", True)]):
        torch.manual_seed(400 + i)
        lin = torch.nn.Linear(inf, outf, bias=has_bias)
        lin.weight.data.uniform_(-1, 1)
        lin = lin.to(DT[dt])
        with contextlib.redirect_stdout(io.StringIO()):
            bc = plf.BlockCirculantLinear.from_linear(lin, block_size=B)
        key = f"fl{i}"
        out[f"{key}/W"] = f32(lin.weight)
        out[f"{key}/c"] = f32(bc.c)
        if has_bias:
            out[f"{key}/lin_bias"] = f32(lin.bias)
            out[f"{key}/bc_bias"] = f32(bc.bias)
        fl.append(f"{key}|{inf}|{outf}|{B}|{dt}|{int(has_bias)}|{str(bc.c.dtype).replace('torch.', '')}")
    out["from_linear_meta"] = np.array(fl)

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)

    # ---- a checkpoint written by the reference's save_bc_params -------------
    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = torch.nn.ModuleList()
            for _ in range(2):
                blk = torch.nn.Module()
                blk.gate_proj = plf.BlockCirculantLinear(128, 256, block_size=64, bias=False)
                blk.down_proj = plf.BlockCirculantLinear(256, 128, block_size=64, bias=True)
                blk.dense = torch.nn.Linear(8, 8)
                self.layers.append(blk)

    torch.manual_seed(55)
    tiny = Tiny()
    with contextlib.redirect_stdout(io.StringIO()):
        plf.save_bc_params(tiny, os.path.join(HERE, "reference_bc_params.pt"))
    n_bytes = sum(os.path.getsize(os.path.join(HERE, f)) for f in ("reference_vectors.npz", "reference_bc_params.pt"))
    print(f"wrote {len(out)} arrays, {n_bytes / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
