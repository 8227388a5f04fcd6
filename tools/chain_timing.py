#!/usr/bin/env python3
"""Decode latency at the launch granularities a real decoder has (B200, CUDA-graph replay, fp16 activations, fp32 spectra, B = 128,
Llama-2-7B shapes, 32 layers of distinct weights): per-module kernel vs the chain kernel, the fused MLP, the dependent layer chain.
Also dumps the chain kernel's clock stamps (CTA 0) for the fused MLP.   python tools/chain_timing.py [--bf16]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DEV = torch.device("cuda:0")
dbg = torch.zeros(8 * 8, device=DEV, dtype=torch.int64)
os.environ["BCL_CHAIN_DBG_PTR"] = str(dbg.data_ptr())              # read once by the library at its first chain launch

import llama_fft_b200 as bc                                        # noqa: E402
from llama_fft_b200 import _cabi                                   # noqa: E402
from llama_fft_b200.patch_llama_fft import bc_forward             # noqa: E402

B = 128
SDT = torch.bfloat16 if "--bf16" in sys.argv else torch.float32
LAYERS = 32
L = _cabi.lib()


def spectrum(p, q):
    c = torch.randn(p, q, B, device=DEV) * 0.01
    spec = torch.empty(p, q, B, device=DEV, dtype=SDT)
    _cabi.check(L.bcl_precompute_spectrum(c.data_ptr(), p, q, B, spec.data_ptr(), _cabi.DTYPE[SDT], _cabi.LAYOUT_STREAM, _cabi.stream_ptr(DEV)), "spec")
    return spec


def time_graph(fn, reps=20):
    s = torch.cuda.Stream(DEV)
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            fn()
        for _ in range(3):
            graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            graph.replay()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3          # us per replay


torch.manual_seed(0)
H, I = 4096, 11008
layers = []
for _ in range(LAYERS):
    layers.append(dict(qkv=spectrum(96, 32), o=spectrum(32, 32), gu=spectrum(172, 32), down=spectrum(32, 86)))
x0 = torch.randn(1, H, device=DEV).half()
f16 = dict(device=DEV, dtype=torch.float16)
out = {"spectrum_dtype": str(SDT)}

if "--stamps-only" in sys.argv:
    hb, yb = torch.empty(1, I, **dict(device=DEV, dtype=torch.float16)), torch.empty(1, H, **dict(device=DEV, dtype=torch.float16))
    hfb = torch.empty(1, I, device=DEV, dtype=torch.float32)
    ch = bc.BCChain([[(x0, layers[0]["gu"], hb, _cabi.EPILOGUE_SWIGLU, None, hfb)], [(hfb, layers[0]["down"], yb, 0, None, None, True)]])
    for _ in range(3):
        ch.run()
    torch.cuda.synchronize()
    dbg.zero_()
    ch.run()
    torch.cuda.synchronize()
    st = dbg.view(8, 8)[:2].tolist()
    t0 = st[0][0]
    fin = dbg[16:20].tolist()
    print(json.dumps({"legend": "[loop top, barrier passed, x ready, parked, finished, arrived, x loads landed (warp 1), first rfft done (warp 1)]",
                      "step0": [v - t0 if v else None for v in st[0]], "step1": [v - t0 if v else None for v in st[1]],
                      "finish_step0 [sums done, irfft done, epilogue done, emit done]": [v - t0 for v in fin]}))
    sys.exit(0)

# ---- (a) one small projection per launch, 32 distinct weights back to back (independent) ------------------------------------------
ys = [torch.empty(1, H, **f16) for _ in range(LAYERS)]
def per_module_o():
    for l, y in zip(layers, ys):
        bc_forward(x0, l["o"], None, 32, 32, B, out=y)
chains_o = [bc.BCChain([[(x0, l["o"], y, 0)]]) for l, y in zip(layers, ys)]
out["o_proj_us_per_launch"] = {"stream_single_kernel": time_graph(per_module_o) / LAYERS, "chain_kernel_1_step": time_graph(lambda: [c.run() for c in chains_o]) / LAYERS}
yg = [torch.empty(1, 2 * I, **f16) for _ in range(LAYERS)]
def per_module_gu():
    for l, y in zip(layers, yg):
        bc_forward(x0, l["gu"], None, 172, 32, B, out=y)
chains_gu = [bc.BCChain([[(x0, l["gu"], y, 0)]]) for l, y in zip(layers, yg)]
out["gate_up_stacked_us_per_launch"] = {"stream_single_kernel": time_graph(per_module_gu) / LAYERS, "chain_kernel_1_step": time_graph(lambda: [c.run() for c in chains_gu]) / LAYERS}

# ---- (b) the MLP: per-module (gate/up stacked launch, silu*mul, down) vs ONE fused launch -----------------------------------------
hbuf = [torch.empty(1, I, **f16) for _ in range(LAYERS)]
ymlp = [torch.empty(1, H, **f16) for _ in range(LAYERS)]
def mlp_per_module():
    for l, g, y in zip(layers, yg, ymlp):
        bc_forward(x0, l["gu"], None, 172, 32, B, out=g)
        h = torch.nn.functional.silu(g[:, :I]) * g[:, I:]
        bc_forward(h, l["down"], None, 32, 86, B, out=y)
mlp_chains = [bc.BCChain([[(x0, l["gu"], h, _cabi.EPILOGUE_SWIGLU)], [(h, l["down"], y, 0)]]) for l, h, y in zip(layers, hbuf, ymlp)]
hfbuf = [torch.empty(1, I, device=DEV, dtype=torch.float32) for _ in range(LAYERS)]
mlp_chains_hf = [bc.BCChain([[(x0, l["gu"], h, _cabi.EPILOGUE_SWIGLU, None, hf)], [(hf, l["down"], y, 0, None, None, True)]]) for l, h, hf, y in zip(layers, hbuf, hfbuf, ymlp)]
mlp_two = [(bc.BCChain([[(x0, l["gu"], h, _cabi.EPILOGUE_SWIGLU)]]), bc.BCChain([[(h, l["down"], y, 0)]])) for l, h, y in zip(layers, hbuf, ymlp)]
out["mlp_us_per_layer"] = {"per_module_3_launches": time_graph(mlp_per_module) / LAYERS, "fused_chain_1_launch": time_graph(lambda: [c.run() for c in mlp_chains]) / LAYERS,
                           "fused_chain_1_launch_spectrum_handoff": time_graph(lambda: [c.run() for c in mlp_chains_hf]) / LAYERS,
                           "chain_kernel_2_launches_swiglu_then_down": time_graph(lambda: [(a.run(), b.run()) for a, b in mlp_two]) / LAYERS}

# ---- (c) 7 independent projections of a layer in one launch ----------------------------------------------------------------------
shapes = [(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)]
grp, chn = [], []
for _ in range(LAYERS):
    specs = [spectrum(p, q) for (p, q) in shapes]
    xs = [torch.randn(1, q * B, device=DEV).half() for (p, q) in shapes]
    g = bc.GroupedBCForward(xs, specs)
    grp.append(g)
    chn.append(bc.BCChain([[(x, s, y, 0) for x, s, y in zip(xs, specs, g.ys)]]))
out["layer_7_independent_us_per_launch"] = {"grouped_ws_kernel": time_graph(lambda: [g.run() for g in grp]) / LAYERS,
                                            "chain_kernel_1_step": time_graph(lambda: [c.run() for c in chn]) / LAYERS}
del grp, chn

# ---- (d) the dependent layer chain: q/k/v -> o -> gate/up -> down, each launch reads the previous one's output ---------------------
bufs = dict(qkv=torch.empty(1, 3 * H, **f16), a=torch.empty(1, H, **f16), gu=torch.empty(1, 2 * I, **f16), h=torch.empty(1, I, **f16), hf=torch.empty(1, I, device=DEV, dtype=torch.float32),
            y=[torch.empty(1, H, **f16) for _ in range(2)])
def chain_per_module():
    cur = x0
    for k, l in enumerate(layers):
        bc_forward(cur, l["qkv"], None, 96, 32, B, out=bufs["qkv"])
        bc_forward(bufs["qkv"][:, :H], l["o"], None, 32, 32, B, out=bufs["a"])
        bc_forward(bufs["a"], l["gu"], None, 172, 32, B, out=bufs["gu"])
        bc_forward(bufs["gu"][:, :I], l["down"], None, 32, 86, B, out=bufs["y"][k & 1])
        cur = bufs["y"][k & 1]
single, fused3, whole = [], [], []
cur = x0
for k, l in enumerate(layers):
    y = bufs["y"][k & 1]
    single += [bc.BCChain([[(cur, l["qkv"], bufs["qkv"], 0)]]), bc.BCChain([[(bufs["qkv"][:, :H], l["o"], bufs["a"], 0)]]),
               bc.BCChain([[(bufs["a"], l["gu"], bufs["gu"], 0)]]), bc.BCChain([[(bufs["gu"][:, :I], l["down"], y, 0)]])]
    fused3 += [bc.BCChain([[(cur, l["qkv"], bufs["qkv"], 0)]]), bc.BCChain([[(bufs["qkv"][:, :H], l["o"], bufs["a"], 0)]]),
               bc.BCChain([[(bufs["a"], l["gu"], bufs["h"], _cabi.EPILOGUE_SWIGLU, None, bufs["hf"])], [(bufs["hf"], l["down"], y, 0, None, None, True)]])]
    whole += [bc.BCChain([[(cur, l["qkv"], bufs["qkv"], 0)], [(bufs["qkv"][:, :H], l["o"], bufs["a"], 0)],
                          [(bufs["a"], l["gu"], bufs["h"], _cabi.EPILOGUE_SWIGLU, None, bufs["hf"])], [(bufs["hf"], l["down"], y, 0, None, None, True)]])]
    cur = y
out["dependent_chain_us_per_layer"] = {
    "per_module_kernel_4_launches": time_graph(chain_per_module, 10) / LAYERS,
    "chain_kernel_4_launches": time_graph(lambda: [c.run() for c in single], 10) / LAYERS,
    "qkv_o_fusedMLP_3_launches": time_graph(lambda: [c.run() for c in fused3], 10) / LAYERS,
    "whole_layer_1_launch_4_steps": time_graph(lambda: [c.run() for c in whole], 10) / LAYERS}

# ---- clock stamps of CTA 0 for one fused MLP launch ---------------------------------------------------------------------------------
dbg.zero_()
mlp_chains_hf[0].run()
torch.cuda.synchronize()
st = dbg.view(8, 8)[:2].tolist()
t0 = st[0][0]
out["fused_mlp_stamps_cycles_cta0"] = {"legend": "per step: [loop top, barrier passed, x spectrum ready, partials parked, finished, arrived]",
                                       "step0": [v - t0 if v else None for v in st[0][:8]], "step1": [v - t0 if v else None for v in st[1][:8]]}
print(json.dumps(out, indent=1))
