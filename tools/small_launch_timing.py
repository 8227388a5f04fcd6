#!/usr/bin/env python3
"""Small grouped launches (one decoder layer = 7 projections, Llama-2-7B shapes, B=128, 1 token): eager back-to-back and CUDA-graph
replay time per launch.  With BCL_LIB_PATH pointing at a -DBCL_WS_TIMING build the kernel prints its per-warp wait breakdown."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_fft_b200 as bc
from llama_fft_b200 import _cabi
DEV = torch.device("cuda:0")
B = 128
LAYER = [(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)]
sms = torch.cuda.get_device_properties(DEV).multi_processor_count
g = torch.Generator(device=DEV).manual_seed(0)
groups = []
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    specs = []
    for (p, q) in LAYER:
        c = torch.randn(p, q, B, device=DEV, generator=g) * 0.01
        spec = torch.empty(p, q, B, device=DEV, dtype=torch.float32)
        _cabi.check(_cabi.lib().bcl_precompute_spectrum(c.data_ptr(), p, q, B, spec.data_ptr(), _cabi.F32, _cabi.LAYOUT_STREAM, _cabi.stream_ptr(DEV)), "spec")
        specs.append(spec)
    grp = bc.GroupedBCForward.from_packed(LAYER, specs, 1, torch.float16, DEV, num_ctas=sms)
    grp.x_all.copy_(torch.randn(grp.x_all.numel(), device=DEV, generator=g).to(torch.float16))
    groups.append(grp)
torch.cuda.synchronize()
if os.environ.get("ONE_SHOT"):
    groups[0].run(); torch.cuda.synchronize(); groups[1].run(); torch.cuda.synchronize(); sys.exit(0)
def ev(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
def all_layers():
    for grp in groups: grp.run()
eager = ev(all_layers, 20) / len(groups)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    all_layers(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        all_layers()
    graph = ev(gr.replay, 20) / len(groups)
print(json.dumps({"launches": len(groups), "bytes_per_launch_MB": round(sum(p * q * B * 4 for p, q in LAYER) / 1e6, 2), "eager_us_per_launch": round(eager, 2), "graph_us_per_launch": round(graph, 2)}))
