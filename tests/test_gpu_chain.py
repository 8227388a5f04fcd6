"""GPU tests (-m gpu) of the chain kernel (llama-fft_b200/csrc/bcl_stream_chain.cu, ``bcl_forward_chain``): dependent BC linears of one
token in ONE launch with grid barriers in between, the SWIGLU epilogue and the whole-MLP fusion built on it.
Oracle: oracle/bc_oracle.py (the reference's forward, src/patch_llama_fft.py:226-275) composed the way HF LlamaMLP.forward composes the
three modules the reference swaps (src/patch_llama_fft.py:379): down_proj(act_fn(gate_proj(x)) * up_proj(x))."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

import llama_fft_b200 as bc
from llama_fft_b200 import _cabi
from llama_fft_b200.patch_llama_fft import bc_forward
from oracle import bc_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def spectrum(c, sdt):
    p, q, B = c.shape
    cd = c.to(DEV).contiguous()
    spec = torch.empty(p, q, B, device=DEV, dtype=sdt)
    _cabi.check(_cabi.lib().bcl_precompute_spectrum(cd.data_ptr(), p, q, B, spec.data_ptr(), _cabi.DTYPE[sdt], _cabi.LAYOUT_STREAM,
                                                    _cabi.stream_ptr(torch.device(DEV))), "bcl_precompute_spectrum")
    return spec


def assert_close(y, y_ref, tol, io_dtype, ulps=2):
    yd, rd = y.double().cpu(), y_ref.double().cpu()
    assert yd.shape == rd.shape
    scale = max(rd.abs().max().item(), 1e-30)
    bound = tol * scale + (ulps * G.ulp(io_dtype) * rd.abs() if io_dtype != torch.float32 else 0.0)
    bad = (yd - rd).abs() > bound
    assert not bool(torch.as_tensor(bad).any()), f"max rel err {(yd - rd).abs().max().item() / scale:.3e} (tol {tol})"


SHAPES = [(4096, 4096, 128), (4096, 11008, 128), (11008, 4096, 128), (512, 1024, 64), (1024, 512, 256), (8192, 1024, 256)]


@pytest.mark.parametrize("inf,outf,B", SHAPES)
@pytest.mark.parametrize("dtype,sdt,tol", [(torch.float16, torch.float32, 1e-4), (torch.float32, torch.float32, 1e-4), (torch.bfloat16, torch.bfloat16, 1e-2)])
def test_single_step_matches_oracle_and_the_per_module_kernel(inf, outf, B, dtype, sdt, tol):
    g = torch.Generator().manual_seed(inf + outf + B)
    c = torch.randn(outf // B, inf // B, B, generator=g) * 0.02
    x = torch.randn(1, inf, generator=g).to(dtype)
    spec = spectrum(c, sdt)
    xd = x.to(DEV)
    y = torch.zeros(1, outf, device=DEV, dtype=dtype)
    bc.BCChain([[(xd, spec, y, _cabi.EPILOGUE_LINEAR)]]).run()
    torch.cuda.synchronize()
    assert_close(y, O.bc_forward_torch(c, x), tol, dtype)
    y_mod = bc_forward(xd, spec, None, outf // B, inf // B, B)           # stream_single_kernel: same split-q order, same FFT core
    assert_close(y, y_mod, 1e-6, dtype, ulps=1)


def test_single_step_with_bias_promotion():
    g = torch.Generator().manual_seed(5)
    c = torch.randn(8, 8, 128, generator=g) * 0.05
    b = torch.randn(1024, generator=g) * 0.1
    x = torch.randn(1, 1024, generator=g).half()
    spec = spectrum(c, torch.float32)
    y = torch.zeros(1, 1024, device=DEV, dtype=torch.float32)              # fp16 x + fp32 bias -> fp32 y (reference :266-271)
    bc.BCChain([[(x.to(DEV), spec, y, _cabi.EPILOGUE_LINEAR, b.to(DEV))]]).run()
    torch.cuda.synchronize()
    y_ref = O.bc_forward_torch(c, x, b)
    assert y_ref.dtype == torch.float32
    # y = round16(y32) + b: the fp16 rounding acted on y32 = y_ref - b, so a flipped rounding moves y by one fp16 ulp of THAT magnitude
    err = (y.double().cpu() - y_ref.double()).abs()
    bound = 1e-4 * y_ref.abs().max().item() + 2 * G.ulp(torch.float16) * (y_ref.double() - b.double()).abs()
    assert bool((err <= bound).all()), f"max err {err.max().item():.3e}"


def test_one_layer_of_independent_projections_in_one_step():
    """q/k/v/o + gate/up + down of a 7B layer as ONE single-step launch (the small-launch replacement of the grouped plan)."""
    g = torch.Generator().manual_seed(9)
    shapes = [(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)]
    cs = [torch.randn(p, q, 128, generator=g) * 0.02 for (p, q) in shapes]
    xs = [torch.randn(1, q * 128, generator=g).half() for (p, q) in shapes]
    specs = [spectrum(c, torch.float32) for c in cs]
    ys = [torch.zeros(1, p * 128, device=DEV, dtype=torch.float16) for (p, q) in shapes]
    chain = bc.BCChain([[(x.to(DEV), s, y, _cabi.EPILOGUE_LINEAR) for x, s, y in zip(xs, specs, ys)]])
    chain.run()
    torch.cuda.synchronize()
    for c, x, y in zip(cs, xs, ys):
        assert_close(y, O.bc_forward_torch(c, x), 1e-4, torch.float16)


def mlp_reference(cg, cu, cd, x):
    """HF LlamaMLP.forward over the reference's modules, on the CPU oracle: every module output is rounded to x.dtype (:266)."""
    gate, up = O.bc_forward_torch(cg, x), O.bc_forward_torch(cu, x)
    h = F.silu(gate) * up
    return h, O.bc_forward_torch(cd, h)


@pytest.mark.parametrize("hidden,inter,B", [(4096, 11008, 128), (512, 1536, 64), (1024, 2048, 256)])
@pytest.mark.parametrize("dtype,sdt,tol", [(torch.float16, torch.float32, 2e-4), (torch.float32, torch.float32, 2e-4), (torch.bfloat16, torch.bfloat16, 2e-2)])
def test_fused_mlp_chain_matches_the_composed_oracle(hidden, inter, B, dtype, sdt, tol):
    g = torch.Generator().manual_seed(hidden + inter + B)
    cg = torch.randn(inter // B, hidden // B, B, generator=g) * 0.03
    cu = torch.randn(inter // B, hidden // B, B, generator=g) * 0.03
    cd = torch.randn(hidden // B, inter // B, B, generator=g) * 0.03
    x = torch.randn(1, hidden, generator=g).to(dtype)
    h_ref, y_ref = mlp_reference(cg, cu, cd, x)
    spec_gu = spectrum(torch.cat([cg, cu], 0), sdt)
    spec_d = spectrum(cd, sdt)
    h = torch.zeros(1, inter, device=DEV, dtype=dtype)
    y = torch.zeros(1, hidden, device=DEV, dtype=dtype)
    chain = bc.BCChain([[(x.to(DEV), spec_gu, h, _cabi.EPILOGUE_SWIGLU)], [(h, spec_d, y, _cabi.EPILOGUE_LINEAR)]])
    chain.run()
    torch.cuda.synchronize()
    # h: products of two rounded values: a 1-ulp flip of gate or up moves h by ~2 ulp; absolute floor from the 1e-4 / 1e-2 class
    assert_close(h, h_ref, tol, dtype, ulps=4)
    assert_close(y, y_ref, tol, dtype, ulps=4)
    # the barrier counter is shared and never reset: many launches back to back, then under CUDA-graph replay, stay bit-identical
    y0, h0 = y.clone(), h.clone()
    for _ in range(40):
        chain.run()
    torch.cuda.synchronize()
    assert torch.equal(y, y0) and torch.equal(h, h0)
    # spectrum hand-off: step 0 also emits rfft of the rounded h, step 1 consumes it instead of transforming h again: same bits
    hf = torch.zeros(1, inter, device=DEV, dtype=torch.float32)
    h2, y2 = torch.zeros_like(h), torch.zeros_like(y)
    bc.BCChain([[(x.to(DEV), spec_gu, h2, _cabi.EPILOGUE_SWIGLU, None, hf)], [(hf, spec_d, y2, _cabi.EPILOGUE_LINEAR, None, None, True)]]).run()
    torch.cuda.synchronize()
    assert torch.equal(h2, h0) and torch.equal(y2, y0)
    hf_ref = torch.fft.rfft(h0.float().view(inter // B, B), dim=-1)
    packed = torch.view_as_real(hf_ref[:, :B // 2]).reshape(inter // B, B).clone()
    packed[:, 1] = hf_ref[:, B // 2].real
    assert (hf.view(inter // B, B) - packed).abs().max().item() <= 1e-5 * packed.abs().max().item()
    s = torch.cuda.Stream(DEV)
    with torch.cuda.stream(s):
        chain.run()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            chain.run()
            chain.run()
        y.zero_(); h.zero_()
        for _ in range(5):
            graph.replay()
        torch.cuda.synchronize()
    assert torch.equal(y, y0) and torch.equal(h, h0)


def test_long_dependent_chain_equals_step_by_step_launches():
    """Eight dependent steps (each reads the previous step's output, buffers ping-pong so an address is re-read after it was rewritten by
    other SMs): one launch == eight single-step launches, bit for bit; mixed grids share the one barrier counter."""
    g = torch.Generator().manual_seed(77)
    B, p = 128, 16
    cs = [torch.randn(p, p, B, generator=g) * (1.0 / (p * B) ** 0.5) for _ in range(8)]       # unit gain per step: no fp16 overflow
    specs = [spectrum(c, torch.float32) for c in cs]
    x = torch.randn(1, p * B, generator=g).half().to(DEV)
    buf = [torch.zeros(1, p * B, device=DEV, dtype=torch.float16) for _ in range(2)]
    steps, cur = [], x
    for k, s in enumerate(specs):
        steps.append([(cur, s, buf[k & 1], _cabi.EPILOGUE_LINEAR)])
        cur = buf[k & 1]
    bc.BCChain(steps).run()
    torch.cuda.synchronize()
    got = buf[1].clone()
    cur = x
    for k, s in enumerate(specs):
        bc.BCChain([[(cur, s, buf[k & 1], _cabi.EPILOGUE_LINEAR)]]).run()
        cur = buf[k & 1]
    torch.cuda.synchronize()
    assert torch.equal(got, buf[1])
    ref = x.cpu()
    for c in cs:
        ref = O.bc_forward_torch(c, ref)
    assert_close(got, ref, 5e-3, torch.float16, ulps=8)                      # eight fp16 roundings deep


def test_chain_argument_validation():
    L = _cabi.lib()
    spec = torch.zeros(4, 4, 128, device=DEV)
    x2 = torch.zeros(2, 512, device=DEV)
    y2 = torch.zeros(2, 512, device=DEV)
    probs = (_cabi.Problem * 1)(_cabi.Problem(x2.data_ptr(), spec.data_ptr(), None, y2.data_ptr(), 2, 512, 512, 4, 4, 0, 0, 0, 0))
    first = (ctypes.c_int * 2)(0, 1)
    st = _cabi.stream_ptr(torch.device(DEV))
    assert L.bcl_forward_chain(probs, first, 1, 128, _cabi.F32, None, st) == -2          # N != 1
    probs[0].N = 1
    assert L.bcl_forward_chain(probs, first, 1, 128, _cabi.F32, None, st) == 0
    spec3 = torch.zeros(3, 4, 128, device=DEV)
    probs[0] = _cabi.Problem(x2.data_ptr(), spec3.data_ptr(), None, y2.data_ptr(), 1, 512, 512, 3, 4, 0, 0, 0, _cabi.EPILOGUE_SWIGLU)
    assert L.bcl_forward_chain(probs, first, 1, 128, _cabi.F32, None, st) == -2          # odd number of stacked rows
    first2 = (ctypes.c_int * 3)(0, 1, 1)
    assert L.bcl_forward_chain(probs, first2, 2, 128, _cabi.F32, None, st) == -1         # several steps need a counter / empty step
    torch.cuda.synchronize()


def test_fuse_mlp_in_a_tiny_llama_decode_step():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      vocab_size=320, max_position_embeddings=512)
    torch.manual_seed(0)
    model = LlamaForCausalLM(cfg).eval()
    bc.patch_llama_fft(model, block_size=64)
    model = model.to(DEV)
    ids = torch.randint(0, 320, (1, 1), device=DEV)
    with torch.no_grad():
        ref1 = model(input_ids=ids).logits
        assert bc.fuse_mlp(model) == 2
        fused = [m._bcl_fused_mlp for m in model.modules() if hasattr(m, "_bcl_fused_mlp")]
        got1 = model(input_ids=ids).logits
        assert [f.fused_calls for f in fused] == [1, 1]                     # one token: the chain kernel served both MLPs
        ids9 = torch.randint(0, 320, (1, 9), device=DEV)
        a9 = model(input_ids=ids9).logits                                    # nine tokens: the original forward
        assert [f.fused_calls for f in fused] == [1, 1]
        out = model.generate(input_ids=ids9, max_new_tokens=4, do_sample=False, use_cache=True, pad_token_id=0)
        assert [f.fused_calls for f in fused] == [4, 4] and out.shape == (1, 13)   # prefill unfused, 3 + ... single-token decode steps fused
        assert bc.unfuse_mlp(model) == 2
        b9 = model(input_ids=ids9).logits
        back1 = model(input_ids=ids).logits
    assert torch.equal(a9, b9) and torch.equal(back1, ref1)
    assert (got1 - ref1).abs().max().item() <= 2e-4 * ref1.abs().max().item()


# ---------------------------------------------------------------------------------------------------------
# whole-MLP fusion at prefill: bcl_forward_swiglu (gate | up stacked, SWIGLU epilogue in the tensor-core inverse stage)
# ---------------------------------------------------------------------------------------------------------
def _gemm_spectrum(c, sd=_cabi.BF16):
    p, q, B = c.shape
    L = _cabi.lib()
    cd = c.to(DEV).contiguous()
    spec = torch.empty(L.bcl_spectrum_bytes(p, q, B, sd, _cabi.LAYOUT_GEMM), device=DEV, dtype=torch.uint8)
    _cabi.check(L.bcl_precompute_spectrum(cd.data_ptr(), p, q, B, spec.data_ptr(), sd, _cabi.LAYOUT_GEMM, _cabi.stream_ptr(torch.device(DEV))),
                "bcl_precompute_spectrum")
    return spec


def _forward_swiglu(x, spec, p, q, B, sd=_cabi.BF16, ws_bytes=None):
    L = _cabi.lib()
    N = x.shape[0]
    h = torch.full((N, p // 2 * B), float("nan"), device=DEV, dtype=x.dtype)
    ws = torch.zeros(ws_bytes or L.bcl_workspace_bytes(N, p, q, B, sd, _cabi.LAYOUT_GEMM), device=DEV, dtype=torch.uint8)
    rc = L.bcl_forward_swiglu(x.data_ptr(), _cabi.DTYPE[x.dtype], N, x.stride(0), spec.data_ptr(), sd, p, q, B,
                              h.data_ptr(), _cabi.DTYPE[h.dtype], h.stride(0), ws.data_ptr(), ws.numel(), _cabi.stream_ptr(torch.device(DEV)))
    _cabi.check(rc, "bcl_forward_swiglu")
    torch.cuda.synchronize()
    assert int(ws[:8].view(torch.int32)[0].item()) == 0 and int(ws[:8].view(torch.int32)[1].item()) == 0      # no pipeline stall
    return h


@pytest.mark.parametrize("hidden,inter,B,N,dtype", [(4096, 11008, 128, 1500, torch.float16), (4096, 11008, 128, 640, torch.bfloat16),
                                                     (512, 1536, 64, 700, torch.float16), (1024, 2048, 128, 128 * 9 + 1, torch.bfloat16)])
def test_prefill_swiglu_stage_matches_the_unfused_modules_and_the_oracle(hidden, inter, B, N, dtype):
    """h = act_fn(gate_proj(x)) * up_proj(x) in one three-stage forward.  Against (a) the same two projections through the library's own
    tensor-core path followed by torch's silu and multiply -- the stacked forward is bit-identical per member, so only the fp32 silu
    evaluation (__expf / __fdividef vs torch) can flip a last bit of the rounded silu: h within 2 ulp everywhere, bit-equal almost
    everywhere -- and (b) the reference's modules composed on the CPU oracle, bf16-spectrum class."""
    g = torch.Generator().manual_seed(hidden + inter + B + N)
    cg = torch.randn(inter // B, hidden // B, B, generator=g) * 0.03
    cu = torch.randn(inter // B, hidden // B, B, generator=g) * 0.03
    x = torch.randn(N, hidden, generator=g).to(dtype)
    xd = x.to(DEV)
    p, q = 2 * (inter // B), hidden // B
    h = _forward_swiglu(xd, _gemm_spectrum(torch.cat([cg, cu], 0)), p, q, B)
    assert not torch.isnan(h.float()).any()
    gate = bc_forward(xd, _gemm_spectrum(cg), None, p // 2, q, B, layout=_cabi.LAYOUT_GEMM, spectrum_dtype=torch.bfloat16)
    up = bc_forward(xd, _gemm_spectrum(cu), None, p // 2, q, B, layout=_cabi.LAYOUT_GEMM, spectrum_dtype=torch.bfloat16)
    h_unfused = F.silu(gate) * up
    diff = (h.double() - h_unfused.double()).abs()
    assert bool((diff <= 2 * G.ulp(dtype) * h_unfused.double().abs() + 1e-30).all()), f"max diff {diff.max().item():.3e}"
    assert (h == h_unfused).float().mean().item() >= 0.98
    rows = torch.tensor([0, 1, 127, 128, N // 2, N - 1])
    gate_ref, up_ref = O.bc_forward_torch(cg, x[rows].float()), O.bc_forward_torch(cu, x[rows].float())
    h_ref = F.silu(gate_ref) * up_ref
    assert_close(h[rows], h_ref, 2e-2, dtype, ulps=4)
    h2 = _forward_swiglu(xd, _gemm_spectrum(torch.cat([cg, cu], 0)), p, q, B)
    assert torch.equal(h, h2)


@pytest.mark.parametrize("hidden,inter,B,N,dtype", [(4096, 11008, 128, 1500, torch.float16), (4096, 11008, 128, 640, torch.bfloat16),
                                                     (512, 1536, 64, 700, torch.float16), (1024, 2048, 128, 128 * 9 + 1, torch.bfloat16)])
def test_prefill_swiglu_stage_with_fp32_spectra(hidden, inter, B, N, dtype):
    """The same fusion for the default precision (fp32 spectra = (hi, lo) bf16 pairs, split stages; dft_inv_swiglu_tc_split_kernel):
    (a) within 2 ulp of the two unfused projections through the library's own split tensor-core path followed by torch's silu and
    multiply, (b) against the reference's modules composed on the CPU oracle in the fp32-spectrum class (1e-4 on gate and up; h is a
    product of two rounded 16-bit values, hence the ulps), (c) forced chunking (a workspace of two 128-token tiles) changes nothing."""
    g = torch.Generator().manual_seed(hidden + inter + B + N + 1)
    cg = torch.randn(inter // B, hidden // B, B, generator=g) * 0.03
    cu = torch.randn(inter // B, hidden // B, B, generator=g) * 0.03
    x = torch.randn(N, hidden, generator=g).to(dtype)
    xd = x.to(DEV)
    p, q = 2 * (inter // B), hidden // B
    spec = _gemm_spectrum(torch.cat([cg, cu], 0), _cabi.F32)
    h = _forward_swiglu(xd, spec, p, q, B, _cabi.F32)
    assert not torch.isnan(h.float()).any()
    gate = bc_forward(xd, _gemm_spectrum(cg, _cabi.F32), None, p // 2, q, B, layout=_cabi.LAYOUT_GEMM, spectrum_dtype=torch.float32)
    up = bc_forward(xd, _gemm_spectrum(cu, _cabi.F32), None, p // 2, q, B, layout=_cabi.LAYOUT_GEMM, spectrum_dtype=torch.float32)
    h_unfused = F.silu(gate) * up
    diff = (h.double() - h_unfused.double()).abs()
    assert bool((diff <= 2 * G.ulp(dtype) * h_unfused.double().abs() + 1e-30).all()), f"max diff {diff.max().item():.3e}"
    assert (h == h_unfused).float().mean().item() >= 0.98
    rows = torch.tensor([0, 1, 127, 128, N // 2, N - 1])
    gate_ref, up_ref = O.bc_forward_torch(cg, x[rows].float()), O.bc_forward_torch(cu, x[rows].float())
    h_ref = F.silu(gate_ref) * up_ref
    assert_close(h[rows], h_ref, 5e-4, dtype, ulps=4)
    L = _cabi.lib()
    small = L.bcl_workspace_bytes(256, p, q, B, _cabi.F32, _cabi.LAYOUT_GEMM)
    if N > 256:
        assert torch.equal(h, _forward_swiglu(xd, spec, p, q, B, _cabi.F32, ws_bytes=small))


def test_forward_swiglu_argument_validation():
    L = _cabi.lib()
    x = torch.zeros(256, 512, device=DEV, dtype=torch.float16)
    c = torch.zeros(6, 4, 128)
    spec = _gemm_spectrum(c)
    h = torch.zeros(256, 384, device=DEV, dtype=torch.float16)
    ws = torch.zeros(L.bcl_workspace_bytes(256, 6, 4, 128, _cabi.BF16, _cabi.LAYOUT_GEMM), device=DEV, dtype=torch.uint8)
    st = _cabi.stream_ptr(torch.device(DEV))
    args = lambda p=6, sd=_cabi.BF16, B=128, xt=x: (xt.data_ptr(), _cabi.DTYPE[xt.dtype], 256, xt.stride(0), spec.data_ptr(), sd, p, 4, B,
                                                      h.data_ptr(), _cabi.DTYPE[xt.dtype], h.stride(0), ws.data_ptr(), ws.numel(), st)
    assert L.bcl_forward_swiglu(*args()) == 0
    assert L.bcl_forward_swiglu(*args(p=5)) != 0                               # odd number of block-rows
    assert L.bcl_forward_swiglu(*args(sd=_cabi.F16)) != 0                      # spectra are bf16 or f32 (split pairs)
    assert L.bcl_forward_swiglu(*args(B=256)) != 0                             # B <= 128
    assert L.bcl_forward_swiglu(*args(xt=x.float()[:, :512].contiguous())) != 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("sdt,tol", [(torch.bfloat16, 1e-2), (torch.float32, 2e-3)])
def test_fuse_mlp_prefill_path_in_a_tiny_llama(monkeypatch, sdt, tol):
    """fuse_mlp serves many-token eval calls with bcl_forward_swiglu + down_proj when the stacked gate | up layer has enough work for
    the tensor-core path: logits agree with the unfused model to the spectrum class (fp16 model: the fp32-spectrum arm is bounded by
    the 16-bit activations between the layers), and the counter shows the path really ran."""
    import importlib
    from transformers import LlamaConfig, LlamaForCausalLM
    P = importlib.import_module("llama_fft_b200.patch_llama_fft")
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      vocab_size=320, max_position_embeddings=512)
    torch.manual_seed(0)
    model = LlamaForCausalLM(cfg).eval()
    bc.patch_llama_fft(model, block_size=64, spectrum_dtype=sdt)
    model = model.to(DEV).half()
    ids = torch.randint(0, 320, (4, 96), device=DEV)
    monkeypatch.setattr(P, "PREFILL_MIN_TOKENS", 256)
    monkeypatch.setattr(P, "PREFILL_MIN_TOKENS_F32", 256)
    with torch.no_grad():
        ref = model(input_ids=ids).logits.float()
        assert bc.fuse_mlp(model) == 2
        fused = [m._bcl_fused_mlp for m in model.modules() if hasattr(m, "_bcl_fused_mlp")]
        got = model(input_ids=ids).logits.float()
        assert [f.fused_prefill_calls for f in fused] == [1, 1] and [f.fused_calls for f in fused] == [0, 0]
        bc.unfuse_mlp(model)
    assert P.prefill_status(DEV) == 0
    assert (got - ref).abs().max().item() <= tol * ref.abs().max().item()


def test_prefill_swiglu_stage_at_full_size():
    """BASELINE configs[2] size (16384 tokens, Llama-2-7B MLP, B = 128): the fused SWIGLU stage against the unfused modules on 256 random
    rows plus the tile edges, determinism, and linearity in x of the PRE-activation path via the down projection is not available here
    (SWIGLU is not linear): instead gate == 0 rows must give exactly h == 0 (silu(0) * up)."""
    hidden, inter, B, N, dtype = 4096, 11008, 128, 16384, torch.float16
    g = torch.Generator().manual_seed(5)
    cg = torch.randn(inter // B, hidden // B, B, generator=g) * 0.03
    cu = torch.randn(inter // B, hidden // B, B, generator=g) * 0.03
    x = torch.randn(N, hidden, generator=g).to(dtype)
    x[777] = 0                                                      # a zero token: gate = up = 0 -> h = 0 exactly
    xd = x.to(DEV)
    p, q = 2 * (inter // B), hidden // B
    spec = _gemm_spectrum(torch.cat([cg, cu], 0))
    h = _forward_swiglu(xd, spec, p, q, B)
    assert torch.equal(h, _forward_swiglu(xd, spec, p, q, B))
    assert not torch.isnan(h.float()).any() and bool((h[777] == 0).all())
    rows = torch.unique(torch.cat([torch.tensor([0, 127, 128, 8191, 8192, N - 1]), torch.randint(0, N, (256,), generator=g)]))
    xr = xd[rows].contiguous()
    gate = bc_forward(xr, _gemm_spectrum(cg), None, p // 2, q, B, layout=_cabi.LAYOUT_GEMM, spectrum_dtype=torch.bfloat16)
    up = bc_forward(xr, _gemm_spectrum(cu), None, p // 2, q, B, layout=_cabi.LAYOUT_GEMM, spectrum_dtype=torch.bfloat16)
    h_unfused = F.silu(gate) * up
    diff = (h[rows].double() - h_unfused.double()).abs()
    assert bool((diff <= 2 * G.ulp(dtype) * h_unfused.double().abs() + 1e-30).all()), f"max diff {diff.max().item():.3e}"
    sub = rows[:24]
    h_ref = F.silu(O.bc_forward_torch(cg, x[sub].float())) * O.bc_forward_torch(cu, x[sub].float())
    assert_close(h[sub], h_ref, 2e-2, dtype, ulps=4)
