"""GPU tests (-m gpu) of shared-input fusion (llama-fft_b200/shared_input.py): q/k/v and gate/up as one BC forward over stacked
block-rows.  Block-row j depends on c[j] only, so every member's slice must be BIT-identical to the member's own forward
(reference forward: src/patch_llama_fft.py:226-275), on the decode kernel, the small-N kernel and the tcgen05 prefill path."""
import pytest
import torch

import llama_fft_b200 as bc
from llama_fft_b200.patch_llama_fft import prefill_status

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _layers(inf, outs, B, sdt, bias, seed):
    torch.manual_seed(seed)
    ls = []
    for o in outs:
        m = bc.BlockCirculantLinear(inf, o, block_size=B, bias=bias, spectrum_dtype=sdt)
        if bias:
            with torch.no_grad():
                m.bias.normal_(0, 0.1)
        ls.append(m.to(DEV).eval())
    return ls


@pytest.mark.parametrize("inf,outs,B,N,dtype,sdt,bias", [
    (4096, (4096, 4096, 4096), 128, 1, torch.float16, torch.float32, False),          # q/k/v decode
    (4096, (11008, 11008), 128, 1, torch.float16, torch.float32, False),              # gate/up decode
    (1024, (1024, 256, 256), 64, 7, torch.float32, torch.float32, True),              # GQA-like widths, bias, small N
    (4096, (4096, 4096, 4096), 128, 384, torch.float16, torch.bfloat16, False),       # q/k/v on the tcgen05 prefill path
    (4096, (11008, 11008), 128, 300, torch.bfloat16, torch.bfloat16, False),          # gate/up prefill, ragged tile
    (512, (512, 1024), 256, 5, torch.bfloat16, torch.bfloat16, True),
])
def test_fused_members_are_bit_identical_to_their_own_forward(inf, outs, B, N, dtype, sdt, bias):
    ls = _layers(inf, outs, B, sdt, bias, seed=inf + N)
    if bias:
        for m in ls:
            m.bias.data = m.bias.data.to(dtype)
    x = torch.randn(2, N, inf, device=DEV).to(dtype) if N > 1 else torch.randn(N, inf, device=DEV).to(dtype)
    with torch.no_grad():
        own = [m(x) for m in ls]
    group = bc.SharedInputGroup(ls)
    calls = []
    inner = group.forward_all
    group.forward_all = lambda xf: (calls.append(1), inner(xf))[1]
    for m in ls:
        m._shared_group = group
    with torch.no_grad():
        fused = [m(x) for m in ls]
    torch.cuda.synchronize()
    assert prefill_status(DEV) == 0
    assert len(calls) == 1                                           # one fused launch served every member
    for a, b in zip(own, fused):
        assert a.shape == b.shape and a.dtype == b.dtype
        assert torch.equal(a, b)
    # a different activation, or the same member twice, recomputes; grad mode bypasses the group
    with torch.no_grad():
        y0 = ls[0](x)
        y0b = ls[0](x)
        y1 = ls[1](x * 2)
    assert len(calls) == 4 and torch.equal(y0, y0b) and not torch.equal(y1, fused[1])
    xg = x.float().requires_grad_(True)
    ls[0](xg).sum().backward()
    assert len(calls) == 4 and xg.grad is not None


def test_fused_spectrum_follows_member_updates():
    ls = _layers(512, (512, 512), 64, torch.float32, False, seed=3)
    group = bc.SharedInputGroup(ls)
    for m in ls:
        m._shared_group = group
    x = torch.randn(3, 512, device=DEV)
    with torch.no_grad():
        a = [m(x).clone() for m in ls]
        ls[1].c.mul_(2.0)                                            # an optimizer step / load_bc_params bumps the version
        b = [m(x).clone() for m in ls]
    assert torch.equal(a[0], b[0])
    assert torch.allclose(b[1], 2 * a[1], rtol=1e-5, atol=1e-6)


def test_tiny_llama_logits_unchanged_by_fusion():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      vocab_size=320, max_position_embeddings=512)
    torch.manual_seed(0)
    model = LlamaForCausalLM(cfg).eval()
    bc.patch_llama_fft(model, block_size=64)
    model = model.to(DEV)
    ids = torch.randint(0, 320, (2, 9), device=DEV)
    assert not any(m.training for m in model.modules())             # from_linear keeps the replaced module's mode (model was in eval)
    with torch.no_grad():
        ref = model(input_ids=ids).logits
        assert bc.fuse_shared_inputs(model) == 4                     # GQA: k/v are narrower than q, still one group
        groups = list({id(m._shared_group): m._shared_group for m in model.modules() if getattr(m, "_shared_group", None) is not None}.values())
        got = model(input_ids=ids).logits
        assert [g.fused_calls for g in groups] == [1, 1, 1, 1]       # one fused launch per group per forward: the group path really ran
        ids_long = torch.randint(0, 320, (2, 200), device=DEV)       # 400 tokens: the split (fp32-spectrum) tensor-core chain, fused
        l1 = model(input_ids=ids_long).logits
        with torch.inference_mode():                                 # inference tensors have no version counter (the bench's eval functions)
            l1i = model(input_ids=ids_long).logits
        assert [g.fused_calls for g in groups] == [3, 3, 3, 3]
        bc.unfuse_shared_inputs(model)
        l2 = model(input_ids=ids_long).logits
    assert torch.equal(ref, got) and torch.equal(l1, l2) and torch.equal(l1i, l2)


def test_cache_shims_and_data_writes():
    """`.data` writes do not move c._version: the cached spectrum goes stale until drop_spectrum(); disable_cfft_cache() restores the
    reference's rebuild-on-every-forward semantics (src/patch_llama_fft.py:256-257), enable_cfft_cache() the cached one."""
    torch.manual_seed(5)
    m = bc.BlockCirculantLinear(256, 256, block_size=64, bias=False).to(DEV).eval()
    holder = torch.nn.Sequential(m)
    x = torch.randn(3, 256, device=DEV)
    with torch.no_grad():
        y0 = m(x).clone()
        m.c.data.mul_(2.0)                                           # invisible to the version-keyed cache
        assert torch.equal(m(x), y0)
        m.drop_spectrum()
        assert torch.allclose(m(x), 2 * y0, rtol=1e-5, atol=1e-6)
        assert bc.disable_cfft_cache(holder) == 1
        m.c.data.mul_(0.5)
        assert torch.allclose(m(x), y0, rtol=1e-5, atol=1e-6)        # rebuilt on every forward now
        assert bc.enable_cfft_cache(holder) == 1
        m.c.data.mul_(3.0)
        assert torch.allclose(m(x), y0, rtol=1e-5, atol=1e-6)        # cached again


def test_from_linear_keeps_training_mode():
    lin = torch.nn.Linear(128, 128, bias=False).to(DEV)
    assert bc.BlockCirculantLinear.from_linear(lin, block_size=64).training
    assert not bc.BlockCirculantLinear.from_linear(lin.eval(), block_size=64).training
