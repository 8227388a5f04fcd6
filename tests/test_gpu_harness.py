"""GPU tests (-m gpu) of the harness-level pieces of SURVEY 8f item 4: CUDA-graph greedy decode inside the unified bench
(reference src/bench_all_in_one.py:393-427), the ``llama_fft`` lm-eval adapter (commands.txt:144-149) and the layer microbenchmark
(src/microbench_dense_vs_bc.py) -- all on a tiny random-init Llama, all through the CUDA path."""
import math
import os

import pytest
import torch

import llama_fft_b200 as bc

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tiny(seed=0, layers=2):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=layers, num_attention_heads=4, num_key_value_heads=4,
                      vocab_size=320, max_position_embeddings=256)
    torch.manual_seed(seed)
    return LlamaForCausalLM(cfg).eval()


@pytest.mark.parametrize("fuse", [False, True])
def test_graph_decoder_matches_hf_generate(fuse):
    from llama_fft_b200.bench_all_in_one import GraphDecoder
    model = tiny()
    bc.patch_llama_fft(model, block_size=64)
    model = model.to(DEV).eval()
    if fuse:
        assert bc.fuse_shared_inputs(model) == 4 and bc.fuse_mlp(model) == 2
    dec = GraphDecoder(model, DEV, 64)
    for T in (9, 5, 12):                                             # the captured graph is reused across prompts of any length
        ids = torch.randint(3, 320, (1, T), device=DEV)
        with torch.no_grad():
            ref = model.generate(input_ids=ids, max_new_tokens=8, do_sample=False, use_cache=True, pad_token_id=0)[:, T:]
        got = dec.generate(ids, 8)
        assert got.shape == (1, 8)
        agree = (got == ref).float().mean().item()
        assert agree == 1.0 or (fuse and agree >= 0.75), (got.tolist(), ref.tolist())   # fused MLP rounds silu with fast intrinsics: argmax ties may differ
    assert dec.graph is not None
    if fuse:
        assert sum(m._bcl_fused_mlp.fused_calls for m in model.modules() if hasattr(m, "_bcl_fused_mlp")) > 0


def test_bench_all_in_one_graph_decode_row(tmp_path):
    from llama_fft_b200 import bench_all_in_one as bai
    rows = bai.main(["--random_init", "tiny", "--prompts_file", os.path.join(ROOT, "tools", "prompts_synthetic.txt"), "--limit", "3", "--max_len", "48",
                     "--block_sizes", "64", "--num_layers_to_patch", "2", "--targets", "all", "--batch_size", "3", "--warmup", "1", "--runs", "2",
                     "--do_generate", "1", "--max_new_tokens", "6", "--graph_decode", "1", "--fuse_mlp", "1", "--fuse_shared_inputs", "1"])
    for r in rows:
        assert "decode_graph_error" not in r, r.get("decode_graph_error")
        assert r["decode_graph_new_tokens_per_s"] > 0
    assert rows[1]["fused_mlps"] == 2 and rows[1]["fused_mlp_calls"] > 0 and rows[1]["decode_graph_speedup_vs_dense"] > 0


def test_lm_eval_adapter_scores_match_a_direct_computation():
    from llama_fft_b200.lm_eval_adapter import LlamaFftLM
    lm = LlamaFftLM(pretrained="", random_init="tiny", dtype="float32", device="cuda:0", block_size=64, num_layers=2, targets="all", cache_cfft=1)
    assert lm.bc_layers == 14
    ctx, cont = "The capital of France is", " Paris"
    (ll, greedy), = lm.loglikelihood([(ctx, cont)])
    ids = lm.tok_encode(ctx) + lm.tok_encode(cont)
    with torch.no_grad():
        logits = lm.model(input_ids=torch.tensor([ids[:-1]], device=DEV)).logits[0].float()
    n = len(lm.tok_encode(cont))
    want = torch.log_softmax(logits[-n:], -1).gather(-1, torch.tensor(ids[-n:], device=DEV)[:, None]).sum().item()
    assert math.isclose(ll, want, rel_tol=1e-5, abs_tol=1e-4) and isinstance(greedy, bool)
    (roll,) = lm.loglikelihood_rolling([("hello world, hello block-circulant world",)])
    assert roll < 0
    (text,) = lm.generate_until([("Once upon a time", {"until": ["\n"], "max_gen_toks": 5})])
    assert isinstance(text, str)

    class Req:                                                      # the harness passes Instance objects with .args
        def __init__(self, *a):
            self.args = a
    assert lm.loglikelihood([Req(ctx, cont)])[0][0] == ll


def test_microbench_mirror_runs_and_reports_the_reference_rows():
    from llama_fft_b200 import microbench_dense_vs_bc as mb
    res = mb.main(["--batch", "1", "--seq", "4", "--block_sizes", "128", "--warmup", "1", "--iters", "2", "--bench_mlp", "--config1", "--json"])
    assert set(res["dense"]) == {"gate 4096->11008", "up   4096->11008", "down 11008->4096", "MLP"}
    assert set(res["bc"]["128"]) == set(res["dense"]) and all(v > 0 for v in res["bc"]["128"].values())
    assert res["config1"]["bc_ms"] > 0
    with pytest.raises(SystemExit):
        mb.main(["--device", "cpu"])
