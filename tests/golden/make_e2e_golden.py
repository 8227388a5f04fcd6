#!/usr/bin/env python3
"""End-to-end golden fixture: a tiny random-init Llama whose seven projections per layer are block-circulant, run by the
UNMODIFIED reference (/root/reference/src/patch_llama_fft.py) on CPU, on lines of the reference's own prompt files.

Build container only (the GPU box has no /root/reference):   python tests/golden/make_e2e_golden.py

Records to tests/golden/e2e_tiny_llama.pt
    config     LlamaConfig kwargs (Llama-2 architecture, tiny dims; B=64)
    state      state_dict of the patched reference model (embeddings, norms, lm_head, every BC layer's `c`)
    input_ids  (P, T) byte-level token ids (stand-in tokenizer: UTF-8 byte b -> id 3+b, 0 = pad),
    lengths    (P,) unpadded lengths
    logits     (P, T, vocab) float32 logits of the reference model (full sequence)
    last_logits / nll   the quantities the reference's benches report (last-token logits, src/bench_all_in_one.py:297-345;
                        mean next-token negative log-likelihood -> perplexity)
The reference only patches the MLP (src/patch_llama_fft.py:379); q/k/v/o are swapped here with the reference's own
BlockCirculantLinear.from_linear, which is what BASELINE.json's "all 7 projections" configuration means.
"""
import contextlib
import io
import os
import sys

import torch

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
import patch_llama_fft as plf  # noqa: E402
from transformers import LlamaConfig, LlamaForCausalLM  # noqa: E402

CFG = dict(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
           num_key_value_heads=4, max_position_embeddings=128, rms_norm_eps=1e-5, hidden_act="silu", tie_word_embeddings=False,
           attention_bias=False, mlp_bias=False)
B = 64
T = 48


def byte_tokenize(lines, T):
    ids = torch.zeros(len(lines), T, dtype=torch.long)
    lens = []
    for r, s in enumerate(lines):
        b = list(s.encode("utf-8"))[:T - 1]
        row = [1] + [3 + v for v in b]                     # 1 = BOS like Llama
        ids[r, :len(row)] = torch.tensor(row)
        lens.append(len(row))
    return ids, torch.tensor(lens)


def main():
    torch.manual_seed(1234)
    model = LlamaForCausalLM(LlamaConfig(**CFG)).eval()
    with contextlib.redirect_stdout(io.StringIO()):
        plf.patch_mlp_with_block_circulant(model, num_layers_to_patch=CFG["num_hidden_layers"], block_size=B, patch_position="first")
    for layer in model.model.layers:                          # attention projections with the reference's own constructor
        for name in ("q_proj", "k_proj", "v_proj", "o_proj"):
            setattr(layer.self_attn, name, plf.BlockCirculantLinear.from_linear(getattr(layer.self_attn, name), block_size=B))
    lines = []
    for fn, k in (("prompts_20_quality.txt", 6), ("prompts_100.txt", 6)):
        with open(os.path.join(REF, fn), encoding="utf-8") as f:
            lines += [ln.strip() for ln in f if ln.strip()][:k]
    ids, lens = byte_tokenize(lines, T)
    mask = (torch.arange(T)[None, :] < lens[:, None]).long()
    with torch.no_grad():
        logits = model(input_ids=ids, attention_mask=mask).logits.float()
    last = logits[torch.arange(len(lines)), lens - 1]
    logp = torch.log_softmax(logits[:, :-1], dim=-1)
    tgt = ids[:, 1:]
    tok_nll = -logp.gather(-1, tgt[..., None])[..., 0]
    valid = (torch.arange(T - 1)[None, :] < (lens - 1)[:, None])
    nll = (tok_nll * valid).sum() / valid.sum()
    out = dict(config=CFG, block_size=B, state={k: v.clone() for k, v in model.state_dict().items()}, input_ids=ids, lengths=lens,
               logits=logits, last_logits=last, nll=nll, n_bc=sum(isinstance(m, plf.BlockCirculantLinear) for m in model.modules()))
    path = os.path.join(HERE, "e2e_tiny_llama.pt")
    torch.save(out, path)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB, {len(lines)} prompts, nll {nll.item():.6f}, ppl {nll.exp().item():.3f}, BC layers {out['n_bc']}")


if __name__ == "__main__":
    main()
