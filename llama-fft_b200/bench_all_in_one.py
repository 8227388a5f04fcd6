"""The reference's unified bench (``src/bench_all_in_one.py``) on the B200 path: same metric definitions and result keys, so its
JSON/CSV consumers keep working, with the block-circulant layers served by the CUDA kernels behind ``include/bclinear.h``.

    reference symbol (src/bench_all_in_one.py)          here
    enable_cfft_cache / disable_cfft_cache  :164-254    re-exported from patch_llama_fft (spectrum is always version-cached)
    softmax_kl                              :112-120    softmax_kl
    correctness_last_token                  :297-346    correctness_last_token  (last_mse, last_kl, last_cos, last_top1_acc, ...)
    perf_prefill                            :349-390    perf_prefill            (prefill_avg_ms, prefill_tokens_per_s, peak_mem_bytes)
    perf_generate                           :393-427    perf_generate           (decode_new_tokens_per_s, decode_total_s)
    main                                    :455-       main                    (per block size: patch, [load calib], correctness, perf)

Differences, all additive: ``--random_init {tiny,7b}`` builds a random-init Llama of that architecture when no checkpoint
is on disk (this image has none; the reference needs ``--model_path``), prompts are tokenised with the checkpoint's tokenizer
when there is one and with a byte-level stand-in otherwise, ``--targets all`` patches q/k/v/o as well as the MLP,
``--fuse_shared_inputs`` turns on shared-input fusion, and the timed regions are additionally bracketed by CUDA events.
"""
import argparse
import json
import os
import time
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .patch_llama_fft import (enable_cfft_cache, disable_cfft_cache, patch_llama_fft, patch_mlp_with_block_circulant,  # noqa: F401
                              load_bc_params)
from .calibrate_bc import pad_batch

Tokenized = List[Tuple[torch.Tensor, torch.Tensor]]


def pick_dtype(name: str) -> torch.dtype:
    return {"fp16": torch.float16, "float16": torch.float16, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16,
            "fp32": torch.float32, "float32": torch.float32}[name]


def sync_if_cuda(device: torch.device) -> None:
    if device.type == "cuda":
        torch.cuda.synchronize(device)


def read_prompts(path: str, limit: int) -> List[str]:
    with open(path, "r", encoding="utf-8") as f:
        lines = [ln.strip() for ln in f if ln.strip()]
    return lines[:limit] if limit > 0 else lines


def byte_tokenize(prompts: Sequence[str], max_len: int, vocab_size: int) -> Tokenized:
    """Stand-in tokenizer (no tokenizer files in this image): UTF-8 byte b -> id 3 + b (mod vocab), truncated to max_len."""
    out = []
    for p in prompts:
        ids = torch.tensor([3 + (b % max(vocab_size - 3, 1)) for b in p.encode("utf-8")[:max_len]] or [3], dtype=torch.long)
        out.append((ids, torch.ones_like(ids)))
    return out


def iter_batches(tokenized: Tokenized, batch_size: int, pad_id: int):
    for i in range(0, len(tokenized), batch_size):
        yield pad_batch(tokenized[i:i + batch_size], pad_id)


def softmax_kl(student_logits: torch.Tensor, teacher_logits: torch.Tensor, temperature: float = 1.0) -> torch.Tensor:
    """KL(teacher || student) over the vocabulary; float32 logits; shape (B,)."""
    t = float(temperature)
    log_p = torch.log_softmax(teacher_logits / t, dim=-1)
    log_q = torch.log_softmax(student_logits / t, dim=-1)
    return (log_p.exp() * (log_p - log_q)).sum(dim=-1)


@torch.inference_mode()
def compute_teacher_last_cache(teacher, tokenized: Tokenized, device: torch.device) -> List[torch.Tensor]:
    """Last-token logits of the dense model, CPU float32 [(V,), ...] (reference :260-294 takes a path; a model object here)."""
    out = []
    for ids, attn in tokenized:
        logits = teacher(input_ids=ids.unsqueeze(0).to(device), attention_mask=attn.unsqueeze(0).to(device)).logits
        out.append(logits[0, int(attn.sum().item()) - 1, :].to(torch.float32).cpu())
    return out


@torch.inference_mode()
def correctness_last_token(student, tokenized: Tokenized, teacher_last_cache: List[torch.Tensor], device: torch.device,
                           temperature: float, topk: int) -> Dict[str, float]:
    acc = {"last_mse": 0.0, "last_kl": 0.0, "last_cos": 0.0, "last_top1_acc": 0.0, "last_top1_in_student_topk": 0.0, "last_topk_overlap": 0.0}
    n = 0
    for (ids, attn), t_cpu in zip(tokenized, teacher_last_cache):
        logits = student(input_ids=ids.unsqueeze(0).to(device), attention_mask=attn.unsqueeze(0).to(device)).logits
        s = logits[0, int(attn.sum().item()) - 1, :].to(torch.float32)
        t = t_cpu.to(device)
        acc["last_mse"] += (s - t).pow(2).mean().item()
        acc["last_kl"] += softmax_kl(s, t, temperature=temperature).mean().item()
        acc["last_cos"] += F.cosine_similarity(s.unsqueeze(0), t.unsqueeze(0), dim=-1).mean().item()
        k = min(topk, s.numel())
        s_top, t_top = torch.topk(s, k=k).indices.tolist(), torch.topk(t, k=k).indices.tolist()
        acc["last_top1_acc"] += float(s_top[0] == t_top[0])
        acc["last_top1_in_student_topk"] += float(t_top[0] in s_top)
        acc["last_topk_overlap"] += len(set(s_top) & set(t_top)) / float(k)
        n += 1
    return {k: v / float(max(n, 1)) for k, v in acc.items()}


def _timed(device: torch.device, fn, warmup: int, runs: int) -> Tuple[float, Optional[float]]:
    """(wall seconds, CUDA-event seconds or None) of `runs` calls after `warmup` untimed ones."""
    for _ in range(warmup):
        fn()
    sync_if_cuda(device)
    ev = None
    if device.type == "cuda":
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    t0 = time.perf_counter()
    for _ in range(runs):
        fn()
    if ev is not None:
        ev[1].record()
    sync_if_cuda(device)
    wall = max(time.perf_counter() - t0, 1e-9)
    return wall, (ev[0].elapsed_time(ev[1]) * 1e-3 if ev is not None else None)


@torch.inference_mode()
def perf_prefill(model, tokenized: Tokenized, device: torch.device, pad_id: int, batch_size: int, warmup: int, runs: int) -> Dict[str, float]:
    batches = [(i.to(device), a.to(device), int(a.sum().item())) for i, a, _ in iter_batches(tokenized, batch_size, pad_id)]
    if device.type == "cuda":
        torch.cuda.reset_peak_memory_stats(device)

    def once():
        for input_ids, attention_mask, _ in batches:
            model(input_ids=input_ids, attention_mask=attention_mask).logits
    wall, dev_s = _timed(device, once, warmup, runs)
    total_tokens = runs * sum(n for _, _, n in batches)
    out = {"prefill_avg_ms": wall / max(runs * len(batches), 1) * 1e3, "prefill_tokens_per_s": total_tokens / wall,
           "peak_mem_bytes": float(torch.cuda.max_memory_allocated(device)) if device.type == "cuda" else 0.0}
    if dev_s is not None:
        out["prefill_tokens_per_s_device_clock"] = total_tokens / max(dev_s, 1e-9)
    return out


@torch.inference_mode()
def perf_generate(model, tokenized: Tokenized, device: torch.device, pad_id: int, batch_size: int, warmup: int, runs: int,
                  max_new_tokens: int) -> Dict[str, float]:
    batches = [(i.to(device), a.to(device)) for i, a, _ in iter_batches(tokenized, batch_size, pad_id)]
    kw = dict(max_new_tokens=max_new_tokens, min_new_tokens=max_new_tokens, do_sample=False, use_cache=True, pad_token_id=pad_id)

    def once():
        for input_ids, attention_mask in batches:
            model.generate(input_ids=input_ids, attention_mask=attention_mask, **kw)
    wall, _ = _timed(device, once, warmup, runs)
    return {"decode_new_tokens_per_s": runs * len(tokenized) * max_new_tokens / wall, "decode_total_s": wall}


class GraphDecoder:
    """Greedy single-sequence decode with the per-token model step captured in ONE CUDA graph (HF StaticCache, static token / position
    buffers; token selection and position increment are part of the graph, so a generated token costs one graph replay and no host
    synchronisation).  Why: HF ``generate`` spends ~25 ms of Python per step on a 7B-shaped model while the GPU needs 0.7-2 ms
     -- per-layer microseconds only reach tokens/s once the host is out of the loop.  The reference's
    ``perf_generate`` (src/bench_all_in_one.py:393-427) times ``model.generate``; this is the same greedy decode, same metric."""

    def __init__(self, model, device: torch.device, max_cache_len: int):
        from transformers import StaticCache
        self.model, self.device, self.max_cache_len = model, device, max_cache_len
        self.cache = StaticCache(config=model.config, max_cache_len=max_cache_len)
        self.tok = torch.zeros(1, 1, dtype=torch.long, device=device)
        self.pos = torch.zeros(1, dtype=torch.long, device=device)
        self.out = torch.zeros(1, max_cache_len, dtype=torch.long, device=device)
        self.slot = torch.zeros(1, dtype=torch.long, device=device)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.stream = torch.cuda.Stream(device) if device.type == "cuda" else None

    def _step(self) -> None:
        logits = self.model(input_ids=self.tok, past_key_values=self.cache, cache_position=self.pos, use_cache=True).logits[:, -1]
        self.tok.copy_(logits.argmax(-1, keepdim=True))
        self.out.index_copy_(1, self.slot, self.tok)
        self.pos.add_(1)
        self.slot.add_(1)

    @torch.inference_mode()
    def generate(self, input_ids: torch.Tensor, max_new_tokens: int) -> torch.Tensor:
        """input_ids (1, T) -> (1, max_new_tokens) greedy continuation."""
        T = input_ids.shape[1]
        if T + max_new_tokens > self.max_cache_len:
            raise ValueError("prompt + new tokens exceed the static cache")
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _nullcontext()
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with ctx:
            self.cache.reset()
            logits = self.model(input_ids=input_ids, past_key_values=self.cache, cache_position=torch.arange(T, device=self.device), use_cache=True).logits[:, -1]
            self.tok.copy_(logits.argmax(-1, keepdim=True))
            self.out.zero_()
            self.out[:, 0:1] = self.tok
            self.pos.fill_(T)
            self.slot.fill_(1)
            steps = max_new_tokens - 1
            if self.stream is None:
                for _ in range(steps):
                    self._step()
            else:
                if self.graph is None and steps > 0:                # capture once: one eager step to warm every lazy allocation, then record
                    self._step()
                    steps -= 1
                    torch.cuda.synchronize(self.device)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.stream):
                        self._step()
                    self.graph = g
                    steps -= 1                                      # the capture pass does not execute: replay it once for this token
                    self.graph.replay()
                for _ in range(max(steps, 0)):
                    self.graph.replay()
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self.out[:, :max_new_tokens].clone()


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


@torch.inference_mode()
def perf_generate_graph(model, tokenized: Tokenized, device: torch.device, warmup: int, runs: int, max_new_tokens: int) -> Dict[str, float]:
    """The reference's decode metric (decode_new_tokens_per_s = runs * prompts * max_new_tokens / wall, src/bench_all_in_one.py:424-426)
    with the decode step under CUDA-graph replay, one sequence at a time (batch 1, the decode configuration of BASELINE.json)."""
    max_len = max(int(a.sum().item()) for _, a in tokenized) + max_new_tokens
    dec = GraphDecoder(model, device, max_len)
    prompts = [ids[:int(a.sum().item())].unsqueeze(0).to(device) for ids, a in tokenized]

    def once():
        for ids in prompts:
            dec.generate(ids, max_new_tokens)
    wall, _ = _timed(device, once, warmup, runs)
    return {"decode_graph_new_tokens_per_s": runs * len(prompts) * max_new_tokens / wall, "decode_graph_total_s": wall}


RANDOM_INIT = {
    "tiny": dict(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=8, vocab_size=320,
                 max_position_embeddings=512),
    "7b": dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=32, vocab_size=32000,
               max_position_embeddings=4096),      # Llama-2-7B (the reference's MODEL_PATH, src/patch_llama_fft.py:10)
}


def build_model(args, device: torch.device, dtype: torch.dtype):
    from transformers import LlamaConfig, LlamaForCausalLM
    if args.model_path:
        from transformers import AutoModelForCausalLM
        return AutoModelForCausalLM.from_pretrained(args.model_path, torch_dtype=dtype, low_cpu_mem_usage=True, local_files_only=True).to(device).eval()
    cfg = LlamaConfig(**RANDOM_INIT[args.random_init])
    torch.manual_seed(args.seed)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device(device):
            model = LlamaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(prev)
    return model.eval()


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model_path", type=str, default="")
    ap.add_argument("--random_init", choices=sorted(RANDOM_INIT), default="tiny", help="architecture to random-init when --model_path is empty")
    ap.add_argument("--prompts_file", type=str, required=True)
    ap.add_argument("--limit", type=int, default=20)
    ap.add_argument("--max_len", type=int, default=128)
    ap.add_argument("--block_sizes", type=str, default="64,128,256")
    ap.add_argument("--num_layers_to_patch", type=int, default=8)
    ap.add_argument("--patch_position", type=str, default="last", choices=["first", "last"])
    ap.add_argument("--targets", choices=["mlp", "all"], default="mlp", help="mlp = the reference's gate/up/down; all = + q/k/v/o")
    ap.add_argument("--spectrum", choices=["float32", "bfloat16"], default="float32")
    ap.add_argument("--fuse_shared_inputs", type=int, default=0)
    ap.add_argument("--fuse_mlp", type=int, default=0, help="single-token MLP calls as one chain launch (gate|up + SiLU*mul + down)")
    ap.add_argument("--graph_decode", type=int, default=0, help="also time greedy decode with the per-token step under CUDA-graph replay")
    ap.add_argument("--calib_dir", type=str, default="", help="directory with bc_calibrated_B{B}.pt (src/calibrate_bc.py:302)")
    ap.add_argument("--dtype", type=str, default="fp16")
    ap.add_argument("--device", type=str, default="cuda")
    ap.add_argument("--temperature", type=float, default=1.0)
    ap.add_argument("--topk", type=int, default=10)
    ap.add_argument("--batch_size", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--runs", type=int, default=5)
    ap.add_argument("--do_generate", type=int, default=0)
    ap.add_argument("--max_new_tokens", type=int, default=32)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out_json", type=str, default="")
    args = ap.parse_args(argv)

    device, dtype = torch.device(args.device), pick_dtype(args.dtype)
    sdt = torch.float32 if args.spectrum == "float32" else torch.bfloat16
    prompts = read_prompts(args.prompts_file, args.limit)
    rows = []
    dense_row = None
    teacher_cache = None
    for B in [0] + [int(b) for b in args.block_sizes.split(",") if b]:       # 0 = the dense model (teacher + dense timings)
        model = build_model(args, device, dtype)
        pad_id = 0
        if args.model_path:
            from transformers import AutoTokenizer
            tok = AutoTokenizer.from_pretrained(args.model_path, local_files_only=True)
            enc = [tok(p, truncation=True, max_length=args.max_len, return_tensors="pt") for p in prompts]
            tokenized = [(e["input_ids"][0], e["attention_mask"][0]) for e in enc]
            pad_id = tok.pad_token_id if tok.pad_token_id is not None else 0
        else:
            tokenized = byte_tokenize(prompts, args.max_len, model.config.vocab_size)
        row: Dict[str, float] = {"B": B, "calib_loaded": 0}
        if B:
            if args.targets == "all":
                row["bc_layers"] = patch_llama_fft(model, num_layers_to_patch=args.num_layers_to_patch, block_size=B,
                                                   patch_position=args.patch_position, spectrum_dtype=sdt)
            else:
                patch_mlp_with_block_circulant(model, num_layers_to_patch=args.num_layers_to_patch, block_size=B,
                                               patch_position=args.patch_position, spectrum_dtype=sdt, verbose=False)
            ck = os.path.join(args.calib_dir, f"bc_calibrated_B{B}.pt") if args.calib_dir else ""
            if ck and os.path.exists(ck):
                load_bc_params(model, ck, strict_shapes=True)
                row["calib_loaded"] = 1
            model.eval()                                   # swapped-in layers follow the model's mode; the cached spectrum is eval-only
            row["cfft_cached_layers"] = enable_cfft_cache(model)
            groups = []
            if args.fuse_shared_inputs:
                from .shared_input import fuse_shared_inputs
                row["shared_input_groups"] = fuse_shared_inputs(model)
                groups = list({id(m._shared_group): m._shared_group for m in model.modules() if getattr(m, "_shared_group", None) is not None}.values())
            if args.fuse_mlp:
                from .chain import fuse_mlp
                row["fused_mlps"] = fuse_mlp(model)
        if B == 0:
            teacher_cache = compute_teacher_last_cache(model, tokenized, device)
        else:
            row.update(correctness_last_token(model, tokenized, teacher_cache, device, args.temperature, args.topk))
        row.update(perf_prefill(model, tokenized, device, pad_id, args.batch_size, args.warmup, args.runs))
        if args.do_generate:
            row.update(perf_generate(model, tokenized, device, pad_id, args.batch_size, max(1, args.warmup // 2), max(1, args.runs // 2),
                                     args.max_new_tokens))
        if args.do_generate and args.graph_decode:
            try:
                row.update(perf_generate_graph(model, tokenized, device, max(1, args.warmup // 2), max(1, args.runs // 2), args.max_new_tokens))
            except Exception as e:                                   # noqa: BLE001 -- HF internals that cannot be captured: report, keep the eager figure
                row["decode_graph_error"] = f"{type(e).__name__}: {e}"[:300]
        if B and args.fuse_mlp:
            row["fused_mlp_calls"] = sum(m._bcl_fused_mlp.fused_calls for m in model.modules() if hasattr(m, "_bcl_fused_mlp"))
        if B and args.fuse_shared_inputs:
            row["shared_input_fused_calls"] = sum(g.fused_calls for g in groups)     # > 0: the fused path really served the forwards
        if B == 0:
            dense_row = row
        else:
            row["prefill_speedup_vs_dense"] = row["prefill_tokens_per_s"] / dense_row["prefill_tokens_per_s"]
            if args.do_generate:
                row["decode_speedup_vs_dense"] = row["decode_new_tokens_per_s"] / dense_row["decode_new_tokens_per_s"]
                if "decode_graph_new_tokens_per_s" in row and "decode_graph_new_tokens_per_s" in dense_row:
                    row["decode_graph_speedup_vs_dense"] = row["decode_graph_new_tokens_per_s"] / dense_row["decode_graph_new_tokens_per_s"]
        rows.append(row)
        print(json.dumps(row), flush=True)
        del model
        if device.type == "cuda":
            torch.cuda.empty_cache()
    if args.out_json:
        with open(args.out_json, "w") as f:
            json.dump({"args": vars(args), "rows": rows}, f, indent=1)
    return rows


if __name__ == "__main__":
    main()
