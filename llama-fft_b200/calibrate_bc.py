"""Calibration of block-circulant layers against a dense teacher -- drop-in for the reference's ``src/calibrate_bc.py``.

Same command line (flags of src/calibrate_bc.py:112-136), same helper names, same outputs (``bc_calibrated_B{B}.pt`` written by
``save_bc_params``, the teacher / token caches), same objective: KL(teacher || student) on the last-token logits, AdamW on the
BC parameters only (src/calibrate_bc.py:253-294).  What differs is underneath: the student's BC layers run -- forward AND
backward -- on this repo's sm_100a kernels (autograd.py), and patching uses the GPU projection kernel instead of Python loops.

Additions (all optional, defaults reproduce the reference): ``--model_path``, ``--targets mlp|all`` (the reference only patches
the MLP, src/patch_llama_fft.py:379), ``--device``, and ``--tokenizer byte`` -- a byte-level stand-in (UTF-8 byte b -> id 3+b)
for boxes that have the weights but no tokenizer files.  The accelerate-offload loading of the reference
(``device_map="auto"``, 6 GiB caps for a consumer GPU, :179-188) is replaced by a plain ``.to(device)``: a B200 holds teacher and
student at once.  ``run_calibration`` is the loop itself, importable so it can be driven with any (tiny) model in tests.
"""
from __future__ import annotations

import argparse
import gc
from pathlib import Path
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .patch_llama_fft import (MODEL_PATH, BlockCirculantLinear, patch_llama_fft, patch_mlp_with_block_circulant, save_bc_params)


# ---------------------------------------------------------------------------------------------------------
# helpers (names and contracts of src/calibrate_bc.py:23-107)
# ---------------------------------------------------------------------------------------------------------
def pick_input_device(m: nn.Module) -> torch.device:
    dm = getattr(m, "hf_device_map", None)
    if dm:
        d = dm.get("model.embed_tokens", next(iter(dm.values())))
        return torch.device(f"cuda:{d}") if isinstance(d, int) else torch.device(str(d))
    return next(m.parameters()).device


def default_prompts_file() -> str:
    """The reference defaults to ``prompts_100.txt`` next to the script (src/calibrate_bc.py:113).  Its prompt files are data of the
    reference repo and are not shipped here: a copy placed next to this module is picked up, otherwise the synthetic prompt set that
    ships with this repo (tools/prompts_synthetic.txt) is the default."""
    here = Path(__file__).resolve()
    for cand in (here.with_name("prompts_100.txt"), here.parent.parent / "tools" / "prompts_synthetic.txt"):
        if cand.exists():
            return str(cand)
    return str(here.with_name("prompts_100.txt"))


def load_texts(path: Path, limit: int) -> List[str]:
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"Prompts file not found: {path} (pass --prompts_file; the reference's src/prompts_*.txt are not shipped "
                                f"with this package, tools/prompts_synthetic.txt is)")
    lines = (ln.strip() for ln in path.read_text(encoding="utf-8").splitlines())
    texts = [ln for ln in lines if ln and not ln.startswith("#")][:limit]
    if not texts:
        raise ValueError(f"No prompts loaded from: {path}")
    return texts


def freeze_all_params(model: nn.Module) -> None:
    for prm in model.parameters():
        prm.requires_grad = False


def enable_bc_trainable(model: nn.Module, *, train_bias: bool = True) -> int:
    """Unfreeze c (and bias) of every BC layer; returns the number of trainable scalars."""
    count = 0
    for mod in model.modules():
        if not isinstance(mod, BlockCirculantLinear):
            continue
        mod.c.requires_grad = True
        count += mod.c.numel()
        if train_bias and mod.bias is not None:
            mod.bias.requires_grad = True
            count += mod.bias.numel()
    return count


def iter_trainable_params(model: nn.Module) -> Iterable[nn.Parameter]:
    return (prm for prm in model.parameters() if prm.requires_grad)


def pad_batch(items: Sequence[Tuple[torch.Tensor, torch.Tensor]], pad_id: int):
    """[(ids, mask)] of 1-D tensors -> (input_ids (B,T), attention_mask (B,T), last_idx (B,)), right padded."""
    width = max(int(ids.numel()) for ids, _ in items)
    input_ids = torch.full((len(items), width), pad_id, dtype=torch.long)
    attention_mask = torch.zeros((len(items), width), dtype=torch.long)
    for r, (ids, mask) in enumerate(items):
        input_ids[r, :ids.numel()] = ids
        attention_mask[r, :mask.numel()] = mask
    return input_ids, attention_mask, attention_mask.sum(dim=1) - 1


def gather_last_logits(logits: torch.Tensor, last_idx: torch.Tensor) -> torch.Tensor:
    """(B,T,V), (B,) -> (B,V): the logits at each sample's last real token."""
    return logits[torch.arange(logits.shape[0], device=logits.device), last_idx.to(logits.device)]


# ---------------------------------------------------------------------------------------------------------
# tokenisation / teacher cache
# ---------------------------------------------------------------------------------------------------------
class ByteTokenizer:
    """Stand-in when no tokenizer files exist: BOS=1, EOS/pad=2, UTF-8 byte b -> 3+b (Llama's byte-fallback range)."""
    pad_token_id = 2
    eos_token_id = 2

    def __call__(self, text: str, return_tensors: str = "pt", truncation: bool = True, max_length: int = 128):
        ids = [1] + [3 + b for b in text.encode("utf-8")]
        if truncation:
            ids = ids[:max_length]
        t = torch.tensor([ids], dtype=torch.long)
        return {"input_ids": t, "attention_mask": torch.ones_like(t)}


def tokenize_prompts(tokenizer, texts: Sequence[str], max_len: int, cache: Optional[Path] = None):
    if cache is not None and Path(cache).exists():
        blob = torch.load(cache, map_location="cpu")
        return blob["input_ids"], blob["attention_mask"]
    ids, masks = [], []
    for t in texts:
        enc = tokenizer(t, return_tensors="pt", truncation=True, max_length=max_len)
        ids.append(enc["input_ids"][0].cpu())
        masks.append(enc["attention_mask"][0].cpu())
    if cache is not None:
        torch.save({"input_ids": ids, "attention_mask": masks}, cache)
    return ids, masks


@torch.inference_mode()
def teacher_last_logits(teacher: nn.Module, tok_ids, tok_attn, cache: Optional[Path] = None) -> List[torch.Tensor]:
    """fp32 last-token logits of the dense teacher, one (V,) tensor per prompt (src/calibrate_bc.py:192-212)."""
    if cache is not None and Path(cache).exists():
        return torch.load(cache, map_location="cpu")
    dev = pick_input_device(teacher)
    out = []
    for ids, mask in zip(tok_ids, tok_attn):
        logits = teacher(input_ids=ids[None].to(dev), attention_mask=mask[None].to(dev)).logits
        out.append(logits[0, int(mask.sum()) - 1].float().cpu())
    if cache is not None:
        torch.save(out, cache)
    return out


# ---------------------------------------------------------------------------------------------------------
# the calibration loop (src/calibrate_bc.py:253-294)
# ---------------------------------------------------------------------------------------------------------
def run_calibration(student: nn.Module, tok_ids, tok_attn, teacher_last_cache, *, pad_id: int, steps: int = 200, lr: float = 1e-3,
                    batch_size: int = 4, train_bias: bool = False, log_every: int = 10, verbose: bool = True) -> List[float]:
    """Freeze everything but the BC parameters and minimise KL(teacher || student) of the last-token distribution with
    AdamW over round-robin padded batches.  Returns the per-step losses."""
    freeze_all_params(student)
    n_train = enable_bc_trainable(student, train_bias=train_bias)
    if verbose:
        print(f"Trainable parameters (BC only): {n_train}")
    optim = torch.optim.AdamW(list(iter_trainable_params(student)), lr=lr)
    dev = pick_input_device(student)
    n_prompts = len(tok_ids)
    batch_size = max(1, batch_size)
    losses: List[float] = []
    for step in range(steps):
        picks = [(step * batch_size + k) % n_prompts for k in range(batch_size)]
        input_ids, attention_mask, last_idx = pad_batch([(tok_ids[i], tok_attn[i]) for i in picks], pad_id)
        logits = student(input_ids=input_ids.to(dev), attention_mask=attention_mask.to(dev)).logits
        student_last = gather_last_logits(logits, last_idx).float()
        teacher_last = torch.stack([teacher_last_cache[i] for i in picks]).to(dev)
        loss = F.kl_div(F.log_softmax(student_last, dim=-1), F.softmax(teacher_last, dim=-1), reduction="batchmean")
        optim.zero_grad(set_to_none=True)
        loss.backward()
        optim.step()
        losses.append(float(loss.item()))
        if verbose and (step + 1) % log_every == 0:
            recent = losses[-log_every:]
            print(f"Step {step + 1:4d}/{steps} | KL(last-token) avg{log_every}={sum(recent) / len(recent):.6f}")
    return losses


def _load_causal_lm(model_path: str, device: torch.device, dtype: torch.dtype):
    from transformers import AutoModelForCausalLM
    model = AutoModelForCausalLM.from_pretrained(model_path, dtype=dtype, low_cpu_mem_usage=True, local_files_only=True)
    return model.to(device).eval()


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="Calibrate block-circulant layers against the dense teacher (B200 kernels).")
    ap.add_argument("--prompts_file", type=str, default=default_prompts_file())
    ap.add_argument("--limit", type=int, default=50)
    ap.add_argument("--block_sizes", type=str, default="64,128,256")
    ap.add_argument("--num_layers_to_patch", type=int, default=1)
    ap.add_argument("--patch_position", type=str, default="first", choices=["first", "last"])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--train_bias", action="store_true")
    ap.add_argument("--batch_size", type=int, default=4)
    ap.add_argument("--max_len", type=int, default=128)
    ap.add_argument("--out_dir", type=str, default="calib_out")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--teacher_cache", type=str, default="teacher_last_cache.pt")
    ap.add_argument("--token_cache", type=str, default="tokenized_prompts_cache.pt")
    # additions
    ap.add_argument("--model_path", type=str, default=MODEL_PATH)
    ap.add_argument("--targets", type=str, default="mlp", choices=["mlp", "all"], help="mlp: gate/up/down (reference); all: + q/k/v/o")
    ap.add_argument("--device", type=str, default="cuda:0")
    ap.add_argument("--tokenizer", type=str, default="auto", choices=["auto", "byte"])
    return ap


def main(argv: Optional[Sequence[str]] = None) -> None:
    args = build_parser().parse_args(argv)
    torch.manual_seed(args.seed)
    device = torch.device(args.device)
    out_dir = Path(args.out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)

    texts = load_texts(Path(args.prompts_file), args.limit)
    if args.tokenizer == "byte":
        tokenizer = ByteTokenizer()
    else:
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(args.model_path, local_files_only=True)
        if tokenizer.pad_token_id is None:
            tokenizer.pad_token = tokenizer.eos_token
    pad_id = int(tokenizer.pad_token_id)
    tok_ids, tok_attn = tokenize_prompts(tokenizer, texts, args.max_len, Path(args.token_cache))

    cache_path = Path(args.teacher_cache)
    if cache_path.exists():
        teacher_cache = torch.load(cache_path, map_location="cpu")
    else:
        print("Loading teacher (original) model...")
        teacher = _load_causal_lm(args.model_path, device, torch.float16)
        print("Precomputing teacher last-token cache...")
        teacher_cache = teacher_last_logits(teacher, tok_ids, tok_attn, cache_path)
        del teacher
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
        print("Teacher freed.")

    block_sizes = [int(v) for v in args.block_sizes.split(",") if v.strip()]
    if not block_sizes:
        raise ValueError("No block sizes specified")
    for B in block_sizes:
        print(f"\n=== Calibrating block_size={B} ===")
        print("Loading student (will be patched) model...")
        student = _load_causal_lm(args.model_path, device, torch.float16)
        if args.targets == "all":
            patch_llama_fft(student, num_layers_to_patch=args.num_layers_to_patch, block_size=B, patch_position=args.patch_position, verbose=True)
        else:
            patch_mlp_with_block_circulant(student, num_layers_to_patch=args.num_layers_to_patch, block_size=B, patch_position=args.patch_position)
        print("Starting calibration...")
        run_calibration(student, tok_ids, tok_attn, teacher_cache, pad_id=pad_id, steps=args.steps, lr=args.lr,
                        batch_size=args.batch_size, train_bias=args.train_bias)
        out_path = out_dir / f"bc_calibrated_B{B}.pt"
        save_bc_params(student, str(out_path))
        print(f"Saved: {out_path}")
        del student
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
    print("\nDone.")


if __name__ == "__main__":
    main()
