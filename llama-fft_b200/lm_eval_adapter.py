"""lm-evaluation-harness model adapter ``llama_fft`` on the B200 path.

The reference's evaluation runs use a model adapter of that name (``lm-eval --model llama_fft --model_args
"pretrained=...,dtype=float16,device=cuda,block_size=128,num_layers=1,calib_path=...,cache_cfft=1"``, commands.txt:144-149; the tracebacks
in docs/result_16.02.2025.txt:208-209 name ``lm_eval/models/llama_fft.py``, class ``LlamaFftLM``) -- but that file lives in the author's
private fork of lm-evaluation-harness and is not part of the reference repository.  This module restates it
from its command line: same model name, same ``model_args`` keys, and the three request types of the harness' ``LM`` interface
(``loglikelihood``, ``loglikelihood_rolling``, ``generate_until``), with the block-circulant layers served by the CUDA kernels behind
include/bclinear.h.

``lm_eval`` itself is not installed in this image: when it is importable the class derives from ``lm_eval.api.model.LM`` and registers
itself as ``llama_fft``; otherwise it is a plain class with the same methods (requests are objects with an ``.args`` tuple, or tuples).
Extras beyond the reference's keys (all optional): ``patch_position``, ``targets=mlp|all``, ``spectrum=float32|bfloat16``,
``fuse_shared_inputs``, ``fuse_mlp``, ``max_length``, and ``random_init=tiny|7b`` for boxes without a checkpoint (byte-level tokenizer).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .patch_llama_fft import enable_cfft_cache, load_bc_params, patch_llama_fft, patch_mlp_with_block_circulant

try:                                                # pragma: no cover - lm_eval is not installed in the build image
    from lm_eval.api.model import LM as _LMBase
    from lm_eval.api.registry import register_model as _register
except Exception:                                   # noqa: BLE001
    class _LMBase:                                  # the part of lm_eval.api.model.LM this adapter relies on
        def __init__(self) -> None:
            self._rank, self._world_size = 0, 1

        @property
        def rank(self) -> int:
            return self._rank

        @property
        def world_size(self) -> int:
            return self._world_size

    def _register(*names):
        def deco(cls):
            return cls
        return deco


def _args(req) -> tuple:
    return tuple(req.args) if hasattr(req, "args") else tuple(req)


class _ByteTokenizer:
    """Stand-in when no tokenizer files exist (random-init models): UTF-8 byte b -> id 3 + b, 1 = BOS, 2 = EOS."""
    bos_token_id, eos_token_id, pad_token_id = 1, 2, 0

    def __init__(self, vocab_size: int):
        self.vocab_size = vocab_size

    def encode(self, text: str, add_special_tokens: bool = False) -> List[int]:
        ids = [3 + (b % max(self.vocab_size - 3, 1)) for b in text.encode("utf-8")]
        return ([self.bos_token_id] if add_special_tokens else []) + ids

    def decode(self, ids: Sequence[int], skip_special_tokens: bool = True) -> str:
        return bytes([(i - 3) % 256 for i in ids if i >= 3]).decode("utf-8", "replace")


@_register("llama_fft")
class LlamaFftLM(_LMBase):
    def __init__(self, pretrained: str = "", dtype: str = "float16", device: str = "cuda", block_size: int = 128, num_layers: int = 1,
                 patch_position: str = "first", calib_path: str = "", cache_cfft: int = 1, targets: str = "mlp", spectrum: str = "float32",
                 fuse_shared_inputs: int = 0, fuse_mlp: int = 0, batch_size: int = 1, max_length: int = 2048, random_init: str = "",
                 seed: int = 0, **_ignored) -> None:
        super().__init__()
        from .bench_all_in_one import RANDOM_INIT, pick_dtype
        self._device = torch.device(device)
        self._dtype = pick_dtype(dtype)
        self.max_length = int(max_length)
        self.batch_size = int(batch_size) if str(batch_size).isdigit() else 1
        if pretrained:
            from transformers import AutoModelForCausalLM, AutoTokenizer
            self.tokenizer = AutoTokenizer.from_pretrained(pretrained, local_files_only=True)
            self.model = AutoModelForCausalLM.from_pretrained(pretrained, torch_dtype=self._dtype, low_cpu_mem_usage=True, local_files_only=True)
        else:
            from transformers import LlamaConfig, LlamaForCausalLM
            cfg = LlamaConfig(**RANDOM_INIT[random_init or "tiny"])
            torch.manual_seed(int(seed))
            self.model = LlamaForCausalLM(cfg).to(self._dtype)
            self.tokenizer = _ByteTokenizer(cfg.vocab_size)
        self.model = self.model.to(self._device).eval()
        n = int(num_layers)
        self.bc_layers = 0
        if n > 0:
            sdt = torch.float32 if spectrum == "float32" else torch.bfloat16
            if targets == "all":
                self.bc_layers = patch_llama_fft(self.model, num_layers_to_patch=n, block_size=int(block_size), patch_position=patch_position, spectrum_dtype=sdt)
            else:                                                   # the reference's adapter patches the MLP of the first `num_layers` layers
                patch_mlp_with_block_circulant(self.model, num_layers_to_patch=n, block_size=int(block_size), patch_position=patch_position,
                                               spectrum_dtype=sdt, verbose=False)
                self.bc_layers = 3 * min(n, self.model.config.num_hidden_layers)
            if calib_path:
                load_bc_params(self.model, calib_path, strict_shapes=True)
            self.model.eval()
            if int(cache_cfft):
                enable_cfft_cache(self.model)
            if int(fuse_shared_inputs):
                from .shared_input import fuse_shared_inputs as _fuse
                _fuse(self.model)
            if int(fuse_mlp):
                from .chain import fuse_mlp as _fuse_mlp
                _fuse_mlp(self.model)

    # -- helpers ----------------------------------------------------------------------------------------------------
    @property
    def eot_token_id(self) -> int:
        return self.tokenizer.eos_token_id

    def tok_encode(self, text: str) -> List[int]:
        return list(self.tokenizer.encode(text, add_special_tokens=False))

    @torch.inference_mode()
    def _score(self, context: List[int], continuation: List[int]) -> Tuple[float, bool]:
        """log P(continuation | context) and whether the continuation is the greedy choice at every position."""
        ids = (context + continuation)[-self.max_length:]
        n_cont = min(len(continuation), len(ids) - 1)
        inp = torch.tensor([ids[:-1]], dtype=torch.long, device=self._device)
        logits = self.model(input_ids=inp).logits[0, -n_cont:].float()
        tgt = torch.tensor(ids[-n_cont:], dtype=torch.long, device=self._device)
        logp = F.log_softmax(logits, dim=-1)
        ll = logp.gather(-1, tgt[:, None]).sum().item()
        greedy = bool((logits.argmax(-1) == tgt).all().item())
        return ll, greedy

    # -- the harness' three request types ------------------------------------------------------------------------------
    def loglikelihood(self, requests) -> List[Tuple[float, bool]]:
        out = []
        for req in requests:
            context, continuation = _args(req)[:2]
            ctx = self.tok_encode(context) if context else [self.eot_token_id]
            out.append(self._score(ctx, self.tok_encode(continuation)))
        return out

    def loglikelihood_rolling(self, requests) -> List[float]:
        out = []
        for req in requests:
            (text,) = _args(req)[:1]
            toks = self.tok_encode(text)
            total, prev = 0.0, [self.eot_token_id]
            for i in range(0, len(toks), self.max_length - 1):      # disjoint windows, each conditioned on the previous window's last token
                window = toks[i:i + self.max_length - 1]
                total += self._score(prev, window)[0]
                prev = [window[-1]]
            out.append(total)
        return out

    @torch.inference_mode()
    def generate_until(self, requests) -> List[str]:
        out = []
        for req in requests:
            context, gen_kwargs = _args(req)[:2]
            gen_kwargs = dict(gen_kwargs or {})
            until = gen_kwargs.get("until", [])
            until = [until] if isinstance(until, str) else list(until)
            max_new = int(gen_kwargs.get("max_gen_toks", gen_kwargs.get("max_new_tokens", 64)))
            ids = self.tok_encode(context)[-(self.max_length - max_new):] or [self.eot_token_id]
            inp = torch.tensor([ids], dtype=torch.long, device=self._device)
            gen = self.model.generate(input_ids=inp, max_new_tokens=max_new, do_sample=False, use_cache=True,
                                      pad_token_id=self.tokenizer.pad_token_id if self.tokenizer.pad_token_id is not None else 0)
            text = self.tokenizer.decode(gen[0, len(ids):].tolist(), skip_special_tokens=True)
            for stop in until:
                if stop and stop in text:
                    text = text.split(stop)[0]
            out.append(text)
        return out
