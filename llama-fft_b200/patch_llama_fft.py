"""Drop-in for the reference's ``src/patch_llama_fft.py``: same names, signatures, parameters, state-dict keys and
error behaviour -- but ``BlockCirculantLinear.forward`` calls hand-written sm_100a kernels through the C ABI of
``libbclinear.so`` instead of torch.fft/einsum.

Reference map (paths relative to /root/reference/):
    BlockCirculantLinear.__init__/forward     src/patch_llama_fft.py:142-275
    cached-spectrum semantics                 src/bench_all_in_one.py:164-254
    from_linear                               src/patch_llama_fft.py:180-224
    projections                               src/patch_llama_fft.py:47-139
    save/load_bc_params                       src/patch_llama_fft.py:287-340
    patch_mlp_with_block_circulant            src/patch_llama_fft.py:343-394
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, Optional, Sequence

import torch
import torch.nn as nn

from . import _cabi

# Module constants the reference exports (src/patch_llama_fft.py:10-15); other scripts import them.
MODEL_PATH = os.environ.get("LLAMA_FFT_MODEL_PATH", "Llama-2-7b-hf")
BLOCK_SIZE = 256
NUM_LAYERS_TO_PATCH = 1

# Where a layer switches from the CUDA-core stream kernel to the tcgen05 path.  Measured per-module on B200 (tools/prefill_vs_dense.py,
# profiles/r1_prefill_notes.md): the stream kernel costs ~17 us + 1.05e-4 us per (token x block pair), the tensor-core chain a flat
# ~24 us (bf16 spectra) / ~33 us (fp32 spectra, split operands) up to a few hundred tokens.  So the switch is on WORK, tokens x p x q:
#   4096->4096 (1024 block pairs): from 64 / 156 tokens;  4096<->11008 (2752 pairs): from 24 / 58 tokens.
# BCL_PREFILL_MIN_TOKENS / BCL_PREFILL_MIN_TOKENS_F32, when set, replace the rule by a fixed token floor.
PREFILL_MIN_WORK = 65536            # tokens x p x q, bf16 spectra
PREFILL_MIN_WORK_F32 = 160000       # fp32 spectra (split-bf16 tcgen05 GEMM + fp32-accurate DFT stages)
PREFILL_MIN_TOKENS = int(os.environ["BCL_PREFILL_MIN_TOKENS"]) if "BCL_PREFILL_MIN_TOKENS" in os.environ else None
PREFILL_MIN_TOKENS_F32 = int(os.environ["BCL_PREFILL_MIN_TOKENS_F32"]) if "BCL_PREFILL_MIN_TOKENS_F32" in os.environ else None


def prefill_token_floor(p: int, q: int, spectrum_dtype: torch.dtype) -> int:
    """Tokens per call from which a (p x q)-block layer takes the tensor-core path."""
    fixed = PREFILL_MIN_TOKENS if spectrum_dtype == torch.bfloat16 else PREFILL_MIN_TOKENS_F32
    if fixed is not None:
        return fixed
    work = PREFILL_MIN_WORK if spectrum_dtype == torch.bfloat16 else PREFILL_MIN_WORK_F32
    return max(16, -(-work // (p * q)))

MLP_TARGETS = ("gate_proj", "up_proj", "down_proj")
ATTN_TARGETS = ("q_proj", "k_proj", "v_proj", "o_proj")


# ---------------------------------------------------------------------------------------------------------
# dense -> circulant projections (construction-time host logic; vectorised, no Python loop over B)
# ---------------------------------------------------------------------------------------------------------
def _wrapped_diagonals(W_blocks: torch.Tensor, sign: int) -> torch.Tensor:
    """(..., B, B) -> (..., t, r) with element W[r, (r + sign*t) % B]."""
    B = W_blocks.shape[-1]
    r = torch.arange(B, device=W_blocks.device)
    cols = (r[None, :] + sign * r[:, None]) % B
    return W_blocks[..., r[None, :].expand(B, B), cols]


def dense_block_to_circulant_column(W_block: torch.Tensor, *, convention: str = "diag") -> torch.Tensor:
    """Legacy Frobenius projections (reference :47-79).
    "diag": c[k] = mean_i W[i,(i+k)%B];  "diag_inv": c[k] = mean_i W[(i+k)%B, i]."""
    assert W_block.dim() == 2 and W_block.shape[0] == W_block.shape[1]
    if convention not in ("diag", "diag_inv"):
        raise ValueError(f"Unknown convention: {convention}")
    src = W_block if convention == "diag" else W_block.transpose(0, 1)
    return _wrapped_diagonals(src, +1).mean(-1)


def _detect_best_convention_for_layer(W: torch.Tensor, block_size: int, *, num_probes: int = 3) -> str:
    """Legacy sanity check (reference :17-44; its call sites are commented out there, :210/:218): which of the two legacy
    projections of the top-left BxB block reproduces the dense matvec better on `num_probes` random vectors (seed 0, :30).
    Returns "diag" or "diag_inv".  Construction-time host logic on W's device; the circulant product is formed explicitly."""
    B = block_size
    block = W[:B, :B].detach().to(dtype=torch.float32)
    torch.manual_seed(0)
    xs = torch.stack([torch.randn(B, device=block.device, dtype=torch.float32) for _ in range(num_probes)], dim=1)   # (B, probes)
    y_dense = block @ xs

    def score(convention: str) -> float:
        C = circulant_from_first_col(dense_block_to_circulant_column(block, convention=convention).to(torch.float32))
        return (y_dense - C @ xs).pow(2).mean(dim=0).sum().item() / float(num_probes)
    return "diag" if score("diag") <= score("diag_inv") else "diag_inv"


def dense_block_to_circulant_column_loss_aware(W_block: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """First column c of the circulant C[i,j] = c[(i-j)%B] closest to W in Frobenius norm, times the
    energy-matching scalar alpha = <W,C>/<C,C> (reference :82-128).  Accepts (..., B, B); arithmetic stays in
    W's dtype like the reference's."""
    assert W_block.dim() >= 2 and W_block.shape[-1] == W_block.shape[-2]
    B = W_block.shape[-1]
    diag = _wrapped_diagonals(W_block, -1)
    sums = diag.sum(-1)
    c = diag.mean(-1)
    alpha = (c * sums).sum(-1, keepdim=True) / (B * (c * c).sum(-1, keepdim=True)).clamp_min(eps)
    return (alpha * c).to(W_block.dtype)


def circulant_from_first_col(c: torch.Tensor) -> torch.Tensor:
    """Explicit BxB circulant with first column c: C[i,j] = c[(i-j)%B] (reference :132-139)."""
    B = c.numel()
    idx = torch.arange(B, device=c.device)
    return c[(idx[:, None] - idx[None, :]) % B]


def _project_weight(weight: torch.Tensor, B: int) -> torch.Tensor:
    """(out,in) dense weight -> c (p,q,B) float32, loss-aware projection of every BxB block."""
    out_f, in_f = weight.shape
    p, q = out_f // B, in_f // B
    if weight.is_cuda and B in (64, 128, 256) and weight.dtype in _cabi.DTYPE:
        W = weight.detach()
        if W.stride(1) != 1:
            W = W.contiguous()
        c = torch.empty(p, q, B, device=W.device, dtype=torch.float32)
        with torch.cuda.device(W.device):
            rc = _cabi.lib().bcl_project_from_dense(W.data_ptr(), _cabi.DTYPE[W.dtype], W.stride(0), p, q, B,
                                                    c.data_ptr(), _cabi.stream_ptr(W.device))
        _cabi.check(rc, "bcl_project_from_dense")
        return c
    # CPU weights / block sizes the kernels do not cover: vectorised host construction, chunked over p to bound memory
    Wb = weight.detach().reshape(p, B, q, B).permute(0, 2, 1, 3)
    rows = [dense_block_to_circulant_column_loss_aware(Wb[j]) for j in range(p)]
    return torch.stack(rows, 0).to(torch.float32)


# ---------------------------------------------------------------------------------------------------------
# the layer
# ---------------------------------------------------------------------------------------------------------
class BlockCirculantLinear(nn.Module):
    """y = W_bc x + b with W_bc a (out_blocks x in_blocks) grid of BxB circulant blocks; c[j,i,:] is the first
    column of the block mapping input block i to output block j (reference :142-178).

    Extra, keyword-only knob (not in the reference): ``spectrum_dtype`` -- torch.float32 (default; parity
    <= 1e-4 vs the reference) or torch.bfloat16 (half the HBM stream; parity <= 1e-2)."""

    def __init__(self, in_features: int, out_features: int, block_size: int = 256, bias: bool = True, *,
                 spectrum_dtype: torch.dtype = torch.float32):
        super().__init__()
        assert in_features % block_size == 0, "in_features must be divisible by block_size"
        assert out_features % block_size == 0, "out_features must be divisible by block_size"
        self.in_features = in_features
        self.out_features = out_features
        self.block_size = block_size
        self.in_blocks = in_features // block_size
        self.out_blocks = out_features // block_size
        self.c = nn.Parameter(torch.randn(self.out_blocks, self.in_blocks, self.block_size) * 0.01)
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        if spectrum_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("spectrum_dtype must be torch.float32 or torch.bfloat16")
        self.spectrum_dtype = spectrum_dtype
        self._spec: Optional[torch.Tensor] = None          # STREAM layout (decode / small N)
        self._spec_gemm: Optional[torch.Tensor] = None     # GEMM layout (tcgen05 prefill path)
        self._spec_key = None
        self._cache_spectrum = True                        # False after disable_cfft_cache(): rfft(c) is recomputed on every forward
        self._shared_group = None                          # shared_input.SharedInputGroup when fused with siblings (q/k/v, gate/up)

    def extra_repr(self) -> str:
        return (f"in_features={self.in_features}, out_features={self.out_features}, block_size={self.block_size}, "
                f"bias={self.bias is not None}, spectrum_dtype={self.spectrum_dtype}")

    # -- construction ---------------------------------------------------------------------------------------
    @classmethod
    def from_linear(cls, linear: nn.Linear, block_size: int = 256, **kwargs) -> "BlockCirculantLinear":
        """Approximate a dense nn.Linear (reference :180-224): parameters are created in fp32 on the weight's
        device, every BxB block is projected with the loss-aware projection, the bias is copied."""
        layer = cls(linear.in_features, linear.out_features, block_size=block_size, bias=(linear.bias is not None), **kwargs)
        layer = layer.to(linear.weight.device)
        with torch.no_grad():
            layer.c.copy_(_project_weight(linear.weight.data, block_size))
            if layer.bias is not None and linear.bias is not None:
                layer.bias.copy_(linear.bias.data)
        # A swapped-in module must behave like the one it replaces: a model that was put in eval() before patching stays in eval
        # (otherwise the cached spectrum and the shared-input groups would never be used: both are eval / no-grad only).
        layer.train(linear.training)
        return layer

    # -- spectrum lifecycle ---------------------------------------------------------------------------------
    def _spectrum_key(self):
        c = self.c
        return (c.data_ptr(), c._version, c.device, c.dtype, self.spectrum_dtype)

    def build_spectrum(self, *, force: bool = False, layout: int = _cabi.LAYOUT_STREAM) -> torch.Tensor:
        """rfft(c) on c's device (replaces reference :256-257 / the _Cf_cache of src/bench_all_in_one.py:204-219), in the
        STREAM layout (decode kernel) or the GEMM layout (tcgen05 prefill kernels; fp32 spectra -> split bf16 pairs).  Re-used until c is
        written to (its version counter moves), moved, or re-typed -- so load_bc_params / optimizer steps invalidate
        it, which the reference's cache key does not.
        Writes that bypass autograd's version counter (``layer.c.data.copy_(...)``, ``layer.c.data[j, i] = ...`` as the reference's own
        from_linear does, :214-219) are NOT seen: call ``drop_spectrum()`` after them, or ``disable_cfft_cache(model)`` to get the
        reference's default behaviour (rfft(c) recomputed on every forward)."""
        key = self._spectrum_key()
        if self._spec_key != key:
            self._spec, self._spec_gemm, self._spec_key = None, None, key
        cached = self._spec if layout == _cabi.LAYOUT_STREAM else self._spec_gemm
        if not force and cached is not None:
            return cached
        c = self.c.detach()
        _cabi.require_cuda(c, "BlockCirculantLinear.c")
        c32 = c.to(torch.float32).contiguous()
        L = _cabi.lib()
        if layout == _cabi.LAYOUT_STREAM:
            spec = torch.empty(self.out_blocks, self.in_blocks, self.block_size, device=c.device, dtype=self.spectrum_dtype)
        else:
            # bf16: one bf16 matrix per frequency; fp32: (hi, lo) bf16 pairs, v = hi + lo to ~2^-17 (the 1e-4 parity class)
            nbytes = L.bcl_spectrum_bytes(self.out_blocks, self.in_blocks, self.block_size, _cabi.DTYPE[self.spectrum_dtype], layout)
            spec = torch.empty(nbytes, device=c.device, dtype=torch.uint8)      # Cmat + DFT matrices (include/bclinear.h)
        with torch.cuda.device(c.device):
            rc = L.bcl_precompute_spectrum(c32.data_ptr(), self.out_blocks, self.in_blocks, self.block_size,
                                           spec.data_ptr(), _cabi.DTYPE[self.spectrum_dtype], layout,
                                           _cabi.stream_ptr(c.device))
        _cabi.check(rc, "bcl_precompute_spectrum")
        if layout == _cabi.LAYOUT_STREAM:
            self._spec = spec
        else:
            self._spec_gemm = spec
        return spec

    def drop_spectrum(self) -> None:
        self._spec, self._spec_gemm, self._spec_key = None, None, None

    def _use_prefill_path(self, n_tokens: int) -> bool:
        """Enough tokens to fill 128-row MMA tiles -> tcgen05 per-frequency GEMM path (bf16 spectra: all three stages on the
        tensor cores; fp32 spectra: split-bf16 GEMM between fp32 CUDA-core FFT stages)."""
        return n_tokens >= prefill_token_floor(self.out_blocks, self.in_blocks, self.spectrum_dtype) and self.block_size in (64, 128, 256)

    # -- forward --------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 3:
            bs, sl, in_f = x.shape
            if in_f != self.in_features:
                raise ValueError("in_features mismatch")
            x_flat = x.reshape(bs * sl, in_f)
            lead = (bs, sl)
        elif x.dim() == 2:
            if x.shape[1] != self.in_features:
                raise ValueError("in_features mismatch")
            x_flat, lead = x, None
        else:
            raise ValueError("Unsupported input dim")
        _cabi.require_cuda(x_flat, "input x")
        if x_flat.dtype not in _cabi.DTYPE:
            raise TypeError(f"unsupported activation dtype {x_flat.dtype} (float32, float16, bfloat16)")

        needs_grad = torch.is_grad_enabled() and (
            x_flat.requires_grad or self.c.requires_grad or (self.bias is not None and self.bias.requires_grad))
        if needs_grad:
            from .autograd import bc_linear_autograd
            y_flat = bc_linear_autograd(self, x_flat)
        else:
            # reference rule (src/bench_all_in_one.py:212-219): a cached spectrum is served only in eval/no-grad
            reuse = (not self.training) and (not torch.is_grad_enabled()) and self._cache_spectrum
            if reuse and self._shared_group is not None:
                # siblings reading the same activation (q/k/v, gate/up): one fused forward, this layer's column slice
                y_flat = self._shared_group.member_forward(self, x_flat)
                return y_flat.view(lead[0], lead[1], self.out_features) if lead is not None else y_flat
            layout = _cabi.LAYOUT_GEMM if self._use_prefill_path(x_flat.shape[0]) else _cabi.LAYOUT_STREAM
            spec = self.build_spectrum(force=not reuse, layout=layout)
            y_flat = bc_forward(x_flat, spec, self.bias, self.out_blocks, self.in_blocks, self.block_size, layout=layout,
                                spectrum_dtype=self.spectrum_dtype)
        if lead is not None:
            return y_flat.view(lead[0], lead[1], self.out_features)
        return y_flat


_workspaces: Dict[tuple, torch.Tensor] = {}      # (device index, raw stream) -> staging buffer of the GEMM-layout path
_captured_workspaces: list = []                   # buffers a CUDA graph may still point at: kept alive for the process lifetime
WS_HEADER = 256


def prefill_workspace(device, nbytes: int) -> torch.Tensor:
    """Staging buffer of the GEMM-layout path (status words + Xf/Yf of one token chunk) for the CURRENT stream of `device`;
    grows, never shrinks.  Kernels of one stream run in order, so one buffer per (device, stream) serves every layer; forwards on
    different streams get different buffers and cannot race on Xf / Yf / the status words.  A buffer that was handed out while the
    stream was capturing a CUDA graph is never freed (the graph keeps its address): growing allocates a new one next to it."""
    dev = device if isinstance(device, torch.device) else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, _cabi.stream_ptr(dev))
    capturing = torch.cuda.is_current_stream_capturing()
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None and getattr(ws, "_bcl_captured", False):
            _captured_workspaces.append(ws)
        ws = torch.empty(max(nbytes, WS_HEADER), device=dev, dtype=torch.uint8)
        ws[:WS_HEADER].zero_()                          # status words: [0] per call, [1] sticky (include/bclinear.h)
        _workspaces[key] = ws
    if capturing:
        ws._bcl_captured = True
    return ws


def bc_forward(x_flat: torch.Tensor, spec: torch.Tensor, bias: Optional[torch.Tensor], p: int, q: int, B: int,
               out: Optional[torch.Tensor] = None, *, layout: int = _cabi.LAYOUT_STREAM,
               spectrum_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """One ``bcl_forward`` call: (N,in) -> (N,out).  Output dtype follows the reference: x.dtype, or the PyTorch
    promotion of x.dtype and bias.dtype when a bias is present (reference :266-271)."""
    if x_flat.stride(-1) != 1 or (x_flat.stride(0) * x_flat.element_size()) % 16 or x_flat.data_ptr() % 16:
        x_flat = x_flat.contiguous()
    N = x_flat.shape[0]
    y_dtype = x_flat.dtype if bias is None else torch.promote_types(x_flat.dtype, bias.dtype)
    if y_dtype not in _cabi.DTYPE or (bias is not None and bias.dtype not in _cabi.DTYPE):
        raise TypeError(f"unsupported bias/output dtype {y_dtype}")
    y = out if out is not None else torch.empty(N, p * B, device=x_flat.device, dtype=y_dtype)
    L = _cabi.lib()
    ws_ptr, ws_bytes = None, 0
    if layout == _cabi.LAYOUT_GEMM:                      # an opaque byte buffer: bf16 matrices, or (hi, lo) bf16 pairs for fp32 spectra
        spec_dt = _cabi.DTYPE[spectrum_dtype] if spectrum_dtype is not None else _cabi.BF16
    else:
        spec_dt = _cabi.DTYPE[spec.dtype]
    if layout == _cabi.LAYOUT_GEMM and N > 0:
        ws_bytes = L.bcl_workspace_bytes(N, p, q, B, spec_dt, layout)
        ws = prefill_workspace(x_flat.device, ws_bytes)
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
    with _cabi.on_device(x_flat.device):
        rc = L.bcl_forward(x_flat.data_ptr(), _cabi.DTYPE[x_flat.dtype], N, x_flat.stride(0),
                           spec.data_ptr(), spec_dt, layout, p, q, B,
                           None if bias is None else bias.data_ptr(), 0 if bias is None else _cabi.DTYPE[bias.dtype],
                           y.data_ptr(), _cabi.DTYPE[y_dtype], y.stride(0), ws_ptr, ws_bytes, _cabi.stream_ptr(x_flat.device))
    _cabi.check(rc, "bcl_forward")
    return y


def prefill_status(device, *, sticky: bool = True) -> int:
    """Pipeline status of the GEMM-layout forwards on `device` (0 = no tcgen05 pipeline wait ever timed out).  Synchronises.
    sticky=True: the first stall recorded in ANY workspace of the device since it was allocated (a later healthy call does not
    erase it); sticky=False: the per-call word of the most recent forward on the current stream."""
    dev = device if isinstance(device, torch.device) else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if not sticky:
        ws = _workspaces.get((idx, _cabi.stream_ptr(dev)))
        return 0 if ws is None else int(ws[:4].view(torch.int32).item())
    for (d, _), ws in list(_workspaces.items()) + [((w.device.index, 0), w) for w in _captured_workspaces]:
        if d == idx:
            words = ws[:8].view(torch.int32).tolist()
            if words[1] or words[0]:
                return int(words[1] or words[0])
    return 0


# ---------------------------------------------------------------------------------------------------------
# spectrum cache shims (reference: src/bench_all_in_one.py:164-254)
# ---------------------------------------------------------------------------------------------------------
def enable_cfft_cache(model: nn.Module) -> int:
    """Reference semantics: 'reuse rFFT(c) across eval/no-grad forwards'.  Here the spectrum is always cached (keyed
    on c's version), so this just builds it eagerly.  Returns the number of BC layers, like the reference."""
    n = 0
    for m in model.modules():
        if isinstance(m, BlockCirculantLinear):
            m._cache_spectrum = True
            if m.c.is_cuda:
                m.build_spectrum()
            n += 1
    return n


def disable_cfft_cache(model: nn.Module) -> int:
    """Reference semantics (src/bench_all_in_one.py:239-254): back to the module's own forward, which transforms c on EVERY call
    (src/patch_llama_fft.py:256-257).  Drops the cached spectra and makes every later forward rebuild them, so that writes the
    version counter cannot see (``c.data`` updates) are always picked up.  Returns the number of BC layers."""
    n = 0
    for m in model.modules():
        if isinstance(m, BlockCirculantLinear):
            m.drop_spectrum()
            m._cache_spectrum = False
            n += 1
    return n


# ---------------------------------------------------------------------------------------------------------
# checkpoint I/O (reference :278-340) -- file format kept: flat {"<module path>.c"/".bias": cpu tensor}
# ---------------------------------------------------------------------------------------------------------
def _get_submodule_by_name(model: nn.Module, name: str) -> nn.Module:
    cur = model
    for part in (name.split(".") if name else []):
        cur = getattr(cur, part)
    return cur


def save_bc_params(model: nn.Module, path: str) -> None:
    bc_state: Dict[str, torch.Tensor] = {}
    for module_name, module in model.named_modules():
        if isinstance(module, BlockCirculantLinear):
            bc_state[f"{module_name}.c"] = module.c.detach().cpu()
            if module.bias is not None:
                bc_state[f"{module_name}.bias"] = module.bias.detach().cpu()
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    torch.save(bc_state, path)
    print(f"Saved BC params: {len(bc_state)} tensors -> {path}")


def load_bc_params(model: nn.Module, path: str, *, strict_shapes: bool = True) -> None:
    state: Dict[str, torch.Tensor] = torch.load(path, map_location="cpu")
    loaded = skipped = 0
    for full_name, tensor in state.items():
        if not (full_name.endswith(".c") or full_name.endswith(".bias")):
            skipped += 1
            continue
        module_name, param_name = full_name.rsplit(".", 1)
        try:
            module = _get_submodule_by_name(model, module_name)
        except AttributeError:
            skipped += 1
            continue
        param = getattr(module, param_name, None)
        if not isinstance(param, torch.Tensor) or (strict_shapes and param.shape != tensor.shape):
            skipped += 1
            continue
        with torch.no_grad():
            param.copy_(tensor.to(param.device, dtype=param.dtype))   # bumps c._version -> spectrum rebuilt on next forward
        loaded += 1
    print(f"Loaded BC params from {path}: loaded={loaded}, skipped={skipped}")


# ---------------------------------------------------------------------------------------------------------
# model surgery (reference :343-394)
# ---------------------------------------------------------------------------------------------------------
def _layer_indices(n_layers: int, num_layers_to_patch: int, patch_position: str):
    if num_layers_to_patch < 0:
        raise ValueError("num_layers_to_patch must be >= 0")
    k = min(num_layers_to_patch, n_layers)
    if patch_position not in ("first", "last"):
        raise ValueError("patch_position must be 'first' or 'last'")
    return list(range(k)) if patch_position == "first" else list(range(n_layers - k, n_layers))


def _swap(parent: nn.Module, names: Iterable[str], block_size: int, spectrum_dtype: torch.dtype, verbose: bool) -> int:
    n = 0
    for name in names:
        old = getattr(parent, name, None)
        if not isinstance(old, nn.Linear):
            if verbose:
                print(f"    Skip {name}: not an nn.Linear")
            continue
        if verbose:
            print(f"    Replacing {name}: in_features={old.in_features}, out_features={old.out_features}")
        setattr(parent, name, BlockCirculantLinear.from_linear(old, block_size=block_size, spectrum_dtype=spectrum_dtype))
        n += 1
    return n


def patch_mlp_with_block_circulant(model, *, num_layers_to_patch: int = NUM_LAYERS_TO_PATCH, block_size: int = BLOCK_SIZE,
                                   patch_position: str = "first", targets: Sequence[str] = MLP_TARGETS,
                                   spectrum_dtype: torch.dtype = torch.float32, verbose: bool = True) -> None:
    """Replace ``mlp.{gate,up,down}_proj`` of the first/last k decoder layers in place (reference :343-394).
    ``targets`` / ``spectrum_dtype`` / ``verbose`` are keyword-only extensions with reference-compatible defaults."""
    n_layers = model.config.num_hidden_layers
    if verbose:
        print(f"Model has {n_layers} transformer layers.")
    idx = _layer_indices(n_layers, num_layers_to_patch, patch_position)
    if verbose:
        print(f"Patching {len(idx)} layer(s) from {patch_position}: {idx}")
    for li in idx:
        if verbose:
            print(f"  Patching layer {li} MLP...")
        _swap(model.model.layers[li].mlp, targets, block_size, spectrum_dtype, verbose)


def patch_llama_fft(model, *, num_layers_to_patch: Optional[int] = None, block_size: int = BLOCK_SIZE,
                    patch_position: str = "first", spectrum_dtype: torch.dtype = torch.float32,
                    verbose: bool = False) -> int:
    """All seven projections (self_attn.{q,k,v,o}_proj + mlp.{gate,up,down}_proj) of the selected decoder layers --
    the configuration BASELINE.json's configs name; the reference itself only ships the MLP subset.
    Returns the number of replaced linears."""
    n_layers = model.config.num_hidden_layers
    idx = _layer_indices(n_layers, n_layers if num_layers_to_patch is None else num_layers_to_patch, patch_position)
    n = 0
    for li in idx:
        layer = model.model.layers[li]
        n += _swap(layer.self_attn, ATTN_TARGETS, block_size, spectrum_dtype, verbose)
        n += _swap(layer.mlp, MLP_TARGETS, block_size, spectrum_dtype, verbose)
    return n
