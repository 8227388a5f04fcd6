"""Differentiable BC linear for calibration (reference: autograd through src/patch_llama_fft.py:226-275, used by
src/calibrate_bc.py:281-294).  Placeholder until the backward kernels land: fails loudly, never falls back."""
import torch


def bc_linear_autograd(layer, x_flat: torch.Tensor) -> torch.Tensor:
    raise NotImplementedError(
        "BlockCirculantLinear backward kernels are not built yet (SURVEY.md section 8f item 2); "
        "run the forward under torch.no_grad() / inference_mode().")
