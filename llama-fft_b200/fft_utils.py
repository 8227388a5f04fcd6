"""Drop-in for the reference's ``src/fft_utils.py``: the single-block primitive that fixes the convention
C[i,j] = c[(i-j) % n].  Runs the same sm_100a kernels as the layer (a 1x1 grid of blocks)."""
import torch

from . import _cabi
from .patch_llama_fft import bc_forward


def circulant_matvec_fft(c: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """y = C(c) x for one length-n circulant (reference src/fft_utils.py:3-34): fp32 internally, result cast
    back to x.dtype.  n must be a block size the kernels support (64, 128, 256); tensors must be on a CUDA device."""
    assert c.dim() == 1 and x.dim() == 1
    n = c.shape[0]
    assert x.shape[0] == n
    _cabi.require_cuda(x, "x")
    _cabi.require_cuda(c, "c")
    if n not in (64, 128, 256):
        raise ValueError(f"block size {n} not supported by the CUDA kernels (64, 128, 256)")
    c32 = c.detach().to(torch.float32).contiguous().view(1, 1, n)
    spec = torch.empty_like(c32)
    with torch.cuda.device(c.device):
        rc = _cabi.lib().bcl_precompute_spectrum(c32.data_ptr(), 1, 1, n, spec.data_ptr(), _cabi.F32, _cabi.LAYOUT_STREAM,
                                                 _cabi.stream_ptr(c.device))
    _cabi.check(rc, "bcl_precompute_spectrum")
    return bc_forward(x.detach().view(1, n), spec, None, 1, 1, n).view(n)
