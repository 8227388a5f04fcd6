"""GPU tests (-m gpu) of the prefill path: bf16 GEMM-layout spectrum + rfft rows -> tcgen05 per-frequency GEMM -> irfft rows
(llama-fft_b200/csrc/bcl_gemm.cu), called through the C ABI.  Each stage is checked against the reference step it
replaces (src/patch_llama_fft.py:252 / :260 / :263-271) and the chained forward against the oracle.

Tolerance (BASELINE.json north_star): bf16 spectra -> max|d| <= 1e-2 * max|y_ref| (+ 2 ulp of a 16-bit I/O dtype)."""
import numpy as np
import pytest
import torch

import llama_fft_b200 as bc
from llama_fft_b200 import _cabi
from llama_fft_b200.patch_llama_fft import bc_forward, prefill_status
from oracle import bc_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF = torch.bfloat16


def make_layer(c, bias=None):
    p, q, B = c.shape
    m = bc.BlockCirculantLinear(q * B, p * B, block_size=B, bias=bias is not None, spectrum_dtype=BF)
    with torch.no_grad():
        m.c.copy_(c)
        if bias is not None:
            m.bias.data = bias.clone()
    return m.to(DEV).eval()


def assert_close(y, y_ref, tol, io_dtype=torch.float32, mag=None):
    """mag: magnitude the 16-bit rounding acted on when it is not |y_ref| itself (y rounded BEFORE a bias is added)."""
    assert y.shape == y_ref.shape
    yd, rd = y.double().cpu(), y_ref.double().cpu()
    scale = max(rd.abs().max().item(), 1e-30)
    mg = rd.abs() if mag is None else torch.maximum(rd.abs(), mag.double().cpu())
    bound = tol * scale + (2 * G.ulp(io_dtype) * mg if io_dtype != torch.float32 else 0.0)
    bad = (yd - rd).abs() > bound
    assert not bool(torch.as_tensor(bad).any()), f"max rel err {(yd - rd).abs().max().item() / scale:.3e} (tol {tol})"


def geom(p, q, B):
    F, Kp, Np = _cabi.prefill_geometry(p, q, B)
    return F, Kp, Np, Kp // 8, F // 4


def gemm_spec(m):
    """(whole GEMM-layout buffer, its Cmat part viewed as (F, Np, Kp) bf16)."""
    buf = m.build_spectrum(layout=_cabi.LAYOUT_GEMM)
    F, Kp, Np, _, _ = geom(m.out_blocks, m.in_blocks, m.block_size)
    return buf, buf[:F * Np * Kp * 2].view(BF).view(F, Np, Kp)


def ref_cmat(c):
    """The GEMM-layout spectrum from torch.fft.rfft(c) (reference :256-257), fp64: [F][Np][Kp]."""
    p, q, B = c.shape
    F, Kp, Np, _, _ = geom(p, q, B)
    Cf = torch.fft.rfft(c.double(), dim=-1)                      # (p,q,B/2+1)
    M = torch.zeros(F, Np, Kp, dtype=torch.float64)
    cr, ci = Cf.real[..., :F].permute(2, 0, 1), Cf.imag[..., :F].permute(2, 0, 1)     # (F,p,q)
    M[:, 0:2 * p:2, 0:2 * q:2] = cr
    M[:, 0:2 * p:2, 1:2 * q:2] = -ci
    M[:, 1:2 * p:2, 0:2 * q:2] = ci
    M[:, 1:2 * p:2, 1:2 * q:2] = cr
    M[0] = 0
    M[0, 0:2 * p:2, 0:2 * q:2] = Cf.real[..., 0]
    M[0, 1:2 * p:2, 1:2 * q:2] = Cf.real[..., F]              # Nyquist
    return M


def xf_natural(Xf_chunks):
    """stage layout [F][KG][N][8] -> [F][N][Kp]"""
    F, KG, N, _ = Xf_chunks.shape
    return Xf_chunks.permute(0, 2, 1, 3).reshape(F, N, KG * 8)


def xf_chunks(Xf_nat):
    F, N, Kp = Xf_nat.shape
    return Xf_nat.reshape(F, N, Kp // 8, 8).permute(0, 2, 1, 3).contiguous()


def yf_natural(Yf_chunks):
    """stage layout [FG][p][N][8] (8 = 4 frequencies x {re,im}) -> [F][N][2p]"""
    FG, p, N, _ = Yf_chunks.shape
    return Yf_chunks.reshape(FG, p, N, 4, 2).permute(0, 3, 2, 1, 4).reshape(FG * 4, N, 2 * p)


def yf_chunks(Yf_nat, p):
    F, N, _ = Yf_nat.shape
    return Yf_nat[:, :, :2 * p].reshape(F // 4, 4, N, p, 2).permute(0, 3, 2, 1, 4).reshape(F // 4, p, N, 8).contiguous()


def ref_xf(x, q, B, Kp):
    """rfft(x) (reference :252), natural layout [F][N][Kp], fp64."""
    N = x.shape[0]
    F = B // 2
    Xf = torch.fft.rfft(x.double().view(N, q, B), dim=-1)       # (N,q,F+1)
    out = torch.zeros(F, N, Kp, dtype=torch.float64)
    out[:, :, 0:2 * q:2] = Xf.real[..., :F].permute(2, 0, 1)
    out[:, :, 1:2 * q:2] = Xf.imag[..., :F].permute(2, 0, 1)
    out[0, :, 1:2 * q:2] = Xf.real[..., F]
    return out


def pitch(N):
    return (N + 31) // 32 * 32                                  # token pitch of the stage layouts (include/bclinear.h)


def pad_tokens(t, N):
    """(.., .., N, 8) -> (.., .., pitch(N), 8), zero padded"""
    out = torch.zeros(t.shape[0], t.shape[1], pitch(N), 8, dtype=t.dtype, device=t.device)
    out[:, :, :N] = t
    return out


def new_status():
    return torch.zeros(1, device=DEV, dtype=torch.int32)


@pytest.mark.parametrize("p,q,B", [(3, 5, 64), (32, 32, 128), (86, 32, 128), (4, 9, 256)])
def test_gemm_layout_spectrum_matches_rfft(p, q, B):
    g = torch.Generator().manual_seed(p * 100 + q + B)
    c = torch.randn(p, q, B, generator=g)
    m = make_layer(c)
    buf, spec = gemm_spec(m)
    assert buf.numel() == _cabi.lib().bcl_spectrum_bytes(p, q, B, _cabi.BF16, _cabi.LAYOUT_GEMM)
    ref = ref_cmat(c)
    got = spec.double().cpu()
    assert (got - ref).abs().max() <= 2.0 ** -8 * ref.abs().max()
    assert (got[:, 2 * p:, :] == 0).all() and (got[:, :, 2 * q:] == 0).all()         # zero padding is what the MMA multiplies
    # the DFT matrices behind Cmat: forward rows = packed spectrum reals, inverse = its scaled transpose
    o16, obf, oinv = _cabi.prefill_dft_offsets(p, q, B)
    wf = buf[o16:o16 + B * B * 2].view(torch.float16).view(B, B).double().cpu()
    wb = buf[obf:obf + B * B * 2].view(BF).view(B, B).double().cpu()
    wi = buf[oinv:oinv + B * B * 2].view(BF).view(B, B).double().cpu()
    eye = torch.eye(B, dtype=torch.float64)
    Fm = torch.fft.rfft(eye, dim=-1)                              # row t: spectrum of e_t
    Wref = torch.zeros(B, B, dtype=torch.float64)
    Wref[0], Wref[1] = Fm.real[:, 0], Fm.real[:, B // 2]
    Wref[2::2], Wref[3::2] = Fm.real[:, 1:B // 2].T, Fm.imag[:, 1:B // 2].T
    assert (wf - Wref).abs().max() <= 2.0 ** -11 and (wb - Wref).abs().max() <= 2.0 ** -8
    assert ((wi @ Wref) - eye).abs().max() <= 2e-2                # Winv * Wfwd = I up to bf16 rounding of both


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("q,B,N,dtype", [(5, 64, 7, torch.float32), (32, 128, 200, torch.float16), (86, 128, 33, torch.bfloat16),
                                         (9, 256, 130, torch.float32), (5, 64, 300, torch.float16), (32, 128, 1024, torch.bfloat16),
                                         (9, 256, 333, torch.float16), (32, 256, 200, torch.bfloat16)])
def test_stage_rfft(q, B, N, dtype, impl):
    if impl == 2 and dtype == torch.float32:
        pytest.skip("tensor-core DFT covers 16-bit activations")
    g = torch.Generator().manual_seed(q + B + N)
    x = torch.randn(N, q * B, generator=g).to(dtype)
    p = 3
    m = make_layer(torch.randn(p, q, B, generator=g))
    buf, _ = gemm_spec(m)
    F, Kp, Np, KG, FG = geom(p, q, B)
    Xf = torch.full((F, KG, pitch(N), 8), float("nan"), device=DEV, dtype=BF)
    xd = x.to(DEV)
    st = new_status()
    rc = _cabi.lib().bcl_prefill_rfft(xd.data_ptr(), _cabi.DTYPE[dtype], N, xd.stride(0), p, q, B, buf.data_ptr(), impl,
                                      Xf.data_ptr(), st.data_ptr(), _cabi.stream_ptr(DEV))
    _cabi.check(rc, "bcl_prefill_rfft")
    torch.cuda.synchronize()
    assert st.item() == 0
    ref = ref_xf(x, q, B, Kp)
    got = xf_natural(Xf[:, :, :N]).double().cpu()
    assert torch.isfinite(got).all()
    # impl 2 multiplies by a 16-bit DFT matrix (fp16: 2^-11, bf16: 2^-8 per coefficient) before the bf16 rounding of Xf
    tol = 2.0 ** -8 if impl == 1 else (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -8 + 2.0 ** -10)
    assert (got - ref).abs().max() <= tol * ref.abs().max()


def test_stage_rfft_tc_rejects_unsupported():
    x = torch.zeros(8, 512, device=DEV)
    Xf = torch.zeros(128, 4, 8, 8, device=DEV, dtype=BF)
    rc = _cabi.lib().bcl_prefill_rfft(x.data_ptr(), _cabi.F32, 8, 512, 2, 2, 256, None, 2, Xf.data_ptr(), new_status().data_ptr(), _cabi.stream_ptr(DEV))
    assert rc == -8                                                # fp32 activations: no tensor-core DFT (NULL spectrum buffer too)


@pytest.mark.parametrize("p,q,B,N", [(32, 32, 128, 128), (32, 32, 128, 300), (86, 32, 128, 256), (32, 86, 128, 129), (3, 5, 64, 40),
                                     (16, 16, 256, 512), (172, 8, 64, 128), (224, 4, 64, 260)])
def test_stage_tcgen05_gemm(p, q, B, N):
    """Yf[f] = Xf[f] @ Cmat[f]^T per frequency, bf16 operands, fp32 accumulate (the einsum of reference :260)."""
    g = torch.Generator().manual_seed(p + q + B + N)
    F, Kp, Np, KG, FG = geom(p, q, B)
    Xf = torch.zeros(F, N, Kp)
    Xf[:, :, :2 * q] = torch.randn(F, N, 2 * q, generator=g)
    Cm = torch.zeros(F, Np, Kp)
    Cm[:, :2 * p, :2 * q] = torch.randn(F, 2 * p, 2 * q, generator=g)
    Xd, Cd = pad_tokens(xf_chunks(Xf), N).to(DEV, BF), Cm.to(DEV, BF)
    Yf = torch.full((FG, p, pitch(N), 8), float("nan"), device=DEV, dtype=BF)
    status = new_status()
    rc = _cabi.lib().bcl_prefill_gemm(Xd.data_ptr(), Cd.data_ptr(), Yf.data_ptr(), N, p, q, B, status.data_ptr(), _cabi.stream_ptr(DEV))
    _cabi.check(rc, "bcl_prefill_gemm")
    torch.cuda.synchronize()
    assert status.item() == 0, f"tcgen05 pipeline stalled (status {status.item()})"
    ref = torch.einsum("fnk,fmk->fnm", xf_natural(Xd[:, :, :N]).double().cpu(), Cd.double().cpu())[:, :, :2 * p]
    got = yf_natural(Yf[:, :, :N]).double().cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs()
    assert (err <= 2.0 ** -8 * ref.abs() + 1e-3 * ref.abs().max()).all(), f"max err {err.max().item():.3e} of {ref.abs().max().item():.3e}"


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("p,B,N,dtype,bias_dt", [(5, 64, 7, torch.float32, None), (32, 128, 200, torch.float16, torch.float16),
                                                 (86, 128, 33, torch.bfloat16, None), (9, 256, 130, torch.float16, torch.float32),
                                                 (5, 64, 300, torch.bfloat16, torch.bfloat16), (32, 128, 1024, torch.float16, None),
                                                 (9, 256, 333, torch.float16, torch.float16), (32, 256, 200, torch.bfloat16, None)])
def test_stage_irfft(p, B, N, dtype, bias_dt, impl):
    if impl == 2 and (dtype == torch.float32 or (bias_dt is not None and bias_dt != dtype)):
        pytest.skip("tensor-core inverse DFT covers 16-bit activations with output dtype == activation dtype")
    g = torch.Generator().manual_seed(p + B + N)
    q = 2
    m = make_layer(torch.randn(p, q, B, generator=g))
    buf, _ = gemm_spec(m)
    F, Kp, Np, KG, FG = geom(p, q, B)
    Yc = torch.fft.rfft(torch.randn(N, p, B, generator=g).double(), dim=-1)            # a valid half-spectrum per block
    Yn = torch.zeros(F, N, 2 * p)
    Yn[:, :, 0::2] = Yc.real[..., :F].permute(2, 0, 1).float()
    Yn[:, :, 1::2] = Yc.imag[..., :F].permute(2, 0, 1).float()
    Yn[0, :, 1::2] = Yc.real[..., F].float()
    Yd = pad_tokens(yf_chunks(Yn, p), N).to(DEV, BF)
    # reference on the SAME bf16-rounded spectrum: irfft (:263), round to x.dtype (:266), + bias in the promoted dtype (:270-271)
    Yb = yf_natural(Yd[:, :, :N]).double().cpu()
    spec = torch.zeros(N, p, F + 1, dtype=torch.complex128)
    spec.real[..., :F] = Yb[:, :, 0::2].permute(1, 2, 0)
    spec.imag[..., 1:F] = Yb[1:, :, 1::2].permute(1, 2, 0)
    spec.real[..., F] = Yb[0, :, 1::2]
    y32 = torch.fft.irfft(spec, n=B, dim=-1).float().reshape(N, p * B)
    bias = None if bias_dt is None else torch.randn(p * B, generator=g).to(bias_dt)
    y_ref = y32.to(dtype)
    if bias is not None:
        y_ref = y_ref + bias
    y = torch.empty(N, p * B, device=DEV, dtype=y_ref.dtype)
    bd = None if bias is None else bias.to(DEV)
    st = new_status()
    rc = _cabi.lib().bcl_prefill_irfft(Yd.data_ptr(), N, p, q, B, buf.data_ptr(), impl, _cabi.DTYPE[dtype], None if bd is None else bd.data_ptr(),
                                       0 if bd is None else _cabi.DTYPE[bias_dt], y.data_ptr(), _cabi.DTYPE[y_ref.dtype], y.stride(0),
                                       st.data_ptr(), _cabi.stream_ptr(DEV))
    _cabi.check(rc, "bcl_prefill_irfft")
    torch.cuda.synchronize()
    assert st.item() == 0
    tol = 1e-5 if impl == 1 else 6e-3                          # impl 2: bf16 inverse DFT matrix (2^-9 per coefficient, B terms)
    assert_close(y, y_ref, tol, dtype, mag=y32.abs() + (0 if bias is None else bias.float().abs()))


SHAPES = [(4096, 4096, 128), (4096, 11008, 128), (11008, 4096, 128), (4096, 4096, 64), (4096, 4096, 256), (1024, 512, 64)]


@pytest.mark.parametrize("inf,outf,B", SHAPES)
@pytest.mark.parametrize("N", [256, 333])
def test_prefill_forward_vs_oracle(inf, outf, B, N):
    g = torch.Generator().manual_seed(inf + outf + B + N)
    c = torch.randn(outf // B, inf // B, B, generator=g) * 0.01
    x = torch.randn(N, inf, generator=g)
    m = make_layer(c)
    assert m._use_prefill_path(N)
    with torch.no_grad():
        y = m(x.to(DEV))
    assert prefill_status(DEV) == 0
    assert_close(y, O.bc_forward_torch(c, x), 1e-2)


@pytest.mark.parametrize("dtype,bias_dt", [(torch.float16, None), (torch.float16, torch.float16), (torch.float16, torch.float32), (torch.bfloat16, torch.bfloat16)])
def test_prefill_forward_3d_16bit_io_and_bias(dtype, bias_dt):
    g = torch.Generator().manual_seed(11)
    c = torch.randn(32, 32, 128, generator=g) * 0.01
    bias = None if bias_dt is None else (torch.randn(4096, generator=g) * 0.1).to(bias_dt)
    x = torch.randn(4, 600, 4096, generator=g).to(dtype)
    m = make_layer(c, bias)
    with torch.no_grad():
        y = m(x.to(DEV))
    y_ref = O.bc_forward_torch(c, x, bias)
    assert y.shape == (4, 600, 4096) and y.dtype == y_ref.dtype
    assert prefill_status(DEV) == 0
    mag = None if bias is None else O.bc_forward_torch(c, x).float().abs() + bias.float().abs()
    assert_close(y, y_ref, 1e-2, dtype, mag=mag)


def test_prefill_matches_stream_kernel_with_same_spectrum_dtype():
    """Both paths hold bf16 spectra; they differ only by the bf16 rounding of Xf/Yf, so they agree far inside 1e-2."""
    g = torch.Generator().manual_seed(3)
    c = torch.randn(86, 32, 128, generator=g) * 0.01
    x = torch.randn(512, 4096, generator=g)
    m = make_layer(c)
    with torch.no_grad():
        y_gemm = m(x.to(DEV))
        spec = m.build_spectrum()
        y_stream = bc_forward(x.to(DEV), spec, None, 86, 32, 128)
    assert_close(y_gemm, y_stream, 1e-2)


def test_prefill_chunking_small_workspace_and_full_size_properties():
    """BASELINE configs[2] size (N = 8 x 2048) through a workspace that forces several chunks + a ragged tail; checked by
    linearity and by agreement of a row subset with the oracle."""
    g = torch.Generator().manual_seed(9)
    p = q = 32
    B = 128
    c = torch.randn(p, q, B, generator=g) * 0.01
    m = make_layer(c)
    spec = m.build_spectrum(layout=_cabi.LAYOUT_GEMM)
    N = 16384 + 77
    x1 = torch.randn(N, 4096, generator=g).to(DEV, torch.float16)
    x2 = torch.randn(N, 4096, generator=g).to(DEV, torch.float16)
    L = _cabi.lib()
    per_tok = (B // 2) * (64 + 64) * 2
    ws = torch.empty(256 + 512 + 1000 * per_tok, device=DEV, dtype=torch.uint8)           # 896-token chunks
    def run(x):
        y = torch.empty(N, 4096, device=DEV, dtype=torch.float16)
        rc = L.bcl_forward(x.data_ptr(), _cabi.F16, N, 4096, spec.data_ptr(), _cabi.BF16, _cabi.LAYOUT_GEMM, p, q, B, None, 0,
                           y.data_ptr(), _cabi.F16, 4096, ws.data_ptr(), ws.numel(), _cabi.stream_ptr(DEV))
        _cabi.check(rc, "bcl_forward")
        return y
    y1, y2, y12 = run(x1), run(x2), run((x1.float() + x2.float()).half())
    assert int(ws[:4].view(torch.int32).item()) == 0
    lin = (y12.float() - (y1.float() + y2.float())).abs().max().item()
    assert lin <= 2e-2 * y12.float().abs().max().item()
    rows = torch.unique(torch.cat([torch.tensor([0, 127, 128, 895, 896, 897, 8000, 16383, 16384, N - 1]),
                                   torch.randint(0, N, (256,), generator=torch.Generator().manual_seed(N))]))      # tile / chunk edges + 256 random rows
    y_ref = O.bc_forward_torch(c, x1[rows].cpu())
    assert_close(y1[rows], y_ref, 1e-2, torch.float16)
    with torch.no_grad():
        assert torch.equal(m(x1), m(x1))                                                     # deterministic


def test_gemm_layout_rejects_bad_spectrum_dtype_and_tiny_workspace():
    L = _cabi.lib()
    assert L.bcl_spectrum_bytes(4, 4, 128, _cabi.F16, _cabi.LAYOUT_GEMM) == 0
    # fp32 = the split layout: Cmat and the three DFT matrices twice (hi, lo)
    assert (L.bcl_spectrum_bytes(4, 4, 128, _cabi.F32, _cabi.LAYOUT_GEMM) - L.bcl_spectrum_bytes(4, 4, 128, _cabi.BF16, _cabi.LAYOUT_GEMM)
            == 64 * 64 * 16 * 2 + 3 * 128 * 128 * 2)
    x = torch.zeros(256, 512, device=DEV)
    spec = torch.zeros(L.bcl_spectrum_bytes(4, 4, 128, _cabi.BF16, _cabi.LAYOUT_GEMM), device=DEV, dtype=torch.uint8)
    y = torch.zeros(256, 512, device=DEV)
    ws = torch.zeros(1024, device=DEV, dtype=torch.uint8)
    rc = L.bcl_forward(x.data_ptr(), _cabi.F32, 256, 512, spec.data_ptr(), _cabi.F16, _cabi.LAYOUT_GEMM, 4, 4, 128, None, 0,
                       y.data_ptr(), _cabi.F32, 512, ws.data_ptr(), ws.numel(), _cabi.stream_ptr(DEV))
    assert rc == -5
    rc = L.bcl_forward(x.data_ptr(), _cabi.F32, 256, 512, spec.data_ptr(), _cabi.BF16, _cabi.LAYOUT_GEMM, 4, 4, 128, None, 0,
                       y.data_ptr(), _cabi.F32, 512, ws.data_ptr(), ws.numel(), _cabi.stream_ptr(DEV))
    assert rc == -6


@pytest.mark.parametrize("N,dtype,sdt", [(1, torch.float16, torch.float32), (5, torch.float32, torch.bfloat16), (300, torch.float16, BF), (300, torch.float32, BF)])
def test_forward_bcast_writes_every_destination(N, dtype, sdt):
    """bcl_forward_bcast (the fused gather of sharded layers): one forward, identical results in all destinations --
    stream kernel (N small), tensor-core path (16-bit I/O falls back to plain stores for peers) and CUDA-core stage 3."""
    from llama_fft_b200.sharded import bc_forward_bcast
    g = torch.Generator().manual_seed(N)
    c = torch.randn(6, 4, 128, generator=g) * 0.05
    m = bc.BlockCirculantLinear(512, 768, block_size=128, bias=True, spectrum_dtype=sdt)
    with torch.no_grad():
        m.c.copy_(c)
        m.bias.normal_(0, 0.1, generator=g)
    m = m.to(DEV).eval()
    x = torch.randn(N, 512, generator=g).to(dtype).to(DEV)
    with torch.no_grad():
        y_ref = m(x)
    big = [torch.full((N, 2048), 7.0, device=DEV, dtype=y_ref.dtype) for _ in range(3)]       # slices of wider buffers, like peers' y
    dsts = [b[:, 256:1024] for b in big]
    bc_forward_bcast(m, x, dsts)
    torch.cuda.synchronize()
    for b, d in zip(big, dsts):
        assert_close(d, y_ref, 1e-2 if (sdt == BF and N >= 256) else 1e-6, dtype)
        assert (b[:, :256] == 7.0).all() and (b[:, 1024:] == 7.0).all()


_FUSED_CHILD = r"""
import sys, torch
import llama_fft_b200 as bc
from llama_fft_b200.patch_llama_fft import prefill_status
inf, outf, B, N, dt, has_bias, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], int(sys.argv[6]), sys.argv[7]
dtype = {"f16": torch.float16, "bf16": torch.bfloat16}[dt]
g = torch.Generator().manual_seed(inf + outf + B + N)
c = torch.randn(outf // B, inf // B, B, generator=g) * 0.01
x = torch.randn(N, inf, generator=g).to(dtype)
m = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=bool(has_bias), spectrum_dtype=torch.bfloat16)
with torch.no_grad():
    m.c.copy_(c)
    if has_bias:
        m.bias.copy_(torch.randn(outf, generator=g) * 0.1)
m = m.to("cuda:0").eval()
if has_bias:
    m.bias.data = m.bias.data.to(dtype)               # same dtype as the activations: the tensor-core inverse stage adds it
with torch.no_grad():
    y = m(x.to("cuda:0"))
    y2 = m(x.to("cuda:0"))
torch.cuda.synchronize()
assert prefill_status("cuda:0") == 0, prefill_status("cuda:0")
assert torch.equal(y, y2)
torch.save(y.cpu(), out)
"""


@pytest.mark.parametrize("inf,outf,B,N,dt,has_bias", [(4096, 4096, 128, 2048 + 77, "f16", 0), (4096, 11008, 128, 1500, "bf16", 1),
                                                      (11008, 4096, 128, 1024, "f16", 0), (1024, 512, 64, 3000, "f16", 1)])
def test_prefill_fused_single_launch_matches_three_launch_chain(tmp_path, inf, outf, B, N, dt, has_bias):
    """The opt-in single-launch pipeline (bcl_prefill_fused.cu, BCL_FUSED_MIN_TILES) runs the same MMAs on the same tiles as the
    three stage launches, only handed over through per-tile counters: its output must be BIT-identical to the chain's, and
    within the bf16-spectrum tolerance of the oracle.  The switch is read once per process, hence the child process."""
    import os
    import subprocess
    import sys
    outs = {}
    for tag, tiles in (("chain", "0"), ("fused", "4")):
        out = str(tmp_path / f"{tag}.pt")
        env = dict(os.environ, BCL_FUSED_MIN_TILES=tiles, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        r = subprocess.run([sys.executable, "-c", _FUSED_CHILD, str(inf), str(outf), str(B), str(N), dt, str(has_bias), out],
                           env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = torch.load(out)
    assert torch.equal(outs["chain"], outs["fused"])
    dtype = {"f16": torch.float16, "bf16": torch.bfloat16}[dt]
    g = torch.Generator().manual_seed(inf + outf + B + N)
    c = torch.randn(outf // B, inf // B, B, generator=g) * 0.01
    x = torch.randn(N, inf, generator=g).to(dtype)
    rows = torch.tensor([0, 1, 127, 128, N // 2, N - 2, N - 1])
    if not has_bias:
        assert_close(outs["fused"][rows], O.bc_forward_torch(c, x[rows].float()), 1e-2, dtype)


@pytest.mark.parametrize("inf,outf,B,N,dt,has_bias,lag,ring", [(4096, 4096, 128, 6144 + 77, "f16", 0, 0, 0), (4096, 4096, 128, 4096, "bf16", 1, 6, 14),
                                                               (4096, 11008, 128, 1500, "bf16", 1, 3, 8), (11008, 4096, 128, 1024, "f16", 0, 2, 6),
                                                               (1024, 512, 64, 3000, "f16", 1, 4, 9)])
def test_prefill_wave_single_launch_matches_three_launch_chain(monkeypatch, inf, outf, B, N, dt, has_bias, lag, ring):
    """The opt-in wave-schedule kernel (bcl_prefill_wave.cu, BCL_PREFILL_WAVE=1): every CTA interleaves forward-DFT / GEMM / inverse-DFT
    items `lag` tiles apart over rings of `ring` tiles of Xf / Yf.  Same MMAs on the same tiles as the three stage launches: the output
    must be BIT-identical, also when ring slots are reused (ring < tiles) and with a ragged last tile.  The switch is read per call."""
    dtype = {"f16": torch.float16, "bf16": torch.bfloat16}[dt]
    g = torch.Generator().manual_seed(inf + outf + B + N)
    c = torch.randn(outf // B, inf // B, B, generator=g) * 0.01
    x = torch.randn(N, inf, generator=g).to(dtype).to(DEV)
    m = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=bool(has_bias), spectrum_dtype=torch.bfloat16)
    with torch.no_grad():
        m.c.copy_(c)
        if has_bias:
            m.bias.copy_(torch.randn(outf, generator=g) * 0.1)
    m = m.to(DEV).eval()
    if has_bias:
        m.bias.data = m.bias.data.to(dtype)
    monkeypatch.setenv("BCL_PREFILL_WAVE", "0")
    with torch.no_grad():
        y_chain = m(x).clone()
    monkeypatch.setenv("BCL_PREFILL_WAVE", "1")
    monkeypatch.setenv("BCL_PREFILL_WAVE_MIN_TILES", "1")
    if lag:
        monkeypatch.setenv("BCL_WAVE_LAG", str(lag))
        monkeypatch.setenv("BCL_WAVE_RING", str(ring))
    with torch.no_grad():
        y_wave = m(x).clone()
        y_wave2 = m(x).clone()
    torch.cuda.synchronize()
    assert prefill_status(DEV) == 0
    assert torch.equal(y_wave, y_chain)
    assert torch.equal(y_wave2, y_chain)


# ---------------------------------------------------------------------------------------------------------
# fp32 spectra on the tensor cores: split (hi, lo) bf16 operands, three MMAs per k-step, fp32 CUDA-core FFT stages
# (north star: max relative error <= 1e-4 for fp32 spectra)
# ---------------------------------------------------------------------------------------------------------
def make_layer_f32(c, bias=None):
    p, q, B = c.shape
    m = bc.BlockCirculantLinear(q * B, p * B, block_size=B, bias=bias is not None, spectrum_dtype=torch.float32)
    with torch.no_grad():
        m.c.copy_(c)
        if bias is not None:
            m.bias.data = bias.clone()
    return m.to(DEV).eval()


@pytest.mark.parametrize("inf,outf,B", SHAPES + [(8192, 1024, 256)])
@pytest.mark.parametrize("N", [256, 333])
def test_split_prefill_forward_fp32_class_vs_oracle(inf, outf, B, N):
    g = torch.Generator().manual_seed(inf + outf + B + N + 1)
    c = torch.randn(outf // B, inf // B, B, generator=g) * 0.01
    x = torch.randn(N, inf, generator=g)
    m = make_layer_f32(c)
    assert m._use_prefill_path(N)
    with torch.no_grad():
        y = m(x.to(DEV))
    assert prefill_status(DEV) == 0
    assert_close(y, O.bc_forward_torch(c, x), 1e-4)


def test_split_prefill_matches_fp64_dense_oracle_and_stream_kernel():
    g = torch.Generator().manual_seed(77)
    p, q, B, N = 5, 9, 64, 300
    c = torch.randn(p, q, B, generator=g) * 0.05
    bias = torch.randn(p * B, generator=g) * 0.1
    x = torch.randn(N, q * B, generator=g)
    m = make_layer_f32(c, bias)
    with torch.no_grad():
        y = m(x.to(DEV))
        ys = bc_forward(x.to(DEV), m.build_spectrum(), m.bias, p, q, B)              # the CUDA-core stream kernel, fp32 spectrum
    y_ref = torch.from_numpy(O.bc_forward_dense64(c, x, bias))
    assert_close(y, y_ref, 1e-4)
    assert_close(y, ys, 1e-4)


@pytest.mark.parametrize("dtype,bias_dt", [(torch.float16, None), (torch.float16, torch.float32), (torch.bfloat16, torch.bfloat16)])
def test_split_prefill_16bit_io_3d_and_bias(dtype, bias_dt):
    g = torch.Generator().manual_seed(5)
    p, q, B = 8, 16, 128
    c = torch.randn(p, q, B, generator=g) * 0.01
    bias = None if bias_dt is None else (torch.randn(p * B, generator=g) * 0.1).to(bias_dt)
    x = torch.randn(3, 100, q * B, generator=g).to(dtype)
    m = make_layer_f32(c, bias)
    with torch.no_grad():
        y = m(x.to(DEV))
    y_ref = O.bc_forward_torch(c, x, bias)
    assert y.shape == (3, 100, p * B) and y.dtype == y_ref.dtype
    y32 = O.bc_forward_torch(c, x.float())
    assert_close(y, y_ref, 1e-4, dtype, mag=y32.abs() + (0 if bias is None else bias.float().abs()))


def test_split_prefill_chunked_workspace_full_size():
    """N = 8 x 2048 + ragged tail through a workspace that forces several chunks; row subset against the oracle at 1e-4."""
    g = torch.Generator().manual_seed(10)
    p = q = 32
    B = 128
    c = torch.randn(p, q, B, generator=g) * 0.01
    m = make_layer_f32(c)
    spec = m.build_spectrum(layout=_cabi.LAYOUT_GEMM)
    N = 16384 + 77
    x = torch.randn(N, 4096, generator=g).to(DEV)
    L = _cabi.lib()
    full = L.bcl_workspace_bytes(N, p, q, B, _cabi.F32, _cabi.LAYOUT_GEMM)
    per_tok = (B // 2) * (64 + 64) * 2 * 2
    ws = torch.empty(4096 + 1000 * per_tok, device=DEV, dtype=torch.uint8)
    assert ws.numel() < full
    y = torch.empty(N, 4096, device=DEV)
    rc = L.bcl_forward(x.data_ptr(), _cabi.F32, N, 4096, spec.data_ptr(), _cabi.F32, _cabi.LAYOUT_GEMM, p, q, B, None, 0,
                       y.data_ptr(), _cabi.F32, 4096, ws.data_ptr(), ws.numel(), _cabi.stream_ptr(DEV))
    _cabi.check(rc, "bcl_forward")
    torch.cuda.synchronize()
    assert int(ws[:4].view(torch.int32).item()) == 0
    rows = torch.unique(torch.cat([torch.tensor([0, 127, 128, 895, 896, 897, 8000, 16383, 16384, N - 1]),
                                   torch.randint(0, N, (256,), generator=torch.Generator().manual_seed(N))]))      # tile / chunk edges + 256 random rows
    assert_close(y[rows], O.bc_forward_torch(c, x[rows].cpu()), 1e-4)
    with torch.no_grad():
        assert torch.equal(m(x), y)                                                        # unchunked == chunked, deterministic


@pytest.mark.parametrize("inf,outf,B", SHAPES + [(8192, 1024, 256), (8192, 8192, 256), (1024, 3584, 256)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_split_prefill_16bit_activations_tensor_core_dft_stages(inf, outf, B, dtype):
    """fp32 spectra + 16-bit activations (the reference's default Llama configuration, commands.txt:22): all three stages run on the
    tensor cores with (hi, lo) operands (B = 256, the 70B block size: quarter-N launches of the DFT stages); the result must stay in
    the fp32-spectrum class (1e-4 + I/O rounding) and agree with the CUDA-core FFT stages (BCL_PREFILL_DFT=1 equivalent: fp32 activations)."""
    N = 300
    g = torch.Generator().manual_seed(inf + outf + B + 3)
    c = torch.randn(outf // B, inf // B, B, generator=g) * 0.01
    x = torch.randn(N, inf, generator=g).to(dtype)
    m = make_layer_f32(c)
    with torch.no_grad():
        y = m(x.to(DEV))
        y32 = m(x.float().to(DEV))                                # fp32 activations: CUDA-core fp32 FFT stages + split GEMM
    assert prefill_status(DEV) == 0
    assert y.dtype == dtype
    y_ref = O.bc_forward_torch(c, x.float())
    assert_close(y32, y_ref, 1e-4)
    assert_close(y, y_ref, 1e-4, dtype)


def test_prefill_chain_kernel_matches_three_launches(tmp_path):
    """Opt-in single-launch chain kernel (BCL_PREFILL_CHAIN=1: three stages separated by grid barriers) against the default three
    PDL-chained launches: bit-identical, bf16 and fp32 (split) spectra, B = 128 and the two-half B = 256 stages."""
    import os
    import subprocess
    import sys
    child = r"""
import sys, torch
import llama_fft_b200 as bc
from llama_fft_b200.patch_llama_fft import prefill_status
out = sys.argv[1]
res = {}
for (inf, outf, B, N, sdt) in [(4096, 4096, 128, 2048 + 77, torch.bfloat16), (1024, 2048, 128, 700, torch.float32), (1024, 512, 256, 900, torch.bfloat16)]:
    g = torch.Generator().manual_seed(inf + outf + B + N)
    m = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False, spectrum_dtype=sdt)
    with torch.no_grad():
        m.c.copy_(torch.randn(outf // B, inf // B, B, generator=g) * 0.01)
    m = m.to("cuda:0").eval()
    x = torch.randn(N, inf, generator=g).half().to("cuda:0")
    with torch.no_grad():
        y = m(x)
    torch.cuda.synchronize()
    assert prefill_status("cuda:0") == 0
    res[f"{inf}x{outf}x{B}"] = y.cpu()
torch.save(res, out)
"""
    outs = {}
    for tag, flag in (("launches", "0"), ("chain", "1")):
        out = str(tmp_path / f"{tag}.pt")
        env = dict(os.environ, BCL_PREFILL_CHAIN=flag, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        r = subprocess.run([sys.executable, "-c", child, out], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = torch.load(out)
    assert outs["launches"].keys() == outs["chain"].keys()
    for k in outs["launches"]:
        assert torch.equal(outs["launches"][k], outs["chain"][k]), k


def test_workspace_is_per_stream_survives_capture_and_status_is_sticky():
    """one staging buffer per (device, stream); a buffer handed out during CUDA-graph capture is kept alive when a larger
    call regrows the workspace, so replaying the graph still writes into memory it owns; the status word pair is (per call, sticky)."""
    import importlib
    P = importlib.import_module("llama_fft_b200.patch_llama_fft")
    torch.manual_seed(11)
    m = bc.BlockCirculantLinear(512, 512, block_size=128, bias=False, spectrum_dtype=torch.bfloat16).to(DEV).eval()
    x_small = torch.randn(256, 512, device=DEV).half()
    x_big = torch.randn(2048, 512, device=DEV).half()
    side = torch.cuda.Stream(DEV)
    with torch.no_grad():
        y_ref = m(x_small).clone()
        with torch.cuda.stream(side):
            m(x_small)                                                # warm-up on the side stream: its own workspace
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                y_cap = m(x_small)
        keys = [k for k in P._workspaces if k[0] == torch.device(DEV).index]
        assert len(keys) >= 2                                         # current stream + side stream
        with torch.cuda.stream(side):
            ws_before = P._workspaces[(torch.device(DEV).index, side.cuda_stream)]
            assert getattr(ws_before, "_bcl_captured", False)
            y_big = m(x_big)                                          # regrows the side stream's workspace
            ws_after = P._workspaces[(torch.device(DEV).index, side.cuda_stream)]
            assert ws_after.data_ptr() != ws_before.data_ptr() and any(w is ws_before for w in P._captured_workspaces)
            junk = [torch.full((ws_before.numel(),), 7, device=DEV, dtype=torch.uint8) for _ in range(2)]   # would reuse a freed buffer
            graph.replay()
            torch.cuda.synchronize()
        assert torch.equal(y_cap, y_ref)
        assert all(int(j[0].item()) == 7 and int(j[-1].item()) == 7 for j in junk)
        assert y_big.shape == (2048, 512)
    assert prefill_status(DEV) == 0 and prefill_status(DEV, sticky=False) == 0
