"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
  (1) outputs recorded from the unmodified reference (tests/golden), (2) the torch restatement of the reference
  (oracle.bc_forward_torch) on the same seeded inputs, (3) the FFT-free fp64 oracles, and
  (4) size-independent properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star): fp32 spectra  max|d| <= 1e-4 * max|y_ref|;  bf16 spectra <= 1e-2.
16-bit I/O rounds both sides to the activation dtype, so one output ulp may flip: the per-element
bound is 1e-4*max|y| + 2 ulp(dtype)*|y_ref|."""
import ctypes
import io
import contextlib

import numpy as np
import pytest
import torch

import llama_fft_b200 as bc
from llama_fft_b200 import _cabi
from llama_fft_b200.patch_llama_fft import bc_forward
from oracle import bc_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_layer(c, bias, spectrum_dtype=torch.float32):
    p, q, B = c.shape
    m = bc.BlockCirculantLinear(q * B, p * B, block_size=B, bias=bias is not None, spectrum_dtype=spectrum_dtype)
    with torch.no_grad():
        m.c.copy_(c)
        if bias is not None:
            m.bias.data = bias.clone()
    return m.to(DEV).eval()


def assert_close(y, y_ref, tol, io_dtype):
    assert y.shape == y_ref.shape
    yd, rd = y.double().cpu(), y_ref.double().cpu()
    scale = max(rd.abs().max().item(), 1e-30)
    bound = tol * scale + (2 * G.ulp(io_dtype) * rd.abs() if io_dtype != torch.float32 else 0.0)
    bad = (yd - rd).abs() > bound
    assert not bool(torch.as_tensor(bad).any()), f"max rel err {(yd - rd).abs().max().item() / scale:.3e} (tol {tol})"


CASES = list(G.forward_cases())
GPU_CASES = [c for c in CASES if c["B"] in (64, 128, 256)]


@pytest.mark.parametrize("case", GPU_CASES, ids=[c["name"] for c in GPU_CASES])
def test_forward_matches_recorded_reference_outputs(case):
    c, x, bias, y_ref, _ = G.load_case(case)
    m = make_layer(c, bias)
    with torch.no_grad():
        y = m(x.to(DEV))
    assert y.dtype == case["y_dtype"]                     # incl. fp16 x + fp32 bias -> fp32
    assert_close(y, y_ref, 1e-4, case["dtype"])


@pytest.mark.parametrize("case", [c for c in GPU_CASES if c["bias_mode"] != "fp32"][::3], ids=lambda c: c["name"])
def test_bf16_spectrum_within_1e2(case):
    c, x, bias, y_ref, _ = G.load_case(case)
    m = make_layer(c, bias, torch.bfloat16)
    with torch.no_grad():
        y = m(x.to(DEV))
    assert_close(y, y_ref, 1e-2, case["dtype"])


def test_unsupported_block_size_raises_not_falls_back():
    m = bc.BlockCirculantLinear(128, 64, block_size=32, bias=False).to(DEV).eval()
    with torch.no_grad(), pytest.raises(ValueError, match="block size 32"):
        m(torch.zeros(4, 128, device=DEV))


LLAMA = [(4096, 4096, 64), (4096, 4096, 128), (4096, 11008, 128), (11008, 4096, 128), (4096, 4096, 256),
         (11008, 11008, 64), (11008, 11008, 128), (11008, 11008, 256),
         (8192, 8192, 256), (8192, 1024, 256), (8192, 28672, 256), (28672, 8192, 256)]


@pytest.mark.parametrize("inf,outf,B", LLAMA)
@pytest.mark.parametrize("N", [1, 2, 7])
def test_llama_shapes_vs_oracles(inf, outf, B, N):
    """The ten shapes of SURVEY 8(a) (+ the 11008^2 sweep) against the torch restatement AND the fp64 dense oracle."""
    g = torch.Generator().manual_seed(inf + outf + B + N)
    c = torch.randn(outf // B, inf // B, B, generator=g) * 0.01
    x = torch.randn(N, inf, generator=g)
    m = make_layer(c, None)
    with torch.no_grad():
        y = m(x.to(DEV))
    assert_close(y, O.bc_forward_torch(c, x), 1e-4, torch.float32)
    if inf * outf <= 4096 * 11008:
        assert_close(y, torch.from_numpy(O.bc_forward_dense64(c, x)), 1e-4, torch.float32)


@pytest.mark.parametrize("N", [128, 2048])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_many_tokens_3d_and_16bit_io(N, dtype):
    g = torch.Generator().manual_seed(N)
    c = torch.randn(32, 32, 128, generator=g) * 0.01
    x = torch.randn(N // 64, 64, 4096, generator=g).to(dtype)
    m = make_layer(c, None)
    with torch.no_grad():
        y = m(x.to(DEV))
    assert y.shape == (N // 64, 64, 4096) and y.dtype == dtype
    assert_close(y, O.bc_forward_torch(c, x), 1e-4, dtype)


def test_non_contiguous_input_and_row_strides():
    g = torch.Generator().manual_seed(5)
    c = torch.randn(3, 4, 64, generator=g)
    big = torch.randn(6, 512, generator=g)
    m = make_layer(c, None)
    x = big[:, 128:384]                                    # row stride 512, offset 128 elements
    with torch.no_grad():
        assert_close(m(x.to(DEV)[:, :]), O.bc_forward_torch(c, x), 1e-4, torch.float32)
        xs = big.to(DEV)[:, 128:384]
        assert_close(m(xs), O.bc_forward_torch(c, x), 1e-4, torch.float32)
        xt = torch.randn(256, 5, generator=g).t()          # inner stride != 1 -> the reference's reshape copies too
        assert_close(m(xt.to(DEV)), O.bc_forward_torch(c, xt), 1e-4, torch.float32)


@pytest.mark.parametrize("B", [64, 128, 256])
@pytest.mark.parametrize("sdt", [torch.float32, torch.bfloat16])
def test_spectrum_precompute_matches_rfft(B, sdt):
    g = torch.Generator().manual_seed(B)
    c = torch.randn(5, 3, B, generator=g)
    m = make_layer(c, None, sdt)
    spec = m.build_spectrum().float().cpu().double().numpy()
    ref = O.packed_spectrum(c)
    tol = 1e-6 if sdt == torch.float32 else 2.0 ** -8
    assert np.abs(spec - ref).max() <= tol * np.abs(ref).max()
    assert m.build_spectrum().dtype == sdt and m.build_spectrum().shape == (5, 3, B)


@pytest.mark.parametrize("B", [64, 128, 256])
def test_known_answers_impulse_constant_and_golden_c(B):
    """x = e_k in input block i gives, in every output block j, the column c[j,i] rolled by k (C[r,s]=c[(r-s)%B]);
    constant x gives sum(c); the reference's golden c=[1..8] pattern is embedded periodically."""
    g = torch.Generator().manual_seed(B)
    p, q = 3, 2
    c = torch.randn(p, q, B, generator=g)
    m = make_layer(c, None)
    with torch.no_grad():
        for (i, k) in [(0, 0), (1, 5), (1, B - 1)]:
            x = torch.zeros(1, q * B)
            x[0, i * B + k] = 1.0
            y = m(x.to(DEV)).cpu().view(p, B)
            assert (y - torch.roll(c[:, i, :], shifts=k, dims=-1)).abs().max() <= 1e-5 * c.abs().max()
        y = m(torch.ones(1, q * B, device=DEV)).cpu().view(p, B)
        want = c.sum(dim=(1, 2))[:, None].expand(p, B)
        assert (y - want).abs().max() <= 1e-4 * want.abs().max()
        c8 = torch.arange(1.0, 9.0).repeat(B // 8).view(1, 1, B)
        x = torch.randn(1, B, generator=g)
        y = make_layer(c8, None)(x.to(DEV)).cpu()
        assert_close(y, torch.from_numpy(O.bc_forward_dense64(c8, x)), 1e-5, torch.float32)


def test_primitive_matches_recorded_reference():
    v = G.vectors()
    for row in v["prim_meta"]:
        key, n, dt = str(row).split("|")
        c, x, y_ref = G.t(v[f"{key}/c"]), G.t(v[f"{key}/x"], G.DT[dt]), G.t(v[f"{key}/y"], G.DT[dt])
        if int(n) not in (64, 128, 256):
            with pytest.raises(ValueError):
                bc.circulant_matvec_fft(c.to(DEV), x.to(DEV))
            continue
        y = bc.circulant_matvec_fft(c.to(DEV), x.to(DEV))
        assert y.dtype == x.dtype
        assert_close(y, y_ref, 1e-5, G.DT[dt])


def test_reference_validation_suite_on_the_cuda_path():
    """Tests 1, 3, 5 of src/validate_implementation.py (2-norm tolerances 1e-5 / 1e-4 / 1e-4), B lifted to a supported size."""
    torch.manual_seed(42)
    B = 64
    c, x = torch.randn(B), torch.randn(B)
    idx = torch.arange(B)
    C = c[(idx[:, None] - idx[None, :]) % B]
    y = bc.circulant_matvec_fft(c.to(DEV), x.to(DEV)).cpu()
    assert ((C @ x) - y).norm() / (C @ x).norm() < 1e-5
    c_proj = bc.dense_block_to_circulant_column_loss_aware(C)
    y3 = bc.circulant_matvec_fft(c_proj.to(DEV), x.to(DEV)).cpu()
    assert ((C @ x) - y3).norm() / (C @ x).norm() < 1e-4
    m = bc.BlockCirculantLinear(256, 128, block_size=64, bias=False)
    with torch.no_grad():
        m.c.copy_(torch.randn(2, 4, 64))
    W = torch.from_numpy(O.dense_bc_weight64(m.c.detach()))
    xx = torch.randn(4, 256)
    with torch.no_grad():
        yb = m.to(DEV).eval()(xx.to(DEV)).cpu().double()
    yd = xx.double() @ W.T
    assert (yd - yb).norm() / yd.norm() < 1e-4


def test_from_linear_on_gpu_matches_reference_golden():
    v = G.vectors()
    for row in v["from_linear_meta"]:
        key, inf, outf, B, dt, has_bias, _ = str(row).split("|")
        lin = torch.nn.Linear(int(inf), int(outf), bias=bool(int(has_bias))).to(G.DT[dt])
        with torch.no_grad():
            lin.weight.copy_(G.t(v[f"{key}/W"], G.DT[dt]))
        m = bc.BlockCirculantLinear.from_linear(lin.to(DEV), block_size=int(B))
        c_ref = G.t(v[f"{key}/c"], torch.float32)
        assert m.c.is_cuda and m.c.dtype == torch.float32
        assert (m.c.cpu() - c_ref).abs().max() <= 4 * G.ulp(G.DT[dt]) * c_ref.abs().max()


def test_cache_is_invalidated_by_load_and_inplace_updates(tmp_path):
    g = torch.Generator().manual_seed(9)
    c1, c2 = torch.randn(2, 2, 128, generator=g), torch.randn(2, 2, 128, generator=g)
    x = torch.randn(3, 256, generator=g)
    holder = torch.nn.Module()
    holder.proj = make_layer(c1, None)
    with torch.no_grad():
        y1 = holder.proj(x.to(DEV))
        s1 = holder.proj.build_spectrum()
        assert holder.proj.build_spectrum() is s1                       # cached in eval / no-grad
        src = torch.nn.Module()
        src.proj = make_layer(c2, None)
        with contextlib.redirect_stdout(io.StringIO()):
            bc.save_bc_params(src, str(tmp_path / "p.pt"))
            bc.load_bc_params(holder, str(tmp_path / "p.pt"))
        y2 = holder.proj(x.to(DEV))
        assert_close(y2, O.bc_forward_torch(c2, x), 1e-4, torch.float32)
        assert (y1 - y2).abs().max() > 0
        holder.proj.c.mul_(0.5)                                          # an optimizer-style in-place step
        assert_close(holder.proj(x.to(DEV)), O.bc_forward_torch(c2 * 0.5, x), 1e-4, torch.float32)
    # cached vs uncached equivalence (train mode rebuilds the spectrum on every call, reference semantics)
    holder.proj.train()
    with torch.no_grad():
        y3 = holder.proj(x.to(DEV))
        assert torch.equal(y3, holder.proj.eval()(x.to(DEV)))
    assert_close(y3, O.bc_forward_torch(c2 * 0.5, x), 1e-4, torch.float32)


def test_grouped_launch_equals_per_module_calls():
    g = torch.Generator().manual_seed(11)
    shapes = [(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)]
    B = 128
    layers, xs = [], []
    for li in range(3):
        for (p, q) in shapes:
            layers.append(make_layer(torch.randn(p, q, B, generator=g) * 0.01, None))
            xs.append(torch.randn(1, q * B, generator=g).to(torch.float16).to(DEV))
    specs = [m.build_spectrum() for m in layers]
    grp = bc.GroupedBCForward(xs, specs)
    ys = grp.run()
    torch.cuda.synchronize()
    with torch.no_grad():
        for m, x, y in zip(layers, xs, ys):
            assert_close(y, m(x), 1e-5, torch.float16)                   # same math, different partial-sum order
            assert_close(y, O.bc_forward_torch(m.c.detach().cpu(), x.cpu()), 1e-4, torch.float16)
    # with N = 3 tokens, a bias, and bf16 spectra
    layers = [make_layer(torch.randn(4, 6, 64, generator=g), torch.randn(256, generator=g), torch.bfloat16) for _ in range(5)]
    xs = [torch.randn(3, 384, generator=g).to(DEV) for _ in layers]
    grp = bc.GroupedBCForward(xs, [m.build_spectrum() for m in layers], [m.bias.detach() for m in layers], num_ctas=7)
    ys = grp.run()
    with torch.no_grad():
        for m, x, y in zip(layers, xs, ys):
            assert_close(y, m(x), 1e-5, torch.float32)
            assert_close(y, O.bc_forward_torch(m.c.detach().cpu(), x.cpu(), m.bias.detach().cpu()), 1e-2, torch.float32)


def test_host_buffer_entry_point():
    g = torch.Generator().manual_seed(13)
    c = torch.randn(32, 32, 128, generator=g) * 0.01
    m = make_layer(c, None)
    spec = m.build_spectrum()
    N = 4
    hx = torch.randn(N, 4096, generator=g).to(torch.float16).pin_memory()
    hy = torch.empty(N, 4096, dtype=torch.float16).pin_memory()
    dx, dy = torch.empty(N, 4096, dtype=torch.float16, device=DEV), torch.empty(N, 4096, dtype=torch.float16, device=DEV)
    rc = _cabi.lib().bcl_forward_host(hx.data_ptr(), _cabi.F16, N, spec.data_ptr(), _cabi.F32, _cabi.LAYOUT_STREAM, 32, 32, 128,
                                      None, 0, hy.data_ptr(), _cabi.F16, dx.data_ptr(), dy.data_ptr(), None, 0,
                                      _cabi.stream_ptr(torch.device(DEV)))
    _cabi.check(rc, "bcl_forward_host")
    assert_close(hy, O.bc_forward_torch(c, hx), 1e-4, torch.float16)


# ---- size-independent properties at BASELINE.json's full sizes (the oracle is too slow there) --------------------
@pytest.mark.parametrize("inf,outf,B,N", [(4096, 11008, 128, 16384), (11008, 4096, 128, 4096), (8192, 28672, 256, 512)])
def test_full_size_properties(inf, outf, B, N):
    g = torch.Generator(device=DEV).manual_seed(1)
    p, q = outf // B, inf // B
    c = torch.randn(p, q, B, device=DEV, generator=g) * 0.01
    m = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False).to(DEV).eval()
    with torch.no_grad():
        m.c.copy_(c)
        x1 = torch.randn(N, inf, device=DEV, generator=g)
        x2 = torch.randn(N, inf, device=DEV, generator=g)
        y1, y2 = m(x1), m(x2)
        scale = y1.abs().max().item()
        # linearity
        assert (m(2.0 * x1 - 0.5 * x2) - (2.0 * y1 - 0.5 * y2)).abs().max().item() <= 1e-4 * scale * 3
        # circulant structure: rolling every input block by s rolls every output block by s
        s = 37
        xr = torch.roll(x1.view(N, q, B), shifts=s, dims=-1).reshape(N, inf)
        yr = torch.roll(y1.view(N, p, B), shifts=s, dims=-1).reshape(N, outf)
        assert (m(xr) - yr).abs().max().item() <= 1e-4 * scale
        # spot rows against the torch restatement of the reference
        rows = torch.unique(torch.cat([torch.tensor([0, N // 2, N - 1]), torch.randint(0, N, (256,), generator=torch.Generator().manual_seed(N))]))
        ref = O.bc_forward_torch(c.cpu(), x1[rows].cpu())
        assert_close(y1[rows], ref, 1e-4, torch.float32)
        # a zero input block-row of tokens gives exact zeros; determinism across calls
        assert torch.equal(m(torch.zeros(3, inf, device=DEV)), torch.zeros(3, outf, device=DEV))
        assert torch.equal(m(x1), y1)


def test_hypothesis_small_shapes():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(B=st.sampled_from([64, 128, 256]), p=st.integers(1, 5), q=st.integers(1, 5), n=st.integers(1, 9),
           dt=st.sampled_from([torch.float32, torch.float16, torch.bfloat16]), bias=st.booleans(), seed=st.integers(0, 2 ** 16))
    def run(B, p, q, n, dt, bias, seed):
        g = torch.Generator().manual_seed(seed)
        c = torch.randn(p, q, B, generator=g)
        b = (torch.randn(p * B, generator=g) * 0.1).to(dt) if bias else None
        x = torch.randn(n, q * B, generator=g).to(dt)
        m = make_layer(c, b)
        with torch.no_grad():
            y = m(x.to(DEV))
        assert_close(y, O.bc_forward_torch(c, x, b), 1e-4, dt)

    run()


def test_pipelined_host_entry_matches_serial_host_entry():
    """bcl_forward_grouped_host_pipelined (3 streams, 2 slots) delivers the same bytes as the synchronous host call."""
    g = torch.Generator().manual_seed(21)
    shapes = [(4, 3), (2, 5), (3, 3)]
    slots, hxs, hys = [], [], []
    for s in range(2):
        specs = []
        for (p, q) in shapes:
            m = make_layer(torch.randn(p, q, 128, generator=g) * 0.1, None)
            specs.append(m.build_spectrum())
        grp = bc.GroupedBCForward.from_packed(shapes, specs, 1, torch.float16, torch.device(DEV))
        slots.append(grp)
        hxs.append(torch.randn(grp.x_all.numel(), generator=g).half().pin_memory())
        hys.append(torch.zeros(grp.y_all.numel(), dtype=torch.float16).pin_memory())
    ref = []
    for grp, hx in zip(slots, hxs):
        out = torch.zeros(grp.y_all.numel(), dtype=torch.float16).pin_memory()
        grp.run_host(hx, out)
        ref.append(out.clone())
    pipe = bc.HostPipe(torch.device(DEV))
    for it in range(7):                                     # slots are reused: exercises the dev_x / dev_y hazards
        k = it % 2
        slots[k].wait_host()
        if it >= 2:
            assert torch.equal(hys[k], ref[k])
            hys[k].zero_()
        slots[k].submit_host(hxs[k], hys[k], pipe)
    for k in range(2):
        slots[k].wait_host()
        assert torch.equal(hys[k], ref[k])


@pytest.mark.parametrize("B,shapes,reps,num_ctas,bias", [
    (128, [(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)], 8, None, False),       # 56 problems over 148 CTAs: ring wraps, group-buffer reuse
    (128, [(32, 32), (32, 86), (86, 32)], 4, 5, True),                         # few CTAs: many segments per CTA (X double buffer), bias
    (64, [(16, 24), (40, 8), (9, 70)], 5, 9, False),
    (256, [(8, 8), (12, 20), (20, 12)], 5, 11, True),
])
def test_grouped_decode_bf16_row_per_warp_kernel(B, shapes, reps, num_ctas, bias):
    """bf16 spectra, one token per problem: the grouped launch runs the row-per-warp kernel (bcl_stream_rw.cu), whose row warps drift
    apart (sequence-number guarded barriers).  Every problem must match its own per-module forward (same math, other summation
    order) and the oracle in the bf16 class; a second launch of the same plan must be bit-identical (no ordering effects)."""
    g = torch.Generator().manual_seed(B + len(shapes) + reps)
    layers, xs = [], []
    for _ in range(reps):
        for (p, q) in shapes:
            b = torch.randn(p * B, generator=g) * 0.1 if bias else None
            layers.append(make_layer(torch.randn(p, q, B, generator=g) * 0.02, b, torch.bfloat16))
            xs.append(torch.randn(1, q * B, generator=g).to(torch.float16).to(DEV))
    specs = [m.build_spectrum() for m in layers]
    kw = {} if num_ctas is None else {"num_ctas": num_ctas}
    grp = bc.GroupedBCForward(xs, specs, [m.bias.detach() for m in layers] if bias else None, **kw)
    ys = [y.clone() for y in grp.run()]
    ys2 = grp.run()
    torch.cuda.synchronize()
    with torch.no_grad():
        for m, x, y, y2 in zip(layers, xs, ys, ys2):
            assert torch.equal(y, y2)
            # the spectrum product is rounded to x.dtype (fp16) BEFORE the bias is added (reference :266-271): another summation
            # order may move that rounding by one fp16 ulp even when y itself is fp32 (fp32 bias)
            # (with a bias the flipped ulp belongs to the value before the bias was added, which |y| no longer bounds: absolute 1e-3)
            assert_close(y, m(x), 1e-3 if bias else 1e-5, torch.float16)
            assert_close(y, O.bc_forward_torch(m.c.detach().cpu(), x.cpu(), None if m.bias is None else m.bias.detach().cpu()), 1e-2, torch.float16)


@pytest.mark.parametrize("B,shapes,reps,num_ctas,bias,dtype", [
    (128, [(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)], 3, None, False, torch.float16),      # Llama-2-7B layer x 3, one CTA per SM
    (128, [(32, 32), (86, 32), (32, 86), (5, 3), (9, 1), (40, 17)], 2, 11, True, torch.float16),   # ragged rows / K padding, few CTAs
    (128, [(32, 32)] * 2 + [(3, 64)], 1, 5, False, torch.bfloat16),
    (64, [(64, 64), (10, 7), (172, 64)], 2, 37, False, torch.float32)])
def test_grouped_decode_tensor_core_kernel(B, shapes, reps, num_ctas, bias, dtype):
    """BCL_LAYOUT_DECODE_TC: the grouped batch-1 launch with the per-frequency contraction on tcgen05 (bcl_stream_tc.cu): bf16 spectrum
    image streamed as the UMMA A operand, bf16 x spectrum as the B operand, fp32 accumulation in TMEM, inverse FFT + reference epilogue
    on the CUDA cores.  Every problem must match the oracle in the bf16 class (the x spectrum is rounded to bf16 too, like Xf of the
    prefill path) and its own per-module forward with bf16 spectra; a second launch of the same plan must be bit-identical."""
    g = torch.Generator().manual_seed(B + len(shapes) + reps)
    layers, xs, tcs = [], [], []
    for _ in range(reps):
        for (p, q) in shapes:
            b = torch.randn(p * B, generator=g) * 0.1 if bias else None
            layers.append(make_layer(torch.randn(p, q, B, generator=g) * 0.02, b, torch.bfloat16))
            xs.append(torch.randn(1, q * B, generator=g).to(dtype).to(DEV))
            tcs.append(bc.decode_tc_spectrum(layers[-1].c.detach()))
    kw = {} if num_ctas is None else {"num_ctas": num_ctas}
    grp = bc.GroupedBCForward(xs, tcs, [m.bias.detach() for m in layers] if bias else None, layout=_cabi.LAYOUT_DECODE_TC,
                              shapes=shapes * reps, block_size=B, **kw)
    ys = [y.clone() for y in grp.run()]
    ys2 = grp.run()
    torch.cuda.synchronize()
    with torch.no_grad():
        for m, x, y, y2 in zip(layers, xs, ys, ys2):
            assert torch.equal(y, y2)
            assert not torch.isnan(y.float()).any()
            assert_close(y, O.bc_forward_torch(m.c.detach().cpu(), x.cpu(), None if m.bias is None else m.bias.detach().cpu()), 1e-2, dtype)
            assert_close(y, m(x), 1e-2, dtype)


@pytest.mark.parametrize("sdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,shapes,num_ctas", [(128, [(32, 16), (7, 3), (40, 8), (86, 32), (5, 1), (32, 86)], 9),
                                                (64, [(64, 64), (9, 5), (33, 32), (16, 70)], None),
                                                (256, [(16, 8), (12, 20), (5, 2), (32, 16)], 13)])
def test_grouped_decode_split_q_kernel_two_rows_per_stage(monkeypatch, sdt, B, shapes, num_ctas):
    """The warp-specialised split-q kernel (bcl_stream_ws.cu) puts TWO block-rows of a narrow segment (2 q <= blocks per stage) into one
    ring stage and multiplies them with independent accumulators; 16-bit spectra use 16 KB stages (twice the blocks).  Forced for bf16
    spectra too (BCL_GROUP_VARIANT=ws; the default there is the row-per-warp kernel).  Odd row counts, q not a multiple of the MAC
    warps, segments cut in the middle of a row pair: every problem must match its per-module forward and the oracle."""
    monkeypatch.setenv("BCL_GROUP_VARIANT", "ws")
    g = torch.Generator().manual_seed(B + len(shapes))
    layers, xs = [], []
    for rep in range(3):
        for (p, q) in shapes:
            layers.append(make_layer(torch.randn(p, q, B, generator=g) * 0.02, None, sdt))
            xs.append(torch.randn(1, q * B, generator=g).to(torch.float16).to(DEV))
    kw = {} if num_ctas is None else {"num_ctas": num_ctas}
    grp = bc.GroupedBCForward(xs, [m.build_spectrum() for m in layers], None, **kw)
    ys = [y.clone() for y in grp.run()]
    ys2 = grp.run()
    torch.cuda.synchronize()
    tol = 1e-4 if sdt == torch.float32 else 1e-2
    with torch.no_grad():
        for m, x, y, y2 in zip(layers, xs, ys, ys2):
            assert torch.equal(y, y2)
            assert_close(y, m(x), 1e-5, torch.float16)
            assert_close(y, O.bc_forward_torch(m.c.detach().cpu(), x.cpu()), tol, torch.float16)


def test_grouped_plans_marked_independent_skip_the_dependency_wait(monkeypatch):
    """BCL_GROUP_INDEPENDENT=1 at plan-encode time: launches of the plan do not wait for the previous launch of the stream
    (griddepcontrol.wait skipped), so consecutive launches that use DIFFERENT buffers overlap on the SMs.  Two alternating groups,
    many launches back to back: every output must stay bit-identical to the ordinary plans' and to the per-module forward."""
    g = torch.Generator().manual_seed(99)
    shapes = [(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)]
    B = 128
    layers = [make_layer(torch.randn(p, q, B, generator=g) * 0.01, None) for _ in range(2) for (p, q) in shapes]
    xs = [torch.randn(1, q * B, generator=g).to(torch.float16).to(DEV) for _ in range(2) for (p, q) in shapes]
    specs = [m.build_spectrum() for m in layers]
    n = len(shapes)
    ref = [bc.GroupedBCForward(xs[k * n:(k + 1) * n], specs[k * n:(k + 1) * n]) for k in range(2)]
    ref_out = [[y.clone() for y in grp.run()] for grp in ref]
    monkeypatch.setenv("BCL_GROUP_INDEPENDENT", "1")
    ind = [bc.GroupedBCForward(xs[k * n:(k + 1) * n], specs[k * n:(k + 1) * n]) for k in range(2)]
    monkeypatch.delenv("BCL_GROUP_INDEPENDENT")
    for i in range(200):
        ind[i & 1].run()
    torch.cuda.synchronize()
    for k in range(2):
        for y, yr in zip(ind[k].ys, ref_out[k]):
            assert torch.equal(y, yr)


@pytest.mark.parametrize("spec", ["10,1", "40,2", "25,4", "90,3"])
def test_grouped_decode_dynamic_tail_is_bit_identical_and_rearms_its_counters(monkeypatch, spec):
    """Dynamic tail of the fp32 decode kernel (bcl_api.cu / bcl_stream_ws.cu): part of the rows is drawn by the CTAs from a ticket
    counter in the device plan.  Which CTA computes a row must not change a bit of it; the counters must be back at zero after every
    launch (200 launches back to back, two plans alternating, also with a small grid so that every CTA draws many items); and a CTA
    without a static share (more CTAs than static rows) must work from tickets alone."""
    g = torch.Generator().manual_seed(1234)
    shapes = ([(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)]) * 2 + [(5, 3), (1, 40)]
    B = 128
    layers = [make_layer(torch.randn(p, q, B, generator=g) * 0.01, None) for (p, q) in shapes]
    xs = [[torch.randn(1, q * B, generator=g).to(torch.float16).to(DEV) for (p, q) in shapes] for _ in range(2)]
    specs = [m.build_spectrum() for m in layers]
    monkeypatch.setenv("BCL_GROUP_DYNAMIC", "0")
    ref_out = [[y.clone() for y in bc.GroupedBCForward(xs[k], specs).run()] for k in range(2)]
    monkeypatch.setenv("BCL_GROUP_DYNAMIC", spec)
    for num_ctas in (None, 8, 40):
        grp = [bc.GroupedBCForward(xs[k], specs, num_ctas=num_ctas) for k in range(2)]
        for i in range(200):
            grp[i & 1].run()
        torch.cuda.synchronize()
        for k in range(2):
            for y, yr in zip(grp[k].ys, ref_out[k]):
                assert torch.equal(y, yr)
    # the oracle, once
    for m, x, y in zip(layers[:3], xs[0][:3], ref_out[0][:3]):
        assert_close(y, O.bc_forward_torch(m.c.detach().cpu(), x.cpu()), 2e-3, torch.float16)
