"""CPU-side tests (-m "not gpu"): the C-ABI library loads and exports every symbol the header declares, argument
validation and plan encoding (no compute), and the host logic of the drop-in module against the golden vectors."""
import ctypes
import io
import os
import re
import contextlib

import numpy as np
import pytest
import torch

import llama_fft_b200 as bc
from llama_fft_b200 import _cabi
from tests import _golden as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "bclinear.h")).read()
    declared = set(re.findall(r"\b(bcl_[a-z_]+)\s*\(", header))
    assert declared == set(_cabi.SYMBOLS)
    L = _cabi.lib()
    for name in declared:
        assert hasattr(L, name)
    assert L.bcl_abi_version() == _cabi.ABI_VERSION
    m = re.search(r"#define BCL_ABI_VERSION (\d+)", header)
    assert int(m.group(1)) == _cabi.ABI_VERSION


def test_spectrum_bytes_and_argument_errors():
    L = _cabi.lib()
    assert L.bcl_spectrum_bytes(32, 32, 128, _cabi.F32, _cabi.LAYOUT_STREAM) == 32 * 32 * 128 * 4
    assert L.bcl_spectrum_bytes(86, 32, 128, _cabi.BF16, _cabi.LAYOUT_STREAM) == 86 * 32 * 128 * 2
    assert L.bcl_spectrum_bytes(1, 1, 100, _cabi.F32, _cabi.LAYOUT_STREAM) == 0
    buf = (ctypes.c_float * 64)()
    a = ctypes.addressof(buf)
    # unsupported block size -> BCL_ERR_BLOCK -> ValueError, before anything touches a device
    rc = L.bcl_precompute_spectrum(a, 1, 1, 32, a, _cabi.F32, _cabi.LAYOUT_STREAM, None)
    assert rc == -3
    with pytest.raises(ValueError, match="block size 32"):
        _cabi.check(rc, "bcl_precompute_spectrum")
    rc = L.bcl_forward(a, _cabi.F32, 1, 64, a, _cabi.F32, _cabi.LAYOUT_STREAM, 1, 1, 64, None, 0, a, _cabi.F16, 64, None, 0, None)
    assert rc == -5 and b"promotion" in L.bcl_last_error()
    rc = L.bcl_forward(a + 4, _cabi.F32, 1, 64, a, _cabi.F32, _cabi.LAYOUT_STREAM, 1, 1, 64, None, 0, a, _cabi.F32, 64, None, 0, None)
    assert rc == -4
    rc = L.bcl_forward(None, _cabi.F32, 0, 64, None, _cabi.F32, _cabi.LAYOUT_STREAM, 1, 1, 64, None, 0, None, _cabi.F32, 64, None, 0, None)
    assert rc == 0      # N == 0 is a no-op
    rc = L.bcl_forward(a, _cabi.F32, 1, 32, a, _cabi.F32, _cabi.LAYOUT_STREAM, 1, 1, 64, None, 0, a, _cabi.F32, 64, None, 0, None)
    assert rc == -2


def _encode(shapes, B, num_ctas, N=1):
    L = _cabi.lib()
    probs = (_cabi.Problem * len(shapes))()
    for g, (p, q) in enumerate(shapes):
        probs[g] = _cabi.Problem(0x1000, 0x2000, None, 0x3000, N, q * B, p * B, p, q, _cabi.F16, _cabi.F16, 0, 0)
    n = L.bcl_group_plan_bytes_for(probs, len(shapes), B, _cabi.F32, num_ctas)
    assert n >= L.bcl_group_plan_bytes(len(shapes), num_ctas)
    buf = (ctypes.c_uint8 * n)()
    rc = L.bcl_group_plan_encode(probs, len(shapes), B, _cabi.F32, num_ctas, ctypes.addressof(buf), n)
    return rc, bytes(buf)


def _decode(blob):
    hdr = np.frombuffer(blob[:64], dtype=np.uint8)
    magic, count, num_ctas, num_segs, B, sdt, nt, qmax = np.frombuffer(hdr[:32].tobytes(), dtype=np.int32)
    variant = int(np.frombuffer(blob[32:36], dtype=np.int32)[0])
    off_p, off_c, off_s, total = np.frombuffer(blob[40:72], dtype=np.uint64)
    cta_ptr = np.frombuffer(blob[int(off_c):int(off_c) + 4 * (num_ctas + 1)], dtype=np.int32)
    segs = np.frombuffer(blob[int(off_s):int(off_s) + 16 * num_segs], dtype=np.int32).reshape(-1, 4)
    num_dyn, dyn_rows = (int(v) for v in np.frombuffer(blob[72:80], dtype=np.int32))
    off_ctr = int(np.frombuffer(blob[80:88], dtype=np.uint64)[0])
    dyn = np.frombuffer(blob[int(off_s) + 16 * num_segs:int(off_s) + 16 * num_segs + 32 * num_dyn], dtype=np.int32).reshape(-1, 8)[:, :4]
    return dict(count=count, num_ctas=num_ctas, num_segs=num_segs, nt=nt, qmax=qmax, variant=variant, num_dyn=num_dyn, dyn_rows=dyn_rows,
                off_ctr=off_ctr, dyn=dyn, total=int(total)), cta_ptr, segs


def _variant(shape, B, sdt):
    L = _cabi.lib()
    p, q = shape
    probs = (_cabi.Problem * 32)()
    for g in range(32):
        probs[g] = _cabi.Problem(0x1000, 0x2000, None, 0x3000, 1, q * B, p * B, p, q, _cabi.F16, _cabi.F16, 0, 0)
    n = L.bcl_group_plan_bytes_for(probs, 32, B, sdt, 148)
    buf = (ctypes.c_uint8 * n)()
    assert L.bcl_group_plan_encode(probs, 32, B, sdt, 148, ctypes.addressof(buf), n) == 0
    return _decode(bytes(buf))[0]["variant"]


def test_group_plan_selects_the_pipelined_kernels_for_every_sweep_shape(monkeypatch):
    """BASELINE configs[3] (block-size sweep: in = out in {4096, 11008}, B in {64, 128, 256}) and the Llama shapes: the grouped decode
    launch must run the warp-specialised TMA kernels (variant 1 = split-q, fp32 spectra; 2 = row-per-warp, bf16 spectra), never the plain
    fallback (variant 0) -- 11008 at B = 64 (q = 172) silently fell back once (144 us against 42 us per 32-layer launch)."""
    monkeypatch.delenv("BCL_GROUP_VARIANT", raising=False)
    for dim in (4096, 11008):
        for B in (64, 128, 256):
            q = dim // B
            assert _variant((q, q), B, _cabi.F32) == 1, (dim, B)
            assert _variant((q, q), B, _cabi.BF16) == 2, (dim, B)
    # (70B down_proj, 28672 -> 8192 at B = 256, is refused by the grouped call -- its 112 KB x spectrum leaves no room for a ring -- and
    #  goes through bcl_forward per module)
    for shape, B in (((32, 86), 128), ((86, 32), 128), ((32, 32), 256), ((112, 32), 256), ((4, 32), 256)):
        assert _variant(shape, B, _cabi.F32) == 1 and _variant(shape, B, _cabi.BF16) == 2, (shape, B)


def test_group_plan_covers_every_row_once_and_balances_bytes(monkeypatch):
    # one Llama-2-7B decoder layer x 4: q,k,v,o (32x32), gate, up (86x32), down (32x86) at B=128
    layer = [(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)]
    shapes = layer * 4
    monkeypatch.delenv("BCL_PLAN_COST", raising=False)            # the default: bytes only
    monkeypatch.setenv("BCL_GROUP_DYNAMIC", "0")                  # every row in a static share
    rc, blob = _encode(shapes, 128, 296)
    assert rc == 0
    H, cta_ptr, segs = _decode(blob)
    assert H["count"] == len(shapes) and H["qmax"] == 86 and H["nt"] == 1 and H["variant"] == 1
    assert cta_ptr[0] == 0 and cta_ptr[-1] == H["num_segs"] and np.all(np.diff(cta_ptr) >= 0)
    seen = {g: np.zeros(p, dtype=int) for g, (p, q) in enumerate(shapes)}
    for g, j0, j1, _ in segs:
        assert 0 <= j0 < j1 <= shapes[g][0]
        seen[g][j0:j1] += 1
    assert all((v == 1).all() for v in seen.values())
    w = np.zeros(H["num_ctas"])
    for b in range(H["num_ctas"]):
        for g, j0, j1, _ in segs[cta_ptr[b]:cta_ptr[b + 1]]:
            w[b] += (j1 - j0) * shapes[g][1]
    total = sum(p * q for p, q in shapes)
    assert w.sum() == total
    assert w.max() <= total / H["num_ctas"] + 86 + 1          # within one row of the ideal share



def test_group_plan_cost_model_override(monkeypatch):
    """BCL_PLAN_COST="arow,aseg": a row weighs q + arow blocks and every problem start aseg blocks (an A/B switch: the default is bytes
    only, which measured faster on B200, bcl_api.cu) -- CTAs with many short rows or many segments get fewer bytes.  Every row once."""
    layer = [(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)]
    shapes = layer * 32
    monkeypatch.setenv("BCL_PLAN_COST", "3,105")
    monkeypatch.setenv("BCL_GROUP_DYNAMIC", "0")
    rc, blob = _encode(shapes, 128, 148)
    assert rc == 0
    H, cta_ptr, segs = _decode(blob)
    seen = {g: np.zeros(p, dtype=int) for g, (p, q) in enumerate(shapes)}
    w, nbytes = np.zeros(H["num_ctas"]), np.zeros(H["num_ctas"])
    for b in range(H["num_ctas"]):
        for g, j0, j1, _ in segs[cta_ptr[b]:cta_ptr[b + 1]]:
            seen[g][j0:j1] += 1
            w[b] += (j1 - j0) * (shapes[g][1] + 3) + (105 if j0 == 0 else 0)
            nbytes[b] += (j1 - j0) * shapes[g][1]
    assert all((v == 1).all() for v in seen.values())
    total = sum(p * (q + 3) + 105 for p, q in shapes)
    assert abs(w.sum() - total) < 1e-6
    assert w.max() <= total / H["num_ctas"] + (86 + 3) + 105 + 1   # within one row and one segment start of the ideal share
    assert nbytes.max() > nbytes.min() * 1.05                      # ... and the bytes are no longer equal

def test_group_plan_dynamic_tail(monkeypatch):
    """BCL_GROUP_DYNAMIC=20 (opt-in: measured slower than the static plan) for the fp32 decode kernel at 7B scale: ~20 % of
    the weight is not assigned to CTAs but cut into 4-row items that the CTAs draw from a ticket counter (bcl_api.cu: dynamic tail).
    Static segments + items cover every row once, the static shares are byte-balanced, the items come from the narrow-input problems,
    counters start at zero.  Without the request every row is in a static share."""
    layer = [(32, 32)] * 4 + [(86, 32)] * 2 + [(32, 86)]
    shapes = layer * 32
    monkeypatch.delenv("BCL_PLAN_COST", raising=False)
    monkeypatch.delenv("BCL_GROUP_DYNAMIC", raising=False)
    monkeypatch.delenv("BCL_GROUP_INDEPENDENT", raising=False)
    assert _decode(_encode(shapes, 128, 148)[1])[0]["num_dyn"] == 0
    monkeypatch.setenv("BCL_GROUP_DYNAMIC", "20")
    rc, blob = _encode(shapes, 128, 148)
    assert rc == 0
    H, cta_ptr, segs = _decode(blob)
    assert H["variant"] == 1 and H["num_dyn"] > 148 and H["dyn_rows"] == 4
    assert H["off_ctr"] % 16 == 0 and H["off_ctr"] + 8 <= H["total"] and blob[H["off_ctr"]:H["off_ctr"] + 16] == bytes(16)
    seen = {g: np.zeros(p, dtype=int) for g, (p, q) in enumerate(shapes)}
    for g, j0, j1, _ in list(segs) + list(H["dyn"]):
        assert 0 <= j0 < j1 <= shapes[g][0]
        seen[g][j0:j1] += 1
    assert all((v == 1).all() for v in seen.values())
    total = sum(p * q for p, q in shapes)
    wd = sum((j1 - j0) * shapes[g][1] for g, j0, j1, _ in H["dyn"])
    assert 0.19 * total <= wd <= 0.22 * total
    assert all(j1 - j0 <= 4 and shapes[g][1] == 32 for g, j0, j1, _ in H["dyn"])
    w = np.zeros(H["num_ctas"])
    for b in range(H["num_ctas"]):
        for g, j0, j1, _ in segs[cta_ptr[b]:cta_ptr[b + 1]]:
            w[b] += (j1 - j0) * shapes[g][1]
    assert w.sum() == total - wd and w.max() <= (total - wd) / 148 + 86 + 1 and w.min() > 0
    # a percentage and an item size on request; plans marked independent keep every row static (their launches overlap: no shared counter)
    monkeypatch.setenv("BCL_GROUP_DYNAMIC", "10,2")
    H2 = _decode(_encode(shapes, 128, 148)[1])[0]
    assert H2["dyn_rows"] == 2 and 0.09 * total <= sum((j1 - j0) * shapes[g][1] for g, j0, j1, _ in H2["dyn"]) <= 0.11 * total
    monkeypatch.setenv("BCL_GROUP_DYNAMIC", "50,1")               # more items than the plan has room for: the item size doubles
    H3 = _decode(_encode(shapes, 128, 148)[1])[0]
    assert H3["num_dyn"] <= 16 * 148 and H3["dyn_rows"] > 1
    monkeypatch.setenv("BCL_GROUP_DYNAMIC", "20")
    monkeypatch.setenv("BCL_GROUP_INDEPENDENT", "1")
    assert _decode(_encode(shapes, 128, 148)[1])[0]["num_dyn"] == 0


def test_group_plan_rejects_bad_requests():
    rc, _ = _encode([(4, 4)], 128, 8, N=9)
    assert rc == -2
    rc, _ = _encode([(4, 4)], 100, 8)
    assert rc == -3
    rc, _ = _encode([(2, 112)], 256, 8, N=8)       # 8 x 112 x 256 x 4 B of x-spectrum cannot fit shared memory
    assert rc == -8


# ---- host logic of the drop-in module -------------------------------------------------------------------------
def test_module_surface_matches_reference():
    m = bc.BlockCirculantLinear(256, 128, block_size=64, bias=True)
    assert (m.in_features, m.out_features, m.block_size, m.in_blocks, m.out_blocks) == (256, 128, 64, 4, 2)
    assert m.c.shape == (2, 4, 64) and m.c.dtype == torch.float32 and m.bias.shape == (128,)
    assert set(m.state_dict().keys()) == {"c", "bias"}
    assert bc.BlockCirculantLinear(64, 64, 64, bias=False).bias is None
    assert abs(m.c.std().item() - 0.01) < 2e-3                      # reference init randn * 0.01
    with pytest.raises(AssertionError):
        bc.BlockCirculantLinear(100, 64, block_size=64)
    assert (bc.BLOCK_SIZE, bc.NUM_LAYERS_TO_PATCH) == (256, 1) and isinstance(bc.MODEL_PATH, str)
    import inspect
    sig = inspect.signature(bc.patch_mlp_with_block_circulant)
    for k in ("num_layers_to_patch", "block_size", "patch_position"):
        assert sig.parameters[k].kind is inspect.Parameter.KEYWORD_ONLY
    assert sig.parameters["block_size"].default == 256 and sig.parameters["patch_position"].default == "first"


def test_forward_fails_loudly_without_cuda_and_on_bad_dims():
    m = bc.BlockCirculantLinear(64, 64, block_size=64, bias=False).eval()
    with pytest.raises(ValueError, match="Unsupported input dim"):
        m(torch.zeros(64))
    with pytest.raises(ValueError, match="Unsupported input dim"):
        m(torch.zeros(1, 1, 1, 64))
    with pytest.raises(ValueError, match="in_features mismatch"):
        m(torch.zeros(2, 3, 128))
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(2, 64))                                       # CPU tensor: no fallback
    with pytest.raises(RuntimeError, match="no CPU"):
        bc.circulant_matvec_fft(torch.zeros(64), torch.zeros(64))


def test_projections_match_reference_golden():
    v = G.vectors()
    for row in v["proj_meta"]:
        key, B, dt = str(row).split("|")
        W = G.t(v[f"{key}/W"], G.DT[dt])
        tol = 4 * G.ulp(G.DT[dt])
        for name, fn in [("loss_aware", bc.dense_block_to_circulant_column_loss_aware),
                         ("diag", lambda w: bc.dense_block_to_circulant_column(w, convention="diag")),
                         ("diag_inv", lambda w: bc.dense_block_to_circulant_column(w, convention="diag_inv"))]:
            ref = G.t(v[f"{key}/{name}"], G.DT[dt]).float()
            got = fn(W)
            assert got.dtype == W.dtype and got.shape == ref.shape
            assert (got.float() - ref).abs().max() <= tol * ref.abs().max()
    with pytest.raises(ValueError, match="Unknown convention"):
        bc.dense_block_to_circulant_column(torch.eye(4), convention="x")
    c_true = G.t(v["known/c_true"])
    assert torch.equal(bc.circulant_from_first_col(c_true), G.t(v["known/circ_from_first_col"]))
    rec = bc.dense_block_to_circulant_column_loss_aware(G.t(v["known/C"]))
    assert (rec / c_true).std() < 1e-5                              # reference test 2


def test_from_linear_cpu_matches_reference_golden():
    v = G.vectors()
    for row in v["from_linear_meta"]:
        key, inf, outf, B, dt, has_bias, _ = str(row).split("|")
        lin = torch.nn.Linear(int(inf), int(outf), bias=bool(int(has_bias)))
        lin = lin.to(G.DT[dt])
        with torch.no_grad():
            lin.weight.copy_(G.t(v[f"{key}/W"], G.DT[dt]))
            if int(has_bias):
                lin.bias.copy_(G.t(v[f"{key}/lin_bias"], G.DT[dt]))
        m = bc.BlockCirculantLinear.from_linear(lin, block_size=int(B))
        c_ref = G.t(v[f"{key}/c"], torch.float32)
        assert m.c.dtype == torch.float32 and m.c.shape == c_ref.shape
        assert (m.c - c_ref).abs().max() <= 4 * G.ulp(G.DT[dt]) * c_ref.abs().max()
        if int(has_bias):
            assert torch.equal(m.bias.detach().float(), G.t(v[f"{key}/bc_bias"]).float())


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.layers = torch.nn.ModuleList()
        for _ in range(2):
            blk = torch.nn.Module()
            blk.gate_proj = bc.BlockCirculantLinear(128, 256, block_size=64, bias=False)
            blk.down_proj = bc.BlockCirculantLinear(256, 128, block_size=64, bias=True)
            blk.dense = torch.nn.Linear(8, 8)
            self.layers.append(blk)


def test_checkpoint_format_is_the_references(tmp_path):
    ref_state = torch.load(G.GOLDEN_PT, map_location="cpu")
    model = _Tiny()
    with contextlib.redirect_stdout(io.StringIO()) as out:
        bc.load_bc_params(model, G.GOLDEN_PT)                       # a file the REFERENCE wrote
    assert "loaded=6, skipped=0" in out.getvalue()
    for k, t in ref_state.items():
        mod, par = k.rsplit(".", 1)
        cur = model
        for part in mod.split("."):
            cur = getattr(cur, part)
        assert torch.equal(getattr(cur, par).detach(), t)
    path = str(tmp_path / "sub" / "bc.pt")
    with contextlib.redirect_stdout(io.StringIO()):
        bc.save_bc_params(model, path)
    mine = torch.load(path, map_location="cpu")
    assert list(mine.keys()) == list(ref_state.keys())
    assert all(torch.equal(mine[k], ref_state[k]) and mine[k].device.type == "cpu" for k in mine)
    # strict_shapes: mismatching tensors are skipped, not an error (reference :332-334)
    bad = dict(ref_state)
    bad["layers.0.gate_proj.c"] = torch.zeros(1, 1, 64)
    bad["something.else"] = torch.zeros(1)
    bad["layers.9.gate_proj.c"] = torch.zeros(1)
    torch.save(bad, path)
    with contextlib.redirect_stdout(io.StringIO()) as out:
        bc.load_bc_params(model, path)
    assert "loaded=5, skipped=3" in out.getvalue()


def test_spectrum_key_tracks_parameter_version():
    m = bc.BlockCirculantLinear(64, 64, block_size=64, bias=False)
    k0 = m._spectrum_key()
    with torch.no_grad():
        m.c.mul_(2.0)
    assert m._spectrum_key() != k0                                  # optimizer steps / load_bc_params invalidate


def _tiny_llama(layers=3, hidden=128, inter=256, heads=2):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers, num_attention_heads=heads,
                      num_key_value_heads=heads, vocab_size=320, max_position_embeddings=64)
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg).eval()


def test_patch_api_on_a_random_init_llama():
    model = _tiny_llama()
    with pytest.raises(ValueError, match="num_layers_to_patch"):
        bc.patch_mlp_with_block_circulant(model, num_layers_to_patch=-1, block_size=64, verbose=False)
    with pytest.raises(ValueError, match="patch_position"):
        bc.patch_mlp_with_block_circulant(model, num_layers_to_patch=1, block_size=64, patch_position="middle", verbose=False)
    with contextlib.redirect_stdout(io.StringIO()) as out:
        assert bc.patch_mlp_with_block_circulant(model, num_layers_to_patch=2, block_size=64, patch_position="last") is None
    assert "Patching 2 layer(s) from last: [1, 2]" in out.getvalue()
    kinds = [[type(getattr(l.mlp, n)).__name__ for n in ("gate_proj", "up_proj", "down_proj")] for l in model.model.layers]
    assert kinds[0] == ["Linear"] * 3 and kinds[1] == kinds[2] == ["BlockCirculantLinear"] * 3
    assert type(model.model.layers[2].self_attn.q_proj).__name__ == "Linear"
    assert bc.patch_llama_fft(model, block_size=64) == 4 * 3 + 3          # q,k,v,o everywhere + layer 0's MLP
    assert bc.enable_cfft_cache(model) == 21 and bc.disable_cfft_cache(model) == 21
    # the projection of layer 2's gate_proj equals the oracle's from_linear on the same dense weight
    from oracle import bc_oracle as O
    ref = _tiny_llama()
    W = ref.model.layers[2].mlp.gate_proj.weight.data
    got = model.model.layers[2].mlp.gate_proj.c.detach()
    assert (got - O.from_linear_c(W, 64)).abs().max() <= 1e-6 * got.abs().max()


def test_shared_input_groups_attach_without_changing_the_module_tree():
    """fuse_shared_inputs: q/k/v and gate/up of every patched layer form one group each; the members stay the
    model's modules (state-dict keys, checkpoint I/O unchanged) and unfuse detaches the groups again."""
    model = _tiny_llama(layers=2)
    bc.patch_llama_fft(model, block_size=64)
    keys = sorted(model.state_dict().keys())
    assert bc.fuse_shared_inputs(model) == 2 * 2
    assert sorted(model.state_dict().keys()) == keys
    attn, mlp = model.model.layers[1].self_attn, model.model.layers[1].mlp
    g = attn.q_proj._shared_group
    assert g is attn.k_proj._shared_group is attn.v_proj._shared_group and attn.o_proj._shared_group is None
    assert g.out_blocks == 3 * (128 // 64) and g.offsets == [0, 128, 256, 384]
    assert mlp.gate_proj._shared_group is mlp.up_proj._shared_group and mlp.down_proj._shared_group is None
    assert mlp.gate_proj._shared_group.offsets == [0, 256, 512]
    with pytest.raises(ValueError, match="share"):
        bc.SharedInputGroup([attn.q_proj, mlp.down_proj])
    with pytest.raises(RuntimeError):                                # still no CPU path: the fused call fails as loudly as a single layer
        with torch.no_grad():
            attn.q_proj(torch.randn(3, 128))
    assert bc.unfuse_shared_inputs(model) == 4 and attn.q_proj._shared_group is None


def test_bench_all_in_one_metric_helpers_cpu():
    """Metric definitions of the unified bench (reference src/bench_all_in_one.py:112-120, :297-346) on a stand-in model."""
    from llama_fft_b200 import bench_all_in_one as bai
    torch.manual_seed(0)
    s, t = torch.randn(3, 50), torch.randn(3, 50)
    want = torch.nn.functional.kl_div(torch.log_softmax(s / 2, -1), torch.softmax(t / 2, -1), reduction="none").sum(-1)
    assert torch.allclose(bai.softmax_kl(s, t, temperature=2.0), want, atol=1e-6)
    toks = bai.byte_tokenize(["ab", "", "x" * 500], max_len=8, vocab_size=320)
    assert [int(i.numel()) for i, _ in toks] == [2, 1, 8] and toks[0][0].tolist() == [3 + 97, 3 + 98]
    batches = list(bai.iter_batches(toks, 2, pad_id=0))
    assert batches[0][0].shape == (2, 2) and batches[0][2].tolist() == [1, 0] and batches[1][0].shape == (1, 8)

    class Fake(torch.nn.Module):
        def __init__(self, shift):
            super().__init__()
            self.emb = torch.nn.Embedding(320, 16)
            self.head = torch.nn.Linear(16, 40)
            self.shift = shift

        def forward(self, input_ids, attention_mask=None):
            class Out:
                pass
            o = Out()
            o.logits = self.head(self.emb(input_ids)) + self.shift
            return o
    torch.manual_seed(1)
    teacher = Fake(0.0)
    cache = bai.compute_teacher_last_cache(teacher, toks, torch.device("cpu"))
    same = bai.correctness_last_token(teacher, toks, cache, torch.device("cpu"), 1.0, 5)
    assert same["last_mse"] == 0.0 and abs(same["last_kl"]) < 1e-6 and same["last_top1_acc"] == 1.0 and same["last_topk_overlap"] == 1.0
    teacher.shift = 0.5                                  # a constant logit shift: MSE 0.25, distribution unchanged
    moved = bai.correctness_last_token(teacher, toks, cache, torch.device("cpu"), 1.0, 5)
    assert abs(moved["last_mse"] - 0.25) < 1e-5 and abs(moved["last_kl"]) < 1e-5 and moved["last_top1_acc"] == 1.0
    perf = bai.perf_prefill(teacher, toks, torch.device("cpu"), 0, 2, 1, 2)
    assert perf["prefill_tokens_per_s"] > 0 and set(perf) >= {"prefill_avg_ms", "prefill_tokens_per_s", "peak_mem_bytes"}


def test_legacy_convention_detector():
    """reference src/patch_llama_fft.py:17-44.  "diag_inv" (c[k] = mean_i W[(i+k)%B, i]) is exact for every circulant block, a
    transposed one included, so the detector must pick it there; on random blocks it must agree with a slow restatement."""
    torch.manual_seed(3)
    B = 32
    c = torch.randn(B)
    C = bc.circulant_from_first_col(c)
    W = torch.randn(2 * B, 3 * B)
    for blk in (C, C.t()):
        W[:B, :B] = blk
        assert bc._detect_best_convention_for_layer(W, B) == "diag_inv"
    assert torch.allclose(bc.dense_block_to_circulant_column(C, convention="diag_inv"), c, atol=1e-6)
    assert torch.allclose(bc.dense_block_to_circulant_column(C, convention="diag"), c[(-torch.arange(B)) % B], atol=1e-6)

    def slow(Wb):
        torch.manual_seed(0)
        xs = [torch.randn(B) for _ in range(3)]
        errs = {}
        for conv in ("diag", "diag_inv"):
            col = torch.stack([torch.stack([Wb[i, (i + k) % B] if conv == "diag" else Wb[(i + k) % B, i] for i in range(B)]).mean() for k in range(B)])
            Cm = torch.stack([torch.stack([col[(i - j) % B] for j in range(B)]) for i in range(B)])
            errs[conv] = sum(((Wb @ x) - (Cm @ x)).pow(2).mean().item() for x in xs) / 3.0
        return "diag" if errs["diag"] <= errs["diag_inv"] else "diag_inv"
    seen = set()
    for seed in range(6):
        Wr = torch.randn(B, B, generator=torch.Generator().manual_seed(seed))
        got = bc._detect_best_convention_for_layer(Wr, B)
        assert got == slow(Wr)
        seen.add(got)
    assert seen <= {"diag", "diag_inv"}


def test_work_based_switch_to_the_tensor_core_path(monkeypatch):
    """Without the BCL_PREFILL_MIN_TOKENS* overrides a layer takes the tcgen05 path from a token count that shrinks with its size
    (stream kernel: ~17 us + 1.05e-4 us per token x block pair; tensor-core chain: flat ~24 / ~33 us)."""
    import importlib
    P = importlib.import_module("llama_fft_b200.patch_llama_fft")      # (the package attribute of that name is the function)
    monkeypatch.setattr(P, "PREFILL_MIN_TOKENS", None)
    monkeypatch.setattr(P, "PREFILL_MIN_TOKENS_F32", None)
    assert P.prefill_token_floor(32, 32, torch.bfloat16) == 64 and P.prefill_token_floor(86, 32, torch.bfloat16) == 24
    assert P.prefill_token_floor(32, 32, torch.float32) == 157 and P.prefill_token_floor(32, 86, torch.float32) == 59
    assert P.prefill_token_floor(2, 2, torch.bfloat16) == 16384 and P.prefill_token_floor(500, 500, torch.bfloat16) == 16
    m = bc.BlockCirculantLinear(4096, 4096, block_size=128, bias=False, spectrum_dtype=torch.bfloat16)
    assert not m._use_prefill_path(63) and m._use_prefill_path(64)
    monkeypatch.setattr(P, "PREFILL_MIN_TOKENS", 256)                      # the override is a plain token floor
    assert not m._use_prefill_path(255) and m._use_prefill_path(256)


def test_from_linear_follows_the_replaced_modules_mode_and_disable_cache_flag():
    """a model put in eval() before patching must end up with eval-mode BC layers (the cached spectrum and the
    shared-input groups are eval-only); disable_cfft_cache switches to rebuild-on-every-forward."""
    import torch.nn as nn
    lin = nn.Linear(128, 64, bias=True)
    assert bc.BlockCirculantLinear.from_linear(lin, block_size=32).training
    m = bc.BlockCirculantLinear.from_linear(lin.eval(), block_size=32)
    assert not m.training and m._cache_spectrum
    holder = nn.Sequential(m)
    assert bc.disable_cfft_cache(holder) == 1 and not m._cache_spectrum
    assert bc.enable_cfft_cache(holder) == 1 and m._cache_spectrum


def test_calibrate_default_prompts_file_exists():
    from llama_fft_b200 import calibrate_bc
    import os
    assert os.path.exists(calibrate_bc.build_parser().parse_args([]).prompts_file)


def test_lm_eval_adapter_on_cpu_without_bc_layers_matches_direct_scores():
    """num_layers=0 leaves the model dense, so the adapter's request handling (tokenisation, windowing, greedy flag) can be checked on CPU."""
    import math
    from llama_fft_b200.lm_eval_adapter import LlamaFftLM
    lm = LlamaFftLM(pretrained="", random_init="tiny", dtype="float32", device="cpu", num_layers=0, max_length=16)
    ctx, cont = "abc", "defg"
    (ll, greedy), = lm.loglikelihood([(ctx, cont)])
    ids = lm.tok_encode(ctx) + lm.tok_encode(cont)
    with torch.no_grad():
        logits = lm.model(input_ids=torch.tensor([ids[:-1]])).logits[0]
    want = torch.log_softmax(logits[-4:], -1).gather(-1, torch.tensor(ids[-4:])[:, None]).sum().item()
    assert math.isclose(ll, want, rel_tol=1e-5, abs_tol=1e-5)
    long_text = "x" * 40                                              # 40 tokens through windows of 15
    (roll,) = lm.loglikelihood_rolling([(long_text,)])
    assert roll < 0 and lm.loglikelihood([("", "a")])[0][0] < 0


def test_fuse_mlp_attaches_and_detaches_on_cpu():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2, vocab_size=64)
    model = LlamaForCausalLM(cfg).eval()
    bc.patch_llama_fft(model, block_size=64)
    assert bc.fuse_mlp(model) == 2 and bc.fuse_mlp(model) == 0
    mlp = model.model.layers[0].mlp
    assert not mlp._bcl_fused_mlp.eligible(torch.zeros(1, 1, 128))        # CPU tensor: the original forward (which then refuses: no CPU path)
    assert bc.unfuse_mlp(model) == 2 and "forward" not in mlp.__dict__
