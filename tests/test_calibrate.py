"""calibrate_bc drop-in (llama-fft_b200/calibrate_bc.py vs the reference's src/calibrate_bc.py): CLI surface and helpers on
CPU; the KL calibration loop itself -- forward and backward through the CUDA kernels -- on the GPU with a tiny model."""
import importlib

import pytest
import torch

import llama_fft_b200 as bc

cal = importlib.import_module("llama_fft_b200.calibrate_bc")

# flags and defaults of the reference parser, src/calibrate_bc.py:112-136
REFERENCE_FLAGS = {"prompts_file": None, "limit": 50, "block_sizes": "64,128,256", "num_layers_to_patch": 1, "patch_position": "first",
                   "steps": 200, "lr": 1e-3, "train_bias": False, "batch_size": 4, "max_len": 128, "out_dir": "calib_out", "seed": 0,
                   "teacher_cache": "teacher_last_cache.pt", "token_cache": "tokenized_prompts_cache.pt"}


def test_cli_keeps_reference_flags_and_defaults():
    ns = cal.build_parser().parse_args([])
    for k, v in REFERENCE_FLAGS.items():
        assert hasattr(ns, k), k
        if v is not None:
            assert getattr(ns, k) == v, k
    with pytest.raises(SystemExit):
        cal.build_parser().parse_args(["--patch_position", "middle"])


def test_helpers_follow_reference_contracts(tmp_path):
    items = [(torch.tensor([5, 6, 7]), torch.ones(3, dtype=torch.long)), (torch.tensor([9]), torch.ones(1, dtype=torch.long))]
    ids, mask, last = cal.pad_batch(items, pad_id=2)                     # src/calibrate_bc.py:73-93
    assert ids.tolist() == [[5, 6, 7], [9, 2, 2]] and mask.tolist() == [[1, 1, 1], [1, 0, 0]] and last.tolist() == [2, 0]
    logits = torch.arange(2 * 3 * 4, dtype=torch.float32).view(2, 3, 4)
    assert torch.equal(cal.gather_last_logits(logits, last), torch.stack([logits[0, 2], logits[1, 0]]))   # :96-104
    f = tmp_path / "p.txt"
    f.write_text("# comment\n\n first \nsecond\nthird\n", encoding="utf-8")
    assert cal.load_texts(f, 2) == ["first", "second"]                    # :34-47
    with pytest.raises(FileNotFoundError):
        cal.load_texts(tmp_path / "missing.txt", 1)
    (tmp_path / "e.txt").write_text("#only\n")
    with pytest.raises(ValueError):
        cal.load_texts(tmp_path / "e.txt", 1)
    m = torch.nn.Sequential(torch.nn.Linear(4, 4), bc.BlockCirculantLinear(64, 128, block_size=64, bias=True))
    cal.freeze_all_params(m)
    assert not any(p.requires_grad for p in m.parameters())
    assert cal.enable_bc_trainable(m, train_bias=False) == 2 * 64       # :55-65
    assert cal.enable_bc_trainable(m, train_bias=True) == 2 * 64 + 128
    assert sum(p.numel() for p in cal.iter_trainable_params(m)) == 2 * 64 + 128
    tok = cal.ByteTokenizer()
    enc = tok("hé", max_length=8)
    assert enc["input_ids"].tolist() == [[1, 3 + 0x68, 3 + 0xC3, 3 + 0xA9]] and enc["attention_mask"].sum() == 4


@pytest.mark.gpu
def test_calibration_loop_reduces_kl_on_tiny_llama(tmp_path):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = dict(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=4,
               max_position_embeddings=64, tie_word_embeddings=False)
    torch.manual_seed(3)
    teacher = LlamaForCausalLM(LlamaConfig(**cfg)).to("cuda:0").eval()
    student = LlamaForCausalLM(LlamaConfig(**cfg))
    student.load_state_dict(teacher.state_dict())
    student = student.to("cuda:0").eval()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        bc.patch_mlp_with_block_circulant(student, num_layers_to_patch=2, block_size=64, patch_position="first")
    tok = cal.ByteTokenizer()
    texts = ["the quick brown fox", "jumps over", "the lazy dog again and again", "block circulant", "fft", "llama two"]
    ids, masks = cal.tokenize_prompts(tok, texts, 24, tmp_path / "tok.pt")
    ids2, _ = cal.tokenize_prompts(tok, ["ignored"], 24, tmp_path / "tok.pt")       # second call reads the cache
    assert len(ids2) == len(texts)
    tl = cal.teacher_last_logits(teacher, ids, masks, tmp_path / "teacher.pt")
    assert len(tl) == len(texts) and tl[0].shape == (512,) and tl[0].dtype == torch.float32
    losses = cal.run_calibration(student, ids, masks, tl, pad_id=tok.pad_token_id, steps=40, lr=2e-3, batch_size=3, verbose=False)
    assert all(l == l for l in losses)                                   # finite
    assert sum(losses[-6:]) / 6 < 0.7 * sum(losses[:6]) / 6, (losses[:6], losses[-6:])
    trainable = [n for n, p in student.named_parameters() if p.requires_grad]
    assert trainable and all(n.endswith(".c") for n in trainable)
    out = tmp_path / "bc.pt"
    with contextlib.redirect_stdout(io.StringIO()):
        bc.save_bc_params(student, str(out))
    state = torch.load(out)
    assert len(state) == 6 and all(k.endswith(".c") for k in state)     # 2 layers x gate/up/down, reference checkpoint format
