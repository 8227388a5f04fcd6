"""GPU test (-m gpu) of the unified bench harness (llama-fft_b200/bench_all_in_one.py) on a random-init tiny Llama: the CLI runs
end to end through the CUDA path, emits the reference's result keys, and fused / unfused runs agree."""
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_all_in_one_cli_on_tiny_random_init_llama(tmp_path):
    from llama_fft_b200 import bench_all_in_one as bai
    out = str(tmp_path / "r.json")
    common = ["--random_init", "tiny", "--prompts_file", os.path.join(ROOT, "tools", "prompts_synthetic.txt"), "--limit", "6", "--max_len", "96",
              "--block_sizes", "64,128", "--num_layers_to_patch", "2", "--targets", "all", "--batch_size", "3", "--warmup", "1", "--runs", "2",
              "--do_generate", "1", "--max_new_tokens", "4"]
    rows = bai.main(common + ["--out_json", out])
    assert [r["B"] for r in rows] == [0, 64, 128] and os.path.exists(out)
    for r in rows[1:]:
        for k in ("last_mse", "last_kl", "last_cos", "last_top1_acc", "last_top1_in_student_topk", "last_topk_overlap", "prefill_avg_ms",
                  "prefill_tokens_per_s", "peak_mem_bytes", "decode_new_tokens_per_s"):
            assert k in r
        assert r["bc_layers"] == 14 and r["cfft_cached_layers"] == 14 and r["last_mse"] > 0 and r["prefill_tokens_per_s"] > 0
    fused = bai.main(common + ["--fuse_shared_inputs", "1"])
    assert fused[1]["shared_input_groups"] == 4
    assert all(r["shared_input_fused_calls"] > 0 for r in fused[1:])       # under torch.inference_mode, layers in eval mode: the group path ran
    for a, b in zip(rows[1:], fused[1:]):                 # fusion is bit-exact: identical quality metrics
        assert a["last_mse"] == b["last_mse"] and a["last_kl"] == b["last_kl"]
