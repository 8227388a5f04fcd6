#!/usr/bin/env python3
"""Stage the UNMODIFIED reference where the GPU box can reach it.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference (Pinguhat/llama-fft) is pure Python: there is nothing to compile, "building" it is a byte-for-byte copy of the
few modules that hold the hot path from /root/reference/src into the git-ignored oracle/_ref/:

    patch_llama_fft.py      BlockCirculantLinear (:142-275), from_linear, patch/save/load
    fft_utils.py            circulant_matvec_fft (:3-34)
    bench_all_in_one.py     enable_cfft_cache / forward_cached (:164-254) -- the configuration every published number used
    microbench_dense_vs_bc.py, validate_implementation.py      the reference's own harness / tests for this path
    prompts_*.txt           inputs of the end-to-end parity tests

/root/reference does not exist on the GPU box; oracle/_ref/ travels there with the gpurun snapshot like the built .so files
(git-ignored, NOT gpurun-ignored), so the parity tests can run the reference itself on the same B200 and bench.py can time the
reference's own CPU path (cpu_baseline.kind = "reference") and its torch CUDA path (detail.reference_cuda).  Nothing under
llama-fft_b200/ ever imports it; a MANIFEST with sha256 sums is written so a test can prove the copies are unmodified.

    python oracle/make_ref.py            (run by __graft_entry__.build() when /root/reference is present)
"""
import hashlib
import json
import os
import shutil
import sys

SRC = os.environ.get("BCL_REFERENCE_SRC", "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["patch_llama_fft.py", "fft_utils.py", "bench_all_in_one.py", "microbench_dense_vs_bc.py", "validate_implementation.py",
         "prompts_100.txt", "prompts_20_quality.txt", "prompts_calib_200.txt", "prompts_calib_50_long.txt", "prompts_eval_20_quality_long.txt"]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def main():
    if not os.path.isdir(SRC):
        print(f"make_ref: {SRC} not present (GPU box?): keeping whatever oracle/_ref/ holds")
        return 0
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name in FILES:
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        manifest[name] = sha256(os.path.join(DST, name))
        assert manifest[name] == sha256(os.path.join(SRC, name))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1, sort_keys=True)
    print(f"make_ref: {len(FILES)} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
