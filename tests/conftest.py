"""pytest configuration: the `gpu` marker, repo-root imports, shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests never silently pass on a CPU box: they are skipped (not run) there, and on a GPU box
    they fail loudly if the CUDA library is missing (see llama_fft_b200._cabi)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# The tests use small layers; the library switches to the tensor-core path on WORK (tokens x p x q), which such layers only reach at
# thousands of tokens.  Pin the fixed token floors (the documented override) so that "N >= 256" keeps meaning "tensor-core path" here,
# in this process and in the child processes some tests start.  tests/test_host_cpu.py checks the work-based rule itself.
import os  # noqa: E402

os.environ.setdefault("BCL_PREFILL_MIN_TOKENS", "256")
os.environ.setdefault("BCL_PREFILL_MIN_TOKENS_F32", "256")
