"""Compatibility names of the reference's ``src/bench_all_in_one.py`` that touch the hot path (its :164-254).
The HF end-to-end harness around them is out of scope."""
from .patch_llama_fft import enable_cfft_cache, disable_cfft_cache  # noqa: F401
