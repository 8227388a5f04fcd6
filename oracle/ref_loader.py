"""Import the UNMODIFIED reference modules staged in oracle/_ref/ (oracle/make_ref.py).  TEST / BASELINE INFRASTRUCTURE ONLY:
used by tests/, __graft_entry__.smoke() and bench.py's reference legs; never by the product package."""
import hashlib
import importlib
import json
import os
import sys

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "patch_llama_fft.py")) and os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


def verify() -> dict:
    """sha256 of every staged file against the manifest written when it was copied from /root/reference/src."""
    with open(os.path.join(REF_DIR, "MANIFEST.json")) as f:
        man = json.load(f)["sha256"]
    for name, want in man.items():
        with open(os.path.join(REF_DIR, name), "rb") as f:
            got = hashlib.sha256(f.read()).hexdigest()
        if got != want:
            raise RuntimeError(f"oracle/_ref/{name} differs from the reference it was copied from")
    return man


def load(name: str = "patch_llama_fft"):
    """The reference module `name` (its siblings import each other by bare name, so oracle/_ref goes on sys.path)."""
    if not available():
        raise ImportError("oracle/_ref/ is empty: run `python oracle/make_ref.py` in the build container (needs /root/reference)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    mod = importlib.import_module(name)
    if os.path.dirname(os.path.abspath(mod.__file__)) != REF_DIR:
        raise ImportError(f"{name} resolved to {mod.__file__}, not to oracle/_ref/")
    return mod


def prompt_lines(filename: str, limit: int):
    with open(os.path.join(REF_DIR, filename), encoding="utf-8") as f:
        return [ln.strip() for ln in f if ln.strip() and not ln.startswith("#")][:limit]
