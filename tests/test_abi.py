"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol include/rgb200.h declares, and the product
path fails loudly without a GPU (no CPU fallback, no route through oracle/)."""
import os
import re
import subprocess

import pytest

import raygun_b200 as rg

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = rg.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert set(declared) == set(rg.ABI_SYMBOLS)


def test_no_unexpected_dependencies():
    """The product .so links only the CUDA runtime + libc/libstdc++: nothing from oracle/, no torch."""
    out = subprocess.run(["ldd", rg.LIB_PATH], capture_output=True, text=True).stdout
    assert "liboracle" not in out and "torch" not in out
    nm = subprocess.run(["nm", "-D", "--undefined-only", rg.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in nm


def test_product_sources_never_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "raygun_b200")):
        if "_obj" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and '#include "orc_' not in text, f


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rg.RaygunError, match="no CPU fallback"):
        rg.Raytracer(64, 36)


def test_sm100a_cubin_embedded():
    out = subprocess.run(["cuobjdump", "-lelf", rg.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
