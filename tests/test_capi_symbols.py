"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/gdtb.h declares,
and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gdtb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gdtb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(gdt):
    names = declared_symbols()
    assert len(names) > 40
    assert set(names) == set(gdt.capi.PROTOTYPES), set(names) ^ set(gdt.capi.PROTOTYPES)


def test_library_exports_every_declared_symbol(gdt):
    handle = C.CDLL(gdt.capi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), f"libgdtb.so does not export {name}"


def test_library_is_built_for_sm_100a_only(gdt):
    import subprocess

    out = subprocess.run(["cuobjdump", "--list-elf", gdt.capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_descriptor_layouts_match_the_header(gdt):
    D = gdt.descriptors
    assert C.sizeof(D.GridDesc) == 4 + 4 + 24 + 24 + 24
    assert C.sizeof(D.Function) == 16 + 72 + 64 + 8 + 16
    assert C.sizeof(D.Integrand) == 16 + 2 * C.sizeof(D.Function)
    assert C.sizeof(D.Form) == 16 + 4 * C.sizeof(D.Integrand)
    assert C.sizeof(D.Flux) == 8 + 32


def test_no_cpu_fallback(gdt):
    """Without a GPU the product must refuse to compute instead of silently falling back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gdt.capi.CudaError):
        gdt.Context(0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under dune-gdt_b200/ or dune_gdt_b200/ may reference it."""
    bad = []
    for top in ("dune-gdt_b200", "dune_gdt_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".hpp", ".hh", ".h", ".cc", ".cpp")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    uses = (
                        re.search(r"^\s*(import|from)\s+oracle\b", text, re.M)
                        or re.search(r"#\s*include\s*[<\"][^>\"]*oracle", text)
                        or "liboracle" in text
                        or re.search(r"import_module\(\s*[\"']oracle", text)
                    )
                    if uses:
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
