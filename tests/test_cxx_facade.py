"""The C++ facade (dune-gdt_b200/include/dune/gdt/b200.hh): drivers written with dune-gdt's own class / function names
compile against the C ABI with the host compiler alone, and (on a GPU box) reproduce size-independent properties."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "dune-gdt_b200", "examples")
PROGS = ["stationary-heat-equation", "linear-transport-fv", "elliptic-swipdg", "generic-function-check", "parallel-slabs", "euler-2d"]


def build_examples():
    subprocess.check_call(["make", "-C", EXAMPLES], stdout=subprocess.DEVNULL)


def test_examples_compile_against_the_c_abi(gdt):
    build_examples()
    for prog in PROGS:
        path = os.path.join(EXAMPLES, prog)
        assert os.access(path, os.X_OK)
        needed = subprocess.check_output(["readelf", "-d", path], text=True)
        assert "libgdtb.so" in needed  # the drivers go through the C ABI, nothing else


def test_examples_fail_loudly_without_a_gpu(gdt):
    import torch

    if torch.cuda.is_available():
        pytest.skip("only meaningful on a box without a GPU")
    build_examples()
    r = subprocess.run([os.path.join(EXAMPLES, PROGS[0]), "8"], capture_output=True, text=True)
    assert r.returncode != 0 and "DUNE reported error" in r.stderr  # Exceptions::device_error, no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("args", [["stationary-heat-equation", "128", "2"], ["stationary-heat-equation", "48", "3"],
                                  ["linear-transport-fv", "1024"],
                                  ["elliptic-swipdg"],  # the reference's ESV2007 H^1 table on 8^2, 16^2, 32^2 (3 digits)
                                  ["elliptic-swipdg", "256"],
                                  # GenericFunction lambdas (sampled by the facade) vs built-in / constant coefficients
                                  ["generic-function-check"],
                                  # multi-GPU entry points of the facade with one rank (a periodic slab is its own neighbour)
                                  ["parallel-slabs"],
                                  # the reference's 2d_euler driver (systems, m = 4): conservation, positivity, symmetry
                                  ["euler-2d", "64", "0.25"]])
def test_examples_run(gdt, args):
    build_examples()
    r = subprocess.run([os.path.join(EXAMPLES, args[0])] + args[1:], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK")


@pytest.mark.gpu
def test_parallel_facade_two_gpus(gdt):
    """Parallel::SlabAssembler / HaloSlabAssembler / PeerMemoryRungeKuttaTimeStepper / PeerMemoryEulerTimeLoop with one
    C++ process per GPU (tools/mprun.sh; handles exchanged through Parallel::FileRendezvous)"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    build_examples()
    r = subprocess.run([os.path.join(ROOT, "tools", "mprun.sh"), "2", os.path.join(EXAMPLES, "parallel-slabs")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout and "FAILED" not in r.stdout
