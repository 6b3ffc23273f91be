"""examples/maxcut3.c: the C ABI used from plain C (no Python, no torch) — builds against include/clrs_b200.h and the shared library;
on a B200 it must reach the reference's known answer 9/4 (README.md:70-72) through the dense and the triplet upload."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "clusteredlowranksolver.jl_b200", "csrc")


def _build(tmp_path):
    exe = str(tmp_path / "maxcut3")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "maxcut3.c"), "-o", exe,
                    "-L", CSRC, "-lclrs_b200", "-lm", f"-Wl,-rpath,{CSRC}"], check=True)
    return exe


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_c_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_c_example_reaches_nine_quarters_on_the_device(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "MAXCUT3 OK" in out.stdout, (out.stdout, out.stderr)
