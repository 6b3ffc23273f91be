"""The C-ABI library loads without a GPU, exports every symbol include/clrs_b200.h declares, and
fails loudly (no CPU fallback) when asked to compute without an sm_100 device."""
import ctypes as C
import os
import re

import pytest

import clrs_b200
from clrs_b200 import workloads, Options

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "clrs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clrs_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = clrs_b200.load_library("device")
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_options_struct_matches_header_layout():
    lib = clrs_b200.load_library("device")
    o = Options()
    lib.clrs_default_options(C.byref(o))
    assert o.prec == 256 and o.gamma == 0.9 and o.beta_infeasible == 0.3 and o.beta_feasible == 0.1
    assert o.omega_p == 1e10 and o.duality_gap_threshold == 1e-15 and o.step_length_threshold == 1e-7 and o.safe_step == 1


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_a_gpu():
    with pytest.raises(RuntimeError) as e:
        clrs_b200.Solver(workloads.maxcut(workloads.laplacian_cycle(3)), lib="device")
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_package_cannot_reach_the_oracle_without_registration():
    """The host mirror binds only the CUDA library; lib="oracle" exists only after test infrastructure registers it."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, '.'); import clrs_b200; from clrs_b200 import workloads, Solver\n"
            "try:\n    Solver(workloads.maxcut(workloads.laplacian_cycle(3)), lib='oracle'); print('reachable')\n"
            "except RuntimeError as e:\n    print('refused')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=300)
    assert out.stdout.strip() == "refused", (out.stdout, out.stderr[-500:])


def test_ctypes_structs_have_the_layout_the_c_compiler_gives_the_header(tmp_path):
    """sizeof / offsetof of clrs_options and clrs_iter_info as gcc sees include/clrs_b200.h, against the ctypes mirrors in api.py
    (the Julia structs in INTEGRATION.md list the same fields in the same order)."""
    import subprocess
    from clrs_b200 import IterInfo
    fields_o = [n for n, _ in Options._fields_]
    fields_i = [n for n, _ in IterInfo._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "clrs_b200.h")}"', 'int main(void) {',
           '  printf("%zu %zu\\n", sizeof(clrs_options), sizeof(clrs_iter_info));']
    src += [f'  printf("%zu\\n", offsetof(clrs_options, {n}));' for n in fields_o]
    src += [f'  printf("%zu\\n", offsetof(clrs_iter_info, {n}));' for n in fields_i]
    src += ['  return 0; }']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(c)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert [int(out[0]), int(out[1])] == [C.sizeof(Options), C.sizeof(IterInfo)]
    offs = [int(v) for v in out[2:]]
    assert offs[:len(fields_o)] == [getattr(Options, n).offset for n in fields_o]
    assert offs[len(fields_o):] == [getattr(IterInfo, n).offset for n in fields_i]
