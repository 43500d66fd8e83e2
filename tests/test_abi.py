"""CPU tests of the drop-in boundary: libsarw_b200.so loads and exports every symbol include/sarw_abi.h declares;
without a GPU every entry point fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re
import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sarw_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sarw_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(sarw):
    lib = sarw.load_library()
    names = _declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(sarw.ABI_SYMBOLS) == names


def test_target_count_needs_no_gpu(sarw):
    assert sarw.target_count((10, 10, 40), 0.92) == 158
    assert sarw.target_count((632.2, 632.2, 632.2), 0.92) == 9999171
    assert sarw.target_count((2934.5, 2934.5, 2934.5), 0.92) > 2 ** 31 / 3   # 64-bit volume (sarw.h:103 overflows here)


def test_parameter_validation_and_no_cpu_fallback(sarw):
    with pytest.raises(sarw.SarwError):
        sarw.Context(nChains=0)
    with pytest.raises(sarw.SarwError):
        sarw.Context(nChains=1, boxSize=(1.5, 10, 10), minDist=1.0)   # the reference needs floor(L/minDist) >= 2
    for bad in (dict(slabRank=2, slabRanks=2), dict(slabRank=-1, slabRanks=3), dict(slabRank=0, slabRanks=9), dict(slabRank=0, slabRanks=2, slabHalo=-1)):
        with pytest.raises(sarw.SarwError, match="slab"):          # checked before any device is touched
            sarw.Context(nChains=4, boxSize=(10, 10, 10), minDist=1.0, **bad)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(sarw.SarwError, match="no CUDA device|CUDA"):
            sarw.Context(nChains=4, boxSize=(10, 10, 10), minDist=1.0)
        with pytest.raises(sarw.SarwError, match="no CUDA device|CUDA"):   # a slab context has no CPU fallback either
            sarw.Context(nChains=4, boxSize=(10, 10, 10), minDist=1.0, slabRank=0, slabRanks=2)


def test_product_never_imports_oracle():
    """The product package must not include, import or load anything under oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "polymer-sarw_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                continue
            for line in open(os.path.join(dirpath, f), errors="ignore").read().splitlines():
                code = line.split("//")[0].split("#include")[-1] if "#include" in line else line.split("//")[0]
                if re.search(r"#include|^\s*(import|from)\s|dlopen|CDLL", line):
                    assert "oracle" not in code.lower(), (f, line)


def test_header_compiles_as_c_and_cpp(tmp_path):
    """include/sarw_abi.h is the drop-in boundary: it must be usable from plain C (cgo/ctypes-style bindings) and from C++."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src_c = tmp_path / "t.c"; src_cpp = tmp_path / "t.cpp"
    body = '#include "sarw_abi.h"\nint main(void) { sarw_params p; sarw_stats s; (void)p; (void)s; return (int)sizeof(sarw_params) == 0; }\n'
    src_c.write_text(body); src_cpp.write_text(body)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(root, "include"), str(src_c)], check=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(root, "include"), str(src_cpp)], check=True)


def test_python_struct_layouts_match_the_header(sarw, tmp_path):
    """sizeof / field offsets of sarw_params and sarw_stats as the C compiler sees them == the ctypes mirrors"""
    import subprocess
    import ctypes as C
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "l.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sarw_abi.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(sarw_params), '
                   'offsetof(sarw_params, seed), offsetof(sarw_params, launchMode), sizeof(sarw_stats), offsetof(sarw_stats, growMs), offsetof(sarw_stats, cycSub)); return 0; }\n')
    exe = tmp_path / "l"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    P, S = sarw.Params, sarw.Stats
    assert out == [C.sizeof(P), P.seed.offset, P.launchMode.offset, C.sizeof(S), S.growMs.offset, S.cycSub.offset]


def test_c_example_compiles_and_links_against_the_library(tmp_path):
    """examples/minimal.c: a plain-C99 program over include/sarw_abi.h links against libsarw_b200.so (it is run by the GPU suite)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "polymer-sarw_b200", "lib")
    for name in ("minimal", "slabs"):      # slabs.c: one melt over N slab contexts (sarw_mr_grow_local)
        exe = tmp_path / name
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "examples", name + ".c"),
                        "-L", lib, "-lsarw_b200", "-Wl,-rpath," + lib, "-o", str(exe)], check=True)
        assert exe.exists()
    import torch
    if not torch.cuda.is_available():       # without a device the example says so and fails: no CPU fallback
        r = subprocess.run([str(tmp_path / "slabs"), "2"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_c_example_runs(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "polymer-sarw_b200", "lib")
    exe = tmp_path / "minimal"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "minimal.c"),
                    "-L", lib, "-lsarw_b200", "-Wl,-rpath," + lib, "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("176 beads, 156 bonds in 7 rounds (target 158)"), r.stdout
