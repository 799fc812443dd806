"""The C-ABI library loads without a GPU and exports every symbol include/goldilocks_b200.h declares
(no compute calls here); without a GPU its entry points fail loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "goldilocks_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    funcs = re.findall(r"GOLDILOCKS_B200_API\s+[^;(]*?\b(goldilocks_\w+)\s*\(", text)
    data = re.findall(r"GOLDILOCKS_B200_API\s+extern\s+[^;]*?\b(goldilocks_\w+)\s*(?:\[[^\]]*\])?\s*[;,]", text)
    more = re.findall(r"GOLDILOCKS_B200_API extern const goldilocks_448_scalar_p (\w+), (\w+);", text)
    for a, b in more:
        data += [a, b]
    return sorted(set(funcs)), sorted(set(data))


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "goldilocks_b200.h"\nint main(void){return sizeof(goldilocks_448_point_s)==256 && sizeof(goldilocks_448_scalar_s)==56 ? 0 : 1;}\n')
    exe = tmp_path / "t"
    assert os.system("gcc -std=c99 -Wall -Werror -I%s -o %s %s" % (os.path.join(ROOT, "include"), exe, src)) == 0
    assert os.system(str(exe)) == 0


def test_library_exports_every_declared_symbol():
    import libgoldilocks_b200 as g
    assert os.path.exists(g.LIB_PATH), "run `make lib` (or __graft_entry__.build()) first"
    lib = C.CDLL(g.LIB_PATH)
    funcs, data = declared_symbols()
    assert len(funcs) >= 80 and "goldilocks_ed448_verify_batch" in funcs and "goldilocks_448_point_base" in data
    missing = [s for s in funcs + data if not hasattr(lib, s)]
    assert not missing, "declared but not exported: %s" % missing


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import libgoldilocks_b200 as g
    lib = g.load()
    a = np.zeros((4, 56), np.uint8)
    with pytest.raises(RuntimeError):
        lib.gf_mul(a, a)
    with pytest.raises(RuntimeError):
        lib.ed448_verify(np.zeros((1, 114), np.uint8), np.zeros((1, 57), np.uint8), [b""])


def test_product_never_links_the_oracle():
    import subprocess
    import libgoldilocks_b200 as g
    out = subprocess.run(["ldd", g.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "goldilocks_ref" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "libgoldilocks_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle/" not in text and "liboracle" not in text and "_hostsim" not in text, f
