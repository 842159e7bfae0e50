"""CPU tier: the C-ABI library loads and exports every symbol include/ronk_b200.h declares; the
host package fails loudly without a GPU (no CPU fallback); metadata entry points (host-side
O(log p) code, no GPU needed) return the reference's constants."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ronk_b200.h")).read()
    return sorted(set(re.findall(r"\b(ronk_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from ronkathon_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    lib = C.CDLL(built.LIB_PATH)
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ronk_b200.h but not exported"
    assert set(built.SIGNATURES) == set(names), set(built.SIGNATURES) ^ set(names)


def test_library_has_sm100a_code_only(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", built.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_metadata_matches_reference_constants(built, kats, gold64):
    lib = built.lib()
    g = C.c_uint64()
    for p, exp in kats["field"]["generator"].items():
        assert lib.ronk_field_generator(int(p), C.byref(g)) == 0 and g.value == exp
    assert lib.ronk_field_generator(built.GOLDILOCKS, C.byref(g)) == 0 and g.value == 7
    assert lib.ronk_field_generator(100, C.byref(g)) == built.EINVAL  # non-prime modulus panics
    w = C.c_uint64()
    for k, exp in gold64["roots"].items():
        assert lib.ronk_root_of_unity(built.GOLDILOCKS, 7, 1 << int(k), C.byref(w)) == 0 and w.value == exp
    for p, n in kats["field"]["no_root_of_unity"]:
        lib.ronk_field_generator(p, C.byref(g))
        assert lib.ronk_root_of_unity(p, g.value, n, C.byref(w)) == built.EINVAL
    assert lib.ronk_root_of_unity(101, 2, 4, C.byref(w)) == 0 and w.value == 10  # ω4 in F101


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.RonkError):
        built.Context()
    import ronkathon_b200 as r
    with pytest.raises(built.RonkError):
        r.PlutoBaseField(40) * r.PlutoBaseField(61)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ronkathon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "ronk_oracle" not in text, f


def test_rust_sys_crate_declares_every_header_symbol():
    """bindings/rust is shipped as source only (no rustc in the build image), so at least keep its extern block
    in step with include/ronk_b200.h: every exported function must be declared there."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "ronk_b200.h")).read()
    names = sorted(set(re.findall(r"\b(ronk_[a-z0-9_]+)\s*\(", header)))
    rs = open(os.path.join(root, "bindings", "rust", "ronkathon-b200-sys", "src", "lib.rs")).read()
    missing = [n for n in names if not re.search(r"\bfn\s+" + n + r"\s*\(", rs)]
    assert not missing, missing
