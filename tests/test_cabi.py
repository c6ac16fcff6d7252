"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
fails loudly without a GPU, and the C++ header compiles against it. No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ikd-tree_b200", "libikd_b200.so")
HDR = os.path.join(ROOT, "include", "ikd_b200.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "ikd-tree_b200"), "-j8"])
    return ctypes.CDLL(LIB)


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ikd_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/ikd_b200.h but not exported: {missing}"


def test_python_binding_covers_the_header():
    import ikd_ctypes as I
    assert sorted(I.SIGNATURES) == declared_symbols()


def test_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    lib.ikd_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float]
    st = lib.ikd_create(ctypes.byref(h), -1, 0.5, 0.6, 0.2)
    lib.ikd_last_error.restype = ctypes.c_char_p
    assert st != 0 and not h.value
    assert b"no CPU fallback" in lib.ikd_last_error()


def test_product_does_not_touch_the_oracle():
    """The product path (library sources, headers, binding) must not reference anything under oracle/."""
    bad = []
    for d in ("ikd-tree_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, d)):
            for f in fs:
                if f.endswith((".cu", ".cuh", ".h", ".py", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"ikdo_|ref_ctypes|libikd_oracle|libikd_ref|#include\s*[\"<][^\">]*oracle|import\s+.*oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_cpp_header_compiles_and_links(lib, tmp_path):
    exe = tmp_path / "demo_api_test"
    cmd = ["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "demo_api_test.cpp"),
           "-o", str(exe), "-L", os.path.join(ROOT, "ikd-tree_b200"), "-likd_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "ikd-tree_b200"), "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-fopenmp", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "concurrent_api_test.cpp"), "-o", str(tmp_path / "concurrent_api_test"),
                           "-L", os.path.join(ROOT, "ikd-tree_b200"), "-likd_b200", "-Wl,-rpath," + os.path.join(ROOT, "ikd-tree_b200"),
                           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stdout  # loud failure, not a silent fallback
