"""CPU: the C-ABI shared library builds, loads without a GPU and exports exactly the symbols include/tris_sm100.h
declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tris_sm100.h")


@pytest.fixture(scope="module")
def lib_path():
    from tris_b200.build import build_lib
    return build_lib()


def declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tris_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_and_library_exports(lib_path):
    names = declared()
    assert len(names) >= 40 and "tris_gemm" in names and "tris_stage1_loss_fwd" in names
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (tris_[a-z0-9_]+)$", out, flags=re.M)))
    undeclared = [n for n in exported if n not in names]
    assert not undeclared, f"exported but not declared in include/tris_sm100.h: {undeclared}"


def test_library_loads_and_reports_without_gpu(lib_path):
    lib = ctypes.CDLL(lib_path)
    lib.tris_last_error.restype = ctypes.c_char_p
    assert lib.tris_abi_version() >= 1
    import torch
    if not torch.cuda.is_available():
        assert lib.tris_check_device() != 0          # fails loudly: no device
        assert len(lib.tris_last_error()) > 0


def test_library_is_sm100a_only(lib_path):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_streaming_kernels_fit_one_resident_wave(lib_path):
    """The BatchNorm streaming kernels are launched as ONE resident wave of 4 CTAs/SM x 256 threads: that only holds at <= 64
    registers per thread (a 74-register build of bn_apply_fwd ran 3 CTAs/SM and cost 0.4 ms per step), and the GEMM
    kernel is capped at 96 registers (352 threads + 214 KB of shared memory per SM).  Checked on the cubin, no GPU needed."""
    import re
    import shutil
    import subprocess
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-res-usage", lib_path], capture_output=True, text=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    assert regs, "no resource usage parsed"
    seen = 0
    for name, (r, stack) in regs.items():
        if re.search(r"bn_apply_fwd(_unpair)?_kernel|bn_bwd_reduce_kernel|bn_bwd_apply_kernel", name):
            seen += 1
            assert r <= 64, (name, r)
        if "tris_umma_gemm_kernel" in name:
            assert r <= 96 and stack <= 32, (name, r, stack)
    assert seen >= 6


def test_product_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import argparse
    from tris_b200 import _lib as L
    from tris_b200.model_stage1 import TRIS
    args = argparse.Namespace(bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024,
                              attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)
    m = TRIS(args)
    with pytest.raises(L.TrisLibError):
        m(torch.zeros(1, 3, 320, 320), torch.zeros(1, 20, dtype=torch.int32))
    with pytest.raises(L.TrisLibError):
        L.require_device()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under tris_b200/ may import it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tris_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
