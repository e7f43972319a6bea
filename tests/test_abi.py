"""not-gpu: the C-ABI library loads and exports every symbol include/nafgpu.h declares; without a GPU
the product fails loudly instead of falling back to anything."""
import ctypes
import os
import re

import pytest

import naf_b200
from naf_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "nafgpu.h")).read()
    return sorted(set(re.findall(r"\b(nafgpu_[a-z0-9_]+)\s*\(", h)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(api.EXPORTS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(naf_b200.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(naf_b200.library_path())
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert lib.nafgpu_version is not None


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(naf_b200.NafGpuError) as e:
        naf_b200.NafGpu(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_oracle():
    """nothing under naf_b200/ may import, link or execute oracle/ (parity would be void)"""
    for dp, _, fs in os.walk(os.path.join(ROOT, "naf_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in src and "oracle/" not in src and "import oracle" not in src, os.path.join(dp, f)


def test_cli_threaded_file_io(tmp_path):
    """cli/cli_common.hpp: pieces written with pwrite by four threads behind a stdio prefix, read back with pread in pieces of
    several sizes (short read at the end of the file), and the stdio path for pipes (tests/emu/cli_io_check.cpp)"""
    import subprocess
    exe = os.path.join(ROOT, "tests", "_build", "cli_io_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "emu", "cli_io_check.cpp"), "-L", os.path.join(ROOT, "naf_b200"), "-lnafgpu",
                    "-Wl,-rpath," + os.path.join(ROOT, "naf_b200"), "-pthread"], check=True)
    p = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.strip() == "ok", (p.returncode, p.stdout, p.stderr)


def test_interleaved_upload_plan_invariants():
    """naf_b200/csrc/duo_plan.hpp (the order in which a FASTQ .naf goes up for the interleaved decode): every byte exactly once,
    whole blocks per piece, bases before their qualities -- 3,000 random block lists (tests/emu/duo_plan_check.cpp)"""
    import subprocess
    exe = os.path.join(ROOT, "tests", "_build", "duo_plan_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "emu", "duo_plan_check.cpp")], check=True)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.strip() == "ok", (p.returncode, p.stdout, p.stderr)
