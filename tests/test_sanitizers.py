"""ASan + UBSan over the device routines compiled for the host (the CUDA kernels wrap exactly this source)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _lib(name):
    p = subprocess.run(["gcc", f"-print-file-name={name}"], capture_output=True, text=True).stdout.strip()
    return p if os.path.isabs(p) and os.path.exists(p) else None


def test_device_routines_under_asan_ubsan(tmp_path):
    asan, ubsan = _lib("libasan.so"), _lib("libubsan.so")
    if not asan or not ubsan:
        pytest.skip("libasan / libubsan not installed")
    so = str(tmp_path / "libkzemu_asan.so")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-mfma", "-pthread", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                           "-shared", "-o", so, os.path.join(HERE, "hostemu", "emu.cpp")])
    env = dict(os.environ, LD_PRELOAD=f"{asan}:{ubsan}", ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([sys.executable, os.path.join(HERE, "hostemu", "asan_tour.py"), so], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "asan tour done" in r.stdout
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr
