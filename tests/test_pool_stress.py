"""Builds tests/pool_stress.cpp (host pool + NoiseModel::fold_run, no CUDA) with AddressSanitizer / UBSan and runs it:
20 000 back-to-back parallel_for calls of different sizes with every item checked, then runs of frames folded with
helper threads against the frame-by-frame fold on a stream with scene cuts (tables compared byte for byte)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("flags", [["-O2"], ["-O1", "-g", "-fsanitize=address,undefined"]], ids=["plain", "asan_ubsan"])
def test_pool_and_fold_run_stress(tmp_path, flags):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "pool_stress")
    cuda_inc = "/usr/local/cuda/include"   # g1s_model.cpp includes the kernels' header for shared constants only
    cmd = [cxx, *flags, "-std=c++17", "-ffp-contract=off", "-fno-math-errno", "-I", os.path.join(ROOT, "include"),
           "-I", cuda_inc, "-o", exe, os.path.join(ROOT, "tests", "pool_stress.cpp"),
           os.path.join(ROOT, "grav1synth_b200", "csrc", "g1s_model.cpp"), "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in " ".join(flags):
        pytest.skip("sanitizer runtime not available: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert r.stdout.strip().endswith("ok")
