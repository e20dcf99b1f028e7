"""CPU: the PRMT look-up network of the fused kernels (csrc/nf4_lut.cuh) compiled for the host and checked
against the plain decode (tests/host/lut_check.cu)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not found")
def test_prmt_lut_matches_plain_decode(tmp_path):
    exe = str(tmp_path / "lut_check")
    src = os.path.join(HERE, "host", "lut_check.cu")
    build = subprocess.run(["nvcc", "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe, src],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0 and "mismatches=0" in run.stdout, run.stdout + run.stderr
