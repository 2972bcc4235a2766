"""The C ABI is usable from plain C (gcc + dlopen, no Python in the call path):
tests/c_harness/capi_demo.c.  Without a GPU it must stop at cfmm_create with
CFMM_ERR_CUDA (exit code 3); with one it reproduces the reference KAT."""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, "cfmmrouter.jl_b200", "libcfmm_b200.so")


def _build(tmp_path):
    exe = str(tmp_path / "capi_demo")
    subprocess.run(["/usr/bin/gcc", os.path.join(HERE, "c_harness", "capi_demo.c"), "-I",
                    os.path.join(ROOT, "include"), "-ldl", "-o", exe], check=True)
    return exe


def test_c_caller_without_gpu_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([_build(tmp_path), LIB], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU path" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_c_caller_reference_kat(tmp_path):
    r = subprocess.run([_build(tmp_path), LIB], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "bad index rc=-1" in r.stdout and "outside 1..2" in r.stdout
    m = re.search(r"unit pool at v=\(2,1\): D=\(([^,]+), ([^)]+)\) L=\(([^,]+), ([^)]+)\)", r.stdout)
    d1, d2, l1, l2 = (float(x) for x in m.groups())
    # test/cfmms.jl:82-86: Δ = [≈0, √2−1], Λ = [1−√½, ≈0]
    assert d1 == 0.0 and d2 == 0.41421356237309515 and l1 == 0.2928932188134524 and l2 == 0.0
