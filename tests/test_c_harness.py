"""The C program tests/c_abi_harness.c calls the C ABI exactly as the Fortran ISO_C_BINDING shim does
(no Fortran compiler exists in the image).  Without a GPU it must fail loudly; with one its output field
checksums must equal those of the oracle on the same inputs."""
import os
import subprocess

import numpy as np
import pytest

import wrf_model_cuda_sample_b200 as wrf
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_harness")
    libdir = os.path.join(ROOT, "wrf_model_cuda_sample_b200")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, os.path.join(ROOT, "tests", "c_abi_harness.c"), "-I", os.path.join(ROOT, "include"),
                    "-L", libdir, "-lwrfb200", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    return exe


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _fnv1a(a: np.ndarray) -> str:
    h = 1469598103934665603
    for b in a.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


@pytest.mark.skipif(_have_gpu(), reason="GPU present")
def test_harness_fails_loudly_without_gpu(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_harness_matches_oracle(tmp_path):
    from oracle import loader
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(line.split()[1::2] for line in r.stdout.strip().splitlines())
    g = wrf.Grid(1, 40, 1, 30, 12, -2, 43, -2, 33, 1, 12, 1, 40, 1, 30, 1, 12, False, True, False)
    f = wrf.synth_fields(g)
    loader.oracle_c(f, g, cases.SCALARS_12KM)
    for name in cases.OUTPUTS:
        assert got[str(wrf.FIELD_ID[name])] == _fnv1a(f[name]), name
