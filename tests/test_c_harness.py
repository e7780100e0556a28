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


def _build(tmp_path, name="c_abi_harness"):
    exe = str(tmp_path / name)
    libdir = os.path.join(ROOT, "wrf_model_cuda_sample_b200")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-O1", os.path.join(ROOT, "tests", name + ".c"), "-I", os.path.join(ROOT, "include"),
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


def test_comm_harness_builds_and_rejects_bad_usage(tmp_path):
    exe = _build(tmp_path, "c_comm_harness")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("px,py", [(1, 2), (2, 2)])
def test_comm_harness_two_processes_match_oracle(tmp_path, px, py):
    """One PROCESS per rank (fork), CUDA IPC between them, no Python and no NCCL in the loop.  With fewer
    GPUs than ranks the ranks share a device (the driver time-slices their contexts)."""
    import torch
    from wrf_model_cuda_sample_b200 import parallel
    exe = _build(tmp_path, "c_comm_harness")
    nx, ny, nz, nsteps = 200, 96, 14, 3
    out = tmp_path / "out"
    out.mkdir()
    env = dict(os.environ, WRFB200_FLAG_TIMEOUT_MS="20000")
    r = subprocess.run([exe, str(px), str(py), str(nx), str(ny), str(nz), str(nsteps),
                        str(max(1, torch.cuda.device_count())), str(out)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    G = cases.grid(nx, ny, nz, halo=5, variant="specified")
    whole = wrf.synth_fields(G, seed=99)
    want = cases.oracle_loop(G, whole, cases.SCALARS_3KM, nsteps, c=0.25)
    decomp = parallel.Decomposition(G, px, py, halo=3)
    for rank in range(px * py):
        pg = decomp.patch_grid(rank)
        f = {}
        for n in cases.OUTPUTS + ("u", "v"):
            a = np.fromfile(out / f"rank{rank}_field{wrf.FIELD_ID[n]}.bin", dtype=np.float32)
            f[n] = a.reshape(pg.shape_of(n))
        bad = cases.patch_mismatches(f, want, G, pg, decomp.patch_extents(rank))
        assert not bad, f"rank {rank}: {bad}"
