"""Behaviour of include/dev_type.h (reference src/dev_type.c:13-150), MALLOC_ATTACH_WORKBUF (reference src/dev_type.h:63-88) and the
caller-provided work-buffer path of mat_redist (reference src/mat_redist.c:239-267), through a C caller (tests/c/test_dev_type.c)
compiled against include/ + libcrpspmm.so exactly like a reference driver would be."""
import os
import subprocess

import pytest

from util import MINIMPIRUN, PKG, ROOT, run_cmd

SRC = os.path.join(ROOT, "tests", "c", "test_dev_type.c")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("devtype") / "test_dev_type.exe")
    lib = os.path.join(PKG, "lib")
    cmd = ["gcc", "-O1", "-g", "-std=gnu11", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "minimpi"), SRC, "-o", out,
           "-L" + lib, "-lcrpspmm", "-lminimpi", "-lm", "-Wl,-rpath," + lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def run(exe, nproc, dev_type, env=None):
    r = run_cmd([MINIMPIRUN, "-np", str(nproc), exe, str(dev_type)], env=dict(os.environ, **(env or {})), timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("DEVTYPE OK") == nproc, r.stdout + r.stderr


@pytest.mark.parametrize("nproc", [1, 3])
def test_dev_type_host(exe, nproc):
    run(exe, nproc, 0)


def test_cuda_types_invalid_without_device(exe):
    """No GPU here: is_dev_type_valid(DEV_TYPE_CUDA) is 0 and the C caller's CHECK on it fails (exit 3) - never a silent host fallback."""
    import ctypes
    try:
        ctypes.CDLL("libcuda.so.1")
        have = subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0
    except OSError:
        have = False
    if have:
        pytest.skip("a GPU is visible")
    r = run_cmd([MINIMPIRUN, "-np", "1", exe, "1"], timeout=120)
    assert r.returncode != 0 and "not usable here" in r.stderr, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("nproc,dev_type", [(1, 1), (2, 1), (3, 2)])
def test_dev_type_cuda(exe, nproc, dev_type):
    run(exe, nproc, dev_type)
