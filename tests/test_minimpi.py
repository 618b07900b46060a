"""The bundled mini-MPI (control plane only): its own self-test on 1, 2, 4 and 7 ranks."""
import os
import subprocess

import pytest

from util import MINIMPIRUN, PKG

ROOT = os.path.dirname(PKG)


@pytest.fixture(scope="module")
def selftest(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("minimpi") / "selftest")
    subprocess.check_call(["gcc", "-O1", "-g", "-o", out, os.path.join(PKG, "minimpi", "selftest.c"), "-I" + os.path.join(PKG, "minimpi"),
                           "-L" + os.path.join(PKG, "lib"), "-lminimpi", "-Wl,-rpath," + os.path.join(PKG, "lib")])
    return out


@pytest.mark.parametrize("nproc", [1, 2, 4, 7])
def test_selftest(selftest, nproc):
    r = subprocess.run([MINIMPIRUN, "-np", str(nproc), selftest], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"OK on {nproc} ranks" in r.stdout


def test_failing_rank_does_not_hang(tmp_path):
    """A rank that dies takes the job down with a non-zero status instead of leaving the others waiting."""
    src = os.path.join(str(tmp_path), "die.c")
    with open(src, "w") as f:
        f.write('#include <mpi.h>\n#include <stdlib.h>\nint main(int c,char**v){int r;MPI_Init(&c,&v);MPI_Comm_rank(MPI_COMM_WORLD,&r);'
                'if(r==1)exit(3);MPI_Barrier(MPI_COMM_WORLD);MPI_Barrier(MPI_COMM_WORLD);MPI_Finalize();return 0;}\n')
    exe = os.path.join(str(tmp_path), "die")
    subprocess.check_call(["gcc", "-o", exe, src, "-I" + os.path.join(PKG, "minimpi"), "-L" + os.path.join(PKG, "lib"), "-lminimpi",
                           "-Wl,-rpath," + os.path.join(PKG, "lib")])
    r = subprocess.run([MINIMPIRUN, "-np", "3", exe], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
