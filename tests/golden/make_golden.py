"""Mint the golden vectors in tests/golden/ by running the UNMODIFIED reference
sources (built under oracle/_ref by oracle/Makefile against the mini-MPI, the
OpenMP stand-in for MKL and the METIS stub) on the small cases of tests/cases.py.

    python tests/golden/make_golden.py          # needs /root/reference (build container only)

The reference has no golden vectors of its own; these files are
what pins the oracle (oracle/crp_oracle.c) and the library's host planner.
Each <case>.npz holds the input CSR, the run parameters and, per rank r, every
array ref_dump wrote as "r<r>/<field>".
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from pycrp import gen  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def run(cmd, env=None):
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed:\n{r.stdout}\n{r.stderr}")
    return r.stdout


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    with tempfile.TemporaryDirectory() as tmp:
        for name, spec, n, mode, nproc, layout, reidx in cases.SPMM_CASES:
            m, k, rp, ci, v = cases.build_matrix(spec)
            csr = os.path.join(tmp, name + ".bin")
            gen.write_csr_bin(csr, m, k, rp, ci, v)
            prefix = os.path.join(tmp, name)
            out = run([os.path.join(REF, "minimpirun"), "-np", str(nproc), "-x", f"RP_SPMM_REIDX={reidx}", "-x", "OMP_NUM_THREADS=1",
                       os.path.join(REF, "ref_dump.exe"), csr, str(n), "1", mode, prefix, str(layout)])     # 1 timed exec: the stat table is printed
            rec = dict(m=m, k=k, n=n, nproc=nproc, layout=layout, reidx=reidx, mode=np.array(mode),
                       csr_rowptr=rp, csr_colidx=ci, csr_val=v, stdout=np.array(out))
            for r in range(nproc):
                for key, arr in gen.read_dump(f"{prefix}.r{r}.bin").items():
                    rec[f"r{r}/{key}"] = arr
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
            print("golden", name, out.splitlines()[0] if out else "")
        for name in cases.REDIST_CASES:
            lay = cases.redist_layout(name)
            gr, gc = cases.REDIST_DIMS[name]
            path = os.path.join(tmp, name + ".txt")
            with open(path, "w") as f:
                f.write(f"{len(lay)} {gr} {gc}\n")
                for row in lay:
                    f.write(" ".join(str(x) for x in row) + "\n")
            prefix = os.path.join(tmp, "rd_" + name)
            run([os.path.join(REF, "minimpirun"), "-np", str(len(lay)), os.path.join(REF, "ref_redist_dump.exe"), path, prefix])
            rec = dict(layout=np.array(lay, dtype=np.int32), dims=np.array([gr, gc]))
            for r in range(len(lay)):
                for key, arr in gen.read_dump(f"{prefix}.r{r}.bin").items():
                    rec[f"r{r}/{key}"] = arr
            np.savez_compressed(os.path.join(HERE, "redist_" + name + ".npz"), **rec)
            print("golden redist", name)


if __name__ == "__main__":
    main()
