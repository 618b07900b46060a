"""Mint tests/golden/crpspmm_tables.json: the "Communicated Matrix Elements" table and the C error the UNMODIFIED deprecated
composite engine prints (reference deprecated/src/crpspmm.c:715-772 + deprecated/examples/test_crpspmm.c, built by
oracle/Makefile as oracle/_ref/test_crpspmm.exe) for the cases of tests/cases.py:CRPSPMM_CASES, in both exchange modes
(A2A_B_FINEGRAIN=0: whole blocks travel, the reference's default; =1: only the needed rows, what this library always does).

    python tests/golden/make_golden_crpspmm.py          # needs /root/reference (build container only)
"""
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "crp-spmm_b200"))
import cases  # noqa: E402
from pycrp import gen  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
ROWS = ("Redist A", "Allgatherv A", "Redist B", "Alltoallv B", "Alltoallv B necessary")


def parse_tables(out):
    """{row label: [min, max, sum]}, grid (pm, pn), C error from a test_crpspmm.exe stdout."""
    tab = {}
    body = out.split("Communicated Matrix Elements")[1]
    for label in ROWS:
        mobj = re.search(r"^" + re.escape(label) + r"\s+(\d+)\s+(\d+)\s+(\d+)\s*$", body, re.M)
        tab[label] = [int(mobj.group(i)) for i in (1, 2, 3)]
    g = re.search(r"2D partition: (\d+) \* (\d+)", out)
    e = re.search(r"\|\|C_ref - C\|\|_f / \|\|C_ref\|\|_f = ([0-9.eE+-]+)", out)
    return tab, (int(g.group(1)), int(g.group(2))), float(e.group(1))


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, spec, n, nproc in cases.CRPSPMM_CASES:
            m, k, rp, ci, v = cases.build_matrix(spec)
            mtx = os.path.join(tmp, name + ".mtx")
            gen.write_mtx(mtx, m, k, rp, ci, v)
            entry = {"n": n, "nproc": nproc, "m": m, "k": k, "nnz": int(rp[-1])}
            for fine in (0, 1):
                env = dict(os.environ, OMP_NUM_THREADS="2", A2A_B_FINEGRAIN=str(fine))
                r = subprocess.run([os.path.join(REF, "minimpirun"), "-np", str(nproc), os.path.join(REF, "test_crpspmm.exe"), mtx, str(n), "2", "1"],
                                   capture_output=True, text=True, env=env, timeout=300)
                assert r.returncode == 0, r.stdout + r.stderr
                tab, grid, err = parse_tables(r.stdout)
                assert err <= 1e-12
                entry[f"finegrain{fine}"] = tab
                entry["grid"] = list(grid)
            res[name] = entry
            print(name, entry)
    with open(os.path.join(HERE, "crpspmm_tables.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
