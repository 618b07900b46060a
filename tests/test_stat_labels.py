"""The stat tables are observable behaviour: every line the reference printed in the golden runs
(tests/golden/*.npz keep its stdout) must be producible by the library's print_stat format strings."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
NUM = r"[-+0-9.eE]+"


def format_regexes(path, fn):
    src = open(os.path.join(ROOT, path)).read()
    body = src[src.index(fn):]
    body = body[:body.index("\n}\n")]
    out = []
    for f in re.findall(r'printf\("((?:[^"\\]|\\.)*)"', body):
        f = f.replace("\\n", "")
        parts = re.split(r"%[0-9.]*(?:zu|f|d|e)", f)
        out.append(re.compile("^" + (r"\s*" + NUM).join(re.escape(x) for x in parts) + "$"))
    return out


def check(case, path, fn, first_line):
    g = np.load(os.path.join(GOLD, case + ".npz"))
    lines = str(g["stdout"]).splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith(first_line))
    table = [ln for ln in lines[start:] if ln.strip()]
    rxs = format_regexes(path, fn)
    assert len(table) >= 8
    for ln in table:
        assert any(rx.match(ln) for rx in rxs), f"reference line not reproducible: {ln!r}"


def test_para2d_table_matches_reference_output():
    check("rand300_2d_np4_n16", "crp-spmm_b200/csrc/host/para2d_spmm.c", "void para2d_spmm_print_stat", "para2d_spmm_init() time")


def test_rp_table_matches_reference_output():
    check("rand300_rp_np4_n8", "crp-spmm_b200/csrc/host/rowpara_spmm.c", "void rp_spmm_print_stat", "rp_spmm_init() time")
