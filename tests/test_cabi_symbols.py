"""The drop-in boundary: libcrpspmm.so loads on a box without GPU and exports every symbol
include/*.h declares (no compute is called here)."""
import ctypes
import os
import re

from pycrp import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    """Function names declared in a header (a C identifier followed by '(' at declaration level)."""
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*?(?<!\\)$", "", src, flags=re.M | re.S) if False else "\n".join(
        line for line in src.splitlines() if not line.strip().startswith("#") and not line.rstrip().endswith("\\"))
    names = set()
    for mobj in re.finditer(r"^[A-Za-z_][\w \*]*?\b([A-Za-z_]\w*)\s*\(", src, flags=re.M):
        name = mobj.group(1)
        if name not in ("defined", "sizeof", "while", "if", "do", "for", "fprintf", "free_func", "attach_func"):
            names.add(name)
    return names


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    headers = [h for h in os.listdir(os.path.join(ROOT, "include")) if h.endswith(".h")]
    assert {"rowpara_spmm.h", "para2d_spmm.h", "mat_redist.h", "spmat_part.h", "dev_type.h", "utils.h", "crp_cuda.h", "crp_ext.h"} <= set(headers)
    missing = []
    total = 0
    for h in headers:
        for name in sorted(declared_functions(h)):
            total += 1
            try:
                getattr(lib, name)
            except AttributeError:
                missing.append(f"{h}:{name}")
    assert total > 70
    assert not missing, missing


def test_reference_api_names_present():
    lib = capi.load()
    for h, names in capi.EXPORTS.items():
        for name in names:
            assert hasattr(lib, name), (h, name)


def test_struct_prefix_matches_reference_layout():
    """Field order of the public structs == the reference headers (src/rowpara_spmm.h:8-40, src/para2d_spmm.h:6-14)."""
    names = [f[0] for f in capi.RowparaSpmm._fields_]
    assert names[:21] == ["nproc", "my_rank", "glb_n", "A_nrow", "rB_nrow", "rB_self_src_offset", "rB_self_dst_offset", "rB_self_nrow",
                          "rB_p2p", "rB_reidx", "A_rowptr", "A_colidx", "rB_self_src_ridxs", "rB_scnts", "rB_sridxs", "rB_sdispls",
                          "rB_rcnts", "rB_rridxs", "rB_rdispls", "A_val", "comm"]
    assert names[21:30] == ["rB_recv_size", "n_exec", "t_init", "t_pack", "t_a2a", "t_unpack", "t_spmm", "t_exec", "dev"]


def test_no_gpu_means_loud_failure_not_fallback(tmp_path):
    """Without a device and without CRP_SPMM_PLAN_ONLY the engine must abort, never compute on the CPU."""
    import subprocess
    import sys
    lib = capi.load()
    if lib.crp_cuda_device_count() > 0:
        return
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np, ctypes as C\n"
        "from pycrp import capi\n"
        "capi.mpi_init(); L = capi.load()\n"
        "rp = np.array([0, 1, 2], np.int32); ci = np.array([0, 1], np.int32); v = np.ones(2)\n"
        "d = np.array([0, 2], np.int32); e = C.POINTER(capi.RowparaSpmm)()\n"
        "L.rp_spmm_init(0, 2, capi.ptr(rp), capi.ptr(ci), capi.ptr(v), capi.ptr(d), 4, 0, C.byref(e))\n"
        "print('SURVIVED')\n" % os.path.join(ROOT, "crp-spmm_b200"))
    env = {k: v for k, v in os.environ.items() if k != "CRP_SPMM_PLAN_ONLY"}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode != 0 and "SURVIVED" not in r.stdout
    assert "no CPU fallback" in r.stderr
