"""Helpers shared by the tests: launching the library's driver flow on P mini-MPI ranks."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "crp-spmm_b200")
MINIMPIRUN = os.path.join(PKG, "bin", "minimpirun")


class Result:
    def __init__(self, returncode, stdout, stderr):
        self.returncode, self.stdout, self.stderr = returncode, stdout, stderr


def run_cmd(cmd, env=None, timeout=600):
    """subprocess.run that, on timeout, kills the whole process group (mini-MPI ranks included) instead of leaving orphans on the GPU."""
    import signal
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(p.pid, signal.SIGKILL)
        except ProcessLookupError:
            pass
        out, err = p.communicate()
        return Result(-9, out, err + f"\n[timeout after {timeout} s: process group killed]")
    return Result(p.returncode, out, err)


def run_flow(tmp_path, csr_path, n, mode, nproc, layout=0, reidx=1, plan_only=False, device=False, f32=False, extra_env=None, timeout=600):
    """python -m pycrp.flow on `nproc` ranks; returns the list of per-rank dump dicts."""
    prefix = os.path.join(str(tmp_path), "dump")
    env = dict(os.environ)
    env["PYTHONPATH"] = PKG + os.pathsep + env.get("PYTHONPATH", "")
    env["RP_SPMM_REIDX"] = str(reidx)
    env["OMP_NUM_THREADS"] = "2"
    if plan_only:
        env["CRP_SPMM_PLAN_ONLY"] = "1"
    env.update(extra_env or {})
    cmd = [MINIMPIRUN, "-np", str(nproc), sys.executable, "-m", "pycrp.flow", csr_path, str(n), mode, "--layout", str(layout), "--dump", prefix]
    if plan_only:
        cmd.append("--no-exec")
    if device:
        cmd.append("--device")
    if f32:
        cmd.append("--f32")
    r = run_cmd(cmd, env=env, timeout=timeout)
    assert r.returncode == 0, f"{' '.join(cmd)}\nstdout:\n{r.stdout}\nstderr:\n{r.stderr}"
    return [dict(np.load(f"{prefix}.r{i}.npz")) for i in range(nproc)]


def rel_err(C, Cref):
    """||Cref - C||_F / ||Cref||_F, the figure the reference drivers print (src/utils.c:75-89)."""
    num = np.linalg.norm((C.astype(np.float64) - Cref.astype(np.float64)).ravel())
    den = np.linalg.norm(Cref.astype(np.float64).ravel())
    return num / den if den > 0 else num
