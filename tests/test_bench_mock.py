"""bench.py's GPU arm end to end on a box WITHOUT GPU, with every CUDA entry point of the C-ABI replaced by a stub
(the engine itself runs in plan-only mode): guards the JSON contract and the Python plumbing of the bench line -
a NameError there would cost the round's measurement.  Nothing is measured here."""
import ctypes as C
import importlib.util
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeLib:
    """Real library for the host planner, stubs for everything that needs a device."""

    def __init__(self, real):
        self._real = real
        self._handles = 0
        self._launches = 0
        self._host = {}

    def __getattr__(self, name):
        return getattr(self._real, name)

    def _new(self):
        self._handles += 1
        return self._handles

    def crp_cuda_device_count(self):
        return 1

    def crp_cuda_stream_create(self):
        return self._new()

    def crp_cuda_event_create(self):
        return self._new()

    def crp_cuda_event_elapsed_ms(self, a, b):
        return 0.5

    def crp_kernel_launch_count(self):
        self._launches += 2
        return self._launches

    def rp_spmm_kernel_name(self, rp):
        return b"mock_kernel"

    def crp_cuda_malloc_host(self, pp, nbytes):
        buf = (C.c_char * max(int(nbytes), 1))()
        self._host[C.addressof(buf)] = buf
        C.cast(pp, C.POINTER(C.c_void_p))[0] = C.addressof(buf)

    def crp_cuda_free_host(self, p):
        self._host.pop(p.value if hasattr(p, "value") else p, None)


    def crp_cuda_measure_dfma_tflops(self):
        return 37.0

    def rp_spmm_plan_info(self, rp, out):
        pass


for _name in ("rp_spmm_sync_stats", "crp_set_stream", "crp_set_blocking", "crp_cuda_memset_async", "crp_cuda_event_record", "crp_cuda_stream_sync", "crp_cuda_device_sync",
              "crp_cuda_event_sync", "rp_spmm_set_kernel"):
    setattr(FakeLib, _name, lambda self, *a: None)


class FakeDevBuf:
    def __init__(self, nbytes):
        self.arr = np.zeros(max(int(nbytes), 1), np.uint8)
        self.p = C.c_void_p(self.arr.ctypes.data)

    @classmethod
    def from_numpy(cls, a):
        b = cls(a.nbytes)
        b.arr[:a.nbytes] = np.frombuffer(np.ascontiguousarray(a).tobytes(), np.uint8)
        return b

    def to_numpy(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return np.frombuffer(self.arr[:n].tobytes(), dtype).reshape(shape).copy()

    def free(self):
        pass


def test_bench_line_contract(monkeypatch, tmp_path):
    monkeypatch.setenv("CRP_SPMM_PLAN_ONLY", "1")
    monkeypatch.setenv("TMPDIR", str(tmp_path))
    import tempfile
    tempfile.tempdir = None
    from pycrp import capi, flow
    fake = FakeLib(capi.load())
    monkeypatch.setattr(capi, "load", lambda: fake)
    monkeypatch.setattr(capi, "DevBuf", FakeDevBuf)
    monkeypatch.setattr(capi, "mpi_finalize", lambda: None)
    monkeypatch.setattr(flow.Problem, "exec_ptr", lambda self, B, Cc: None)
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self, t0, t1: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []})
    monkeypatch.setattr(bench, "run_reference_cpu", lambda *a, **k: {"gflops": 1.0, "sec": 1.0, "cores": 8, "ranks": 4, "threads": 2, "grid": "4x1",
                                                                   "local_spmm_s": 0.5, "steps": 3, "warmup": 1})
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "pwtk_small", "--steps", "3", "--warmup", "3"])
    out = io.StringIO()
    with redirect_stdout(out):
        assert bench.main() == 0
    tempfile.tempdir = None
    lines = [ln for ln in out.getvalue().splitlines() if ln.strip()]
    assert len(lines) == 1, lines                        # exactly ONE JSON line on stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "GFLOP/s" and d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"]) and d["roofline"]["bound"] == "hbm"
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "reference"
    assert d["gpu_launches"] > 0 and abs(d["ms_per_step"] - 0.5) < 1e-9
    # both arms describe the workload with the same config keys (the driver compares them)
    assert sorted(d["config"]) == ["driver", "n", "workload"]
    assert set(("rel_err_max_over_ranks", "rows_checked", "tol", "ok")) <= set(d["parity"])
    assert "fp64" in d["roofline"] and d["roofline"]["fp64"]["peak_tflops"] == 37.0
    assert len(d["per_rank"]["rows"]) == 1 and len(d["per_rank"]["rows"][0]) == len(d["per_rank"]["columns"])


def test_reference_arm_line(monkeypatch, tmp_path):
    spec = importlib.util.spec_from_file_location("bench_under_test2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_dump.exe")):
        pytest.skip("oracle/_ref not built")
    monkeypatch.setenv("TMPDIR", str(tmp_path))
    import tempfile
    tempfile.tempdir = None
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--workload", "pwtk_small", "--steps", "2", "--warmup", "3"])
    out = io.StringIO()
    with redirect_stdout(out):
        assert bench.main() == 0
    tempfile.tempdir = None
    lines = [ln for ln in out.getvalue().splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 3
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert sorted(d["config"]) == ["driver", "n", "workload"]
