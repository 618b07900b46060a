# round 2 session 2, run 8 (2 GPUs): per-block time stamps of the fused exchange kernel (CRP_PANEL_TRACE), trace overhead check
mkdir -p gpurun_out
timeout 200 python tools/kbench.py --variants "auto" --iters 10 2>&1 | cut -c1-160
run_bench() {  # name, nproc, args...
  name=$1; np=$2; shift 2
  env $ENVV timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2s2_trace_${name}.json 2> gpurun_out/r2s2_trace_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s2_trace_${name}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step")}, d["detail"]["kernel"], d["phases_ms"], [r[0] for r in d["per_rank"]["rows"]], d.get("nvlink"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2s2_trace_${name}.err").read()[-1500:])
PY
}
ENVV="CRP_PANEL_TRACE=gpurun_out/r2s2_trace_n2" run_bench n2 2
ENVV="CRP_X=1" run_bench n2_notrace 2
ls gpurun_out | grep trace_n2 | head
( timeout 600 python -m pytest tests/test_gpu_panel.py tests/test_gpu_transports.py -m gpu -q --tb=short --timeout 240 -x -k "not np8" 2>&1 | tail -n 4 ) > gpurun_out/r2s2_pytest_put.log; tail -n 3 gpurun_out/r2s2_pytest_put.log
