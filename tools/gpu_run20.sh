# round 2 session 2, run 5 (2 GPUs): multi-rank pipelined e2e over real NVLink peer stores, GPU-built plans, panel-count sweep
mkdir -p gpurun_out
nvidia-smi -L
( timeout 900 python -m pytest tests/test_gpu_plan.py tests/test_gpu_e2e_pipeline.py -m gpu -q --tb=short --timeout 240 2>&1 | tail -n 12 ) > gpurun_out/r2s2_pytest_plan.log; tail -n 6 gpurun_out/r2s2_pytest_plan.log
( timeout 900 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_transports.py -m gpu -q --tb=short --timeout 240 -k "np2 or np1 or np3" 2>&1 | tail -n 8 ) > gpurun_out/r2s2_pytest_np2.log; tail -n 4 gpurun_out/r2s2_pytest_np2.log
run_bench() {  # name, nproc, args...
  name=$1; np=$2; shift 2
  env $ENVV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2s2_bench_${name}.json 2> gpurun_out/r2s2_bench_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s2_bench_${name}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["parity"]["rel_err_max_over_ranks"], d["detail"]["grid"], d["detail"]["kernel"], d["detail"]["transport"], d["e2e"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2s2_bench_${name}.err").read()[-2500:])
PY
}
ENVV="CRP_X=1" run_bench n2_e2e 2
ENVV="CRP_SPMM_E2E_PANELS=1" run_bench n2_e2e_p1 2
for P in 3 5 6; do
  CRP_SPMM_E2E_PANELS=$P timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s2_bench_e2e_p$P.json 2> gpurun_out/r2s2_bench_e2e_p$P.err; echo "bench P=$P rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2s2_bench_e2e_p$P.json").read().strip().splitlines()[-1])
print("P=$P", d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["matches_device_result"])
PY
done
