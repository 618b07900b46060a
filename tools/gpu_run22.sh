# round 2 session 2, run 7 (2 GPUs): where does the multi-GPU step lose against the same kernel on the slice alone?
mkdir -p gpurun_out
for sl in 0.0:0.5 0.5:1.0; do
timeout 200 python tools/kbench.py --rows $sl --variants "auto" --iters 10 2>&1 | cut -c1-200
done
run_bench() {  # name, nproc, args...
  name=$1; np=$2; shift 2
  env $ENVV timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2s2_diag_${name}.json 2> gpurun_out/r2s2_diag_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s2_diag_${name}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step")}, d["detail"]["kernel"], d["detail"]["transport"], d["phases_ms"], [r[0] for r in d["per_rank"]["rows"]], [r[-1] for r in d["per_rank"]["rows"]])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2s2_diag_${name}.err").read()[-1500:])
PY
}
ENVV="CRP_X=1" run_bench n2_p2p 2
ENVV="CRP_SPMM_TRANSPORT=0" run_bench n2_nccl 2
ENVV="CRP_PANEL_REST=0" run_bench n2_p2p_norest 2
( timeout 600 python -m pytest tests/test_gpu_transports.py tests/test_gpu_plan.py tests/test_dev_type.py -m gpu -q --tb=short --timeout 240 -x 2>&1 | tail -n 5 ) > gpurun_out/r2s2_pytest_adv.log; tail -n 3 gpurun_out/r2s2_pytest_adv.log
