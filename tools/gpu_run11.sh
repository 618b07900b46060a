mkdir -p gpurun_out
nvidia-smi -L | head -3
run_bench() {  # name, nproc, extra env...
  name=$1; np=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_bench_${name}.json 2> gpurun_out/r2_bench_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_${name}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["parity"]["rel_err_max_over_ranks"], d["parity"]["ok"], d["detail"]["grid"], d["detail"]["kernel"], d["phases_ms"])
    for r in d["per_rank"]["rows"]: print("   ", r)
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2_bench_${name}.err").read()[-1500:])
PY
}
run_bench n2_p2p 2 CRP_X=1
run_bench n2_nccl 2 CRP_SPMM_TRANSPORT=0
run_bench n2_p2p_ov1 2 CRP_SPMM_OVERLAP=1
timeout 600 python -m pytest tests/test_gpu_spmm.py -m gpu -q --maxfail=5 --tb=short --timeout 180 -k "golden and (np2 or np1)" 2>&1 | tail -n 4
