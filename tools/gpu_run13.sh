mkdir -p gpurun_out
run_bench() {  # name, nproc, args..., env via ENVV
  name=$1; np=$2; shift 2
  env $ENVV timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 20 --warmup 5 --no-e2e "$@" > gpurun_out/r2_bench_${name}.json 2> gpurun_out/r2_bench_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_${name}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["parity"]["rel_err_max_over_ranks"], d["parity"]["ok"], d["detail"]["grid"], d["detail"]["kernel"], d["phases_ms"], d["gpu_launches"], d["nvlink"])
    for r in d["per_rank"]["rows"]: print("   ", r)
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2_bench_${name}.err").read()[-2500:])
PY
}
ENVV="CRP_X=1" run_bench n8_p2p 8
ENVV="CRP_X=1" run_bench n4_p2p 4
ENVV="CRP_SPMM_TRANSPORT=0" run_bench n8_nccl 8
ENVV="CRP_X=1" run_bench n8_er2d 8 --workload er2d
grep -h "t_ag_A\|replicate" gpurun_out/r2_bench_n8_er2d.err | head -5
