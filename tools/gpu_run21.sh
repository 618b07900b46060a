# round 2 session 2, run 6 (8 GPUs): scaling line with e2e, the other BASELINE shapes at N = 8, the composite engine with device buffers
mkdir -p gpurun_out
run_bench() {  # name, nproc, args...
  name=$1; np=$2; shift 2
  env $ENVV timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2s2_bench_${name}.json 2> gpurun_out/r2s2_bench_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s2_bench_${name}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["parity"]["rel_err_max_over_ranks"], d["detail"]["grid"], d["detail"]["kernel"], d["detail"]["transport"], d["phases_ms"], d["e2e"], d.get("nvlink"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2s2_bench_${name}.err").read()[-1500:])
PY
}
ENVV="CRP_X=1" run_bench n8_e2e 8
ENVV="CRP_X=1" run_bench n8_stencil 8 --workload stencil --no-e2e
ENVV="CRP_X=1" run_bench n8_er 8 --workload er --no-e2e
# composite engine (deprecated API): caller's A in even row blocks, B / C in an even 4 x 2 block layout on the device, stencil 128^3, n = 512 fp64
CSR=$(python - <<PY
import bench
print(bench.matrix_path("stencil"))
PY
)
PYTHONPATH=crp-spmm_b200 timeout 400 crp-spmm_b200/bin/minimpirun -np 8 python -m pycrp.composite_flow $CSR 512 /tmp/cmp --device --ntest 6 --no-dump --json gpurun_out/r2s2_composite_stencil_n8.json 2>&1 | tail -n 30 | cut -c1-400
