# round 2 session 2, run 12 (8 GPUs): final scaling point after removing the per-exec host collective
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_final3_n8.json 2> gpurun_out/r2s2_final3_n8.err
echo "== n8 rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s2_final3_n8.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["parity"]["rel_err_max_over_ranks"], d["phases_ms"], [r[0] for r in d["per_rank"]["rows"]], d["e2e"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2s2_final3_n8.err").read()[-1500:])
PY
