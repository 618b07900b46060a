# round 2 session 2, run 9 (8 GPUs): final scaling line of the pwtk-shaped workload (fused exchange with the put off the critical path)
mkdir -p gpurun_out
for np in 8 4; do
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_final_n$np.json 2> gpurun_out/r2s2_final_n$np.err
echo "== n$np rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s2_final_n$np.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["parity"]["rel_err_max_over_ranks"], d["detail"]["grid"], d["detail"]["kernel"], d["phases_ms"], [r[0] for r in d["per_rank"]["rows"]], d["e2e"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2s2_final_n$np.err").read()[-1500:])
PY
done
