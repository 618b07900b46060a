# round 2 session 2, run 11 (2 GPUs): no per-exec host collective any more - step time and the multi-rank pipelined host path
mkdir -p gpurun_out
timeout 200 python tools/kbench.py --rows 0.0:0.5 --variants "auto" --iters 10 2>&1 | cut -c1-160
for np in 2; do
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s2_final2_n$np.json 2> gpurun_out/r2s2_final2_n$np.err
echo "== n$np rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s2_final2_n$np.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["parity"]["rel_err_max_over_ranks"], d["phases_ms"], [r[0] for r in d["per_rank"]["rows"]], d["e2e"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2s2_final2_n$np.err").read()[-1500:])
PY
done
( timeout 600 python -m pytest tests/test_gpu_e2e_pipeline.py tests/test_gpu_transports.py -m gpu -q --tb=short --timeout 240 -x -k "not np8" 2>&1 | tail -n 4 ) > gpurun_out/r2s2_pytest_fin.log; tail -n 3 gpurun_out/r2s2_pytest_fin.log
