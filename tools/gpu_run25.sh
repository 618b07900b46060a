# round 2 session 2, run 10 (8 GPUs): per-block trace of the fused kernel at N = 8 + the same kernel on one rank's slice alone, same box
mkdir -p gpurun_out
timeout 120 python tools/kbench.py --rows 0.25:0.375 --variants "auto" --iters 10 2>&1 | cut -c1-160
CRP_PANEL_TRACE=gpurun_out/r2s2_trace_slice timeout 120 python tools/kbench.py --rows 0.25:0.375 --variants "auto" --iters 3 2>&1 | cut -c1-160
CRP_PANEL_TRACE=gpurun_out/r2s2_trace_n8 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2s2_trace_n8.json 2> gpurun_out/r2s2_trace_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2s2_trace_n8.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, [r[0] for r in d["per_rank"]["rows"]])
PY
ls gpurun_out | grep -c "trace_n8\.[0-9]"
