# round 2 session 2, run 13 (1 GPU): final validation of HEAD - whole -m gpu suite, default bench line, e2e panel shape A/B, ncu launch list
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 400 -n 4 2>&1 | tail -n 15 ) > gpurun_out/r2s2_pytest_gpu_final.log
tail -n 5 gpurun_out/r2s2_pytest_gpu_final.log
timeout 600 python bench.py > gpurun_out/r2s2_bench_final_n1.json 2> gpurun_out/r2s2_bench_final_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2s2_bench_final_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["fp64"]["frac"], d["e2e"], d["cpu_baseline"]["value"], d["parity"]["rel_err_max_over_ranks"])
PY
CRP_SPMM_E2E_TAPER=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('E2E equal panels', d['ms_per_step'], d['e2e']['ms_per_step'])"
CRP_SPMM_E2E_TAPER=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('E2E tapered     ', d['ms_per_step'], d['e2e']['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2s2_ncu_bench_launches_final.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2s2_ncu_bench_final.log 2>&1; echo "ncu list rc=$?"
