# round 2 session 2, run 4: headline back to the fast configuration, pipelined e2e, composite table pin, stencil ncu
mkdir -p gpurun_out
timeout 200 python tools/kbench.py --variants "auto" --check --iters 10 2>&1 | cut -c1-200
( timeout 900 python -m pytest tests/test_gpu_e2e_pipeline.py tests/test_composite.py tests/test_dev_type.py tests/test_gpu_panel.py -m gpu -q --tb=short --timeout 240 2>&1 | tail -n 12 ) > gpurun_out/r2s2_pytest_e2e.log; tail -n 6 gpurun_out/r2s2_pytest_e2e.log
for P in 4 1 2 8; do
  CRP_SPMM_E2E_PANELS=$P timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s2_bench_e2e_p$P.json 2> gpurun_out/r2s2_bench_e2e_p$P.err; echo "bench P=$P rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2s2_bench_e2e_p$P.json").read().strip().splitlines()[-1])
print("P=$P", d["ms_per_step"], d["e2e"])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmm_panel|spmm_rowsplit" -s 2 -c 2 -o gpurun_out/r2s2_prof_stencil -f python tools/kbench.py --workload stencil --variants "auto,rowsplit" --iters 1 > gpurun_out/r2s2_ncu_stencil.log 2>&1; echo "ncu stencil rc=$?"
tail -n 3 gpurun_out/r2s2_ncu_stencil.log | cut -c1-200
