mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_panel.py -m gpu -x -q -s -k "test_panel_fp64 and 256" --timeout 90 > gpurun_out/r2_t0.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_t0.log
if grep -q "1 passed" gpurun_out/r2_t0.log; then
  timeout 900 python -m pytest tests/test_gpu_panel.py tests/test_gpu_spmm.py -m gpu -q --maxfail=15 --tb=short --timeout 120 > gpurun_out/r2_t1.log 2>&1
  echo "tests rc=$?" >> gpurun_out/r2_t1.log
  timeout 400 python tools/kbench.py --variants "auto,panel:CRP_PANEL_CR=24:CRP_PANEL_EMAX=96,panel:CRP_PANEL_CR=16:CRP_PANEL_EMAX=64,panel:CRP_PANEL_K=11,panel:CRP_PANEL_K=11:CRP_PANEL_CR=24:CRP_PANEL_EMAX=96,panel:CRP_PANEL_STAGES=2,panel:CRP_PANEL_GRID=296,rowgroup,rowsplit" --check --iters 10 > gpurun_out/r2_kbench1.log 2>&1
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
  echo "bench rc=$?" >> gpurun_out/r2_bench_n1.err
fi
tail -n 3 gpurun_out/r2_t0.log gpurun_out/r2_t1.log; cat gpurun_out/r2_kbench1.log | cut -c1-250
timeout 300 python tools/kbench.py --workload rmat --n 128 --variants "mergepath,rowsplit,auto" --check --iters 5 > gpurun_out/r2_kbench_rmat.log 2>&1
cat gpurun_out/r2_kbench_rmat.log | cut -c1-250
