mkdir -p gpurun_out
timeout 300 python tools/kbench.py --variants "auto,panel:CRP_PANEL_CR=24:CRP_PANEL_EMAX=96,panel:CRP_PANEL_CR=16:CRP_PANEL_EMAX=64,panel:CRP_PANEL_CR=8:CRP_PANEL_EMAX=32,panel:CRP_PANEL_STAGES=2,rowgroup" --check --iters 10 > gpurun_out/r2_kbench2.log 2>&1
cut -c1-200 gpurun_out/r2_kbench2.log
timeout 900 python -m pytest tests/test_gpu_panel.py tests/test_gpu_spmm.py -m gpu -q --maxfail=15 --tb=short --timeout 120 > gpurun_out/r2_t1.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_t1.log
tail -n 15 gpurun_out/r2_t1.log
timeout 900 python -m pytest tests/test_gpu_transports.py -m gpu -q --maxfail=10 --tb=short --timeout 180 > gpurun_out/r2_t2.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_t2.log
tail -n 25 gpurun_out/r2_t2.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_bench_n1.json
ncu --set full --clock-control none --import-source on -k regex:spmm_panel -s 2 -c 1 -f -o gpurun_out/r2_prof_panel2 python tools/kbench.py --variants auto --iters 3 > gpurun_out/r2_ncu_panel2.log 2>&1
