mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_panel.py tests/test_gpu_spmm.py tests/test_gpu_transports.py -m gpu -q --maxfail=8 --tb=short --timeout 180 -k "not golden or pwtk600 or np4" 2>&1 | tail -n 6
for sl in 0.25:0.375 0.0:1.0; do
timeout 200 python tools/kbench.py --rows $sl --variants "auto,auto:CRP_PANEL_REST=0" --check --iters 10 2>&1 | cut -c1-220
done
