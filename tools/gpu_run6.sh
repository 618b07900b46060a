mkdir -p gpurun_out
timeout 300 python tools/kbench.py --variants "auto,panel:CRP_PANEL_CR=24:CRP_PANEL_EMAX=96,panel:CRP_PANEL_CR=16:CRP_PANEL_EMAX=64,panel:CRP_PANEL_CR=8:CRP_PANEL_EMAX=32,panel:CRP_PANEL_K=11,panel:CRP_PANEL_K=11:CRP_PANEL_CR=24:CRP_PANEL_EMAX=96,panel:CRP_PANEL_K=11:CRP_PANEL_CR=16:CRP_PANEL_EMAX=64,rowgroup" --check --iters 10 > gpurun_out/r2_kbench4.log 2>&1
cut -c1-200 gpurun_out/r2_kbench4.log
timeout 300 python -m pytest tests/test_gpu_panel.py -m gpu -q --maxfail=5 --tb=short --timeout 120 2>&1 | tail -n 5
