mkdir -p gpurun_out
timeout 300 python tools/kbench.py --variants "auto,panel:CRP_PANEL_K=12,panel:CRP_PANEL_K=12:CRP_PANEL_CR=24:CRP_PANEL_EMAX=96,panel:CRP_PANEL_K=12:CRP_PANEL_CR=16:CRP_PANEL_EMAX=64,panel:CRP_PANEL_CR=24:CRP_PANEL_EMAX=96,panel:CRP_PANEL_CR=16:CRP_PANEL_EMAX=64" --check --iters 10 > gpurun_out/r2_kbench5.log 2>&1
cut -c1-200 gpurun_out/r2_kbench5.log
