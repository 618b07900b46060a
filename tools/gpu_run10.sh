mkdir -p gpurun_out
timeout 300 python tools/kbench.py --variants "auto,auto:CRP_PANEL_SPLIT=2,panel:CRP_PANEL_SPLIT=2:CRP_PANEL_CR=24:CRP_PANEL_EMAX=96,panel:CRP_PANEL_SPLIT=2:CRP_PANEL_CR=16:CRP_PANEL_EMAX=64,panel:CRP_PANEL_SPLIT=2:CRP_PANEL_L2PF=0" --check --iters 10 > gpurun_out/r2_kbench8.log 2>&1
cut -c1-200 gpurun_out/r2_kbench8.log
