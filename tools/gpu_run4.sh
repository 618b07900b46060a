mkdir -p gpurun_out
timeout 300 python tools/kbench.py --variants "auto,panel:CRP_PANEL_CR=24:CRP_PANEL_EMAX=96,panel:CRP_PANEL_CR=16:CRP_PANEL_EMAX=64,panel:CRP_PANEL_CR=48:CRP_PANEL_EMAX=192,panel:CRP_PANEL_STAGES=2,panel:CRP_PANEL_GRID=296,rowgroup" --check --iters 10 > gpurun_out/r2_kbench3.log 2>&1
cut -c1-200 gpurun_out/r2_kbench3.log
timeout 600 python tools/kbench.py --workload stencil --variants "auto,auto:CRP_SPMM_RG_FILL=0.3,auto:CRP_SPMM_RG_FILL=0.3:CRP_SPMM_ROWGROUP_R=8,auto:CRP_SPMM_RG_FILL=0.3:CRP_SPMM_ROWGROUP_R=4" --check --iters 5 > gpurun_out/r2_kbench_stencil.log 2>&1
cut -c1-250 gpurun_out/r2_kbench_stencil.log
timeout 400 python tools/kbench.py --workload er --variants "auto,mergepath" --check --iters 5 > gpurun_out/r2_kbench_er.log 2>&1
cut -c1-250 gpurun_out/r2_kbench_er.log
