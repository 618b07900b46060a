mkdir -p gpurun_out
timeout 300 python tools/kbench.py --variants "auto,auto:CRP_PANEL_L2PF=0,panel:CRP_PANEL_CR=24:CRP_PANEL_EMAX=96,panel:CRP_PANEL_CR=16:CRP_PANEL_EMAX=64,panel:CRP_PANEL_STAGES=2" --check --iters 10 > gpurun_out/r2_kbench7.log 2>&1
cut -c1-200 gpurun_out/r2_kbench7.log
timeout 300 python -m pytest tests/test_gpu_panel.py -m gpu -q --maxfail=5 --tb=short --timeout 120 2>&1 | tail -n 3
CRP_PANEL_K=12 timeout 120 compute-sanitizer --tool memcheck --print-limit 8 python tools/dbg_panel.py 256 > gpurun_out/r2_sanitize_k12.log 2>&1
grep -v "^$" gpurun_out/r2_sanitize_k12.log | head -40 | cut -c1-200
