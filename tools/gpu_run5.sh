mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:spmm_panel -s 2 -c 1 -f -o gpurun_out/r2_prof_panel3 python tools/kbench.py --variants auto --iters 3 > gpurun_out/r2_ncu_panel3.log 2>&1
tail -n 3 gpurun_out/r2_ncu_panel3.log
