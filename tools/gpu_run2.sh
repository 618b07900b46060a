mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:spmm_panel -s 2 -c 2 -f -o gpurun_out/r2_prof_panel python tools/kbench.py --variants auto --iters 3 > gpurun_out/r2_ncu_panel.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2_ncu_panel.log
tail -n 5 gpurun_out/r2_ncu_panel.log
