# round 2 session 2, run 3: clustered tiles + next-header prefetch in the panel kernel (A/B), stencil, dev_type fix
mkdir -p gpurun_out
timeout 200 python tools/kbench.py --variants "auto,auto:CRP_PANEL_CLUSTER=0,rowsplit" --check --iters 10 > gpurun_out/r2s2_kbench_cluster.log 2>&1; echo "kbench pwtk rc=$?"
cut -c1-260 gpurun_out/r2s2_kbench_cluster.log
( timeout 900 python -m pytest tests/test_gpu_panel.py tests/test_dev_type.py tests/test_column_passes.py -m gpu -q --tb=short --timeout 200 -x 2>&1 | tail -n 8 ) > gpurun_out/r2s2_pytest_panel.log; tail -n 4 gpurun_out/r2s2_pytest_panel.log
for sl in 0.25:0.375; do
timeout 200 python tools/kbench.py --rows $sl --variants "auto,auto:CRP_PANEL_CLUSTER=0" --check --iters 10 2>&1 | cut -c1-230
done
timeout 900 python tools/kbench.py --workload stencil --variants "auto,auto:CRP_SPMM_RG_FILL=0.3,auto:CRP_SPMM_RG_FILL=0.3:CRP_PANEL_CLUSTER=0,rowsplit" --check --iters 5 > gpurun_out/r2s2_kbench_stencil2.log 2>&1
cut -c1-260 gpurun_out/r2s2_kbench_stencil2.log
( timeout 900 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_transports.py -m gpu -q --tb=short --timeout 200 -x 2>&1 | tail -n 8 ) > gpurun_out/r2s2_pytest_spmm.log; tail -n 4 gpurun_out/r2s2_pytest_spmm.log
