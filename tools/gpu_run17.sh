# round 2 session 2, run 2: new tests (dev_type, column passes, full-size configs 2-4), stencil column-pass sweep, bench lines of configs 2-4
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_dev_type.py tests/test_column_passes.py tests/test_gpu_fullsize.py -m gpu -q --tb=short --timeout 600 2>&1 | tail -n 25 ) > gpurun_out/r2s2_pytest_new.log
tail -n 12 gpurun_out/r2s2_pytest_new.log
timeout 900 python tools/kbench.py --workload stencil --variants "auto,auto:CRP_SPMM_RG_FILL=0.3,rowsplit" --passes 1,2,4,8,0 --iters 5 > gpurun_out/r2s2_kbench_stencil.log 2>&1
cut -c1-230 gpurun_out/r2s2_kbench_stencil.log
for w in stencil er; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2s2_bench_${w}_n1.json 2> gpurun_out/r2s2_bench_${w}_n1.err; echo "bench $w rc=$?"
  cut -c1-200 gpurun_out/r2s2_bench_${w}_n1.json
done
timeout 1500 python bench.py --workload rmat --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2s2_bench_rmat_n1.json 2> gpurun_out/r2s2_bench_rmat_n1.err; echo "bench rmat rc=$?"
cut -c1-200 gpurun_out/r2s2_bench_rmat_n1.json
