# round 2, session 2: state-of-HEAD evidence run (1 GPU): full -m gpu suite, default bench, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi -L | head -2
( timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 300 -x 2>&1 | tail -n 15 ) > gpurun_out/r2s2_pytest_gpu.log
tail -n 5 gpurun_out/r2s2_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2s2_bench_n1.json 2> gpurun_out/r2s2_bench_n1.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r2s2_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s2_bench_ref.json 2> gpurun_out/r2s2_bench_ref.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/r2s2_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s2_ncu_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2s2_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_panel -s 3 -c 1 -o gpurun_out/r2s2_prof_panel -f python tools/kbench.py --variants auto --iters 3 > gpurun_out/r2s2_ncu_full.log 2>&1; echo "ncu full rc=$?"
