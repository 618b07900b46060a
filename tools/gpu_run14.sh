mkdir -p gpurun_out
for sl in 0.25:0.375 0.25:0.5 0.0:0.5; do
timeout 200 python tools/kbench.py --rows $sl --variants "auto,rowgroup" --iters 10 2>&1 | cut -c1-220
done
