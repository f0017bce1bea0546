set -x
timeout 200 python bench.py --workload products --steps 5 --warmup 3 --no-cpu-baseline --no-dims > gpurun_out/r03i_bench_products_n1.json 2> gpurun_out/r03i_bench_products_n1.err; echo "rc=$?"; tail -c 300 gpurun_out/r03i_bench_products_n1.json
