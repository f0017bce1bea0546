set -x
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "matmul_pitched_tn" > gpurun_out/r02f_tn.log 2>&1; tail -6 gpurun_out/r02f_tn.log
timeout 300 python tools/debug_wide.py > gpurun_out/r02f_debug_wide.log 2>&1; tail -22 gpurun_out/r02f_debug_wide.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02f_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_gputests.log; tail -12 gpurun_out/r02f_gputests.log
timeout 900 python bench.py --workload products --steps 5 --warmup 3 --no-cpu-baseline --save-trajectory gpurun_out/traj_products_n1.json > gpurun_out/r02f_bench_products_n1.json 2> gpurun_out/r02f_bench_products_n1.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r02f_bench_products_n1.json; tail -5 gpurun_out/r02f_bench_products_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02f_bench_n1.json
