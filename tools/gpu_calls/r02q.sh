set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02q_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02q_gputests.log; tail -6 gpurun_out/r02q_gputests.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --save-trajectory gpurun_out/traj_n1.json > gpurun_out/r02q_bench_n1.json 2> gpurun_out/r02q_bench_n1.err; echo "rc=$?"; tail -c 700 gpurun_out/r02q_bench_n1.json
timeout 600 python bench.py --workload products --steps 5 --warmup 3 --no-cpu-baseline --save-trajectory gpurun_out/traj_products_n1.json > gpurun_out/r02q_bench_products_n1.json 2> gpurun_out/r02q_bench_products_n1.err; echo "rc=$?"; tail -c 700 gpurun_out/r02q_bench_products_n1.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dims"
$CMD > gpurun_out/r02q_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 12 -c 4 -o gpurun_out/r02q_prof_gather $CMD > gpurun_out/r02q_ncu_gather.log 2>&1
