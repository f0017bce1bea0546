set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02k_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02k_gputests.log; tail -8 gpurun_out/r02k_gputests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err; echo "rc=$?"; tail -c 900 gpurun_out/r02k_bench_n1.json
GCN_TC_TRANSFORM=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims > gpurun_out/r02k_bench_n1_tc.json 2> gpurun_out/r02k_bench_n1_tc.err; echo "rc=$?"; tail -c 900 gpurun_out/r02k_bench_n1_tc.json
GCN_TREE_LOSS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims > gpurun_out/r02k_bench_n1_tree.json 2> gpurun_out/r02k_bench_n1_tree.err; echo "rc=$?"; tail -c 500 gpurun_out/r02k_bench_n1_tree.json
