set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02s_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_gputests.log; tail -6 gpurun_out/r02s_gputests.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err; echo "rc=$?"; tail -c 500 gpurun_out/r02s_bench_n1.json
GCN_TREE_LOSS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims > gpurun_out/r02s_bench_n1_tree.json 2> gpurun_out/r02s_bench_n1_tree.err; echo "rc=$?"
GCN_NO_RNG_OVERLAP=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims > gpurun_out/r02s_bench_n1_norng.json 2> gpurun_out/r02s_bench_n1_norng.err; echo "rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims > gpurun_out/r02s_bench_n1_b.json 2> gpurun_out/r02s_bench_n1_b.err; echo "rc=$?"
