set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_gputests.log; tail -5 gpurun_out/r02c_gputests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims --save-trajectory gpurun_out/traj_n1.json > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/r02c_bench_n1.json
