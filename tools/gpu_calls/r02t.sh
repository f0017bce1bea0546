set -x
for w in 1 0 2; do
GCN_SEQ_WHEN=$w timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims > gpurun_out/r02t_bench_n1_when$w.json 2> gpurun_out/r02t_bench_n1_when$w.err; echo "rc=$?"
done
GCN_SEQ_WHEN=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims > gpurun_out/r02t_bench_n1_when1b.json 2> gpurun_out/r02t_bench_n1_when1b.err; echo "rc=$?"
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/r02t_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02t_gputests.log; tail -4 gpurun_out/r02t_gputests.log
