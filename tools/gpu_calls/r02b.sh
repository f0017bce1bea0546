set -x
nvidia-smi -L
python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02b_dist_tests_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_dist_tests_2gpu.log; tail -5 gpurun_out/r02b_dist_tests_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r02b_bench_n2.json; tail -5 gpurun_out/r02b_bench_n2.err
GCN_EXCHANGE=barrier python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02b_bench_n2_barrier.json 2> gpurun_out/r02b_bench_n2_barrier.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02b_bench_n2_barrier.json
