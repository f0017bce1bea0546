set -x
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02g_dist_tests_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_dist_tests_2gpu.log; tail -8 gpurun_out/r02g_dist_tests_2gpu.log
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train.py -m gpu -x -q -k "matmul or wide or products" > gpurun_out/r02g_ops.log 2>&1; tail -4 gpurun_out/r02g_ops.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --workload products --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02g_bench_products_n2.json 2> gpurun_out/r02g_bench_products_n2.err; echo "bench rc=$?"; tail -c 2000 gpurun_out/r02g_bench_products_n2.json; tail -5 gpurun_out/r02g_bench_products_n2.err
