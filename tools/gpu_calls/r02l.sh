set -x
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02l_dist_tests_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02l_dist_tests_2gpu.log; tail -8 gpurun_out/r02l_dist_tests_2gpu.log
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 600 gpurun_out/$name.json; tail -2 gpurun_out/$name.err; PORT=$((PORT+10)); }
PORT=29700
run r02l_bench_n2_fused X=1
run r02l_bench_n2_unfused GCN_FUSED_EXCHANGE=0
