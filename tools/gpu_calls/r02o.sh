set -x
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "sequential_sum" > gpurun_out/r02o_seq.log 2>&1; tail -3 gpurun_out/r02o_seq.log
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02o_dist_tests_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02o_dist_tests_2gpu.log; tail -5 gpurun_out/r02o_dist_tests_2gpu.log
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 $EXTRA > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 500 gpurun_out/$name.json; tail -2 gpurun_out/$name.err; PORT=$((PORT+10)); }
PORT=30000
EXTRA="--steps 20 --warmup 5"; run r02o_bench_n2 X=1
EXTRA="--steps 20 --warmup 5"; run r02o_bench_n2_unfused GCN_FUSED_EXCHANGE=0
