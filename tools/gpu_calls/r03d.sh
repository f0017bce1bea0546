set -x
timeout 1500 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/r03d_dist2.log 2>&1; echo "rc=$?" >> gpurun_out/r03d_dist2.log; tail -n 8 gpurun_out/r03d_dist2.log
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 $EXTRA > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 300 gpurun_out/$name.json; tail -n 2 gpurun_out/$name.err; PORT=$((PORT+10)); }
PORT=30500
EXTRA="--steps 20 --warmup 5 --no-dims"; run r03d_bench_n2 X=1
EXTRA="--workload products --steps 10 --warmup 3 --no-dims"; run r03d_bench_products_n2 X=1
