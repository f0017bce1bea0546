set -x
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NP $EXTRA > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 400 gpurun_out/$name.json; tail -2 gpurun_out/$name.err; PORT=$((PORT+10)); }
PORT=30100; NP=8; EXTRA="--steps 20 --warmup 5"
run r02p_bench_n8 X=1
run r02p_bench_n8_unfused GCN_FUSED_EXCHANGE=0
EXTRA="--workload products --steps 10 --warmup 3"; run r02p_bench_products_n8 X=1
NP=4; EXTRA="--steps 20 --warmup 5"; run r02p_bench_n4 X=1
run r02p_bench_n4_unfused GCN_FUSED_EXCHANGE=0
