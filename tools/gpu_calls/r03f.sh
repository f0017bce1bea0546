set -x
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NP $EXTRA > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 300 gpurun_out/$name.json; tail -n 2 gpurun_out/$name.err; PORT=$((PORT+10)); }
PORT=30700; NP=8; EXTRA="--steps 20 --warmup 5 --no-dims"
run r03f_bench_n8 X=1
run r03f_bench_n8_late GCN_SEQ_WHEN=1
EXTRA="--workload products --steps 10 --warmup 3 --no-dims"; run r03f_bench_products_n8 X=1
NP=4; EXTRA="--steps 20 --warmup 5 --no-dims"; run r03f_bench_n4 X=1
