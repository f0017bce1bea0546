set -x
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NP $EXTRA > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 400 gpurun_out/$name.json; tail -2 gpurun_out/$name.err; PORT=$((PORT+10)); }
PORT=30300; NP=8; EXTRA="--steps 20 --warmup 5"
run r02r_bench_n8 X=1
run r02r_bench_n8_tree GCN_TREE_LOSS=1
EXTRA="--workload products --steps 10 --warmup 3"; run r02r_bench_products_n8 X=1
NP=4; EXTRA="--steps 20 --warmup 5"; run r02r_bench_n4 X=1
NP=2; run r02r_bench_n2 X=1
