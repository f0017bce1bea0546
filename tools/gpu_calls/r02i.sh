set -x
nvidia-smi -L | head -8
timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/r02i_dist_tests_8gpu_box.log 2>&1; echo "rc=$?" >> gpurun_out/r02i_dist_tests_8gpu_box.log; tail -6 gpurun_out/r02i_dist_tests_8gpu_box.log
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $NP $EXTRA --steps 20 --warmup 5 > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 700 gpurun_out/$name.json; PORT=$((PORT+10)); }
PORT=29600; NP=8; EXTRA=""
run r02i_bench_n8 X=1
run r02i_bench_n8_overlap GCN_OVERLAP=1
run r02i_bench_n8_treeloss GCN_TREE_LOSS=1
run r02i_bench_n8_barrier GCN_EXCHANGE=barrier
EXTRA="--workload products"; run r02i_bench_products_n8 X=1
EXTRA=""; NP=4; run r02i_bench_n4 X=1
