set -x
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "sequential" > gpurun_out/r02y_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02y_tests.log; tail -n 5 gpurun_out/r02y_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims"
run() { n=$1; shift; env "$@" timeout 300 $B > gpurun_out/r02y_$n.json 2> gpurun_out/r02y_$n.err; echo "$n rc=$?"; tail -n 2 gpurun_out/r02y_$n.err; }
run default X=1
run when2 GCN_SEQ_WHEN=2
run when0 GCN_SEQ_WHEN=0
