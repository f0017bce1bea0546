set -x
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "dense or feature or transform" > gpurun_out/r03c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r03c_tests.log; tail -n 5 gpurun_out/r03c_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims"
run() { n=$1; shift; env "$@" timeout 300 $B > gpurun_out/r03c_$n.json 2> gpurun_out/r03c_$n.err; echo "$n rc=$?"; tail -n 2 gpurun_out/r03c_$n.err; }
run default X=1
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/r03c_train.log 2>&1; echo "rc=$?" >> gpurun_out/r03c_train.log; tail -n 5 gpurun_out/r03c_train.log
