set -x
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "dropout_stream or rng or sequential" > gpurun_out/r02v_rngtests.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_rngtests.log; tail -5 gpurun_out/r02v_rngtests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims"
run() { n=$1; shift; env "$@" timeout 300 $B > gpurun_out/r02v_$n.json 2> gpurun_out/r02v_$n.err; echo "$n rc=$?"; tail -2 gpurun_out/r02v_$n.err; }
run bs X=1
run scalar GCN_RNG_SCALAR=1
run bs_norng GCN_NO_RNG_OVERLAP=1
run bs2 X=1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02v_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_gputests.log; tail -5 gpurun_out/r02v_gputests.log
