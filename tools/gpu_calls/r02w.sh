set -x
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims"
run() { n=$1; shift; env "$@" timeout 300 $B > gpurun_out/r02w_$n.json 2> gpurun_out/r02w_$n.err; echo "$n rc=$?"; tail -2 gpurun_out/r02w_$n.err; }
run ls7 GCN_RNG_LS=7
run ls8 GCN_RNG_LS=8
run ls9 GCN_RNG_LS=9
run ls10 GCN_RNG_LS=10
run default X=1
run when2 GCN_SEQ_WHEN=2
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "dropout_stream" > gpurun_out/r02w_rngtests.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_rngtests.log; tail -3 gpurun_out/r02w_rngtests.log
GCN_RNG_LS=8 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "dropout_stream" > gpurun_out/r02w_rngtests8.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_rngtests8.log; tail -3 gpurun_out/r02w_rngtests8.log
GCN_RNG_LS=9 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "dropout_stream" > gpurun_out/r02w_rngtests9.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_rngtests9.log; tail -3 gpurun_out/r02w_rngtests9.log
