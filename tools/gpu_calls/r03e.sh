set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r03e_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r03e_gputests.log; tail -n 5 gpurun_out/r03e_gputests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims"
run() { n=$1; shift; env "$@" timeout 300 $B > gpurun_out/r03e_$n.json 2> gpurun_out/r03e_$n.err; echo "$n rc=$?"; tail -n 2 gpurun_out/r03e_$n.err; }
run default X=1
run w16 GCN_FW_WARPS=16
run w13 GCN_FW_WARPS=13
