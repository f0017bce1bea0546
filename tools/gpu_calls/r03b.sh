set -x
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims"
run() { n=$1; shift; env "$@" timeout 300 $B > gpurun_out/r03b_$n.json 2> gpurun_out/r03b_$n.err; echo "$n rc=$?"; tail -n 2 gpurun_out/r03b_$n.err; }
run nobitloads GCN_DEBUG_NO_BITLOADS=1
run default X=1
