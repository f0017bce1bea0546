set -x
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dims"
run() { n=$1; shift; env "$@" timeout 300 $B > gpurun_out/r02u_$n.json 2> gpurun_out/r02u_$n.err; echo "$n rc=$?"; }
run base X=1
run c30 GCN_CARVEOUT=30
run c30_when0 GCN_CARVEOUT=30 GCN_SEQ_WHEN=0
run c15 GCN_CARVEOUT=15
run c50 GCN_CARVEOUT=50
run c0 GCN_CARVEOUT=0
run c30_norng GCN_CARVEOUT=30 GCN_NO_RNG_OVERLAP=1
