set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dims"
$CMD > gpurun_out/r03g_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 12 -c 3 -o gpurun_out/r03g_prof_gather $CMD > gpurun_out/r03g_ncu_gather.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -n 2
