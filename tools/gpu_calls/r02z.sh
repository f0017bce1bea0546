set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dims"
GCN_SEQ_WHEN=2 $CMD > gpurun_out/r02z_plain.log 2>&1 && GCN_SEQ_WHEN=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches.csv $CMD > gpurun_out/r02z_ncu.log 2>&1
echo "rc=$?"
grep -c . gpurun_out/r02z_launches.csv
