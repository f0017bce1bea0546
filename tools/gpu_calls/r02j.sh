set -x
timeout 300 python tools/time_wide_gemms.py > gpurun_out/r02j_wide_gemms.txt 2>&1; cat gpurun_out/r02j_wide_gemms.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dims"
$CMD > gpurun_out/r02j_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02j_launches.csv $CMD > gpurun_out/r02j_ncu_launches.log 2>&1
$CMD > gpurun_out/r02j_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 12 -c 4 -o gpurun_out/r02j_prof_gather $CMD > gpurun_out/r02j_ncu_gather.log 2>&1
CMD2="python tools/time_wide_gemms.py 612257 1"
$CMD2 > gpurun_out/r02j_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"matmul_tc_kernel|matmul_tn_kernel" -s 15 -c 5 -o gpurun_out/r02j_prof_matmul $CMD2 > gpurun_out/r02j_ncu_matmul.log 2>&1
ls -la gpurun_out/ | tail -8
