set -x
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r03h_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r03h_gputests.log; tail -n 4 gpurun_out/r03h_gputests.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r03h_bench_n1.json 2> gpurun_out/r03h_bench_n1.err; echo "rc=$?"; tail -c 300 gpurun_out/r03h_bench_n1.json
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03h_smoke.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/r03h_smoke.log
