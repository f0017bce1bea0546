timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/r03j_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r03j_gputests.log; tail -n 3 gpurun_out/r03j_gputests.log
