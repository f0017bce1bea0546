set -x
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "matmul_pitched or wide_rowlocal or matmul_tcgen05" > gpurun_out/r02d_new_ops.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_new_ops.log; tail -25 gpurun_out/r02d_new_ops.log
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "products or wide" > gpurun_out/r02d_wide_train.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_wide_train.log; tail -25 gpurun_out/r02d_wide_train.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_gputests.log; tail -5 gpurun_out/r02d_gputests.log
