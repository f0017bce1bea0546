set -x
timeout 300 python tools/debug_wide.py > gpurun_out/r02e_debug_wide.log 2>&1; tail -30 gpurun_out/r02e_debug_wide.log
GCNK_NO_TCGEN05=1 timeout 300 python tools/debug_wide.py > gpurun_out/r02e_debug_wide_simt.log 2>&1; tail -30 gpurun_out/r02e_debug_wide_simt.log
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "matmul_pitched_tn" > gpurun_out/r02e_tn.log 2>&1; tail -15 gpurun_out/r02e_tn.log
