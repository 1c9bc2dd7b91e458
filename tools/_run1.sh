timeout 900 python -m pytest tests/test_dpt_gpu.py -x -q -m gpu > gpurun_out/r2b_test_dpt.log 2>&1; tail -15 gpurun_out/r2b_test_dpt.log
