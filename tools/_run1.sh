timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_gpu_suite.log 2>&1; tail -4 gpurun_out/r2c_gpu_suite.log
python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -c 600 gpurun_out/r2c_bench_n1.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2c_bench_reference_arm.json 2> gpurun_out/r2c_bench_reference_arm.err; tail -c 400 gpurun_out/r2c_bench_reference_arm.json
