timeout 900 python -m pytest tests/test_model_gpu.py tests/test_precision_gpu.py tests/test_gemm_gpu.py -x -q -m gpu > gpurun_out/r2c_test_cam.log 2>&1; tail -5 gpurun_out/r2c_test_cam.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incumbent --sequence-frames 0 > gpurun_out/r2c_bench_cam.json 2> gpurun_out/r2c_bench_cam.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_cam.json').read().strip().splitlines()[-1]); kc=d['kernel_classes']
print(round(d['value'],1), round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2), v['launches_per_step']) for k,v in kc.items()}, round(d['attention']['tflops'],1), d['clocks']['sm_mhz'])
PY
