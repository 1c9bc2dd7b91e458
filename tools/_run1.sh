timeout 600 python -m pytest tests/test_attention_gpu.py -x -q -m gpu > gpurun_out/r2b_test_attn_warp.log 2>&1; tail -5 gpurun_out/r2b_test_attn_warp.log
python tools/ub_attn.py 2>&1 | grep "temporal" | sed "s/^/warp /"
LSVS_ATTN_WARP=0 python tools/ub_attn.py 2>&1 | grep "temporal" | sed "s/^/tcgen05 /"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_gpu_suite.log 2>&1; tail -5 gpurun_out/r2b_gpu_suite.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incumbent --sequence-frames 0 > gpurun_out/r2b_bench_quick.json 2> gpurun_out/r2b_bench_quick.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_quick.json').read().strip().splitlines()[-1]); kc=d['kernel_classes']
print(round(d['value'],1), round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in kc.items()}, d['attention']['tflops'])
PY
