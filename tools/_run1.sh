timeout 600 python -m pytest tests/test_attention_gpu.py -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
for p in 0 default; do
if [ $p = default ]; then unset LSVS_ATTN_POLY; else export LSVS_ATTN_POLY=$p; fi
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incumbent --sequence-frames 0 > gpurun_out/r2b_bench_poly_${p}_$rep.json 2> gpurun_out/r2b_bench_poly_${p}_$rep.err
done; done
unset LSVS_ATTN_POLY
python tools/ub_attn.py 2>&1 | grep "frame S\|global S=32" | cut -c1-200
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_poly_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); kc=d['kernel_classes']
    print(f, round(d['value'],1), round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in kc.items()}, round(d['attention']['tflops'],1), d['clocks']['sm_mhz'])
PY
