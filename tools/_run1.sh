for rep in 1 2; do
for p in 0 1; do
LSVS_ATTN_PERSIST=$p python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incumbent --sequence-frames 0 > gpurun_out/r2b_bench_persist_${p}_$rep.json 2> gpurun_out/r2b_bench_persist_${p}_$rep.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_persist_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); kc=d['kernel_classes']
    print(f, round(d['value'],1), round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in kc.items()}, round(d['attention']['tflops'],1), d['clocks']['sm_mhz'])
PY
