for rep in 1 2; do
LSVS_B200_LIB=variants/base0/liblsvs_b200.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incumbent --sequence-frames 0 > gpurun_out/r2b_bench_ab_base_$rep.json 2> gpurun_out/r2b_bench_ab_base_$rep.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incumbent --sequence-frames 0 > gpurun_out/r2b_bench_ab_nt1_$rep.json 2> gpurun_out/r2b_bench_ab_nt1_$rep.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); kc=d['kernel_classes']
        print(f, round(d['value'],1), round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in kc.items()}, d['attention']['tflops'])
    except Exception as e: print(f, 'ERR', e)
PY
