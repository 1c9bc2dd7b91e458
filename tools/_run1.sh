LSVS_B200_LIB=variants/hangdbg/liblsvs_b200.so timeout 300 python -m pytest tests/test_attention_gpu.py -x -q -m gpu > gpurun_out/r2c_test_attn_qbuf_dbg.log 2>&1; echo "dbg rc=$?"; tail -3 gpurun_out/r2c_test_attn_qbuf_dbg.log
if grep -q "passed" gpurun_out/r2c_test_attn_qbuf_dbg.log && ! grep -q "failed" gpurun_out/r2c_test_attn_qbuf_dbg.log; then
timeout 300 python -m pytest tests/test_attention_gpu.py -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
for v in qbuf2 qbuf1; do
if [ $v = qbuf1 ]; then export LSVS_B200_LIB=variants/qbuf1/liblsvs_b200.so; else unset LSVS_B200_LIB; fi
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-incumbent --sequence-frames 0 > gpurun_out/r2c_bench_${v}_$rep.json 2> gpurun_out/r2c_bench_${v}_$rep.err
done; done
unset LSVS_B200_LIB
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_bench_qbuf*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); kc=d['kernel_classes']
    print(f, round(d['value'],1), round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in kc.items()}, d['clocks']['sm_mhz'])
PY
fi
