mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
for L in 3 4 6 8; do
  python bench.py --frames 32 --steps 3 --warmup 3 --skip-cpu --skip-e2e --lanes $L > gpurun_out/r2e_bench_lanes$L.json 2> gpurun_out/r2e_bench_lanes$L.err
done
tail -3 gpurun_out/r2e_pytest.log
python - <<'PY'
import json
for L in (3,4,6,8):
    try:
        d=json.loads(open(f'gpurun_out/r2e_bench_lanes{L}.json').read().strip().splitlines()[-1])
        print(L, round(d['value']), d['ms_per_step'], d['roofline']['in_step']['us_per_launch'], d['roofline']['us_per_launch'])
    except Exception as e: print(L, 'failed', e)
PY
