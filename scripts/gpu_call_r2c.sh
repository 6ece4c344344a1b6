# round 2, third session, first call: state of HEAD on a fresh box (tests, bench, block timeline, ncu of the default K2 geometry)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_gpus.log
python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python scripts/timing_probe.py 2:16 2:14 > gpurun_out/r2c_timing_probe.log 2>&1
python scripts/k2_variants.py 13 2:16 2:16:2 > gpurun_out/r2c_k2_variants.log 2>&1
python scripts/sharp_probe.py > gpurun_out/r2c_sharp_probe.log 2>&1
export STK_LOOP_MODE=host
CMD="python bench.py --frames 8 --steps 2 --warmup 1 --skip-cpu --skip-e2e"
$CMD > gpurun_out/r2c_plain_host.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2c_launches_host.csv $CMD > gpurun_out/r2c_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ecc_iter_v2 -s 12 -c 2 -o gpurun_out/prof_ecc_r2c $CMD > gpurun_out/r2c_ncu_ecc.log 2>&1
ncu -i gpurun_out/prof_ecc_r2c.ncu-rep --page raw --csv > gpurun_out/r2c_ncu_raw_ecc.csv 2>/dev/null
ncu -i gpurun_out/prof_ecc_r2c.ncu-rep --page source --csv --print-source sass > gpurun_out/r2c_ncu_src_ecc.csv 2>/dev/null
unset STK_LOOP_MODE
tail -3 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_timing_probe.log gpurun_out/r2c_k2_variants.log; tail -2 gpurun_out/r2c_bench.err; cat gpurun_out/r2c_bench.json
