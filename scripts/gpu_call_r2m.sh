# 1 GPU: K4 trims (packed weights / guard, clamped partial tiles), full suite, config bench, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
python scripts/config5_scale.py --frames 64 --steps 3 > gpurun_out/r2m_cfg5.json 2> gpurun_out/r2m_cfg5.err
python scripts/config_bench.py > gpurun_out/r2m_config_bench.log 2>&1
python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
export STK_LOOP_MODE=host
CMD="python bench.py --frames 17 --steps 2 --warmup 1 --skip-cpu --skip-e2e"
ncu --set full --clock-control none --import-source on -k regex:warp_accumulate_v2 -s 4 -c 1 -o gpurun_out/prof_warp_r2m $CMD > gpurun_out/r2m_ncu_warp.log 2>&1
ncu -i gpurun_out/prof_warp_r2m.ncu-rep --page raw --csv > gpurun_out/r2m_ncu_raw_warp.csv 2>/dev/null
ncu -i gpurun_out/prof_warp_r2m.ncu-rep --page source --csv --print-source sass > gpurun_out/r2m_ncu_src_warp.csv 2>/dev/null
unset STK_LOOP_MODE
tail -4 gpurun_out/r2m_pytest.log; cat gpurun_out/r2m_cfg5.json; cat gpurun_out/r2m_config_bench.log; cut -c1-200 gpurun_out/r2m_bench.json
